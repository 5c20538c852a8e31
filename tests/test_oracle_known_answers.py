"""CPU tests of the oracle against everything the reference's surviving code and tests pin
(SURVEY.md §8(c)), plus analytic known answers for the formulas docs/ORACLE_SPEC.md fixes."""
import ctypes as C
import math

import numpy as np
import pytest
from scipy import signal

from groove_b200 import abi
from tests import scenes
from tests.oracle_binding import OracleEngine, oracle_lib, pcm16


def render(scene, sr=44100.0):
    o = OracleEngine(sr)
    n = scene(o)
    return o.render(n)


# ---- pinned by the reference ------------------------------------------------------------------
def test_tuning_ratio_known_answers():
    """settings/src/patches.rs:754-796 oscillator_tuning_helpers."""
    lib = oracle_lib()
    assert lib.go_tune_ratio(5, 0.0) == pytest.approx(1.3348398541700344, abs=1e-15)
    assert lib.go_tune_ratio(12, 0.0) == 2.0
    assert lib.go_tune_ratio(-12, 0.0) == 0.5
    assert lib.go_tune_ratio(0, 1200.0) == 2.0
    assert lib.go_tune_ratio(0, 0.0) == 1.0


def test_gather_audio_branch_sum():
    """orchestrator.rs:1640-1668 gather_audio_with_branches: 0.1 + 0.5*(0.3+0.5)."""
    out = render(scenes.scene_graph_toys)
    assert np.allclose(out, 0.1 + 0.5 * (0.3 + 0.5), atol=1e-15)


def test_gather_audio_basic_cases():
    """orchestrator.rs:1444-1541: nothing patched -> silence; gain multiplies; chains compose; an
    effect with no input -> silence; patch order does not matter."""
    o = OracleEngine()
    o.finalize()
    assert np.all(o.render(8) == 0.0)

    o = OracleEngine()
    a = o.add_instrument(abi.INST_TOY_SOURCE, abi.ToySourceParams(0.1, 0.1))
    g1 = o.add_effect(abi.FX_GAIN, abi.GainParams(0.5))
    g2 = o.add_effect(abi.FX_GAIN, abi.GainParams(0.25))
    lonely = o.add_effect(abi.FX_GAIN, abi.GainParams(0.9))
    o.patch_chain([a, g1, g2, abi.MAIN_MIXER])
    o.patch(lonely, abi.MAIN_MIXER)
    o.finalize()
    assert np.allclose(o.render(4), 0.1 * 0.5 * 0.25, atol=1e-17)

    def two(order):
        o = OracleEngine()
        x = o.add_instrument(abi.INST_TOY_SOURCE, abi.ToySourceParams(0.1, 0.2))
        y = o.add_instrument(abi.INST_TOY_SOURCE, abi.ToySourceParams(0.3, 0.4))
        for u in (x, y)[::order]:
            o.patch(u, abi.MAIN_MIXER)
        o.finalize()
        return o.render(3)
    assert np.allclose(two(1), two(-1), atol=1e-16)
    assert np.allclose(two(1), [[0.4, 0.6]] * 3, atol=1e-16)


def test_patch_rules():
    """orchestrator.rs:263-304: the input of a patch must be an effect; unknown uids are errors."""
    o = OracleEngine()
    a = o.add_instrument(abi.INST_TOY_SOURCE, abi.ToySourceParams(0.1, 0.1))
    b = o.add_instrument(abi.INST_TOY_SOURCE, abi.ToySourceParams(0.1, 0.1))
    with pytest.raises(abi.GrooveError) as ei:
        o.patch(a, b)
    assert ei.value.code == abi.EGRAPH
    with pytest.raises(abi.GrooveError) as ei:
        o.patch(a, 999)
    assert ei.value.code == abi.ENOENT
    g1 = o.add_effect(abi.FX_GAIN, abi.GainParams(0.5))
    g2 = o.add_effect(abi.FX_GAIN, abi.GainParams(0.5))
    o.patch(g1, g2)
    o.patch(g2, g1)
    o.patch(g1, abi.MAIN_MIXER)
    with pytest.raises(abi.GrooveError) as ei:
        o.finalize()
    assert ei.value.code == abi.EGRAPH


def test_render_is_chunk_size_independent():
    """orchestrator.rs:1683: a prime buffer size (17) must give the same audio as any other."""
    def run(chunk):
        o = OracleEngine()
        n = scenes.scene_cello_chord(o)
        parts = []
        done = 0
        while done < n:
            k = min(chunk, n - done)
            parts.append(o.render(k).copy())
            done += k
        return np.concatenate(parts)
    a, b = run(17), run(4096)
    assert np.array_equal(a, b)


def test_pcm16_truncates_and_saturates():
    """orchestration/src/helpers.rs:78,90-91: (sample * 32767.0) as i16."""
    x = np.array([0.0, 0.5, -0.5, 1.0, -1.0, 1.5, -1.5, 0.99999, -0.99999, 1e-9, float("nan")])
    want = np.array([0, 16383, -16383, 32767, -32767, 32767, -32768, 32766, -32766, 0, 0], dtype=np.int16)
    assert np.array_equal(pcm16(x), want)


# ---- analytic known answers for the spec'd DSP ---------------------------------------------------
def test_note_frequencies():
    lib = oracle_lib()
    assert lib.go_note_hz(69) == 440.0
    assert lib.go_note_hz(81) == 880.0
    assert lib.go_note_hz(60) == pytest.approx(261.6255653005986, rel=1e-15)


def test_percent_to_frequency_range():
    lib = oracle_lib()
    assert lib.go_pct_to_hz(0.0) == 25.0
    assert lib.go_pct_to_hz(1.0) == pytest.approx(20000.0, rel=1e-13)
    assert scenes.hz_to_pct(lib.go_pct_to_hz(0.37)) == pytest.approx(0.37, abs=1e-14)


def test_envelope_segments_and_continuity():
    lib = oracle_lib()
    p = abi.env(0.01, 0.02, 0.6, 0.03)
    sr = 10000.0
    na, nd, nr = 100, 200, 300
    held = 1 << 60
    lv = lambda n, off=held: lib.go_envelope_level(C.byref(p), sr, 1000, off, n)
    assert lv(999) == 0.0 and lv(1000) == 0.0
    assert lv(1000 + na) == 1.0                              # attack ends exactly at 1
    assert lv(1000 + na // 2) == pytest.approx(0.75)          # t(2-t) at t = 0.5
    assert lv(1000 + na + nd) == 0.6 and lv(10**6) == 0.6     # sustain
    assert lv(1000 + na + nd // 2) == pytest.approx(0.6 + 0.4 * 0.25)
    off = 1000 + na + nd + 50
    assert lv(off, off) == pytest.approx(0.6)                 # release starts from the current level
    assert lv(off + nr // 2, off) == pytest.approx(0.6 * 0.25)
    assert lv(off + nr, off) == 0.0
    seq = [lv(n, off) for n in range(990, off + nr + 5)]
    assert max(abs(np.diff(seq))) < 0.03                      # no jumps anywhere


def test_lp24_is_unity_at_dc_and_matches_scipy():
    lib = oracle_lib()
    c = np.zeros(10)
    lib.go_lp24_coefficients(1000.0, 0.8, 44100.0, c.ctypes.data)
    for s in range(2):
        b0, b1, b2, a1, a2 = c[5 * s:5 * s + 5]
        assert (b0 + b1 + b2) / (1.0 - a1 - a2) == pytest.approx(1.0, rel=1e-12)
    # oracle render of an impulse-ish input through the effect vs scipy sosfilt with the same coefficients
    o = OracleEngine()
    a = o.add_instrument(abi.INST_TOY_SOURCE, abi.ToySourceParams(1.0, -0.5))
    f = o.add_effect(abi.FX_LOW_PASS_24DB, abi.Lowpass24Params(1000.0, 0.8))
    o.patch_chain([a, f, abi.MAIN_MIXER])
    o.finalize()
    y = o.render(600)
    sos = np.array([[c[0], c[1], c[2], 1.0, -c[3], -c[4]], [c[5], c[6], c[7], 1.0, -c[8], -c[9]]])
    ref = signal.sosfilt(sos, np.ones(600))
    assert np.allclose(y[:, 0], ref, atol=1e-12)
    assert np.allclose(y[:, 1], -0.5 * ref, atol=1e-12)
    assert y[-1, 0] == pytest.approx(1.0, abs=1e-6)           # step response settles at the DC gain


@pytest.mark.parametrize("kind,p2", [(abi.FX_LOW_PASS_12DB, 0.707), (abi.FX_HIGH_PASS_12DB, 2.0),
                                     (abi.FX_BAND_PASS_12DB, 1.5), (abi.FX_BAND_STOP_12DB, 0.7),
                                     (abi.FX_ALL_PASS_12DB, 5.0), (abi.FX_PEAKING_EQ_12DB, 6.0),
                                     (abi.FX_LOW_SHELF_12DB, -4.0), (abi.FX_HIGH_SHELF_12DB, 5.0)])
def test_rbj_biquads_match_cookbook_and_lfilter(kind, p2):
    """doc/Audio-EQ-Cookbook.txt:74-198: coefficients recomputed here in numpy, response via lfilter."""
    lib = oracle_lib()
    sr, fc = 44100.0, 1200.0
    c = np.zeros(5)
    lib.go_rbj_coefficients(kind, fc, p2, sr, c.ctypes.data)
    w0 = 2 * math.pi * fc / sr
    cs, sn = math.cos(w0), math.sin(w0)
    if kind in (abi.FX_LOW_PASS_12DB, abi.FX_HIGH_PASS_12DB, abi.FX_ALL_PASS_12DB):
        al = sn / (2 * p2)
        a = [1 + al, -2 * cs, 1 - al]
        b = {abi.FX_LOW_PASS_12DB: [(1 - cs) / 2, 1 - cs, (1 - cs) / 2],
             abi.FX_HIGH_PASS_12DB: [(1 + cs) / 2, -(1 + cs), (1 + cs) / 2],
             abi.FX_ALL_PASS_12DB: [1 - al, -2 * cs, 1 + al]}[kind]
    elif kind in (abi.FX_BAND_PASS_12DB, abi.FX_BAND_STOP_12DB):
        al = sn * math.sinh(math.log(2) / 2 * p2 * w0 / sn)
        a = [1 + al, -2 * cs, 1 - al]
        b = [al, 0, -al] if kind == abi.FX_BAND_PASS_12DB else [1, -2 * cs, 1]
    elif kind == abi.FX_PEAKING_EQ_12DB:
        A = 10 ** (p2 / 40)
        al = sn / (2 * math.sqrt(0.5))
        b = [1 + al * A, -2 * cs, 1 - al * A]
        a = [1 + al / A, -2 * cs, 1 - al / A]
    else:
        A = 10 ** (p2 / 40)
        al = sn / 2 * math.sqrt(2.0)
        t = 2 * math.sqrt(A) * al
        if kind == abi.FX_LOW_SHELF_12DB:
            b = [A * ((A + 1) - (A - 1) * cs + t), 2 * A * ((A - 1) - (A + 1) * cs), A * ((A + 1) - (A - 1) * cs - t)]
            a = [(A + 1) + (A - 1) * cs + t, -2 * ((A - 1) + (A + 1) * cs), (A + 1) + (A - 1) * cs - t]
        else:
            b = [A * ((A + 1) + (A - 1) * cs + t), -2 * A * ((A - 1) + (A + 1) * cs), A * ((A + 1) + (A - 1) * cs - t)]
            a = [(A + 1) - (A - 1) * cs + t, 2 * ((A - 1) - (A + 1) * cs), (A + 1) - (A - 1) * cs - t]
    want = np.array([b[0], b[1], b[2], a[1], a[2]]) / a[0]
    assert np.allclose(c, want, rtol=1e-12, atol=1e-15)
    # the oracle's DF1 loop vs scipy.signal.lfilter on a noise burst
    o = OracleEngine(sr)
    src = o.add_instrument(abi.INST_SAMPLER, abi.SamplerParams(440.0, 1, 0))
    x = np.random.default_rng(1).uniform(-0.9, 0.9, 2000)
    o.load_sample(src, 0, x, sr, 440.0)
    f = o.add_effect(kind, abi.BiquadParams(fc, p2))
    o.patch_chain([src, f, abi.MAIN_MIXER])
    o.finalize()
    o.note_on(0, src, 69)
    y = o.render(2000)[:, 0]
    ref = signal.lfilter(want[:3], [1.0, want[3], want[4]], x)
    assert np.allclose(y, ref, atol=1e-11)


def test_sine_carrier_frequency_and_level():
    """An FM voice with beta = 0 is a pure sine at the note frequency: checks the phase accumulator,
    MIDI key -> Hz and the centre pan gain (0.75) against the analytic signal."""
    sr = 48000.0
    o = OracleEngine(sr)
    u = o.add_instrument(abi.INST_FM, scenes.fm_params(beta=0.0, depth=0.0, car=(0, 0, 1, 0), mod=(0, 0, 1, 0), voices=1))
    o.patch(u, abi.MAIN_MIXER)
    o.finalize()
    o.note_on(0, u, 69)
    y = o.render(4800)[:, 0]
    n = np.arange(4800)
    assert np.allclose(y, 0.75 * np.sin(2 * np.pi * 440.0 * n / sr), atol=1e-9)


@pytest.mark.parametrize("wave,shape", [
    (abi.WAVE_SQUARE, lambda p: np.where(p < 0.5, 1.0, -1.0)),
    (abi.WAVE_SAWTOOTH, lambda p: np.where(p < 0.5, 2 * p, 2 * p - 2)),
    (abi.WAVE_TRIANGLE, lambda p: np.where(p < 0.5, 4 * p - 1, 3 - 4 * p)),
    (abi.WAVE_PULSE_WIDTH, lambda p: np.where(p < 0.25, 1.0, -1.0)),
    (abi.WAVE_DEBUG_MAX, lambda p: np.ones_like(p)),
])
def test_naive_waveforms_through_an_open_filter(wave, shape):
    """Naive waveforms (README.md:113-116).  The voice filter cannot be bypassed, so the expected
    signal is the analytic waveform run through scipy's sosfilt with the oracle's own coefficients."""
    sr = 48000.0
    lib = oracle_lib()
    o = OracleEngine(sr)
    p = scenes.generic_welsh(w1=wave, pw1=0.25, w2=abi.WAVE_NONE, mix=1.0, cutoff_end=0.0, routing=abi.LFO_NONE,
                             cutoff_hz=5000.0, amp=(0, 0, 1, 0), voices=1)
    u = o.add_instrument(abi.INST_WELSH, p)
    o.patch(u, abi.MAIN_MIXER)
    o.finalize()
    o.note_on(0, u, 45)  # 110 Hz
    y = o.render(3000)[:, 0]
    n = np.arange(3000)
    phase = (n * (110.0 / sr)) % 1.0
    c = np.zeros(10)
    lib.go_lp24_coefficients(5000.0, 0.707, sr, c.ctypes.data)
    sos = np.array([[c[0], c[1], c[2], 1.0, -c[3], -c[4]], [c[5], c[6], c[7], 1.0, -c[8], -c[9]]])
    ref = signal.sosfilt(sos, shape(phase)) * 0.5 * 0.75      # amp-LFO idle factor 0.5, centre pan 0.75
    # phase landmarks can differ by one frame from the float phase above; allow isolated edge frames
    bad = np.abs(y - ref) > 1e-6
    assert bad.mean() < 0.02


def test_dca_pan_law():
    for pan, (l, r) in ((0.0, (0.75, 0.75)), (-1.0, (1.0, 0.0)), (1.0, (0.0, 1.0))):
        o = OracleEngine(48000.0)
        u = o.add_instrument(abi.INST_FM, scenes.fm_params(beta=0.0, car=(0, 0, 1, 0), pan=pan, voices=1))
        o.patch(u, abi.MAIN_MIXER)
        o.finalize()
        o.note_on(0, u, 69)
        y = o.render(200)
        k = np.argmax(np.abs(np.sin(2 * np.pi * 440.0 * np.arange(200) / 48000.0)))
        s = math.sin(2 * math.pi * 440.0 * k / 48000.0)
        assert y[k, 0] == pytest.approx(l * s, abs=1e-9) and y[k, 1] == pytest.approx(r * s, abs=1e-9)


def test_delay_reverb_chorus_impulse_responses():
    sr = 10000.0
    def impulse_through(kind, params, n):
        o = OracleEngine(sr)
        s = o.add_instrument(abi.INST_SAMPLER, abi.SamplerParams(440.0, 1, 0))
        o.load_sample(s, 0, np.array([1.0]), sr, 440.0)
        f = o.add_effect(kind, params)
        o.patch_chain([s, f, abi.MAIN_MIXER])
        o.finalize()
        o.note_on(0, s, 69)
        return o.render(n)[:, 0]
    y = impulse_through(abi.FX_DELAY, abi.DelayParams(0.01), 300)
    assert y[100] == 1.0 and np.count_nonzero(y) == 1
    y = impulse_through(abi.FX_CHORUS, abi.ChorusParams(4, 0.04, 1.0), 600)
    assert np.allclose(y[[100, 200, 300, 400]], 0.25) and np.count_nonzero(y) == 4
    # Schroeder comb: impulses every D frames decaying by g; the all-passes only rearrange them, so the
    # energy decays by 60 dB over `seconds`
    y = impulse_through(abi.FX_REVERB, abi.ReverbParams(1.0, 0.5), 12000)
    assert np.all(y[:297] == 0.0) and y[297 + 50 + 17] != 0.0  # first comb echo, through both all-passes
    e_early = np.sum(y[:2500] ** 2)
    e_late = np.sum(y[-2500:] ** 2)
    assert e_late < e_early * 1e-3


def test_voice_stealing_and_retrigger_are_deterministic():
    a = render(scenes.scene_welsh_variants)
    b = render(scenes.scene_welsh_variants)
    assert np.array_equal(a, b) and np.abs(a).max() > 0.1 and not np.isnan(a).any()


def test_fanout_evaluates_each_node_once():
    """Deliberate, documented deviation: a node on several patch paths is evaluated once per frame."""
    def build(paths):
        o = OracleEngine()
        u = o.add_instrument(abi.INST_FM, scenes.fm_params(voices=1))
        for _ in range(paths):
            g = o.add_effect(abi.FX_GAIN, abi.GainParams(0.5))
            o.patch_chain([u, g, abi.MAIN_MIXER])
        o.finalize()
        o.note_on(0, u, 60)
        return o.render(500)
    assert np.allclose(build(2), 2 * build(1), atol=1e-15)


def test_mma_transforms_match_reference_bounds():
    """orchestration/src/util.rs:286-318 (mma_concave_transform / mma_convex_transform), verbatim bounds."""
    lib = oracle_lib()
    cc, cv = lib.go_mma_concave, lib.go_mma_convex
    assert cc(0.001) < 0.0002 and cc(0.01) < 0.019 and cc(0.1) < 0.02
    assert 0.12 < cc(0.5) < 0.13 and cc(0.9) > 0.40 and cc(0.99) > 0.83 and cc(0.995) > 0.95
    assert cv(0.995) > 0.999 and cv(0.99) > 0.998 and cv(0.9) > 0.98
    assert 0.87 < cv(0.5) < 0.88 and cv(0.1) < 0.59 and cv(0.01) < 0.17 and cv(0.001) < 0.0005
    for i in range(101):
        x = i / 100.0
        assert cc(x) <= x + 1e-15 and cv(x) >= x - 1e-15


def test_transport_advances_exactly_one_beat_per_second_at_60_bpm():
    """src/mini/transport.rs:157-188: frame-by-frame MusicalTime deltas over one second sum to
    UNITS_IN_BEAT (65536) at every sample rate of the reference test."""
    from groove_b200.project import UNITS_IN_BEAT, frames_to_units
    for sr in (100, 997, 22050, 44100, 48000, 88200, 98689, 100000, 262144):
        covered, prev = 0, 0
        for f in range(1, sr + 1):
            now = frames_to_units(f, 60.0, sr)
            covered += now - prev
            prev = now
        assert covered == UNITS_IN_BEAT and frames_to_units(sr, 60.0, sr) == UNITS_IN_BEAT


def test_sidechain_link_known_answer():
    """Signal-passthrough control link (include/groove_b200.h, gb_link_control): a constant 0.3 source through a
    passthrough node drives the ceiling of a gain on a constant 0.5 source.  Controllers work once per
    64-frame buffer (orchestrator.rs:631-708), so the first buffer still has the initial ceiling 1.0:
    0.3 + 0.5 * 1.0, and every later frame 0.3 + 0.5 * 0.3.  An unpatched source never fires its link."""
    o = OracleEngine()
    a = o.add_instrument(abi.INST_TOY_SOURCE, abi.ToySourceParams(0.3, 0.3))
    b = o.add_instrument(abi.INST_TOY_SOURCE, abi.ToySourceParams(0.5, 0.5))
    c = o.add_instrument(abi.INST_TOY_SOURCE, abi.ToySourceParams(0.25, 0.25))
    tap = o.add_effect(abi.FX_SIGNAL_PASSTHROUGH)
    dead = o.add_effect(abi.FX_SIGNAL_PASSTHROUGH)
    g = o.add_effect(abi.FX_GAIN, abi.GainParams(1.0))
    g2 = o.add_effect(abi.FX_GAIN, abi.GainParams(1.0))
    o.patch_chain([a, tap, abi.MAIN_MIXER])
    o.patch_chain([b, g, abi.MAIN_MIXER])
    o.patch_chain([c, g2, abi.MAIN_MIXER])
    o.patch(a, dead)                       # dead has an input but no path to the main mixer
    o.link_control(tap, g, 0)
    o.link_control(dead, g2, 0)
    o.finalize()
    y = o.render(200)
    assert np.allclose(y[:64], 0.3 + 0.5 * 1.0 + 0.25, atol=1e-15)
    assert np.allclose(y[64:], 0.3 + 0.5 * 0.3 + 0.25, atol=1e-15)
    # errors: wrong source kind, unsupported target, second link on one target
    o = OracleEngine()
    tap = o.add_effect(abi.FX_SIGNAL_PASSTHROUGH)
    g = o.add_effect(abi.FX_GAIN, abi.GainParams(1.0))
    d = o.add_effect(abi.FX_DELAY, abi.DelayParams(0.1))
    with pytest.raises(Exception):
        o.link_control(g, g, 0)
    with pytest.raises(Exception):
        o.link_control(tap, d, 0)
    o.link_control(tap, g, 0)
    with pytest.raises(Exception):
        o.link_control(tap, g, 0)
