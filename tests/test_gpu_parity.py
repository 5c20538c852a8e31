"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI, against the CPU oracle.

Tolerance (BASELINE.json north_star): per-sample absolute error <= 1e-6 of full scale and 16-bit
PCM within +-1 LSB.  The kernels actually land around 1e-11; TIGHT is asserted too so a regression
in numerical quality is caught long before the contractual bound.
"""
import numpy as np
import pytest

from groove_b200 import abi, workloads
from tests import scenes
from tests.oracle_binding import OracleEngine, pcm16

pytestmark = pytest.mark.gpu

TOL = 1e-6
TIGHT = 1e-9


def gpu_engine(sr=44100.0, **kw):
    from groove_b200 import Engine
    return Engine(sr, **kw)


def both(scene, sr=44100.0, max_block=0, chunk=None):
    o = OracleEngine(sr)
    n = scene(o)
    ref = o.render(n)
    g = gpu_engine(sr, max_block=max_block)
    scene(g)
    if chunk is None:
        out = g.render(n)
    else:
        parts, done = [], 0
        while done < n:
            k = min(chunk, n - done)
            parts.append(g.render(k).copy())
            done += k
        out = np.concatenate(parts)
    st = g.stats()
    g.close()
    assert st.kernel_launches > 0
    return out, ref


def check(out, ref, tol=TOL, tight=TIGHT):
    err = float(np.abs(out - ref).max())
    assert err <= tol, f"max abs error {err}"
    assert err <= tight, f"numerical quality regressed: {err}"
    a = pcm16(np.clip(out, -1.0, 1.0)).astype(np.int32)
    b = pcm16(np.clip(ref, -1.0, 1.0)).astype(np.int32)
    assert int(np.abs(a - b).max()) <= 1


@pytest.mark.parametrize("name", list(scenes.ALL_SCENES))
@pytest.mark.parametrize("max_block", [0, 1000, 257])
def test_scene_parity(name, max_block):
    out, ref = both(scenes.ALL_SCENES[name], max_block=max_block)
    check(out, ref)


@pytest.mark.parametrize("chunk", [17, 64, 4099])
def test_caller_buffer_size_does_not_change_audio(chunk):
    """orchestrator.rs:1683 uses a prime buffer (17): state must carry exactly across render calls."""
    out, ref = both(scenes.scene_welsh_variants, chunk=chunk)
    check(out, ref)
    out, ref = both(scenes.scene_effects_rack, chunk=chunk)
    check(out, ref)


@pytest.mark.parametrize("lookahead", [64, 4096])
def test_lookahead_serves_small_buffers_with_the_same_audio(lookahead):
    """gb_set_lookahead (include/groove_b200.h): the Orchestrator's 64-frame tick buffers (orchestrator.rs:1696)
    are served from one big device render per `lookahead` frames; audio and position as without it."""
    for scene in (scenes.scene_welsh_variants, scenes.scene_effects_rack):
        o = OracleEngine()
        n = scene(o)
        ref = o.render(n)
        g = gpu_engine()
        scene(g)
        g.set_lookahead(lookahead)
        parts, done = [], 0
        sizes = [64, 64, 17, 64, 3 * lookahead + 5]   # small calls, a ragged one, one larger than the look-ahead
        i = 0
        while done < n:
            k = min(sizes[i % len(sizes)], n - done)
            parts.append(g.render(k).copy())
            done += k
            i += 1
            assert g.position == done
            if i == 3:                              # 145 frames handed out: frames rendered ahead are pending
                with pytest.raises(abi.GrooveError):
                    g.set_lookahead(0)
                with pytest.raises(abi.GrooveError):
                    g.save_state()
        launches = g.stats().kernel_launches
        g.close()
        check(np.concatenate(parts), ref)
        g2 = gpu_engine()
        scene(g2)
        done = 0
        while done < n:
            k = min(64, n - done)
            g2.render(k)
            done += k
        if lookahead > 64:
            assert launches < g2.stats().kernel_launches / 4
        g2.close()


def test_effect_batching_and_fusion(monkeypatch):
    """Independent effects of one kind and graph level share a launch, and a memoryless effect behind an IIR
    stage runs as its post-op: same audio as one launch per node (GB_FX_BATCH=0 GB_FX_FUSE=0), fewer launches."""
    out, ref = both(scenes.scene_fx_chains, max_block=4096)
    check(out, ref)
    def run():
        g = gpu_engine(max_block=4096)
        n = scenes.scene_fx_chains(g)
        y = g.render(n).copy()
        st = g.stats()
        g.close()
        return y, st.kernel_launches, st.fx_batched_nodes
    y1, l1, b1 = run()
    monkeypatch.setenv("GB_FX_BATCH", "0")
    monkeypatch.setenv("GB_FX_FUSE", "0")
    y0, l0, b0 = run()
    assert b0 == 0 and b1 > 0
    assert l1 < l0 - 5 * 6            # per chunk: 4 lp24 + 4 biquad stages in 2 launches, 6 fused post-ops
    assert float(np.abs(y1 - y0).max()) <= 1e-15
    check(y1, ref)


def test_graph_semantics_on_gpu():
    """orchestrator.rs:1444-1668 restated on the GPU engine: silence, sums, gain chains, branch."""
    g = gpu_engine()
    g.finalize()
    assert np.all(g.render(100) == 0.0)
    g.close()
    out, ref = both(scenes.scene_graph_toys)
    assert np.allclose(out, 0.1 + 0.5 * (0.3 + 0.5), atol=1e-15)
    assert np.array_equal(out, ref)


def test_pcm16_output_matches_oracle_conversion():
    g = gpu_engine()
    n = scenes.scene_drums_and_sampler(g)
    pcm = g.render_pcm16(n)
    g.close()
    o = OracleEngine()
    scenes.scene_drums_and_sampler(o)
    ref = pcm16(o.render(n))
    assert pcm.dtype == np.int16 and pcm.shape == (n, 2)
    assert int(np.abs(pcm.astype(np.int32) - ref.astype(np.int32)).max()) <= 1
    assert np.abs(ref).max() == 32767  # the scene clips: saturation is exercised


def test_device_resident_render_and_readback():
    g = gpu_engine()
    n = scenes.scene_fm(g)
    done = g.render_device(n)
    assert done == n
    ptr, frames = g.last_device_buffer()
    assert ptr and frames == n
    out = g.read_last(n)
    g.close()
    o = OracleEngine()
    scenes.scene_fm(o)
    check(out, o.render(n))


def test_save_restore_state_roundtrip():
    g = gpu_engine()
    n = scenes.scene_effects_rack(g)
    first = g.render(5000).copy()
    blob = g.save_state()
    rest_a = g.render(n - 5000).copy()
    g.restore_state(blob)
    assert g.position == 5000
    rest_b = g.render(n - 5000).copy()
    g.close()
    assert np.array_equal(rest_a, rest_b)
    o = OracleEngine()
    scenes.scene_effects_rack(o)
    check(np.concatenate([first, rest_a]), o.render(n))


def test_high_q_low_cutoff_filter_scan():
    """SURVEY.md §7 hard part: pole radius ~0.997+ (40 Hz, high resonance) and the 0.04 Hz / Q 0.05
    low-pass of test-data/perf-1.json; the scan must not lose precision."""
    def scene(r):
        p = scenes.generic_welsh(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_SQUARE, cutoff_start=scenes.hz_to_pct(40.0),
                                 cutoff_end=0.05, ripple=denorm(1.0), voices=2, filt=(0.0, 1.5, 0.3, 1.5))
        u = r.add_instrument(abi.INST_WELSH, p)
        f1 = r.add_effect(abi.FX_LOW_PASS_12DB, abi.BiquadParams(0.04, 0.05))
        f2 = r.add_effect(abi.FX_ALL_PASS_12DB, abi.BiquadParams(40.0, 20.0))
        r.patch_chain([u, f2, abi.MAIN_MIXER])
        r.patch_chain([u, f1, abi.MAIN_MIXER])
        r.finalize()
        r.note_on(0, u, 40)
        r.note_off(30000, u, 40)
        return 44100
    denorm = lambda q: q * q * 10.0 + 0.707
    out, ref = both(scene)
    check(out, ref)


def test_empty_and_ragged_inputs():
    g = gpu_engine()
    u = g.add_instrument(abi.INST_DRUMKIT, abi.DrumkitParams())       # no samples loaded
    s = g.add_instrument(abi.INST_SAMPLER, abi.SamplerParams(440.0, 2, 0))
    g.patch(u, abi.MAIN_MIXER)
    g.patch(s, abi.MAIN_MIXER)
    g.finalize()
    g.note_on(0, u, 35)
    g.note_on(0, s, 60)
    assert g.render(0).shape == (0, 2)
    assert np.all(g.render(1) == 0.0)
    assert np.all(g.render(1001) == 0.0)
    g.close()


def test_cfg4_slice_parity_and_linearity():
    """Config 4 on a slice the oracle finishes in seconds, then a size-independent property at a
    larger size: rendering voices A and B together equals rendering them apart and summing."""
    cfgA = workloads.cfg4_slice(128, 20000)
    o = OracleEngine(48000.0)
    workloads.build_cfg4(o, cfgA)
    ref = o.render(20000)
    g = gpu_engine(48000.0)
    workloads.build_cfg4(g, cfgA)
    out = g.render(20000)
    g.close()
    check(out, ref)

    frames = 200_000
    def run(voices, offset):
        e = gpu_engine(48000.0)
        workloads.build_cfg4(e, workloads.cfg4_slice(voices, frames, offset))
        y = e.render(frames)
        e.close()
        return y
    whole = run(256, 0)
    parts = run(128, 0) + run(128, 128)
    # cfg4_slice(256) groups voices by i mod 128, the halves by i mod 128 as well: same voices
    assert np.abs(whole - parts).max() < 1e-12
    assert np.abs(whole).max() > 1e-3


@pytest.mark.parametrize("solo_min", [None, "0"])
def test_cfg5_variant_batch_parity(solo_min, monkeypatch):
    """BASELINE config 5 on a batch the oracle finishes in seconds: 96 randomised one-voice
    subtractive / FM variants (solo-warp work items, short envelopes -> exact and general paths).
    GB_SOLO_MIN=0 forces the job-list kernel (welsh_solo_kernel) that larger batches take by themselves."""
    if solo_min is not None:
        monkeypatch.setenv("GB_SOLO_MIN", solo_min)
    frames, note_off = 24000, 12000
    o = OracleEngine(48000.0)
    workloads.build_cfg5(o, 96, first=1000, frames=frames, note_off=note_off)
    ref = o.render(frames)
    g = gpu_engine(48000.0, max_block=frames)
    workloads.build_cfg5(g, 96, first=1000, frames=frames, note_off=note_off)
    out = g.render(frames)
    st = g.stats()
    g.close()
    assert (st.solo_kernel_launches > 0) == (solo_min == "0")
    assert np.abs(ref).max() > 1e-3
    check(out, ref)


def test_solo_job_list_classes(monkeypatch):
    """welsh_solo_kernel: the host sorts every (solo voice, 2048-frame sub-chunk) into idle / resting /
    sweeping / general.  Instruments of 1-3 voices with envelopes slow enough for the sweeping class in
    attack, decay AND release, a released voice whose filter envelope runs out before its amplitude
    envelope (released resting tables), amplitude LFOs, a fixed filter, sine / noise oscillators and hard
    sync (general class), a retrigger, several render chunks with a ragged tail and idle stretches whose
    buffers held audio in the chunk before."""
    monkeypatch.setenv("GB_SOLO_MIN", "0")
    slow = dict(filt=(1.6, 1.0, 0.5, 1.6), amp=(0.3, 0.4, 0.7, 1.9), cutoff_start=0.3, cutoff_end=0.5)
    cfgs = [
        (1, dict(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_TRIANGLE, tune2=1.0029, **slow)),
        (2, dict(w1=abi.WAVE_PULSE_WIDTH, pw1=0.2, w2=abi.WAVE_SQUARE, routing=abi.LFO_AMPLITUDE, depth=0.3, lfo_hz=6.0, **slow)),
        (3, dict(w1=abi.WAVE_TRIANGLE, w2=abi.WAVE_SAWTOOTH, mix=0.3, filt=(0.002, 0.03, 0.4, 0.05), amp=(0.01, 0.02, 0.8, 1.2),
                 cutoff_start=0.25, cutoff_end=0.7)),                                   # filter release ends long before the amp's
        (1, dict(w1=abi.WAVE_SQUARE, w2=abi.WAVE_PULSE_WIDTH, pw2=0.3, cutoff_end=0.0, cutoff_hz=1200.0, amp=(0.2, 0.3, 0.6, 0.8),
                 routing=abi.LFO_AMPLITUDE, depth=0.2, lfo_hz=9.0)),                    # fixed filter
        (1, dict(w1=abi.WAVE_SINE, w2=abi.WAVE_NOISE, **slow)),                          # general class throughout
        (2, dict(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_SQUARE, sync=1, tune2=1.5, **slow)),
        (1, dict(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_SAWTOOTH, tune2=0.5, filt=(0.01, 0.05, 0.3, 0.1), amp=(0.001, 0.05, 0.5, 0.05),
                 cutoff_start=0.2, cutoff_end=0.9)),                                   # fast sweeps: exact coefficients
    ]
    frames = 230_000

    def scene(r):
        uids = []
        for rep in range(3):
            for i, (nv, c) in enumerate(cfgs):
                u = r.add_instrument(abi.INST_WELSH, scenes.generic_welsh(voices=nv, gain=0.04, pan=-0.8 + 0.2 * i + 0.05 * rep, **c))
                r.patch(u, abi.MAIN_MIXER)
                uids.append((u, nv, rep))
        r.finalize()
        for k, (u, nv, rep) in enumerate(uids):
            for v in range(nv):
                on = 300 + 977 * rep + 131 * k + 4000 * v
                r.note_on(on, u, 40 + 5 * v + k % 7)
                r.note_off(on + 140_000 - 9000 * rep, u, 40 + 5 * v + k % 7)
            if rep == 1:
                r.note_on(60_000 + k, u, 40 + k % 7)        # retrigger of voice 0 inside its decay / sustain
        return frames

    o = OracleEngine(48000.0)
    scene(o)
    ref = o.render(frames)
    for max_block, chunk in ((65536, None), (10_000, None), (65536, 33_333)):
        g = gpu_engine(48000.0, max_block=max_block)
        scene(g)
        if chunk is None:
            out = g.render(frames)
        else:
            out = np.concatenate([g.render(min(chunk, frames - a)).copy() for a in range(0, frames, chunk)])
        st = g.stats()
        g.close()
        assert st.solo_kernel_launches > 0
        assert all(c > 0 for c in st.solo_class_items), list(st.solo_class_items)   # rest / sweep / general / exact
        assert st.idle_voice_samples > 0
        check(out, ref)


def test_cfg5_batch_is_sum_of_its_variants():
    """Size-independent property at a larger batch: the bus of 512 variants equals the sum of the
    buses of its two halves (independent renders)."""
    frames, note_off = 16000, 8000
    def run(n, first):
        e = gpu_engine(48000.0, max_block=frames)
        workloads.build_cfg5(e, n, first=first, frames=frames, note_off=note_off)
        y = e.render(frames)
        e.close()
        return y
    whole = run(512, 0)
    parts = run(256, 0) + run(256, 256)
    assert np.abs(whole).max() > 1e-2
    assert np.abs(whole - parts).max() < 1e-12


@pytest.mark.parametrize("vpc", [None, "24", "16", "41/4", "32/4"])
def test_resting_voices_kernel(vpc, monkeypatch):
    """Chunks in which every voice of a CTA rests (note held, both envelopes at sustain, no events in the
    chunk) go to welsh_rest_kernel: voice state cached in shared memory across the chunk, two voices per
    warp.  Instruments with 41 / 16 / 9 voices and forced voices-per-CTA splits cover the paired, the
    single and the accumulate variants, all four (LFO, flat-oscillator) classes, and the hand-over to and
    from the general kernel at note-on / note-off chunks.  "vpc/4": the four-voices-per-warp instantiation
    (GB_REST_NV=4; quads, then a pair or a single voice for the remainder of a CTA)."""
    if vpc:
        monkeypatch.setenv("GB_VPC", vpc.split("/")[0])
        if vpc.endswith("/4"):
            monkeypatch.setenv("GB_REST_NV", "4")
    fast_filt = (0.0, 0.01, 0.6, 0.05)
    cfgs = [
        (41, dict(w1=abi.WAVE_PULSE_WIDTH, pw1=0.1, w2=abi.WAVE_SQUARE, mix=0.5, routing=abi.LFO_AMPLITUDE, depth=0.05,
                  lfo_hz=7.5, filt=fast_filt, amp=(0.02, 0.0, 1.0, 0.0), cutoff_start=scenes.hz_to_pct(40.0), cutoff_end=0.9)),
        (16, dict(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_TRIANGLE, tune2=1.0029, routing=abi.LFO_NONE, filt=fast_filt,
                  amp=(0.001, 0.02, 0.8, 0.02), ripple=2.5, cutoff_start=0.2, cutoff_end=0.4)),
        (9, dict(w1=abi.WAVE_TRIANGLE, w2=abi.WAVE_PULSE_WIDTH, pw2=0.8, mix=0.3, cutoff_end=0.0, cutoff_hz=700.0,
                 routing=abi.LFO_AMPLITUDE, depth=0.5, lfo_hz=11.0, amp=(0.0, 0.01, 0.9, 0.01))),
        (12, dict(w1=abi.WAVE_SQUARE, w2=abi.WAVE_PULSE_WIDTH, pw2=0.3, routing=abi.LFO_NONE, filt=fast_filt,
                  amp=(0.0, 0.0, 1.0, 0.0), cutoff_start=0.3, cutoff_end=0.8)),
    ]
    frames = 6 * 4096 + 700

    def scene(r):
        uids = []
        for i, (nv, c) in enumerate(cfgs):
            u = r.add_instrument(abi.INST_WELSH, scenes.generic_welsh(voices=nv, gain=0.05, pan=-0.6 + 0.4 * i, **c))
            r.patch(u, abi.MAIN_MIXER)
            uids.append(u)
        r.finalize()
        for i, (nv, _) in enumerate(cfgs):
            for v in range(nv):
                r.note_on(5 + 7 * v + i, uids[i], 30 + v)
                r.note_off(4 * 4096 + 100 + 3 * v, uids[i], 30 + v)
        return frames

    o = OracleEngine(48000.0)
    scene(o)
    ref = o.render(frames)
    g = gpu_engine(48000.0, max_block=4096)
    scene(g)
    out = g.render(frames)
    st = g.stats()
    g.close()
    assert st.rest_kernel_launches >= 6, "chunks 1..3 should have taken the resting-voice kernel"
    assert st.rest_voice_samples > 0
    check(out, ref)


def test_voice_range_resting_chunks(monkeypatch):
    """Chunks in which EVERY grouped CTA rests run over voice ranges of 14 that ignore instrument boundaries
    (welsh_rest_vr_kernel: a CTA mixes the tail of one instrument with the head of the next, each warp taking its
    own instrument record and pan).  12 instruments x 16 voices with different cutoffs, pans and LFO depths
    = 14 ranges (the last one of 10 voices) against 24 instrument CTAs; GB_REST_VR=0 is the instrument-CTA
    render of the same scene.  Also covers the overlapped mixdown (12 split instruments, one consumer)."""
    monkeypatch.setenv("GB_MIN_CUT_VOICES", "1")
    frames = 10 * 4096 + 300

    def scene(r):
        uids = []
        for i in range(12):
            p = scenes.generic_welsh(voices=16, gain=0.04, pan=-0.9 + 0.16 * i, w1=abi.WAVE_PULSE_WIDTH, pw1=0.1 + 0.02 * i,
                                     w2=abi.WAVE_SQUARE, mix=0.5, routing=abi.LFO_AMPLITUDE, depth=0.05 + 0.01 * i,
                                     lfo_hz=7.5, filt=(0.0, 0.01, 0.6, 0.05), amp=(0.02, 0.0, 1.0, 0.0),
                                     cutoff_start=scenes.hz_to_pct(40.0), cutoff_end=0.5 + 0.03 * i)
            u = r.add_instrument(abi.INST_WELSH, p)
            r.patch(u, abi.MAIN_MIXER)
            uids.append(u)
        r.finalize()
        for i, u in enumerate(uids):
            for v in range(16):
                r.note_on(5 + 3 * v + i, u, 30 + 2 * v + i % 2)
                r.note_off(8 * 4096 + 100 + 3 * v, u, 30 + 2 * v + i % 2)
        return frames

    o = OracleEngine(48000.0)
    scene(o)
    ref = o.render(frames)
    outs, stats = [], []
    # voice ranges with 16 frames per lane (welsh_rest_vr16_kernel: 512-frame blocks, the passes' intermediate values
    # parked in the tile rows), with 8 frames per lane (welsh_rest_vr_kernel), and the instrument-CTA layout
    # (the 16-frame kernel twice: with the sign-bit form of the symmetric pulse / square oscillators and without)
    for vr, r16, sym in (("1", "1", "1"), ("1", "0", "1"), ("0", "1", "1"), ("1", "1", "0")):
        monkeypatch.setenv("GB_REST_VR", vr)
        monkeypatch.setenv("GB_REST16", r16)
        monkeypatch.setenv("GB_OSC_SYM", sym)
        g = gpu_engine(48000.0, max_block=4096)
        scene(g)
        outs.append(g.render(frames))
        stats.append(g.stats())
        g.close()
    for k in (0, 1, 3):
        assert stats[k].rest_kernel_launches >= 5 and stats[k].rest_ctas == 14 * stats[k].rest_kernel_launches
        assert stats[k].rest_vr_launches == stats[k].rest_kernel_launches
    assert stats[2].rest_ctas == 24 * stats[2].rest_kernel_launches and stats[2].rest_vr_launches == 0
    for y in outs:
        check(y, ref)
    assert np.abs(outs[0] - outs[2]).max() < 1e-12 and np.abs(outs[1] - outs[2]).max() < 1e-12
    assert np.array_equal(outs[0], outs[3])      # the sign-bit oscillator form yields the same bits


def test_voice_range_sweeping_chunks(monkeypatch):
    """Chunks in which EVERY grouped CTA sweeps (all voices inside the filter envelope's decay) also run over voice
    ranges (welsh_sweep_vr_kernel: a warp's two voices may belong to different instruments).  12 instruments x 16
    voices with a 1.5 s filter decay: 14 ranges against 24 instrument CTAs per sweeping launch."""
    monkeypatch.setenv("GB_MIN_CUT_VOICES", "1")
    frames = 20 * 4096

    def scene(r):
        uids = []
        for i in range(12):
            p = scenes.generic_welsh(voices=16, gain=0.04, pan=-0.9 + 0.16 * i, w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_PULSE_WIDTH,
                                     pw2=0.2 + 0.02 * i, mix=0.5, routing=abi.LFO_AMPLITUDE, depth=0.05 + 0.01 * i, lfo_hz=6.0,
                                     filt=(0.0, 1.5, 0.6, 0.05), amp=(0.01, 0.0, 1.0, 0.0),
                                     cutoff_start=scenes.hz_to_pct(60.0), cutoff_end=0.45 + 0.02 * i)
            u = r.add_instrument(abi.INST_WELSH, p)
            r.patch(u, abi.MAIN_MIXER)
            uids.append(u)
        r.finalize()
        for i, u in enumerate(uids):
            for v in range(16):
                r.note_on(5 + 3 * v + i, u, 30 + 2 * v + i % 2)
        return frames

    o = OracleEngine(48000.0)
    scene(o)
    ref = o.render(frames)
    outs, stats = [], []
    for vr in ("1", "0"):
        monkeypatch.setenv("GB_REST_VR", vr)
        g = gpu_engine(48000.0, max_block=4096)
        scene(g)
        outs.append(g.render(frames))
        stats.append(g.stats())
        g.close()
    assert stats[0].sweep_kernel_launches >= 10 and stats[0].sweep_ctas == 14 * stats[0].sweep_kernel_launches
    assert stats[1].sweep_kernel_launches >= 10 and stats[1].sweep_ctas == 24 * stats[1].sweep_kernel_launches
    check(outs[0], ref)
    check(outs[1], ref)
    assert np.abs(outs[0] - outs[1]).max() < 1e-12


@pytest.mark.parametrize("max_block", [0, 100, 64])
def test_sidechain_link_known_answer_on_gpu(max_block):
    """The oracle's sidechain known answer (tests/test_oracle_known_answers.py) on the CUDA engine: the
    target's parameter table is built on the device from the source's rendered signal, across chunk
    boundaries that do and do not coincide with the 64-frame control boundaries."""
    g = gpu_engine(max_block=max_block)
    a = g.add_instrument(abi.INST_TOY_SOURCE, abi.ToySourceParams(0.3, 0.3))
    b = g.add_instrument(abi.INST_TOY_SOURCE, abi.ToySourceParams(0.5, 0.5))
    tap = g.add_effect(abi.FX_SIGNAL_PASSTHROUGH)
    gain = g.add_effect(abi.FX_GAIN, abi.GainParams(1.0))
    g.patch_chain([b, gain, abi.MAIN_MIXER])        # the target is patched first: the link must reorder the plan
    g.patch_chain([a, tap, abi.MAIN_MIXER])
    g.link_control(tap, gain, 0)
    with pytest.raises(Exception):
        g.link_control(tap, gain, 0)
    with pytest.raises(Exception):
        g.link_control(gain, gain, 0)
    g.finalize()
    y = np.concatenate([g.render(37), g.render(163)])
    g.close()
    assert np.allclose(y[:64], 0.3 + 0.5 * 1.0, atol=1e-15)
    assert np.allclose(y[64:], 0.3 + 0.5 * 0.3, atol=1e-15)


def test_resting_paths_agree_at_size(monkeypatch):
    """Size-independent property at a larger size than the oracle can check: the three ways the engine
    renders a resting voice — welsh_rest_kernel (whole chunks), welsh_block_lti inside welsh_kernel (per
    block, GB_REST_KERNEL=0) and the knot-interpolated moving-cutoff block evaluated at a cutoff that
    happens not to move (GB_LTI=0) — agree far inside the parity tolerance on 2048 config-4 voices over
    a stretch that runs from the filter decay into the resting state."""
    frames = 5 * 65536
    cfg = workloads.cfg4_slice(2048, frames)      # 128 instruments x 16 voices: grouped CTAs
    cfg.note_off_base = 4 * 65536 + 1000

    def run(env):
        for k in ("GB_REST_KERNEL", "GB_LTI", "GB_SWEEP_KERNEL", "GB_CHUNK_CUTS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        e = gpu_engine(48000.0)
        workloads.build_cfg4(e, cfg)
        y = e.render(frames)
        st = e.stats()
        e.close()
        return y, st

    a, sa = run({})
    b, sb = run({"GB_REST_KERNEL": "0"})
    c, sc = run({"GB_LTI": "0"})
    d, sd = run({"GB_SWEEP_KERNEL": "0"})
    f, sf = run({"GB_CHUNK_CUTS": "0"})       # chunks of max_block instead of cuts at the transition clusters
    assert sa.rest_kernel_launches > 0 and sb.rest_kernel_launches == 0 and sc.rest_kernel_launches == 0
    assert sa.sweep_kernel_launches > 0 and sd.sweep_kernel_launches == 0   # chunks 1-2: filter decay, no events
    assert np.abs(a).max() > 1e-3
    assert np.abs(a - b).max() < 1e-12
    assert np.abs(a - c).max() < 1e-10
    assert np.abs(a - d).max() < 1e-10      # one-sided vs centred coefficient knots
    assert np.abs(a - f).max() < 1e-10
    assert sf.kernel_launches < sa.kernel_launches


def test_split_instruments_summed_by_their_consumer(monkeypatch):
    """Instruments split over several CTAs write partial buffers.  When their only consumer sums more than
    kMaxSources inputs from a pointer table, the table lists the partial buffers themselves and the
    separate reduce pass is skipped (GB_FUSED_SUMS=0 restores it).  Both against the oracle; idle CTAs at
    the end of the render are not launched at all (their buffers are zeroed once)."""
    monkeypatch.setenv("GB_VPC", "8")
    frames = 9000

    def scene(r):
        uids = []
        for i in range(12):
            p = scenes.generic_welsh(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_PULSE_WIDTH, pw2=0.3, voices=20, gain=0.05,
                                     pan=-0.9 + 0.15 * i, filt=(0.0, 0.02, 0.5, 0.02), amp=(0.002, 0.01, 0.8, 0.01))
            u = r.add_instrument(abi.INST_WELSH, p)
            r.patch(u, abi.MAIN_MIXER)
            uids.append(u)
        r.finalize()
        for i, u in enumerate(uids):
            for v in range(20):
                r.note_on(3 * i + 11 * v, u, 30 + v)
                r.note_off(5000 + 7 * v + i, u, 30 + v)
        # gain / pan automation on a split instrument (its node buffer is needed for the DCA pass: the
        # shortcut must step aside for that chunk), and an instrument that wakes up again after idling
        r.control(2048 + 64, uids[3], 0, 0.5)      # GB_CTL_INST_DCA_GAIN
        r.control(3000, uids[3], 1, 0.9)           # GB_CTL_INST_DCA_PAN
        r.note_on(7300, uids[5], 50)
        r.note_off(7900, uids[5], 50)
        return frames

    o = OracleEngine(48000.0)
    scene(o)
    ref = o.render(frames)
    outs = []
    for fused in ("1", "0"):
        monkeypatch.setenv("GB_FUSED_SUMS", fused)
        g = gpu_engine(48000.0, max_block=1024)
        scene(g)
        outs.append(g.render(frames))
        g.close()
    check(outs[0], ref)
    check(outs[1], ref)
    assert np.abs(outs[0] - outs[1]).max() < 1e-14
    assert np.all(outs[0][-512:] == 0.0)      # idle tail


def test_save_restore_across_resting_chunks_and_sidechain():
    """State save / restore in the middle of a render whose chunks use the resting-voice kernel and a
    sidechain link (device-side link state, cached filter state written back per chunk), and the
    multi-chunk PCM16 path: the resumed render continues bit-identically."""
    def scene(r):
        u = r.add_instrument(abi.INST_WELSH, scenes.generic_welsh(
            w1=abi.WAVE_PULSE_WIDTH, pw1=0.2, w2=abi.WAVE_SQUARE, voices=16, gain=0.1, routing=abi.LFO_AMPLITUDE,
            depth=0.1, lfo_hz=6.0, filt=(0.0, 0.01, 0.6, 0.05), amp=(0.002, 0.0, 1.0, 0.01)))
        tap = r.add_effect(abi.FX_SIGNAL_PASSTHROUGH)
        d = r.add_instrument(abi.INST_FM, scenes.fm_params(car=(0.0, 0.08, 0.0, 0.08), gain=0.8))
        comp = r.add_effect(abi.FX_COMPRESSOR, abi.CompressorParams(1.0, 0.3, 0.0, 0.0))
        r.patch_chain([d, tap, abi.MAIN_MIXER])
        r.patch_chain([u, comp, abi.MAIN_MIXER])
        r.link_control(tap, comp, 0)
        r.finalize()
        for v in range(16):
            r.note_on(2 + v, u, 36 + v)
            r.note_off(14000 + v, u, 36 + v)
        for k in range(5):
            r.note_on(100 + 3000 * k, d, 40 + k)
            r.note_off(1500 + 3000 * k, d, 40 + k)
        return 16384

    g = gpu_engine(48000.0, max_block=2048)
    n = scene(g)
    first = g.render(6144).copy()
    blob = g.save_state()
    rest_a = g.render(n - 6144).copy()
    st = g.stats()
    g.restore_state(blob)
    rest_b = g.render(n - 6144).copy()
    g.close()
    assert st.rest_kernel_launches > 0
    assert np.array_equal(rest_a, rest_b)
    o = OracleEngine(48000.0)
    scene(o)
    ref = o.render(n)
    check(np.concatenate([first, rest_a]), ref)
    g = gpu_engine(48000.0, max_block=2048)
    scene(g)
    pcm = g.render_pcm16(n)
    g.close()
    assert int(np.abs(pcm.astype(np.int32) - pcm16(np.clip(ref, -1.0, 1.0)).astype(np.int32)).max()) <= 1


@pytest.mark.parametrize("vpc", [None, "16"])
def test_sweeping_voices_kernel(vpc, monkeypatch):
    """Chunks in which every voice of a CTA is held inside one moving stage of the filter envelope (and
    one stage of the amplitude envelope) with no note event go to welsh_sweep_kernel: one exact
    coefficient knot per lane, one-sided quadratic, b0 from the unity DC gain.  Attack and decay stages,
    LFO and flat-oscillator variants, a retriggered note (attack from a non-zero level), against the oracle."""
    if vpc:
        monkeypatch.setenv("GB_VPC", vpc)
    cfgs = [
        (24, dict(w1=abi.WAVE_PULSE_WIDTH, pw1=0.1, w2=abi.WAVE_SQUARE, mix=0.5, routing=abi.LFO_AMPLITUDE, depth=0.05,
                  lfo_hz=7.5, filt=(0.0, 2.0, 0.7, 2.0), amp=(0.02, 0.0, 1.0, 0.0), cutoff_start=scenes.hz_to_pct(40.0),
                  cutoff_end=0.9)),
        (16, dict(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_TRIANGLE, tune2=1.0029, routing=abi.LFO_NONE, filt=(3.0, 0.5, 0.3, 0.5),
                  amp=(0.3, 0.4, 0.6, 0.1), ripple=2.5, cutoff_start=0.2, cutoff_end=0.6)),
        (9, dict(w1=abi.WAVE_SQUARE, w2=abi.WAVE_PULSE_WIDTH, pw2=0.3, routing=abi.LFO_NONE, filt=(2.5, 0.4, 0.5, 0.3),
                 amp=(0.0, 0.0, 1.0, 0.0), cutoff_start=0.3, cutoff_end=0.8)),
    ]
    frames = 10 * 4096 + 300

    def scene(r):
        uids = []
        for i, (nv, c) in enumerate(cfgs):
            u = r.add_instrument(abi.INST_WELSH, scenes.generic_welsh(voices=nv, gain=0.05, pan=-0.5 + 0.5 * i, **c))
            r.patch(u, abi.MAIN_MIXER)
            uids.append(u)
        r.finalize()
        for i, (nv, _) in enumerate(cfgs):
            for v in range(nv):
                r.note_on(5 + 7 * v + i, uids[i], 30 + v)
                r.note_off(9 * 4096 + 100 + 3 * v, uids[i], 30 + v)
            r.note_on(3 * 4096 + 17, uids[i], 30)      # retrigger: the attack restarts from the current level
        return frames

    o = OracleEngine(48000.0)
    scene(o)
    ref = o.render(frames)
    g = gpu_engine(48000.0, max_block=4096)
    scene(g)
    out = g.render(frames)
    st = g.stats()
    g.close()
    assert st.sweep_kernel_launches >= 6
    assert st.sweep_voice_samples > 0
    check(out, ref)


def test_cfg4_full_length_against_oracle(monkeypatch):
    """The headline workload pinned end to end: a config-4 slice (2 instruments x 16 voices: grouped CTAs)
    for ALL 2 880 000 frames on the GPU and on the oracle — note-ons, the 3.29 s filter decay (sweeping
    kernel), the decay -> rest hand-over, 46.7 s of resting state (resting kernel; oscillator phases, LFO
    rotation re-seeded per chunk, scan under constant span maps over ~45 chunks), note-offs at 50 s and the
    silent tail.  Chunk cuts are switched on as in the 4096-voice configuration (state carried across
    every kernel hand-over, orchestrator.rs:856-877)."""
    monkeypatch.setenv("GB_MIN_CUT_VOICES", "1")
    cfg = workloads.Cfg4(total_voices=32, groups=2)
    assert cfg.frames == 2_880_000
    o = OracleEngine(48000.0)
    workloads.build_cfg4(o, cfg)
    ref = o.render(cfg.frames)
    g = gpu_engine(48000.0)
    workloads.build_cfg4(g, cfg)
    out = g.render(cfg.frames)
    st = g.stats()
    g.close()
    assert st.rest_kernel_launches > 0 and st.sweep_kernel_launches > 0
    assert st.rest_voice_samples > 0.7 * 32 * cfg.frames      # the resting stretch dominates, as in the benchmark
    assert np.abs(ref[2_000_000:2_400_000]).max() > 1e-4      # still sounding after 41 s
    assert np.all(ref[2_600_000:] == 0.0) and np.all(out[2_600_000:] == 0.0)
    check(out, ref)


def test_held_chord_scene_uses_rest_and_sweep_kernels():
    """The scene smoke() renders: both specialised kernels must actually launch (max_block 4096)."""
    o = OracleEngine(44100.0)
    n = scenes.scene_cello_held_chord(o)
    ref = o.render(n)
    g = gpu_engine(44100.0, max_block=4096)
    scenes.scene_cello_held_chord(g)
    out = g.render(n)
    st = g.stats()
    g.close()
    assert st.rest_kernel_launches > 0 and st.sweep_kernel_launches > 0
    check(out, ref)


@pytest.mark.parametrize("lfo", [False, True])
def test_hard_sync_patch_takes_the_sync_kernels(lfo, monkeypatch):
    """Hard-sync patches (settings/src/patches.rs:122; `piano`, 18 of the 106 Welsh patches) in grouped CTAs:
    the filter decay runs in welsh_sweep_kernel<.., SYNC>, the sustain in welsh_rest_kernel<.., SYNC> —
    oscillator 2 restarting at every wrap of oscillator 1 — and must match the general kernel's render
    (GB_SYNC_KERNELS=0) and the oracle.  Held 5.8 s so that both stages and the release are covered."""
    frames = 330_000

    def scene(r):
        p = workloads.piano_params(16, 1.0 / 16.0, 0.2)
        if lfo:
            p.lfo = abi.osc(abi.WAVE_SINE, frequency=5.5)
            p.lfo_routing = abi.LFO_AMPLITUDE
            p.lfo_depth = 0.3
        u = r.add_instrument(abi.INST_WELSH, p)
        r.patch(u, abi.MAIN_MIXER)
        r.finalize()
        ev = []
        for i in range(16):
            ev.append((100 + 37 * i, u, abi.EV_NOTE_ON, 30 + 3 * i, 127, 0.0))
            ev.append((280_000 + 64 * i, u, abi.EV_NOTE_OFF, 30 + 3 * i, 0, 0.0))
        r.push_events(ev)
        return frames

    o = OracleEngine(48000.0)
    scene(o)
    ref = o.render(frames)

    def run():
        g = gpu_engine(48000.0, max_block=8192)
        scene(g)
        y = g.render(frames).copy()
        st = g.stats()
        g.close()
        return y, st
    out, st = run()
    assert st.rest_kernel_launches > 0 and st.sweep_kernel_launches > 0
    check(out, ref)
    monkeypatch.setenv("GB_SYNC_KERNELS", "0")
    gen, st0 = run()
    assert st0.rest_kernel_launches == 0 and st0.sweep_kernel_launches == 0
    check(gen, ref)
    assert float(np.abs(out - gen).max()) <= 1e-10


@pytest.mark.parametrize("vpc,max_block", [(2, 8192), (3, 4352), (8, 65536), (4, 2048)])
def test_time_parallel_resting_kernel(vpc, max_block, monkeypatch):
    """welsh_rest_tp_kernel: CTAs of 2..8 voices whose 8 warps take consecutive 256-frame blocks of the same
    voice pair (a strong-scaling shard's shape).  Forced here with GB_VPC on the 16-voice held chord; chunk
    sizes cover full rounds (8 blocks), partial rounds (4352 = 17 blocks) and single-round chunks; vpc 3 gives
    CTAs with a pair and a single voice.  Must match the oracle and the per-warp kernel (GB_REST_TP=0)."""
    monkeypatch.setenv("GB_VPC", str(vpc))
    monkeypatch.setenv("GB_MIN_CUT_VOICES", "1")
    o = OracleEngine(44100.0)
    n = scenes.scene_cello_held_chord(o)
    ref = o.render(n)

    def run():
        g = gpu_engine(44100.0, max_block=max_block)
        scenes.scene_cello_held_chord(g)
        y = g.render(n).copy()
        st = g.stats()
        g.close()
        return y, st
    out, st = run()
    assert st.rest_tp_launches > 0
    check(out, ref)
    monkeypatch.setenv("GB_REST_TP", "0")
    base, st0 = run()
    assert st0.rest_tp_launches == 0 and st0.rest_kernel_launches > 0
    assert float(np.abs(out - base).max()) <= 1e-12


def test_time_parallel_resting_kernel_hard_sync(monkeypatch):
    """The SYNC instantiation of the time-parallel kernel (oscillator 2's restarts seen from a block offset)."""
    monkeypatch.setenv("GB_VPC", "2")
    frames = 330_000

    def scene(r):
        p = workloads.piano_params(16, 1.0 / 16.0, -0.4)
        u = r.add_instrument(abi.INST_WELSH, p)
        r.patch(u, abi.MAIN_MIXER)
        r.finalize()
        ev = []
        for i in range(16):
            ev.append((50 + 41 * i, u, abi.EV_NOTE_ON, 33 + 2 * i, 127, 0.0))
            ev.append((300_000 + 64 * i, u, abi.EV_NOTE_OFF, 33 + 2 * i, 0, 0.0))
        r.push_events(ev)
        return frames

    o = OracleEngine(48000.0)
    scene(o)
    ref = o.render(frames)
    g = gpu_engine(48000.0, max_block=16384)
    scene(g)
    out = g.render(frames)
    st = g.stats()
    g.close()
    assert st.rest_tp_launches > 0
    check(out, ref)


def test_two_gpu_bus_reduce_matches_single_gpu_render(tmp_path):
    """N > 1 on real GPUs: two ranks (torchrun, NCCL) each render their shard of a config-4 slice, the
    stereo buses are summed onto rank 0 — by the product's bus exchange (peer_sum_kernel reading the ranks' CUDA
    IPC buffers over NVLink) and by one NCCL f64 reduce (the fallback) — and rank 0 compares both results with
    its own single-GPU render of the union and with the oracle."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "result.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(root, "tests", "nccl_bus_check.py"), str(out)]
    proc = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    import json
    res = json.loads(out.read_text())
    assert res["world"] == 2
    assert res["peak"] > 1e-3
    assert res["max_abs_vs_single_gpu"] < 1e-12
    assert res["max_abs_vs_oracle"] < TIGHT
    # the product's exchange: the root's peer_sum_kernel over the ranks' CUDA IPC buffers (NVLink P2P loads)
    assert res["p2p_error"] is None, res["p2p_error"]
    assert res["p2p_max_abs_vs_nccl"] < 1e-15 and res["p2p_max_abs_vs_oracle"] < TIGHT
