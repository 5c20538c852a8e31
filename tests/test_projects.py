"""Reference projects (BASELINE configs 1-3 and the feedback-effect demos) as compiled plans.

CPU: the oracle reproduces the committed decimated renders; the loader reproduces the committed plans
when the reference tree is present.  GPU (-m gpu): the CUDA engine matches the oracle on every plan.
"""
import os

import numpy as np
import pytest

from groove_b200 import abi, project
from tests import plans
from tests.oracle_binding import OracleEngine, pcm16

REF = "/root/reference"


@pytest.mark.parametrize("name", plans.PLAN_NAMES)
def test_oracle_reproduces_committed_plan_renders(name):
    plan = plans.load_plan(name)
    o = OracleEngine(plan.sample_rate)
    project.build_plan(o, plan, plans.sample)
    y = o.render(plan.frames)
    gold = plans.oracle_renders()
    stats = gold[name + "/stats"]
    assert plan.frames == int(stats[0])
    assert np.allclose(y[::41], gold[name + "/stride41"], atol=1e-11, rtol=0)
    assert (y * y).sum() == pytest.approx(stats[2], rel=1e-9)


def test_config_lengths_match_the_survey():
    """config 1 = 165 375 frames (8 beats @128 bpm); configs 2/3 = 41 344 frames when
    the control trips (4 whole-note steps = 16 beats) hold the song open."""
    assert plans.load_plan("drums-filtered-24db").frames == 165375
    assert plans.load_plan("perf-1").frames == 41344
    assert plans.load_plan("kitchen-sink").frames == 41344


def test_perf1_and_kitchen_sink_are_the_same_song():
    """Same topology with older key spellings (SURVEY.md §5 schema drift) -> identical audio."""
    def run(name):
        plan = plans.load_plan(name)
        o = OracleEngine(plan.sample_rate)
        project.build_plan(o, plan, plans.sample)
        return o.render(plan.frames)
    assert np.array_equal(run("perf-1"), run("kitchen-sink"))


def test_events_are_quantised_to_64_frame_buffers():
    for name in plans.PLAN_NAMES:
        plan = plans.load_plan(name)
        assert all(ev[0] % 64 == 0 for ev in plan.events), name


def test_pattern_notes_are_velocity_127_one_step_long():
    """settings/src/lib.rs:55-77; drums-filtered: 16th notes at 128 bpm = 0.1171875 s = 5167.97 frames."""
    plan = plans.load_plan("drums-filtered-24db")
    ons = [ev for ev in plan.events if ev[2] == abi.EV_NOTE_ON]
    offs = [ev for ev in plan.events if ev[2] == abi.EV_NOTE_OFF]
    assert len(ons) == 44 and len(offs) == 44 and all(ev[4] == 127 for ev in ons)
    assert sorted({ev[3] for ev in ons}) == [35, 38, 42, 44]
    first_hat_off = min(ev[0] for ev in offs if ev[3] == 42)
    assert first_hat_off == (5167 // 64) * 64


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")
@pytest.mark.parametrize("name,rel", [("drums-filtered-24db", "projects/demos/effects/drums-filtered-24db.json"),
                                      ("perf-1", "test-data/perf-1.json"), ("kitchen-sink", "test-data/kitchen-sink.json")])
def test_loader_reproduces_committed_plans(name, rel):
    loader = project.ProjectLoader(os.path.join(REF, "assets"))
    plan = loader.load(os.path.join(REF, rel))
    gold = plans.load_plan(name)
    assert plan.frames == gold.frames and plan.cables == gold.cables
    assert [tuple(e) for e in plan.events] == [tuple(e) for e in gold.events]
    assert [e.uvid for e in plan.entities] == [e.uvid for e in gold.entities]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")
def test_every_reference_project_loads_or_fails_as_the_reference_does():
    """All 96 project fixtures of the reference (projects/**, test-data/*.json*): 94 compile to plans; the two that
    do not are Welsh patches the reference itself cannot deserialise (an LFO routing that is not in the enum,
    settings/src/patches.rs:269-278, and an LFO depth outside its range).  The older sampler schema of
    projects/tests/load-stereo-wav.json (midi-in and filename in one object) loads and renders a STEREO sample."""
    import glob
    files = sorted(glob.glob(os.path.join(REF, "projects", "**", "*.json*"), recursive=True) +
                   glob.glob(os.path.join(REF, "test-data", "*.json*")))
    assert len(files) == 96
    loader = project.ProjectLoader(os.path.join(REF, "assets"))
    bad = {}
    for f in files:
        try:
            loader.load(f)
        except Exception as ex:   # noqa: BLE001
            bad[os.path.relpath(f, REF)] = str(ex)
    assert sorted(bad) == ["projects/demos/instruments/welsh-harmonica.json", "projects/demos/instruments/welsh-octave-switch.json"], bad
    plan = loader.load(os.path.join(REF, "projects", "tests", "load-stereo-wav.json"))
    o = OracleEngine(plan.sample_rate)
    project.build_plan(o, plan, loader.sample)
    y = o.render(min(plan.frames, 60000))
    assert np.abs(y).max() > 0.1 and np.abs(y[:, 0] - y[:, 1]).max() > 1e-3


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")
def test_welsh_patch_mapping_quirks():
    """settings/src/patches.rs:87-170: release := decay for both envelopes; mix = m1/(m1+m2);
    cutoff_start from the 12 dB preset, cutoff_hz from the 24 dB preset; every patch file loads
    unless its LFO routing does not deserialise (README.md:79-80)."""
    import glob
    import json
    ok = bad = 0
    for path in sorted(glob.glob(os.path.join(REF, "assets/patches/welsh/*.json"))):
        patch = json.load(open(path))
        try:
            p = project.welsh_params_from_patch(patch)
        except ValueError:
            bad += 1
            continue
        ok += 1
        assert p["amp"][3] == p["amp"][1] and p["filt"][3] == p["filt"][1]
        assert 0.0 <= p["mix"] <= 1.0 and 0.0 <= p["cutoff_start"] <= 1.0
    assert ok >= 80 and ok + bad == 106   # 88 load; the rest use routings/waveforms/depths that do not deserialise
    cello = project.welsh_params_from_patch(json.load(open(os.path.join(REF, "assets/patches/welsh/cello.json"))))
    assert cello["mix"] == 0.5 and cello["lfo_depth"] == pytest.approx(0.05) and cello["cutoff_end"] == pytest.approx(0.9)
    assert cello["amp"] == [pytest.approx(0.06), 0.0, 1.0, 0.0] and cello["filt"][3] == pytest.approx(3.29)
    # oscillator-2-track=false (patches.rs:92-101): the fixed frequency lands on a throwaway oscillator only;
    # the returned voice params keep oscillator 2 tracking the note at ratio 1.0 (tune = Note)
    untracked = 0
    for path in sorted(glob.glob(os.path.join(REF, "assets/patches/welsh/*.json"))):
        patch = json.load(open(path))
        if patch.get("oscillator-2-track", True) or patch["oscillator-2"]["waveform"] == "none":
            continue
        try:
            p = project.welsh_params_from_patch(patch)
        except ValueError:
            continue
        untracked += 1
        assert p["fixed2"] == 0.0 and p["tune2"] == 1.0, path
    assert untracked >= 3   # the other untracked patches panic in the reference too (tune is not a Note)


def test_json5_subset_parser():
    text = """{ // comment
      title: 'x', "clock": {"bpm": 120, /* inline */ "time-signature": {top: 3, bottom: 4},},
      devices: [],
    }"""
    d = project.parse_json5(text)
    assert d["title"] == "x" and d["clock"]["time-signature"]["top"] == 3 and d["devices"] == []


def test_wav_roundtrip(tmp_path):
    pcm = (np.arange(200, dtype=np.int16).reshape(100, 2) - 100) * 300
    path = str(tmp_path / "t.wav")
    project.write_wav16(path, pcm, 44100)
    x, sr = project.read_wav(path)
    assert sr == 44100 and x.shape == (100, 2)
    assert np.array_equal(np.round(x * 32768).astype(np.int16), pcm)


# ------------------------------------------------------------------------------------- GPU ---
@pytest.mark.gpu
@pytest.mark.parametrize("name", plans.PLAN_NAMES)
def test_plan_parity_on_gpu(name):
    from groove_b200 import Engine
    plan = plans.load_plan(name)
    o = OracleEngine(plan.sample_rate)
    project.build_plan(o, plan, plans.sample)
    ref = o.render(plan.frames)
    g = Engine(plan.sample_rate)
    project.build_plan(g, plan, plans.sample)
    out = g.render(plan.frames)
    g.close()
    err = float(np.abs(out - ref).max())
    assert err <= 1e-6, err
    assert err <= 1e-9, f"numerical quality regressed: {err}"
    a = pcm16(np.clip(out, -1, 1)).astype(np.int32)
    b = pcm16(np.clip(ref, -1, 1)).astype(np.int32)
    assert int(np.abs(a - b).max()) <= 1


@pytest.mark.gpu
def test_config1_wav_is_within_one_lsb(tmp_path):
    """groove-cli --wav path: 16-bit stereo WAV of config 1 rendered by the GPU vs the oracle."""
    from groove_b200 import Engine
    plan = plans.load_plan("drums-filtered-24db")
    g = Engine(plan.sample_rate)
    project.build_plan(g, plan, plans.sample)
    pcm = g.render_pcm16(plan.frames)
    g.close()
    path = str(tmp_path / "drums.wav")
    project.write_wav16(path, pcm, plan.sample_rate)
    o = OracleEngine(plan.sample_rate)
    project.build_plan(o, plan, plans.sample)
    ref = pcm16(o.render(plan.frames))
    x, sr = project.read_wav(path)
    back = np.round(x * 32768).astype(np.int32)
    assert sr == 44100 and int(np.abs(back - ref.astype(np.int32)).max()) <= 1


def test_wav_root_note_metadata(tmp_path):
    """README.md:82-84: the sampler takes its root frequency from the WAV metadata when it can.  A synthetic
    file with a `smpl` chunk (MIDI unity note 64), one with an `acid` chunk (root 57, flag bit 1), one whose
    acid chunk says the root is not valid, and one without metadata."""
    import struct
    def wav(chunks):
        fmt = struct.pack("<HHIIHH", 1, 1, 44100, 88200, 2, 16)
        data = struct.pack("<4h", 0, 1000, -1000, 0)
        body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(data)) + data
        for cid, payload in chunks:
            body += cid + struct.pack("<I", len(payload)) + payload + (b"\0" if len(payload) & 1 else b"")
        return b"RIFF" + struct.pack("<I", len(body)) + body
    smpl = struct.pack("<9I", 0, 0, 22675, 64, 0, 0, 0, 0, 0)
    acid = lambda flags, note: struct.pack("<IHHfIHHf", flags, note, 0x8000, 0.0, 4, 4, 4, 120.0)
    cases = {"smpl.wav": ([(b"smpl", smpl)], 64), "acid.wav": ([(b"LIST", b"abc"), (b"acid", acid(2, 57))], 57),
             "acid-invalid.wav": ([(b"acid", acid(1, 57))], None), "plain.wav": ([], None),
             "both.wav": ([(b"acid", acid(2, 57)), (b"smpl", smpl)], 64)}
    for name, (chunks, want) in cases.items():
        path = tmp_path / name
        path.write_bytes(wav(chunks))
        assert project.wav_root_note(str(path)) == want, name
        x, sr = project.read_wav(str(path))      # the PCM reader is not confused by the extra chunks
        assert sr == 44100 and x.shape == (4,)
    loader = project.ProjectLoader(str(tmp_path))
    assert loader.sample_root_hz("acid.wav") == pytest.approx(220.0)
    assert loader.sample_root_hz("plain.wav") == 0.0
    ref = os.path.join(REF, "test-data", "samples", "riff-acidized.wav")
    if os.path.exists(ref):      # only in the build container
        assert project.wav_root_note(ref) == 57
        assert project.wav_root_note(os.path.join(REF, "test-data", "samples", "riff-not-acidized.wav")) is None
