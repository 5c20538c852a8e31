"""CPU tests: the oracle against the committed golden vectors; the C ABI library's exports and
struct layout; host-side workload construction.  No GPU compute happens here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from groove_b200 import abi, workloads
from groove_b200.engine import LIB_PATH
from tests import scenes
from tests.oracle_binding import OracleEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "scenes.npz"))


@pytest.mark.parametrize("name", list(scenes.ALL_SCENES))
def test_oracle_matches_golden_vectors(name):
    o = OracleEngine(44100.0)
    n = scenes.ALL_SCENES[name](o)
    y = o.render(n)
    stats = GOLD[name + "/stats"]
    assert n == int(stats[0])
    # transcendental libm calls may differ in the last ulp between hosts: 1e-12 absolute
    assert np.allclose(y[:256], GOLD[name + "/head"], atol=1e-12, rtol=0)
    assert np.allclose(y[::61], GOLD[name + "/stride61"], atol=1e-12, rtol=0)
    assert y.sum() == pytest.approx(stats[1], abs=1e-8)
    assert (y * y).sum() == pytest.approx(stats[2], rel=1e-10)


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "groove_b200.h")).read()
    return sorted(set(re.findall(r"\b(gb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    """The drop-in boundary: libgroove_b200.so loads without a GPU and exports all of include/*.h."""
    assert os.path.exists(LIB_PATH), "build it first: python -m groove_b200.build"
    lib = C.CDLL(LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted("gb_" + s for s in abi.ABI_SYMBOLS) == declared


def test_create_fails_loudly_without_a_device():
    from tests.conftest import HAVE_GPU
    if HAVE_GPU:
        pytest.skip("a CUDA device is present")
    from groove_b200 import Engine
    with pytest.raises(abi.GrooveError) as ei:
        Engine(44100.0)
    assert ei.value.code == abi.ENODEV and "no CPU fallback" in str(ei.value)


def test_struct_layout_matches_header():
    """ctypes mirrors must have the C sizes (checked against sizes the oracle build reports through
    accepting/rejecting params_size)."""
    o = OracleEngine()
    for kind, st in ((abi.INST_WELSH, abi.WelshParams()), (abi.INST_FM, abi.FmParams()),
                     (abi.INST_SAMPLER, abi.SamplerParams()), (abi.INST_TOY_SOURCE, abi.ToySourceParams())):
        assert o.add_instrument(kind, st) > 1
    for kind, st in ((abi.FX_GAIN, abi.GainParams()), (abi.FX_LIMITER, abi.LimiterParams()),
                     (abi.FX_BITCRUSHER, abi.BitcrusherParams()), (abi.FX_COMPRESSOR, abi.CompressorParams()),
                     (abi.FX_DELAY, abi.DelayParams()), (abi.FX_CHORUS, abi.ChorusParams(2, 0.01, 1.0)),
                     (abi.FX_REVERB, abi.ReverbParams(0.5, 0.5)), (abi.FX_LOW_PASS_12DB, abi.BiquadParams(100, 1)),
                     (abi.FX_LOW_PASS_24DB, abi.Lowpass24Params(100, 1))):
        assert o.add_effect(kind, st) > 1
    with pytest.raises(abi.GrooveError):
        o.add_effect(abi.FX_GAIN, abi.LimiterParams())       # wrong size is rejected, not misread
    assert C.sizeof(abi.Event) == 32 and C.sizeof(abi.WelshParams) == 280


def test_event_errors():
    o = OracleEngine()
    u = o.add_instrument(abi.INST_FM, scenes.fm_params())
    o.patch(u, abi.MAIN_MIXER)
    o.finalize()
    o.render(10)
    with pytest.raises(abi.GrooveError):
        o.note_on(5, u, 60)          # in the past
    with pytest.raises(abi.GrooveError):
        o.note_on(20, 999, 60)       # unknown uid
    with pytest.raises(abi.GrooveError):
        o.patch(u, abi.MAIN_MIXER)   # frozen graph


def test_cfg4_recipe_shape():
    """4096 voices, distinct keys per instrument, staggered on/off."""
    cfg = workloads.Cfg4()
    assert cfg.voice_samples == 11_796_480_000
    o = OracleEngine(48000.0)
    small = workloads.cfg4_slice(256, 9000)
    assert workloads.build_cfg4(o, small) == 9000
    y = o.render(9000)
    assert np.all(y[0] == 0.0) and np.abs(y[8200:]).max() > 0    # attack starts from level 0 at frame 0
    per = {}
    for i in range(4096):
        per.setdefault(i % 128, set()).add(36 + i % 49)
    assert all(len(v) == 32 for v in per.values())


def test_run_length_formula():
    """orchestrator.rs:1723-1737: frames = ceil(beats * 60 / bpm * sample_rate); 4 beats @240 bpm @24 kHz = 24000."""
    from groove_b200.project import song_frames
    assert song_frames(4, 240.0, 24000.0) == 24000
    assert song_frames(8, 128.0, 44100.0) == 165375
    assert song_frames(4, 1024.0, 44100.0) == 10336


def test_rust_sys_crate_matches_header():
    """groove-b200-sys/ (source only: no Rust toolchain in this image) must not drift from the header: every
    entry point, every entity kind and every field of every params struct, in order."""
    hdr = open(os.path.join(ROOT, "include", "groove_b200.h")).read()
    rs = open(os.path.join(ROOT, "groove-b200-sys", "src", "lib.rs")).read()
    for sym in _header_symbols():
        assert re.search(r"pub fn %s\(" % sym, rs), sym
    m = re.search(r"#define GB_ABI_VERSION (\d+)", hdr)
    assert re.search(r"GB_ABI_VERSION: u32 = %s;" % m.group(1), rs)
    for name, value in re.findall(r"\b(GB_(?:INST|FX)_[A-Z0-9_]+) = (\d+)", hdr):
        assert re.search(r"pub const %s: i32 = %s;" % (name, value), rs), name
    ctype = {"int32_t": "i32", "uint32_t": "u32", "double": "f64", "int64_t": "i64", "uint64_t": "u64"}
    for body, name in re.findall(r"typedef struct \{(.*?)\} (gb_[a-z0-9_]+);", hdr, flags=re.S):
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            typ, rest = decl.split(None, 1)
            for f in rest.split(","):
                f = f.strip()
                arr = re.match(r"(\w+)\[(\d+)\]", f)
                fields.append((arr.group(1) if arr else f, typ, int(arr.group(2)) if arr else 0))
        rm = re.search(r"pub struct %s \{(.*?)\n    \}" % name, rs, flags=re.S)
        assert rm, name
        rfields = re.findall(r"pub (\w+): ([^,\n]+),", rm.group(1))
        assert len(rfields) == len(fields), (name, fields, rfields)
        for (fname, typ, n), (rname, rtyp) in zip(fields, rfields):
            assert rname == (fname + "_" if fname == "type" else fname), (name, fname, rname)
            want = ctype.get(typ, typ)
            assert rtyp == (f"[{want}; {n}]" if n else want), (name, fname, rtyp)
