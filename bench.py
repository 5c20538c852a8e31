#!/usr/bin/env python
"""bench.py — voice-samples/s of the Groove synthesis hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the CPU implementation (oracle) on host cores

A *step* is one full render of BASELINE config 4 — 4096 Welsh `cello` voices, 60 s at 48 kHz stereo
= 1.17965e10 voice-samples — through the engine.  At N > 1 every rank renders its own 4096-voice
shard (weak scaling) and the stereo buses are summed onto rank 0 by the engine's bus exchange: one kernel on
rank 0 whose loads of the other ranks' buses cross NVLink (CUDA IPC peer buffers; NCCL's f64 reduce only if the
peers cannot be mapped — `config.parallelism` says which).

Printed JSON (one line, rank 0):
  value      whole-job voice-samples/s, device-timed (CUDA events on the engine's stream around
             each render call; at N > 1 plus CUDA events around the bus exchange, max over
             ranks), inputs resident in HBM, result left in HBM;
  e2e        the same metric through the C ABI with host buffers: the note events are pushed from
             host memory, rendered, and the f64 stereo result is copied back to a host buffer, all
             inside the timed region (engine construction / plan / allocation stay outside);
  roofline   dominant kernel (config 4: the resting-voice kernel — welsh_rest_vr16_kernel when every voice of
             a chunk rests, `roofline.kernel` names what ran) against the FP64
             vector pipe, which is what binds this path: achieved = the ALGORITHMIC
             150 FLOP x voice-samples its launches covered / their CUDA-event time; peak = FP64 FMA
             microbenchmark measured live on the same GPU.  `roofline.executed` is the same with the FP64
             FLOP the kernel actually executes (it proves the per-frame coefficient work redundant
             while a voice rests: DESIGN.md §3.1a);
  legs       secondary workloads, each with its own value and roofline:
               cfg5          (every N) BASELINE config 5: 8192 one-shot FM / subtractive patch variants per
                             GPU (65 536 over 8), per-variant buffers + NCCL-reduced bus;
               strong        (N > 1) config 4 with its 4096 voices SPLIT over the N GPUs,
                             efficiency against a single-GPU render timed in the same run;
               time_varying  (N = 1) config 4 with the filter decay stretched past the note: the cutoff
                             moves on every frame, no voice rests;
               general       (N = 1) config 4's score on the hard-sync `piano` patch;
               stream_fx     (N = 1) 1024 copies of config 1's effect chain, HBM-bound;
               small_blocks  (N = 1) config 4 through 64-frame gb_render_block calls (the reference's own
                             call size, orchestrator.rs:1696);
  cpu_baseline  the CPU oracle (reference-structured restatement, -O3 -march=native) on a bounded
             sample, 1 core, in a subprocess.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "voice_samples_per_sec"
UNIT = "voice-samples/s"
SR = 48000.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--voices", type=int, default=4096, help="voices per GPU (config 4: 4096)")
    ap.add_argument("--seconds", type=float, default=60.0, help="audio seconds per step (config 4: 60)")
    ap.add_argument("--groups", type=int, default=0, help="instruments per GPU (config 4: 128; 0 = min(128, voices))")
    ap.add_argument("--max-block", type=int, default=1 << 16)
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg5"],
                    help="headline workload: cfg4 = BASELINE config 4; cfg5 = batch of one-shot patch variants")
    ap.add_argument("--variants", type=int, default=8192, help="cfg5: variants per GPU (config 5: 65536 over 8 GPUs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip every secondary leg")
    ap.add_argument("--legs", default="cfg5,strong,time_varying,general,stream_fx,small_blocks",
                    help="comma-separated secondary legs to run")
    ap.add_argument("--no-time-varying", action="store_true", help="(kept for older command lines) skip that leg")
    ap.add_argument("--cpu-sample-voices", type=int, default=64)
    ap.add_argument("--cpu-sample-seconds", type=float, default=15.0)
    ap.add_argument("--_cpu-child", default=None, help=argparse.SUPPRESS)
    return ap.parse_args()


# --------------------------------------------------------------------------- CPU legs ---
def native_oracle() -> str:
    """The oracle compiled -O3 -march=native for THIS host, into a temp file: the committed
    build recipe (oracle/Makefile, -O2) is what the parity tests load; the CPU timings use this one."""
    out = os.path.join(tempfile.gettempdir(), f"libgroove_oracle_native_{os.getuid()}.so")
    src = os.path.join(ROOT, "oracle", "groove_oracle.cpp")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        tmp = out + f".{os.getpid()}"
        subprocess.check_call(["g++", "-O3", "-march=native", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-o", tmp, src])
        os.replace(tmp, out)
    return out


def _oracle_render_slice(args):
    """Worker: render `voices` voices x `frames` frames of config 4 on the CPU oracle; returns seconds."""
    voices, frames, voice_offset, note_off = args
    from groove_b200 import workloads
    from tests.oracle_binding import OracleEngine
    o = OracleEngine(SR)
    cfg = workloads.cfg4_slice(voices, frames, voice_offset)
    cfg.note_off_base = note_off
    n = workloads.build_cfg4(o, cfg)
    t = time.perf_counter()
    done = 0
    out = np.empty((4800, 2))
    while done < n:                       # 64-frame-multiple buffers like the reference CLI loop
        k = min(4800, n - done)
        o.render(k, out)
        done += k
    return time.perf_counter() - t


def cpu_child(spec: str) -> None:
    """Subprocess body of the cpu_baseline leg: the GPU arm's process never maps oracle/."""
    voices, seconds = spec.split(",")
    voices, seconds = int(voices), float(seconds)
    frames = int(seconds * SR)
    dt = _oracle_render_slice((voices, frames, 0, int(frames * 2_400_000 / 2_880_000)))
    print(json.dumps({
        "value": voices * frames / dt, "unit": UNIT, "cores": 1, "kind": "port",
        "sample": f"config-4 recipe, voices 0..{voices - 1}, {seconds:g} s with the full render's time profile "
                  f"(note-off at 5/6 of the length; {voices * frames:.3e} voice-samples), oracle/groove_oracle.cpp "
                  f"per-frame graph walk, g++ -O3 -march=native -ffp-contract=off, {dt:.2f} s wall",
    }), flush=True)


def cpu_baseline_subprocess(voices: int, seconds: float) -> dict:
    env = dict(os.environ, GROOVE_ORACLE_SO=native_oracle())
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--_cpu-child", f"{voices},{seconds}"],
                         env=env, capture_output=True, text=True, timeout=600)
    if out.returncode != 0:
        return {"error": out.stderr[-400:]}
    return json.loads(out.stdout.strip().splitlines()[-1])


def run_reference_arm(a) -> None:
    """--impl reference: the CPU implementation of the path on all host cores.

    The reference itself cannot be built (no Rust toolchain, DSP source absent: SURVEY.md §0), so this
    arm times the oracle port (-O3 -march=native).  Each step = `cores` independent renders in parallel
    processes, each a slice of config 4 (distinct voices) over the FULL 60 s, i.e. with the same time
    profile as the GPU arm's step (50 s of sounding notes, 10 s silent tail)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["GROOVE_ORACLE_SO"] = native_oracle()
    cores = os.cpu_count() or 1
    voices_each = 4
    frames = int(round(a.seconds * SR))
    note_off = int(frames * 2_400_000 / 2_880_000)
    jobs = [(voices_each, frames, i * voices_each, note_off) for i in range(cores)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for _ in range(min(a.warmup, 1)):
            pool.map(_oracle_render_slice, jobs)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            pool.map(_oracle_render_slice, jobs)
        dt = time.perf_counter() - t0
    total = cores * voices_each * frames * a.steps
    value = total / dt
    sample = (f"{cores} parallel processes x ({voices_each} config-4 voices x {a.seconds:g} s, the full render length) per "
              "step; oracle port built -O3 -march=native -ffp-contract=off (reference is not buildable here)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": min(a.warmup, 1), "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config-4: 4096-voice Welsh cello subtractive synth, 60 s at 48 kHz stereo (bounded sample "
                               "per step: a slice of the voices over the full length)",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "realtime_factor": value / (4096 * SR),
    }), flush=True)


# ------------------------------------------------------------------------------ GPU leg ---
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(gpu_index)],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, power = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.tmp.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


STAT_SUMS = ("render_ms", "voice_kernel_ms", "fx_kernel_ms", "kernel_launches", "voice_kernel_launches",
             "rest_kernel_ms", "rest_kernel_launches", "rest_voice_samples", "rest_ctas",
             "sweep_kernel_ms", "sweep_kernel_launches", "sweep_voice_samples", "sweep_ctas",
             "solo_kernel_ms", "solo_kernel_launches", "solo_voice_samples", "solo_jobs",
             "fm_kernel_ms", "fm_kernel_launches", "idle_voice_samples", "voice_samples", "h2d_bytes", "d2h_bytes",
             "rest_tp_launches", "rest_vr_launches", "rest_vr16_launches")


class Acc(dict):
    """Sums of engine stats over the timed steps of one leg."""

    def add(self, wall: float, st, reduce_ms: float = 0.0):
        for k in STAT_SUMS:
            self[k] = self.get(k, 0) + getattr(st, k)
        cls = list(st.solo_class_items)
        self["solo_class_items"] = [x + y for x, y in zip(self.get("solo_class_items", [0, 0, 0, 0]), cls)]
        self["wall"] = self.get("wall", 0.0) + wall
        self["reduce_ms"] = self.get("reduce_ms", 0.0) + reduce_ms
        self["steps"] = self.get("steps", 0) + 1
        self.setdefault("walls_ms", []).append(round(wall * 1e3, 3))


def run_ours(a) -> None:
    import torch
    import torch.distributed as dist

    from groove_b200 import Engine, abi, parallel, workloads

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this repo has no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    legs = set() if a.no_legs else {x for x in a.legs.split(",") if x}
    if a.no_time_varying:
        legs.discard("time_varying")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # The bus mixdown at N > 1: the root sums the ranks' buses with one kernel whose loads cross NVLink (CUDA IPC
    # peer buffers, groove_b200.parallel.BusExchange); NCCL's reduce only if the peers cannot be mapped.
    exchange = None
    exchange_note = "single GPU"
    if world > 1:
        try:
            if os.environ.get("GB_BUS_NCCL") == "1":
                raise RuntimeError("GB_BUS_NCCL=1")
            exchange = parallel.BusExchange(local, max(int(round(a.seconds * SR)), workloads.CFG5_FRAMES))
            exchange_note = "peer_sum_kernel on rank 0: P2P loads of the ranks' CUDA IPC exchange buffers over NVLink (no NCCL data-path call)"
        except Exception as ex:   # noqa: BLE001 - any failure to map peers keeps the NCCL path
            exchange_note = f"ncclReduce(f64 sum) onto rank 0 (peer mapping unavailable: {ex})"

    def bus_reduce(eng, frames):
        """Sum the per-rank stereo buses onto rank 0."""
        if world == 1:
            return None
        if exchange is not None:
            return exchange.reduce(eng, frames)
        return parallel.reduce_bus(parallel.device_bus_tensor(eng, local), dst=0)

    class Workload:
        """One benchmark workload: `graph(eng)` builds the (finalized) engine and returns the events to push."""

        def __init__(self, name, frames, max_block, graph, reduce=True, collective=True):
            # collective=False: a workload that ONE rank runs on its own (no barrier, no bus reduce)
            self.name, self.frames, self.max_block, self.graph = name, frames, max_block, graph
            self.reduce = reduce and collective
            self.collective = collective
            self.host_out = np.empty((frames, 2), dtype=np.float64)
            self.pinned_out = (torch.empty((frames, 2), dtype=torch.float64, pin_memory=True)
                               if world > 1 and rank == 0 and reduce else None)

        def step(self, mode: str, block: int = 0, lookahead: int = 0):
            """Engine construction (allocation, plan) is setup and stays outside the timed span; pushing the
            step's events from host memory, the per-chunk uploads and the result download are inside."""
            eng = Engine(SR, device=local, max_block=self.max_block)
            eng.set_timing(True)
            ev = self.graph(eng)
            if lookahead:
                eng.set_lookahead(lookahead)
            flush.zero_()
            if self.collective:
                barrier()
            else:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            reduce_ms = 0.0
            eng.push_events(ev)
            if mode == "device":
                eng.render_device(self.frames)    # returns with the engine's stream drained
                if world > 1 and self.reduce:     # the bus reduce, device-timed on the stream NCCL is enqueued from
                    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    ev0.record()
                    bus_reduce(eng, self.frames)
                    ev1.record()
                    torch.cuda.synchronize()
                    reduce_ms = ev0.elapsed_time(ev1)
            elif mode == "blocks":               # the reference's call pattern: fixed small caller buffers
                import ctypes
                fn, h, base = eng._f("render_block"), eng._h, self.host_out.ctypes.data   # the bare C ABI call per buffer
                got, done = ctypes.c_size_t(), 0
                while done < self.frames:
                    k = min(block, self.frames - done)
                    if fn(h, base + 16 * done, k, ctypes.byref(got)) != 0:
                        raise RuntimeError(eng._f("last_error")(h))
                    done += k
            elif world == 1 or not self.reduce:
                eng.render(self.frames, self.host_out)
            else:
                eng.render_device(self.frames)
                t = bus_reduce(eng, self.frames)
                if rank == 0:
                    self.pinned_out.copy_(t)      # D2H of the reduced bus into pinned host memory
                torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            st = eng.stats()
            eng.close()
            return wall, st, reduce_ms

        def run(self, mode: str, steps: int, warmup: int, block: int = 0, lookahead: int = 0) -> Acc:
            sync = barrier if self.collective else torch.cuda.synchronize
            for _ in range(warmup):
                self.step(mode, block, lookahead)
            sync()
            acc = Acc()
            for _ in range(steps):
                acc.add(*self.step(mode, block, lookahead))
            sync()
            return acc

    def max_over_ranks(*vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    # ---- workloads -------------------------------------------------------------------------------
    frames4 = int(round(a.seconds * SR))
    off4 = int(frames4 * 2_400_000 / 2_880_000)

    def cfg4_workload(name, cfg, params=None, reduce=True, collective=True):
        def graph(eng):
            return workloads.cfg4_events(cfg, workloads.build_cfg4_graph(eng, cfg, params))
        return Workload(name, cfg.frames, a.max_block, graph, reduce, collective)

    weak_cfg = parallel.shard_cfg4(workloads.Cfg4(total_voices=a.voices, frames=frames4, note_off_base=off4,
                                                  groups=a.groups or min(128, a.voices)), rank, world, weak=True)
    cfg5_variants = None

    def cfg5_workload():
        nonlocal cfg5_variants
        if cfg5_variants is None:   # the draws are deterministic in the variant index: made once, outside every span
            cfg5_variants = [workloads.cfg5_variant(rank * a.variants + i) for i in range(a.variants)]

        def graph(eng):
            _, _, ev = workloads.build_cfg5(eng, a.variants, first=rank * a.variants, variants=cfg5_variants, push=False)
            return ev
        return Workload("cfg5", workloads.CFG5_FRAMES, workloads.CFG5_FRAMES, graph)

    headline5 = a.workload == "cfg5"
    main = cfg5_workload() if headline5 else cfg4_workload("cfg4", weak_cfg)
    per_rank_vs = (a.variants * workloads.CFG5_FRAMES) if headline5 else weak_cfg.voice_samples

    # ---- headline: warm-up (both modes), then the timed steps -------------------------------------
    W = max(a.warmup, 1)
    main.run("device", 0, W)
    main.run("e2e", 0, W)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    t_region = time.perf_counter()
    dev = main.run("device", a.steps, 0)
    region_s = time.perf_counter() - t_region
    e2e = main.run("e2e", a.steps, 0)
    clocks = sampler.stop() if sampler else None

    dev_ms, e2e_wall = max_over_ranks(dev["render_ms"] + dev["reduce_ms"], e2e["wall"])
    step_s = dev_ms * 1e-3 / a.steps
    total_vs = per_rank_vs * world
    value = total_vs / step_s
    e2e_value = total_vs / (e2e_wall / a.steps)

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json" if peaks else "fallback 6650 (B200_PROFILING.md)"
    fp64_peak = fp32_peak = None
    if rank == 0:
        eng = Engine(SR, device=local)
        fp64_peak = eng.measure_fma_peak(True)
        fp32_peak = eng.measure_fma_peak(False)
        eng.close()
    WV, WF = workloads.W_VOICE_FLOP, workloads.W_FM_FLOP

    def kernel_roofline(name, ms, launches, vs, flop, extra=None):
        """Roofline entry of one kernel family from the engine's per-launch CUDA events."""
        if not launches or ms <= 0:
            return None
        launch_s = ms * 1e-3 / launches
        ach = flop * (vs / launches) / launch_s / 1e12
        r = {"bound": "fp64", "kernel": name, "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
             "frac": ach / fp64_peak if fp64_peak else None, "launch_ms": launch_s * 1e3,
             "voice_samples_per_launch": vs / launches, "algorithmic_flop_per_voice_sample": flop, "traffic": None}
        if extra:
            r.update(extra)
        return r

    def cfg5_rooflines(acc, steps, ms_step):
        wvs = acc["solo_voice_samples"] + 0     # non-idle Welsh (voice, sub-chunk) items x sub-chunk frames
        n_w = a.variants // 2 + a.variants % 2
        n_f = a.variants // 2
        out = {
            "welsh": kernel_roofline("welsh_solo_kernel<8> (job list: resting / sweeping / general items)",
                                     acc["solo_kernel_ms"], acc["solo_kernel_launches"],
                                     n_w * workloads.CFG5_FRAMES * steps, WV,
                                     {"sounding_voice_samples_per_launch": wvs / max(acc["solo_kernel_launches"], 1),
                                      "items_by_class": dict(zip(("resting", "sweeping", "general", "exact"),
                                                                 [int(x // steps) for x in acc["solo_class_items"]])),
                                      "jobs_per_launch": acc["solo_jobs"] / max(acc["solo_kernel_launches"], 1)}),
            "fm": kernel_roofline("fm_kernel<8>", acc["fm_kernel_ms"], acc["fm_kernel_launches"],
                                  n_f * workloads.CFG5_FRAMES * steps, WF),
        }
        step_flop = (n_w * WV + n_f * WF) * workloads.CFG5_FRAMES
        out["step"] = {"bound": "fp64", "achieved": step_flop / (ms_step * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                       "frac": step_flop / (ms_step * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                       "what": "whole step: (150 x subtractive + 67 x FM) FLOP x 96 000 frames over the step's device time; "
                               "the step also writes and re-reads 16 B per frame per variant (per-variant buffers + bus sum)",
                       "hbm_bytes_per_step": 2 * 16.0 * a.variants * workloads.CFG5_FRAMES}
        return out

    out = None
    if rank == 0:
        steps = a.steps
        if headline5:
            roof = cfg5_rooflines(dev, steps, step_s * 1e3)
            roofline = dict(roof["welsh"] or {}, fm=roof["fm"], step=roof["step"])
        else:
            # Dominant kernel of config 4: welsh_rest_kernel (resting voices); its launches, CTAs and the
            # voice-samples they covered are counted by the engine.
            use_rest = dev["rest_kernel_launches"] > 0 and dev["rest_kernel_ms"] > 0.5 * dev["voice_kernel_ms"]
            if use_rest:
                k_ms, k_l, k_vs, k_ctas = dev["rest_kernel_ms"], dev["rest_kernel_launches"], dev["rest_voice_samples"], dev["rest_ctas"]
                # which resting kernel the launches were: instrument CTAs, voice ranges (all-resting chunks of an
                # engine whose CTAs then cover every SM evenly) or the time-parallel variant (small shards)
                n_vr, n_vr16, n_tp = dev["rest_vr_launches"], dev["rest_vr16_launches"], dev["rest_tp_launches"]
                k_name = ("welsh_rest_vr16_kernel<8,lfo,flat>" if 2 * n_vr16 > k_l else
                          "welsh_rest_vr_kernel<8,lfo,flat>" if 2 * n_vr > k_l else
                          "welsh_rest_tp_kernel<8,lfo,flat>" if 2 * n_tp > k_l else "welsh_rest_kernel<8,lfo,flat>")
                k_name += (f" ({int(n_vr // steps)} voice-range [{int(n_vr16 // steps)} with 16 frames per lane], "
                           f"{int(n_tp // steps)} time-parallel of {int(k_l // steps)} resting launches per step)")
            else:
                k_ms, k_l, k_vs, k_ctas = dev["voice_kernel_ms"], dev["voice_kernel_launches"], per_rank_vs * steps, 0
                k_name = "welsh_kernel<8,2>"
            # executed-instruction facts of the dominant kernel: from the committed ncu --set full capture
            prof = {}
            facts = ("r2f_rest_vr16_kernel_facts.json" if use_rest and 2 * dev["rest_vr16_launches"] > k_l
                     else "r2f_rest_vr_kernel_facts.json" if use_rest and 2 * dev["rest_vr_launches"] > k_l
                     else "r2_rest_kernel_facts.json")
            try:
                with open(os.path.join(ROOT, "profiles", facts)) as f:
                    prof = json.load(f)
            except (OSError, ValueError):
                pass
            vs_l = k_vs / max(k_l, 1)
            frames_l = vs_l / a.voices
            ctas_l = k_ctas / max(k_l, 1)
            extra = {
                "ctas_per_launch": ctas_l,
                "algorithmic_bytes_per_launch": 16.0 * frames_l * ctas_l if ctas_l else None,
                "traffic": (prof["dram_bytes_per_launch"] * vs_l / prof["voice_samples_per_launch"]
                            if "dram_bytes_per_launch" in prof else None),
                "traffic_source": "profiles/" + facts + " (ncu --set full dram__bytes_read+write of the same "
                                  "kernel, scaled to this run's voice-samples per launch; not measured live)",
                "peak_source": "FP64 FMA microbenchmark measured live on this GPU (gb_measure_fma_peak); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "kernel_share_of_step": k_ms * 1e-3 / steps / step_s if world == 1 else None,
                "all_voice_kernels_share_of_step": dev["voice_kernel_ms"] * 1e-3 / steps / step_s if world == 1 else None,
                "note": "achieved counts the ALGORITHMIC 150 FLOP per voice-sample of the reference's per-frame loop "
                        "(SURVEY.md 8d).  While a voice rests its cutoff does not move, so this kernel takes the "
                        "coefficient sets and scan maps from per-instrument tables instead of re-deriving them per "
                        "frame: it executes fewer FP64 operations than the algorithmic count — see `executed` (what "
                        "the pipe actually does) and legs.time_varying (the cutoff moving on every frame).",
                "fp32_peak_tflops": fp32_peak,
                "hbm": {"peak_gbs": hbm_peak, "peak_source": hbm_src},
            }
            if use_rest and "fp64_flop_per_voice_sample" in prof:
                ex = prof["fp64_flop_per_voice_sample"] * vs_l / (k_ms * 1e-3 / k_l) / 1e12
                extra["executed"] = {"fp64_flop_per_voice_sample": prof["fp64_flop_per_voice_sample"], "tflops": ex,
                                     "frac": ex / fp64_peak, "fp64_pipe_active_pct": prof.get("fp64_pipe_active_pct"),
                                     "issue_active_pct": prof.get("issue_active_pct"),
                                     "what": "FP64 FLOP the kernel executes per voice-sample (ncu opcode mix: 2 per DFMA, "
                                             "1 per DMUL / DADD) over the same launch time"}
            roofline = kernel_roofline(k_name, k_ms, k_l, k_vs, WV, extra)
        sounding = 1.0 - dev["idle_voice_samples"] / max(dev["voice_samples"], 1)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": W,
            "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"config-5 recipe: {a.variants} one-shot FM/subtractive patch variants per GPU, 2 s at 48 kHz" if headline5
                            else "config-4: 4096-voice Welsh-cookbook cello subtractive synth (dual osc + LFO + 2 ADSR + "
                            "per-frame 24 dB LPF), 60 s at 48 kHz stereo" if (a.voices, a.seconds) == (4096, 60.0)
                            else f"config-4 recipe scaled: {a.voices} voices x {a.seconds:g} s at 48 kHz stereo",
                "voices_per_gpu": a.variants if headline5 else a.voices, "frames": main.frames,
                "voice_samples_per_step": total_vs,
                "coefficients": "per-instrument tables while a voice rests; otherwise quadratic through exact knots "
                                "(every 8 frames in welsh_sweep_kernel, every 4 in welsh_kernel) when the cutoff moves <= "
                                + os.environ.get("GB_KNOT_MAX_RATE", "1e-5") + "/frame, else exact per frame",
                "max_block": main.max_block, "parallelism": f"voices sharded over {world} GPU(s); one bus exchange: " + exchange_note,
                "l2": "256 MiB device memset between steps (L2 flush); fresh engine per step",
            },
            "realtime_factor": value / ((a.variants if headline5 else a.voices) * world * SR),
            "sounding_fraction": sounding,
            "sounding_voice_samples_per_sec": value * sounding,
            "gpu_launches": int(dev["kernel_launches"] // a.steps),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_wall / a.steps * 1e3,
                    "ms_steps_rank0": e2e["walls_ms"],   # wall clock of each timed step (host effects show up here)
                    "h2d_bytes_per_step": int(e2e["h2d_bytes"] // a.steps),
                    "d2h_bytes_per_step": int(e2e["d2h_bytes"] // a.steps) if world == 1 else main.frames * 16},
            "roofline": roofline,
            "clocks": clocks,
            "timed_region_s": region_s,
            "legs": {},
        }
        if not headline5 and world == 1 and dev["fx_kernel_ms"] > 0:
            # the mixdown side of the step (bus-level partial sums + the mixer), HBM-bound
            n_part = dev["rest_ctas"] / max(dev["rest_kernel_launches"], 1) if dev["rest_kernel_launches"] else 0
            mix_bytes = (n_part + 1) * 16.0 * main.frames
            ms = dev["fx_kernel_ms"] / a.steps
            overlapped = os.environ.get("GB_OVERLAP", "1") != "0"
            out["mixdown"] = {"kernel": "sum_table_kernel (+ pointwise_kernel)", "bound": "hbm", "bytes_per_step": mix_bytes,
                              "ms_per_step": ms, "achieved": None if overlapped else mix_bytes / (ms * 1e-3) / 1e9,
                              "peak": hbm_peak, "unit": "GB/s",
                              "frac": None if overlapped else mix_bytes / (ms * 1e-3) / 1e9 / hbm_peak,
                              "bus_bytes": 16.0 * main.frames, "overlapped": overlapped,
                              "note": ("the table sum of chunk i runs on the engine's mix stream WHILE chunk i+1's voice kernels "
                                       "run on the voice stream (partial buffers alternate by chunk parity): ms_per_step is the "
                                       "sum kernels' elapsed time on the SM slots the voice kernels leave free, not time added "
                                       "to the step; alone (GB_OVERLAP=0) the same sums take 2.35 ms at 5.05 TB/s = 0.77 of the "
                                       "HBM peak (profiles/r2_overlap_ab.txt)") if overlapped else None}

    # ---- secondary legs --------------------------------------------------------------------------
    def put(name, d):
        if rank == 0 and d is not None:
            out["legs"][name] = d

    if "cfg5" in legs and not headline5:
        w5 = cfg5_workload()
        acc = w5.run("device", 3, 1)
        acc_e = w5.run("e2e", 2, 1)
        ms5, wall5 = max_over_ranks(acc["render_ms"] + acc["reduce_ms"], acc_e["wall"])
        ms5 /= 3
        vs5 = a.variants * workloads.CFG5_FRAMES * world
        if rank == 0:
            roof = cfg5_rooflines(acc, 3, ms5)
            put("cfg5", {
                "what": f"BASELINE config 5: {a.variants} one-shot FM / subtractive patch variants per GPU "
                        f"({a.variants * world} in all; 65 536 at 8 GPUs), 2 s at 48 kHz, per-variant stereo buffers in HBM + "
                        "the summed bus (NCCL f64 reduce onto rank 0 at N > 1)",
                "value": vs5 / (ms5 * 1e-3), "unit": UNIT, "ms_per_step": ms5, "scaling": "weak",
                "e2e": {"value": vs5 / (wall5 / 2), "unit": UNIT, "ms_per_step": wall5 / 2 * 1e3},
                "gpu_launches": int(acc["kernel_launches"] // 3),
                "roofline": roof["step"], "kernels": {"welsh": roof["welsh"], "fm": roof["fm"]},
                "mixdown_ms": acc["fx_kernel_ms"] / 3,
            })

    if "strong" in legs and not headline5 and world > 1:
        whole = workloads.Cfg4(total_voices=a.voices, frames=frames4, note_off_base=off4, groups=a.groups or min(128, a.voices))
        ws = cfg4_workload("strong", parallel.shard_cfg4(whole, rank, world, weak=False))
        acc = ws.run("device", 3, 1)
        (ms_n,) = max_over_ranks(acc["render_ms"] + acc["reduce_ms"])
        ms_n /= 3
        # the single-GPU time of the same total work, in the same run: rank 0 alone renders all the voices
        ms_1 = 0.0
        if rank == 0:
            w1 = cfg4_workload("strong-n1", whole, reduce=False, collective=False)
            for _ in range(2):
                _, st, _ = w1.step("device")
            ms_1 = st.render_ms
        barrier()
        if rank == 0:
            put("strong", {
                "what": f"config 4 with its {a.voices} voices split over {world} GPUs ({a.voices // world} per GPU, SURVEY.md 8(e)), "
                        "one bus exchange (config.parallelism); efficiency against rank 0 rendering all the voices alone in this run",
                "value": whole.voice_samples / (ms_n * 1e-3), "unit": UNIT, "ms_per_step": ms_n, "scaling": "strong",
                "ms_per_step_1gpu": ms_1, "efficiency_vs_n1": ms_1 / (world * ms_n),
                "reduce_ms": acc["reduce_ms"] / 3, "voice_kernel_ms": acc["voice_kernel_ms"] / 3, "mix_ms": acc["fx_kernel_ms"] / 3,
            })

    if world == 1 and not headline5:
        from dataclasses import replace
        if "time_varying" in legs:
            tv_frames = min(frames4, 12 * 48000)
            tv_cfg = replace(weak_cfg, frames=tv_frames, note_off_base=int(tv_frames * 2_400_000 / 2_880_000), filter_decay=120.0)
            acc = cfg4_workload("tv", tv_cfg).run("device", 2, 1)
            ms = acc["render_ms"] / 2
            if acc["sweep_kernel_launches"]:
                roof = kernel_roofline("welsh_sweep_kernel<8,lfo,flat>", acc["sweep_kernel_ms"], acc["sweep_kernel_launches"],
                                       acc["sweep_voice_samples"], WV,
                                       {"kernel_share_of_step": acc["sweep_kernel_ms"] / 2 / ms})
            else:
                roof = kernel_roofline("welsh_kernel<8,2>", acc["voice_kernel_ms"], acc["voice_kernel_launches"],
                                       tv_cfg.voice_samples * 2, WV)
            put("time_varying", {
                "what": f"config-4 recipe, {tv_frames / SR:g} s, filter-envelope decay stretched to 120 s: the cutoff moves on "
                        "every frame of every note, no voice rests (per-frame coefficient sets from exact knots: "
                        "welsh_sweep_kernel for chunks inside one envelope stage, welsh_kernel for the others)",
                "value": tv_cfg.voice_samples / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "roofline": roof})
        if "general" in legs:
            g_frames = min(frames4, 12 * 48000)
            g_cfg = replace(weak_cfg, frames=g_frames, note_off_base=int(g_frames * 2_400_000 / 2_880_000))
            acc = cfg4_workload("general", g_cfg, params=workloads.piano_params).run("device", 2, 1)
            ms = acc["render_ms"] / 2
            put("general", {
                "what": f"config 4's score ({a.voices} voices, {g_frames / SR:g} s) on the Welsh `piano` patch: hard sync "
                        "(settings/src/patches.rs:122), sawtooth + square — the paths real patches take outside the "
                        "piecewise-linear / no-sync specialisations",
                "value": g_cfg.voice_samples / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                "roofline": kernel_roofline("welsh voice kernels (all)", acc["voice_kernel_ms"], acc["voice_kernel_launches"],
                                            g_cfg.voice_samples * 2, WV),
                "rest_launches": acc["rest_kernel_launches"] // 2, "sweep_launches": acc["sweep_kernel_launches"] // 2})
        if "stream_fx" in legs and hasattr(workloads, "build_fx_chains"):
            n_chains, fx_frames = 1024, 1 << 16

            def graph(eng):
                return workloads.build_fx_chains(eng, n_chains)
            wfx = Workload("stream_fx", fx_frames, fx_frames, graph)
            acc = wfx.run("device", 3, 1)
            ms = acc["fx_kernel_ms"] / 3
            by = 32.0 * fx_frames * n_chains
            put("stream_fx", {
                "what": f"{n_chains} copies of config 1's effect chain (24 dB low-pass with a stepped cutoff trip -> gain) over "
                        f"{fx_frames} frames: ONE launch for all the chains (the gain runs as the filter's post-op), then the "
                        "mixer's table sum; counted 16 B read + 16 B written per frame per chain (SURVEY.md 8(d)) over the "
                        "effect kernels' CUDA-event time (the mixer's own 16 B read per chain is not counted)",
                "value": n_chains * fx_frames / (ms * 1e-3), "unit": "chain-frames/s", "ms_per_step": ms,
                "render_ms": acc["render_ms"] / 3,
                "roofline": {"bound": "hbm", "kernel": "lp24_batch_kernel<2> (gain fused as post-op) + sum_table_kernel", "achieved": by / (ms * 1e-3) / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "frac": by / (ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                             "algorithmic_bytes_per_launch": by, "traffic": None}})
        if "small_blocks" in legs:
            sb_frames = min(frames4, 6 * 48000)
            sb_cfg = replace(weak_cfg, frames=sb_frames, note_off_base=int(sb_frames * 2_400_000 / 2_880_000))
            wsb = cfg4_workload("small_blocks", sb_cfg)
            big = wsb.run("e2e", 2, 1)
            small = wsb.run("blocks", 2, 1, block=64, lookahead=a.max_block)
            plain = wsb.run("blocks", 1, 0, block=64)
            put("small_blocks", {
                "what": f"config-4 recipe, {sb_frames / SR:g} s, rendered through gb_render_block in 64-frame caller buffers "
                        "(the reference's own call size: orchestrator.rs:1696, audio_panel.rs:69) with "
                        f"gb_set_lookahead({a.max_block}), against one big call; `without_lookahead` = the same calls "
                        "each going to the device",
                "value": sb_cfg.voice_samples / (small["wall"] / 2), "unit": UNIT, "ms_per_step": small["wall"] / 2 * 1e3,
                "big_buffer_ms_per_step": big["wall"] / 2 * 1e3, "fraction_of_big_buffer_e2e": big["wall"] / small["wall"],
                "calls_per_step": -(-sb_frames // 64), "gpu_launches": int(small["kernel_launches"] // 2),
                "without_lookahead": {"ms_per_step": plain["wall"] * 1e3, "gpu_launches": int(plain["kernel_launches"]),
                                      "fraction_of_big_buffer_e2e": big["wall"] / 2 / plain["wall"]}})

    if rank == 0:
        if not a.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_subprocess(a.cpu_sample_voices, a.cpu_sample_seconds)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    import faulthandler
    faulthandler.dump_traceback_later(1500, exit=True)   # a hung collective must not hang the caller: trace and exit
    a = parse_args()
    if a._cpu_child:
        cpu_child(a._cpu_child)
    elif a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
