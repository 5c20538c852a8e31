#!/usr/bin/env python
"""bench.py — voice-samples/s of the Groove synthesis hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the CPU implementation (oracle) on host cores

A *step* is one full render of BASELINE config 4 — 4096 Welsh `cello` voices, 60 s at 48 kHz stereo
= 1.17965e10 voice-samples — through the engine.  At N > 1 every rank renders its own 4096-voice
shard (weak scaling) and the stereo buses are summed onto rank 0 with one NCCL f64 reduce.

Printed JSON (one line, rank 0):
  value      whole-job voice-samples/s, device-timed (CUDA events on the engine's stream around
             each render call; at N > 1 plus CUDA events around the NCCL bus reduce, max over
             ranks), inputs resident in HBM, result left in HBM;
  e2e        the same metric through the C ABI with host buffers: events pushed from host memory
             and the f64 stereo result copied back to a host buffer inside the timed region;
  roofline   dominant kernel (config 4: welsh_rest_kernel, the resting-voice kernel) against the FP64
             vector pipe, which is what binds this path: achieved = the ALGORITHMIC
             150 FLOP x voice-samples its launches covered / their CUDA-event time; peak = FP64 FMA
             microbenchmark measured live on the same GPU.  The kernel executes fewer FP64 instructions
             than the algorithmic count (the resting cutoff makes the per-frame coefficient work
             redundant: DESIGN.md §3.1a), so the executed count and the ncu pipe utilisation are
             reported beside it;
  time_varying  (N = 1) the same recipe with the filter decay stretched past the note — the cutoff moves
             on every frame, no voice rests: value and roofline of welsh_kernel's moving-cutoff path;
  cpu_baseline  the CPU oracle (reference-structured restatement) on a bounded sample, 1 core.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--voices", type=int, default=4096, help="voices per GPU (config 4: 4096)")
    ap.add_argument("--seconds", type=float, default=60.0, help="audio seconds per step (config 4: 60)")
    ap.add_argument("--max-block", type=int, default=1 << 16)
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg5"],
                    help="cfg4 = BASELINE config 4 (headline); cfg5 = batch of one-shot patch variants")
    ap.add_argument("--variants", type=int, default=8192, help="cfg5: variants per GPU (config 5: 65536 over 8 GPUs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-time-varying", action="store_true", help="skip the secondary always-moving-cutoff leg")
    ap.add_argument("--cpu-sample-voices", type=int, default=256)
    ap.add_argument("--cpu-sample-seconds", type=float, default=4.0)
    return ap.parse_args()


# --------------------------------------------------------------------------- CPU legs ---
def _oracle_render_slice(args):
    """Worker: render `voices` voices x `frames` frames of config 4 on the CPU oracle; returns seconds."""
    voices, frames, voice_offset = args
    from groove_b200 import workloads
    from tests.oracle_binding import OracleEngine
    o = OracleEngine(48000.0)
    n = workloads.build_cfg4(o, workloads.cfg4_slice(voices, frames, voice_offset))
    t = time.perf_counter()
    done = 0
    out = np.empty((4800, 2))
    while done < n:                       # 64-frame-multiple buffers like the reference CLI loop
        k = min(4800, n - done)
        o.render(k, out)
        done += k
    return time.perf_counter() - t


def cpu_baseline_single(voices: int, seconds: float) -> dict:
    frames = int(seconds * 48000)
    dt = _oracle_render_slice((voices, frames, 0))
    return {
        "value": voices * frames / dt, "unit": UNIT, "cores": 1, "kind": "port",
        "sample": f"config-4 voices 0..{voices - 1}, first {seconds:g} s ({voices * frames:.3e} voice-samples), "
                  f"oracle/groove_oracle.cpp per-frame graph walk, {dt:.2f} s wall",
    }


def run_reference_arm(a) -> None:
    """--impl reference: the CPU implementation of the path on all host cores.

    The reference itself cannot be built (no Rust toolchain, DSP source absent: SURVEY.md §0), so this
    arm times the oracle port.  Each step = `cores` independent renders in parallel processes, each a
    bounded slice of config 4 (distinct voices).
    """
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    voices_each, seconds = 32, 3.0
    frames = int(seconds * 48000)
    jobs = [(voices_each, frames, i * voices_each) for i in range(cores)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for _ in range(a.warmup):
            pool.map(_oracle_render_slice, jobs)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            pool.map(_oracle_render_slice, jobs)
        dt = time.perf_counter() - t0
    total = cores * voices_each * frames * a.steps
    value = total / dt
    sample = (f"{cores} parallel processes x ({voices_each} config-4 voices x {seconds:g} s) per step; "
              "oracle port (reference is not buildable here)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config-4: 4096-voice Welsh cello subtractive synth, 48 kHz stereo (bounded sample per step)",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "realtime_factor": value / (4096 * 48000.0),
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


def run_ours(a) -> None:
    import torch
    import torch.distributed as dist

    from groove_b200 import Engine, parallel, workloads

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this repo has no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg5 = a.workload == "cfg5"
    frames = workloads.CFG5_FRAMES if cfg5 else int(round(a.seconds * 48000))
    if cfg5:
        a.max_block = frames          # one chunk: every variant's node buffer is its full 2 s output
    cfg = parallel.shard_cfg4(
        workloads.Cfg4(total_voices=a.voices, frames=frames, note_off_base=int(frames * 2_400_000 / 2_880_000),
                       groups=min(128, a.voices)), rank, world, weak=True)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    host_out = np.empty((frames, 2), dtype=np.float64)
    pinned_out = torch.empty((frames, 2), dtype=torch.float64, pin_memory=True) if world > 1 and rank == 0 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def bus_reduce(eng):
        """Sum the per-rank stereo buses onto rank 0: one NCCL f64 reduce over NVLink."""
        if world == 1:
            return None
        return parallel.reduce_bus(parallel.device_bus_tensor(eng, local), dst=0)

    def one_step(mode: str, step_cfg=None):
        """Build config 4 and render it.  Engine construction (allocation, plan, host-side event list)
        is setup and stays outside the timed span; the per-chunk event upload (H2D) and the result
        download (D2H) happen inside the render call, i.e. inside the e2e span."""
        eng = Engine(48000.0, device=local, max_block=a.max_block)
        eng.set_timing(True)
        if cfg5:
            workloads.build_cfg5(eng, a.variants, first=rank * a.variants)
        else:
            workloads.build_cfg4(eng, step_cfg or cfg)
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        reduce_ms = 0.0
        if mode == "device":
            eng.render_device(frames)        # returns with the engine's stream drained
            if world > 1:                    # the bus reduce, device-timed on the stream NCCL is enqueued from
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                bus_reduce(eng)
                ev1.record()
                torch.cuda.synchronize()
                reduce_ms = ev0.elapsed_time(ev1)
        else:
            if world == 1:
                eng.render(frames, host_out)
            else:
                eng.render_device(frames)
                t = bus_reduce(eng)
                if rank == 0:
                    pinned_out.copy_(t)      # D2H of the reduced bus into pinned host memory
                torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        st = eng.stats()
        eng.close()
        st.reduce_ms = reduce_ms
        return wall, st

    # warm-up (both modes), then the timed steps
    for _ in range(max(a.warmup, 1)):
        one_step("device")
    for _ in range(max(a.warmup, 1)):   # the host-buffer path warms up separately (copy stream, host pages)
        one_step("e2e")
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    dev_ms, kern_ms, launches, vlaunches, wall_dev = 0.0, 0.0, 0, 0, 0.0
    rest_ms, rest_launches, rest_vs, fx_ms = 0.0, 0, 0, 0.0
    t_region = time.perf_counter()
    for _ in range(a.steps):
        wall, st = one_step("device")
        dev_ms += st.render_ms + st.reduce_ms   # device time of the step: render (engine events) + bus reduce
        kern_ms += st.voice_kernel_ms
        launches += st.kernel_launches
        vlaunches += st.voice_kernel_launches
        rest_ms += st.rest_kernel_ms
        rest_launches += st.rest_kernel_launches
        rest_vs += st.rest_voice_samples
        fx_ms += st.fx_kernel_ms
        wall_dev += wall
    barrier()
    region_s = time.perf_counter() - t_region
    e2e_wall, h2d, d2h = 0.0, 0, 0
    e2e_steps_ms = []
    for _ in range(a.steps):
        wall, st = one_step("e2e")
        e2e_wall += wall
        e2e_steps_ms.append(round(wall * 1e3, 3))
        h2d += st.h2d_bytes
        d2h += st.d2h_bytes if world == 1 else (frames * 16 if rank == 0 else 0)
    barrier()
    clocks = sampler.stop() if sampler else None
    # secondary leg (N = 1, config 4 only): the same recipe with the filter-envelope decay stretched past
    # the note, so the cutoff moves on every frame of every note and no voice ever rests — the
    # time-varying path (welsh_kernel's knot-interpolated block) on its own
    tv = None
    if world == 1 and not cfg5 and not a.no_time_varying:
        from dataclasses import replace
        tv_frames = min(frames, 12 * 48000)
        tv_cfg = replace(cfg, frames=tv_frames, note_off_base=int(tv_frames * 2_400_000 / 2_880_000), filter_decay=120.0)
        saved = frames
        frames = tv_frames
        one_step("device", tv_cfg)
        t_ms, t_kern, t_vl, t_sw_ms, t_sw_l, t_sw_vs = 0.0, 0.0, 0, 0.0, 0, 0
        for _ in range(2):
            _, st = one_step("device", tv_cfg)
            t_ms += st.render_ms; t_kern += st.voice_kernel_ms; t_vl += st.voice_kernel_launches
            t_sw_ms += st.sweep_kernel_ms; t_sw_l += st.sweep_kernel_launches; t_sw_vs += st.sweep_voice_samples
        frames = saved
        tv = {"ms": t_ms / 2, "kern_ms": t_kern / 2, "launches": t_vl // 2, "voice_samples": tv_cfg.voice_samples,
              "sounding_voice_samples": a.voices * tv_cfg.note_off_base, "seconds": tv_frames / 48000.0,
              "sweep_ms": t_sw_ms, "sweep_launches": t_sw_l, "sweep_vs": t_sw_vs}

    # max over ranks
    vals = torch.tensor([dev_ms, wall_dev, e2e_wall, kern_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    dev_ms, wall_dev, e2e_wall, kern_ms = [float(x) for x in vals.tolist()]
    # device-timed at every N (max over ranks): CUDA events around the render on the engine's stream plus,
    # at N > 1, CUDA events around the NCCL bus reduce; e2e below is wall clock
    step_s = dev_ms * 1e-3 / a.steps
    per_rank_vs = a.variants * frames if cfg5 else cfg.voice_samples
    total_vs = per_rank_vs * world
    value = total_vs / step_s
    e2e_value = total_vs / (e2e_wall / a.steps)

    if rank == 0:
        eng = Engine(48000.0, device=local)
        fp64_peak = eng.measure_fma_peak(True)
        fp32_peak = eng.measure_fma_peak(False)
        eng.close()
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # Dominant kernel.  Config 4: welsh_rest_kernel (resting voices; 90+ % of the step) — its launches and
        # the voice-samples they covered are counted by the engine.  Otherwise: all voice-kernel launches.
        use_rest = (not cfg5) and rest_launches > 0 and rest_ms > 0.5 * kern_ms
        if use_rest:
            k_ms, k_launches, k_vs = rest_ms, rest_launches, rest_vs
            k_name = "welsh_rest_kernel<8,lfo,flat>"
        else:
            k_ms, k_launches, k_vs = kern_ms, vlaunches, per_rank_vs * a.steps
            k_name = "welsh_kernel<8,2>" + (" + fm_kernel<8>" if cfg5 else "")
        vs_per_launch = k_vs / max(k_launches, 1)
        frames_per_launch = vs_per_launch / (a.variants if cfg5 else a.voices)
        ctas_per_launch = 2 * min(128, a.voices)   # config 4: each 32-voice instrument is split over 2 CTAs
        launch_s = k_ms * 1e-3 / max(k_launches, 1)
        # cfg5: half the variants are FM voices (67 FLOP), launched as a second kernel per chunk
        flop_per_vs = 0.5 * (workloads.W_VOICE_FLOP + workloads.W_FM_FLOP) if cfg5 else workloads.W_VOICE_FLOP
        achieved_tflops = flop_per_vs * vs_per_launch / launch_s / 1e12
        # DRAM traffic and executed-instruction facts of the dominant kernel: from the committed ncu --set full
        # capture (cannot be measured live), traffic scaled to this run's voice-samples per launch
        traffic, prof = None, {}
        try:
            with open(os.path.join(ROOT, "profiles", "r1_welsh_traffic.json")) as f:
                prof = json.load(f)
            if not cfg5:
                traffic = prof["dram_bytes_per_launch"] * vs_per_launch / prof["voice_samples_per_launch"]
        except (OSError, KeyError, ValueError):
            pass
        # algorithmic HBM bytes: 16 B stereo f64 out per frame per CTA partial + voice state in/out
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 1),
            "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"config-5 recipe: {a.variants} one-shot FM/subtractive patch variants per GPU, 2 s at 48 kHz" if cfg5
                            else "config-4: 4096-voice Welsh-cookbook cello subtractive synth (dual osc + LFO + 2 ADSR + "
                            "per-frame 24 dB LPF), 60 s at 48 kHz stereo" if (a.voices, a.seconds) == (4096, 60.0)
                            else f"config-4 recipe scaled: {a.voices} voices x {a.seconds:g} s at 48 kHz stereo",
                "voices_per_gpu": a.voices, "frames": frames, "voice_samples_per_step": total_vs,
                "coefficients": ("exact per frame (GB_KNOT_MAX_RATE=0)" if os.environ.get("GB_KNOT_MAX_RATE", "") in ("0", "0.0")
                                 else "per-instrument tables while a voice rests; otherwise quadratic through exact knots "
                                      "(every 8 frames in welsh_sweep_kernel, every 4 in welsh_kernel) when the cutoff moves <= "
                                      + os.environ.get("GB_KNOT_MAX_RATE", "1e-5") + "/frame, else exact per frame"),
                "max_block": a.max_block, "parallelism": f"voices sharded over {world} GPU(s); one NCCL f64 bus reduce",
                "l2": "256 MiB device memset between steps (L2 flush); fresh engine per step",
            },
            "realtime_factor": value / ((a.variants if cfg5 else a.voices) * world * 48000.0),
            "gpu_launches": int(launches // a.steps),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_wall / a.steps * 1e3,
                    "ms_steps_rank0": e2e_steps_ms,   # wall clock of each timed step (host effects show up here)
                    "h2d_bytes_per_step": int(h2d // a.steps), "d2h_bytes_per_step": int(d2h // a.steps)},
            "roofline": {
                "bound": "fp64", "kernel": k_name, "achieved": achieved_tflops, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": achieved_tflops / fp64_peak, "traffic": traffic,
                "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read+write, profiles/r1_welsh_traffic.json)",
                "algorithmic_bytes_per_launch": 16.0 * frames_per_launch * ctas_per_launch if not cfg5 else None,
                "peak_source": "FP64 FMA microbenchmark measured live on this GPU (gb_measure_fma_peak); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "algorithmic_flop_per_voice_sample": flop_per_vs,
                "voice_samples_per_launch": vs_per_launch, "launch_ms": launch_s * 1e3,
                "kernel_share_of_step": k_ms * 1e-3 / a.steps / step_s if world == 1 else None,
                "all_voice_kernels_share_of_step": kern_ms * 1e-3 / a.steps / step_s if world == 1 else None,
                "note": "achieved counts the ALGORITHMIC 150 FLOP per voice-sample of the reference's per-frame loop "
                        "(SURVEY.md 8d). While a voice rests (envelopes at sustain) its cutoff does not move, so this "
                        "kernel takes the coefficient sets and scan maps from per-instrument tables instead of "
                        "re-deriving them per frame: it executes fewer FP64 instructions than the algorithmic count, "
                        "which is why frac can exceed what the pipe utilisation alone would give; see "
                        "executed_fp64_instr_per_voice_sample / fp64_pipe_active_pct (ncu) and the time_varying leg.",
                "executed": ({"fp64_flop_per_voice_sample": prof["fp64_flop_per_voice_sample"],
                              "tflops": prof["fp64_flop_per_voice_sample"] * vs_per_launch / launch_s / 1e12,
                              "frac": prof["fp64_flop_per_voice_sample"] * vs_per_launch / launch_s / 1e12 / fp64_peak,
                              "what": "FP64 FLOP the kernel actually executes per voice-sample (ncu opcode mix: "
                                      "2 per DFMA, 1 per DMUL / DADD) over the same launch time"}
                             if use_rest and "fp64_flop_per_voice_sample" in prof else None),
                "executed_fp64_instr_per_voice_sample": prof.get("fp64_instr_per_voice_sample"),
                "fp64_pipe_active_pct": prof.get("fp64_pipe_active_pct"),
                "issue_active_pct": prof.get("issue_active_pct"),
                "fp32_peak_tflops": fp32_peak,
                "hbm": {"peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650"},
            },
            "clocks": clocks,
            "timed_region_s": region_s,
        }
        if not cfg5 and world == 1 and fx_ms > 0:
            # the mixdown side of the step: split instruments' partial buffers summed into the mixer bus
            # (sum_table_kernel + the mixer's pointwise pass), HBM-bound: 16 B per partial per frame
            chunks = -(-frames // a.max_block)
            mix_bytes = (ctas_per_launch + 1 + 2) * 16.0 * frames
            out["mixdown"] = {"kernel": "sum_table_kernel + pointwise_kernel", "bound": "hbm",
                              "bytes_per_step": mix_bytes, "ms_per_step": fx_ms / a.steps,
                              "achieved": mix_bytes / (fx_ms / a.steps * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                              "frac": mix_bytes / (fx_ms / a.steps * 1e-3) / 1e9 / hbm_peak, "chunks": chunks}
        if tv:
            if tv["sweep_launches"]:   # dominant kernel of this leg: the sweeping-voice kernel, by its own launches
                tv_launch_s = tv["sweep_ms"] * 1e-3 / tv["sweep_launches"]
                tv_vs_launch = tv["sweep_vs"] / tv["sweep_launches"]
                tv_kernel = "welsh_sweep_kernel<8,lfo,flat>"
            else:
                tv_launch_s = tv["kern_ms"] * 1e-3 / max(tv["launches"], 1)
                tv_vs_launch = tv["voice_samples"] / max(tv["launches"], 1)
                tv_kernel = "welsh_kernel<8,2>"
            tv_ach = workloads.W_VOICE_FLOP * tv_vs_launch / tv_launch_s / 1e12
            out["time_varying"] = {
                "what": f"config-4 recipe, {tv['seconds']:g} s, filter-envelope decay stretched to 120 s: the cutoff moves "
                        "on every frame of every note, no voice rests (per-frame coefficient sets from exact knots: "
                        "welsh_sweep_kernel for chunks inside one envelope stage, welsh_kernel for the others)",
                "value": tv["voice_samples"] / (tv["ms"] * 1e-3), "unit": UNIT, "ms_per_step": tv["ms"],
                "roofline": {"bound": "fp64", "kernel": tv_kernel, "achieved": tv_ach, "peak": fp64_peak,
                             "unit": "TFLOP/s", "frac": tv_ach / fp64_peak, "launch_ms": tv_launch_s * 1e3,
                             "voice_samples_per_launch": tv_vs_launch,
                             "kernel_share_of_step": (tv["sweep_ms"] / 2 if tv["sweep_launches"] else tv["kern_ms"]) / tv["ms"]},
            }
        if not a.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_single(a.cpu_sample_voices, a.cpu_sample_seconds)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
