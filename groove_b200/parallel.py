"""Multi-GPU plumbing: shard independent voices/instruments over ranks, one f64 bus reduce.

The path shards with a single exchange step: every rank renders its shard of the
voices into a stereo bus (`double2[frames]`) and the buses are summed onto rank 0 with one
`reduce(SUM)` — NCCL over NVLink on GPUs, gloo in the CPU tests.  No per-block collectives.
"""
from __future__ import annotations

from dataclasses import replace

import numpy as np

from . import workloads


def shard_cfg4(cfg: workloads.Cfg4, rank: int, world: int, weak: bool = True) -> workloads.Cfg4:
    """Rank-local slice of config 4.

    weak scaling (bench.py): every rank renders `cfg.total_voices` voices of a `world`-times larger
    ensemble.  strong scaling: the `cfg.total_voices` voices are split evenly.  When the instruments divide
    evenly they are dealt out WHOLE (rank r renders instruments r G/world .. (r+1) G/world - 1 with all their
    voices: independent tracks), so every rank keeps full voice groups for its CTAs;
    otherwise the voices are cut by index (voice i belongs to rank i // (total/world)) and the groups shrink.
    """
    if weak:
        return replace(cfg, voice_offset=cfg.voice_offset + rank * cfg.total_voices)
    assert cfg.total_voices % world == 0
    per = cfg.total_voices // world
    if cfg.groups % world == 0 and not cfg.group_total:
        g = cfg.groups // world
        return replace(cfg, total_voices=per, groups=g, group_first=rank * g, group_total=cfg.groups)
    groups = min(cfg.groups, per)
    while per % groups:
        groups -= 1
    return replace(cfg, total_voices=per, groups=groups, voice_offset=cfg.voice_offset + rank * per)


def device_bus_tensor(engine, device_index: int):
    """Wrap the engine's device-resident result (f64 L,R interleaved in HBM) as a torch tensor, zero-copy."""
    import torch
    ptr, n = engine.last_device_buffer()

    class _Wrap:
        __cuda_array_interface__ = {"shape": (n, 2), "typestr": "<f8", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Wrap(), device=torch.device("cuda", device_index))


def reduce_bus(bus, dst: int = 0):
    """Sum the per-rank buses onto `dst` in place.  `bus` is a torch tensor (CUDA for NCCL, CPU for gloo)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(bus, dst=dst, op=dist.ReduceOp.SUM)
    return bus


def reduce_bus_numpy(bus: np.ndarray, dst: int = 0) -> np.ndarray:
    """gloo / CPU variant used by the multi-rank tests."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(bus))
    reduce_bus(t, dst)
    return t.numpy()
