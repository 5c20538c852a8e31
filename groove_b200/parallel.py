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


class BusExchange:
    """The bus mixdown over NVLink peer memory (include/groove_b200.h, "multi-GPU bus exchange").

    Every rank owns an exchange buffer in its HBM; the CUDA IPC handles are gathered once over
    torch.distributed; `reduce()` has every rank copy its engine's device-resident render into its buffer, the
    ranks meet at a barrier, and the root sums all buffers with ONE kernel whose remote loads cross NVLink
    (peer_sum_kernel) — no NCCL data-path call.  Construction raises if the peers' buffers cannot be mapped (no P2P
    between the GPUs, IPC not permitted in the sandbox): callers then keep `reduce_bus` (NCCL).
    """

    def __init__(self, device_index: int, frames: int, root: int = 0):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import abi, engine
        self._lib = engine.load_library()
        vp = C.c_void_p
        for name, res, args in (
                ("create", C.c_int, [C.c_int32, C.c_size_t, C.POINTER(vp)]), ("destroy", None, [vp]),
                ("export", C.c_int, [vp, vp]), ("open", C.c_int, [vp, vp, C.c_int32, C.c_int32]),
                ("publish", C.c_int, [vp, vp, C.c_size_t, vp]), ("reduce", C.c_int, [vp, C.c_size_t, vp]),
                ("result", C.c_int, [vp, C.POINTER(vp), C.POINTER(C.c_size_t)])):
            fn = getattr(self._lib, "gb_bus_exchange_" + name)
            fn.restype, fn.argtypes = res, args
        self._lib.gb_last_error.restype, self._lib.gb_last_error.argtypes = C.c_char_p, [vp]
        self._err = abi.GrooveError
        self.device, self.frames, self.root = device_index, int(frames), root
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self._h = vp()
        self._check(self._lib.gb_bus_exchange_create(device_index, self.frames, C.byref(self._h)))
        handle = C.create_string_buffer(64)
        self._check(self._lib.gb_bus_exchange_export(self._h, handle))
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).cuda(device_index)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine)
        ok = 1
        if self.rank == root:   # only the root reads peer memory
            blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in gathered)
            ok = 1 if self._lib.gb_bus_exchange_open(self._h, blob, self.world, self.rank) == 0 else 0
        flag = torch.tensor([ok], dtype=torch.int32, device=torch.device("cuda", device_index))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            msg = self._lib.gb_last_error(None)
            self.close()
            raise RuntimeError("bus exchange unavailable: " + ((msg or b"").decode() or "peer mapping failed on the root"))

    def _check(self, rc: int):
        if rc != 0:
            msg = self._lib.gb_last_error(None)
            raise self._err(rc, (msg or b"").decode())

    def reduce(self, eng, frames: int):
        """Publish this rank's render, meet, and (root) sum the buses; everything is enqueued on torch's current
        stream, so CUDA events recorded around this call time it on the device.  Returns the summed bus (root)
        as a torch tensor view of the exchange's result buffer, else None."""
        import ctypes as C

        import torch
        import torch.distributed as dist
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = self._lib.gb_bus_exchange_publish(self._h, eng._h, frames, stream)
        if rc != 0:
            msg = self._lib.gb_last_error(eng._h)
            raise self._err(rc, (msg or b"").decode())
        dist.barrier()          # every rank's copy is ordered before the root's loads (NCCL barrier on the same stream order)
        if self.rank != self.root:
            return None
        self._check(self._lib.gb_bus_exchange_reduce(self._h, frames, stream))
        ptr, n = C.c_void_p(), C.c_size_t()
        self._check(self._lib.gb_bus_exchange_result(self._h, C.byref(ptr), C.byref(n)))

        class _Wrap:
            __cuda_array_interface__ = {"shape": (int(frames), 2), "typestr": "<f8", "data": (ptr.value, False), "version": 2}
        return torch.as_tensor(_Wrap(), device=torch.device("cuda", self.device))

    def close(self):
        if self._h:
            self._lib.gb_bus_exchange_destroy(self._h)
            self._h = None
