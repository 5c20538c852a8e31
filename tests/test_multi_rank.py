"""world_size-2 gloo test (CPU) of the N > 1 path: shard config 4's voices over ranks, render each
shard (with the CPU oracle standing in for the engine — the sharding and the bus reduce are host
logic), reduce the stereo buses onto rank 0, compare with the unsharded render."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, frames, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from groove_b200 import parallel, workloads
    from tests.oracle_binding import OracleEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = workloads.cfg4_slice(64, frames)
    mine = parallel.shard_cfg4(cfg, rank, world, weak=False)
    o = OracleEngine(48000.0)
    workloads.build_cfg4(o, mine)
    bus = o.render(frames)
    total = parallel.reduce_bus_numpy(bus, dst=0)
    if rank == 0:
        q.put(total.copy())
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_shard_and_bus_reduce_matches_single_rank():
    from groove_b200 import workloads
    from tests.oracle_binding import OracleEngine
    frames = 6000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = OracleEngine(48000.0)
    workloads.build_cfg4(o, workloads.cfg4_slice(64, frames))
    ref = o.render(frames)
    assert np.abs(ref).max() > 1e-4
    assert np.abs(total - ref).max() < 1e-15 + 1e-12 * np.abs(ref).max()


def test_shard_helpers():
    from groove_b200 import parallel, workloads
    cfg = workloads.Cfg4()
    w = [parallel.shard_cfg4(cfg, r, 8) for r in range(8)]
    assert [c.voice_offset for c in w] == [4096 * r for r in range(8)] and all(c.total_voices == 4096 for c in w)
    s = [parallel.shard_cfg4(cfg, r, 8, weak=False) for r in range(8)]
    assert sum(c.total_voices for c in s) == 4096 and [c.group_first for c in s] == [16 * r for r in range(8)]
    assert all(c.groups == 16 and c.group_total == 128 and c.voice_offset == 0 for c in s)
    # the shards' voices partition the whole ensemble's (same events, instrument by instrument)
    whole = workloads.cfg4_events(cfg, list(range(100, 228)))
    parts = [workloads.cfg4_events(c, list(range(100 + c.group_first, 100 + c.group_first + 16))) for c in s]
    key = lambda ev: sorted(map(tuple, ev.tolist()))
    assert key(whole) == key(np.concatenate(parts))
    odd = [parallel.shard_cfg4(workloads.Cfg4(total_voices=96, groups=3), r, 2, weak=False) for r in range(2)]
    assert [c.voice_offset for c in odd] == [0, 48] and all(c.total_voices == 48 for c in odd)
