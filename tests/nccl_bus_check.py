"""Run under torchrun with 2+ ranks (tests/test_gpu_parity.py::test_two_gpu_bus_reduce_...): every rank renders
its strong-scaling shard of a config-4 slice on its own GPU, the buses are reduced onto rank 0 over NCCL, and
rank 0 writes the comparison against a single-GPU render of the union and against the CPU oracle."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from groove_b200 import Engine, parallel, workloads  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    frames = 3 * 65536 + 1234
    whole = workloads.Cfg4(total_voices=512, frames=frames, note_off_base=2 * 65536 + 999, groups=32)
    mine = parallel.shard_cfg4(whole, rank, world, weak=False)
    eng = Engine(48000.0, device=local)
    workloads.build_cfg4(eng, mine)
    eng.render_device(frames)
    # the product path: the root sums the ranks' exchange buffers over NVLink peer memory (peer_sum_kernel) ...
    p2p, p2p_err = None, None
    try:
        xch = parallel.BusExchange(local, frames)
        t = xch.reduce(eng, frames)
        torch.cuda.synchronize()
        if rank == 0:
            p2p = t.cpu().numpy().copy()
        dist.barrier()
        xch.close()
    except Exception as ex:   # noqa: BLE001
        p2p_err = str(ex)
    # ... and the NCCL reduce of the same buses (the fallback when peers cannot be mapped)
    bus = parallel.device_bus_tensor(eng, local)
    parallel.reduce_bus(bus, dst=0)
    torch.cuda.synchronize()
    if rank == 0:
        mixed = bus.cpu().numpy().copy()
    eng.close()
    if rank == 0:
        single = Engine(48000.0, device=local)
        workloads.build_cfg4(single, whole)
        alone = single.render(frames)
        single.close()
        from tests.oracle_binding import OracleEngine
        o = OracleEngine(48000.0)
        workloads.build_cfg4(o, whole)
        ref = o.render(frames)
        with open(sys.argv[1], "w") as f:
            json.dump({"world": world, "peak": float(np.abs(alone).max()),
                       "max_abs_vs_single_gpu": float(np.abs(mixed - alone).max()),
                       "max_abs_vs_oracle": float(np.abs(mixed - ref).max()),
                       "p2p_error": p2p_err,
                       "p2p_max_abs_vs_nccl": None if p2p is None else float(np.abs(p2p - mixed).max()),
                       "p2p_max_abs_vs_oracle": None if p2p is None else float(np.abs(p2p - ref).max())}, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
