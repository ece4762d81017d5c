"""Multi-rank path on CPU: world_size-2 gloo processes each solve THEIR shard of a batch (kernel bodies through the
host emulation, since there is no GPU here), gather, and the result must equal the single-process solve of the whole
batch -- robots are independent, so sharding may not change a single bit."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q_out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from dyros_robot_controller_b200.shard import gather_batch, shard_range
    from tests.conftest import LINK, SRDF, URDF, workload
    from tests.emu import Emu
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        emu = Emu(URDF, SRDF)
        ma = emu.model_arrays()

        class M:
            q_lo, q_hi, v_lim = ma["q_lo"], ma["q_hi"], ma["v_lim"]
        q, qd, q_t, xd = workload(M, total, 77)
        f = emu.frame_id(LINK)
        x_t = emu.update_and_get(q_t, qd, f)["pose"]
        lo, hi = shard_range(total, world, rank)
        r = emu.cycle(1, q[lo:hi], qd[lo:hi], x_t[lo:hi], xd[lo:hi], f)
        out = gather_batch(torch.from_numpy(r["out"]), total)
        its = gather_batch(torch.from_numpy(r["iters"]), total)
        if rank == 0:
            q_out.put((out.numpy(), its.numpy()))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_tile_the_batch():
    from dyros_robot_controller_b200.shard import shard_range, shard_sizes
    for total in (0, 1, 7, 65536, 1048577):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(total, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sz = shard_sizes(total, world)
            assert max(sz) - min(sz) <= 1 and sum(sz) == total
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_two_rank_gloo_equals_single_process():
    import torch.multiprocessing as mp
    from tests.conftest import LINK, SRDF, URDF, workload
    from tests.emu import Emu
    total, world = 203, 2   # odd on purpose: shards of 101 and 102 robots
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    out, its = q_out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    emu = Emu(URDF, SRDF)
    ma = emu.model_arrays()

    class M:
        q_lo, q_hi, v_lim = ma["q_lo"], ma["q_hi"], ma["v_lim"]
    q, qd, q_t, xd = workload(M, total, 77)
    f = emu.frame_id(LINK)
    x_t = emu.update_and_get(q_t, qd, f)["pose"]
    ref = emu.cycle(1, q, qd, x_t, xd, f)
    assert np.array_equal(its, ref["iters"])
    assert np.array_equal(out, ref["out"])
