"""The N>1 path on CPU: two processes over gloo shard a batch by instance, each solves its block, rank 0 gathers.
The solver stand-in is the host emulation of the device code (tests/emu -- test infrastructure); what is under
test is the host-side logic that the GPU bench uses: shard ranges, the global instance offset that keys the
perturbStep generator, the gather in rank order and the max/sum reductions."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from lcqpow_b200 import sharding


def test_shard_ranges_cover_the_batch():
    for total in (0, 1, 7, 37, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharding.shard_counts(total, world)
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _worker(rank, world, port, total, q):
    import ctypes as C
    import sys
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here)):
        if p not in sys.path:
            sys.path.insert(0, p)
    from emu import EmuLib
    from lcqpow_b200 import problems as P
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        emu = EmuLib()
        emu.lib.lcqp_emu_set_instance_offset.argtypes = [C.c_ulonglong]
        pb = P.dense_random_batch(total, n=12, nComp=5, nC=3, seed0=71000)
        lo, hi = sharding.shard_range(total, rank, world)
        emu.lib.lcqp_emu_set_instance_offset(lo)
        s = emu.solve_batch(pb.slice(lo, hi), emu.default_options(perturbStep=1, perturb_seed=5))
        x = sharding.gather_rows(s.x, total)
        it = sharding.gather_rows(s.res["iterTotal"].astype(np.int64), total)
        tmax = sharding.reduce_scalars([float(rank + 1)], "max")[0]
        nsum = sharding.reduce_scalars([float(hi - lo)], "sum")[0]
        if rank == 0:
            emu.lib.lcqp_emu_set_instance_offset(0)
            full = emu.solve_batch(pb, emu.default_options(perturbStep=1, perturb_seed=5))
            q.put((bool(np.array_equal(x, full.x)), bool(np.array_equal(it, full.res["iterTotal"])), tmax, nsum,
                   int((full.res["ret"] == 0).sum())))
        else:
            assert x is None and it is None
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_is_bit_identical_to_one_rank():
    world, total = 2, 9
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    same_x, same_it, tmax, nsum, solved = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same_x and same_it
    assert tmax == 2.0 and nsum == total
    assert solved >= total - 1


def test_bench_family_shard_is_generated_without_the_rest():
    """bench.py gives every rank only its own shard of the C2 family (8 x 2^20 instances would be 27 GB of g / x0 per rank):
    circle_batch_fast(n, lo, hi) must be instances [lo, hi) of circle_batch_fast(n), instance 0 the shipped one."""
    import numpy as np
    from lcqpow_b200 import problems as P
    from lcqpow_b200 import sharding
    n, world = 4096, 4
    full = P.circle_batch_fast(n)
    assert tuple(full.x0[0, :2]) == (0.5, -0.6)
    for rank in range(world):
        lo, hi = sharding.shard_range(n, rank, world)
        part = P.circle_batch_fast(n, lo=lo, hi=hi)
        assert part.batch == hi - lo
        assert np.array_equal(part.g, full.g[lo:hi]) and np.array_equal(part.x0, full.x0[lo:hi])
        assert part.shared == full.shared
