"""N > 1 host logic on CPU (gloo, world_size 2): contiguous slab partition, per-rank computation on the slab only,
final gather -- compared with the unsharded result.  The per-rank computation here is the CPU oracle of Stage A
(cheap); on the GPU box the same partition/gather code runs with NCCL around the CUDA path (bench.py --gpus N)."""
import os
import numpy as np
import pytest
import torch.multiprocessing as mp

from aeroapps_b200 import swath


def test_slab_bounds_cover_and_balance():
    for nobs in (0, 1, 7, 48600, 1000003):
        for world in (1, 2, 3, 8):
            b = [swath.slab_bounds(nobs, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == nobs
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        swath.slab_bounds(10, 2, 2)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aeroapps_b200 import mietables as mt
    from oracle import mie_oracle as orc
    nobs, km = 37, 12                                  # ragged split: 19 + 18
    st = mt.synthetic_state(nobs, km, seed=5)
    o = orc.OracleMie([550e-9], 0, mt.synthetic_tables(nmom=0))
    ch = np.array([550.0], np.float32)
    qm = np.asfortranarray(swath.slab_of(st["qm"], 2, world, rank)); rh = np.asfortranarray(swath.slab_of(st["rh"], 1, world, rank))
    tau, ssa, g, rc = o.getaopscalar(ch, st["vnames"], qm, rh)
    full = swath.gather_swath(tau, nobs, axis=2, dist=dist)
    if rank == 0:
        tau0, _, _, _ = o.getaopscalar(ch, st["vnames"], np.asfortranarray(st["qm"]), np.asfortranarray(st["rh"]))
        q.put(bool(full.shape == tau0.shape and np.array_equal(full, tau0)))
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slabs_and_gather(built):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert ok and all(p.exitcode == 0 for p in procs)
