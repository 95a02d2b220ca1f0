"""CPU test of the multi-GPU host logic (not gpu): world_size 2 over gloo.  Checks the sharding plan the NCCL driver
uses (eagen_lhs_witness_sharded in libeagen_msm.so; host mirror in halo2-liam-eagen-msm_b200/sharded.py): position ranges
(the library's own eagen_position_range) are a partition, per-position all-gathers of the per-rank digit planes land as the global
position-major planes, and per-rank partial digit sums combine to the global sums
(the oracle plays the ranks' arithmetic here; on the GPU box the same plan runs over NCCL)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pyref


def _worker(rank, world, port, n_local, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import load_eagen
    import oracle_lib
    eg = load_eagen()
    from eagen_b200.sharded import gather_planes_rowwise, position_range
    cv = pyref.Curve("pallas")
    base, d = 5, pyref.num_digits(cv, 5)
    rng = pyref.SplitMix64(99)
    p0, dl = pyref.random_point(rng, cv), pyref.random_point(rng, cv)
    pts = [p0]
    for _ in range(world * n_local - 1):
        pts.append(cv.add(pts[-1], dl))
    sc = [pyref.random_scalar(rng, cv) for _ in pts]
    lo, hi = rank * n_local, (rank + 1) * n_local
    # this rank's stage A on its point range (oracle stands in for the kernels): digits, and per-position sums
    S, P = oracle_lib.pack_felts(sc[lo:hi], cv.q), oracle_lib.pack_points(pts[lo:hi], cv.p)
    r = oracle_lib.lhs_witness(0, S, P, base, with_functions=False)
    planes_local = torch.from_numpy(np.ascontiguousarray(r.digits.T))          # (d, n_local) position-major
    sums = []
    for i in range(d):
        acc = None
        for j in range(n_local):
            dg = int(r.digits[j][i])
            if dg:
                acc = cv.add(acc, cv.mul(dg, pts[lo + j]))
        sums.append(acc)
    packed = torch.from_numpy(oracle_lib.pack_points(sums, cv.p).astype(np.int64))
    all_sums = torch.empty(world * packed.numel(), dtype=torch.int64)
    dist.all_gather_into_tensor(all_sums, packed.reshape(-1))
    all_sums = all_sums.view(world, d, 12)
    planes = gather_planes_rowwise(dist, planes_local)   # what the library does with one grouped ncclAllGather per position row
    # global truth
    Sg, Pg = oracle_lib.pack_felts(sc, cv.q), oracle_lib.pack_points(pts, cv.p)
    g = oracle_lib.lhs_witness(0, Sg, Pg, base, with_functions=False)
    ok = bool((planes.numpy() == g.digits.T).all())
    # carry chain over gathered partials == oracle carries
    carry = None
    gsums = all_sums.numpy().astype(np.uint64)
    for i in range(d):
        carry = cv.mul(base, cv.neg(carry))
        for w in range(world):
            part = oracle_lib.unpack_affine(gsums[w, i, :8], cv.p)[0] if gsums[w, i, 8:].any() else None
            carry = cv.add(carry, part)
        ok &= oracle_lib.unpack_affine(g.carries[i], cv.p)[0] == carry
    ranges = [position_range(k, world, d) for k in range(world)]
    ok &= ranges[0][0] == 0 and ranges[-1][1] == d and all(ranges[k][1] == ranges[k + 1][0] for k in range(world - 1))
    ret[rank] = ok
    dist.destroy_process_group()


def test_two_rank_sharding_plan_over_gloo():
    world, n_local = 2, 24
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_local, ret), nprocs=world, join=True)
    assert ret[0] and ret[1]


def test_position_ranges_balanced():
    from conftest import load_eagen
    load_eagen()
    from eagen_b200.sharded import position_range
    for d in (33, 56, 65, 129):
        for world in (1, 2, 4, 8):
            rs = [position_range(r, world, d) for r in range(world)]
            sizes = [b - a for a, b in rs]
            assert sum(sizes) == d and max(sizes) - min(sizes) <= 1
