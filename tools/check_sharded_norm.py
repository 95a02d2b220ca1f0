#!/usr/bin/env python3
"""Full-size property check of the SHARDED path (run under torchrun on a multi-GPU box):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29512 tools/check_sharded_norm.py --log-n 20
Every rank runs the sharded step and, for the first and last digit position it owns, checks the norm identity
   f(Q) f(-Q) = (-1)^n prod_i (x_Q - x(P_i))      over ALL n points of that position's list (points of every rank),
the left side from its device-resident functions (eagen_result_eval), the right side with Python integers from the all-gathered
digit plane, multiples table and carries.  A wrong coefficient anywhere fails (Schwartz-Zippel)."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=16)
    ap.add_argument("--curve", default="vesta")
    args = ap.parse_args()
    eg = load_package()
    import pyref
    from eagen_b200.sharded import ShardedWitness, merge_planes
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = eg.Context(args.curve, local)
    cv = pyref.Curve(args.curve)
    p, R = cv.p, pyref.R
    n_local, base = 1 << args.log_n, 5
    n_total = n_local * world
    s = torch.empty(n_local * 32, dtype=torch.uint8, device=dev)
    pts = torch.empty(n_local * 96, dtype=torch.uint8, device=dev)
    ctx.dev_synth_inputs(0xEA6E0003 + rank, n_local, s.data_ptr(), pts.data_ptr())
    sw = ShardedWitness(ctx, dist, n_local, base, dev)
    keep = []
    ms = sw.step(s, pts, keep)
    mine = keep[0]
    d = sw.d
    p0, p1 = sw.pos
    rng = pyref.SplitMix64(77 + rank)
    Q = pyref.random_point(rng, cv)

    def to_words(v):
        return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
    one = to_words(R % p)
    QJ = np.array([to_words(Q[0] * R % p) + to_words(Q[1] * R % p) + one,
                   to_words(Q[0] * R % p) + to_words((-Q[1]) % p * R % p) + one], dtype=np.uint64)
    vals = mine.ev(QJ)
    rinv = pow(R, -1, p)
    from_m = lambda w: (int(w[0]) | int(w[1]) << 64 | int(w[2]) << 128 | int(w[3]) << 192) * rinv % p
    planes = merge_planes(sw.all_planes, world, d, n_local)                    # (d, n_total) uint8, MSD first
    table = sw.all_table.view(n_total, base - 1, 64)
    carries = sw.carries.view(d, 64).cpu().numpy().view(np.uint64)             # (d, 8) affine Montgomery
    ok = True
    for i in sorted({p0, p1 - 1}):
        dg = planes[i].to(torch.int64)
        idx = torch.nonzero(dg).flatten()
        xw = table[idx, dg[idx] - 1, :32].contiguous().cpu().numpy().view(np.uint64).astype(object)   # (m, 4) x of every list point
        xm = xw[:, 0] + (xw[:, 1] << 64) + (xw[:, 2] << 128) + (xw[:, 3] << 192)
        xq_m = Q[0] * R % p
        acc = 1
        for v in xm:
            acc = acc * (xq_m - v) % p
        acc = acc * pow(rinv, len(xm), p) % p
        npts = len(xm)
        if i and carries[i - 1].any():
            acc = acc * pow(Q[0] - from_m(carries[i - 1][:4]), base, p) % p
            npts += base
        if carries[i].any():
            acc = acc * (Q[0] - from_m(carries[i][:4])) % p
            npts += 1
        want = acc if npts % 2 == 0 else (-acc) % p
        slot = p1 - 1 - i
        got = from_m(vals[slot, 0]) * from_m(vals[slot, 1]) % p
        ok &= got == want
        print("rank %d position %d: %d points, norm identity %s" % (rank, i, npts, "OK" if got == want else "MISMATCH"), flush=True)
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("sharded norm identity (%s, %d ranks, 2^%d points per rank, step %.1f ms): %s" % (args.curve, world, args.log_n, ms, "OK" if int(t) else "MISMATCH"))
    mine.free()
    dist.destroy_process_group()
    sys.exit(0 if int(t) else 1)


if __name__ == "__main__":
    main()
