#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun on a GPU box):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py --log-n 12
Every rank runs the sharded path (its point range, NCCL all-gathers, its digit positions) and compares its functions and
the carries byte-for-byte with a single-GPU run of the whole job on the same inputs."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=12)
    ap.add_argument("--curve", default="vesta")
    args = ap.parse_args()
    eg = load_package()
    from eagen_b200.sharded import ShardedWitness
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = eg.Context(args.curve, local)
    n_local = 1 << args.log_n
    n_total = n_local * world
    base = 5
    # global inputs: rank r's shard is generated with seed + r, exactly like bench.py
    shards_s, shards_p = [], []
    for r in range(world):
        s = torch.empty(n_local * 32, dtype=torch.uint8, device=dev)
        p = torch.empty(n_local * 96, dtype=torch.uint8, device=dev)
        ctx.dev_synth_inputs(0xEA6E0003 + r, n_local, s.data_ptr(), p.data_ptr())
        shards_s.append(s)
        shards_p.append(p)
    all_s, all_p = torch.cat(shards_s), torch.cat(shards_p)
    full = ctx.compute_lhs_witness_ptr(all_s.data_ptr(), all_p.data_ptr(), n_total, base, eg.CANONICAL, device=True)
    sw = ShardedWitness(ctx, dist, n_local, base, dev)
    keep = []
    sw.step(shards_s[rank], shards_p[rank], keep)
    mine = keep[0]
    ok = bool((mine.carries == full.carries).all()) and bool((mine.carry == full.carry).all())
    d = full.d
    p0, p1 = sw.pos
    assert mine.num_functions == p1 - p0
    # slot s of the ranged result is position p1-1-s, i.e. function index k = d-1-(p1-1-s)
    for s in range(mine.num_functions):
        k = d - 1 - (p1 - 1 - s)
        fa, fb = mine.poly(s, 0), mine.poly(s, 1)
        ga, gb = full.poly(k, 0), full.poly(k, 1)
        ok &= fa.shape == ga.shape and bool((fa == ga).all()) and fb.shape == gb.shape and bool((fb == gb).all())
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("sharded parity (%s, %d ranks, 2^%d points per rank): %s" % (args.curve, world, args.log_n, "OK" if int(t) else "MISMATCH"))
    dist.destroy_process_group()
    sys.exit(0 if int(t) else 1)


if __name__ == "__main__":
    main()
