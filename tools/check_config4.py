#!/usr/bin/env python3
"""Parity evidence for BASELINE config 4 (Vesta, 2^24 points point-sharded over 8 B200s) -- run under torchrun on a multi-GPU box:

   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \\
          tools/check_config4.py --log-n 21 > profiles/r02_config4_parity_8gpu.log

The oracle cannot run 2^24 points (hours of CPU), so the full-size evidence is layered:
  (1) every rank runs the SHARDED call (eagen_dev_lhs_witness_sharded: NCCL inside the library) and, on its own GPU, the SINGLE-GPU
      call over all 2^24 points; the SHA-256 of each of its functions, of the carries and of the final carry must agree.  The
      single-GPU path is the one pinned to the oracle (all functions at 2^16, SHA-256 of the oracle's full 2^20 run).
  (2) ranks 0 and world-1 check the norm identity  f(Q) f(-Q) = (-1)^n prod_i (x_Q - x(P_i))  of one of their functions over
      ALL ~13.4 M points of that position's list with Python integers: independent of every kernel (Schwartz-Zippel).
  (3) the final carry equals the library's independent windowed-bucket MSM (eagen_msm) of the same scalars and points.
Test infrastructure; prints one line per check and a verdict; exit code 0 only if everything agrees on every rank."""
import argparse
import hashlib
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=21, help="points per rank = 2^log_n")
    ap.add_argument("--curve", default="vesta")
    ap.add_argument("--no-norm", action="store_true")
    args = ap.parse_args()
    eg = load_package()
    import pyref
    from eagen_b200.sharded import ShardedWitness
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = eg.Context(args.curve, local)
    n_local, base = 1 << args.log_n, 5
    n_total = n_local * world
    # global inputs: rank r's shard comes from seed + r (as in bench.py); every rank generates all shards for its single-GPU run
    shards_s, shards_p = [], []
    for r in range(world):
        s = torch.empty(n_local * 32, dtype=torch.uint8, device=dev)
        p = torch.empty(n_local * 96, dtype=torch.uint8, device=dev)
        ctx.dev_synth_inputs(0xEA6E0003 + r, n_local, s.data_ptr(), p.data_ptr())
        shards_s.append(s)
        shards_p.append(p)
    all_s, all_p = torch.cat(shards_s), torch.cat(shards_p)
    sw = ShardedWitness(ctx, dist, n_local, base, dev)
    keep = []
    t0 = time.time()
    ms = sw.step(shards_s[rank], shards_p[rank], keep)
    mine = keep[0]
    dist.barrier()
    full = ctx.compute_lhs_witness_ptr(all_s.data_ptr(), all_p.data_ptr(), n_total, base, eg.CANONICAL | eg.KEEP_DIGITS, device=True)
    ok = True
    k0 = mine.first_function
    b0, b1 = sw.pos
    ok &= k0 == full.d - b1 and mine.num_functions == b1 - b0
    same_c = sha(mine.carries) == sha(full.carries) and sha(mine.carry) == sha(full.carry)
    ok &= same_c
    bad = []
    for s in range(mine.num_functions):
        hs = sha(mine.poly(s, 0), mine.poly(s, 1))
        hf = sha(full.poly(k0 + s, 0), full.poly(k0 + s, 1))
        if hs != hf:
            bad.append(k0 + s)
    ok &= not bad
    print("rank %d: sharded step %.1f ms; functions %d..%d of %d: sha256 %s the single-GPU run; carries %s" % (
        rank, ms, k0, k0 + mine.num_functions - 1, full.d, "==" if not bad else "MISMATCH at %s vs" % bad, "==" if same_c else "MISMATCH"), flush=True)
    if rank == 0:   # (3) independent MSM
        S = all_s.cpu().numpy().view(np.uint64).reshape(-1, 4)
        P = all_p.cpu().numpy().view(np.uint64).reshape(-1, 12)
        msm = ctx.best_multiexp(S, P)
        m_ok = bool((np.asarray(msm).reshape(-1)[:8] == mine.carry).all())
        ok &= m_ok
        print("rank 0: final carry %s eagen_msm (windowed bucket MSM over all %d points)" % ("==" if m_ok else "MISMATCH vs", n_total), flush=True)
    if not args.no_norm and rank in (0, world - 1):   # (2) norm identity over a whole list
        cv = pyref.Curve(args.curve)
        p, R = cv.p, pyref.R
        rinv = pow(R, -1, p)
        words = lambda v: [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
        from_m = lambda w: (int(w[0]) | int(w[1]) << 64 | int(w[2]) << 128 | int(w[3]) << 192) * rinv % p
        rng = pyref.SplitMix64(2024 + rank)
        Q = pyref.random_point(rng, cv)
        one = words(R % p)
        QJ = np.array([words(Q[0] * R % p) + words(Q[1] * R % p) + one, words(Q[0] * R % p) + words((-Q[1]) % p * R % p) + one], dtype=np.uint64)
        vals = mine.ev(QJ)
        slot = mine.num_functions // 2
        k = k0 + slot
        i = full.d - 1 - k                                    # iteration position of function k
        dg = full.digits[:, i].astype(np.int64)               # (n_total,)
        idx = np.nonzero(dg)[0]
        P = all_p.cpu().numpy().view(np.uint64).reshape(-1, 12)
        mult = ctx.precompute_multiplicities(P, base)         # (n_total, base-1, 8) affine
        xw = mult[idx, dg[idx] - 1, :4].astype(object)
        del mult
        xm = xw[:, 0] + (xw[:, 1] << 64) + (xw[:, 2] << 128) + (xw[:, 3] << 192)
        xq_m = Q[0] * R % p
        acc = 1
        for v in xm:
            acc = acc * (xq_m - v) % p
        acc = acc * pow(rinv, len(xm), p) % p
        npts = len(xm)
        carries = full.carries
        if i and carries[i - 1].any():
            acc = acc * pow(Q[0] - from_m(carries[i - 1][:4]), base, p) % p
            npts += base
        if carries[i].any():
            acc = acc * (Q[0] - from_m(carries[i][:4])) % p
            npts += 1
        want = acc if npts % 2 == 0 else (-acc) % p
        got = from_m(vals[slot, 0]) * from_m(vals[slot, 1]) % p
        ok &= got == want
        print("rank %d: function %d (list of %d points): norm identity f(Q) f(-Q) = (-1)^n prod (x_Q - x_i) %s" % (rank, k, npts, "OK" if got == want else "MISMATCH"), flush=True)
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("config-4 parity (%s, %d ranks x 2^%d = %d points, base %d): %s  [%.0f s]" % (
            args.curve, world, args.log_n, n_total, base, "OK" if int(t) else "MISMATCH", time.time() - t0), flush=True)
    mine.free()
    full.free()
    dist.destroy_process_group()
    sys.exit(0 if int(t) else 1)


if __name__ == "__main__":
    main()
