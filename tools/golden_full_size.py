#!/usr/bin/env python3
"""One-off full-size run of the CPU oracle: per-function SHA-256 of the canonical witness at BASELINE sizes.

  python tools/golden_full_size.py --log-n 20            # ~25-40 min on 8 cores, ~8 GB of RAM

Inputs are the product's synthetic inputs (seed 0xEA6E0002, the seed bench.py and tests/test_gpu_b_large.py use) restated
on the CPU by oracle_synth_inputs.  Output:
  tests/golden/lhs_<curve>_2p<log_n>_b<base>_hashes.json   sha256(digits), sha256(carries), sha256(a_k || b_k) for every k
  profiles/r02_cpu_2p<log_n>.json                          wall time of the oracle (= the CPU baseline at that size)
The -m gpu test tests/test_gpu_b_large.py::test_full_size_hashes compares the CUDA path with the committed hashes, which
pins every later kernel change at full size.  Test infrastructure: nothing here is imported by the product.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import oracle_lib  # noqa: E402
from hashes import witness_hashes  # noqa: E402

CURVES = {"pallas": 0, "vesta": 1, "grumpkin": 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--curve", default="pallas", choices=sorted(CURVES))
    ap.add_argument("--base", type=int, default=5)
    ap.add_argument("--seed", type=lambda s: int(s, 0), default=0xEA6E0002)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    oracle_lib.lib()
    oracle_lib.set_threads(args.threads)
    cid, n = CURVES[args.curve], 1 << args.log_n
    t0 = time.time()
    S, P = oracle_lib.synth_inputs(cid, args.seed, n)
    t_in = time.time() - t0
    t0 = time.time()
    r = oracle_lib.lhs_witness(cid, S, P, args.base)
    wall = time.time() - t0
    rec = witness_hashes(r.digits, r.carries, r.ca, r.cb)
    rec.update({"curve": args.curve, "log_n": args.log_n, "base": args.base, "seed": "0x%X" % args.seed, "d": int(r.d),
                "generator": "tools/golden_full_size.py (CPU oracle, canonical form)",
                "hash": "sha256 over little-endian Montgomery bytes: digits n x d (MSD first); carries d x 64; per function a_k then b_k"})
    out = os.path.join(ROOT, "tests", "golden", "lhs_%s_2p%d_b%d_hashes.json" % (args.curve, args.log_n, args.base))
    with open(out, "w") as f:
        json.dump(rec, f, indent=1)
    prof = {"what": "CPU oracle (C++ restatement of the reference algorithm), full compute_lhs_witness with all %d divisor witnesses" % r.d,
            "curve": args.curve, "log_n": args.log_n, "base": args.base, "threads": args.threads, "cores": os.cpu_count(),
            "oracle_seconds": r.seconds, "wall_seconds_incl_marshalling": wall, "input_seconds": t_in,
            "points_per_second": n / r.seconds, "host": "build container (no GPU)" if not os.path.exists("/dev/nvidia0") else "GPU box host"}
    with open(os.path.join(ROOT, "profiles", "r02_cpu_2p%d.json" % args.log_n), "w") as f:
        json.dump(prof, f, indent=1)
    print(json.dumps(prof))


if __name__ == "__main__":
    main()
