#!/usr/bin/env python3
"""BASELINE.json config 5: stand-alone sweeps of the two HBM-side kernels against the measured copy bandwidth.
   K1 negbase decomposition (scalars -> digit planes), 2^12 .. 2^24 scalars: algorithmic bytes = n * (32 + d)
   K6 forward + inverse NTT of ONE transform of 2^12 .. 2^24 elements over Fp and Fq: algorithmic bytes = 64 * T per transform
      (ideal single pass; the kernel needs ceil((log T - 10) / 8) + 1 passes)
Usage (GPU box): python tools/sweep.py > profiles/rNN_sweep.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    eg = load_package()
    dev = torch.device("cuda", 0)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    out = {"hbm_peak_gbs": peak, "negbase": [], "ntt": []}
    ctx = eg.Context("pallas", 0)
    d = eg.num_digits(eg.PALLAS, 5)
    for log_n in range(12, 25, 2):
        n = 1 << log_n
        s = torch.empty(n * 32, dtype=torch.uint8, device=dev)
        p = torch.empty(n * 96, dtype=torch.uint8, device=dev)
        ctx.dev_synth_inputs(1234, n, s.data_ptr(), p.data_ptr())
        del p
        planes = torch.empty(n * d, dtype=torch.uint8, device=dev)
        for _ in range(3):
            ctx.dev_negbase(s.data_ptr(), n, 5, planes.data_ptr())
        ms = min(ctx.dev_negbase(s.data_ptr(), n, 5, planes.data_ptr()) for _ in range(5))
        gbs = n * (32 + d) / (ms * 1e-3) / 1e9
        out["negbase"].append({"log_n": log_n, "ms": ms, "GBps": gbs, "frac_of_hbm": gbs / peak, "scalars_per_s": n / (ms * 1e-3)})
        del s, planes
    for curve in ("pallas", "vesta"):
        c = eg.Context(curve, 0)
        for log_n in range(12, 25, 2):
            T = 1 << log_n
            batch = max(1, (1 << 22) >> log_n)  # keep at least 4M elements in flight for the small sizes
            buf = torch.empty(T * batch * 32, dtype=torch.uint8, device=dev)
            c.dev_synth_inputs(99, T * batch, buf.data_ptr(), torch.empty(T * batch * 96, dtype=torch.uint8, device=dev).data_ptr())
            for _ in range(2):
                c.dev_ntt(buf.data_ptr(), log_n, batch, False)
                c.dev_ntt(buf.data_ptr(), log_n, batch, True)
            f = min(c.dev_ntt(buf.data_ptr(), log_n, batch, False) for _ in range(4))
            i = min(c.dev_ntt(buf.data_ptr(), log_n, batch, True) for _ in range(4))
            passes = 1 if log_n <= 10 else 1 + -(-(log_n - 10) // 8)
            for name, ms in (("forward", f), ("inverse", i)):
                ideal = 64.0 * T * batch / (ms * 1e-3) / 1e9
                out["ntt"].append({"field": "Fp" if curve == "pallas" else "Fq", "dir": name, "log_n": log_n, "batch": batch, "ms": ms, "passes": passes,
                                   "ideal_GBps": ideal, "frac_ideal": ideal / peak, "per_pass_GBps": ideal * passes, "frac_per_pass": ideal * passes / peak,
                                   "modmul_per_s": T * batch * log_n / 2 / (ms * 1e-3)})
            del buf
        c.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
