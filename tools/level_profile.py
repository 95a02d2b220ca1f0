#!/usr/bin/env python3
"""Per-level, per-kernel-group device time of one witness step (profiling mode 2: CUDA events around every scope).

  python tools/level_profile.py [--log-n 20] [--steps 2] [--out gpurun_out/levels.json]

Prints a table: rows = tree level of the merge loop (level l builds the parents of level l+1, transform size 2^(l+1)),
columns = kernel groups, cells = ms per step; plus the groups outside the level loop.  The events serialise nothing (one stream),
so the sum of the cells is the step's device time minus the launch gaps.
"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--curve", default="pallas")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    from __graft_entry__ import load_package
    eg = load_package()
    ctx = eg.Context(args.curve, 0)
    n = 1 << args.log_n
    dev = torch.device("cuda", 0)
    d_s = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    d_p = torch.empty(n * 96, dtype=torch.uint8, device=dev)
    ctx.dev_synth_inputs(0xEA6E0002, n, d_s.data_ptr(), d_p.data_ptr())
    for _ in range(2):
        ctx.compute_lhs_witness_ptr(d_s.data_ptr(), d_p.data_ptr(), n, 5, eg.CANONICAL, device=True).free()
    ctx.set_profiling(2)
    ctx.profile_reset()
    tot = 0.0
    for _ in range(args.steps):
        r = ctx.compute_lhs_witness_ptr(d_s.data_ptr(), d_p.data_ptr(), n, 5, eg.CANONICAL, device=True)
        tot += r.device_ms
        r.free()
    prof = ctx.profile()
    levels = collections.defaultdict(dict)
    groups = []
    for e in prof:
        name, _, lv = e["kernel"].partition("@L")
        lvl = int(lv) if lv else -1
        levels[lvl][name] = (e["ms"] / args.steps, e["modmul"] / args.steps, e["bytes"] / args.steps, e["launches"] / args.steps)
        if name not in groups:
            groups.append(name)
    print("step %.2f ms (device events around the whole call), sum of scopes %.2f ms" % (tot / args.steps, sum(e["ms"] for e in prof) / args.steps))
    print("level " + " ".join("%15s" % g[:15] for g in groups) + "    total")
    for lvl in sorted(levels):
        row = levels[lvl]
        print("%5s " % ("-" if lvl < 0 else lvl) + " ".join("%15s" % ("%.3f" % row[g][0] if g in row else "") for g in groups)
              + "  %7.3f" % sum(v[0] for v in row.values()))
    print("  sum " + " ".join("%15s" % ("%.2f" % sum(levels[l][g][0] for l in levels if g in levels[l])) for g in groups))
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"step_ms": tot / args.steps, "log_n": args.log_n, "curve": args.curve,
                       "levels": {str(l): {g: {"ms": v[0], "modmul": v[1], "bytes": v[2], "launches": v[3]} for g, v in row.items()}
                                  for l, row in levels.items()}}, f, indent=1)


if __name__ == "__main__":
    main()
