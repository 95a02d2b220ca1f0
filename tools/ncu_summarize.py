#!/usr/bin/env python3
"""Summarise an `ncu --csv` launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum) per kernel.
Writes a text table and profiles/ncu_traffic.json {profile group: average DRAM bytes per launch}, which bench.py reads for
roofline.traffic.   usage: tools/ncu_summarize.py gpurun_out/launches.csv profiles/rNN"""
import collections
import csv
import json
import re
import sys

GROUP = [("k_ntt_pass", "0>", "ntt_forward"), ("k_ntt_pass", "1>", "ntt_inverse"), ("k_pointwise", "", "merge_pointwise"), ("k_den", "", "merge_den"),
         ("k_binv", "", "batch_invert"), ("k_digit_sums", "", "digit_sums"), ("k_reduce_partials", "", "digit_sums"), ("k_negbase", "", "negbase"),
         ("k_multiples_proj", "", "multiples"), ("k_scatter_points", "", "scatter_points"), ("k_fixup", "", "merge_fixup"),
         ("k_pair_", "", "pair_points"), ("k_leaf_lines", "", "pair_points"), ("k_merge_desc", "", "pair_points"), ("k_carry_chain", "", "carry_chain")]


def group_of(name):
    for key, tag, grp in GROUP:
        if key in name and (not tag or re.search(r"\(bool\)%s|, %s" % (tag[0], tag), name) or name.rstrip().endswith(tag)):
            return grp
    return None


def main():
    src, prefix = sys.argv[1], sys.argv[2]
    lines = [l for l in open(src) if not l.startswith("==")]
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    ids = collections.defaultdict(set)
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per[name][row["Metric Name"]] += v * scale
        ids[name].add(row["ID"])
    tot = sum(m["gpu__time_duration.sum"] for m in per.values())
    out = ["# per-kernel totals of one bench step under ncu (cold-cache, serialised launches: compare SHARES with bench.py kernel_shares)",
           "total_ms %.3f" % tot]
    traffic = collections.defaultdict(lambda: [0.0, 0])
    for name, m in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n = len(ids[name])
        dram = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        short = re.sub(r"\(.*", "", name)[:60]
        out.append("%-62s launches %5d  ms %9.3f  share %.4f  dram_GB %8.3f" % (short, n, m["gpu__time_duration.sum"], m["gpu__time_duration.sum"] / tot, dram / 1e9))
        g = group_of(name)
        if g:
            traffic[g][0] += dram
            traffic[g][1] += n
    open(prefix + "_launches_summary.txt", "w").write("\n".join(out) + "\n")
    tj = {g: v[0] / max(v[1], 1) for g, v in traffic.items()}
    nf, ni = traffic.get("ntt_forward", [0.0, 0]), traffic.get("ntt_inverse", [0.0, 0])
    if nf[1] + ni[1]:   # both directions are the same kernel template: bench.py's roofline group
        tj["k_ntt_pass"] = (nf[0] + ni[0]) / (nf[1] + ni[1])
    json.dump(tj, open("profiles/ncu_traffic.json", "w"), indent=1)
    print("\n".join(out[:14]))
    print(tj)


if __name__ == "__main__":
    main()
