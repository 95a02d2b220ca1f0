#!/usr/bin/env python3
"""Per-kernel SASS opcode histograms of the built objects (the evidence behind every instruction-mix claim in DESIGN.md).

  python tools/sass_mix.py [-o profiles/r02_sass_mix.txt] [--filter k_ntt] [objects...]

Runs `cuobjdump -sass` (no GPU needed) on halo2-liam-eagen-msm_b200/csrc/*.o and reports, for every kernel, the static
instruction count, IMAD.WIDE / IMAD.HI / other IMAD, IADD3 (+ IADD.64 forms), LOP3/SHF/LEA/PRMT, SEL/FSEL, predicate ops, shared and
global memory operations, local-memory spills (LDL/STL), barriers, and whether any tensor-core instruction is present
(HMMA / IMMA / UTC*MMA / QGMMA ... must be absent: the path is carry-chain arithmetic, not a contraction).
"non-mul per IMAD.WIDE" is the ratio VERDICT r01 asked to bring down in k_ntt_pass (was 3.35; 2.0 in the register loop).
"""
import argparse
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUOBJDUMP = os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")
CUFILT = os.environ.get("CUFILT", "/usr/local/cuda/bin/cu++filt")

GROUPS = [
    ("IMAD.WIDE", lambda op: op.startswith("IMAD.WIDE")),
    ("IMAD.HI", lambda op: op.startswith("IMAD.HI")),
    ("IMAD.MOV/SHL/IADD", lambda op: op.startswith("IMAD.MOV") or op.startswith("IMAD.SHL") or op.startswith("IMAD.IADD")),
    ("IMAD other", lambda op: op.startswith("IMAD")),
    ("IADD3/IADD", lambda op: op.startswith("IADD")),
    ("LOP3/SHF/LEA/PRMT", lambda op: op.split(".")[0] in ("LOP3", "SHF", "LEA", "PRMT", "BREV", "FLO", "POPC")),
    ("SEL", lambda op: op.split(".")[0] in ("SEL", "FSEL", "SELP")),
    ("ISETP/PLOP3/P2R", lambda op: op.split(".")[0] in ("ISETP", "PLOP3", "P2R", "R2P", "PSETP")),
    ("MOV/UMOV/S2R", lambda op: op.split(".")[0] in ("MOV", "UMOV", "S2R", "S2UR", "CS2R", "R2UR", "LDC", "ULDC", "LDCU")),
    ("LDS/STS", lambda op: op.split(".")[0] in ("LDS", "STS", "LDSM", "STSM")),
    ("LDG/STG", lambda op: op.split(".")[0] in ("LDG", "STG", "LD", "ST", "LDGSTS", "ATOM", "ATOMG", "RED")),
    ("LDL/STL", lambda op: op.split(".")[0] in ("LDL", "STL")),
    ("BAR", lambda op: op.split(".")[0] in ("BAR", "WARPSYNC", "DEPBAR", "MEMBAR")),
    ("DFMA/DADD/DMUL", lambda op: op.split(".")[0] in ("DFMA", "DADD", "DMUL")),
    ("BRA/EXIT/...", lambda op: op.split(".")[0] in ("BRA", "EXIT", "BSSY", "BSYNC", "CALL", "RET", "NOP", "BRX", "JMP", "WARPSYNC")),
]
TENSOR = re.compile(r"^(HMMA|IMMA|DMMA|BMMA|QGMMA|HGMMA|IGMMA|UTCHMMA|UTCIMMA|UTCQMMA|UTCMXQMMA|UTCOMMA|UTC.*MMA)")


def demangle(names):
    try:
        out = subprocess.run([CUFILT] + names, capture_output=True, text=True).stdout.split("\n")
        return [o if o else n for o, n in zip(out, names)]
    except Exception:
        return names


def kernels_of(obj):
    """yield (mangled name, [opcodes])"""
    txt = subprocess.run([CUOBJDUMP, "-sass", obj], capture_output=True, text=True).stdout
    name, ops = None, []
    for ln in txt.split("\n"):
        m = re.search(r"Function : (\S+)", ln)
        if m:
            if name:
                yield name, ops
            name, ops = m.group(1), []
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m and name:
            ops.append(m.group(1))
    if name:
        yield name, ops


def short(name):
    # keep the kernel name and its template arguments readable
    name = re.sub(r"eagen::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(bool\)", "", name)
    m = re.match(r"([A-Za-z_0-9]+(?:<[^()]*?>)?)\(", name)   # kernel name + template arguments, parameter list dropped
    return m.group(1) if m else re.sub(r"\(.*$", "", name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("objects", nargs="*")
    ap.add_argument("-o", "--out")
    ap.add_argument("--filter", default="", help="regex on the demangled kernel name")
    ap.add_argument("--top", type=int, default=0, help="also list the N most frequent opcodes per kernel")
    args = ap.parse_args()
    objs = args.objects or [os.path.join(ROOT, "halo2-liam-eagen-msm_b200", "csrc", "engine_pallas.o")]
    lines = []
    hdr = ["kernel", "instr"] + [g[0] for g in GROUPS] + ["other", "non-mul/IMAD.WIDE", "tensor"]
    for obj in objs:
        ks = list(kernels_of(obj))
        names = demangle([k[0] for k in ks])
        lines.append("# %s" % os.path.relpath(obj, ROOT))
        lines.append(" | ".join(hdr))
        for (mangled, ops), dn in sorted(zip(ks, names), key=lambda z: short(z[1])):
            sn = short(dn)
            if args.filter and not re.search(args.filter, sn):
                continue
            cnt = collections.Counter()
            rest = collections.Counter()
            tensor = 0
            for op in ops:
                if TENSOR.match(op):
                    tensor += 1
                for gname, pred in GROUPS:
                    if pred(op):
                        cnt[gname] += 1
                        break
                else:
                    rest[op.split(".")[0]] += 1
            wide = cnt["IMAD.WIDE"]
            ratio = "%.2f" % ((len(ops) - wide) / wide) if wide else "-"
            row = [sn, str(len(ops))] + [str(cnt[g[0]]) for g in GROUPS] + [str(sum(rest.values())), ratio, "NONE" if tensor == 0 else str(tensor)]
            lines.append(" | ".join(row))
            if args.top:
                allc = collections.Counter(ops)
                lines.append("    top: " + ", ".join("%s %d" % kv for kv in allc.most_common(args.top)))
        lines.append("")
    text = "\n".join(lines)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")
    sys.stdout.write(text + "\n")


if __name__ == "__main__":
    main()
