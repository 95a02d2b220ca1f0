"""Builds libeagen_msm.so in-tree with nvcc for sm_100a (no torch extension machinery needed: the
library is a plain C-ABI shared object).  python halo2-liam-eagen-msm_b200/build.py [--force]"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("EAGEN_OUT", os.path.join(HERE, "libeagen_msm.so"))
OBJ_SUFFIX = os.environ.get("EAGEN_OBJ_SUFFIX", "")   # variant builds keep their own objects (tools/variant.sh)
SOURCES = ["capi.cu", "engine_pallas.cu", "engine_vesta.cu", "engine_grumpkin.cu"]
HEADERS = ["field.cuh", "curve.cuh", "kernels.cuh", "engine.cuh", "eagen_params.h", os.path.join("..", "..", "include", "eagen_msm.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"] + (os.environ.get("EAGEN_NVCC_EXTRA", "").split())


def stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(CSRC, s.replace(".cu", OBJ_SUFFIX + ".o"))
        objs.append(obj)
        if force or stale(obj, [src] + hdrs):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        r = subprocess.run([NVCC] + FLAGS + ["-c", src, "-o", obj], capture_output=True, text=True)
        log = os.path.join(CSRC, os.path.basename(obj) + ".ptxas.log")
        with open(log, "w") as f:
            f.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-6000:]))
        return src

    with ThreadPoolExecutor(max_workers=4) as ex:
        for done in ex.map(compile_one, jobs):
            if verbose:
                print("compiled", done)
    if jobs or force or stale(OUT, objs):
        r = subprocess.run([NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
