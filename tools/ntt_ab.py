"""stand-alone timing of k_ntt_pass for A/B decisions (low noise: one kernel, many repetitions):
   python tools/ntt_ab.py lib1.so lib2.so ...   -> median / min ms of forward and inverse batched transforms per library"""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] != "--child":
    for lib in sys.argv[1:]:
        dst = os.path.join(ROOT, "halo2-liam-eagen-msm_b200", "libeagen_msm.so")
        if os.path.abspath(lib) != os.path.abspath(dst):
            subprocess.run(["cp", lib, dst], check=True)
        out = subprocess.run([sys.executable, __file__, "--child"], capture_output=True, text=True)
        print(lib, out.stdout.strip(), out.stderr.strip()[-300:])
    sys.exit(0)
sys.path.insert(0, ROOT)
import torch
from __graft_entry__ import load_package
eg = load_package()
ctx = eg.Context("pallas", 0)
dev = torch.device("cuda", 0)
res = {}
for log_n, batch in ((10, 27000), (8, 108000), (19, 53), (14, 1700)):
    n = (1 << log_n) * batch
    buf = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    buf.zero_()   # timing is data independent (branch-free field arithmetic)
    for inv in (0, 1):
        ts = []
        for it in range(12):
            ts.append(ctx.dev_ntt(buf.data_ptr(), log_n, batch, inv))
        ts = sorted(ts[2:])
        res["2^%d x%d %s" % (log_n, batch, "inv" if inv else "fwd")] = "%.3f/%.3f" % (ts[len(ts) // 2], ts[0])
print(json.dumps(res))
