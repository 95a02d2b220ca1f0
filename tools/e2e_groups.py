"""end-to-end time of eagen_lhs_witness_stream for several group schedules (eagen_ctx_set_stream_split): python tools/e2e_groups.py"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
eg = load_package()
n = 1 << 20
dev = torch.device("cuda", 0)
ctx = eg.Context("pallas", 0)
d_s = torch.empty(n * 32, dtype=torch.uint8, device=dev); d_p = torch.empty(n * 96, dtype=torch.uint8, device=dev)
ctx.dev_synth_inputs(1, n, d_s.data_ptr(), d_p.data_ptr())
h_s = torch.empty(n * 32, dtype=torch.uint8).pin_memory(); h_p = torch.empty(n * 96, dtype=torch.uint8).pin_memory()
h_s.copy_(d_s); h_p.copy_(d_p)
a, b, tot = ctx.stream_layout(n, 5)
h_out = torch.empty(tot, dtype=torch.uint8).pin_memory()
for rnd in range(2):
    for pct in ((100,), (70, 30), (60, 30, 10), (55, 30, 15), (50, 30, 20), (45, 30, 15, 10), (65, 25, 10), (75, 25), (80, 20)):
        ctx.set_stream_split(list(pct))
        ts = []
        for it in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = ctx.compute_lhs_witness_stream(h_s.data_ptr(), h_p.data_ptr(), n, 5, h_out.data_ptr(), tot)
            ms = r.device_ms; r.free()
            torch.cuda.synchronize(); t1 = time.perf_counter()
            if it: ts.append(((t1 - t0) * 1e3, ms))
        print("split %-18s e2e %s ms (device part %s)" % (pct, " ".join("%.1f" % t[0] for t in ts), " ".join("%.1f" % t[1] for t in ts)), flush=True)
