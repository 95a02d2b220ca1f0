import os, sys, time, torch
sys.path.insert(0, '/root/repo')
from __graft_entry__ import load_package
eg = load_package()
n = 1 << 20
dev = torch.device("cuda", 0)
for ng, pct in ((0, "70,30"), (0, "60,30,10"), (0, "55,30,15"), (0, "50,25,15,10"), (0, "70,30"), (0, "60,30,10"), (0, "65,25,10")):
    os.environ["EAGEN_STREAM_SPLIT"] = pct
    ctx = eg.Context("pallas", 0)
    d_s = torch.empty(n * 32, dtype=torch.uint8, device=dev); d_p = torch.empty(n * 96, dtype=torch.uint8, device=dev)
    ctx.dev_synth_inputs(1, n, d_s.data_ptr(), d_p.data_ptr())
    h_s = torch.empty(n * 32, dtype=torch.uint8).pin_memory(); h_p = torch.empty(n * 96, dtype=torch.uint8).pin_memory()
    h_s.copy_(d_s); h_p.copy_(d_p)
    a, b, tot = ctx.stream_layout(n, 5)
    h_out = torch.empty(tot, dtype=torch.uint8).pin_memory()
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = ctx.compute_lhs_witness_stream(h_s.data_ptr(), h_p.data_ptr(), n, 5, h_out.data_ptr(), tot)
        ms = r.device_ms; r.free()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        if it: print("split %s: e2e %.1f ms (device part %.1f)" % (pct, (t1 - t0) * 1e3, ms))
    ctx.close(); del h_out
