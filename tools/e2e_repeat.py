"""per-call end-to-end time of eagen_lhs_witness_stream over many calls of ONE context (looks for allocation hiccups)"""
import os, sys, time, torch
sys.path.insert(0, '/root/repo')
from __graft_entry__ import load_package
eg = load_package()
n = 1 << 20
dev = torch.device("cuda", 0)
ctx = eg.Context("pallas", 0)
d_s = torch.empty(n * 32, dtype=torch.uint8, device=dev); d_p = torch.empty(n * 96, dtype=torch.uint8, device=dev)
ctx.dev_synth_inputs(1, n, d_s.data_ptr(), d_p.data_ptr())
# the bench's resident steps first, like bench.py does
for it in range(4):
    r = ctx.compute_lhs_witness_ptr(d_s.data_ptr(), d_p.data_ptr(), n, 5, eg.CANONICAL, device=True); ms = r.device_ms; r.free()
    print("resident %d: %.1f ms, free %.1f GB" % (it, ms, torch.cuda.mem_get_info()[0] / 1e9))
h_s = torch.empty(n * 32, dtype=torch.uint8).pin_memory(); h_p = torch.empty(n * 96, dtype=torch.uint8).pin_memory()
h_s.copy_(d_s); h_p.copy_(d_p)
a, b, tot = ctx.stream_layout(n, 5)
h_out = torch.empty(tot, dtype=torch.uint8).pin_memory()
for it in range(10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = ctx.compute_lhs_witness_stream(h_s.data_ptr(), h_p.data_ptr(), n, 5, h_out.data_ptr(), tot)
    ms = r.device_ms; r.free()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("stream %d: e2e %.1f ms (device part %.1f), free %.1f GB" % (it, (t1 - t0) * 1e3, ms, torch.cuda.mem_get_info()[0] / 1e9))
