import sys, time, torch
sys.path.insert(0, '/root/repo')
from __graft_entry__ import load_package
eg = load_package()
ctx = eg.Context("pallas", 0)
n = 1 << 20
dev = torch.device("cuda", 0)
d_s = torch.empty(n * 32, dtype=torch.uint8, device=dev); d_p = torch.empty(n * 96, dtype=torch.uint8, device=dev)
ctx.dev_synth_inputs(1, n, d_s.data_ptr(), d_p.data_ptr())
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = ctx.compute_lhs_witness_ptr(d_s.data_ptr(), d_p.data_ptr(), n, 5, 0, device=True)
    t1 = time.perf_counter()
    ms = r.device_ms
    r.free()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("call %.1f ms (device %.1f)  free %.1f ms" % ((t1 - t0) * 1e3, ms, (t2 - t1) * 1e3))
t0 = time.perf_counter(); x = torch.empty(2 * 10**9, dtype=torch.uint8, device=dev); torch.cuda.synchronize(); t1 = time.perf_counter()
print("torch alloc 2GB %.1f ms" % ((t1 - t0) * 1e3))
