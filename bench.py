#!/usr/bin/env python3
"""bench.py -- Eagen-MSM witness points/sec (BASELINE.json metric) on N B200s.

A step = one full compute_lhs_witness-equivalent pass (digits, multiples, digit sums + carry chain, all d divisor
witnesses in canonical form) over one batch of synthetic scalars/points.

  python bench.py [--gpus N --steps K --warmup W]          this repo's CUDA path
  python bench.py --impl reference [...]                    the CPU restatement of the reference on the host cores

N=1 workload: BASELINE.json configs[2] (Pallas, 2^20 points, base 5).  N>1 (torchrun, one rank per GPU): the point
range is sharded across ranks (n_total = N * 2^20, weak scaling), partial digit sums are all-gathered over NCCL,
and the d independent divisor trees are sharded by digit position.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "Eagen-MSM witness points/sec at 2^20 Pallas, 1/2/4/8 B200 vs host CPU"

# stdout carries exactly ONE JSON line (rank 0): everything else that libraries print to fd 1 (e.g. NCCL's version banner)
# is redirected to stderr for the lifetime of the process.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()

UNIT = "points/s"
BASE = 5


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region"""

    def __init__(self, index, period_ms=500):
        self.index, self.proc, self.lines, self.period_ms = index, None, [], period_ms

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def oracle_sample_step(oracle_lib, eg_inputs, log_n, curve_id=0):
    """one bounded CPU step: the oracle's full compute_lhs_witness on 2^log_n points of the same generator"""
    S, P = eg_inputs
    n = 1 << log_n
    t0 = time.perf_counter()
    oracle_lib.lhs_witness(curve_id, S[:n], P[:n], BASE)
    return time.perf_counter() - t0


def full_size_cpu_record():
    """the one-off full-size oracle run (tools/golden_full_size.py): seconds, threads, points/s at 2^20 -- quoted beside the bounded samples"""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_cpu_2p20.json")) as f:
            r = json.load(f)
        return "one-off FULL 2^%d run of the same oracle: %.0f s on %d threads = %.0f points/s (%s; profiles/r02_cpu_2p20.json)" % (
            r["log_n"], r["oracle_seconds"], r["threads"], r["points_per_second"], r.get("host", "?"))
    except Exception:
        return "no full-size CPU record committed"


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Rust crate cannot be built here) on all host cores.
    A step is the FULL compute_lhs_witness (all d divisor witnesses) on 2^ref_log_n points of the same generator; the size is
    BASELINE config 2 (2^16) unless K steps of it would not fit the time budget, in which case it is reduced -- and the line's
    config.workload always names the size that actually ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib
    oracle_lib.lib()
    cores = os.cpu_count() or 1
    oracle_lib.set_threads(cores)
    dd = oracle_lib.num_digits(0, BASE)
    # probe the host's rate on 2^12 points, then size the step: per-point cost grows like log^2 n
    probe_n = 12
    S, P = oracle_lib.synth_inputs(0, 0xEA6E0002, 1 << max(args.ref_log_n, probe_n))
    t_probe = oracle_sample_step(oracle_lib, (S, P), probe_n)
    log_n = args.ref_log_n
    def est(ln):
        return t_probe * (1 << (ln - probe_n)) * (ln / float(probe_n)) ** 2 * 0.6    # calibrated on the 16-core GPU box: 2^12 in 1.5 s, 2^16 in 26 s
    while log_n > probe_n and est(log_n) * max(args.steps, 1) > args.ref_budget_s:
        log_n -= 1
    for _ in range(args.warmup):
        oracle_sample_step(oracle_lib, (S, P), max(log_n - 3, 4))
    t = [oracle_sample_step(oracle_lib, (S, P), log_n) for _ in range(args.steps)]
    total = sum(t)
    value = (1 << log_n) * args.steps / total
    sample = "full compute_lhs_witness (%d divisor witnesses) on 2^%d Pallas points per step, %d threads, %.1f s per step; %s" % (
        dd, log_n, cores, total / args.steps, full_size_cpu_record())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64x4 (256-bit Montgomery, CPU)",
        "data": "synthetic",
        "config": {"workload": "Pallas MSM witness 2^%d points per step on the HOST CPU (%d threads), base 5, d=%d, canonical (a,b) for all %d digit positions "
                               "-- a bounded sample of the GPU arm's 2^%d-point workload (per-point CPU cost grows ~log^2 n, so the full-size CPU rate is lower: see cpu_baseline.sample)"
                               % (log_n, cores, dd, dd, args.log_n),
                   "reference_sample_log_n": log_n, "gpu_arm_log_n": args.log_n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--log-n", type=int, default=20, help="points per GPU = 2^log_n")
    ap.add_argument("--ref-log-n", type=int, default=16, help="points per CPU reference step: BASELINE config 2 (about 35 s per step on 16 cores)")
    ap.add_argument("--ref-budget-s", type=float, default=600.0, help="the reference arm shrinks its step until K steps fit this many seconds")
    ap.add_argument("--cpu-log-n", type=int, default=14, help="points of the cpu_baseline sample (about 12 s on 16 cores)")
    ap.add_argument("--curve", default="pallas", choices=["pallas", "vesta", "grumpkin"],
                    help="BASELINE config 4 is --curve vesta --log-n 21 under torchrun with 8 ranks (2^24 points)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, what the driver runs): 2^log_n points per GPU; strong: 2^log_n points in TOTAL, 2^log_n / N per GPU")
    ap.add_argument("--clock-period-ms", type=int, default=500, help="nvidia-smi polling period during the timed region (0: no sampling; diagnostic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from __graft_entry__ import load_package
    eg = load_package()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = eg.Context(args.curve, local)
    ctx.set_profiling(True)
    if args.scaling == "strong":
        if (1 << args.log_n) % world:
            raise SystemExit("--scaling strong needs 2^log_n divisible by the number of ranks")
        n_local = (1 << args.log_n) // world
    else:
        n_local = 1 << args.log_n
    n_total = n_local * world
    d = eg.num_digits({"pallas": eg.PALLAS, "vesta": eg.VESTA, "grumpkin": eg.GRUMPKIN}[args.curve], BASE)

    # synthetic inputs generated on the device (resident in HBM before the timed region)
    d_s = torch.empty(n_local * 32, dtype=torch.uint8, device=dev)
    d_p = torch.empty(n_local * 96, dtype=torch.uint8, device=dev)
    ctx.dev_synth_inputs(0xEA6E0002 + rank, n_local, d_s.data_ptr(), d_p.data_ptr())

    if world > 1:
        from eagen_b200.sharded import ShardedWitness
        sw = ShardedWitness(ctx, dist, n_local, BASE, dev)

    def step_resident():
        if world == 1:
            r = ctx.compute_lhs_witness_ptr(d_s.data_ptr(), d_p.data_ptr(), n_local, BASE, eg.CANONICAL, device=True)
            ms = r.device_ms
            r.free()
            return ms
        return sw.step(d_s, d_p)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    ctx.profile_reset()
    sampler = ClockSampler(local, args.clock_period_ms)
    barrier()
    if rank == 0 and args.clock_period_ms > 0:
        sampler.start()
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        dev_ms += step_resident()
    barrier()
    wall = time.perf_counter() - t0
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    prof = ctx.profile()
    # the step time is the device time between CUDA events recorded on the launching stream (max over ranks);
    # wall clock around the same region is kept as a cross-check
    tt = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(tt[0]), float(tt[1])
    ms_per_step = dev_ms / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # ---- e2e: through the C ABI with HOST buffers (pinned), H2D and D2H inside the timed region --------------------
    e2e = None
    if not args.no_e2e and world == 1:
        h_s = torch.empty(n_local * 32, dtype=torch.uint8).pin_memory()
        h_p = torch.empty(n_local * 96, dtype=torch.uint8).pin_memory()
        h_s.copy_(d_s)
        h_p.copy_(d_p)
        a_stride, b_stride, out_bytes = ctx.stream_layout(n_local, BASE)
        h_out = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()
        r = ctx.compute_lhs_witness_stream(h_s.data_ptr(), h_p.data_ptr(), n_local, BASE, h_out.data_ptr(), out_bytes)  # warm-up
        r.free()
        torch.cuda.synchronize()
        k = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(k):
            # the public call a user makes: host scalars/points in, all functions streamed to host memory, carries read back
            r = ctx.compute_lhs_witness_stream(h_s.data_ptr(), h_p.data_ptr(), n_local, BASE, h_out.data_ptr(), out_bytes)
            got = r.total_bytes()
            carries = r.carries
            r.free()
        torch.cuda.synchronize()
        et = (time.perf_counter() - t0) / k
        e2e = {"value": n_local / et, "unit": UNIT, "h2d_bytes_per_step": n_local * 128, "d2h_bytes_per_step": int(got + carries.nbytes),
               "ms_per_step": et * 1e3, "steps": k}
    elif world > 1:
        e2e = sw.e2e(d_s, d_p, UNIT, n_total, steps=max(3, min(args.steps, 5)))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -------------------------------------------------------------------------------
    # Scopes issued to the side stream (the point pyramid, "name~side") overlap the main stream's kernels: their event times are
    # not additive, so shares are taken over the main-stream scopes and the side-stream time is reported separately.
    hbm_peak, peak_src = peaks()
    main = [e for e in prof if "~" not in e["kernel"]]
    side_ms = sum(e["ms"] for e in prof if e["kernel"].endswith("~side"))
    idle_ms = sum(e["ms"] for e in prof if e["kernel"] == "idle~gaps")
    tot_ms = sum(e["ms"] for e in main) or 1.0
    shares = {e["kernel"]: round(e["ms"] / tot_ms, 4) for e in sorted(main, key=lambda e: -e["ms"])}
    # the forward (coset) and inverse transforms are the same kernel template, k_ntt_pass: one group for the roofline
    ntt = [e for e in main if e["kernel"] in ("ntt_forward", "ntt_inverse")]
    top = {"kernel": "k_ntt_pass (ntt_forward + ntt_inverse)", "ms": sum(e["ms"] for e in ntt), "launches": sum(e["launches"] for e in ntt),
           "bytes": sum(e["bytes"] for e in ntt), "modmul": sum(e["modmul"] for e in ntt)}
    other = max((e for e in main if e not in ntt), key=lambda e: e["ms"])
    if other["ms"] > top["ms"]:
        top = dict(other)
    per_launch_ms = top["ms"] / max(top["launches"], 1)
    achieved = top["bytes"] / (top["ms"] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("k_ntt_pass") if top["kernel"].startswith("k_ntt_pass") else tj.get(top["kernel"])
        except Exception:
            traffic = None
    roofline = {"kernel": top["kernel"], "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "peak_source": peak_src, "launches": top["launches"], "avg_launch_ms": per_launch_ms,
                "share_of_step": top["ms"] / tot_ms, "algorithmic_bytes_per_launch": top["bytes"] / max(top["launches"], 1),
                "note": "SURVEY 8d prescribes the HBM view for the transform; the kernel is NOT HBM bound: 256-bit modular arithmetic is bound by the "
                        "IMAD.WIDE multiplier pipe (ncu sm__pipe_fmaheavy_cycles_active, profiles/r02_*_full_raw.csv) -- int_roofline is the primary figure"}
    # PRIMARY: Montgomery products per second against two measured ceilings
    #   (1) the multiplier-pipe ceiling: IMAD.WIDE issue rate measured on this device (eagen_microbench 2) / 87 IMAD.WIDE per product
    #       (87 = the count in the SASS of the product, profiles/r02_sass_mix.txt)
    #   (2) the register-resident loop of the product itself (eagen_microbench 1): what the field code reaches with no memory traffic
    imad_peak = ctx.microbench(0)
    modmul_peak = ctx.microbench(1)
    wide_peak = ctx.microbench(2)
    pipe_ceiling = wide_peak / 87.0
    modmul_total = sum(e["modmul"] for e in prof)
    ach_top = top["modmul"] / (top["ms"] * 1e-3)
    ach_step = modmul_total / (ms_per_step * args.steps * 1e-3)
    int_roofline = {"primary": True, "unit": "modmul/s", "bound": "integer multiplier pipe (IMAD.WIDE)",
                    "achieved_dominant_kernel": ach_top, "achieved_whole_step": ach_step,
                    "peak_pipe_ceiling": pipe_ceiling, "peak_register_loop": modmul_peak,
                    "frac_dominant_vs_pipe_ceiling": ach_top / pipe_ceiling, "frac_dominant_vs_register_loop": ach_top / modmul_peak,
                    "frac_whole_step_vs_pipe_ceiling": ach_step / pipe_ceiling, "frac_whole_step_vs_register_loop": ach_step / modmul_peak,
                    "imad_wide_per_s_measured": wide_peak, "imad32_per_s_measured": imad_peak, "imad_wide_per_modmul": 87,
                    "modmul_per_point": modmul_total / args.steps / n_total,
                    "side_stream_scope_ms_per_step": side_ms / args.steps,
                    "main_stream_idle_between_scopes_ms_per_step": idle_ms / args.steps,
                    "note": "whole-step figure divides ALL counted products by the step time (device events around the call)"}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        import oracle_lib
        oracle_lib.lib()
        cores = os.cpu_count() or 1
        oracle_lib.set_threads(cores)
        S, P = ctx.synth_inputs(0xEA6E0002, 1 << args.cpu_log_n)
        sec = oracle_sample_step(oracle_lib, (S, P), args.cpu_log_n, eg.CURVE_IDS[args.curve])
        cpu = {"value": (1 << args.cpu_log_n) / sec, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "oracle (C++ restatement of the reference algorithm; the Rust crate cannot be built here) on the first 2^%d points of the same "
                         "workload, all 56 divisor witnesses, %.1f s on %d threads; %s" % (args.cpu_log_n, sec, cores, full_size_cpu_record())}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u32x8 (256-bit Montgomery integer arithmetic)", "data": "synthetic",
        "config": {"workload": ("%s MSM witness 2^%d points per GPU (%d total), base 5, d=%d, canonical (a,b) for all %d digit positions"
                                % (args.curve.capitalize(), args.log_n, n_total, d, d)) if args.scaling == "weak" else
                               ("%s MSM witness 2^%d points in total (%d per GPU), base 5, d=%d, canonical (a,b) for all %d digit positions"
                                % (args.curve.capitalize(), args.log_n, n_local, d, d)),
                   "l2": "inputs (128 MiB/GPU) and the ~9 GB working set exceed the 126 MB L2; no explicit flush",
                   "parallelism": "single GPU" if world == 1 else "point range sharded x%d, digit-position trees sharded x%d; NCCL all-gathers inside libeagen_msm.so "
                                                                       "(eagen_lhs_witness_sharded), torch.distributed only ships the unique id and reduces the timings" % (world, world)},
        "wall_ms_per_step": wall_ms / args.steps,
        "profiled_kernel_ms_per_step": tot_ms / args.steps,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "e2e": e2e,
        "roofline": roofline,
        "int_roofline": int_roofline,
        "kernel_shares": shares,
        "cpu_baseline": cpu,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
