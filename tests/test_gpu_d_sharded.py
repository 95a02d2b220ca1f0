"""GPU tests of the multi-GPU entry point behind the C ABI (pytest -m gpu): eagen_lhs_witness_sharded with NCCL inside the
library.  One process, one context + one host thread per GPU (eagen_comm_init_all).  World size 1 exercises the whole sharded
code path (communicator, all-gathers, position ranges) on a single-GPU box; larger worlds run when the box has the GPUs.
Parity: every rank's functions against the CPU oracle's witness of ALL points, and byte equality with a single-GPU run."""
import os
import threading

import numpy as np
import pytest

import pyref

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _run_sharded(eagen, cname, world, S, P, base, flags, host_stream=False):
    """returns per-rank dicts {k0, fa, fb, carries, carry}"""
    ctxs = [eagen.Context(cname, r) for r in range(world)]
    out = [None] * world
    err = [None] * world
    try:
        eagen.comm_init_all(ctxs)
        n_local = len(P) // world

        def work(r):
            try:
                s, p = S[r * n_local:(r + 1) * n_local], P[r * n_local:(r + 1) * n_local]
                if host_stream:
                    a_stride, b_stride, nbytes = ctxs[r].sharded_layout(len(P), base, r, world)
                    buf = np.zeros(max(nbytes, 64), dtype=np.uint8)
                    s, p = np.ascontiguousarray(s), np.ascontiguousarray(p)
                    res = ctxs[r].lhs_witness_sharded_ptr(s.ctypes.data, p.ctypes.data, n_local, base, flags, out_ptr=buf.ctypes.data, out_bytes=nbytes)
                    fa, fb = [], []
                    for k in range(res.num_functions):
                        la, lb = len(res.poly(k, 0)), len(res.poly(k, 1))
                        o = k * (a_stride + b_stride) * 32
                        fa.append(buf[o:o + la * 32].view(np.uint64).reshape(-1, 4).copy())
                        fb.append(buf[o + a_stride * 32:o + a_stride * 32 + lb * 32].view(np.uint64).reshape(-1, 4).copy())
                else:
                    res = ctxs[r].lhs_witness_sharded(s, p, base, flags)
                    fa = [res.poly(k, 0) for k in range(res.num_functions)]
                    fb = [res.poly(k, 1) for k in range(res.num_functions)]
                out[r] = {"k0": res.first_function, "fa": fa, "fb": fb, "carries": res.carries, "carry": res.carry, "d": res.d}
                res.free()
            except Exception as e:  # noqa: BLE001
                err[r] = e
        th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
    finally:
        for c in ctxs:
            c.close()
    for e in err:
        if e is not None:
            raise e
    return out


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("cname,n_local,base", [("pallas", 700, 5), ("vesta", 257, 5), ("grumpkin", 130, 3)])
def test_sharded_witness_vs_oracle(oracle, eagen, gpu_ctx, world, cname, n_local, base):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cv = pyref.Curve(cname)
    S, P = gpu_ctx(cname).synth_inputs(0xEA6E0004 + world, world * n_local)
    ro = oracle.lhs_witness(cv.id, S, P, base)
    ranks = _run_sharded(eagen, cname, world, S, P, base, eagen.CANONICAL)
    seen = []
    for r, o in enumerate(ranks):
        b, e = eagen.position_range(r, world, ro.d)
        assert o["k0"] == ro.d - e and len(o["fa"]) == e - b
        assert (o["carries"] == ro.carries).all() and (o["carry"] == ro.carry).all()
        for s in range(e - b):
            k = o["k0"] + s
            seen.append(k)
            assert o["fa"][s].shape == ro.ca[k].shape and (o["fa"][s] == ro.ca[k]).all(), (r, k)
            assert o["fb"][s].shape == ro.cb[k].shape and (o["fb"][s] == ro.cb[k]).all(), (r, k)
    assert sorted(seen) == list(range(ro.d))     # the ranks' shares partition the d functions


@pytest.mark.parametrize("world", [1, 2, 8])
def test_sharded_streamed_equals_single_gpu(eagen, gpu_ctx, world):
    """config-4 shape at test size (Vesta, point-sharded, streamed to host buffers): byte equality with the single-GPU call"""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    ctx = gpu_ctx("vesta")
    n = 1 << 15
    S, P = ctx.synth_inputs(0xEA6E0003, n)
    one = ctx.compute_lhs_witness(S, P, 5, eagen.CANONICAL)
    ranks = _run_sharded(eagen, "vesta", world, S, P, 5, eagen.CANONICAL, host_stream=True)
    for o in ranks:
        assert (o["carry"] == one.carry).all()
        for s in range(len(o["fa"])):
            f = one.function(o["k0"] + s)
            assert f.a.shape == o["fa"][s].shape and (f.a == o["fa"][s]).all()
            assert f.b.shape == o["fb"][s].shape and (f.b == o["fb"][s]).all()
    one.free()


def test_sharded_errors(eagen, gpu_ctx):
    ctx = gpu_ctx("pallas")
    S, P = ctx.synth_inputs(5, 16)
    with pytest.raises(eagen.EagenError) as e:     # no communicator yet
        ctx.lhs_witness_sharded(S, P, 5)
    assert e.value.status == eagen.E_NCCL
    assert eagen.position_range(3, 8, 56) == (21, 28)
    assert len(eagen.comm_unique_id()) == eagen.COMM_ID_BYTES
