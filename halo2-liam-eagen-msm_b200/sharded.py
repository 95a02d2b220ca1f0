"""Multi-GPU launcher of the witness path: one process per GPU.  The data plane is entirely inside libeagen_msm.so
(eagen_lhs_witness_sharded: NCCL all-gathers on the library's own communication stream, see include/eagen_msm.h); this
module only exchanges the NCCL unique id over torch.distributed, calls the library and reduces the timings.

How the path shards:
  * K1-K3 are independent per scalar / point: rank r owns the contiguous point range [r*n_local, (r+1)*n_local).
  * The d per-position digit sums are associative: every rank reduces its own range to d projective points and the
    d x 96-byte partials are ALL-GATHERED (NCCL has no elliptic-curve reduction op); each rank then runs the d-step
    carry chain on the gathered partials (replicated, trivial).
  * The d divisor trees are independent units: rank r builds the trees of the digit positions
    position_range(r, world, d).  A tree spans the points of ALL ranks, so the digit planes (one grouped all-gather per
    position row, landing position-major) and the multiples table are all-gathered once (the one real exchange step).
No CPU fallback: every stage is inside the library, on device pointers.
"""
import time

import torch


def position_range(rank, world, d):
    """contiguous, balanced split of the d digit positions (56 = 8 x 7 at base 5): the library's own plan (eagen_position_range)"""
    from . import position_range as _pr
    return _pr(rank, world, d)


def gather_planes_rowwise(dist, planes_local):
    """(d, n_local) position-major planes of this rank -> (d, world*n_local) over the global point range: one all-gather per
    position row, which is what the library issues as ONE grouped NCCL launch (no transpose pass).  Host-side mirror for the
    gloo test of the plan."""
    d, n_local = planes_local.shape
    world = dist.get_world_size()
    out = torch.empty((d, world * n_local), dtype=planes_local.dtype, device=planes_local.device)
    for pos in range(d):
        dist.all_gather_into_tensor(out[pos], planes_local[pos].contiguous())
    return out


class ShardedWitness:
    def __init__(self, ctx, dist, n_local, base, device):
        from . import num_digits, CANONICAL, comm_unique_id
        self.ctx, self.dist, self.n_local, self.base, self.dev = ctx, dist, n_local, base, device
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.d = num_digits(ctx.curve, base)
        self.flags = CANONICAL
        # the only thing torch.distributed carries: the 128-byte NCCL unique id of the library's communicator
        box = [comm_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(self.world, self.rank, box[0])
        self.pos = position_range(self.rank, self.world, self.d)
        self.last_result_bytes = 0

    def step(self, d_scalars, d_points, keep=None):
        """one whole-job pass on device-resident shards; returns this rank's milliseconds between the CUDA events the library
        records on its compute stream around the call's device work (the collectives are ordered into that stream)"""
        torch.cuda.synchronize()
        res = self.ctx.lhs_witness_sharded_ptr(d_scalars.data_ptr(), d_points.data_ptr(), self.n_local, self.base, self.flags, device=True)
        ms = res.device_ms
        self.last_result_bytes = res.total_bytes()
        if keep is not None:
            keep.append(res)
        else:
            res.free()
        return ms

    def e2e(self, d_scalars, d_points, unit, n_total, steps=3):
        """the same pass through the host-buffer entry point: pinned HOST shards in, this rank's functions streamed into pinned host
        memory while later positions are still being computed; wall clock over `steps` calls, max over ranks"""
        dist = self.dist
        h_s = torch.empty(d_scalars.numel(), dtype=torch.uint8).pin_memory()
        h_p = torch.empty(d_points.numel(), dtype=torch.uint8).pin_memory()
        h_s.copy_(d_scalars)
        h_p.copy_(d_points)
        _, _, out_bytes = self.ctx.sharded_layout(n_total, self.base, self.rank, self.world)
        h_out = torch.empty(max(out_bytes, 64), dtype=torch.uint8).pin_memory()

        def call():
            r = self.ctx.lhs_witness_sharded_ptr(h_s.data_ptr(), h_p.data_ptr(), self.n_local, self.base, self.flags,
                                                 out_ptr=h_out.data_ptr(), out_bytes=out_bytes)
            got = r.total_bytes() + r.carries.nbytes
            r.free()
            return got
        call()   # warm-up (buffers, communicator channels)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            got = call()
        torch.cuda.synchronize()
        dist.barrier()
        et = (time.perf_counter() - t0) / steps
        tt = torch.tensor([et, float(got)], dtype=torch.float64, device=self.dev)
        dist.all_reduce(tt[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(tt[1:], op=dist.ReduceOp.SUM)
        return {"value": n_total / float(tt[0]), "unit": unit, "h2d_bytes_per_step": int(self.n_local * 128 * self.world),
                "d2h_bytes_per_step": int(tt[1]), "ms_per_step": float(tt[0]) * 1e3, "steps": steps}
