"""Multi-GPU driver of the witness path: one process per GPU, torch.distributed (NCCL over NVLink) for the plumbing.

How the path shards:
  * K1-K3 are independent per scalar / point: rank r owns the contiguous point range [r*n_local, (r+1)*n_local).
  * The d per-position digit sums are associative: every rank reduces its own range to d projective points and the
    d x 96-byte partials are ALL-GATHERED (NCCL has no elliptic-curve reduction op); each rank then runs the d-step
    carry chain on the gathered partials (replicated, trivial).
  * The d divisor trees are independent units: rank r builds the trees of the digit positions
    position_range(r, world, d).  A tree spans the points of ALL ranks, so the digit planes and the multiples table
    are all-gathered once (the one real exchange step of the path).
No CPU fallback: every stage is a call into libeagen_msm.so on device pointers.
"""
import time

import torch


def position_range(rank, world, d):
    """contiguous, balanced split of the d digit positions (56 = 8 x 7 at base 5)"""
    q, r = divmod(d, world)
    begin = rank * q + min(rank, r)
    return begin, begin + q + (1 if rank < r else 0)


def merge_planes(gathered, world, d, n_local):
    """all-gathered planes (world, d, n_local) -> position-major planes over the global point range (d, world*n_local)"""
    return gathered.view(world, d, n_local).permute(1, 0, 2).contiguous().view(d, world * n_local)


class ShardedWitness:
    def __init__(self, ctx, dist, n_local, base, device):
        from . import num_digits, CANONICAL
        self.ctx, self.dist, self.n_local, self.base, self.dev = ctx, dist, n_local, base, device
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.d = num_digits(ctx.curve, base)
        self.flags = CANONICAL
        d, w = self.d, self.world
        u8 = dict(dtype=torch.uint8, device=device)
        self.planes = torch.empty(d * n_local, **u8)
        self.table = torch.empty(n_local * (base - 1) * 64, **u8)
        self.sums = torch.empty(d * 96, **u8)
        self.all_sums = torch.empty(w * d * 96, **u8)
        self.all_planes = torch.empty(w * d * n_local, **u8)
        self.all_table = torch.empty(w * n_local * (base - 1) * 64, **u8)
        self.carries = torch.empty(d * 64, **u8)
        self.pos = position_range(self.rank, w, d)
        self.last_result_bytes = 0

    def step(self, d_scalars, d_points, keep=None):
        """one whole-job pass; returns this rank's milliseconds between two CUDA events (device idle on both sides).  The
        engine's calls are synchronous and the collectives run on torch's stream, so the events recorded on that stream before the
        first and after the last call bracket all of the rank's device work."""
        dist, ctx = self.dist, self.ctx
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        ctx.dev_shard_sums(d_scalars.data_ptr(), d_points.data_ptr(), self.n_local, self.base,
                           self.planes.data_ptr(), self.table.data_ptr(), self.sums.data_ptr())
        dist.all_gather_into_tensor(self.all_sums, self.sums)
        dist.all_gather_into_tensor(self.all_planes, self.planes)
        dist.all_gather_into_tensor(self.all_table, self.table)
        planes = merge_planes(self.all_planes, self.world, self.d, self.n_local)
        torch.cuda.synchronize()
        ctx.dev_carry_chain(self.all_sums.data_ptr(), self.world, self.base, self.carries.data_ptr())
        res = ctx.dev_trees(planes.data_ptr(), self.all_table.data_ptr(), self.carries.data_ptr(), self.n_local * self.world,
                            self.base, self.pos[0], self.pos[1], self.flags)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        self.last_result_bytes = res.total_bytes()
        if keep is not None:
            keep.append(res)
        else:
            res.free()
        return ms

    def e2e(self, d_scalars, d_points, unit, n_total):
        """same pass from pinned HOST shards, with this rank's functions read back to pinned host memory"""
        dist = self.dist
        h_s = torch.empty(d_scalars.numel(), dtype=torch.uint8).pin_memory()
        h_p = torch.empty(d_points.numel(), dtype=torch.uint8).pin_memory()
        h_s.copy_(d_scalars)
        h_p.copy_(d_points)
        keep = []
        self.step(d_scalars, d_points, keep)
        nbytes = keep[0].total_bytes()
        keep[0].free()
        h_out = torch.empty(nbytes + 4096, dtype=torch.uint8).pin_memory()
        s2, p2 = torch.empty_like(d_scalars), torch.empty_like(d_points)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        s2.copy_(h_s, non_blocking=True)
        p2.copy_(h_p, non_blocking=True)
        keep = []
        self.step(s2, p2, keep)
        got = keep[0].copy_all_into(h_out.data_ptr(), h_out.numel())
        keep[0].free()
        torch.cuda.synchronize()
        dist.barrier()
        et = time.perf_counter() - t0
        tt = torch.tensor([et, float(got)], dtype=torch.float64, device=self.dev)
        dist.all_reduce(tt[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(tt[1:], op=dist.ReduceOp.SUM)
        return {"value": n_total / float(tt[0]), "unit": unit, "h2d_bytes_per_step": int(self.n_local * 128 * self.world),
                "d2h_bytes_per_step": int(tt[1]), "ms_per_step": float(tt[0]) * 1e3, "steps": 1}
