"""Host-side multi-GPU plumbing for the two places the hot path shards: one process per GPU,
``torch.distributed`` (NCCL on GPUs; gloo in the CPU tests).

* rasterizer, view-parallel: every rank holds a replica of the Gaussians and renders its own views; the per-Gaussian
  gradients of all parameters live in ONE flat ``[N*59]`` bucket (the backward kernel writes straight into views of it),
  so a step costs exactly one all-reduce;
* MASt3R: independent pairs are split contiguously across ranks, no collective.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

GRAD_FIELDS = (("means", 3), ("quats", 4), ("scales", 3), ("opacities", 1), ("sh", 48))
GRAD_FLOATS = sum(m for _, m in GRAD_FIELDS)  # 59


class GradBucket:
    """Flat fp32 buffer + per-parameter views ([N,3], [N,4], [N,3], [N], [N,16,3])."""

    def __init__(self, n_gaussians: int, device):
        self.n = n_gaussians
        self.flat = torch.zeros(n_gaussians * GRAD_FLOATS, dtype=torch.float32, device=device)
        self.views = {}
        o = 0
        for name, m in GRAD_FIELDS:
            v = self.flat[o:o + n_gaussians * m]
            self.views[name] = v.view(n_gaussians, m) if m > 1 else v
            o += n_gaussians * m
        self.views["sh"] = self.views["sh"].view(n_gaussians, 16, 3)

    def all_reduce(self):
        """Sum over ranks (the multi-view step's semantics: loss = sum of the per-view losses)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat)
        return self


def views_for_rank(n_views: int, world: int, rank: int) -> list[int]:
    """GPU g renders views {v : v mod G = g}."""
    return [v for v in range(n_views) if v % world == rank]


def shard_pairs(n_pairs: int, world: int, rank: int) -> range:
    """Contiguous, balanced split of independent MASt3R pairs (sizes differ by at most one)."""
    base, rem = divmod(n_pairs, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))
