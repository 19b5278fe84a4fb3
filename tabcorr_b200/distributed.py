"""Multi-GPU predictions: shard the draws, replicate the tables, gather the results.

Every draw is independent and the tables are read-only, so the path shards with no data-path
collective: rank r of W evaluates the contiguous slice
``[B r / W, B (r + 1) / W)`` of the draws on its own replica of the table, and the per-rank result
slabs ``[B_r, 1 + R]`` are collected with ONE ``gather`` (NCCL over NVLink on GPUs, gloo in the CPU
tests).  One process per GPU, launched with ``torch.distributed.run``.
"""

import numpy as np


def shard_bounds(n_draws, rank, world_size):
    """Contiguous slice of ``n_draws`` owned by ``rank``: sizes differ by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError('rank {} outside world of size {}'.format(rank, world_size))
    return n_draws * rank // world_size, n_draws * (rank + 1) // world_size


def shard_params(params, rank, world_size):
    """Slice a dict of ``[B]`` arrays (or a ``[B, k]`` array) to this rank's draws."""
    if isinstance(params, dict):
        n_draws = max(np.shape(v)[0] for v in params.values() if np.ndim(v) > 0)
        lo, hi = shard_bounds(n_draws, rank, world_size)
        return {k: (v[lo:hi] if np.ndim(v) > 0 else v) for k, v in params.items()}
    lo, hi = shard_bounds(len(params), rank, world_size)
    return params[lo:hi]


def gather_rows(local, n_total, dst=0, group=None):
    """Collect the per-rank row slabs of a ``[n_total, C]`` result on rank ``dst``.

    ``local`` is this rank's ``[hi - lo, C]`` tensor for ``shard_bounds(n_total, rank, world)``.
    Returns the concatenated ``[n_total, C]`` tensor on ``dst`` and ``None`` elsewhere.  Slabs are
    padded to a common row count because ``gather`` needs equal shapes.
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = max(shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0]
               for r in range(world))
    if local.shape[0] != rows:
        padded = torch.zeros((rows,) + tuple(local.shape[1:]), dtype=local.dtype,
                             device=local.device)
        padded[:local.shape[0]] = local
    else:
        padded = local.contiguous()
    slabs = ([torch.empty_like(padded) for _ in range(world)] if rank == dst else None)
    dist.gather(padded, slabs, dst=dst, group=group)
    if rank != dst:
        return None
    pieces = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        pieces.append(slabs[r][:hi - lo])
    return torch.cat(pieces, dim=0)


def predict_batch_sharded(halotab, params, n_gauss_prim=10, model=None, dst=0, group=None,
                          **predict_kwargs):
    """``halotab.predict_batch`` over all ranks of the process group.

    ``params`` holds ALL draws on every rank (dict of ``[B]`` arrays or ``[B, k]`` array); each
    rank evaluates its slice and rank ``dst`` receives ``(ngal [B], xi [B, *tpcf_shape])`` as numpy
    arrays (other ranks get ``None``).  Works for ``TabCorr`` and ``Interpolator`` instances.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if isinstance(params, dict):
        n_total = max(np.shape(v)[0] for v in params.values() if np.ndim(v) > 0)
    else:
        n_total = len(params)
    local = shard_params(params, rank, world)
    ngal, xi = halotab.predict_batch(local, n_gauss_prim=n_gauss_prim, model=model,
                                     as_numpy=False, **predict_kwargs)
    shape = tuple(xi.shape[1:])
    slab = torch.cat([ngal.reshape(-1, 1), xi.reshape(xi.shape[0], -1)], dim=1)
    full = gather_rows(slab, n_total, dst=dst, group=group)
    if full is None:
        return None
    from .tabcorr import _to_host
    full = _to_host(full)
    return full[:, 0], full[:, 1:].reshape((n_total,) + shape)
