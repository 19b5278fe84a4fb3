"""MCMC-scale sweeps: 1e6 .. 1e8+ parameter draws through one table on 1..8 GPUs.

The reference's usage idiom is a Python loop ``for theta in draws: halotab.predict(model)``
(``README.md:72-74``).  A sweep is that loop at scale (BASELINE.json configs[4]): the draws are
generated on the device chunk by chunk from a counter-based generator -- chunk ``c`` depends on
``(seed, c)`` only, so the draw set is the same for any number of GPUs -- every rank evaluates the
chunks ``c = rank, rank + W, ...`` on its replica of the table, and the ``[chunk, 1 + R]`` result
slabs of one round (W chunks) are collected on rank ``dst`` with one ``gather`` per round, issued
asynchronously so that it overlaps the next round's kernels.  No other collective exists on the
path.  What happens to a gathered slab is up to ``consume`` (default: keep it in a host array).
"""

import numpy as np

from .models import ASSEMBIAS_KEYS, ModelSpec, THETA_KEYS, resolve_model


class UniformPrior:
    """Independent uniform priors ``{key: (lo, hi)}`` sampled on the device.

    ``sample(c, n, device)`` returns the ``[n, len(keys)]`` draws of chunk ``c``; torch's Philox
    generator is re-seeded per chunk, so the result does not depend on which rank asks."""

    def __init__(self, bounds, seed=0):
        self.keys = list(bounds.keys())
        self.lo = np.array([bounds[k][0] for k in self.keys], dtype=np.float64)
        self.hi = np.array([bounds[k][1] for k in self.keys], dtype=np.float64)
        self.seed = int(seed)
        self._cache = {}

    def sample(self, chunk_index, n, device):
        import torch
        device = torch.device(device)
        if device not in self._cache:
            self._cache[device] = (torch.Generator(device=device),
                                   torch.from_numpy(self.lo).to(device),
                                   torch.from_numpy(self.hi - self.lo).to(device))
        generator, lo, width = self._cache[device]
        generator.manual_seed(self.seed * 1000003 + int(chunk_index))
        u = torch.rand((n, len(self.keys)), dtype=torch.float64, device=device,
                       generator=generator)
        return lo + width * u


ZHENG07_PRIOR = {'logMmin': (11.0, 14.0), 'sigma_logM': (0.05, 1.0), 'logM0': (10.0, 13.5),
                 'logM1': (12.0, 15.0), 'alpha': (0.5, 1.5)}


def _theta_tensor(prior, sample, keys=THETA_KEYS):
    """``[n, len(prior.keys)]`` samples -> ``[n, len(keys)]`` tensor in kernel order (``keys``: the
    family's ``ModelSpec.theta_keys``; parameters the prior does not name are 0, e.g. the
    assembly-bias strengths of an undecorated model) and the dict of extra columns
    (interpolation coordinates)."""
    import torch
    theta = torch.zeros((sample.shape[0], len(keys)), dtype=torch.float64, device=sample.device)
    extra = {}
    for j, key in enumerate(prior.keys):
        if key in keys:
            theta[:, keys.index(key)] = sample[:, j]
        else:
            extra[key] = sample[:, j]
    return theta, extra


def chunk_bounds(n_draws, chunk):
    return [(lo, min(lo + chunk, n_draws)) for lo in range(0, n_draws, chunk)]


def predict_sweep(halotab, prior, n_draws, chunk=1 << 20, n_gauss_prim=10, model=None,
                  consume=None, dst=0, group=None, device=None, predict=None, xi_shape=None):
    """Evaluate ``n_draws`` prior draws; collect ``(ngal, xi)`` on rank ``dst``.

    Parameters
    ----------
    halotab : TabCorr or Interpolator
        Table replica of this rank.  For an ``Interpolator`` the prior must also name its
        interpolation coordinates (the columns of ``param_dict_table``).
    prior : UniformPrior
    model : model instance or ModelSpec, optional
        Occupation family; default zheng07, decorated when the prior names the two
        ``*_assembias_param1`` strengths.
    n_draws, chunk : int
        Total number of draws and draws per chunk (one fused launch per chunk).
    consume : callable ``(lo, hi, slab)``, optional
        Called on rank ``dst`` for every gathered chunk with its draw range and the
        ``[hi - lo, 1 + R]`` device tensor (column 0 = ngal).  Default: copy into a host array.
    predict : callable ``(theta [n, n_theta]) -> (ngal [n], xi [n, ...])``, optional
        Replaces ``halotab.predict_batch`` (the gloo tests use an oracle-backed stand-in); called
        as ``predict(theta, extra)`` when the prior names columns outside the family's parameters.
    xi_shape : tuple, optional
        Shape of one prediction; default ``halotab.tpcf_shape``.

    Returns ``(ngal [n_draws], xi [n_draws, *tpcf_shape])`` numpy arrays on rank ``dst`` when
    ``consume`` is None, else the number of draws this rank evaluated.
    """
    import torch
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() \
            else torch.device('cpu')
    if model is not None:
        spec = resolve_model(model)
    else:
        spec = ModelSpec(decorated=all(k in prior.keys for k in ASSEMBIAS_KEYS))
    theta_keys = spec.theta_keys
    coordinate_keys = list(getattr(halotab, '_keys', [])) if hasattr(halotab, 'tabcorr_list') else []
    user_predict = predict is not None
    if predict is None:
        missing = [k for k in coordinate_keys if k not in prior.keys]
        if missing:
            raise ValueError('the prior does not name the interpolation coordinates {}'.format(
                ', '.join(missing)))

        def predict(theta, extra=None):
            if coordinate_keys:
                x = torch.stack([extra[k] for k in coordinate_keys], dim=1)
                return halotab.predict_batch_tensors(theta, x, spec, n_gauss_prim=n_gauss_prim)
            return halotab.predict_batch(theta, n_gauss_prim=n_gauss_prim, model=spec,
                                         as_numpy=False)
    bounds = chunk_bounds(int(n_draws), int(chunk))
    n_rounds = -(-len(bounds) // world)
    host = {}
    xi_shape = tuple(halotab.tpcf_shape) if xi_shape is None else tuple(xi_shape)
    width = 1 + int(np.prod(xi_shape))

    def default_consume(lo, hi, slab):
        if 'out' not in host:
            host['out'] = np.empty((n_draws, slab.shape[1]), dtype=np.float64)
        host['out'][lo:hi] = slab.cpu().numpy()

    sink = consume if consume is not None else default_consume
    pending = None   # (work, slabs, round) of the gather in flight
    evaluated = 0

    def finish(entry):
        work, slabs, rnd = entry
        if work is not None:
            work.wait()
        if rank != dst:
            return
        for r in range(world):
            c = rnd * world + r
            if c < len(bounds):
                lo, hi = bounds[c]
                sink(lo, hi, slabs[r][:hi - lo])

    for rnd in range(n_rounds):
        c = rnd * world + rank
        slab = None
        if c < len(bounds):
            lo, hi = bounds[c]
            theta, extra = _theta_tensor(prior, prior.sample(c, hi - lo, device), theta_keys)
            unknown = [k for k in extra if k not in coordinate_keys]
            if unknown and not user_predict:
                raise ValueError('the prior names parameters that neither the occupation family '
                                 'nor the table uses: {}'.format(', '.join(unknown)))
            if user_predict and extra:
                ngal, xi = predict(theta, extra)   # the stand-in receives the other columns too
            else:
                ngal, xi = predict(theta) if user_predict else predict(theta, extra)
            slab = torch.cat([ngal.reshape(-1, 1), xi.reshape(xi.shape[0], -1)], dim=1)
            if slab.shape[1] != width:
                raise ValueError('predict returned {} columns, expected {}'.format(
                    slab.shape[1], width))
            evaluated += hi - lo
        if world == 1:
            finish((None, [slab], rnd))
            continue
        # gather needs equal shapes on every rank: pad the short (or missing) last chunk
        if slab is None or slab.shape[0] != chunk:
            padded = torch.zeros((chunk, width), dtype=torch.float64, device=device)
            if slab is not None:
                padded[:slab.shape[0]] = slab
            slab = padded
        slabs = [torch.empty_like(slab) for _ in range(world)] if rank == dst else None
        work = dist.gather(slab.contiguous(), slabs, dst=dst, group=group, async_op=True)
        if pending is not None:
            finish(pending)   # the previous round's gather overlapped this round's kernels
        pending = (work, slabs, rnd)
    if pending is not None:
        finish(pending)
    if consume is not None:
        return evaluated
    if rank != dst:
        return None
    out = host.get('out')
    if out is None:
        out = np.empty((0, width))
    return out[:, 0], out[:, 1:].reshape((out.shape[0],) + xi_shape)
