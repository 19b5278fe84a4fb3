"""World-size-2 tests of the multi-GPU plumbing on CPU (gloo): draw sharding, the one gather of the
path, and ``predict_batch_sharded`` end to end.  The per-rank "device" here is a stand-in whose
``predict_batch`` evaluates the numpy oracle (allowed in tests only), so that what is tested is
exactly the host logic that runs around the CUDA path on the GPU box: shard bounds, padding of
ragged slabs, row order after the gather, and that only ``dst`` receives the result."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tabcorr_b200 import distributed as tcd  # noqa: E402
from tabcorr_b200 import synthetic  # noqa: E402


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 100003):
        for world in (1, 2, 3, 8):
            bounds = [tcd.shard_bounds(n, r, world) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            for (lo, hi), (lo2, hi2) in zip(bounds[:-1], bounds[1:]):
                assert hi == lo2 and lo <= hi
            sizes = [hi - lo for lo, hi in bounds]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        tcd.shard_bounds(10, 2, 2)


def test_shard_params_dict_and_array():
    draws = synthetic.make_draws(11, seed=3)
    draws['scalar'] = 0.5
    part = tcd.shard_params(draws, 1, 2)
    assert len(part['logMmin']) == 6 and part['scalar'] == 0.5
    assert np.array_equal(part['alpha'], draws['alpha'][5:])
    arr = np.arange(22.0).reshape(11, 2)
    assert np.array_equal(tcd.shard_params(arr, 0, 2), arr[:5])


class OracleBackedTable:
    """Stand-in for a device table in the gloo tests: same ``predict_batch`` contract (torch
    tensors when ``as_numpy=False``), arithmetic by the oracle."""

    def __init__(self):
        from oracle import tabcorr_oracle as orc
        self.orc = orc
        tab = synthetic.make_table(n_mass=5, n_sec=2, n_r=4, seed=5)
        self.table = orc.OracleTable(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], 'auto')
        self.tpcf_shape = self.table.tpcf_shape

    def predict_batch(self, params, n_gauss_prim=10, model=None, as_numpy=True):
        n = len(params['logMmin'])
        ngal, xi = np.empty(n), np.empty((n, 4))
        for i in range(n):
            m = self.orc.Zheng07Oracle({k: float(v[i]) for k, v in params.items()})
            ngal[i], xi[i] = self.orc.predict(
                self.table, self.orc.mean_occupation(self.table, m, n_gauss_prim))
        if as_numpy:
            return ngal, xi
        return torch.from_numpy(ngal), torch.from_numpy(xi)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_draws, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # 1. raw gather of ragged slabs
        lo, hi = tcd.shard_bounds(n_draws, rank, world)
        local = torch.arange(lo, hi, dtype=torch.float64)[:, None] * torch.tensor([[1.0, -2.0]])
        full = tcd.gather_rows(local, n_draws, dst=0)
        if rank == 0:
            assert full.shape == (n_draws, 2)
            assert torch.equal(full[:, 0], torch.arange(n_draws, dtype=torch.float64))
            assert torch.equal(full[:, 1], -2.0 * torch.arange(n_draws, dtype=torch.float64))
        else:
            assert full is None
        # 1b. chunked slab gather: rows are produced range by range and land in place on rank 0
        n_local = 7
        slab = torch.full((n_local, 3), -1.0, dtype=torch.float64)
        full_slab = torch.zeros((world * n_local, 3), dtype=torch.float64) if rank == 0 else None
        launched = []

        def compute_chunk(c0, c1):
            launched.append((c0, c1))
            rows = torch.arange(c0, c1, dtype=torch.float64) + 100.0 * rank
            slab[c0:c1] = torch.stack([rows, rows * rows, -rows], dim=1)

        tcd.gather_slab_chunks(compute_chunk, slab, full_slab, n_chunks=3, dst=0)
        assert launched == [(0, 2), (2, 4), (4, 7)]
        if rank == 0:
            want = torch.cat([torch.arange(n_local, dtype=torch.float64) + 100.0 * r
                              for r in range(world)])
            assert torch.equal(full_slab[:, 0], want) and torch.equal(full_slab[:, 2], -want)
        # 2. the sharded prediction: every rank holds all draws, evaluates its slice
        halotab = OracleBackedTable()
        draws = synthetic.make_draws(n_draws, seed=21)
        result = tcd.predict_batch_sharded(halotab, draws, n_gauss_prim=4, dst=0)
        if rank == 0:
            np.save(os.path.join(out_dir, 'ngal.npy'), result[0])
            np.save(os.path.join(out_dir, 'xi.npy'), result[1])
        else:
            assert result is None
        # 3. the same through the shared host segment (no gather: every rank writes its rows)
        assert tcd.single_node()
        previous = None
        for mode in ('host', 'auto', 'host', 'host'):
            shared = tcd.predict_batch_sharded(halotab, draws, n_gauss_prim=4, dst=0, gather=mode)
            if rank == 0:
                assert np.array_equal(shared[0], result[0]) and np.array_equal(shared[1], result[1])
                # the views of the call before live in the other segment: the writers of this call
                # (who may run ahead of dst) cannot have touched them
                if previous is not None:
                    assert not np.shares_memory(previous[1], shared[1])
                    assert np.array_equal(previous[1], result[1])
                previous = shared
            else:
                assert shared is None
        assert len(tcd._SEGMENTS) == 2   # two alternating segments, cached across calls of one size
        # another batch size drops them (and their /dev/shm entries)
        fewer = {k: v[:n_draws - 2] for k, v in draws.items()}
        tcd.predict_batch_sharded(halotab, fewer, n_gauss_prim=4, dst=0, gather='host')
        assert len(tcd._SEGMENTS) == 1
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_draws', [9, 16])
def test_world_size_2_gather_equals_single_process(tmp_path, n_draws):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_draws, str(tmp_path)), nprocs=2, join=True)
    ngal = np.load(tmp_path / 'ngal.npy')
    xi = np.load(tmp_path / 'xi.npy')
    single = OracleBackedTable().predict_batch(synthetic.make_draws(n_draws, seed=21),
                                               n_gauss_prim=4)
    # same per-draw arithmetic whatever the partition: bitwise equality
    assert np.array_equal(ngal, single[0])
    assert np.array_equal(xi, single[1])


def test_single_process_passthrough():
    local = torch.ones((3, 2), dtype=torch.float64)
    assert tcd.gather_rows(local, 3) is local


# ---------------------------------------------------------------------------------------------
# sweeps (BASELINE.json configs[4]): chunked device-side draws, round-robin chunks, async gather
# ---------------------------------------------------------------------------------------------
def _sweep_predict(table):
    from tabcorr_b200.models import THETA_KEYS

    def predict(theta):
        params = {k: theta[:, j].numpy() for j, k in enumerate(THETA_KEYS[:5])}
        return table.predict_batch(params, n_gauss_prim=3, as_numpy=False)
    return predict


def _sweep_worker(rank, world, port, n_draws, chunk, out_dir):
    from tabcorr_b200 import sweep
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        table = OracleBackedTable()
        prior = sweep.UniformPrior(sweep.ZHENG07_PRIOR, seed=5)
        result = sweep.predict_sweep(None, prior, n_draws, chunk=chunk, dst=0, device='cpu',
                                     predict=_sweep_predict(table), xi_shape=(4,))
        if rank == 0:
            np.save(os.path.join(out_dir, 'sweep_ngal.npy'), result[0])
            np.save(os.path.join(out_dir, 'sweep_xi.npy'), result[1])
        else:
            assert result is None
        # consume mode: rank 0 sees every draw range exactly once
        seen = []
        n_local = sweep.predict_sweep(None, prior, n_draws, chunk=chunk, dst=0, device='cpu',
                                      predict=_sweep_predict(table), xi_shape=(4,),
                                      consume=lambda lo, hi, slab: seen.append((lo, hi, slab.shape)))
        total = torch.tensor([n_local])
        dist.all_reduce(total)
        assert int(total.item()) == n_draws
        if rank == 0:
            assert sorted(s[:2] for s in seen) == sweep.chunk_bounds(n_draws, chunk)
            assert all(s[2] == (s[1] - s[0], 5) for s in seen)
        else:
            assert seen == []
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_draws,chunk', [(23, 4), (16, 8), (5, 8)])
def test_sweep_world_size_2_equals_single_process(tmp_path, n_draws, chunk):
    from tabcorr_b200 import sweep
    port = _free_port()
    mp.spawn(_sweep_worker, args=(2, port, n_draws, chunk, str(tmp_path)), nprocs=2, join=True)
    ngal = np.load(tmp_path / 'sweep_ngal.npy')
    xi = np.load(tmp_path / 'sweep_xi.npy')
    table = OracleBackedTable()
    prior = sweep.UniformPrior(sweep.ZHENG07_PRIOR, seed=5)
    single = sweep.predict_sweep(None, prior, n_draws, chunk=chunk, device='cpu',
                                 predict=_sweep_predict(table), xi_shape=(4,))
    assert ngal.shape == (n_draws,) and xi.shape == (n_draws, 4)
    assert np.array_equal(ngal, single[0]) and np.array_equal(xi, single[1])
    # the draw set does not depend on the chunk-to-rank assignment: chunk c = f(seed, c)
    again = prior.sample(1, min(chunk, n_draws - chunk) if n_draws > chunk else 1, 'cpu')
    assert torch.equal(again, prior.sample(1, again.shape[0], 'cpu'))
    lo, hi = sweep.ZHENG07_PRIOR['logMmin']
    assert float(again[:, 0].min()) >= lo and float(again[:, 0].max()) <= hi


# ---------------------------------------------------------------------------------------------
# OverlappedGather: the gather of step k behind the kernels of step k + 1 (double-buffered slabs)
# ---------------------------------------------------------------------------------------------
def _overlap_worker(rank, world, port, n_local, n_steps):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        pipe = tcd.OverlappedGather(n_local, 3, 'cpu', dst=0, depth=2)
        landed = []
        for step in range(n_steps):
            slab = pipe.begin()
            # every step overwrites the slab it was handed: values encode (step, rank, row)
            slab.copy_(torch.arange(n_local * 3, dtype=torch.float64).reshape(n_local, 3) +
                       1000.0 * step + 100.0 * rank)
            full = pipe.submit()
            if rank == 0:
                pipe.wait(step)   # a consumer reads a step's buffer after waiting for its gather
                landed.append(full.clone())
            else:
                assert full is None
        pipe.finish()
        assert pipe.works == {}
        if rank == 0:
            for step, full in enumerate(landed):
                for r in range(world):
                    expect = (torch.arange(n_local * 3, dtype=torch.float64).reshape(n_local, 3) +
                              1000.0 * step + 100.0 * r)
                    assert torch.equal(full[r * n_local:(r + 1) * n_local], expect)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_overlapped_gather_world_size_2():
    mp.spawn(_overlap_worker, args=(2, _free_port(), 5, 5), nprocs=2, join=True)


def test_overlapped_gather_single_process_passthrough():
    pipe = tcd.OverlappedGather(4, 2, 'cpu')
    slab = pipe.begin()
    slab.fill_(7.0)
    assert pipe.submit() is slab
    assert pipe.begin() is not slab     # the other buffer
    pipe.finish()
