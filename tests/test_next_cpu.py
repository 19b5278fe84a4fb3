"""CPU tests of the host logic next to the hot path: multipole weights, chunk schedules,
TableSet grouping (with an oracle-backed stand-in table)."""

import numpy as np
import pytest

from tabcorr_b200 import multipole
from tabcorr_b200.tabcorr import chunk_schedule
from tabcorr_b200.tableset import TableSet


def test_multipole_weights_match_definition():
    mu_bins = np.linspace(0, 1, 41)
    centres = 0.5 * (mu_bins[1:] + mu_bins[:-1])
    legendre = np.polynomial.legendre.Legendre
    xi = np.stack([3.0 * legendre.basis(0)(centres) - 1.5 * legendre.basis(2)(centres) +
                   0.25 * legendre.basis(4)(centres), np.cos(centres)])
    for order, expected in ((0, 3.0), (2, -1.5), (4, 0.25)):
        value = multipole.tpcf_multipole(xi, mu_bins, order=order)
        # midpoint rule with 40 bins: percent-level agreement with the exact coefficient
        assert abs(value[0] - expected) < 2e-2 * max(1.0, abs(expected))
        # the literal loop of the published definition
        loop = (2 * order + 1) / 2 * sum(
            xi[:, m] * (mu_bins[m + 1] - mu_bins[m]) *
            (legendre.basis(order)(centres[m]) + legendre.basis(order)(-centres[m]))
            for m in range(40))
        np.testing.assert_allclose(value, loop, rtol=1e-13)
    # odd multipoles of an even function vanish identically
    assert np.all(multipole.tpcf_multipole(xi, mu_bins, order=1) == 0)


def test_chunk_schedule_covers_every_draw_once():
    for n in (0, 1, 9999, 30000, 30001, 100000, 2500000):
        for chunk in ('auto', 0, 7, 25000, [10, 20], [5000, 90000, 5000]):
            bounds = chunk_schedule(n, chunk)
            assert [b[0] for b in bounds] == [0] + [b[1] for b in bounds[:-1]] if bounds else n == 0
            assert (bounds[-1][1] if bounds else 0) == n
            assert all(hi > lo for lo, hi in bounds)
    assert chunk_schedule(100000) == [(0, 10000), (10000, 90000), (90000, 100000)]
    assert max(hi - lo for lo, hi in chunk_schedule(5000000)) <= 1 << 20
    with pytest.raises(ValueError):
        chunk_schedule(10, 'fast')
    with pytest.raises(ValueError):
        chunk_schedule(10, [0])


class _Fake:
    """predict_batch stand-in returning torch tensors: result = offset + params."""

    def __init__(self, offset):
        self.offset = offset

    def predict_batch(self, params, separate_gal_type=False, as_numpy=True, **kw):
        import torch
        a = torch.as_tensor(np.asarray(params['a'], dtype=np.float64)) + self.offset
        xi = torch.stack([a, 2 * a], dim=1)
        if separate_gal_type:
            return {'centrals': a, 'satellites': -a}, {'centrals-centrals': xi}
        return a, xi


def test_table_set_groups_and_scatters():
    ts = TableSet([_Fake(0.0), _Fake(100.0), _Fake(1000.0)])
    a = np.arange(12.0)
    index = np.array([2, 0, 1, 1, 0, 2, 2, 2, 0, 1, 0, 0])
    ngal, xi = ts.predict_batch({'a': a, 'scalar': 1.0}, index, as_numpy=False)
    expected = a + np.array([0.0, 100.0, 1000.0])[index]
    assert np.array_equal(ngal.numpy(), expected)
    assert np.array_equal(xi.numpy(), np.stack([expected, 2 * expected], axis=1))
    ngal, xi = ts.predict_batch({'a': a}, index.astype(float), separate_gal_type=True,
                                as_numpy=False)
    assert np.array_equal(ngal['satellites'].numpy(), -expected)
    assert list(xi) == ['centrals-centrals']
    with pytest.raises(ValueError):
        ts.predict_batch({'a': a}, index + 1)
    with pytest.raises(ValueError):
        ts.predict_batch({'a': a}, index + 0.5)
    with pytest.raises(ValueError):
        TableSet([])
