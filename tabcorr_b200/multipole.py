"""(s, mu) tables -> multipole tables.

The database tabulation script turns a table of xi(s, mu) into tables of the multipoles xi_l(s)
by applying halotools' ``tpcf_multipole`` to every column of ``tpcf_matrix``
(``scripts/tabulate_snapshot.py:102-113``).  The transform is linear in the table, so it is applied
once at load time as a weighted sum over the mu bins; by the same linearity the multipoles of a
(s, mu) *prediction* equal the prediction of the multipole table (tests/test_multipole.py).

``tpcf_multipole`` lives in halotools (un-vendored, unpinned; not installed here).  Its published
definition, restated [parity unpinned]:  with mu-bin centres ``mu_c`` and widths ``d_mu``,

    xi_l(s) = (2 l + 1) / 2 * sum_mu xi(s, mu) * d_mu * (P_l(mu_c) + P_l(-mu_c))

i.e. the midpoint rule of (2l+1)/2 int_{-1}^{1} xi P_l dmu for xi even in mu, tabulated on [0, 1].
"""

import copy

import numpy as np


def multipole_weights(mu_bins, order):
    """Weights ``w[mu]`` with ``xi_l(s) = sum_mu w[mu] xi(s, mu)``."""
    mu_bins = np.atleast_1d(np.asarray(mu_bins, dtype=np.float64))
    order = int(order)
    centres = 0.5 * (mu_bins[:-1] + mu_bins[1:])
    legendre = np.polynomial.legendre.Legendre.basis(order)
    return (2.0 * order + 1.0) / 2.0 * np.diff(mu_bins) * (legendre(centres) + legendre(-centres))


def tpcf_multipole(s_mu_tcpf_result, mu_bins, order=0):
    """Multipole of a ``[n_s, n_mu]`` correlation function (halotools' function of that name)."""
    xi = np.atleast_1d(np.asarray(s_mu_tcpf_result, dtype=np.float64))
    return np.sum(xi * multipole_weights(mu_bins, order), axis=-1)


def tabcorr_s_mu_to_multipole(halotab_s_mu, mu_bins, order):
    """``TabCorr`` of xi_l(s) from a ``TabCorr`` of xi(s, mu) with ``tpcf_shape == (n_s, n_mu)``
    (``scripts/tabulate_snapshot.py:102-113``).  The new table is uploaded to the device on its
    first prediction."""
    n_s, n_mu = halotab_s_mu.tpcf_shape
    weights = multipole_weights(mu_bins, order)
    if len(weights) != n_mu:
        raise ValueError('mu_bins describe {} bins, the table has {}'.format(len(weights), n_mu))
    matrix = halotab_s_mu.tpcf_matrix.reshape(n_s, n_mu, -1)
    halotab_mult = copy.copy(halotab_s_mu)
    halotab_mult.attrs = dict(halotab_s_mu.attrs)
    halotab_mult.gal_type = halotab_s_mu.gal_type.copy()
    halotab_mult.tpcf_kwargs = dict(halotab_s_mu.tpcf_kwargs)
    halotab_mult.tpcf_shape = (n_s,)
    halotab_mult.tpcf_matrix = np.einsum('smp,m->sp', matrix, weights)
    halotab_mult._device_group = None
    return halotab_mult
