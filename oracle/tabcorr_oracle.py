"""CPU oracle for the TabCorr prediction hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import this module.  The product (``tabcorr_b200``) never does: it fails
loudly when its CUDA extension is missing.

This is a plain-numpy restatement of what the reference (johannesulf/TabCorr v1.2.0) computes on
the path ``TabCorr.predict`` / ``Interpolator.predict``.  Every function cites the reference
lines it follows.  Parity status:

* TabCorr-side arithmetic (quadrature, packed quadratic form, per-gal-type split, spline):
  PINNED.  ``oracle/make_golden.py`` runs the reference's own, unmodified source
  (``oracle/refstub.py`` imports it from /root/reference with stub modules for the missing
  h5py/astropy/halotools) on the shipped fixtures and on synthetic tables, and
  ``tests/test_oracle.py`` checks this module against those recorded outputs
  (``tests/golden/*.npz``) to 1e-13.
* halotools occupation arithmetic (zheng07 centrals/satellites, Heaviside assembly bias):
  PARITY UNPINNED.  halotools is a third-party dependency that is not vendored in
  /root/reference, is not installed in this image (``pyproject.toml:12`` lists a bare,
  unpinned "halotools") and none of the reference's tests hold numeric values for it.  The
  formulas below restate the published models (Zheng et al. 2007, eqs. 1/3/5; Hearin et al. 2016
  "decorated HOD", eqs. 9-14) as implemented by halotools' ``Zheng07Cens.mean_occupation``,
  ``Zheng07Sats.mean_occupation`` and ``HeavisideAssembias`` from memory of that code; they must
  be re-checked against a real halotools install when one is available.  The call sites anchoring
  them are ``tabcorr/tabcorr.py:556-563``.  The same holds for ``Leauthaud11Oracle``
  (``Leauthaud11Cens`` / ``Leauthaud11Sats`` over ``Behroozi10SmHm``: Leauthaud et al. 2011
  eqs. 8-14, Behroozi et al. 2010 eq. 21, with halotools' unit conversions (h = 0.7 in
  ``Behroozi10SmHm``, h = 0.72 in ``Leauthaud11Sats``) and its numerical
  inversion of the stellar-to-halo-mass relation by a 100-knot interpolating cubic spline).
"""

import itertools

import numpy as np
from scipy.special import erf

__all__ = [
    'OracleTable', 'Zheng07Oracle', 'Leauthaud11Oracle', 'symmetric_matrix_to_array', 'mean_occupation', 'predict',
    'spline_interpolation_matrix', 'spline_interpolate', 'OracleInterpolator']


# --------------------------------------------------------------------------------------------
# occupation model (halotools restatement -- parity unpinned, see module docstring)
# --------------------------------------------------------------------------------------------

class Zheng07Oracle:
    """zheng07 HOD with optional Heaviside assembly-bias decoration.

    Mirrors the part of halotools' ``PrebuiltHodModelFactory('zheng07')`` /
    ``('hearin15')``-style models that ``TabCorr.mean_occupation`` consumes
    (``tabcorr/tabcorr.py:496-563``): ``param_dict``, ``gal_types``, ``redshift``,
    ``mean_occupation_centrals`` and ``mean_occupation_satellites``.

    Parameters in ``param_dict``: ``logMmin, sigma_logM, logM0, logM1, alpha`` and, when
    ``decorated``, ``mean_occupation_centrals_assembias_param1`` and
    ``mean_occupation_satellites_assembias_param1`` (strengths in [-1, 1]).
    """

    gal_types = ['centrals', 'satellites']

    def __init__(self, param_dict=None, decorated=False, split=0.5, redshift=0.0,
                 modulate_with_cenocc=False, strength_abscissa=((), ()), split_abscissa=((), ()),
                 split_ordinates=((), ())):
        self.param_dict = dict(logMmin=12.02, sigma_logM=0.26, logM0=11.38, logM1=13.31,
                               alpha=1.06)
        if decorated:
            self.param_dict['mean_occupation_centrals_assembias_param1'] = 0.5
            self.param_dict['mean_occupation_satellites_assembias_param1'] = 0.5
        if param_dict is not None:
            self.param_dict.update(param_dict)
        self.decorated = decorated
        self.split = split
        self.redshift = redshift
        self.modulate_with_cenocc = modulate_with_cenocc
        # mass-dependent decoration, (centrals, satellites): halotools HeavisideAssembias with
        # assembias_strength_abscissa / split_abscissa (restated from memory of halotools'
        # heaviside_assembias.py, parity unpinned like the rest of this class)
        self.strength_abscissa = strength_abscissa
        self.split_abscissa = split_abscissa
        self.split_ordinates = split_ordinates

    @staticmethod
    def _custom_spline(abscissa, ordinates, x):
        # halotools model_helpers.custom_spline(abscissa, ordinates, k=3): a constant for one
        # point, else scipy's InterpolatedUnivariateSpline of degree min(3, n - 1)
        from scipy.interpolate import InterpolatedUnivariateSpline
        if len(abscissa) == 1:
            return np.zeros_like(x) + ordinates[0]
        return InterpolatedUnivariateSpline(abscissa, ordinates,
                                            k=min(3, len(abscissa) - 1))(x)

    def _strength(self, gal_type, prim_haloprop):
        # HeavisideAssembias.assembias_strength: spline of the param_dict ordinates over
        # log10(prim_haloprop), clipped to [-1, 1]
        t = 0 if gal_type == 'centrals' else 1
        abscissa = self.strength_abscissa[t]
        n = max(1, len(abscissa))
        ordinates = [self.param_dict['mean_occupation_{}_assembias_param{}'.format(gal_type, k + 1)]
                     for k in range(n)]
        if n == 1:
            result = np.zeros_like(prim_haloprop) + ordinates[0]
        else:
            result = self._custom_spline(abscissa, ordinates, np.log10(prim_haloprop))
        return np.clip(result, -1.0, 1.0)

    def _split(self, gal_type, prim_haloprop):
        # HeavisideAssembias.percentile_splitting_function: spline of the fixed ordinates over
        # log10(prim_haloprop), clipped to [0, 1]
        t = 0 if gal_type == 'centrals' else 1
        if len(self.split_abscissa[t]) == 0:
            return np.zeros_like(prim_haloprop) + self.split
        result = self._custom_spline(self.split_abscissa[t], self.split_ordinates[t],
                                     np.log10(prim_haloprop))
        return np.clip(result, 0.0, 1.0)

    # -- baseline zheng07 ---------------------------------------------------------------
    def _baseline_centrals(self, prim_haloprop):
        # halotools Zheng07Cens.mean_occupation: logM = log10(mass);
        # 0.5 * (1 + erf((logM - logMmin) / sigma_logM))
        log_m = np.log10(prim_haloprop)
        return 0.5 * (1.0 + erf((log_m - self.param_dict['logMmin']) /
                                self.param_dict['sigma_logM']))

    def _baseline_satellites(self, prim_haloprop):
        # halotools Zheng07Sats.mean_occupation: M0 = 10**logM0, M1 = 10**logM1;
        # ((M - M0) / M1)**alpha where M - M0 > 0, else 0; optionally times <N_cen>.
        m0 = 10.0**self.param_dict['logM0']
        m1 = 10.0**self.param_dict['logM1']
        out = np.zeros_like(prim_haloprop, dtype=np.float64)
        ok = prim_haloprop - m0 > 0
        out[ok] = ((prim_haloprop[ok] - m0) / m1)**self.param_dict['alpha']
        if self.modulate_with_cenocc:
            out *= self._baseline_centrals(prim_haloprop)
        return out

    # -- Heaviside assembly bias ----------------------------------------------------------
    def _decorate(self, baseline, percentile, strength, lower, upper, split=None):
        # halotools HeavisideAssembias.assembias_decorator: haloes above the split percentile
        # (type 1, fraction 1 - split) get +delta, the rest -delta * (1 - split) / split, which
        # preserves the mean at fixed mass.  delta = strength * (largest perturbation keeping both
        # sub-populations inside [lower, upper]); strength is clipped to [-1, 1].  strength and
        # split are per-halo arrays (functions of the primary halo property) or scalars.
        result = np.array(baseline, dtype=np.float64)
        split = np.broadcast_to(self.split if split is None else split, result.shape)
        strength = np.clip(np.broadcast_to(strength, result.shape), -1.0, 1.0)
        ok = (result > lower) & (result < upper) & (split > 0) & (split < 1)
        f = result[ok]
        frac1 = 1.0 - split[ok]
        frac2 = split[ok]
        a = strength[ok]
        with np.errstate(invalid='ignore'):
            positive = a * np.minimum(upper - f, frac2 / frac1 * (f - lower))
            negative = -a * np.maximum(lower - f, frac2 / frac1 * (f - upper))
        delta = np.where(a > 0, positive, negative)
        type1 = np.asarray(percentile)[ok] > split[ok]
        result[ok] = np.where(type1, f + delta, f - delta * frac1 / frac2)
        return result

    def mean_occupation_centrals(self, prim_haloprop=None, sec_haloprop_percentile=None, **kw):
        prim_haloprop = np.asarray(prim_haloprop, dtype=np.float64)
        f = self._baseline_centrals(prim_haloprop)
        if self.decorated:
            f = self._decorate(f, np.asarray(sec_haloprop_percentile),
                               self._strength('centrals', prim_haloprop), 0.0, 1.0,
                               self._split('centrals', prim_haloprop))
        return f

    def mean_occupation_satellites(self, prim_haloprop=None, sec_haloprop_percentile=None, **kw):
        prim_haloprop = np.asarray(prim_haloprop, dtype=np.float64)
        f = self._baseline_satellites(prim_haloprop)
        if self.decorated:
            f = self._decorate(f, np.asarray(sec_haloprop_percentile),
                               self._strength('satellites', prim_haloprop), 0.0, np.inf,
                               self._split('satellites', prim_haloprop))
        return f


class Leauthaud11Oracle(Zheng07Oracle):
    """leauthaud11 HOD (halotools ``PrebuiltHodModelFactory('leauthaud11')``) with optional
    Heaviside assembly-bias decoration (``'hearin15'``).  PARITY UNPINNED (module docstring).

    ``param_dict``: the ten ``smhm_*`` parameters of ``Behroozi10SmHm``, ``scatter_model_param1``
    (constant log-normal scatter), ``alphasat, bsat, bcut, betacut, betasat`` and, when decorated,
    the two ``*_assembias_param1`` strengths.  ``threshold`` is log10 of the stellar-mass
    threshold, ``redshift`` the redshift at which the SMHM parameters are evaluated.
    """

    littleh = 0.7         # Behroozi10SmHm.littleh (unit conversion of the SMHM relation)
    littleh_sats = 0.72   # Leauthaud11Sats.littleh (its own constant, not the SMHM model's)
    DEFAULTS = dict(smhm_m0_0=10.72, smhm_m0_a=0.59, smhm_m1_0=12.35, smhm_m1_a=0.3,
                    smhm_beta_0=0.43, smhm_beta_a=0.18, smhm_delta_0=0.56, smhm_delta_a=0.18,
                    smhm_gamma_0=1.54, smhm_gamma_a=2.52, scatter_model_param1=0.2,
                    alphasat=1.0, bsat=10.62, bcut=1.47, betacut=-0.13, betasat=0.859)

    def __init__(self, param_dict=None, threshold=10.5, redshift=0.0, decorated=False, split=0.5,
                 modulate_with_cenocc=True, strength_abscissa=((), ()), split_abscissa=((), ()),
                 split_ordinates=((), ()), scatter_abscissa=()):
        self.param_dict = dict(self.DEFAULTS)
        if decorated:
            for t, gal_type in enumerate(('centrals', 'satellites')):
                for k in range(max(1, len(strength_abscissa[t]))):
                    self.param_dict['mean_occupation_{}_assembias_param{}'.format(
                        gal_type, k + 1)] = 0.5
        if param_dict is not None:
            self.param_dict.update(param_dict)
        self.threshold = threshold
        self.redshift = redshift
        self.decorated = decorated
        self.split = split
        # mass-dependent decoration: HeavisideAssembias keywords, see Zheng07Oracle
        self.strength_abscissa = strength_abscissa
        self.split_abscissa = split_abscissa
        self.split_ordinates = split_ordinates
        # LogNormalScatterModel with scatter_abscissa / scatter_ordinates: the scatter of a halo is
        # custom_spline(abscissa, scatter_model_param1..n) at log10(prim_haloprop) (restated from
        # memory of halotools' smhm_components.py, parity unpinned)
        self.scatter_abscissa = scatter_abscissa
        self.modulate_with_cenocc = modulate_with_cenocc

    def mean_scatter(self, prim_haloprop):
        n = len(self.scatter_abscissa)
        if n <= 1:
            return np.zeros_like(prim_haloprop) + self.param_dict['scatter_model_param1']
        ordinates = [self.param_dict['scatter_model_param{}'.format(k + 1)] for k in range(n)]
        return self._custom_spline(self.scatter_abscissa, ordinates, np.log10(prim_haloprop))

    def mean_log_halo_mass(self, log_stellar_mass):
        # Behroozi10SmHm.mean_log_halo_mass: parameters were fit with h = 0.7, inputs/outputs are
        # in h = 1 units: M* -> M* / h^2 on the way in, M_h -> M_h h on the way out.
        p, a1 = self.param_dict, 1.0 / (1.0 + self.redshift) - 1.0
        stellar_mass = 10.0**np.asarray(log_stellar_mass, dtype=np.float64) / self.littleh**2
        logm0 = p['smhm_m0_0'] + p['smhm_m0_a'] * a1
        logm1 = p['smhm_m1_0'] + p['smhm_m1_a'] * a1
        beta = p['smhm_beta_0'] + p['smhm_beta_a'] * a1
        delta = p['smhm_delta_0'] + p['smhm_delta_a'] * a1
        gamma = p['smhm_gamma_0'] + p['smhm_gamma_a'] * a1
        ratio = stellar_mass / 10.0**logm0
        log_halo_mass = (logm1 + beta * np.log10(ratio) + ratio**delta / (1.0 + ratio**(-gamma))
                         - 0.5)
        return np.log10(10.0**log_halo_mass * self.littleh)

    def mean_log_stellar_mass(self, prim_haloprop):
        # Behroozi10SmHm.mean_stellar_mass: tabulate on 100 knots, interpolate the inverse with
        # model_helpers.custom_spline = scipy InterpolatedUnivariateSpline(k=3)
        from scipy.interpolate import InterpolatedUnivariateSpline
        log_stellar_mass_table = np.linspace(8.5, 12.5, 100)
        log_halo_mass_table = self.mean_log_halo_mass(log_stellar_mass_table)
        if not np.all(np.diff(log_halo_mass_table) > 0):
            raise ValueError('the stellar-to-halo-mass relation is not monotonic')
        spline = InterpolatedUnivariateSpline(log_halo_mass_table, log_stellar_mass_table, k=3)
        return spline(np.log10(prim_haloprop))

    def _baseline_centrals(self, prim_haloprop):
        # Leauthaud11Cens.mean_occupation
        logmstar = self.mean_log_stellar_mass(prim_haloprop)
        logscatter = np.sqrt(2.0) * self.mean_scatter(np.asarray(prim_haloprop, dtype=np.float64))
        return 0.5 * (1.0 - erf((self.threshold - logmstar) / logscatter))

    def _baseline_satellites(self, prim_haloprop):
        # Leauthaud11Sats.mean_occupation with _update_satellite_params
        p = self.param_dict
        h = self.littleh_sats
        knee_threshold = 10.0**self.mean_log_halo_mass(self.threshold) * h
        knee_mass = 1.0e12
        msat = knee_mass * p['bsat'] * (knee_threshold / knee_mass)**p['betasat']
        mcut = knee_mass * p['bcut'] * (knee_threshold / knee_mass)**p['betacut']
        mass = np.asarray(prim_haloprop, dtype=np.float64)
        out = np.exp(-mcut / (mass * h)) * (mass * h / msat)**p['alphasat']
        if self.modulate_with_cenocc:
            out = out * self._baseline_centrals(mass)
        return out


# --------------------------------------------------------------------------------------------
# table container
# --------------------------------------------------------------------------------------------

class OracleTable:
    """The state of a reference ``TabCorr`` instance that ``predict`` reads.

    ``gal_type`` is a numpy structured array with the columns written by
    ``tabcorr/tabcorr.py:199-234`` (``n_h, log_prim_haloprop_min/max, sec_haloprop_percentile,
    prim_haloprop_dist_index, gal_type`` ...); ``gal_type['gal_type']`` may be bytes or str.
    """

    def __init__(self, gal_type, tpcf_matrix, tpcf_shape, mode):
        self.gal_type = gal_type
        self.tpcf_matrix = np.asarray(tpcf_matrix, dtype=np.float64)  # tabcorr.py:399
        self.tpcf_shape = tuple(int(s) for s in tpcf_shape)
        self.mode = str(mode)
        names = gal_type['gal_type']
        if names.dtype.kind == 'S':
            names = np.char.decode(names, 'utf-8')
        self.gal_type_names = np.asarray(names, dtype=str)
        self._pack_cache = None

    def __len__(self):
        return len(self.gal_type)


# --------------------------------------------------------------------------------------------
# TabCorr.predict
# --------------------------------------------------------------------------------------------

def symmetric_matrix_to_array(matrix):
    """Row-major lower triangle including the diagonal (``tabcorr/tabcorr.py:770-806``)."""
    n = matrix.shape[0]
    sel = np.zeros((n * n + n) // 2, dtype=int)
    for i in range(n):
        sel[(i * (i + 1)) // 2:(i * (i + 1)) // 2 + (i + 1)] = np.arange(i * n, i * n + i + 1)
    return matrix.ravel()[sel]


def mean_occupation(table, model, n_gauss_prim=10, **occ_kwargs):
    """Gauss-Legendre averaged occupation per table row (``tabcorr/tabcorr.py:537-578``)."""
    gt = table.gal_type
    log_min = gt['log_prim_haloprop_min']
    d_log = gt['log_prim_haloprop_max'] - log_min
    # :543-546 the reference caches the Gauss-Legendre rule on the table object
    cache = getattr(table, '_gauss_cache', None)
    if cache is None or len(cache[0]) != n_gauss_prim:
        x_gauss, w_gauss = np.polynomial.legendre.leggauss(n_gauss_prim)  # :544
        cache = ((x_gauss + 1) / 2, w_gauss)  # :546
        try:
            table._gauss_cache = cache
        except AttributeError:
            pass
    x_gauss, w_gauss = cache
    prim = 10**(log_min + d_log * x_gauss[:, np.newaxis]).T.ravel()  # :548-549
    pct = np.repeat(gt['sec_haloprop_percentile'], n_gauss_prim)  # :550-551
    names = np.repeat(table.gal_type_names, n_gauss_prim)  # :552
    occ = np.zeros(len(prim))
    cen = names == 'centrals'  # :555
    occ[cen] = model.mean_occupation_centrals(
        prim_haloprop=prim[cen], sec_haloprop_percentile=pct[cen], **occ_kwargs)  # :556-559
    occ[~cen] = model.mean_occupation_satellites(
        prim_haloprop=prim[~cen], sec_haloprop_percentile=pct[~cen], **occ_kwargs)  # :560-563
    occ = occ.reshape(len(gt), n_gauss_prim)
    prim = prim.reshape(occ.shape)
    if 'prim_haloprop_dist_index' in gt.dtype.names:
        n = gt['prim_haloprop_dist_index'][:, np.newaxis] + 1  # :570
    else:
        n = 0  # :574
    return (np.sum(w_gauss * occ * prim**n, axis=-1) / np.sum(w_gauss * prim**n, axis=-1))  # :576-578


def _pack_indices(table):
    if table._pack_cache is None:
        n = len(table)
        i1 = symmetric_matrix_to_array(np.repeat(np.arange(n), n).reshape(n, n))  # :628-634
        i2 = symmetric_matrix_to_array(np.tile(np.arange(n), n).reshape(n, n))  # :630-636
        table._pack_cache = (i1, i2, np.where(i1 == i2, 1, 2))  # :638-639
    return table._pack_cache


def predict(table, occupation, separate_gal_type=False):
    """``TabCorr.predict`` for a given occupation vector (``tabcorr/tabcorr.py:623-683``)."""
    ngal = occupation * table.gal_type['n_h']  # :623
    if table.mode == 'auto':
        i1, i2, pref = _pack_indices(table)
        ngal_sq = pref * ngal[i1] * ngal[i2]  # :641-642
    if not separate_gal_type:
        if table.mode == 'auto':
            xi = np.einsum('ij, j', table.tpcf_matrix, ngal_sq) / np.sum(ngal_sq)  # :646-647
        else:
            xi = np.einsum('ij, j', table.tpcf_matrix, ngal) / np.sum(ngal)  # :649
        return np.sum(ngal), xi.reshape(table.tpcf_shape)  # :650

    if table.mode == 'auto':
        xi = (table.tpcf_matrix * ngal_sq) / np.sum(ngal_sq)  # :653
    else:
        xi = (table.tpcf_matrix * ngal) / np.sum(ngal)  # :655
    names = table.gal_type_names
    ngal_dict, xi_dict = {}, {}
    for name in np.unique(names):  # :660-662
        ngal_dict[str(name)] = np.sum(ngal[names == name])
    if table.mode == 'auto':
        for n1, n2 in itertools.combinations_with_replacement(np.unique(names), 2):  # :665-675
            mask = symmetric_matrix_to_array(
                np.outer(n1 == names, n2 == names) | np.outer(n2 == names, n1 == names))
            xi_dict['%s-%s' % (n1, n2)] = np.sum(xi * mask, axis=1).reshape(table.tpcf_shape)
    else:
        for name in np.unique(names):  # :678-681
            xi_dict[str(name)] = np.sum(xi * (names == name), axis=1).reshape(table.tpcf_shape)
    return ngal_dict, xi_dict


# --------------------------------------------------------------------------------------------
# Interpolator
# --------------------------------------------------------------------------------------------

def spline_interpolation_matrix(xp):
    """Not-a-knot cubic spline as a matrix acting on the y-values
    (``tabcorr/interpolator.py:219-272``)."""
    xp = np.asarray(xp, dtype=np.float64)
    if len(xp) < 4:
        raise ValueError('Cannot perform spline interpolation with less than 4 values.')
    n = len(xp) - 1
    m = np.zeros((4 * n, 4 * n))
    p4, p3, p2 = np.arange(4), np.arange(3), np.arange(2)
    for i in range(n):  # :247-249 spline passes through the knots
        m[i, i * 4:(i + 1) * 4] = xp[i]**p4
        m[i + n, i * 4:(i + 1) * 4] = xp[i + 1]**p4
    for i in range(n - 1):  # :252-259 continuity of first and second derivative
        d1 = np.array([1, 2, 3]) * xp[i + 1]**p3
        d2 = np.array([2, 6]) * xp[i + 1]**p2
        m[i + 2 * n, i * 4 + 1:(i + 1) * 4] = d1
        m[i + 2 * n, (i + 1) * 4 + 1:(i + 2) * 4] = -d1
        m[i + 3 * n - 1, i * 4 + 2:(i + 1) * 4] = d2
        m[i + 3 * n - 1, (i + 1) * 4 + 2:(i + 2) * 4] = -d2
    m[-1, 3] = 6 * xp[1]  # :262-265 not-a-knot: third derivative continuous at xp[1], xp[-2]
    m[-1, 7] = -6 * xp[1]
    m[-2, -5] = 6 * xp[-2]
    m[-2, -1] = -6 * xp[-2]
    m = np.linalg.inv(m)  # :268
    a = np.zeros((4 * n, len(xp)))
    a[:, :-1] = m[:, :n]
    a[:, 1:] += m[:, n:2 * n]
    return a.reshape((n, 4, len(xp)))


def spline_interpolate(x, xp, a, yp, extrapolate=False):
    """Tensor-product evaluation, one axis at a time (``tabcorr/interpolator.py:312-331``)."""
    if not isinstance(xp, list):
        xp = [xp]
    if not isinstance(a, list):
        a = [a]
    x = np.atleast_1d(x)
    for xi, ai, xpi in zip(x, a, xp):
        i_spline = np.digitize(xi, xpi) - 1  # :319
        if xi == xpi[-1]:  # :320-321
            i_spline = len(xpi) - 2
        if i_spline < 0 or i_spline >= len(xpi) - 1:  # :322-328
            if not extrapolate:
                raise ValueError('The x-coordinates are outside of the interpolation range and '
                                 'extrapolation is turned off.')
            i_spline = min(max(i_spline, 0), len(xpi) - 2)
        yp = np.einsum('ij,j...,i', ai[i_spline], yp, xi**np.arange(4))  # :329
    return yp


class OracleInterpolator:
    """``Interpolator.__init__`` + ``predict`` (``tabcorr/interpolator.py:14-70,124-216``).

    ``tables`` is a list of :class:`OracleTable`; ``param_table`` a dict ``{key: values[T]}`` whose
    insertion order is the column order of the reference's ``param_dict_table``.
    """

    def __init__(self, tables, param_table):
        keys = list(param_table.keys())
        cols = [np.asarray(param_table[k], dtype=np.float64) for k in keys]
        if any(len(c) != len(tables) for c in cols):  # :32-34
            raise ValueError("The number of TabCorr instances does not match the number of "
                             "entries in 'param_dict_table'.")
        self.keys = keys
        self.xp = [np.sort(np.unique(c)) for c in cols]  # :41
        self.a = [spline_interpolation_matrix(xp) for xp in self.xp]  # :42
        rows = np.stack(cols, axis=1)
        if (np.prod([len(xp) for xp in self.xp]) != len(tables) or
                len(np.unique(rows, axis=0)) != len(rows)):  # :45-57
            raise ValueError("The 'param_dict_table' does not describe a grid.")
        # :59-61 lexicographic sort by all columns, first column most significant
        self.order = np.lexsort([c for c in cols[::-1]])
        self.tables = tables

    def predict(self, model, separate_gal_type=False, n_gauss_prim=10, extrapolate=False,
                **occ_kwargs):
        try:
            x_model = np.array([model.param_dict[k] for k in self.keys])  # :168-177
        except KeyError as err:
            raise ValueError('The key {} is not present in the parameter dictionary of the '
                             'model.'.format(err.args[0]))
        results = []
        for k in self.order:  # :188-194
            occ = mean_occupation(self.tables[k], model, n_gauss_prim, **occ_kwargs)
            results.append(predict(self.tables[k], occ, separate_gal_type))
        shape = [len(xp) for xp in self.xp]
        output = []
        for i in range(2):  # :198-214
            if separate_gal_type:
                output.append({})
                for key in results[0][i]:
                    data = np.array([r[i][key] for r in results])
                    data = data.reshape(shape + list(data.shape[1:]))
                    output[-1][key] = spline_interpolate(x_model, self.xp, self.a, data,
                                                         extrapolate=extrapolate)
            else:
                data = np.array([r[i] for r in results])
                data = data.reshape(shape + list(data.shape[1:]))
                output.append(spline_interpolate(x_model, self.xp, self.a, data,
                                                 extrapolate=extrapolate))
        return tuple(output)


# ---------------------------------------------------------------------------------------------
# tabulation side, the halo-bin table: restated from
# tabcorr/tabcorr.py:192-234 (n_h histogram, bin columns, prim_haloprop_dist_index),
# sort_into_bins :676-737 and distribution_index :740-767.  Pinned by tests/golden/halo_bins.npz,
# which oracle/make_golden_halo_bins.py recorded from the reference's own functions.
# ---------------------------------------------------------------------------------------------
def sort_into_bins(log_prim_haloprop, log_prim_haloprop_bins, sec_haloprop_percentile,
                   sec_haloprop_percentile_bins, x):
    """Values of ``x`` per (secondary bin, primary bin) cell (tabcorr.py:676-737, no gal_type)."""
    n_p = len(log_prim_haloprop_bins) - 1
    n_s = len(sec_haloprop_percentile_bins) - 1
    i_prim = np.digitize(log_prim_haloprop, bins=log_prim_haloprop_bins, right=False) - 1  # :715
    i_sec = np.digitize(sec_haloprop_percentile, bins=sec_haloprop_percentile_bins,
                        right=False) - 1
    inside = ~((i_prim < 0) | (i_prim >= n_p) | (i_sec < 0) | (i_sec >= n_s))  # :721
    # (the reference filters x but not the indices, so it only works when every halo is inside;
    # filtering both is the same thing there and well defined elsewhere)
    cell = (i_prim + i_sec * n_p)[inside]  # :730
    values = np.asarray(x)[inside]
    order = np.argsort(cell, kind='stable')
    counts = np.insert(np.cumsum(np.bincount(cell, minlength=n_p * n_s)), 0, 0)  # :732-735
    values = values[order]
    return [values[counts[i]:counts[i + 1]] for i in range(len(counts) - 1)]


def distribution_index(x_min, x_max, x_mean):
    """tabcorr.py:740-767, with the very scipy call the reference makes."""
    from scipy.interpolate import interp1d
    x_max = x_max / x_min
    x_mean = x_mean / x_min
    n_interp = np.linspace(-10, +10, 100)
    x_interp = ((n_interp + 1) / (n_interp + 2) * (x_max**(n_interp + 2) - 1) /
                (x_max**(n_interp + 1) - 1))
    return interp1d(x_interp, n_interp, kind='cubic', fill_value=(-10, +10),
                    bounds_error=False)(x_mean)


def halo_bin_table(prim_haloprop, sec_haloprop_percentile, log_prim_haloprop_bins,
                   sec_haloprop_percentile_bins):
    """Columns of ``gal_type`` before the centrals/satellites stacking (tabcorr.py:194-227):
    dict with ``n_h``, the four bin-edge columns, ``prim_haloprop``, ``sec_haloprop_percentile``,
    ``prim_haloprop_dist_index``, plus ``mean_prim`` (NaN where empty) for the tests."""
    prim_haloprop = np.asarray(prim_haloprop, dtype=np.float64)
    n_h, pe, se = np.histogram2d(np.log10(prim_haloprop), sec_haloprop_percentile,
                                 bins=[log_prim_haloprop_bins, sec_haloprop_percentile_bins])  # :194
    out = {'n_h': n_h.ravel(order='F')}  # :199
    grid = np.meshgrid(pe, se)  # :201
    out['log_prim_haloprop_min'] = grid[0][:-1, :-1].ravel()
    out['log_prim_haloprop_max'] = grid[0][:-1, 1:].ravel()
    out['sec_haloprop_percentile_min'] = grid[1][:-1, :-1].ravel()
    out['sec_haloprop_percentile_max'] = grid[1][1:, :-1].ravel()
    out['prim_haloprop'] = 10**(0.5 * (out['log_prim_haloprop_min'] +
                                       out['log_prim_haloprop_max']))  # :208
    out['sec_haloprop_percentile'] = 0.5 * (out['sec_haloprop_percentile_min'] +
                                            out['sec_haloprop_percentile_max'])
    members = sort_into_bins(np.log10(prim_haloprop), pe, sec_haloprop_percentile, se,
                             prim_haloprop)  # :214
    dist = np.zeros(len(out['n_h']))
    mean = np.full(len(out['n_h']), np.nan)
    for i in range(len(dist)):  # :219-226
        if len(members[i]) > 0:
            x_min = 10**out['log_prim_haloprop_min'][i]
            x_max = 10**out['log_prim_haloprop_max'][i]
            mean[i] = np.mean(members[i])
            dist[i] = distribution_index(x_min, x_max, mean[i])
    out['prim_haloprop_dist_index'] = dist
    out['mean_prim'] = mean
    return out
