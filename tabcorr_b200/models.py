"""Occupation models understood by the occupation kernel.

The reference evaluates ``model.mean_occupation_centrals/satellites`` of a halotools
``HodModelFactory`` on the CPU (``tabcorr/tabcorr.py:556-563``).  Here the model object only
*describes* the occupation functions: the arithmetic runs in the CUDA occupation kernel
(``csrc/occupation.cuh``, ``csrc/leauthaud11.cuh``).  ``resolve_model`` maps

* a halotools model (recognised by the class names of its occupation components -- halotools
  itself is never imported),
* the light-weight stand-ins below (``PrebuiltHodModelFactory('zheng07' | 'decorated-zheng07' |
  'leauthaud11' | 'hearin15')``),
* or any duck-typed object with a zheng07 or leauthaud11 ``param_dict`` (plus optional
  ``decorated`` / ``split`` / ``modulate_with_cenocc`` / ``threshold`` / ``redshift`` attributes)

to a kernel family descriptor.  Anything else raises ``NotImplementedError`` -- there is no CPU
fallback; occupations computed elsewhere can still be passed to ``predict`` as an ndarray
(``tabcorr/tabcorr.py:616-621``).
"""

from types import SimpleNamespace

import numpy as np

ZHENG07_KEYS = ('logMmin', 'sigma_logM', 'logM0', 'logM1', 'alpha')
ASSEMBIAS_KEYS = ('mean_occupation_centrals_assembias_param1',
                  'mean_occupation_satellites_assembias_param1')
THETA_KEYS = ZHENG07_KEYS + ASSEMBIAS_KEYS

# halotools Leauthaud11Cens / Leauthaud11Sats over Behroozi10SmHm: param_dict keys in the order
# of the kernel's parameter vector (include/tabcorr_b200.h, TC_FAMILY_LEAUTHAUD11)
LEAUTHAUD11_KEYS = ('smhm_m0_0', 'smhm_m0_a', 'smhm_m1_0', 'smhm_m1_a', 'smhm_beta_0',
                    'smhm_beta_a', 'smhm_delta_0', 'smhm_delta_a', 'smhm_gamma_0', 'smhm_gamma_a',
                    'scatter_model_param1', 'alphasat', 'bsat', 'bcut', 'betacut', 'betasat')
# Behroozi et al. (2010) table 2 / Leauthaud et al. (2011) values halotools ships as defaults
LEAUTHAUD11_DEFAULTS = (10.72, 0.59, 12.35, 0.3, 0.43, 0.18, 0.56, 0.18, 1.54, 2.52, 0.2,
                        1.0, 10.62, 1.47, -0.13, 0.859)
FAMILY_ZHENG07, FAMILY_LEAUTHAUD11 = 0, 1
FAMILY_KEYS = {FAMILY_ZHENG07: ZHENG07_KEYS, FAMILY_LEAUTHAUD11: LEAUTHAUD11_KEYS}

# Zheng et al. (2007) table 1 (SDSS), the values halotools' ``PrebuiltHodModelFactory('zheng07',
# threshold=...)`` loads.  Restated from the paper; -18 and -21 are cross-checked against the
# values quoted in SURVEY.md section 8(a4).
ZHENG07_PUBLISHED = {
    -18.0: (11.35, 0.25, 11.20, 12.40, 0.83),
    -18.5: (11.46, 0.24, 10.59, 12.68, 0.97),
    -19.0: (11.60, 0.26, 11.49, 12.83, 1.02),
    -19.5: (11.75, 0.28, 11.69, 13.01, 1.06),
    -20.0: (12.02, 0.26, 11.38, 13.31, 1.06),
    -20.5: (12.30, 0.21, 11.84, 13.58, 1.12),
    -21.0: (12.79, 0.39, 11.92, 13.94, 1.15),
    -21.5: (13.38, 0.51, 13.94, 13.91, 1.04),
    -22.0: (14.22, 0.77, 14.00, 14.69, 0.87),
}


MAX_KNOTS = 4   # TC_MAX_KNOTS of the C ABI


def _knots(values, name):
    """Control points of a mass-dependent strength / split as a tuple of floats."""
    if values is None:
        return ()
    values = tuple(float(v) for v in np.atleast_1d(np.asarray(values, dtype=np.float64)))
    if len(values) > MAX_KNOTS:
        raise NotImplementedError('{}: at most {} control points are implemented (an interpolating '
                                  'polynomial through them; halotools switches to a cubic spline '
                                  'from the fifth)'.format(name, MAX_KNOTS))
    return values


def assembias_keys(gal_type, n):
    """halotools' ``param_dict`` keys of the ``n`` strength ordinates of one galaxy type."""
    return tuple('mean_occupation_{}_assembias_param{}'.format(gal_type, k + 1) for k in range(n))


class ModelSpec:
    """What the kernel needs to know about a model (mirrors ``tc_model`` in the C ABI).

    Mass-dependent decoration (halotools ``HeavisideAssembias`` with ``assembias_strength_abscissa``
    / ``split_abscissa``; either family): ``strength_abscissa = (cen, sat)`` are the log10
    primary-property positions of each type's strength ordinates -- the draw then carries
    ``mean_occupation_<type>_assembias_param1..n`` -- and ``split_abscissa`` / ``split_ordinates``
    give each type's splitting percentile as a function of mass.  An empty tuple means the
    constant of the plain model (one strength per type, ``split``)."""

    def __init__(self, family=0, decorated=False, modulate_with_cenocc=False, split=0.5,
                 threshold=0.0, redshift=0.0, strength_abscissa=((), ()),
                 split_abscissa=((), ()), split_ordinates=((), ()), scatter_abscissa=()):
        self.family = int(family)
        if self.family not in FAMILY_KEYS:
            raise NotImplementedError('unknown occupation family {}'.format(family))
        self.decorated = bool(decorated)
        self.modulate_with_cenocc = bool(modulate_with_cenocc)
        self.split = float(split)
        self.threshold = float(threshold)   # leauthaud11: log10 stellar-mass threshold
        self.redshift = float(redshift)     # leauthaud11: redshift of the SMHM parameters
        self._theta_keys = FAMILY_KEYS[self.family] + ASSEMBIAS_KEYS
        self.mass_dependent = False
        self.n_strength = (1, 1)    # strength ordinates per draw (centrals, satellites)
        # leauthaud11: log10 primary-property positions of the scatter_model_param1..n ordinates
        # (halotools LogNormalScatterModel scatter_abscissa); one or none: the constant scatter
        self.scatter_abscissa = ()
        if len(scatter_abscissa) > 1:
            return self._init_general(strength_abscissa, split_abscissa, split_ordinates,
                                      scatter_abscissa)
        if not (any(len(v) for v in strength_abscissa) or any(len(v) for v in split_abscissa) or
                any(len(v) for v in split_ordinates)):
            # the plain model: one strength per type, one split (the latency paths construct a
            # ModelSpec per call, so nothing is validated or formatted here)
            self.strength_abscissa = self.split_abscissa = self.split_ordinates = ((), ())
            return
        self._init_general(strength_abscissa, split_abscissa, split_ordinates, ())

    def _init_general(self, strength_abscissa, split_abscissa, split_ordinates, scatter_abscissa):
        self.strength_abscissa = tuple(_knots(v, 'assembias_strength_abscissa')
                                       for v in strength_abscissa)
        self.split_abscissa = tuple(_knots(v, 'split_abscissa') for v in split_abscissa)
        self.split_ordinates = tuple(_knots(v, 'split') for v in split_ordinates)
        for absc, ordi in zip(self.split_abscissa, self.split_ordinates):
            if len(absc) != len(ordi):
                raise ValueError('split_abscissa and split ordinates differ in length')
        for absc in self.strength_abscissa + self.split_abscissa:
            if any(b <= a for a, b in zip(absc[:-1], absc[1:])):
                raise ValueError('abscissa must increase strictly')
        self.mass_dependent = (any(len(a) > 1 for a in self.strength_abscissa) or
                               any(len(a) > 0 for a in self.split_abscissa))
        if self.mass_dependent and not self.decorated:
            raise NotImplementedError('mass-dependent assembly bias needs a decorated model')
        self.n_strength = tuple(max(1, len(a)) for a in self.strength_abscissa)
        scatter = _knots(scatter_abscissa, 'scatter_abscissa')
        if len(scatter) > 1:
            if self.family != FAMILY_LEAUTHAUD11:
                raise NotImplementedError('a stellar-mass scatter belongs to the leauthaud11 family')
            if any(b <= a for a, b in zip(scatter[:-1], scatter[1:])):
                raise ValueError('abscissa must increase strictly')
            self.scatter_abscissa = scatter
        self._theta_keys = (FAMILY_KEYS[self.family] +
                            assembias_keys('centrals', self.n_strength[0]) +
                            assembias_keys('satellites', self.n_strength[1]) +
                            tuple('scatter_model_param{}'.format(k + 1)
                                  for k in range(1, len(self.scatter_abscissa))))

    @property
    def latency_paths(self):
        """Whether the zero-copy one-draw / small-batch entries (seven parameters in the launch
        arguments or pinned columns) apply: fused family, one strength per galaxy type."""
        return self.family == FAMILY_ZHENG07 and not self.mass_dependent

    @property
    def occupation_keys(self):
        """The family's occupation parameters (without the assembly-bias strengths)."""
        return FAMILY_KEYS[self.family]

    @property
    def strength_keys(self):
        n = len(FAMILY_KEYS[self.family])
        return self._theta_keys[n:n + self.n_strength[0] + self.n_strength[1]]

    @property
    def scatter_keys(self):
        """``scatter_model_param2..n`` of a mass-dependent stellar-mass scatter (leauthaud11)."""
        return self._theta_keys[len(FAMILY_KEYS[self.family]) + sum(self.n_strength):]

    @property
    def theta_keys(self):
        """param_dict keys of the kernel's parameter vector, in kernel order."""
        return self._theta_keys

    @property
    def n_theta(self):
        return len(self.theta_keys)

    @property
    def n_base(self):
        return len(FAMILY_KEYS[self.family])

    def key(self):
        return (self.family, self.decorated, self.modulate_with_cenocc, self.split,
                self.threshold, self.redshift, self.strength_abscissa, self.split_abscissa,
                self.split_ordinates, self.scatter_abscissa)


def spec_from_params(params):
    """Family of a bare parameter dict (no model object): zheng07, decorated when the two
    assembly-bias strengths are present.  Other families need ``model=`` (they carry constants
    that are not in ``param_dict``)."""
    if all(k in params for k in ZHENG07_KEYS):
        return ModelSpec(FAMILY_ZHENG07, decorated=all(k in params for k in ASSEMBIAS_KEYS))
    if any(k in params for k in LEAUTHAUD11_KEYS):
        raise ValueError('leauthaud11 parameters need model= (a model instance or ModelSpec '
                         'carrying threshold and redshift)')
    return ModelSpec(FAMILY_ZHENG07, decorated=all(k in params for k in ASSEMBIAS_KEYS))


class Zheng07Model:
    """Parameter container with the attributes ``TabCorr.mean_occupation`` inspects.

    Looks like ``halotools.empirical_models.PrebuiltHodModelFactory('zheng07')`` as far as the
    prediction path is concerned: ``param_dict``, ``gal_types``, ``redshift`` and
    ``_input_model_dictionary[...].prim_haloprop_key`` (``tabcorr/tabcorr.py:496-535``).
    """

    tabcorr_b200_family = 0

    def __init__(self, threshold=-20, redshift=0.0, prim_haloprop_key='halo_mvir',
                 sec_haloprop_key='halo_nfw_conc', decorated=False, split=0.5,
                 modulate_with_cenocc=False, assembias_strength=0.5,
                 assembias_strength_abscissa=None, split_abscissa=None, **ignored):
        """``assembias_strength`` / ``assembias_strength_abscissa`` and ``split`` /
        ``split_abscissa`` follow halotools' ``HeavisideAssembias`` keywords: lists make the
        strength (``mean_occupation_*_assembias_param1..n``) or the splitting percentile a
        function of log10 of the primary halo property, for centrals and satellites alike."""
        try:
            values = ZHENG07_PUBLISHED[float(threshold)]
        except KeyError:
            raise KeyError('no published zheng07 parameters for threshold {}'.format(threshold))
        self.param_dict = dict(zip(ZHENG07_KEYS, values))
        self.decorated = bool(decorated)
        self.modulate_with_cenocc = bool(modulate_with_cenocc)
        strengths = np.atleast_1d(np.asarray(assembias_strength, dtype=np.float64))
        knots = _knots(assembias_strength_abscissa, 'assembias_strength_abscissa')
        if len(strengths) > 1 and len(knots) != len(strengths):
            raise ValueError('assembias_strength_abscissa must match assembias_strength in length')
        if len(knots) <= 1 or len(strengths) <= 1:
            knots = ()
        self.strength_abscissa = (knots, knots)
        split_values = _knots(split, 'split')
        split_knots = _knots(split_abscissa, 'split_abscissa')
        if len(split_values) > 1 or len(split_knots) > 0:
            if len(split_knots) != len(split_values):
                raise ValueError('split_abscissa must match split in length')
            self.split = 0.5
            self.split_abscissa = (split_knots, split_knots)
            self.split_ordinates = (split_values, split_values)
        else:
            self.split = float(split_values[0])
            self.split_abscissa = ((), ())
            self.split_ordinates = ((), ())
        if self.decorated:
            for gal_type in ('centrals', 'satellites'):
                for key, value in zip(assembias_keys(gal_type, max(1, len(knots))),
                                      np.broadcast_to(strengths, max(1, len(knots)))):
                    self.param_dict[key] = float(value)
        self.threshold = threshold
        self.redshift = redshift
        self.gal_types = ['centrals', 'satellites']
        cens = SimpleNamespace(prim_haloprop_key=prim_haloprop_key)
        sats = SimpleNamespace(prim_haloprop_key=prim_haloprop_key,
                               modulate_with_cenocc=self.modulate_with_cenocc)
        if self.decorated:
            cens.sec_haloprop_key = sec_haloprop_key
            sats.sec_haloprop_key = sec_haloprop_key
        self._input_model_dictionary = {'centrals_occupation': cens,
                                        'satellites_occupation': sats}


class Leauthaud11Model:
    """Stand-in for ``PrebuiltHodModelFactory('leauthaud11')`` / ``('hearin15')`` as far as the
    prediction path is concerned (``tabcorr/tabcorr.py:496-563``): stellar-mass threshold sample
    over the Behroozi et al. (2010) stellar-to-halo-mass relation."""

    tabcorr_b200_family = FAMILY_LEAUTHAUD11

    def __init__(self, threshold=10.5, redshift=0.0, prim_haloprop_key='halo_mvir',
                 sec_haloprop_key='halo_nfw_conc', decorated=False, split=0.5,
                 modulate_with_cenocc=True, central_assembias_strength=1.0,
                 satellite_assembias_strength=0.2, assembias_strength_abscissa=None,
                 split_abscissa=None, scatter_abscissa=None, scatter_ordinates=None, **ignored):
        """``central_assembias_strength`` / ``satellite_assembias_strength`` with
        ``assembias_strength_abscissa`` and ``split`` / ``split_abscissa`` follow halotools'
        ``HeavisideAssembias`` keywords: lists make a strength
        (``mean_occupation_*_assembias_param1..n``) or the splitting percentile a function of
        log10 of the primary halo property (the same abscissa for centrals and satellites)."""
        self.param_dict = dict(zip(LEAUTHAUD11_KEYS, LEAUTHAUD11_DEFAULTS))
        # halotools LogNormalScatterModel: scatter_abscissa / scatter_ordinates make the stellar-mass
        # scatter a function of log10 of the primary halo property (scatter_model_param1..n)
        scatter_knots = _knots(scatter_abscissa, 'scatter_abscissa')
        scatter_values = _knots(scatter_ordinates, 'scatter_ordinates')
        if len(scatter_knots) != len(scatter_values):
            raise ValueError('scatter_abscissa must match scatter_ordinates in length')
        self.scatter_abscissa = scatter_knots if len(scatter_knots) > 1 else ()
        for k, value in enumerate(scatter_values):
            self.param_dict['scatter_model_param{}'.format(k + 1)] = float(value)
        self.decorated = bool(decorated)
        self.modulate_with_cenocc = bool(modulate_with_cenocc)
        strengths = [np.atleast_1d(np.asarray(v, dtype=np.float64))
                     for v in (central_assembias_strength, satellite_assembias_strength)]
        knots = _knots(assembias_strength_abscissa, 'assembias_strength_abscissa')
        if any(len(v) > 1 and len(v) != len(knots) for v in strengths):
            raise ValueError('assembias_strength_abscissa must match the strengths in length')
        if len(knots) <= 1 or all(len(v) <= 1 for v in strengths):
            knots = ()
        self.strength_abscissa = (knots, knots)
        split_values = _knots(split, 'split')
        split_knots = _knots(split_abscissa, 'split_abscissa')
        if len(split_values) > 1 or len(split_knots) > 0:
            if len(split_knots) != len(split_values):
                raise ValueError('split_abscissa must match split in length')
            self.split = 0.5
            self.split_abscissa = (split_knots, split_knots)
            self.split_ordinates = (split_values, split_values)
        else:
            self.split = float(split_values[0])
            self.split_abscissa = ((), ())
            self.split_ordinates = ((), ())
        if self.decorated:
            for gal_type, values in zip(('centrals', 'satellites'), strengths):
                for key, value in zip(assembias_keys(gal_type, max(1, len(knots))),
                                      np.broadcast_to(values, max(1, len(knots)))):
                    self.param_dict[key] = float(value)
        self.threshold = float(threshold)
        self.redshift = float(redshift)
        self.gal_types = ['centrals', 'satellites']
        cens = SimpleNamespace(prim_haloprop_key=prim_haloprop_key, threshold=self.threshold,
                               redshift=self.redshift)
        sats = SimpleNamespace(prim_haloprop_key=prim_haloprop_key, threshold=self.threshold,
                               redshift=self.redshift,
                               modulate_with_cenocc=self.modulate_with_cenocc)
        if self.decorated:
            cens.sec_haloprop_key = sec_haloprop_key
            sats.sec_haloprop_key = sec_haloprop_key
        self._input_model_dictionary = {'centrals_occupation': cens,
                                        'satellites_occupation': sats}


def PrebuiltHodModelFactory(model_nickname, **kwargs):
    """Stand-in for ``halotools.empirical_models.PrebuiltHodModelFactory`` (README.md:47,
    tests/conftest.py:29-35) for the model families the occupation kernel implements."""
    name = model_nickname.lower()
    if name == 'zheng07':
        return Zheng07Model(**kwargs)
    if name in ('hearin15-zheng07', 'decorated-zheng07', 'zheng07-decorated'):
        return Zheng07Model(decorated=True, **kwargs)
    if name == 'leauthaud11':
        return Leauthaud11Model(**kwargs)
    if name == 'hearin15':
        return Leauthaud11Model(decorated=True, **kwargs)
    raise NotImplementedError(
        "model '{}' is not implemented by the occupation kernel (available: 'zheng07', "
        "'decorated-zheng07', 'leauthaud11', 'hearin15')".format(model_nickname))


def _scatter_knots(component):
    """``scatter_abscissa`` of a halotools ``Leauthaud11Cens``: its stellar-to-halo-mass model
    (``smhm_model``, a ``Behroozi10SmHm``) keeps a ``LogNormalScatterModel`` (``scatter_model``)
    whose ``abscissa`` are the log10 primary-property positions of ``scatter_model_param1..n``.
    One ordinate is the (default) constant scatter."""
    extra = [k for k in getattr(component, 'param_dict', {})
             if k.startswith('scatter_model_param') and k != 'scatter_model_param1']
    if not extra:
        return ()
    scatter_model = getattr(getattr(component, 'smhm_model', None), 'scatter_model', None)
    abscissa = getattr(scatter_model, 'abscissa', None)
    if abscissa is None or len(np.atleast_1d(abscissa)) != len(extra) + 1:
        raise NotImplementedError('mass-dependent stellar-mass scatter: cannot find the '
                                  'scatter_abscissa of the model (smhm_model.scatter_model.abscissa)')
    return _knots(abscissa, 'scatter_abscissa')


def _constant_split(component):
    """Splitting percentile of a halotools HeavisideAssembias component, if constant."""
    ordinates = getattr(component, '_split_ordinates', None)
    if ordinates is None:
        return 0.5
    ordinates = np.unique(np.asarray(ordinates, dtype=float))
    if len(ordinates) != 1:
        raise NotImplementedError('mass-dependent assembly-bias splits are not implemented for '
                                  'this family')
    return float(ordinates[0])


def _decoration_knots(component):
    """``(strength_abscissa, split_abscissa, split_ordinates)`` of a halotools
    ``HeavisideAssembias`` component: its ``_assembias_strength_abscissa`` (one entry: a constant
    strength) and ``_split_abscissa`` / ``_split_ordinates`` (all ordinates equal: a constant)."""
    strength = _knots(getattr(component, '_assembias_strength_abscissa', None),
                      'assembias_strength_abscissa')
    if len(strength) <= 1:
        strength = ()
    ordinates = getattr(component, '_split_ordinates', None)
    if ordinates is None:
        return strength, (), (0.5,)
    ordinates = _knots(ordinates, 'split')
    if len(set(ordinates)) == 1:
        return strength, (), (ordinates[0],)
    return strength, _knots(getattr(component, '_split_abscissa', None), 'split_abscissa'), ordinates


def resolve_model(model):
    """Return the :class:`ModelSpec` of a model object, or raise ``NotImplementedError``."""
    if isinstance(model, ModelSpec):
        return model
    if hasattr(model, 'tabcorr_b200_family'):
        return ModelSpec(model.tabcorr_b200_family, model.decorated, model.modulate_with_cenocc,
                         model.split, getattr(model, 'threshold', 0.0),
                         getattr(model, 'redshift', 0.0),
                         getattr(model, 'strength_abscissa', ((), ())),
                         getattr(model, 'split_abscissa', ((), ())),
                         getattr(model, 'split_ordinates', ((), ())),
                         getattr(model, 'scatter_abscissa', ()))
    components = getattr(model, '_input_model_dictionary', None)
    if components is not None and 'centrals_occupation' in components:
        cens = components['centrals_occupation']
        sats = components['satellites_occupation']
        names = (type(cens).__name__, type(sats).__name__)
        if names == ('Zheng07Cens', 'Zheng07Sats'):
            return ModelSpec(0, False, getattr(sats, 'modulate_with_cenocc', False), 0.5)
        if names == ('AssembiasZheng07Cens', 'AssembiasZheng07Sats'):
            knots = [_decoration_knots(c) for c in (cens, sats)]
            strength_abscissa = tuple(k[0] for k in knots)
            if all(len(k[1]) == 0 for k in knots) and knots[0][2] == knots[1][2]:
                # one constant split for both galaxy types
                return ModelSpec(0, True, getattr(sats, 'modulate_with_cenocc', False),
                                 knots[0][2][0], strength_abscissa=strength_abscissa)
            # per-type and / or mass-dependent splits: a constant is a single control point
            split_abscissa = tuple(k[1] if len(k[1]) else (0.0,) for k in knots)
            split_ordinates = tuple(k[2] for k in knots)
            return ModelSpec(0, True, getattr(sats, 'modulate_with_cenocc', False), 0.5,
                             strength_abscissa=strength_abscissa, split_abscissa=split_abscissa,
                             split_ordinates=split_ordinates)
        if names in (('Leauthaud11Cens', 'Leauthaud11Sats'),
                     ('AssembiasLeauthaud11Cens', 'AssembiasLeauthaud11Sats')):
            decorated = names[0].startswith('Assembias')
            scatter = _scatter_knots(cens)
            if float(sats.threshold) != float(cens.threshold):
                raise NotImplementedError('different thresholds for centrals and satellites')
            common = (FAMILY_LEAUTHAUD11, decorated, getattr(sats, 'modulate_with_cenocc', True))
            place = (float(cens.threshold), float(getattr(cens, 'redshift', 0.0)))
            if not decorated:
                return ModelSpec(*common, 0.5, *place, scatter_abscissa=scatter)
            knots = [_decoration_knots(c) for c in (cens, sats)]
            strength_abscissa = tuple(k[0] for k in knots)
            if all(len(k[1]) == 0 for k in knots) and knots[0][2] == knots[1][2]:
                # one constant split for both galaxy types
                return ModelSpec(*common, knots[0][2][0], *place,
                                 strength_abscissa=strength_abscissa, scatter_abscissa=scatter)
            # per-type and / or mass-dependent splits: a constant is a single control point
            return ModelSpec(*common, 0.5, *place, strength_abscissa=strength_abscissa,
                             split_abscissa=tuple(k[1] if len(k[1]) else (0.0,) for k in knots),
                             split_ordinates=tuple(k[2] for k in knots), scatter_abscissa=scatter)
        raise NotImplementedError(
            'occupation components {} are not implemented by the CUDA occupation kernel; '
            'pass precomputed occupations as an ndarray instead'.format(names))
    param_dict = getattr(model, 'param_dict', None)
    if param_dict is not None and all(k in param_dict for k in ZHENG07_KEYS):
        return ModelSpec(0, getattr(model, 'decorated', False),
                         getattr(model, 'modulate_with_cenocc', False),
                         getattr(model, 'split', 0.5),
                         strength_abscissa=getattr(model, 'strength_abscissa', ((), ())),
                         split_abscissa=getattr(model, 'split_abscissa', ((), ())),
                         split_ordinates=getattr(model, 'split_ordinates', ((), ())))
    if param_dict is not None and all(k in param_dict for k in LEAUTHAUD11_KEYS):
        if not hasattr(model, 'threshold'):
            raise NotImplementedError('a leauthaud11 model needs a `threshold` attribute')
        return ModelSpec(FAMILY_LEAUTHAUD11, getattr(model, 'decorated', False),
                         getattr(model, 'modulate_with_cenocc', True),
                         getattr(model, 'split', 0.5), float(model.threshold),
                         float(getattr(model, 'redshift', 0.0)),
                         strength_abscissa=getattr(model, 'strength_abscissa', ((), ())),
                         split_abscissa=getattr(model, 'split_abscissa', ((), ())),
                         split_ordinates=getattr(model, 'split_ordinates', ((), ())),
                         scatter_abscissa=getattr(model, 'scatter_abscissa', ()))
    raise NotImplementedError(
        'cannot map {!r} to a kernel occupation family (zheng07, leauthaud11 and their '
        'decorated versions)'.format(model))


def theta_columns(params, spec=None):
    """The columns (arrays ``[B]`` or scalars) of a parameter dict in the kernel order of the
    family (``spec.theta_keys``); raises ``ValueError`` for missing parameters like
    :func:`theta_from_params`."""
    if spec is None:
        spec = ModelSpec()
    missing = [k for k in spec.occupation_keys if k not in params]
    if missing:
        raise ValueError('missing occupation parameters: {}'.format(', '.join(missing)))
    if spec.decorated:
        missing = [k for k in spec.strength_keys if k not in params]
        if missing:
            raise ValueError('missing assembly-bias parameters: {}'.format(', '.join(missing)))
    missing = [k for k in spec.scatter_keys if k not in params]
    if missing:
        raise ValueError('missing scatter parameters: {}'.format(', '.join(missing)))
    return [np.asarray(params.get(k, 0.0), dtype=np.float64) for k in spec.theta_keys]


def theta_from_params(params, n_draws=None, spec=None, alloc=None):
    """``[B, n_theta]`` float64 array in kernel order from a dict of scalars/arrays keyed by
    halotools parameter names (``n_theta`` = 7 for zheng07, 18 for leauthaud11).  Missing
    assembly-bias strengths default to 0.  ``alloc(shape)`` may supply the output buffer (e.g.
    pinned host memory)."""
    columns = theta_columns(params, spec)
    if n_draws is None:
        lengths = [c.shape[0] for c in columns if c.ndim > 0]
        n_draws = max(lengths) if lengths else 1     # 0 for a batch of empty arrays
    shape = (n_draws, len(columns))
    theta = np.empty(shape, dtype=np.float64) if alloc is None else alloc(shape)
    for j, column in enumerate(columns):
        theta[:, j] = column
    return theta
