"""Occupation models understood by the occupation kernel.

The reference evaluates ``model.mean_occupation_centrals/satellites`` of a halotools
``HodModelFactory`` on the CPU (``tabcorr/tabcorr.py:556-563``).  Here the model object only
*describes* the occupation functions: the arithmetic runs in the CUDA occupation kernel
(``csrc/tabcorr_b200.cu``).  ``resolve_model`` maps

* a halotools model (recognised by the class names of its occupation components -- halotools
  itself is never imported),
* the light-weight stand-ins below (``PrebuiltHodModelFactory('zheng07' | 'hearin15-zheng07')``),
* or any duck-typed object with a zheng07 ``param_dict`` (plus optional ``decorated`` / ``split`` /
  ``modulate_with_cenocc`` attributes)

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


class ModelSpec:
    """What the kernel needs to know about a model (mirrors ``tc_model`` in the C ABI)."""

    def __init__(self, family=0, decorated=False, modulate_with_cenocc=False, split=0.5):
        self.family = int(family)
        self.decorated = bool(decorated)
        self.modulate_with_cenocc = bool(modulate_with_cenocc)
        self.split = float(split)

    def key(self):
        return (self.family, self.decorated, self.modulate_with_cenocc, self.split)


class Zheng07Model:
    """Parameter container with the attributes ``TabCorr.mean_occupation`` inspects.

    Looks like ``halotools.empirical_models.PrebuiltHodModelFactory('zheng07')`` as far as the
    prediction path is concerned: ``param_dict``, ``gal_types``, ``redshift`` and
    ``_input_model_dictionary[...].prim_haloprop_key`` (``tabcorr/tabcorr.py:496-535``).
    """

    tabcorr_b200_family = 0

    def __init__(self, threshold=-20, redshift=0.0, prim_haloprop_key='halo_mvir',
                 sec_haloprop_key='halo_nfw_conc', decorated=False, split=0.5,
                 modulate_with_cenocc=False, **ignored):
        try:
            values = ZHENG07_PUBLISHED[float(threshold)]
        except KeyError:
            raise KeyError('no published zheng07 parameters for threshold {}'.format(threshold))
        self.param_dict = dict(zip(ZHENG07_KEYS, values))
        self.decorated = bool(decorated)
        self.split = float(split)
        self.modulate_with_cenocc = bool(modulate_with_cenocc)
        if self.decorated:
            for key in ASSEMBIAS_KEYS:
                self.param_dict[key] = 0.5
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


def PrebuiltHodModelFactory(model_nickname, **kwargs):
    """Stand-in for ``halotools.empirical_models.PrebuiltHodModelFactory`` (README.md:47,
    tests/conftest.py:29-35) for the model families the occupation kernel implements."""
    name = model_nickname.lower()
    if name == 'zheng07':
        return Zheng07Model(**kwargs)
    if name in ('hearin15-zheng07', 'decorated-zheng07', 'zheng07-decorated'):
        return Zheng07Model(decorated=True, **kwargs)
    raise NotImplementedError(
        "model '{}' is not implemented by the occupation kernel (available: 'zheng07', "
        "'decorated-zheng07')".format(model_nickname))


def _constant_split(component):
    """Splitting percentile of a halotools HeavisideAssembias component, if constant."""
    ordinates = getattr(component, '_split_ordinates', None)
    if ordinates is None:
        return 0.5
    ordinates = np.unique(np.asarray(ordinates, dtype=float))
    if len(ordinates) != 1:
        raise NotImplementedError('mass-dependent assembly-bias splits are not implemented')
    return float(ordinates[0])


def resolve_model(model):
    """Return the :class:`ModelSpec` of a model object, or raise ``NotImplementedError``."""
    if isinstance(model, ModelSpec):
        return model
    if hasattr(model, 'tabcorr_b200_family'):
        return ModelSpec(model.tabcorr_b200_family, model.decorated, model.modulate_with_cenocc,
                         model.split)
    components = getattr(model, '_input_model_dictionary', None)
    if components is not None and 'centrals_occupation' in components:
        cens = components['centrals_occupation']
        sats = components['satellites_occupation']
        names = (type(cens).__name__, type(sats).__name__)
        if names == ('Zheng07Cens', 'Zheng07Sats'):
            return ModelSpec(0, False, getattr(sats, 'modulate_with_cenocc', False), 0.5)
        if names == ('AssembiasZheng07Cens', 'AssembiasZheng07Sats'):
            split = _constant_split(cens)
            if _constant_split(sats) != split:
                raise NotImplementedError('different splits for centrals and satellites')
            return ModelSpec(0, True, getattr(sats, 'modulate_with_cenocc', False), split)
        raise NotImplementedError(
            'occupation components {} are not implemented by the CUDA occupation kernel; '
            'pass precomputed occupations as an ndarray instead'.format(names))
    param_dict = getattr(model, 'param_dict', None)
    if param_dict is not None and all(k in param_dict for k in ZHENG07_KEYS):
        return ModelSpec(0, getattr(model, 'decorated', False),
                         getattr(model, 'modulate_with_cenocc', False),
                         getattr(model, 'split', 0.5))
    raise NotImplementedError(
        'cannot map {!r} to a kernel occupation family (zheng07, decorated zheng07)'.format(model))


def theta_columns(params, spec=None):
    """The ``TC_N_THETA`` columns (arrays ``[B]`` or scalars) of a parameter dict in kernel order;
    raises ``ValueError`` for missing parameters like :func:`theta_from_params`."""
    missing = [k for k in ZHENG07_KEYS if k not in params]
    if missing:
        raise ValueError('missing occupation parameters: {}'.format(', '.join(missing)))
    if spec is not None and spec.decorated:
        missing = [k for k in ASSEMBIAS_KEYS if k not in params]
        if missing:
            raise ValueError('missing assembly-bias parameters: {}'.format(', '.join(missing)))
    return [np.asarray(params.get(k, 0.0), dtype=np.float64) for k in THETA_KEYS]


def theta_from_params(params, n_draws=None, spec=None, alloc=None):
    """``[B, 7]`` float64 array in kernel order from a dict of scalars/arrays keyed by halotools
    parameter names.  Missing assembly-bias strengths default to 0.  ``alloc(shape)`` may supply
    the output buffer (e.g. pinned host memory)."""
    columns = theta_columns(params, spec)
    if n_draws is None:
        n_draws = max([c.shape[0] for c in columns if c.ndim > 0] + [1])
    shape = (n_draws, len(THETA_KEYS))
    theta = np.empty(shape, dtype=np.float64) if alloc is None else alloc(shape)
    for j, column in enumerate(columns):
        theta[:, j] = column
    return theta
