"""Database access for the prediction path: path logic + ``Interpolator.read``.

Mirrors the read half of ``tabcorr/database.py`` (``simulation_name`` :161-210, ``directory``
:213-250, ``read`` :253-286 and its alias ``tabcorr`` :290) and the binning presets of
``configuration`` (:14-71) that the reference's tests use to locate the interpolation knots.  The
cosmology half (:74-158) needs astropy and is only metadata for tabulation; it is not rebuilt.
"""

import os
from pathlib import Path

import numpy as np

from .interpolator import Interpolator


def configuration(config_str):
    """Binning presets of a tabulation configuration string (``tabcorr/database.py:14-71``).

    Several configurations are joined with ``_``; earlier ones take precedence and ``default``
    fills what is left.  ``cosmo_obs`` is the astropy ``Planck15`` cosmology when astropy is
    importable and ``None`` otherwise (it is not used by predictions).
    """
    names = config_str.split('_')
    for name in names:
        if name not in ('aemulus', 'default', 'efficient'):
            raise ValueError('Unkown configuration {}.'.format(name))
    names.append('default')
    try:
        from astropy.cosmology import Planck15
    except ImportError:
        Planck15 = None
    presets = {
        's_bins': {'default': np.logspace(-1.0, 1.8, 15), 'aemulus': np.logspace(-1, 1.78, 10)},
        'rp_wp_bins': {'default': np.logspace(-1.0, 1.8, 15),
                       'aemulus': np.logspace(-1, 1.78, 10)},
        'pi_max': {'default': 80},
        'rp_ds_bins': {'default': np.logspace(-1.0, 1.8, 15),
                       'efficient': np.logspace(-1.0, 1.6, 14)},
        'mu_bins': {'default': np.linspace(0, 1, 21), 'aemulus': np.linspace(0, 1, 41)},
        'cosmo_obs': {'default': Planck15, 'aemulus': None},
        'alpha_c_bins': {'default': np.linspace(0.0, 0.4, 4)},
        'alpha_s_bins': {'default': np.linspace(0.8, 1.2, 4)},
        'conc_gal_bias_bins': {'default': np.geomspace(1.0 / 3.0, 3.0, 4)},
        'sats_per_prim_haloprop': {'default': 2e-13, 'efficient': 1e-13},
        'downsample': {'default': 1.0, 'efficient': (lambda x: x / 1e13)},
    }
    config = {}
    for parameter, options in presets.items():
        for name in names:
            if name in options:
                config[parameter] = options[name]
                break
    return config


def cosmology(suite, i_cosmo=0):
    raise NotImplementedError(
        'simulation cosmologies (tabcorr/database.py:74-158) are tabulation metadata that need '
        'astropy; they are outside the accelerated prediction path')


def simulation_name(suite, i_cosmo=0, i_phase=0, config=None):
    """Name of a simulation (``tabcorr/database.py:161-210``)."""
    if suite == 'AbacusSummit':
        return '{}_c{:03d}_ph{:03d}'.format('base' if config is None else config, i_cosmo, i_phase)
    if suite == 'AemulusAlpha':
        if 0 <= i_cosmo < 40:
            return 'Box{:03d}'.format(i_cosmo)
        if 0 <= i_cosmo < 47:
            if i_phase > 6:
                raise ValueError('Unknown phase number {}.'.format(i_phase))
            return 'TestBox{:03d}-{:03d}'.format(i_cosmo - 40, i_phase)
        raise ValueError('Unknown cosmology number {}. '.format(i_cosmo) +
                         'Must be in the range from 0 to 46.')
    raise ValueError('Unkown simulation suite {}.'.format(suite))


def directory(suite, redshift, i_cosmo=0, i_phase=0, config=None):
    """Directory holding all data of a simulation snapshot (``tabcorr/database.py:213-250``)."""
    try:
        root = Path(os.environ['TABCORR_DATABASE'])
    except KeyError:
        raise RuntimeError('You must set the TABCORR_DATABASE environment variable.')
    name = simulation_name(suite, i_cosmo=i_cosmo, i_phase=i_phase, config=config)
    return root / suite / name / '{:.2f}'.format(redshift).replace('.', 'p')


def read(suite, redshift, tpcf, i_cosmo=0, i_phase=0, sim_config=None, tab_config='default',
         device=None):
    """Read the tabulation of a simulation snapshot as a device-resident ``Interpolator``
    (``tabcorr/database.py:253-286``)."""
    path = directory(suite, redshift, i_cosmo=i_cosmo, i_phase=i_phase, config=sim_config)
    return Interpolator.read(path / '{}_{}.hdf5'.format(tpcf, tab_config), device=device)


def read_set(suite, redshift, tpcf, i_cosmo=(0,), i_phase=0, sim_config=None,
             tab_config='default', device=None):
    """One ``Interpolator`` per cosmology as a :class:`~tabcorr_b200.tableset.TableSet`, for
    batches whose draws each name a cosmology (``table_index`` = position in ``i_cosmo``).  Not in
    the reference, where a sampler would call ``read`` once per cosmology
    (``tabcorr/database.py:253-286``)."""
    from .tableset import TableSet
    return TableSet([read(suite, redshift, tpcf, i_cosmo=int(c), i_phase=i_phase,
                          sim_config=sim_config, tab_config=tab_config, device=device)
                     for c in i_cosmo])


# alias kept for backwards compatibility, as in the reference (tabcorr/database.py:290)
tabcorr = read
