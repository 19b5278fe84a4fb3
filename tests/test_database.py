"""``tabcorr_b200.database``: the path logic and ``read`` of ``tabcorr/database.py:161-290``.

The expected names and paths are the reference's own (checked against the live reference source
where the checkout exists, ``-m reference``); the on-disk tree is laid out like the reference's
test database ``tests/AbacusSummit/base_c000_ph000/0p50/ds_efficient.hdf5`` with the shipped
fixture ``tests/golden/ds_efficient.hdf5``.  No GPU is needed: ``read`` only parses the file, the
tables are uploaded on first use.
"""

import os
import shutil
from pathlib import Path

import numpy as np
import pytest

import tabcorr_b200
from tabcorr_b200 import database


@pytest.fixture()
def mini_database(tmp_path, golden_dir, monkeypatch):
    target = tmp_path / 'AbacusSummit' / 'base_c000_ph000' / '0p50'
    target.mkdir(parents=True)
    shutil.copy(os.path.join(golden_dir, 'ds_efficient.hdf5'), target / 'ds_efficient.hdf5')
    monkeypatch.setenv('TABCORR_DATABASE', str(tmp_path))
    return tmp_path


# (suite, i_cosmo, i_phase, config) -> name, tabcorr/database.py:189-207
NAMES = [
    (('AbacusSummit', 0, 0, None), 'base_c000_ph000'),
    (('AbacusSummit', 130, 7, None), 'base_c130_ph007'),
    (('AbacusSummit', 4, 2, 'high'), 'high_c004_ph002'),
    (('AemulusAlpha', 0, 0, None), 'Box000'),
    (('AemulusAlpha', 39, 3, None), 'Box039'),       # the phase is ignored for training boxes
    (('AemulusAlpha', 40, 0, None), 'TestBox000-000'),
    (('AemulusAlpha', 46, 6, None), 'TestBox006-006'),
]


@pytest.mark.parametrize('args, expected', NAMES)
def test_simulation_name(args, expected):
    suite, i_cosmo, i_phase, config = args
    assert database.simulation_name(suite, i_cosmo=i_cosmo, i_phase=i_phase,
                                    config=config) == expected


def test_simulation_name_errors():
    with pytest.raises(ValueError, match='Unkown simulation suite'):      # :209-210 (sic)
        database.simulation_name('Millennium')
    with pytest.raises(ValueError, match='Unknown cosmology number 47'):  # :205-207
        database.simulation_name('AemulusAlpha', i_cosmo=47)
    with pytest.raises(ValueError, match='Unknown cosmology number -1'):
        database.simulation_name('AemulusAlpha', i_cosmo=-1)
    with pytest.raises(ValueError, match='Unknown phase number 7'):       # :200-202
        database.simulation_name('AemulusAlpha', i_cosmo=41, i_phase=7)


def test_directory_needs_the_environment_variable(monkeypatch):
    monkeypatch.delenv('TABCORR_DATABASE', raising=False)
    with pytest.raises(RuntimeError, match='TABCORR_DATABASE'):           # :242-246
        database.directory('AbacusSummit', 0.5)
    with pytest.raises(RuntimeError, match='TABCORR_DATABASE'):
        database.read('AbacusSummit', 0.5, 'ds')


def test_directory_layout(monkeypatch, tmp_path):
    monkeypatch.setenv('TABCORR_DATABASE', str(tmp_path))
    path = database.directory('AbacusSummit', 0.5)
    assert isinstance(path, Path)
    assert path == tmp_path / 'AbacusSummit' / 'base_c000_ph000' / '0p50'   # :247-250
    assert database.directory('AbacusSummit', 1.4, i_cosmo=3, i_phase=1, config='huge').parts[-3:] \
        == ('AbacusSummit', 'huge_c003_ph001', '1p40')
    # two decimals, like '{:.2f}'.format(redshift)
    assert database.directory('AemulusAlpha', 0.249, i_cosmo=42, i_phase=1).parts[-2:] == \
        ('TestBox002-001', '0p25')


def test_read_finds_and_parses_the_reference_test_database(mini_database, golden):
    """The reference's own fixture call: tests/conftest.py:13-19."""
    interp = database.read('AbacusSummit', 0.5, 'ds', tab_config='efficient', i_cosmo=0)
    assert isinstance(interp, tabcorr_b200.Interpolator)
    assert len(interp.tabcorr_list) == 4
    # __init__ appends the sort column (tabcorr/interpolator.py:59-61)
    assert interp.param_dict_table.colnames == ['log_eta', 'tabcorr_index']
    np.testing.assert_allclose(np.asarray(interp.param_dict_table['log_eta']),
                               np.log10(np.geomspace(1 / 3, 3, 4)), rtol=1e-12)
    first = interp.tabcorr_list[0]
    assert first.attrs['mode'] == 'cross' and first.tpcf_shape == (13,)
    assert len(first.gal_type) == 1104 and first.tpcf_matrix.shape == (13, 1104)
    assert first.tpcf_matrix.dtype == np.float64                           # tabcorr.py:399
    # the file that database.read opened is the fixture (same bytes -> same tables)
    direct = tabcorr_b200.Interpolator.read(os.path.join(mini_database, 'AbacusSummit',
                                                         'base_c000_ph000', '0p50',
                                                         'ds_efficient.hdf5'))
    for a, b in zip(interp.tabcorr_list, direct.tabcorr_list):
        assert np.array_equal(a.tpcf_matrix, b.tpcf_matrix)


def test_read_missing_file_propagates_oserror(mini_database):
    with pytest.raises(OSError):
        database.read('AbacusSummit', 0.5, 'wp', tab_config='efficient')   # wp_efficient is absent
    with pytest.raises(OSError):
        database.read('AbacusSummit', 0.5, 'ds', tab_config='efficient', i_cosmo=1)


def test_tabcorr_alias_is_read():
    assert database.tabcorr is database.read                               # :289-290


def test_configuration_presets():
    config = database.configuration('efficient')                           # :37-71
    np.testing.assert_allclose(config['rp_ds_bins'], np.logspace(-1.0, 1.6, 14))
    np.testing.assert_allclose(config['rp_wp_bins'], np.logspace(-1.0, 1.8, 15))   # default fills
    np.testing.assert_allclose(config['conc_gal_bias_bins'], np.geomspace(1 / 3, 3, 4))
    assert config['sats_per_prim_haloprop'] == 1e-13 and config['pi_max'] == 80
    assert database.configuration('aemulus_efficient')['s_bins'].shape == (10,)    # first wins
    assert database.configuration('efficient_aemulus')['sats_per_prim_haloprop'] == 1e-13
    assert database.configuration('default')['downsample'] == 1.0
    with pytest.raises(ValueError, match='Unkown configuration'):
        database.configuration('fast')


def test_out_of_scope_half_raises():
    with pytest.raises(NotImplementedError):
        database.cosmology('AbacusSummit')


@pytest.mark.reference
def test_names_and_paths_against_the_live_reference(monkeypatch, tmp_path):
    from oracle import refstub
    if not refstub.available():
        pytest.skip('reference checkout not present')
    import importlib
    refstub.load()
    try:
        ref_db = importlib.import_module('tabcorr.database')
    except Exception as err:   # the cosmology half needs astropy at import time
        pytest.skip('reference database module does not import here: {}'.format(err))
    monkeypatch.setenv('TABCORR_DATABASE', str(tmp_path))
    for (suite, i_cosmo, i_phase, config), _ in NAMES:
        assert database.simulation_name(suite, i_cosmo, i_phase, config) == \
            ref_db.simulation_name(suite, i_cosmo, i_phase, config)
        assert database.directory(suite, 0.5, i_cosmo, i_phase, config) == \
            ref_db.directory(suite, 0.5, i_cosmo, i_phase, config)
    for name in ('default', 'efficient', 'aemulus', 'aemulus_efficient'):
        ours, theirs = database.configuration(name), ref_db.configuration(name)
        assert set(ours) == set(theirs)
        for key in ours:
            if key in ('cosmo_obs', 'downsample'):
                continue
            np.testing.assert_array_equal(ours[key], theirs[key])
