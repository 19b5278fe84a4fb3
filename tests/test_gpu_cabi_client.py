"""The C ABI driven from plain C (tests/c_client/predict_client.c, compiled with gcc against
include/tabcorr_b200.h and linked to the in-tree library): same inputs as the Python API, results
compared bit for bit.  Shows that the boundary carries no Python or torch types."""

import os
import shutil
import subprocess

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get('CUDA_HOME', '/usr/local/cuda')


@pytest.fixture(scope='module')
def client(tmp_path_factory):
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    if shutil.which('gcc') is None or not os.path.isdir(os.path.join(CUDA, 'include')):
        pytest.skip('gcc or the CUDA runtime headers are not available')
    from tabcorr_b200 import _lib
    _lib.load()
    exe = str(tmp_path_factory.mktemp('c_client') / 'predict_client')
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(
        ['gcc', '-O1', '-std=c99', '-Wall', '-Werror', '-o', exe,
         os.path.join(ROOT, 'tests', 'c_client', 'predict_client.c'),
         '-I', os.path.join(ROOT, 'include'), '-I', os.path.join(CUDA, 'include'),
         '-L', lib_dir, '-l:libtabcorr_b200.so', '-L', os.path.join(CUDA, 'lib64'), '-lcudart',
         '-Wl,-rpath,' + lib_dir, '-Wl,-rpath,' + os.path.join(CUDA, 'lib64')],
        check=True)
    return exe


@pytest.mark.parametrize('name,separate', [('syn36x3', 0), ('syn240dec', 1), ('syncross', 1)])
def test_c_client_matches_python_api(client, tmp_path, name, separate):
    import tabcorr_b200 as tb
    from tabcorr_b200.tabcorr import leggauss01
    kw, _ = cases.SYNTHETIC[name]
    tab = tb.synthetic.make_table(**kw)
    gal_type = tab['gal_type']
    n_rows, mode = len(gal_type), tab['attrs']['mode']
    n_r = int(np.prod(tab['tpcf_shape']))
    n_gauss, n_draws = 10, 777
    draws = tb.synthetic.make_draws(n_draws, seed=6, decorated=True)
    theta = tb.models.theta_from_params(draws, None, tb.models.ModelSpec(decorated=True))
    x01, w = leggauss01(n_gauss)
    names = np.asarray(gal_type['gal_type'])
    if names.dtype.kind == 'S':
        names = np.char.decode(names, 'utf-8')
    with open(tmp_path / 'in.bin', 'wb') as f:
        np.array([0 if mode == 'auto' else 1, n_rows, n_r, n_gauss, n_draws, separate],
                 dtype=np.int32).tofile(f)
        for col in ('n_h', 'log_prim_haloprop_min', 'log_prim_haloprop_max',
                    'sec_haloprop_percentile', 'prim_haloprop_dist_index'):
            np.ascontiguousarray(gal_type[col], dtype=np.float64).tofile(f)
        (names != 'centrals').astype(np.int32).tofile(f)
        np.ascontiguousarray(tab['tpcf_matrix'], dtype=np.float64).tofile(f)
        x01.tofile(f)
        w.tofile(f)
        np.ascontiguousarray(theta).tofile(f)
    run = subprocess.run([client, str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')],
                         capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    assert run.stdout.startswith('ok: 777 draws')
    out = np.fromfile(tmp_path / 'out.bin', dtype=np.float64)
    n_ng = 2 if separate else 1
    n_comp = 1 if not separate else (3 if mode == 'auto' else 2)
    ngal_c = out[:n_draws * n_ng].reshape(n_draws, n_ng)
    xi_c = out[n_draws * n_ng:].reshape(n_draws, n_r, n_comp)
    halotab = tb.TabCorr.from_arrays(gal_type, tab['tpcf_matrix'], tab['tpcf_shape'], tab['attrs'])
    ngal, xi = halotab.predict_batch(draws, separate_gal_type=bool(separate), pipeline_chunk=0,
                                     as_numpy=False)
    if separate:
        ngal_keys, xi_keys = halotab._separate_keys()
        for j, key in ngal_keys:
            assert np.array_equal(ngal[key].cpu().numpy(), ngal_c[:, j])
        for j, key in xi_keys:
            assert np.array_equal(xi[key].cpu().numpy().reshape(n_draws, n_r), xi_c[:, :, j])
    else:
        assert np.array_equal(ngal.cpu().numpy(), ngal_c[:, 0])
        assert np.array_equal(xi.cpu().numpy().reshape(n_draws, n_r), xi_c[:, :, 0])


def test_integration_md_stub_runs():
    """The ctypes stub printed in INTEGRATION.md (what a maintainer of the reference would add)
    is executed as written -- only the library path is substituted -- and agrees with the
    package's own binding bit for bit."""
    import re
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import tabcorr_b200 as tb
    from tabcorr_b200 import _lib
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    block = re.search(r'```python\n(# tabcorr/_b200\.py.*?)```', text, flags=re.S).group(1)
    block = block.replace("'libtabcorr_b200.so'", repr(_lib.LIB_PATH))
    namespace = {}
    exec(compile(block, 'INTEGRATION.md', 'exec'), namespace)
    halotab = tb.TabCorr.read(os.path.join(ROOT, 'tests', 'golden', 'bolplanck_wp.hdf5'))
    handle = namespace['upload'](halotab)
    draws = tb.synthetic.make_draws(300, seed=8)
    theta = tb.models.theta_from_params(draws, None, tb.models.ModelSpec())
    ngal, xi = namespace['predict_batch'](handle, theta, 19)
    ngal_ref, xi_ref = halotab.predict_batch(draws)
    assert np.array_equal(ngal, ngal_ref) and np.array_equal(xi, xi_ref)
