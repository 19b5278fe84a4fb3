"""The C-ABI shared library loads on a CPU-only box and exports every symbol the header declares.
No compute call is made here: without a CUDA device every computing entry point must fail loudly
(TC_ECUDA), never fall back to the host."""

import ctypes
import os
import re

import numpy as np
import pytest

from tabcorr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'tabcorr_b200.h')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(tc_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
    if not os.path.isfile(_lib.LIB_PATH):
        from tabcorr_b200 import build
        build.build()
    return _lib.load()


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_every_declared_symbol_is_exported(lib):
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), name


def test_version_and_constants(lib):
    text = open(HEADER).read()
    assert lib.tc_version() == int(re.search(r'#define TC_VERSION (\d+)', text).group(1))
    assert int(re.search(r'#define TC_N_THETA (\d+)', text).group(1)) == _lib.TC_N_THETA
    assert int(re.search(r'#define TC_N_THETA_LEAUTHAUD11 (\d+)', text).group(1)) == \
        _lib.TC_N_THETA_LEAUTHAUD11
    # 4 x int32 + 3 doubles + the mass-dependent decoration block (2 x 2 int32 + 3 x 2 x 4 doubles)
    assert int(re.search(r'#define TC_MAX_KNOTS (\d+)', text).group(1)) == _lib.TC_MAX_KNOTS == 4
    assert ctypes.sizeof(_lib.tc_model) == 40 + 16 + 3 * 2 * 4 * 8 + 4 * 8
    from tabcorr_b200 import models
    for family, n_theta in ((_lib.TC_FAMILY_ZHENG07, _lib.TC_N_THETA),
                            (_lib.TC_FAMILY_LEAUTHAUD11, _lib.TC_N_THETA_LEAUTHAUD11)):
        spec = models.ModelSpec(family)
        model = _lib.tc_model(family, 0, 0, 0, 0.5, 10.5, 0.0)
        assert lib.tc_model_n_theta(ctypes.byref(model)) == n_theta == spec.n_theta
    assert lib.tc_model_n_theta(ctypes.byref(_lib.tc_model(9, 0, 0, 0, 0.5, 0.0, 0.0))) == -3
    # mass-dependent strengths: the draw grows by the extra ordinates; bad control points are refused
    from tabcorr_b200.tabcorr import DeviceTableGroup
    spec = models.ModelSpec(0, True, strength_abscissa=((11.0, 12.0, 13.0), (11.5, 13.5)))
    model = DeviceTableGroup._model_struct(spec)
    assert lib.tc_model_n_theta(ctypes.byref(model)) == spec.n_theta == 5 + 3 + 2
    model.strength_abscissa[0][1] = 10.0
    assert lib.tc_model_n_theta(ctypes.byref(model)) == -1 and b'increase' in lib.tc_last_error()
    model.n_strength[0] = 7
    assert lib.tc_model_n_theta(ctypes.byref(model)) == -3


def test_argument_errors_do_not_need_a_device(lib):
    handle = ctypes.c_void_p()
    # bad mode / NULL arrays are rejected before any CUDA call
    assert lib.tc_table_create(ctypes.byref(handle), 7, 4, 1, 1, None, None, None, None, None,
                               None, None, 0) == -1
    assert b'mode' in lib.tc_last_error()
    assert lib.tc_table_create(ctypes.byref(handle), 0, 4, 1, 1, None, None, None, None, None,
                               None, None, 0) == -1
    assert lib.tc_table_destroy(None) == 0
    assert lib.tc_interp_destroy(None) == 0


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    n = 4
    ones = np.ones(n)
    is_sat = np.array([0, 0, 1, 1], dtype=np.int32)
    matrix = np.ones((1, n * (n + 1) // 2))
    c_double_p = ctypes.POINTER(ctypes.c_double)
    mats = (c_double_p * 1)(_lib.as_double_p(matrix))
    handle = ctypes.c_void_p()
    status = lib.tc_table_create(
        ctypes.byref(handle), 0, n, 1, 1, _lib.as_double_p(ones), _lib.as_double_p(ones),
        _lib.as_double_p(ones * 2), _lib.as_double_p(ones * 0.5), None, _lib.as_int32_p(is_sat),
        mats, 0)
    assert status == -2  # TC_ECUDA
    assert not handle.value
    with pytest.raises(_lib.TabCorrB200Error):
        _lib.check(status)
    peak = ctypes.c_double()
    assert lib.tc_measure_dmma_peak(0, ctypes.byref(peak)) == -2


def test_python_api_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    import tabcorr_b200
    from tabcorr_b200 import synthetic
    tab = synthetic.make_table(n_mass=4, n_sec=1, n_r=3)
    halotab = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'],
                                               tab['tpcf_shape'], tab['attrs'])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        halotab.predict_batch(synthetic.make_draws(3))


def test_header_prototypes_match_the_binding(lib):
    """Every prototype in the header has as many parameters as the ctypes binding declares
    (a mismatch would corrupt the call without any error on x86-64)."""
    text = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    protos = re.findall(r'\b(tc_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', text, flags=re.S)
    assert len(protos) == len(declared_symbols())
    for name, params in protos:
        params = params.strip()
        n_params = 0 if params in ('', 'void') else len(params.split(','))
        assert len(getattr(lib, name).argtypes or []) == n_params, name
    assert _lib.TC_PRECISION_FP64 == int(re.search(r'#define TC_PRECISION_FP64 (\d+)', text).group(1))
    assert _lib.TC_PRECISION_3XTF32 == int(
        re.search(r'#define TC_PRECISION_3XTF32 (\d+)', text).group(1))
    assert _lib.precision_code('3xTF32') == 1 and _lib.precision_code('fp64') == 0
    with pytest.raises(ValueError):
        _lib.precision_code('bf16')


def test_header_is_valid_c99(tmp_path):
    """include/tabcorr_b200.h and the plain-C client compile as C99 with -Wall -Werror (the ABI is
    C, not C++); linking against the in-tree library resolves every symbol the client uses."""
    import shutil
    import subprocess
    cuda = os.environ.get('CUDA_HOME', '/usr/local/cuda')
    if shutil.which('gcc') is None or not os.path.isfile(os.path.join(cuda, 'include',
                                                                       'cuda_runtime_api.h')):
        pytest.skip('gcc or the CUDA runtime headers are not available')
    root = os.path.dirname(os.path.dirname(HEADER))
    source = os.path.join(root, 'tests', 'c_client', 'predict_client.c')
    obj = str(tmp_path / 'client.o')
    # the header alone is pedantic C99 (CUDA's own runtime headers are not)
    alone = tmp_path / 'header_only.c'
    alone.write_text('#include "tabcorr_b200.h"\nint main(void) { return TC_VERSION > 0 ? 0 : 1; }\n')
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-pedantic', '-c', str(alone), '-o',
                    str(tmp_path / 'header_only.o'), '-I', os.path.dirname(HEADER)], check=True)
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-c', source, '-o', obj,
                    '-I', os.path.dirname(HEADER), '-I', os.path.join(cuda, 'include')], check=True)
    subprocess.run(['gcc', '-o', str(tmp_path / 'client'), obj, '-L', os.path.dirname(_lib.LIB_PATH),
                    '-l:libtabcorr_b200.so', '-L', os.path.join(cuda, 'lib64'), '-lcudart'],
                   check=True)
