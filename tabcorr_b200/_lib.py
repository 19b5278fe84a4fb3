"""ctypes binding of the C ABI declared in ``include/tabcorr_b200.h``.

The shared library is built in-tree by ``__graft_entry__.build()`` (or ``python -m
tabcorr_b200.build``).  There is no CPU implementation behind this module: if the library is
missing, or no CUDA device is present when a computation is requested, the caller gets an
exception -- never a silent fallback.
"""

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libtabcorr_b200.so')

TC_MODE_AUTO = 0
TC_MODE_CROSS = 1
TC_PRECISION_FP64 = 0
TC_PRECISION_3XTF32 = 1


def precision_code(precision):
    """'fp64' | '3xtf32' (or the integer codes) -> TC_PRECISION_*."""
    table = {'fp64': TC_PRECISION_FP64, 'f64': TC_PRECISION_FP64, '3xtf32': TC_PRECISION_3XTF32,
             TC_PRECISION_FP64: TC_PRECISION_FP64, TC_PRECISION_3XTF32: TC_PRECISION_3XTF32}
    key = precision.lower() if isinstance(precision, str) else precision
    if key not in table:
        raise ValueError("precision must be 'fp64' or '3xtf32'")
    return table[key]
TC_N_THETA = 7
TC_MAX_KNOTS = 4
TC_N_THETA_LEAUTHAUD11 = 18
TC_FAMILY_ZHENG07 = 0
TC_FAMILY_LEAUTHAUD11 = 1

# every symbol include/tabcorr_b200.h declares
SYMBOLS = (
    'tc_last_error', 'tc_version', 'tc_model_n_theta', 'tc_table_create', 'tc_table_destroy', 'tc_table_n_rows',
    'tc_table_n_r', 'tc_table_n_tables', 'tc_table_plan', 'tc_occupation_batch',
    'tc_predict_workspace_bytes', 'tc_predict_workspace_bytes_for', 'tc_predict_batch', 'tc_predict_one', 'tc_interp_create', 'tc_interp_destroy',
    'tc_interp_apply_batch', 'tc_measure_dmma_peak', 'tc_measure_dfma_peak', 'tc_profile_enable', 'tc_profile_read', 'tc_debug_math',
    'tc_peer_alloc', 'tc_peer_open', 'tc_peer_close', 'tc_peer_free', 'tc_halo_bins')


class TabCorrB200Error(RuntimeError):
    """A call into libtabcorr_b200 failed (message from ``tc_last_error``)."""


class tc_model(ctypes.Structure):
    _fields_ = [('family', ctypes.c_int32), ('decorated', ctypes.c_int32),
                ('modulate_with_cenocc', ctypes.c_int32), ('n_scatter', ctypes.c_int32),
                ('split', ctypes.c_double), ('threshold', ctypes.c_double),
                ('redshift', ctypes.c_double),
                # mass-dependent decoration, [centrals, satellites] (include/tabcorr_b200.h)
                ('n_strength', ctypes.c_int32 * 2), ('n_split', ctypes.c_int32 * 2),
                ('strength_abscissa', (ctypes.c_double * TC_MAX_KNOTS) * 2),
                ('split_abscissa', (ctypes.c_double * TC_MAX_KNOTS) * 2),
                ('split_ordinates', (ctypes.c_double * TC_MAX_KNOTS) * 2),
                # leauthaud11: control points of a mass-dependent stellar-mass scatter
                ('scatter_abscissa', ctypes.c_double * TC_MAX_KNOTS)]


_lib = None


def load():
    """Load the shared library (once) and declare the argument types."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise TabCorrB200Error(
            'the CUDA extension {} has not been built; run `python -c "import __graft_entry__ as '
            'g; g.build()"` (or `python -m tabcorr_b200.build`) in the repository root. '
            'tabcorr_b200 has no CPU fallback.'.format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    c_double_p = ctypes.POINTER(ctypes.c_double)
    c_int32_p = ctypes.POINTER(ctypes.c_int32)
    vp = ctypes.c_void_p
    lib.tc_last_error.restype = ctypes.c_char_p
    lib.tc_last_error.argtypes = []
    lib.tc_version.restype = ctypes.c_int
    lib.tc_version.argtypes = []
    lib.tc_model_n_theta.restype = ctypes.c_int
    lib.tc_model_n_theta.argtypes = [ctypes.POINTER(tc_model)]
    lib.tc_table_create.restype = ctypes.c_int
    lib.tc_table_create.argtypes = [
        ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p,
        c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p, ctypes.POINTER(c_double_p),
        ctypes.c_int]
    lib.tc_table_destroy.restype = ctypes.c_int
    lib.tc_table_destroy.argtypes = [vp]
    for name in ('tc_table_n_rows', 'tc_table_n_r', 'tc_table_n_tables'):
        getattr(lib, name).restype = ctypes.c_int
        getattr(lib, name).argtypes = [vp]
    lib.tc_table_plan.restype = ctypes.c_int
    lib.tc_table_plan.argtypes = [vp, ctypes.c_int, c_double_p, c_double_p]
    lib.tc_occupation_batch.restype = ctypes.c_int
    lib.tc_occupation_batch.argtypes = [vp, ctypes.POINTER(tc_model), ctypes.c_int, vp,
                                        ctypes.c_int64, ctypes.c_int64, vp, vp]
    lib.tc_predict_workspace_bytes.restype = ctypes.c_size_t
    lib.tc_predict_workspace_bytes.argtypes = [vp, ctypes.c_int64, ctypes.c_int]
    lib.tc_predict_workspace_bytes_for.restype = ctypes.c_size_t
    lib.tc_predict_workspace_bytes_for.argtypes = [vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int]
    lib.tc_predict_batch.restype = ctypes.c_int
    lib.tc_predict_batch.argtypes = [
        vp, ctypes.POINTER(tc_model), ctypes.c_int, vp, ctypes.c_int64, vp, ctypes.c_int64,
        ctypes.c_int, ctypes.c_int, vp, ctypes.c_int64, vp, ctypes.c_int64, vp, ctypes.c_size_t,
        vp]
    lib.tc_predict_one.restype = ctypes.c_int
    lib.tc_predict_one.argtypes = [
        vp, ctypes.POINTER(tc_model), ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, vp,
        ctypes.c_int64, vp, ctypes.c_int64, vp, ctypes.c_size_t, vp]
    lib.tc_interp_create.restype = ctypes.c_int
    lib.tc_interp_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int, c_int32_p, c_double_p,
                                     c_double_p, c_int32_p, ctypes.c_int]
    lib.tc_interp_destroy.restype = ctypes.c_int
    lib.tc_interp_destroy.argtypes = [vp]
    lib.tc_interp_apply_batch.restype = ctypes.c_int
    lib.tc_interp_apply_batch.argtypes = [vp, vp, ctypes.c_int64, vp, ctypes.c_int, vp,
                                          ctypes.c_int, vp, vp]
    lib.tc_measure_dmma_peak.restype = ctypes.c_int
    lib.tc_measure_dmma_peak.argtypes = [ctypes.c_int, c_double_p]
    lib.tc_measure_dfma_peak.restype = ctypes.c_int
    lib.tc_measure_dfma_peak.argtypes = [ctypes.c_int, c_double_p]
    lib.tc_halo_bins.restype = ctypes.c_int
    lib.tc_halo_bins.argtypes = [ctypes.c_int, vp, vp, vp, ctypes.c_int64, c_double_p, ctypes.c_int,
                                 c_double_p, ctypes.c_int, c_double_p, c_double_p, c_double_p, vp]
    lib.tc_profile_enable.restype = ctypes.c_int
    lib.tc_profile_enable.argtypes = [ctypes.c_int]
    lib.tc_profile_read.restype = ctypes.c_int
    lib.tc_profile_read.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    c_ubyte_p = ctypes.POINTER(ctypes.c_ubyte)
    lib.tc_peer_alloc.restype = ctypes.c_int
    lib.tc_peer_alloc.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(vp), c_ubyte_p]
    lib.tc_peer_open.restype = ctypes.c_int
    lib.tc_peer_open.argtypes = [ctypes.c_int, c_ubyte_p, ctypes.POINTER(vp)]
    lib.tc_peer_close.restype = ctypes.c_int
    lib.tc_peer_close.argtypes = [ctypes.c_int, vp]
    lib.tc_peer_free.restype = ctypes.c_int
    lib.tc_peer_free.argtypes = [ctypes.c_int, vp]
    lib.tc_debug_math.restype = ctypes.c_int
    lib.tc_debug_math.argtypes = [ctypes.c_int, vp, vp, vp, ctypes.c_int64, vp]
    _lib = lib
    return lib


def check(status):
    if status != 0:
        message = load().tc_last_error().decode('utf-8', 'replace')
        raise TabCorrB200Error('libtabcorr_b200 error {}: {}'.format(status, message))


def as_double_p(array):
    return array.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def as_int32_p(array):
    return array.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
