"""ctypes binding of libbf_b200.so (include/bf_b200.h).

There is no CPU implementation behind this module: if the CUDA library is
missing it is built with nvcc, and if that is impossible, or no B200 is visible
at call time, the call fails loudly (RuntimeError carrying bf_last_error()).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbf_b200.so")

ALGO_PAD, ALGO_LERP, ALGO_FIR_SEQ, ALGO_FIR_LANES, ALGO_HYBRID = 0, 1, 2, 3, 4


class BfConfig(ctypes.Structure):
    _fields_ = [("n_microphones", ctypes.c_int), ("n_samples", ctypes.c_int),
                ("n_taps", ctypes.c_int), ("max_res_x", ctypes.c_int),
                ("max_res_y", ctypes.c_int), ("mic_gain", ctypes.c_float),
                ("fir_fused", ctypes.c_int)]


class BfHeatInfo(ctypes.Structure):
    _fields_ = [("max_power", ctypes.c_float), ("min_power", ctypes.c_float),
                ("log_span", ctypes.c_float), ("smooth_max", ctypes.c_float),
                ("center_col", ctypes.c_double), ("center_row", ctypes.c_double),
                ("overlay", ctypes.c_int), ("painted", ctypes.c_int),
                ("fallback", ctypes.c_int), ("reserved", ctypes.c_int)]


class BfRecordLayout(ctypes.Structure):
    """include/bf_b200.h: bf_record_layout (size and member offsets of a shared-memory record)."""
    _fields_ = [("size", ctypes.c_size_t), ("off", ctypes.c_size_t * 4)]


HEAT_INFO_DTYPE = np.dtype([("max_power", "<f4"), ("min_power", "<f4"), ("log_span", "<f4"),
                            ("smooth_max", "<f4"), ("center_col", "<f8"), ("center_row", "<f8"),
                            ("overlay", "<i4"), ("painted", "<i4"), ("fallback", "<i4"),
                            ("reserved", "<i4")])

DATA_SOURCE_FN = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_float))

_lib = None


def _build():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bf_b200_build",
                                                  os.path.join(os.path.dirname(_HERE), "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()


def lib():
    """The loaded library (built on first use when the .so is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            _build()
        except Exception as e:  # noqa: BLE001
            raise ImportError(
                "libbf_b200.so is not built and could not be built (%s). Run "
                "`python zybo-rt-sampler-image-detection_b200/build.py`; there is no CPU fallback." % e)
    L = ctypes.CDLL(LIB_PATH)
    L.bf_last_error.restype = ctypes.c_char_p
    L.bf_version.restype = ctypes.c_char_p
    L.bf_kernel_launches.restype = ctypes.c_uint64
    vp, ci, cs = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    L.bf_mimo_dev.argtypes = [ci, vp, vp, ci, vp, ci, ci, ci, vp]
    L.bf_mimo_dev_ex.argtypes = [ci, vp, vp, ci, vp, ci, ci, ci, ctypes.c_long, ctypes.c_long, ci, vp]
    L.bf_miso_dev.argtypes = [ci, vp, vp, ci, vp, ci, ci, ci, vp]
    L.bf_mimo_host_batch.argtypes = [ci, vp, vp, ci, vp, ci]
    cd = ctypes.c_double
    L.bf_fd_setup.argtypes = [ci, ci, cd, cd, ci, ci, vp, ci, vp, ci, cd, vp, vp, vp, ci]
    L.bf_fd_das.argtypes = [vp, vp, ci, ctypes.c_float, ci]
    L.bf_fd_das_dev.argtypes = [vp, vp, ci, ctypes.c_float, ci, vp]
    L.bf_fd_mvdr.argtypes = [vp, vp, ci, cd]
    L.bf_fd_mvdr_dev.argtypes = [vp, vp, ci, cd, vp]
    L.bf_fd_get_covariance.argtypes = [vp, cs]
    L.bf_ingest_dev.argtypes = [vp, vp, ci, ci, ci, ci, cd, ci, vp, vp]
    L.bf_window_dev.argtypes = [vp, ctypes.c_long, vp, ci, vp, vp]
    L.bf_ingest_windows_dev.argtypes = [vp, ctypes.c_long, vp, vp, ci, ci, ci, ci, cd, ci, vp, vp]
    cf, cl = ctypes.c_float, ctypes.c_long
    cll = ctypes.c_longlong
    L.bf_dev_alloc.argtypes = [cs, ctypes.POINTER(vp)]
    L.bf_dev_free.argtypes = [vp]
    L.bf_ipc_export.argtypes = [vp, vp]
    L.bf_ipc_open.argtypes = [vp, ctypes.POINTER(vp)]
    L.bf_ipc_close.argtypes = [vp]
    L.bf_mimo_dev_gather.argtypes = [ci, vp, ci, vp, ci, ci, ci, ci, ci, vp, ctypes.c_long, vp]
    L.bf_mimo_dev_gather_sync.argtypes = [ci, vp, ci, vp, ci, ci, ci, ci, ci, vp, ctypes.c_long, vp, cll, cll, vp, vp]
    L.bf_gather_overlap.argtypes = [ci]
    L.bf_peer_copy.argtypes = [vp, vp, ctypes.c_size_t, vp]
    L.bf_host_batch_schedule.argtypes = [ci, vp, ci]
    L.bf_mimo_plan.argtypes = [ci, ci, ci, ci, ci, vp, vp, vp]
    L.bf_mimo_walk.argtypes = [ci, ci, ci, ci, ci, ci, vp, ctypes.c_long]
    L.bf_mimo_walk.restype = ctypes.c_long
    L.bf_gather_signal.argtypes = [vp, ci, ci, cll, vp]
    L.bf_gather_wait.argtypes = [vp, ci, cll, vp, vp]
    L.bf_peer_scatter.argtypes = [vp, ctypes.c_long, ci, ci, ci, vp, ctypes.c_long, vp]
    L.bf_fd_mvdr_dev_slice.argtypes = [vp, vp, ci, cd, ci, ci, vp]
    L.bf_fd_das_dev_slice.argtypes = [vp, vp, ci, ci, ci, vp]
    L.bf_fd_normalise_dev.argtypes = [vp, ci, ctypes.c_float, ci, vp]
    L.bf_fd_mvdr_factor_dev.argtypes = [vp, ci, cd, ci, ci, vp]
    L.bf_fd_mvdr_operands.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(cs), ctypes.POINTER(vp)]
    L.bf_fd_mvdr_steer_dev.argtypes = [vp, ci, ci, vp]
    L.bf_layout_miso.argtypes = [ci, ci]
    L.bf_layout_miso.restype = BfRecordLayout
    L.bf_layout_padata.argtypes = [ci]
    L.bf_layout_padata.restype = BfRecordLayout
    L.bf_layout_ring_buffer.argtypes = [ci, ci]
    L.bf_layout_ring_buffer.restype = BfRecordLayout
    L.bf_miso_record_listen.argtypes = [vp, vp]
    L.bf_kf_create.restype = vp
    L.bf_kf_destroy.argtypes = [vp]
    L.bf_kf_destroy.restype = None
    L.bf_kf_update.argtypes = [vp, vp]
    L.bf_kf_get_state.argtypes = [vp, vp]
    L.bf_kf_predict.argtypes = [vp, ci, vp]
    L.bf_jet_lut.argtypes = [vp]
    L.bf_heatmap_dev.argtypes = [vp, ci, cl, ci, ci, cf, cf, ci, ci, vp, vp, vp, vp, vp]
    L.bf_resize_linear_u8_dev.argtypes = [vp, ci, ci, ci, ci, vp, ci, ci, vp]
    L.bf_entropy_dev.argtypes = [vp, ci, cl, vp, vp]
    L.bf_heatmap.argtypes = [vp, ci, ci, ci, cf, cf, ci, ci, vp, ci, ci, vp, vp, vp]
    L.bf_load_table_dev.argtypes = [ci, vp, cs]
    L.bf_generate_delays.argtypes = [ctypes.c_double, vp, ci, vp, ci, ctypes.c_double, vp, vp, ci,
                                     vp, vp, vp, ci]
    L.bf_get_lerp_tables.argtypes = [vp, vp, cs]
    L.bf_get_hybrid_tables.argtypes = [vp, vp, cs]
    L.bf_set_data_source.argtypes = [DATA_SOURCE_FN]
    for name in ("mimo_pad", "mimo_lerp", "mimo_convolve_naive", "mimo_convolve_vectorized",
                 "mimo_convolve_hybrid"):
        getattr(L, name).argtypes = [vp, vp, vp, ci]
        getattr(L, name).restype = None
    for name in ("miso_pad", "miso_pad2", "miso_lerp", "miso_convolve_naive",
                 "miso_convolve_vectorized", "miso_convolve_hybrid"):
        getattr(L, name).argtypes = [vp, vp, vp, ci, ci]
        getattr(L, name).restype = None
    for name in ("load_coefficients_pad", "load_coefficients_pad2", "load_coefficients_lerp",
                 "load_coefficients_convolve", "load_coefficients_convolve_hybrid",
                 "load_coefficients2"):
        getattr(L, name).argtypes = [vp, ci]
        getattr(L, name).restype = None
    for name in ("pad_mimo", "lerp_mimo", "convolve_mimo_naive", "convolve_mimo_vectorized",
                 "mimo_truncated"):
        getattr(L, name).argtypes = [vp, vp, ci]
        getattr(L, name).restype = None
    L.miso_steer_listen.argtypes = [vp, vp, ci, ci]
    L.miso_steer_listen.restype = None
    L.pad_delay.argtypes = [vp, vp, ci]
    L.lerp_delay.argtypes = [vp, vp, ctypes.c_float, ci]
    L.convolve_hybrid_delay_add.argtypes = [vp, vp, ci, vp]
    for name in ("convolve_delay_naive_add", "convolve_delay_vectorized",
                 "convolve_delay_vectorized_add", "convolve_delay_naive"):
        getattr(L, name).argtypes = [vp, vp, vp]
    _lib = L
    return L


def check(status=None):
    """Raise if the last library call failed (part-1 functions return void)."""
    L = lib()
    st = L.bf_last_status() if status is None else status
    if st != 0:
        raise RuntimeError("bf_b200: %s (status %d)" % (L.bf_last_error().decode(), st))


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def configure(n_microphones, n_samples, n_taps, max_res_x, max_res_y, mic_gain=128.0,
              fir_fused=-1):
    cfg = BfConfig(int(n_microphones), int(n_samples), int(n_taps), int(max_res_x), int(max_res_y),
                   float(mic_gain), int(fir_fused))
    check(lib().bf_configure(ctypes.byref(cfg)))


def configure_from(config):
    """Push the sizes of an `interface.config`-like module into the library."""
    configure(config.N_MICROPHONES, config.N_SAMPLES, getattr(config, "N_TAPS", 8),
              config.MAX_RES_X, config.MAX_RES_Y, getattr(config, "MIC_GAIN", 128),
              getattr(config, "FIR_FUSED", -1))


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)
