# -*- coding: utf-8 -*-
"""
ctypes binding of ``libqmb200.so`` (C ABI declared in ``include/qmb200.h``).

Mirrors the loading discipline of the reference (``quakemigrate/core/libnames.py:18-47``: a
``ctypes.CDLL`` next to the package, ``ImportError`` with build instructions on failure) and its
argument discipline (``quakemigrate/core/lib.py:27-49``: ``numpy.ctypeslib.ndpointer`` argtypes
enforcing dtype and C-contiguity).  There is no fallback of any kind: if the CUDA library is
missing or no B200 is visible, calls raise.
"""

from __future__ import annotations

import ctypes
import os
import pathlib

import numpy as np
import numpy.ctypeslib as clib

_HERE = pathlib.Path(__file__).resolve().parent
# QMB_LIB: load another build of the library (tuning experiments); default the in-tree one
LIB_PATH = pathlib.Path(os.environ["QMB_LIB"]) if os.environ.get("QMB_LIB") else _HERE / "libqmb200.so"

c_int32 = ctypes.c_int32
c_int64 = ctypes.c_int64
c_double = ctypes.c_double
c_void_p = ctypes.c_void_p
c_dPt = clib.ndpointer(dtype=np.double, flags="C_CONTIGUOUS")
c_i32Pt = clib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
c_i64Pt = clib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
c_f32Pt = clib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")

# numpy mirror of ``stalta_header`` (qmlib.h:34-38; lib.py:33-36)
stalta_header_t = np.dtype([("n", c_int32), ("nsta", c_int32), ("nlta", c_int32)], align=True)
stalta_header_pt = clib.ndpointer(stalta_header_t, flags="C_CONTIGUOUS")

QMB_NCCL_ID_BYTES = 128
STATUS_NAMES = {0: "QMB_OK", 1: "QMB_ERR_CUDA", 2: "QMB_ERR_ARG", 3: "QMB_ERR_UNSUPPORTED",
                4: "QMB_ERR_NCCL", 5: "QMB_ERR_STATE"}
STALTA_POSITION = {"classic": 0, "centred": 1, "recursive": 2}
TRANSFORM = {"none": 0, "energy": 1, "abs": 2, "env": 3, "env_squared": 4}


class LutInfo(ctypes.Structure):
    _fields_ = [("n_nodes_local", c_int64), ("n_nodes_global", c_int64), ("node_offset", c_int64),
                ("n_rows", c_int32), ("max_tau", c_int32), ("max_tile_span", c_int32),
                ("n_tiles", c_int32), ("nodes_per_tile", c_int32), ("device_bytes", c_int64),
                ("duplicate_fraction", c_double), ("variant", c_int32), ("reserved", c_int32)]


class ScanInfo(ctypes.Structure):
    """``qmb_scan_info_t`` (include/qmb200.h)."""
    _fields_ = [("operand_bytes", c_int32), ("fell_back_fp32", c_int32), ("check_windows", c_int32),
                ("check_samples", c_int32), ("check_max_rel_err", c_double),
                ("slow_path_samples", c_int32), ("two_candidate_samples", c_int32),
                ("scan_ms", ctypes.c_float), ("tail_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
                ("check_ms", ctypes.c_float), ("host_syncs_in_timed_region", c_int32),
                ("operand_bits", c_int32)]


class QmbError(RuntimeError):
    """A T2 entry point returned a non-zero status (message from ``qmb_last_error``)."""

    def __init__(self, status: int, message: str):
        self.status = status
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")


# name -> (restype, argtypes); every symbol include/qmb200.h declares
SIGNATURES = {
    # T1 -- reference-compatible (qmlib.h:28-44)
    "migrate": (None, [c_dPt, c_i32Pt, c_dPt, c_int32, c_int32, c_int32, c_int32, c_int32, c_int64,
                       c_int64]),
    "find_max_coa": (None, [c_dPt, c_dPt, c_dPt, c_i64Pt, c_int32, c_int64, c_int64]),
    "overlapping_sta_lta": (None, [c_dPt, stalta_header_pt, c_dPt]),
    "centred_sta_lta": (None, [c_dPt, stalta_header_pt, c_dPt]),
    "recursive_sta_lta": (None, [c_dPt, stalta_header_pt, c_dPt]),
    "qmb_legacy_status": (ctypes.c_int, []),
    # T2
    "qmb_last_error": (ctypes.c_char_p, []),
    "qmb_version": (ctypes.c_char_p, []),
    "qmb_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "qmb_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_void_p)]),
    "qmb_ctx_destroy": (ctypes.c_int, [c_void_p]),
    "qmb_ctx_sync": (ctypes.c_int, [c_void_p]),
    "qmb_ctx_set_variant": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "qmb_ctx_set_scan_precision": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "qmb_ctx_set_scan_qgroup": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "qmb_ctx_last_scan_bytes_per_gather": (ctypes.c_int, [c_void_p, ctypes.POINTER(ctypes.c_int)]),
    "qmb_ctx_last_scan_info": (ctypes.c_int, [c_void_p, ctypes.POINTER(ScanInfo)]),
    "qmb_ctx_launch_count": (ctypes.c_int, [c_void_p, ctypes.POINTER(c_int64)]),
    "qmb_ctx_flush_l2": (ctypes.c_int, [c_void_p]),
    "qmb_ctx_timer_start": (ctypes.c_int, [c_void_p]),
    "qmb_ctx_timer_stop": (ctypes.c_int, [c_void_p, ctypes.POINTER(ctypes.c_float)]),
    "qmb_host_register": (ctypes.c_int, [c_void_p, c_int64]),
    "qmb_host_unregister": (ctypes.c_int, [c_void_p]),
    "qmb_comm_unique_id": (ctypes.c_int, [ctypes.c_char_p]),
    "qmb_comm_init": (ctypes.c_int, [c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]),
    "qmb_comm_destroy": (ctypes.c_int, [c_void_p]),
    "qmb_lut_upload": (ctypes.c_int, [c_void_p, c_i32Pt, c_int32, c_int32, c_int32, c_int32, c_int64,
                                      c_int64, ctypes.POINTER(c_void_p)]),
    "qmb_lut_upload_range": (ctypes.c_int, [c_void_p, c_i32Pt, c_int32, c_int32, c_int32, c_int64, c_int64, c_int64,
                                            ctypes.POINTER(c_void_p)]),
    "qmb_lut_free": (ctypes.c_int, [c_void_p]),
    "qmb_ttgrids_upload": (ctypes.c_int, [c_void_p, c_dPt, c_int32, c_int32, c_int32, c_int32,
                                          ctypes.POINTER(c_void_p)]),
    "qmb_ttgrids_free": (ctypes.c_int, [c_void_p]),
    "qmb_ttgrids_decimate": (ctypes.c_int, [c_void_p, c_void_p, c_i32Pt, ctypes.POINTER(c_void_p)]),
    "qmb_ttgrids_shape": (ctypes.c_int, [c_void_p, c_i32Pt]),
    "qmb_lut_serve": (ctypes.c_int, [c_void_p, c_void_p, c_i32Pt, c_int32, c_double, c_int32, c_int32,
                                     ctypes.POINTER(c_void_p)]),
    "qmb_lut_info": (ctypes.c_int, [c_void_p, ctypes.POINTER(LutInfo)]),
    "qmb_onsets_upload": (ctypes.c_int, [c_void_p, c_dPt, c_int32, c_int64, ctypes.POINTER(c_void_p)]),
    "qmb_onsets_from_waveforms": (ctypes.c_int, [c_void_p, c_dPt, c_int32, c_int64, c_int32, c_i32Pt,
                                                 c_i32Pt, c_i32Pt, c_i32Pt, c_int32, c_int32, c_double,
                                                 c_void_p, ctypes.POINTER(c_void_p)]),
    "qmb_onsets_free": (ctypes.c_int, [c_void_p]),
    "qmb_waveforms_alloc": (ctypes.c_int, [c_void_p, c_int32, c_int64, ctypes.POINTER(c_void_p)]),
    "qmb_waveforms_upload": (ctypes.c_int, [c_void_p, c_dPt, c_int32, c_int64, ctypes.POINTER(c_void_p)]),
    "qmb_waveforms_fetch": (ctypes.c_int, [c_void_p, c_int32, c_int32, c_dPt]),
    "qmb_waveforms_free": (ctypes.c_int, [c_void_p]),
    "qmb_waveforms_filter": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int32,
                                            c_int32, c_double, c_int32, c_int32, c_void_p, c_int32,
                                            c_void_p]),
    "qmb_waveforms_upsample": (ctypes.c_int, [c_void_p, c_void_p, c_int32, ctypes.POINTER(c_void_p)]),
    "qmb_onsets_from_device_waveforms": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_i32Pt, c_i32Pt,
                                                        c_i32Pt, c_i32Pt, c_int32, c_int32, c_double,
                                                        c_void_p, ctypes.POINTER(c_void_p)]),
    "qmb_preprocess": (ctypes.c_int, [c_void_p, c_dPt, c_int32, c_int64, c_void_p, c_int32, c_int32,
                                      c_double, c_int32, c_int32, c_dPt]),
    "qmb_envelope": (ctypes.c_int, [c_void_p, c_dPt, c_int32, c_int64, c_dPt]),
    "qmb_sta_lta": (ctypes.c_int, [c_void_p, c_dPt, c_int32, c_int64, c_int32, c_int32, c_int32, c_dPt]),
    "qmb_detect_device": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                         c_void_p]),
    "qmb_detect_fetch": (ctypes.c_int, [c_void_p, c_dPt, c_dPt, c_i64Pt, c_int32]),
    "qmb_detect": (ctypes.c_int, [c_void_p, c_void_p, c_dPt, c_int32, c_int64, c_int32, c_int32, c_int32,
                                  c_dPt, c_dPt, c_i64Pt]),
    "qmb_detect_fetch_quantised": (ctypes.c_int, [c_void_p, c_i32Pt, c_dPt, c_dPt, c_double, c_void_p,
                                                  c_void_p, c_void_p, c_void_p, c_void_p, c_int32]),
    "qmb_steim_encode": (ctypes.c_int, [c_i32Pt, c_int64, c_int32, c_int32, c_int32, c_void_p,
                                        ctypes.POINTER(c_int64)]),
    "qmb_steim_decode": (ctypes.c_int, [c_void_p, c_int32, c_int32, c_int64, c_i32Pt,
                                        ctypes.POINTER(c_int32)]),
    "qmb_locate_window": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                         c_int32, c_int32, c_void_p, c_void_p]),
    "qmb_locate_window_device": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                                c_int32, c_int32, ctypes.POINTER(c_void_p)]),
    "qmb_map3d_upload": (ctypes.c_int, [c_void_p, c_dPt, c_int32, c_int32, c_int32, ctypes.POINTER(c_void_p)]),
    "qmb_map3d_fetch": (ctypes.c_int, [c_void_p, c_dPt]),
    "qmb_map3d_free": (ctypes.c_int, [c_void_p]),
    "qmb_map3d_normalise": (ctypes.c_int, [c_void_p, c_void_p]),
    "qmb_map3d_gaufilt": (ctypes.c_int, [c_void_p, c_void_p, c_double, ctypes.POINTER(c_void_p)]),
    "qmb_map3d_peak_device": (ctypes.c_int, [c_void_p, c_void_p, ctypes.POINTER(c_int64),
                                             ctypes.POINTER(c_double), ctypes.POINTER(c_double)]),
    "qmb_rbf_cubic_peak": (ctypes.c_int, [c_void_p, c_dPt, c_int32, c_int32, ctypes.POINTER(c_int64),
                                          ctypes.POINTER(c_double)]),
    "qmb_map3d_window": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                        c_int32, c_dPt]),
    "qmb_map3d_covfit": (ctypes.c_int, [c_void_p, c_void_p, c_dPt, c_double, c_int32, c_dPt]),
    "qmb_gaufilt3d": (ctypes.c_int, [c_void_p, c_dPt, c_int32, c_int32, c_int32, c_double, c_dPt]),
    "qmb_map3d_peak": (ctypes.c_int, [c_void_p, c_dPt, c_int64, ctypes.POINTER(c_int64),
                                      ctypes.POINTER(c_double), ctypes.POINTER(c_double)]),
    "qmb_microbench_smem_pipe": (ctypes.c_int, [c_void_p, c_int32, ctypes.POINTER(c_double)]),
    "qmb_microbench_smem_pipe_ex": (ctypes.c_int, [c_void_p, c_int32, c_int32, c_int32, ctypes.POINTER(c_double)]),
    "qmb_covfit3d": (ctypes.c_int, [c_void_p, c_dPt, c_int32, c_int32, c_int32, c_dPt, c_double, c_int32,
                                    c_dPt]),
}

_cdll = None


def load() -> ctypes.CDLL:
    """Open ``libqmb200.so`` and attach signatures.  Raises ImportError if it is not built."""
    global _cdll
    if _cdll is not None:
        return _cdll
    try:
        cdll = ctypes.CDLL(str(LIB_PATH))
    except OSError as e:
        raise ImportError(
            f"Could not load the CUDA library '{LIB_PATH.name}'.\n\n{e}\n\nBuild it with "
            "'python -m quakemigrate_b200.build' (needs nvcc; compiled for sm_100a only). "
            "quakemigrate_b200 has no CPU fallback."
        ) from e
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(cdll, name)  # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _cdll = cdll
    return cdll


def check(status: int) -> None:
    if status != 0:
        raise QmbError(status, load().qmb_last_error().decode(errors="replace"))


def check_legacy() -> None:
    """Raise if the last reference-compatible (void) symbol called from this thread failed."""
    check(load().qmb_legacy_status())


def device_count() -> int:
    n = ctypes.c_int(0)
    status = load().qmb_device_count(ctypes.byref(n))
    return n.value if status == 0 else 0
