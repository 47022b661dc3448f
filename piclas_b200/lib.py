"""ctypes loader of libpiclas_gpu.so — the product path. There is no CPU fallback: a missing library or a
missing CUDA device is an error."""
from __future__ import annotations

import ctypes as C
import os

from .abi import pgpu_mesh_t, pgpu_params_t, c_f64p, c_i32p, c_i64p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PICLAS_GPU_LIB") or os.path.join(HERE, "libpiclas_gpu.so")  # env override: kernel-variant experiments

EXPORTS = [
    "piclas_gpu_init", "piclas_gpu_finalize", "piclas_gpu_last_error", "piclas_gpu_upload_particles",
    "piclas_gpu_deposit", "piclas_gpu_set_field", "piclas_gpu_push_track", "piclas_gpu_num_particles",
    "piclas_gpu_download_particles", "piclas_gpu_exchange_info", "piclas_gpu_exchange_recv_buffer",
    "piclas_gpu_exchange_finish", "piclas_gpu_nodesource_device", "piclas_gpu_deposit_finish",
    "piclas_gpu_last_timing", "piclas_gpu_phase_timing", "piclas_gpu_sf_halo_info", "piclas_gpu_get_charge", "piclas_gpu_kinetic_energy",
    "piclas_gpu_node_halo_info", "piclas_gpu_set_stream", "piclas_gpu_exchange_device_info", "piclas_gpu_emit_lattice",
    "piclas_gpu_get_partsource_async", "piclas_gpu_partsource_wait",
]

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m piclas_b200.build` (nvcc, sm_100a). "
            "piclas_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.piclas_gpu_init.argtypes = [C.POINTER(pgpu_mesh_t), C.POINTER(pgpu_params_t)]
    lib.piclas_gpu_last_error.restype = C.c_char_p
    lib.piclas_gpu_upload_particles.argtypes = [C.c_int64, c_f64p, c_i32p, c_i32p, c_i32p, c_i32p, c_f64p, c_i64p, C.c_int32]
    lib.piclas_gpu_deposit.argtypes = [c_f64p, c_f64p]
    lib.piclas_gpu_set_field.argtypes = [c_f64p]
    lib.piclas_gpu_push_track.argtypes = [C.c_double, C.c_int64, c_i32p]
    lib.piclas_gpu_num_particles.restype = C.c_int64
    lib.piclas_gpu_download_particles.argtypes = [C.c_int64, c_f64p, c_i32p, c_i32p, c_f64p, c_i64p, c_i64p]
    lib.piclas_gpu_exchange_info.argtypes = [c_i32p, c_i64p, C.POINTER(C.c_void_p)]
    lib.piclas_gpu_exchange_recv_buffer.argtypes = [C.c_int64, C.POINTER(C.c_void_p)]
    lib.piclas_gpu_exchange_finish.argtypes = [C.c_int64]
    lib.piclas_gpu_nodesource_device.argtypes = [C.POINTER(C.c_void_p)]
    lib.piclas_gpu_deposit_finish.argtypes = [c_f64p, c_f64p]
    lib.piclas_gpu_last_timing.argtypes = [c_f64p, c_i32p]
    lib.piclas_gpu_phase_timing.argtypes = [c_f64p]
    lib.piclas_gpu_get_charge.argtypes = [c_f64p]
    lib.piclas_gpu_kinetic_energy.argtypes = [c_f64p, c_i64p]
    lib.piclas_gpu_node_halo_info.argtypes = [c_i64p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.piclas_gpu_set_stream.argtypes = [C.c_void_p]
    lib.piclas_gpu_get_partsource_async.argtypes = [c_f64p]
    lib.piclas_gpu_partsource_wait.argtypes = []
    lib.piclas_gpu_emit_lattice.argtypes = [C.c_int32, C.c_int32, c_i32p, C.c_double, C.c_double, c_f64p, C.c_int32, c_i64p]
    lib.piclas_gpu_exchange_device_info.argtypes = [C.POINTER(C.c_void_p), c_i64p, c_i64p]
    lib.piclas_gpu_sf_halo_info.argtypes = [c_i64p, c_i64p, c_i32p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    for name in EXPORTS:
        getattr(lib, name)
    _lib = lib
    return lib
