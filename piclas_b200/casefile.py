"""Case files for the C++ host driver (piclas_b200/host/run_case.cpp).

A case file holds what the Fortran host would hand over at run time — the particle-mesh tables and parsed parameters of
pgpu_mesh_t / pgpu_params_t (include/piclas_gpu.h), the particles, the field — as a flat sequence of named arrays, so that a
compiled host without numpy or HDF5 can drive the C ABI with exactly the inputs of the Python host.  Format (little endian):
magic "PGPUCASE1\\n", then records: u32 name length, name, u8 dtype ('i' int32, 'l' int64, 'd' float64), u32 ndim, ndim x i64
extents, raw data (C order)."""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from .abi import Marshalled, pgpu_mesh_t, pgpu_params_t

MAGIC = b"PGPUCASE1\n"
_CODES = {np.dtype(np.int32): b"i", np.dtype(np.int64): b"l", np.dtype(np.float64): b"d"}
_DTYPES = {v: k for k, v in _CODES.items()}


def write_arrays(path, arrays: dict):
    with open(path, "wb") as f:
        f.write(MAGIC)
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            if a.dtype not in _CODES:
                raise TypeError("%s: dtype %s not supported" % (name, a.dtype))
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)) + nb + _CODES[a.dtype] + struct.pack("<I", a.ndim))
            f.write(struct.pack("<%dq" % a.ndim, *a.shape))
            f.write(a.tobytes())


def read_arrays(path) -> dict:
    b = open(path, "rb").read()
    if not b.startswith(MAGIC):
        raise ValueError("not a PGPUCASE1 file")
    p, out = len(MAGIC), {}
    while p < len(b):
        (n,) = struct.unpack_from("<I", b, p)
        name = b[p + 4:p + 4 + n].decode()
        p += 4 + n
        dt = _DTYPES[b[p:p + 1]]
        (nd,) = struct.unpack_from("<I", b, p + 1)
        shape = struct.unpack_from("<%dq" % nd, b, p + 5)
        p += 5 + 8 * nd
        cnt = int(np.prod(shape)) if nd else 1
        out[name] = np.frombuffer(b, dtype=dt, count=cnt, offset=p).reshape(shape).copy()
        p += cnt * dt.itemsize
    return out


def struct_arrays(prefix, st, keep) -> dict:
    """Every field of a pgpu_* struct as a named array; pointer fields are resolved to the numpy arrays `keep` owns."""
    by_addr = {a.ctypes.data: a for a in keep}
    out = {}
    for name, typ in st._fields_:
        v = getattr(st, name)
        if typ in (C.c_int32, C.c_int64, C.c_double):
            out[prefix + name] = np.array([v], dtype={C.c_int32: np.int32, C.c_int64: np.int64, C.c_double: np.float64}[typ])
        elif issubclass(typ, C.Array):
            out[prefix + name] = np.array(list(v), dtype=np.float64 if typ._type_ is C.c_double else np.int32)
        else:
            addr = C.cast(v, C.c_void_p).value
            if addr is not None:                      # NULL pointers are simply absent from the file
                out[prefix + name] = by_addr[addr]
    return out


def write_case(path, mesh, params, PartState, PartSpecies, GlobalElemID, E, dt, nsteps, IsNewPart=None, ids=None):
    mar = Marshalled(mesh, params)
    arrays = {}
    arrays.update(struct_arrays("mesh.", mar.mesh, mar.keep))
    arrays.update(struct_arrays("params.", mar.params, mar.keep))
    n = len(PartSpecies)
    arrays["part.PartState"] = np.ascontiguousarray(PartState, dtype=np.float64)
    arrays["part.PartSpecies"] = np.ascontiguousarray(PartSpecies, dtype=np.int32)
    arrays["part.GlobalElemID"] = np.ascontiguousarray(GlobalElemID, dtype=np.int32)
    arrays["part.IsNewPart"] = np.ascontiguousarray(IsNewPart if IsNewPart is not None else np.ones(n), dtype=np.int32)
    if ids is not None:
        arrays["part.ids"] = np.ascontiguousarray(ids, dtype=np.int64)
    arrays["field.E"] = np.ascontiguousarray(E, dtype=np.float64)
    arrays["run.dt"] = np.array([dt], dtype=np.float64)
    arrays["run.nsteps"] = np.array([nsteps], dtype=np.int32)
    write_arrays(path, arrays)
    return arrays
