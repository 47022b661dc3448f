"""The three declarations of the drop-in boundary must agree: include/piclas_gpu.h (what the CUDA library is compiled
against), piclas_b200/abi.py (ctypes, what the tests and bench.py call through) and
piclas_b200/fortran/mod_particle_gpu.f90 (ISO_C_BINDING, what a PICLas build binds).  No Fortran compiler exists in this
image, so the Fortran side is checked textually: same fields in the same order with the same C types, one interface per
export with the same number of arguments.  The C side is checked with gcc: offsetof / sizeof of every field against ctypes."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from piclas_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "include", "piclas_gpu.h")).read()
F90 = open(os.path.join(ROOT, "piclas_b200", "fortran", "mod_particle_gpu.f90")).read()
STRUCTS = {"pgpu_mesh_t": ("pgpu_mesh", abi.pgpu_mesh_t), "pgpu_params_t": ("pgpu_params", abi.pgpu_params_t)}


def _strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", "", s, flags=re.S)


def header_fields(tag):
    body = re.search(r"typedef struct %s \{(.*?)\} %s_t;" % (tag, tag), _strip_c_comments(HDR), flags=re.S).group(1)
    out = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"(const )?(int32_t|int64_t|double) (.*)", decl)
        assert m, decl
        base = {"int32_t": "i32", "int64_t": "i64", "double": "f64"}[m.group(2)]
        for item in m.group(3).split(","):
            item = item.strip()
            ptr = item.startswith("*")
            name = item.lstrip("* ")
            cnt = 1
            am = re.match(r"(\w+)\[(\d+)\]", name)
            if am:
                name, cnt = am.group(1), int(am.group(2))
            out.append((name, "ptr" if ptr else base, cnt))
    return out


def ctypes_fields(st):
    kinds = {C.c_int32: "i32", C.c_int64: "i64", C.c_double: "f64"}
    out = []
    for name, typ in st._fields_:
        if typ in kinds:
            out.append((name, kinds[typ], 1))
        elif issubclass(typ, C.Array):
            out.append((name, kinds[typ._type_], typ._length_))
        else:
            out.append((name, "ptr", 1))
    return out


def fortran_fields(tname):
    body = re.search(r"TYPE, BIND\(C\) :: %s\b(.*?)END TYPE" % tname, F90, flags=re.S).group(1)
    kinds = {"INTEGER(C_INT32_T)": "i32", "INTEGER(C_INT64_T)": "i64", "REAL(C_DOUBLE)": "f64", "TYPE(C_PTR)": "ptr"}
    out = []
    for line in body.splitlines():
        line = line.split("!")[0].strip()
        if not line or "::" not in line:
            continue
        typ, names = [x.strip() for x in line.split("::")]
        for item in names.split(","):
            item = item.strip()
            am = re.match(r"(\w+)\((\d+)\)", item)
            out.append((am.group(1), kinds[typ], int(am.group(2))) if am else (item, kinds[typ], 1))
    return out


@pytest.mark.parametrize("tname", sorted(STRUCTS))
def test_struct_fields_agree_in_header_ctypes_and_fortran(tname):
    tag, st = STRUCTS[tname]
    h, c, f = header_fields(tag), ctypes_fields(st), fortran_fields(tname)
    assert len(h) > 20
    assert h == c, [x for x in zip(h, c) if x[0] != x[1]][:3]
    assert h == f, [x for x in zip(h, f) if x[0] != x[1]][:3]


@pytest.mark.parametrize("tname", sorted(STRUCTS))
def test_ctypes_layout_is_the_c_compilers(tname):
    tag, st = STRUCTS[tname]
    names = [n for n, _ in st._fields_]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "piclas_gpu.h"', "int main(void) {",
           '  printf("%%zu\\n", sizeof(%s));' % tname]
    src += ['  printf("%%zu\\n", offsetof(%s, %s));' % (tname, n) for n in names]
    src += ["  return 0;", "}"]
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write("\n".join(src))
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-o",
                        os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        vals = [int(x) for x in subprocess.run([os.path.join(d, "t")], check=True, capture_output=True, text=True).stdout.split()]
    assert vals[0] == C.sizeof(st)
    assert vals[1:] == [getattr(st, n).offset for n in names]


def _c_exports():
    out = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char \*)\s*(piclas_gpu_\w+)\s*\((.*?)\);", _strip_c_comments(HDR), flags=re.S):
        args = " ".join(m.group(2).split())
        out[m.group(1)] = 0 if args in ("void", "") else len(args.split(","))
    return out


def test_fortran_module_binds_every_export_with_the_same_arity():
    exports = _c_exports()
    from piclas_b200 import lib
    assert sorted(exports) == sorted(lib.EXPORTS)
    for name, nargs in exports.items():
        m = re.search(r"FUNCTION %s\((.*?)\)\s*(?:&\s*)?BIND\(C,NAME='%s'\)" % (name, name), F90, flags=re.S)
        assert m, "mod_particle_gpu.f90 has no interface for " + name
        fargs = [a for a in re.sub(r"[&\s]", "", m.group(1)).split(",") if a]
        assert len(fargs) == nargs, (name, fargs, nargs)
        assert re.search(r"PUBLIC ::[^\n]*\b%s\b" % name, F90) or name == "piclas_gpu_last_error", name + " is not PUBLIC"


def test_ctypes_prototypes_have_the_headers_arity():
    from piclas_b200 import lib
    so = lib.load()
    for name, nargs in _c_exports().items():
        fn = getattr(so, name)
        if fn.argtypes is not None:
            assert len(fn.argtypes) == nargs, (name, len(fn.argtypes), nargs)


def _c_prototypes():
    """name -> (return type, [(type, stars, name)]) from the header."""
    out = {}
    for m in re.finditer(r"\b(int|int64_t|const char \*)\s*(piclas_gpu_\w+)\s*\((.*?)\);", _strip_c_comments(HDR), flags=re.S):
        args = []
        text = " ".join(m.group(3).split())
        if text not in ("void", ""):
            for a in text.split(","):
                am = re.match(r"\s*(?:const\s+)?(\w+)\s*(\**)\s*(\w+)\s*$", a)
                assert am, a
                args.append((am.group(1), len(am.group(2)), am.group(3)))
        out[m.group(2)] = (m.group(1).replace(" ", ""), args)
    return out


def test_fortran_argument_types_and_value_attributes_match_the_c_prototypes():
    """Parsed with numpy.f2py's crackfortran (no Fortran compiler in the image): by-value C scalars need VALUE and the same
    kind; C pointers are either a by-reference dummy of the same kind or TYPE(C_PTR),VALUE; `void **` is TYPE(C_PTR) by
    reference; struct pointers are by-reference derived types; return types agree."""
    import numpy.f2py.crackfortran as cf
    cf.verbose = 0
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:           # crackfortran may leave scratch files in the working directory
        os.chdir(d)
        try:
            mod = cf.crackfortran([os.path.join(ROOT, "piclas_b200", "fortran", "mod_particle_gpu.f90")])[0]
        finally:
            os.chdir(cwd)
    funcs = {f["name"]: f for b in mod["body"] if b["block"] == "interface" for f in b["body"]}
    kinds = {"int32_t": ("integer", "c_int32_t"), "int64_t": ("integer", "c_int64_t"), "double": ("real", "c_double"),
             "int": ("integer", "c_int")}
    protos = _c_prototypes()
    assert sorted(protos) == sorted(funcs)
    for name, (ret, cargs) in protos.items():
        f = funcs[name]
        rv = f["vars"][name]
        if ret == "constchar*":
            assert rv["typespec"] == "type" and rv["typename"] == "c_ptr", name
        else:
            assert (rv["typespec"], rv["kindselector"]["kind"]) == kinds[ret], name
        assert len(f["args"]) == len(cargs), name
        for fa, (ctype, stars, cname) in zip(f["args"], cargs):
            v = f["vars"][fa]
            where = "%s(%s)" % (name, cname)
            assert fa == cname.lower(), where                                  # same argument names, same order
            value = "value" in v.get("attrspec", [])
            is_cptr = v["typespec"] == "type" and v.get("typename") == "c_ptr"
            if stars == 0:
                assert value and (v["typespec"], v["kindselector"]["kind"]) == kinds[ctype], where
            elif stars == 2:
                assert ctype == "void" and is_cptr and not value, where
            elif ctype.startswith("pgpu_"):
                assert v["typespec"] == "type" and v["typename"] == ctype and not value, where
            elif is_cptr:
                assert value, where
            else:
                assert not value and (v["typespec"], v["kindselector"]["kind"]) == kinds[ctype], where


def test_ctypes_argument_types_match_the_c_prototypes():
    from piclas_b200 import lib
    so = lib.load()
    scalar = {"double": C.c_double, "int64_t": C.c_int64, "int32_t": C.c_int32, "int": C.c_int}
    pointer = {"double": abi.c_f64p, "int32_t": abi.c_i32p, "int64_t": abi.c_i64p,
               "pgpu_mesh_t": C.POINTER(abi.pgpu_mesh_t), "pgpu_params_t": C.POINTER(abi.pgpu_params_t), "void": C.c_void_p}
    checked = 0
    for name, (ret, cargs) in _c_prototypes().items():
        fn = getattr(so, name)
        if ret == "int64_t":
            assert fn.restype is C.c_int64, name
        elif ret == "constchar*":
            assert fn.restype is C.c_char_p, name
        if fn.argtypes is None:
            assert not cargs, name + " takes arguments but has no ctypes prototype"
            continue
        for at, (ctype, stars, cname) in zip(fn.argtypes, cargs):
            want = scalar[ctype] if stars == 0 else (C.POINTER(C.c_void_p) if stars == 2 else pointer[ctype])
            assert at is want, "%s(%s): ctypes %s, header %s%s" % (name, cname, at, ctype, "*" * stars)
            checked += 1
    assert checked >= 40


def test_missing_library_is_an_error_not_a_fallback(monkeypatch, tmp_path):
    from piclas_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "libpiclas_gpu.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lib.load()
