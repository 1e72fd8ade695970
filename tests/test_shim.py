"""shim/sconeB200_mod.f90 (the iso_c_binding module of the drop-in boundary) against include/scone_b200.h.

There is no Fortran compiler in this image, so the module cannot be compiled here; what can be checked is everything a compiler
would not check anyway - that the module says the same as the header:
  * it is the generator's output for the current header (shim/gen_shim.py),
  * every exported function of the header has an interface block with bind(C, name=...) and the same number of arguments,
  * every struct has a bind(C) type with the same fields in the same order, and the layout those types have under the C rules
    (which bind(C) types follow by definition) - size and offset of every field - is what gcc reports for the header,
  * the plain-C driver of tests/c_driver compiles as strict C99 against the header."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shim"))
import gen_shim  # noqa: E402

F90 = os.path.join(ROOT, "shim", "sconeB200_mod.f90")
HDR = os.path.join(ROOT, "include", "scone_b200.h")
SIZES = {"integer(c_int)": 4, "integer(c_int32_t)": 4, "integer(c_int64_t)": 8, "integer(c_size_t)": 8, "real(c_double)": 8, "type(c_ptr)": 8}


def parse_f90():
    txt = open(F90).read()
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"integer\(c_int32_t\), parameter :: (\w+) = (-?\d+)", txt)}
    types = {}
    for name, body in re.findall(r"type, bind\(C\) :: (\w+)\n(.*?)end type", txt, flags=re.S):
        fields = []
        for ln in body.strip().split("\n"):
            m = re.match(r"\s*(.+?) :: (\w+)(?:\((\w+)\))?\s*$", ln)
            dim = m.group(3)
            fields.append((m.group(2), m.group(1), 1 if dim is None else (int(dim) if dim.isdigit() else consts[dim])))
        types[name] = fields
    funcs = {}
    for m in re.finditer(r"(?:function|subroutine) (\w+)\((.*?)\)\s*(?:&\s*)?(?:result\(\w+\) )?bind\(C, name='(\w+)'\)", txt, flags=re.S):
        args = [a for a in re.sub(r"[&\s]", "", m.group(2)).split(",") if a]
        funcs[m.group(3)] = (m.group(1), args)
    return consts, types, funcs


def layout(types, name):
    """offsets, size, alignment of a bind(C) type by the C rules"""
    off, offs, amax = 0, [], 1
    for fname, ftype, n in types[name]:
        if ftype in SIZES:
            sz = al = SIZES[ftype]
        else:
            sub = re.match(r"type\((\w+)\)", ftype).group(1)
            _, sz, al = layout(types, sub)
        off = (off + al - 1) // al * al
        offs.append((fname, off))
        off += sz * n
        amax = max(amax, al)
    return offs, (off + amax - 1) // amax * amax, amax


def test_module_is_generated_from_the_current_header():
    assert gen_shim.generate() == open(F90).read(), "run python shim/gen_shim.py"


def test_every_export_has_an_interface_with_the_same_arguments():
    H = gen_shim.parse_header(open(HDR).read())
    _, _, funcs = parse_f90()
    assert len(H["funcs"]) >= 60
    for name, ret, args in H["funcs"]:
        assert name in funcs, "no interface for " + name
        fname, fargs = funcs[name]
        assert fname == name and fargs == [a[0] for a in args], name
    # and nothing the library does not export
    assert set(funcs) == {f[0] for f in H["funcs"]}


def test_bind_c_layout_equals_the_headers(tmp_path):
    H = gen_shim.parse_header(open(HDR).read())
    consts, types, _ = parse_f90()
    for k, v in H["defines"] + H["enums"]:
        assert consts[k] == int(v)
    src = ["#include <stdio.h>", "#include <stddef.h>", '#include "scone_b200.h"', "int main(void) {"]
    for name, fields in H["structs"]:
        src.append('  printf("%s size %%zu\\n", sizeof(%s));' % (name, name))
        for fname, ctype, ptr, dim in fields:
            src.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (name, fname, name, fname))
    src += ["  return 0;", "}"]
    c = tmp_path / "abi.c"; c.write_text("\n".join(src))
    exe = str(tmp_path / "abi")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c), "-o", exe])
    got = {}
    for ln in subprocess.check_output([exe], text=True).split("\n"):
        if ln:
            a, b, v = ln.split(); got[(a, b)] = int(v)
    assert len(H["structs"]) >= 11
    for name, fields in H["structs"]:
        assert [f[0] for f in fields] == [f[0] for f in types[name]], name
        offs, size, _ = layout(types, name)
        assert got[(name, "size")] == size, name
        for fname, off in offs:
            assert got[(name, fname)] == off, "%s.%s" % (name, fname)


def test_plain_c_driver_is_strict_c99(tmp_path):
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-c",
                           os.path.join(ROOT, "tests", "c_driver", "run_cycle.c"), "-o", str(tmp_path / "run_cycle.o")])
    src = open(os.path.join(ROOT, "tests", "c_driver", "run_cycle.c")).read()
    assert "sbh_" not in re.sub(r"/\*.*?\*/", "", src, flags=re.S)          # the engine's entry points only
