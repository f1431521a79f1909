"""The Go toolchain is absent from the build image, so go/dfr2d/dfr2d.go (the cgo binding a gocfd maintainer adds,
INTEGRATION.md) cannot be compiled here.  These static checks keep it from drifting away from include/dfr2d.h: every C
symbol it calls is declared there, every dfr2d_problem field it fills exists, and its delimiters balance."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "go", "dfr2d", "dfr2d.go")).read()
HDR = open(os.path.join(ROOT, "include", "dfr2d.h")).read()


def _code_only(src):
    s = re.sub(r"//[^\n]*", "", src)
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    return re.sub(r'"(\\.|[^"\\])*"', '""', s)


def test_delimiters_balance():
    s = _code_only(SRC)
    for a, b in ("{}", "()", "[]"):
        assert s.count(a) == s.count(b), (a, b)


def test_every_c_symbol_is_declared_in_the_header():
    syms = set(re.findall(r"C\.(dfr2d_[a-z0-9_]+|DFR2D_[A-Z0-9_]+)", SRC))
    assert {"dfr2d_create", "dfr2d_step", "dfr2d_multi_step", "dfr2d_get_state", "dfr2d_destroy"} <= syms
    for name in syms:
        assert re.search(r"\b%s\b" % name, HDR), name


def test_problem_fields_exist():
    body = HDR[HDR.index("typedef struct dfr2d_problem"):HDR.index("} dfr2d_problem;")]
    declared = set(re.findall(r"\b([A-Za-z_][A-Za-z_0-9]*)\s*[;,]", body))
    used = set(re.findall(r"\bp\.([A-Za-z_][A-Za-z_0-9]*)", _code_only(SRC)))
    assert len(used) > 30
    assert not sorted(f for f in used if f not in declared)


def test_cgo_preamble_includes_the_header_and_links_the_library():
    pre = SRC[:SRC.index('import "C"')]
    assert '#include "dfr2d.h"' in pre and "-ldfr2d" in pre
