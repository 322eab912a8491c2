"""CPU-side checks of the C-ABI boundary: the shared library loads without a GPU, exports every symbol that
include/ppr_b200.h declares, and the ctypes mirror of ppr_model_desc matches the header field for field."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ppr_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ppr_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    from ppr_diffphys_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = C.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), "libppr_b200.so does not export %s" % n
    assert set(names) == set(_lib.EXPORTS)
    lib.ppr_version.restype = C.c_char_p
    assert b"sm_100a" in lib.ppr_version()


def test_model_desc_mirror_matches_header():
    from ppr_diffphys_b200._capi import ModelDesc
    src = open(HEADER).read()
    body = re.search(r"typedef struct ppr_model_desc \{(.*?)\} ppr_model_desc;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for name in re.split(r",", decl.split(None, 1)[1] if " " in decl else decl):
            name = name.strip().lstrip("*").split("[")[0].strip()
            name = name.split()[-1].lstrip("*")
            fields.append(name)
    assert fields == [f[0] for f in ModelDesc._fields_]


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from ppr_diffphys_b200 import SimEnv, _lib
    with pytest.raises(_lib.PprError):
        SimEnv("laikago")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ppr_diffphys_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                code = "\n".join(l for l in txt.splitlines() if re.match(r"\s*(import|from|#include)\b", l))
                assert "oracle" not in code, f


def test_header_is_plain_c():
    """The boundary is a C ABI: include/ppr_b200.h must compile as C99 (no C++ / torch types in the signatures)."""
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.NamedTemporaryFile("w", suffix=".c", delete=False) as f:
        f.write('#include "include/ppr_b200.h"\nint main(void) { ppr_model_desc d; (void)d; return 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", root, f.name],
                       capture_output=True, text=True)
    os.unlink(f.name)
    assert r.returncode == 0, r.stderr


def test_model_create_validates_before_touching_the_device():
    """ppr_model_create rejects malformed articulations with the documented negative codes (validation runs before
    any CUDA call, so this is checkable without a GPU), and a VALID description fails loudly -- a positive
    cudaError_t, never a silent CPU model -- when no device exists."""
    import copy
    import numpy as np
    import torch
    from ppr_diffphys_b200 import _lib, load_robot
    from ppr_diffphys_b200._capi import make_desc
    lib = C.CDLL(_lib.LIB_PATH)
    lib.ppr_model_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]

    def create(rm):
        d, keep = make_desc(rm)
        h = C.c_void_p()
        rc = lib.ppr_model_create(C.byref(d), C.byref(h))
        return rc, h

    base = load_robot("laikago")
    assert lib.ppr_model_create(None, None) == -1                                   # PPR_E_ARG
    bad = copy.deepcopy(base)
    bad.joint_type = np.array(bad.joint_type).copy()
    bad.joint_type[3] = 2                                                           # prismatic: not a joint of this path
    assert create(bad)[0] == -2                                                     # PPR_E_SHAPE
    bad = copy.deepcopy(base)
    bad.joint_parent = np.array(bad.joint_parent).copy()
    bad.joint_parent[2] = 5                                                         # parent after child
    assert create(bad)[0] == -2
    bad = copy.deepcopy(base)
    bad.contact_body = np.array(bad.contact_body).copy()
    bad.contact_body[0] = 99                                                        # contact on a body that does not exist
    assert create(bad)[0] == -1
    d, keep = make_desc(base)
    d.nb = 33                                                                       # more bodies than a warp has lanes
    assert lib.ppr_model_create(C.byref(d), C.byref(C.c_void_p())) == -2
    if not torch.cuda.is_available():
        rc, h = create(base)
        assert rc > 0 and not h.value, rc                                           # cudaError_t, no handle
