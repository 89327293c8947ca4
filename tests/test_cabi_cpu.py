"""C-ABI checks that need no GPU: the library loads, exports every symbol include/gto_b200.h declares, the ctypes
structures match the header, and the product path fails loudly (no CPU fallback) without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from grasptrajopt_b200 import capi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    with open(os.path.join(REPO, "include", "gto_b200.h")) as fh:
        return fh.read()


def test_library_exports_every_declared_symbol():
    hdr = _header()
    declared = sorted(set(re.findall(r"\b(gto_[a-z_]+)\s*\(", hdr)))
    assert set(declared) == set(capi.SYMBOLS)
    lib = capi.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gto_abi_version() == int(re.search(r"#define GTO_ABI_VERSION (\d+)", hdr).group(1))


def test_default_options_match_oracle():
    import gto_oracle as O

    o = capi.default_options()
    ref = O.SolverOptions()
    for k in ("max_iter", "tol_step", "tol_grad", "lambda0", "lambda_min", "lambda_max", "eta", "noise_rel", "bound_eps", "ftol", "lambda_slow", "slow_window", "slow_ftol"):
        assert getattr(o, k) == getattr(ref, k), k


def test_struct_layouts_follow_header_field_order():
    hdr = _header()

    def fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
        return names

    assert fields("gto_robot_desc") == [f[0] for f in capi.RobotDesc._fields_]
    assert fields("gto_options") == [f[0] for f in capi.Options._fields_]
    assert fields("gto_batch_in") == [f[0] for f in capi.BatchIn._fields_]
    assert fields("gto_batch_out") == [f[0] for f in capi.BatchOut._fields_]
    assert fields("gto_eval_out") == [f[0] for f in capi.EvalOut._fields_]
    assert fields("gto_profile") == [f[0] for f in capi.Profile._fields_]
    assert fields("gto_base_in") == [f[0] for f in capi.BaseIn._fields_]
    assert fields("gto_base_out") == [f[0] for f in capi.BaseOut._fields_]


def test_status_and_flag_constants_match_header():
    hdr = _header()
    for name, val in (("GTO_STATUS_CONVERGED", capi.STATUS_CONVERGED), ("GTO_STATUS_MAX_ITER", capi.STATUS_MAX_ITER),
                      ("GTO_STATUS_NAN", capi.STATUS_NAN), ("GTO_STATUS_STALLED", capi.STATUS_STALLED), ("GTO_STATUS_SLOW", capi.STATUS_SLOW)):
        assert int(re.search(r"#define %s (\d+)" % name, hdr).group(1)) == val
    for name, val in (("GTO_FLAG_NO_JROWS", capi.FLAG_NO_JROWS), ("GTO_FLAG_NO_CULL", capi.FLAG_NO_CULL)):
        assert int(re.search(r"#define %s (\d+)u" % name, hdr).group(1)) == val


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path must raise, never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the loud-failure path is exercised on the CPU-only builder")
    with pytest.raises(capi.GtoError):
        capi.GtoContext(0)


def test_missing_library_is_loud(tmp_path):
    with pytest.raises(capi.GtoLibraryError):
        capi.load_library(str(tmp_path / "nope.so"))


def test_product_code_never_imports_the_oracle():
    bad = []
    for root, _, files in os.walk(os.path.join(REPO, "grasptrajopt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                with open(os.path.join(root, f), errors="ignore") as fh:
                    txt = fh.read()
                if re.search(r"^\s*(import|from)\s+gto_oracle", txt, re.M) or "oracle/" in txt and f.endswith(".py") and "import" in txt and re.search(r"sys\.path.*oracle", txt):
                    bad.append(f)
    assert not bad, bad
