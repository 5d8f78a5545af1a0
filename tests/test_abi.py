"""The C-ABI library loads and exports every symbol include/hadi.h declares; without a GPU the
product fails loudly instead of falling back to anything."""
import ctypes
import os
import re

from conftest import ROOT


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "hadi.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(hadi_[a-z0-9_]+)\s*\(", txt))
    names.discard("hadi_allgather_fn")
    return sorted(names)


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ("hadi_price_batch", "hadi_jacobian_batch", "hadi_lm_update", "hadi_solve5", "hadi_calibrate",
                 "hadi_batch_create", "hadi_batch_launch", "hadi_batch_fetch", "hadi_partition"):
        assert must in names


def test_library_exports_every_declared_symbol(hadi):
    lib = ctypes.CDLL(hadi.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), "libhadi.so does not export %s" % name
    for name in hadi.EXPORTS:
        assert hasattr(lib, name)


def test_struct_layouts_match_the_header(hadi):
    # hadi_point mirrors the reference's CalibrationPoint {double, double, int, double, int}
    assert ctypes.sizeof(hadi.Point) == 40
    assert ctypes.sizeof(hadi.Model) == 64
    assert hadi.Point.delta_t.offset == 24 and hadi.Point.global_index.offset == 32
    assert hadi.Numerics.theta.offset == 8 and hadi.Numerics.dividend_dates.offset == 32


def test_no_cpu_fallback(hadi):
    import torch

    if torch.cuda.is_available():
        return  # covered by the gpu tests
    h = ctypes.c_void_p()
    rc = hadi.lib().hadi_create(ctypes.byref(h), 0)
    assert rc == hadi.ERR_CUDA and not h.value
    try:
        hadi.Context(0)
    except hadi.HadiError as e:
        assert e.code == hadi.ERR_CUDA
    else:
        raise AssertionError("Context() must fail without a CUDA device")


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "pde-based-heston-solver-gpu-accelerated_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                # no include / import / dlopen / exec of anything under oracle/ (prose about "the oracle" in
                # comments is fine; code lines are what is checked)
                for ln in txt.splitlines():
                    code = ln.split("//")[0].split("#include")[-1] if f.endswith((".cpp", ".cu", ".cuh", ".h")) else ln.split("#")[0]
                    assert "oracle/" not in code and "oracle." not in code and "import oracle" not in code, (f, ln)
                    if "#include" in ln:
                        assert "oracle" not in ln and "_ref" not in ln, (f, ln)
                assert "libhadi_oracle" not in txt and "libhadi_ref" not in txt and "reflib" not in txt
