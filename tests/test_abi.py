"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/kokkos_b200.h declares (no compute calls without a GPU), errors are reported loudly,
and the Python host mirror fails -- never falls back -- when no B200 is present."""
import ctypes
import os
import re

import pytest

import kokkos_b200 as kb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kokkos_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path_entry_points():
    syms = declared_symbols()
    assert len(syms) >= 50
    for s in ("b200_init", "b200_fence", "b200_malloc", "b200_reduce_sum_f64", "b200_reduce_minmaxloc_f64",
              "b200_scan_excl_i64", "b200_stream_triad_f64", "b200_stencil7_minmaxloc_f64", "b200_gups_add_i64",
              "b200_spmv_crs_f64", "b200_launch", "b200_occupancy", "b200_scratch_get"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(kb.LIB_PATH), "libkokkos_b200.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(kb.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_signatures_cover_the_library():
    lib = kb.load_library()
    assert b"kokkos_b200" in lib.b200_version()


def test_null_instance_is_an_error_not_a_crash():
    lib = kb.load_library()
    assert lib.b200_fence(None, b"x") == -2
    assert b"not initialised" in lib.b200_last_error_string()
    out = ctypes.c_double()
    assert lib.b200_reduce_sum_f64(None, None, 10, ctypes.byref(out), None) == -2
    assert lib.b200_scan_excl_i64(None, None, None, 10, 0, None, None) == -2


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here; the no-GPU failure mode is tested on the CPU box")
    with pytest.raises(kb.B200Error):
        kb.B200(0)


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under kokkos_b200/ or include/ may name it."""
    bad = []
    for base in ("kokkos_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in d.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".h", ".hpp", "Makefile")):
                    txt = open(os.path.join(d, f), errors="replace").read()
                    if re.search(r"oracle[/_.]|liboracle|kokkos_ref_omp", txt):
                        bad.append(os.path.join(d, f))
    assert not bad, bad


def test_kokkos_arms_library_exports_every_entry_the_binding_uses():
    """benchlib/libkokkos_arms.so (Kokkos user code on the unmodified reference headers + the Kokkos::B200 adapter) is built only
    where /root/reference exists; where it is present every kka_* entry benchlib/arms.py binds must be exported (no compute here)."""
    import subprocess
    so = os.path.join(ROOT, "benchlib", "libkokkos_arms.so")
    if not os.path.exists(so):
        pytest.skip("benchlib/libkokkos_arms.so not built on this machine")
    names = set(re.findall(r"L\.(kka_\w+)", open(os.path.join(ROOT, "benchlib", "arms.py")).read()))
    assert len(names) >= 9
    exported = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    missing = [n for n in sorted(names) if not re.search(rf"\bT {n}\b", exported)]
    assert not missing, missing
    # the adapter's kernels come from libkokkos_b200.so through the C ABI: the arms library must link it, not re-implement it
    needed = subprocess.run(["readelf", "-d", so], capture_output=True, text=True, check=True).stdout
    assert "libkokkos_b200.so" in needed
