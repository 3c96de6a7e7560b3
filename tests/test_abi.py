"""CPU-side checks of the drop-in boundary: the library builds/loads and exports exactly the symbols
include/sla_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sla_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sla_[A-Za-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib(sla):
    if not os.path.exists(sla.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return sla.load()


def test_header_and_binding_agree(sla):
    assert header_symbols() == sorted(sla.SIGNATURES.keys())


def test_library_exports_every_declared_symbol(lib):
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_workspace_queries(lib):
    assert lib.sla_version() >= 100
    b1 = lib.sla_workspace_bytes(0, 256, 1)
    b64 = lib.sla_workspace_bytes(0, 256, 64)
    assert b1 >= 4 * 256 * 256 * 8 and b64 >= 64 * 4 * 256 * 256 * 8
    assert lib.sla_workspace_bytes(1, 256, 64) > b64            # complex needs more
    assert lib.sla_workspace_bytes(0, 0, 4) == 0
    assert lib.sla_greens_chain_workspace_bytes(0, 256, 64, 200, 10) > b64
    assert lib.sla_greens_chain_workspace_bytes(0, 256, 64, 205, 10) == 0   # L % n_stab != 0


def test_argument_errors_without_gpu(lib):
    # argument validation happens before any CUDA call
    assert lib.sla_set_stream(None, None) == -1
    assert lib.sla_ldr(None, 0, 4, 1, None, None, None, None) == -1


def test_no_cpu_fallback(sla):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(Exception):
        sla.api._Handle.current()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "stablelinearalgebra.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f


# ---------------------------------------------------------------------------------------------
# a plain-C client (tests/abi_c/abi_check.c): the header is usable from C, and the calls a Julia `ccall` would make
# -- same argument types -- give right answers on the GPU
# ---------------------------------------------------------------------------------------------
def _build_c_client(sla, tmpdir):
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("CUDA headers not available")
    exe = os.path.join(str(tmpdir), "abi_check")
    libdir = os.path.dirname(sla.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-O1", os.path.join(ROOT, "tests", "abi_c", "abi_check.c"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           "-L", libdir, "-lsla_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_client_compiles_and_links(sla, lib, tmp_path):
    _build_c_client(sla, tmp_path)


@pytest.mark.gpu
def test_c_client(sla, lib, tmp_path):
    import subprocess
    exe = _build_c_client(sla, tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ABI_C_OK" in r.stdout, (r.returncode, r.stdout, r.stderr)
