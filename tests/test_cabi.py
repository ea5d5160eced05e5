"""CPU: the C-ABI shared library loads and exports every symbol include/nem_b200.h declares;
argument validation and file errors of nem() (no compute entry point needs a GPU here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, make_case

HEADER = os.path.join(ROOT, "include", "nem_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"typedef\s+void\s*\(\*\w+\)\s*\([^;]*;", "", src, flags=re.S)   # callback typedefs
    return sorted(set(re.findall(r"\b(nem|nem_b200_ex|nemb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    from pangenomenem_b200 import capi
    lib = capi.load_library()
    syms = declared_symbols()
    assert "nem" in syms and "nemb_fit" in syms and len(syms) >= 25
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_signatures_are_plain_c():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    assert "torch" not in src.lower() and "std::" not in src and "template" not in src
    assert 'extern "C"' in src


def test_version_and_cuda_linkage():
    from pangenomenem_b200 import capi
    lib = capi.load_library()
    assert b"nem-b200" in lib.nemb_version()
    # the kernels are in the library (sm_100a cubin embedded), not behind a Python fallback
    blob = open(capi.LIB_PATH, "rb").read()
    assert b"k_density" in blob and b"sm_100a" in blob


NEM_OK_ARGS = dict(nk=3, algo=b"ncem", beta=0.5, convergence=b"clas", convergence_th=1e-8,
                   format=b"fuzzy", it_max=100, dolog=False, model_family=b"bern",
                   proportion=b"pk", dispersion=b"sk_", init_mode=2)


@pytest.mark.parametrize("override", [
    dict(nk=0), dict(nk=17), dict(algo=b"xyz"), dict(algo=b"gem"), dict(convergence=b"maybe"),
    dict(convergence_th=0.0), dict(format=b"soft"), dict(it_max=-1), dict(model_family=b"norm"),
    dict(proportion=b"pp"), dict(dispersion=b"s"), dict(init_mode=0), dict(init_mode=4)])
def test_nem_rejects_bad_arguments_before_touching_files(tmp_path, override):
    """EXIT_E_ARGS = 2 (lib_io.h:22-34).  The reference overwrites its own error flag
    (nem_exe.c:371-431 vs 472); the replacement validates first."""
    from pangenomenem_b200 import capi
    kw = dict(NEM_OK_ARGS, **override)
    assert capi.nem(Fname=str(tmp_path / "does_not_exist").encode(), **kw) == 2
    assert not os.path.exists(tmp_path / "does_not_exist.uf")


def test_nem_missing_files_is_a_file_error(tmp_path):
    from pangenomenem_b200 import capi
    assert capi.nem(Fname=str(tmp_path / "nothing").encode(), **NEM_OK_ARGS) == 3


def test_nem_without_a_gpu_fails_loudly_and_writes_nothing(tmp_path):
    """No CPU fallback: with valid inputs and no CUDA device nem() returns EXIT_E_SYSTEM (5)
    and leaves no .uf/.mf behind (ppanggolin.py:1886-1889 then reports every family 'U')."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from pangenomenem_b200 import capi, synth
    pg = make_case(200, 12, seed=1)
    base = str(tmp_path / "nem_file")
    synth.write_nem_files(base, pg)
    assert capi.nem(Fname=base.encode(), **NEM_OK_ARGS) == 5
    assert not os.path.exists(base + ".uf") and not os.path.exists(base + ".mf")
    with pytest.raises(capi.NemError):
        capi.Engine(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pangenomenem_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "nem_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_cython_dropin_module_has_the_reference_interface(tmp_path):
    """The top-level module `nem` PPanGGOLiN imports (ppanggolin.py:20, NEM/nem.pyx:1-14):
    keyword-callable with bytes strings, returns the ExitET int."""
    import importlib
    import sys
    from pangenomenem_b200 import build_pyx
    so = build_pyx.build()
    sys.path.insert(0, os.path.dirname(so))
    try:
        mod = importlib.import_module("nem")
        assert os.path.dirname(mod.__file__) == os.path.dirname(so)
        rc = mod.nem(Fname=str(tmp_path / "nothing").encode("ascii") + b"/nem_file", nk=3, algo=b"ncem",
                     beta=0.5, convergence=b"clas", convergence_th=0.00000001, format=b"fuzzy",
                     it_max=100, dolog=True, model_family=b"bern", proportion=b"pk",
                     dispersion=b"sk_", init_mode=2)
        assert rc == 3
        with pytest.raises(TypeError):
            mod.nem(Fname="a str is not bytes", nk=3, algo=b"ncem", beta=0.5, convergence=b"clas",
                    convergence_th=1e-8, format=b"fuzzy", it_max=1, dolog=False,
                    model_family=b"bern", proportion=b"pk", dispersion=b"sk_", init_mode=2)
    finally:
        sys.path.pop(0)
        sys.modules.pop("nem", None)
