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


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """The Python bindings (capi.py) restate the structs of include/nem_b200.h by hand: compile a
    probe against the header and compare sizes and the offsets of the last fields, so that a field
    added on one side only cannot go unnoticed."""
    import ctypes as C
    import subprocess
    from pangenomenem_b200 import capi
    src = tmp_path / "probe.c"
    src.write_text(r'''
#include "nem_b200.h"
#include <stddef.h>
#include <stdio.h>
int main(void) {
    printf("%zu %zu %zu %zu\n", sizeof(nemb_options), offsetof(nemb_options, beta_mode),
           offsetof(nemb_options, grad_step), offsetof(nemb_options, reserved));
    printf("%zu %zu %zu %zu\n", sizeof(nemb_result), offsetof(nemb_result, n_kept),
           offsetof(nemb_result, beta), offsetof(nemb_result, n_beta_tested));
    printf("%zu %zu %zu %zu\n", sizeof(nem_b200_extra), offsetof(nem_b200_extra, seed),
           offsetof(nem_b200_extra, beta_mode), offsetof(nem_b200_extra, heu_lloss));
    printf("%zu %zu\n", sizeof(nemb_beta_heuristic), sizeof(nemb_batch_stats));
    printf("%zu %zu\n", sizeof(nemb_host_problem), offsetof(nemb_host_problem, disp));
    return 0;
}
''')
    exe = tmp_path / "probe"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    rows = [[int(v) for v in line.split()] for line in
            subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()]
    o, r, e = capi.Options, capi.Result, capi.Extra
    assert rows[0] == [C.sizeof(o), o.beta_mode.offset, o.grad_step.offset, o.reserved.offset]
    assert rows[1] == [C.sizeof(r), r.n_kept.offset, r.beta.offset, r.n_beta_tested.offset]
    assert rows[2] == [C.sizeof(e), e.seed.offset, e.beta_mode.offset, e.heu_lloss.offset]
    assert rows[3] == [C.sizeof(capi.BetaHeuristic), C.sizeof(capi.BatchStats)]
    assert rows[4] == [C.sizeof(capi.HostProblem), capi.HostProblem.disp.offset]


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference on the CPU-runnable config: one JSON line with the keys the
    driver reads (runs the compiled reference when oracle/_ref exists, else the C port)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cp = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference",
                         "--workload", "c1", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, timeout=600)
    assert cp.returncode == 0, cp.stderr[-2000:]
    line = json.loads(cp.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["metric"] == "NEM family-iterations/s" and line["unit"] == "family-iterations/s"
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_built_library_uses_the_blackwell_copy_and_barrier_units():
    """SASS of the in-tree sm_100a library: the density kernel moves its tiles with bulk
    asynchronous copies completing on mbarriers (UBLKCP + SYNCS), the fix-up tail and the finalize
    kernel synchronise their thread-block clusters in hardware (UCGABAR) -- profiles/r1_sass_evidence.txt."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    from pangenomenem_b200 import capi
    sass = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    blocks = sass.split("Function : ")
    def body(tag):
        hits = [b for b in blocks if b.split("\n", 1)[0].startswith(tag)]
        assert hits, tag
        return "\n".join(hits)
    den = body("_Z13k_density_tma")
    assert "UBLKCP" in den and "SYNCS" in den and "POPC" in den
    assert "UCGABAR" in body("_Z18k_sweep_ncem_fixup")
    assert "UCGABAR" in body("_Z23k_mstep_finalize_tables")
