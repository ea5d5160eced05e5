"""In-tree build of the drop-in Cython module ``nem`` (pangenomenem_b200/dropin/nem*.so).

``from nem import *`` in the reference's ppanggolin.py:20 resolves to this module once
``pangenomenem_b200/dropin`` is on ``sys.path``.  It links libnem_b200.so with
an $ORIGIN-relative rpath, so it carries no CPU implementation of its own."""
from __future__ import annotations

import glob
import os
import subprocess
import sys
import sysconfig

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
DROPIN = os.path.join(PKG, "dropin")


def module_path() -> str | None:
    hits = glob.glob(os.path.join(DROPIN, "nem*" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so")))
    return hits[0] if hits else None


def build(force: bool = False) -> str:
    pyx = os.path.join(DROPIN, "nem.pyx")
    c_file = os.path.join(PKG, "_build", "nem_dropin.c")
    out = os.path.join(DROPIN, "nem" + sysconfig.get_config_var("EXT_SUFFIX"))
    lib = os.path.join(PKG, "libnem_b200.so")
    if not os.path.exists(lib):
        raise RuntimeError("build libnem_b200.so first (python -m pangenomenem_b200.build)")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(
            os.path.getmtime(pyx), os.path.getmtime(os.path.join(ROOT, "include", "nem_b200.h"))):
        return out
    os.makedirs(os.path.dirname(c_file), exist_ok=True)
    subprocess.run([sys.executable, "-m", "cython", "-3", pyx, "-o", c_file], check=True)
    inc = sysconfig.get_paths()["include"]
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-w", "-I", inc, "-I", os.path.join(ROOT, "include"),
                    c_file, "-o", out, "-L", PKG, "-lnem_b200", "-Wl,-rpath,$ORIGIN/.."], check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
