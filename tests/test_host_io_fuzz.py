"""Host readers (csrc/nem_io.c: the ReadStrFile / ReadMatrixFile / ReadNeiFile / ReadParamFile
replacements, reference nem_exe.c:739-898, 973-1091, 1278-1478) under AddressSanitizer + UBSan on
mutated input files: whatever the bytes, a reader returns a status -- no crash, no hang, no
out-of-range CSR entry.  Regression cases first (found by this fuzzer), then seeded mutations."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pangenomenem_b200", "csrc")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("fuzz") / "harness")
    cmd = ["gcc", "-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-std=gnu11",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           os.path.join(ROOT, "tests", "fuzz_io_harness.c"), os.path.join(CSRC, "nem_io.c"),
           "-o", out, "-lm", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build unavailable: " + r.stderr[-200:])
    return out


def run(harness, base, k=3, threads=1):
    env = dict(os.environ, ASAN_OPTIONS="exitcode=99:detect_leaks=1",
               UBSAN_OPTIONS="halt_on_error=1:exitcode=98")
    p = subprocess.run([harness, base, str(k), str(threads)], capture_output=True, timeout=30, env=env)
    assert p.returncode == 0 and b"runtime error" not in p.stderr and b"Sanitizer" not in p.stderr, \
        (p.returncode, p.stderr[-800:].decode(errors="replace"))
    return dict(l.split(None, 1) for l in p.stdout.decode().splitlines() if " " in l)


def write_valid(base, n, d, k, rng, weighted=True):
    with open(base + ".str", "w") as f:
        f.write(f"S\t{n}\t{d}\n")
    x = (rng.random((n, d)) < 0.4).astype(int)
    with open(base + ".dat", "w") as f:
        for r in x:
            f.write("\t".join(map(str, r)) + "\n")
    with open(base + ".nei", "w") as f:
        f.write("1\n" if weighted else "0\n")
        for i in range(n):
            nb = sorted(set(int(v) for v in rng.integers(1, n + 1, size=int(rng.integers(0, 5)))) - {i + 1})
            line = [str(i + 1), str(len(nb))] + [str(v) for v in nb]
            if weighted:
                line += [str(int(v)) for v in rng.integers(1, 9, size=len(nb))]
            f.write("\t".join(line) + "\n")
    with open(base + ".m", "w") as f:
        f.write("1 " + " ".join(["0.33333"] * (k - 1)) + " "
                + " ".join(["1"] * d + ["0.5"] * d * (k - 2) + ["0"] * d) + " "
                + " ".join(["0.1"] * d * k) + "\n")


def mutate(data, rng):
    b = bytearray(data)
    if not b:
        return bytes(b)
    op, pos = int(rng.integers(0, 8)), int(rng.integers(0, len(b)))
    if op == 0:
        b = b[:pos]
    elif op == 1:
        b[pos] = int(rng.integers(0, 256))
    elif op == 2:
        b[pos:pos] = bytes(rng.integers(32, 127, size=int(rng.integers(1, 20)), dtype=np.uint8).tolist())
    elif op == 3:
        b[pos:pos] = b"99999999999999999999"
    elif op == 4:
        b[pos:pos] = b"-5"
    elif op == 5:
        del b[pos:pos + int(rng.integers(1, 50))]
    elif op == 6:
        b[pos:pos] = b"\n\n\t \t"
    else:
        b[pos:pos] = b"1e400 nan inf 0x10 "
    return bytes(b)


def test_valid_files_read_back(harness, tmp_path):
    rng = np.random.default_rng(5)
    base = str(tmp_path / "f")
    write_valid(base, 40, 33, 3, rng)
    out = run(harness, base, threads=4)
    assert out["str"].split()[0] == "0" and out["dat"] == "0" and out["nei"].split()[0] == "0"
    assert out["m"].split() == ["0", "1"]


def test_neighbour_count_larger_than_the_file(harness, tmp_path):
    """`id nb ...` with an absurd nb: an error, not a doubling loop on an overflowed capacity."""
    rng = np.random.default_rng(6)
    base = str(tmp_path / "f")
    write_valid(base, 4, 5, 3, rng)
    with open(base + ".nei", "w") as f:
        f.write("0\n1\t0\n2\t2\t1\t3\n3\t999999999999999999991\t1\n4\t3\t1\t2\t3\n")
    out = run(harness, base)
    assert out["nei"].split()[0] != "0"


def test_all_records_without_neighbours(harness, tmp_path):
    """every record filtered to zero entries: the per-thread pools stay unallocated (no memcpy from NULL)"""
    rng = np.random.default_rng(7)
    base = str(tmp_path / "f")
    write_valid(base, 3, 5, 3, rng)
    with open(base + ".nei", "w") as f:
        f.write("1\n1\t0\n2\t1\t9\t4\n3\t0\n")     # neighbour 9 is outside 1..3: skipped
    out = run(harness, base)
    assert out["nei"].split()[0] == "0" and out["neisum"].split()[0] == "0"


@pytest.mark.parametrize("seed,big", [(11, False), (12, False), (13, True)])
def test_mutated_files_never_crash(harness, tmp_path, seed, big):
    rng = np.random.default_rng(seed)
    base = str(tmp_path / "f")
    for _ in range(6 if big else 60):
        n = int(rng.integers(3000, 6000)) if big else int(rng.integers(1, 70))
        d = int(rng.choice([1, 7, 32, 33]) if big else rng.choice([1, 5, 31, 32, 33, 64, 100, 130]))
        write_valid(base, n, d, 3, rng, weighted=bool(rng.integers(0, 2)))
        ext = str(rng.choice([".str", ".dat", ".nei", ".m"]))
        with open(base + ext, "rb") as f:
            data = f.read()
        for _ in range(int(rng.integers(1, 4))):
            data = mutate(data, rng)
        with open(base + ext, "wb") as f:
            f.write(data)
        run(harness, base, threads=int(rng.choice([1, 4])))
