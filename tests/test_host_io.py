"""CPU: the host side of the boundary -- NEM file readers (bit packing + CSR) and result
writers of libnem_b200.so, against the file contract of the reference
(readers nem_exe.c:739-898, 973-1091, 1278-1478; SaveResults nem_exe.c:1596-1781; the files
PPanGGOLiN writes, ppanggolin.py:829-930).  No GPU, no compute entry point is called."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_CASES, Golden, make_case


@pytest.fixture(scope="module")
def capi():
    from pangenomenem_b200 import capi as c
    c.load_library()
    return c


def _write(tmp_path, pg, **kw):
    from pangenomenem_b200 import synth
    base = str(tmp_path / "run" / "nem_file")
    synth.write_nem_files(base, pg, **kw)
    return base


@pytest.mark.parametrize("n,d", [(1, 1), (7, 31), (50, 32), (129, 33), (300, 129), (40, 1000)])
def test_dat_packing_bit_exact(tmp_path, capi, oracle, n, d):
    pg = make_case(n, d, seed=n + d, graph="chain" if n > 1 else "none")
    base = _write(tmp_path, pg, spatial=n > 1)
    hp = capi.read_files(base, k=3)
    assert (hp["n"], hp["d"]) == (n, d)
    assert hp["wpr"] % 4 == 0 and hp["wpr"] * 32 >= d
    assert np.array_equal(hp["x_packed"], oracle.pack(pg.x, hp["wpr"]))       # bit-exact
    from pangenomenem_b200 import synth
    assert np.array_equal(hp["x_packed"], synth.pack_rows(pg.x, hp["wpr"]))


@pytest.mark.parametrize("graph,weighted", [("pangenome", True), ("random", True), ("chain", False)])
def test_nei_to_csr_bit_exact(tmp_path, capi, graph, weighted):
    pg = make_case(2000, 40, seed=5, graph=graph, weighted=weighted)
    base = _write(tmp_path, pg, weighted_flag=1 if weighted else 0)
    hp = capi.read_files(base, k=0)
    assert hp["spatial"] and hp["nnz"] == pg.col.shape[0]
    assert np.array_equal(hp["row_ptr"], pg.row_ptr)
    assert np.array_equal(hp["col"], pg.col)                 # file order inside a row is kept
    assert np.array_equal(hp["wgt"], pg.wgt if weighted else np.ones_like(pg.wgt))
    assert hp["max_neigh"] == int(np.diff(pg.row_ptr).max())


def test_default_m_file_as_the_reference_reads_it(tmp_path, capi, oracle):
    pg = make_case(30, 17, seed=1, graph="chain")
    base = _write(tmp_path, pg)
    hp = capi.read_files(base, k=3)
    prop, center, disp = oracle.default_theta(3, 17)
    assert hp["m_flag"] == 1
    # last proportion = 1 - sum of the others, in float (nem_exe.c:1022-1034)
    assert np.array_equal(hp["prop"], prop)
    assert np.array_equal(hp["center"], center) and np.array_equal(hp["disp"], disp)


def test_comments_isolated_points_out_of_range_and_zero_weights(tmp_path, capi):
    base = str(tmp_path / "nem_file")
    with open(base + ".str", "w") as f:
        f.write("# a comment line\n# another\nS 4 3\n")            # lib_io.c:32-87
    with open(base + ".dat", "w") as f:
        f.write("1 0 1\n0\t0\t1\n1 1 1\n0 1 0\n")
    with open(base + ".nei", "w") as f:
        # weighted; point 3 absent => no neighbours; neighbour 9 out of range (nem_exe.c:1416-1420)
        # and a zero weight (nem_exe.c:1441-1445) are dropped
        f.write("1\n1 2 2 9 0.5 1.5\n2 2 1 4 2 0\n4 1 2 3.25\n")
    hp = capi.read_files(base, k=0)
    assert hp["x_packed"][:, 0].tolist() == [0b101, 0b100, 0b111, 0b010]
    assert hp["row_ptr"].tolist() == [0, 1, 2, 2, 3]
    assert hp["col"].tolist() == [1, 0, 1]
    assert hp["wgt"].tolist() == [0.5, 2.0, 3.25]


def test_type_n_has_no_graph(tmp_path, capi):
    pg = make_case(64, 20, seed=2, graph="none")
    base = _write(tmp_path, pg, spatial=False)
    hp = capi.read_files(base, k=3)
    assert not hp["spatial"] and hp["nnz"] == 0 and "row_ptr" not in hp


@pytest.mark.parametrize("breakage,code", [
    ("short_dat", 3), ("missing_dat", 3), ("bad_str", 3), ("m_too_few", 3), ("m_neg_disp", 3),
    ("m_bad_prop", 3)])
def test_malformed_files_are_rejected(tmp_path, capi, breakage, code):
    """ReadMatrixFile short-file detection (nem_exe.c:883-894), .m value count and sign checks
    (nem_exe.c:994-1085): STS_E_FILE, never a crash."""
    pg = make_case(20, 6, seed=3, graph="chain")
    base = _write(tmp_path, pg)
    if breakage == "short_dat":
        txt = open(base + ".dat").read()
        open(base + ".dat", "w").write(txt[: len(txt) // 2])
    elif breakage == "missing_dat":
        os.remove(base + ".dat")
    elif breakage == "bad_str":
        open(base + ".str", "w").write("S twenty 6\n")
    elif breakage == "m_too_few":
        open(base + ".m", "w").write("1 0.3 0.3 1 1 1")
    elif breakage == "m_neg_disp":
        open(base + ".m", "w").write("1 0.3 0.3 " + " ".join(["1"] * 18) + " " + " ".join(["-0.1"] * 18))
    elif breakage == "m_bad_prop":
        open(base + ".m", "w").write("1 0.7 0.7 " + " ".join(["1"] * 18) + " " + " ".join(["0.1"] * 18))
    with pytest.raises(capi.NemError) as ei:
        capi.read_files(base, k=3)
    assert ei.value.code == code


@pytest.mark.parametrize("name", [n for n in GOLDEN_CASES if Golden(n).text("uf")])
def test_writers_reproduce_the_reference_text(tmp_path, capi, name):
    """nemb_write_uf / nemb_write_mf given the reference's own numbers produce the reference's
    own .uf / .mf bytes (printf formats of nem_exe.c:1677, 1708-1773)."""
    g = Golden(name)
    uf_ref, mf_ref = g.text("uf"), g.text("mf")
    t = np.array(uf_ref.split(), dtype=np.float32).reshape(g.n, 3)
    capi.write_uf(str(tmp_path / "o.uf"), t)
    assert open(tmp_path / "o.uf").read() == uf_ref
    # .mf: criteria line is float32 %g of the reference's float accumulators; feed them back
    lines = mf_ref.split("\n")
    vals = lines[2].split()
    crit = {k: float(v) for k, v in zip("UDLM", vals[:4])}
    from pangenomenem_b200 import synth
    mf = synth.read_mf_text(mf_ref, 3, g.d)
    capi.write_mf(str(tmp_path / "o.mf"), crit, g.opt["beta"] if g.spatial else 0.0, mf["p"],
                  mf["mu"], mf["eps"])
    got = open(tmp_path / "o.mf").read().split("\n")
    assert len(got) == len(lines)
    for i, (a, b) in enumerate(zip(got, lines)):
        if i == 2:
            assert a.split()[:4] == b.split()[:4]         # U D L M (err column is nan/-nan)
        else:
            assert a == b, (i, a[:80], b[:80])


def test_cf_writer(tmp_path, capi):
    capi.write_cf(str(tmp_path / "o.cf"), np.array([0, 2, 1, 1], dtype=np.int32))
    assert open(tmp_path / "o.cf").read() == "1 3 2 2 \n"   # nem_exe.c:1683-1700, 1-based


def test_golden_input_files_round_trip(tmp_path, capi, oracle):
    """The loader on the very files the reference consumed."""
    g = Golden("ppanggolin_ncem_sk")
    base = str(tmp_path / "g" / "nem_file")
    g.write_files(base)
    hp = capi.read_files(base, k=3)
    assert np.array_equal(hp["x_packed"], oracle.pack(g.x, hp["wpr"]))
    assert np.array_equal(hp["row_ptr"], g.row_ptr) and np.array_equal(hp["col"], g.col)
    assert np.array_equal(hp["wgt"], g.wgt)


def test_dat_fast_path_validates_every_byte(tmp_path, capi):
    """The vectorised fixed-layout reader (one char per cell + one separator, what
    ppanggolin.py:850 writes) must accept any white-space separator and reject anything else by
    falling back to the tokenizer, which reports the offending value (multi-threaded: n >= 4096)."""
    from pangenomenem_b200 import synth
    pg = make_case(5000, 70, seed=4, graph="none")
    base = _write(tmp_path, pg, spatial=False)
    want = synth.pack_rows(pg.x, capi.read_files(base, k=0)["wpr"])
    raw = bytearray(open(base + ".dat", "rb").read())
    stride = 2 * pg.d
    for r, c in [(0, 0), (17, 31), (4999, 69), (2500, 64)]:      # separators -> space / CR
        raw[r * stride + 2 * c + 1] = ord(" ") if c % 2 else ord("\r")
    open(base + ".dat", "wb").write(bytes(raw))
    assert np.array_equal(capi.read_files(base, k=0)["x_packed"], want)
    for pos, ch in [((3000, 40), b"2"), ((4100, 3), b"x"), ((10, 69), b"\t")]:
        bad = bytearray(raw)
        bad[pos[0] * stride + 2 * pos[1]] = ch[0]                # a cell that is not 0 / 1
        open(base + ".dat", "wb").write(bytes(bad))
        with pytest.raises(capi.NemError) as ei:
            capi.read_files(base, k=0)
        assert ei.value.code == 3
    bad = bytearray(raw)
    bad[1234 * stride + 2 * 33 + 1] = ord("1")                   # a separator that is a digit
    open(base + ".dat", "wb").write(bytes(bad))
    with pytest.raises(capi.NemError):
        capi.read_files(base, k=0)


def test_nei_threaded_reader_matches_the_sequential_semantics(tmp_path, capi):
    """.nei files above 64 KB are parsed line by line on several threads (nem_io.c
    nei_read_lines); whatever the fast path cannot prove line-structured goes to the sequential
    token reader (the reference's fscanf semantics, ReadPtsNeighs nem_exe.c:1342-1478).  Shuffled
    records, repeated records (the last one wins), a record broken over two lines and a short
    record must all behave as before."""
    pg = make_case(20000, 12, seed=6, graph="pangenome")
    base = _write(tmp_path, pg)
    assert os.path.getsize(base + ".nei") > (1 << 16)
    hp = capi.read_files(base, k=0)
    assert np.array_equal(hp["row_ptr"], pg.row_ptr) and np.array_equal(hp["col"], pg.col)
    assert np.array_equal(hp["wgt"], pg.wgt)
    lines = open(base + ".nei").read().split("\n")
    head, recs = lines[0], [l for l in lines[1:] if l]
    rng = np.random.default_rng(0)
    shuffled = [recs[i] for i in rng.permutation(len(recs))]
    open(base + ".nei", "w").write("\n".join([head] + shuffled) + "\n")
    hp = capi.read_files(base, k=0)
    assert np.array_equal(hp["row_ptr"], pg.row_ptr) and np.array_equal(hp["col"], pg.col)
    # a stale first record for three points: the later (real) record wins, wherever the threads cut
    stale = ["5\t1\t6\t9", "12000\t2\t1\t2\t3\t3", "19999\t0"]
    open(base + ".nei", "w").write("\n".join([head] + stale + ["", "  "] + recs) + "\n")
    hp = capi.read_files(base, k=0)
    assert np.array_equal(hp["row_ptr"], pg.row_ptr) and np.array_equal(hp["wgt"], pg.wgt)
    # one record broken over two lines: not line-structured -> sequential reader, same CSR
    k = 7000
    toks = recs[k].split("\t")
    broken = recs[:k] + ["\t".join(toks[:3]), "\t".join(toks[3:])] + recs[k + 1:]
    open(base + ".nei", "w").write("\n".join([head] + broken) + "\n")
    hp = capi.read_files(base, k=0)
    assert np.array_equal(hp["row_ptr"], pg.row_ptr) and np.array_equal(hp["col"], pg.col)
    # a record that announces more neighbours than it lists swallows the next line's tokens in
    # the reference too; here it ends the file early -> error, not a crash
    toks = recs[-1].split("\t")
    short = recs[:-1] + ["\t".join(toks[:2] + toks[2:3])]
    if int(toks[1]) > 1:
        open(base + ".nei", "w").write("\n".join([head] + short) + "\n")
        with pytest.raises(capi.NemError) as ei:
            capi.read_files(base, k=0)
        assert ei.value.code == 3


def test_uf_writer_rounds_like_printf(tmp_path, capi):
    """The .uf writer formats " %5.3f " without printf (nem_exe.c:1677): same bytes as printf on
    random posteriors, on every k/2000 tie candidate and its float neighbours, and at the ends."""
    rng = np.random.default_rng(0)
    grid = (np.arange(0, 2001, dtype=np.float64) / 2000).astype(np.float32)
    v = np.concatenate([
        rng.random(60000, dtype=np.float32), grid, np.nextafter(grid, np.float32(0)),
        np.nextafter(grid, np.float32(2)),
        np.array([0, 1, 1e-45, 1e-30, 0.0005, 0.00049999, 0.9995, 0.99949999, 0.99999994, 0.0625,
                  0.1875, 0.3125], dtype=np.float32)])
    v = v[:len(v) // 3 * 3].reshape(-1, 3)
    capi.write_uf(str(tmp_path / "o.uf"), v)
    got = open(tmp_path / "o.uf").read().split("\n")
    exp = ["".join(" %5.3f " % float(x) for x in r) for r in v]
    assert got[:len(exp)] == exp


def test_str_sizes_beyond_int_are_an_error_not_a_truncation(tmp_path, capi):
    """N and D are `int` in the engine's ABI like in the reference (nem_exe.h:23-35); ReadStrFile
    (nem_exe.c:739-898) would store a truncated count.  Here a header beyond INT_MAX is E_FILE."""
    base = str(tmp_path / "big")
    for head in ["S 4294967301 5\n", "N 10 99999999999\n", "I 70000 70000 3\n", "S -4 5\n"]:
        open(base + ".str", "w").write(head)
        open(base + ".dat", "w").write("0 1 0 1 0\n")
        with pytest.raises(capi.NemError) as ei:
            capi.read_files(base, k=0)
        assert ei.value.code == 3, head
