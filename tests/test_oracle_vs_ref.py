"""Pins the oracle against the UNMODIFIED reference binaries in oracle/_ref on a fresh seeded workload
(CPU only; skipped when oracle/_ref has not been built -- it never reads /root/reference at run time)."""
import os, tempfile
import numpy as np
import pytest
from helpers import SynthDB, build_fmi, run_ref_kaiju, Oracle, make_params, have_ref, read_fastq_packed, make_quirk_db, pack_reads

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def work(built):
    d = tempfile.mkdtemp(prefix="kjref_")
    db = SynthDB(3000, 11)
    db.write(d + "/db.faa", d + "/nodes.dmp")
    fmi = build_fmi(d + "/db.faa", d + "/db", threads=4)
    db.write_fastq(21, 0, 3000, 150, True, d + "/r1.fq", d + "/r2.fq")
    db.write_fastq(22, 0, 3000, 100, False, d + "/s.fq")
    return d, fmi


@pytest.mark.parametrize("kw", [dict(mode="mem"), dict(mode="mem", seg=False), dict(mode="greedy"), dict(mode="greedy", e=5, s=50), dict(mode="greedy", e=1, E=1e-6)])
def test_oracle_equals_reference_cli(work, kw):
    d, fmi = work
    orc = Oracle(fmi, d + "/nodes.dmp")
    for fq1, fq2 in ((d + "/r1.fq", d + "/r2.fq"), (d + "/s.fq", None)):
        ref = run_ref_kaiju(d + "/nodes.dmp", fmi, fq1, fq2, threads=4, **kw)
        n1, s1, o1 = read_fastq_packed(fq1)
        s2 = o2 = None
        if fq2:
            _, s2, o2 = read_fastq_packed(fq2)
        tax, best = orc.classify_batch(make_params(**kw), s1, o1, s2, o2)
        for i, nm in enumerate(n1):
            assert (ref[nm][1], ref[nm][2]) == (int(tax[i]), int(best[i])), (nm, ref[nm], tax[i], best[i])


def _write_fasta(path, seq, off):
    with open(path, "w") as f:
        for i in range(len(off) - 1):
            f.write(">r%d\n%s\n" % (i, seq[int(off[i]):int(off[i + 1])].tobytes().decode()))


@pytest.fixture(scope="module")
def work_long(built):
    """Long DNA reads (400 bp - 12 kb, incl. >127-residue low-complexity runs) and protein reads for -p."""
    d = tempfile.mkdtemp(prefix="kjrefl_")
    db = SynthDB(3000, 11)
    db.write(d + "/db.faa", d + "/nodes.dmp")
    fmi = build_fmi(d + "/db.faa", d + "/db", threads=4)
    ls, lo = db.long_reads(31, 0, 300, 400, 12000)
    ps, po = db.protein_reads(32, 0, 1500, 5, 1500)
    _write_fasta(d + "/long.fa", ls, lo); _write_fasta(d + "/prot.fa", ps, po)
    return d, fmi, (ls, lo), (ps, po)


@pytest.mark.parametrize("kw", [dict(mode="mem"), dict(mode="mem", m=8, seg=False), dict(mode="greedy"), dict(mode="greedy", e=5, s=50), dict(mode="greedy", e=2, E=1e-9)])
def test_oracle_equals_reference_long_reads(work_long, kw):
    d, fmi, (ls, lo), _ = work_long
    ref = run_ref_kaiju(d + "/nodes.dmp", fmi, d + "/long.fa", None, threads=8, **kw)
    tax, best = Oracle(fmi, d + "/nodes.dmp").classify_batch(make_params(**kw), ls, lo)
    for i in range(len(lo) - 1):
        assert (ref["r%d" % i][1], ref["r%d" % i][2]) == (int(tax[i]), int(best[i])), (i, ref["r%d" % i], tax[i], best[i])
    assert (tax > 0).sum() > 100


@pytest.mark.parametrize("kw", [dict(mode="mem"), dict(mode="mem", m=6, seg=False), dict(mode="greedy"), dict(mode="greedy", e=4, s=40)])
def test_oracle_equals_reference_protein_input(work_long, kw):
    d, fmi, _, (ps, po) = work_long
    ref = run_ref_kaiju(d + "/nodes.dmp", fmi, d + "/prot.fa", None, threads=8, protein=True, **kw)
    tax, best = Oracle(fmi, d + "/nodes.dmp").classify_batch(make_params(protein=True, **kw), ps, po)
    for i in range(len(po) - 1):
        assert (ref["r%d" % i][1], ref["r%d" % i][2]) == (int(tax[i]), int(best[i])), (i, ref["r%d" % i], tax[i], best[i])
    assert (tax > 0).sum() > 300


def test_lnfact_table_matches_reference(built):
    """ko_lnfact reproduces every entry of the reference's lnfact[0..10000] (blast_seg.c:53-1306), read from oracle/_ref/libkaijuref.so."""
    import ctypes as C
    from helpers import oracle_lib, REF_DIR
    ref = C.CDLL(os.path.join(REF_DIR, "libkaijuref.so"))
    tab = (C.c_double * 10001).in_dll(ref, "lnfact")
    L = oracle_lib(); L.ko_lnfact.restype = C.c_double; L.ko_lnfact.argtypes = [C.c_int]
    assert all(L.ko_lnfact(n) == tab[n] for n in range(10001))


def test_bwtlen_multiple_of_65536(built):
    """SURVEY.md 8a exactness note (a): with bwtlen = m * 2^16 (m >= 2) the reference's FMindex resolves the last 129 positions to the
    index1 row that holds C[] instead of counts and returns values that are too small by a per-letter constant; matches that pass
    through those rows (e.g. every match ending in the last letter of the alphabet) behave differently from a "clean" FM index.
    The oracle reproduces that: function level (FMindex, get_suffix of the reference's own C code) and end to end (CLI)."""
    import ctypes as C
    from helpers import oracle_lib, REF_DIR
    d = tempfile.mkdtemp(prefix="kjq_")
    fmi, nodes, reads = make_quirk_db(d)
    R = C.CDLL(os.path.join(REF_DIR, "libkaijuref.so")); libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p; libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    R.readIndexes.restype = C.c_void_p; R.readIndexes.argtypes = [C.c_void_p]

    class BWT(C.Structure):   # bwt/bwt.h:13-23
        _fields_ = [("len", C.c_long), ("nseq", C.c_int), ("bwt", C.c_void_p), ("alen", C.c_int), ("alphabet", C.c_char_p), ("f", C.c_void_p), ("s", C.c_void_p)]

    class FMI(C.Structure):   # bwt/compactfmi.h:10-19
        _fields_ = [("alen", C.c_int), ("bwtlen", C.c_long), ("bwt", C.c_void_p), ("N1", C.c_int), ("N2", C.c_int), ("index1", C.c_void_p), ("index2", C.c_void_p), ("startLcode", C.c_void_p)]
    b = BWT.from_address(R.readIndexes(libc.fopen(fmi.encode(), b"r"))); fm = FMI.from_address(b.f); n = fm.bwtlen
    assert n == 131072
    R.FMindex.restype = C.c_long; R.FMindex.argtypes = [C.c_void_p, C.c_ubyte, C.c_long]
    R.get_suffix.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_int), C.POINTER(C.c_long)]
    orc = Oracle(fmi, nodes); L = oracle_lib()
    L.ko_fmindex.restype = C.c_int64; L.ko_fmindex.argtypes = [C.c_void_p, C.c_int, C.c_int64]
    L.ko_get_suffix.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    ks = list(range(n - 400, n + 1)) + list(range(65536 - 200, 65536 + 200)) + list(range(0, n, 997))
    for k in ks:
        for c in range(fm.alen):
            assert L.ko_fmindex(orc.idx, c, k) == R.FMindex(b.f, c, k), (k, c)
    iseq = C.c_int(); pos = C.c_long(); oseq = C.c_int(); opos = C.c_int64()
    for k in list(range(n - 300, n)) + list(range(b.nseq, n, 511)):          # rows below nseq are terminator suffixes: never inside a match interval
        R.get_suffix(b.f, b.s, k, C.byref(iseq), C.byref(pos)); L.ko_get_suffix(orc.idx, k, C.byref(oseq), C.byref(opos))
        assert (iseq.value, pos.value) == (oseq.value, opos.value), k
    with open(d + "/r.fa", "w") as f:
        for i, r in enumerate(reads):
            f.write(">r%d\n%s\n" % (i, r))
    seq, off = pack_reads(reads)
    for kw in (dict(mode="mem"), dict(mode="greedy"), dict(mode="greedy", e=5, s=40), dict(mode="mem", m=5, seg=False)):
        ref = run_ref_kaiju(nodes, fmi, d + "/r.fa", None, threads=4, **kw)
        tax, best = orc.classify_batch(make_params(**kw), seq, off)
        for i in range(len(reads)):
            assert (ref["r%d" % i][1], ref["r%d" % i][2]) == (int(tax[i]), int(best[i])), (kw, i)


def test_reference_reader_rules_for_blank_lines(work):
    """What the device-side parser reproduces (kj_ingest.h, tests/test_gpu_parity.py::test_fastq_with_blank_lines), pinned on the reference itself:
    empty lines before the first record and between FASTQ records are skipped (kaiju.cpp:288-289, 341-348), an empty sequence line inside a
    record is the record's sequence, a missing final newline is fine."""
    d, fmi = work
    lines = open(d + "/r1.fq").read().split("\n")[:4 * 400]; lines2 = open(d + "/r2.fq").read().split("\n")[:4 * 400]
    def dirty(ls):
        out = ["", ""]                                                   # leading blank lines: the file type comes from the first non-empty line
        for r in range(0, len(ls), 4):
            rec = ls[r:r + 4]
            if (r // 4) % 37 == 5:
                rec = [rec[0], "", "+", ""]                               # empty sequence line: a line of the record
            out += rec
            if (r // 4) % 5 == 1:
                out += [""] * (1 + (r // 4) % 3)                          # blank lines between records
        return "\n".join(out)                                            # no newline at the end
    def clean(ls):
        out = []
        for r in range(0, len(ls), 4):
            rec = ls[r:r + 4]
            if (r // 4) % 37 == 5:
                rec = [rec[0], "", "+", ""]
            out += rec
        return "\n".join(out) + "\n"
    for name, fn, src in (("c1", clean, lines), ("c2", clean, lines2), ("b1", dirty, lines), ("b2", dirty, lines2)):
        open(d + "/" + name + ".fq", "w").write(fn(src))
    a = run_ref_kaiju(d + "/nodes.dmp", fmi, d + "/c1.fq", d + "/c2.fq", mode="greedy")
    b = run_ref_kaiju(d + "/nodes.dmp", fmi, d + "/b1.fq", d + "/b2.fq", mode="greedy")
    assert len(a) == 400 and a == b and sum(1 for v in a.values() if v[0] == "C") > 100
    assert run_ref_kaiju(d + "/nodes.dmp", fmi, d + "/c1.fq", mode="mem") == run_ref_kaiju(d + "/nodes.dmp", fmi, d + "/b1.fq", mode="mem")
