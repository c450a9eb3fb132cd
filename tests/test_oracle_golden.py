"""The oracle (oracle/kaiju_oracle.c) against the committed reference outputs in tests/golden/ (CPU only)."""
import json, os
import numpy as np
import pytest
from conftest import GOLDEN_CONFIGS, GOLD
from helpers import Oracle, make_params, oracle_lib
import ctypes as C


@pytest.fixture(scope="module")
def orc(golden, built):
    return Oracle(golden.fmi, golden.nodes)


@pytest.mark.parametrize("cfg", sorted(GOLDEN_CONFIGS))
@pytest.mark.parametrize("tag", ["pe150", "se100"])
def test_oracle_matches_reference_output(orc, golden, cfg, tag):
    names, s1, o1, s2, o2 = golden.reads(tag)
    P = make_params(**GOLDEN_CONFIGS[cfg])
    tax, best = orc.classify_batch(P, s1, o1, s2, o2)
    etax, ebest, eids = golden.expected(cfg, tag)
    assert np.array_equal(tax, etax), "taxon mismatch at reads %s" % [names[i] for i in np.nonzero(tax != etax)[0][:5]]
    assert np.array_equal(best, ebest)
    # id sets (column 5 of -v) on the classified reads
    for i in np.nonzero(etax)[0][:400]:
        a = bytes(s1[int(o1[i]):int(o1[i + 1])]); b = bytes(s2[int(o2[i]):int(o2[i + 1])]) if s2 is not None else None
        t, bst, ids = orc.classify_one(P, a, b)
        assert tuple(ids) == eids[i], names[i]


def test_fmindex_and_get_suffix_known_answers(orc):
    k = np.load(os.path.join(GOLD, "fmindex_kat.npz"))
    L = oracle_lib()
    assert L.ko_index_bwtlen(orc.idx) == int(k["bwtlen"]) and L.ko_index_alen(orc.idx) == int(k["alen"])
    for row, pos in zip(k["fmindex"], k["ks"]):
        for c in range(int(k["alen"])):
            assert L.ko_fmindex(orc.idx, c, int(pos)) == int(row[c])
    iseq = C.c_int(); p = C.c_int64()
    for (es, ep), r in zip(k["suffix"], k["rows"]):
        L.ko_get_suffix(orc.idx, int(r), C.byref(iseq), C.byref(p))
        assert (iseq.value, p.value) == (int(es), int(ep))


def test_seg_and_lnfact_known_answers(built):
    L = oracle_lib()
    kat = json.load(open(os.path.join(GOLD, "seg_kat.json")))
    for n, v in enumerate(kat["lnfact"]):
        assert L.ko_lnfact(n) == v, n                      # bit-identical to the reference's literals
    left = (C.c_int * 64)(); right = (C.c_int * 64)()
    hits = 0
    for s, reg in zip(kat["seqs"], kat["regions"]):
        n = L.ko_seg(s.encode(), len(s), left, right, 64)
        assert [[left[i], right[i]] for i in range(n)] == reg, s
        hits += bool(reg)
    assert hits > 50


def test_known_answer_fragments():
    """SURVEY.md 8c micro-example measured on the reference (MEM -m 5 -X): 9 fragments incl. the N-codon split."""
    L = oracle_lib()
    read = b"ATGGCCAAGCTGACCNGCGTTGAACGTCTGTAAGGCCCTGCACCAGTTTGATCCGGAATGGCTGAAACGTATCGCC"
    buf = C.create_string_buffer(4096)
    n = L.ko_fragments(read, len(read), 5, buf, 4096)
    frags = buf.raw.split(b"\0")[:n]
    assert sorted(frags, key=lambda f: -len(f)) == sorted([b"ALNVCKALHQFDPEWLKRIA", b"GDTFQPFRIKLVQGLTDVQR", b"AIRFSHSGSNWCRALQTFN", b"RYVSAIPDQTGAGPYRRST",
                                                           b"TSVRPCTSLIRNG", b"SGMAETYR", b"MAKLT", b"GPAPV", b"GQLGH"], key=lambda f: -len(f))
    assert n == 9
