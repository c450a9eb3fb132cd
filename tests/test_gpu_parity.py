"""GPU parity: the CUDA path called through the C ABI (libkaijub200.so) vs the committed reference outputs, the oracle on
fresh seeded workloads, edge cases, and size-independent properties at large batch sizes.  Bit-exact (integer work)."""
import ctypes as C
import os, tempfile
import numpy as np
import pytest
from conftest import GOLDEN_CONFIGS
from helpers import Oracle, make_params, SynthDB, build_fmi, have_ref

pytestmark = pytest.mark.gpu


def kb_params(kb, kw):
    return kb.make_params(kw.get("mode", "mem"), m=kw.get("m", 11), e=kw.get("e", 3), s=kw.get("s", 65), E=kw.get("E", 0.01), seg=kw.get("seg", True))


@pytest.fixture(scope="module")
def kb(built):
    import kaiju_b200
    return kaiju_b200


@pytest.fixture(scope="module")
def gclf(kb, golden):
    c = kb.Classifier(golden.fmi, golden.nodes, device=0, params=kb.make_params("mem"))
    yield c
    c.close()


@pytest.mark.parametrize("cfg", sorted(GOLDEN_CONFIGS))
@pytest.mark.parametrize("tag", ["pe150", "se100"])
def test_gpu_matches_reference_golden(kb, gclf, golden, cfg, tag):
    names, s1, o1, s2, o2 = golden.reads(tag)
    gclf.set_params(kb_params(kb, GOLDEN_CONFIGS[cfg]))
    tax, best = gclf.classify(s1, o1, s2, o2)
    etax, ebest, _ = golden.expected(cfg, tag)
    bad = np.nonzero((tax != etax) | (best != ebest))[0]
    assert len(bad) == 0, [(names[i], int(tax[i]), int(etax[i]), int(best[i]), int(ebest[i])) for i in bad[:5]]
    assert gclf.kernel_launches > 0


@pytest.fixture(scope="module")
def fresh(kb):
    if not have_ref():
        pytest.skip("oracle/_ref (index builder) not available")
    d = tempfile.mkdtemp(prefix="kjgpu_")
    db = SynthDB(20000, 5); db.write(d + "/db.faa", d + "/nodes.dmp")
    fmi = build_fmi(d + "/db.faa", d + "/db", threads=min(16, os.cpu_count()))
    return db, fmi, d + "/nodes.dmp"


@pytest.mark.parametrize("kw", [dict(mode="mem"), dict(mode="mem", seg=False), dict(mode="mem", m=14), dict(mode="greedy"),
                                dict(mode="greedy", e=5), dict(mode="greedy", e=2, s=45, seg=False), dict(mode="greedy", E=1e-8)])
def test_gpu_matches_oracle_on_fresh_workload(kb, fresh, kw):
    db, fmi, nodes = fresh
    orc = Oracle(fmi, nodes)
    clf = kb.Classifier(fmi, nodes, device=0, params=kb_params(kb, kw))
    for paired, rl in ((True, 150), (False, 100), (True, 75)):
        s1, o1, s2, o2 = db.reads(31 + rl, 0, 6000, rl, paired)
        otax, obest = orc.classify_batch(make_params(**kw), s1, o1, s2, o2)
        tax, best = clf.classify(s1, o1, s2, o2)
        bad = np.nonzero((tax != otax) | (best != obest))[0]
        assert len(bad) == 0, (kw, paired, rl, [(int(i), int(tax[i]), int(otax[i]), int(best[i]), int(obest[i])) for i in bad[:5]])
    clf.close()


def test_edge_cases(kb, gclf, golden):
    gclf.set_params(kb.make_params("mem"))
    # empty batch
    t, b = gclf.classify(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert len(t) == 0
    # ragged batch: empty reads, reads shorter than 3m, one long read at the supported maximum
    rng = np.random.default_rng(3)
    reads = [b"", b"ACGT", b"ACGTACGTACGTACGTACGTACGTACGTACGT", bytes(rng.choice(list(b"ACGT"), 381).astype(np.uint8)), b"N" * 200, b"acgtn" * 30]
    s = np.frombuffer(b"".join(reads), dtype=np.uint8); o = np.zeros(len(reads) + 1, np.uint64); o[1:] = np.cumsum([len(r) for r in reads])
    t, b = gclf.classify(s, o)
    orc = Oracle(golden.fmi, golden.nodes)
    ot, ob = orc.classify_batch(make_params("mem"), s, o)
    assert np.array_equal(t, ot) and np.array_equal(b, ob)
    # a read beyond KJ_MAX_READ_LEN is refused with an error code, not misclassified
    s2 = np.frombuffer(b"A" * 500, dtype=np.uint8); o2 = np.array([0, 500], np.uint64)
    with pytest.raises(kb.KaijuError):
        gclf.classify(s2, o2)


def test_properties_at_scale(kb, fresh):
    """1 M pairs: (a) results do not depend on batch order/chunking (the persistent-kernel scheduler), (b) device-buffer
    and host-buffer entry points agree, (c) duplicated reads get identical results, (d) checksum of a subsample == oracle."""
    import torch
    db, fmi, nodes = fresh
    n = 1 << 20
    s1, o1, s2, o2 = db.reads(77, 0, n, 150, True)
    clf = kb.Classifier(fmi, nodes, device=0, params=kb.make_params("mem"))
    tax, best = clf.classify(s1, o1, s2, o2)
    # (a) reversed order
    perm = np.arange(n)[::-1]
    l1 = np.diff(o1).astype(np.int64); l2 = np.diff(o2).astype(np.int64)
    assert (l1[l1 != 150].size + l2[l2 != 150].size) < n // 100
    def gather(s, o, perm):
        ln = np.diff(o).astype(np.int64)[perm]; no = np.zeros(len(perm) + 1, np.uint64); no[1:] = np.cumsum(ln)
        idx = np.repeat(o[:-1].astype(np.int64)[perm] - no[:-1].astype(np.int64), ln) + np.arange(int(no[-1]))
        return s[idx], no
    rs1, ro1 = gather(s1, o1, perm); rs2, ro2 = gather(s2, o2, perm)
    rtax, rbest = clf.classify(rs1, ro1, rs2, ro2)
    assert np.array_equal(rtax[::-1], tax) and np.array_equal(rbest[::-1], best)
    # (b) device-resident entry point
    d = [torch.from_numpy(x.view(np.int64) if x.dtype == np.uint64 else x).cuda() for x in (s1, o1, s2, o2)]
    dt = torch.zeros(n, dtype=torch.int64, device="cuda"); dbst = torch.zeros(n, dtype=torch.int32, device="cuda")
    clf.classify_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), n, dt.data_ptr(), dbst.data_ptr())
    torch.cuda.synchronize(); clf.check_errors()
    assert np.array_equal(dt.cpu().numpy().view(np.uint64), tax) and np.array_equal(dbst.cpu().numpy().view(np.uint32), best)
    # (d) oracle on a strided subsample
    orc = Oracle(fmi, nodes); sub = np.arange(0, n, 257)[:3000]
    ss1, so1 = gather(s1, o1, sub); ss2, so2 = gather(s2, o2, sub)
    otax, obest = orc.classify_batch(make_params("mem"), ss1, so1, ss2, so2)
    assert np.array_equal(otax, tax[sub]) and np.array_equal(obest, best[sub])
    assert 0.55 < (tax != 0).mean() < 0.85          # ~70 % of the synthetic reads come from the DB
    clf.close()


def test_wide_index_kernels_and_no_kmer_table(kb, golden, monkeypatch):
    """The 64-bit-interval kernels (indexes >= 2^32 rows) and the table-free chain start, forced on the small golden index."""
    names, s1, o1, s2, o2 = golden.reads("pe150")
    for env in ({"KJ_FORCE_WIDE": "1"}, {"KJ_KMER_K": "0"}, {"KJ_KMER_K": "3"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for cfg in ("mem_default", "greedy_default"):
            clf = kb.Classifier(golden.fmi, golden.nodes, device=0, params=kb_params(kb, GOLDEN_CONFIGS[cfg]))
            tax, best = clf.classify(s1, o1, s2, o2)
            etax, ebest, _ = golden.expected(cfg, "pe150")
            assert np.array_equal(tax, etax) and np.array_equal(best, ebest), (env, cfg)
            clf.close()
        for k in env:
            monkeypatch.delenv(k)


def test_cli_matches_reference_output_format(kb, golden, tmp_path):
    """kaiju-b200 (kj_cli.cpp) on the golden FASTQ files: same C/U lines as the reference, in input order."""
    import gzip, shutil, subprocess
    from conftest import ROOT
    cli = os.path.join(ROOT, "kaiju_b200", "kaiju-b200")
    assert os.path.exists(cli)
    gold = os.path.join(ROOT, "tests", "golden")
    fq = {}
    for name in ("pe150_1", "pe150_2"):
        fq[name] = str(tmp_path / (name + ".fq"))
        with gzip.open(os.path.join(gold, name + ".fq.gz"), "rb") as a, open(fq[name], "wb") as b:
            shutil.copyfileobj(a, b)
    for mode, cfg, extra in (("mem", "mem_default", []), ("greedy", "greedy_default", []), ("mem", "mem_noseg", ["-X"]), ("greedy", "greedy_e5", ["-e", "5"])):
        out = str(tmp_path / (cfg + ".tsv"))
        # gz input for mate 1 (zlib path), plain for mate 2; -z is accepted and ignored
        subprocess.check_call([cli, "-t", golden.nodes, "-f", golden.fmi, "-i", os.path.join(gold, "pe150_1.fq.gz"), "-j", fq["pe150_2"], "-a", mode, "-z", "4", "-v", "-o", out] + extra)
        etax, ebest, _ = golden.expected(cfg, "pe150"); names = golden.reads("pe150")[0]
        lines = open(out).read().splitlines()
        assert len(lines) == len(names)
        for ln, nm, t, b in zip(lines, names, etax, ebest):
            p = ln.split("\t")
            if t:
                assert p[:4] == ["C", nm, str(int(t)), str(int(b))], (ln, nm, t, b)
            else:
                assert p == ["U", nm, "0"], ln
    # error path: -p is refused, missing arguments print usage and exit non-zero
    assert subprocess.call([cli, "-t", golden.nodes, "-f", golden.fmi, "-i", fq["pe150_1"], "-p"], stderr=subprocess.DEVNULL) != 0
    assert subprocess.call([cli, "-f", golden.fmi], stderr=subprocess.DEVNULL) != 0


@pytest.mark.parametrize("cfg", ["mem_default", "mem_m5", "greedy_default", "greedy_e5"])
def test_verbose_id_sets_match_reference_column5(kb, gclf, golden, cfg):
    """kj_classify_verbose: the match-id set (column 5 of `kaiju -v`) incl. the 21-id cap case, ids missing from nodes.dmp, taxon 0."""
    names, s1, o1, s2, o2 = golden.reads("pe150")
    gclf.set_params(kb_params(kb, GOLDEN_CONFIGS[cfg]))
    tax, best, ids = gclf.classify_verbose(s1, o1, s2, o2)
    etax, ebest, eids = golden.expected(cfg, "pe150")
    assert np.array_equal(tax, etax) and np.array_equal(best, ebest)
    for i, nm in enumerate(names):
        assert ids[i] == (eids[i] if etax[i] else ()), (nm, ids[i], eids[i])
    assert max(len(x) for x in ids) == 21            # the capped read is in the fixture
