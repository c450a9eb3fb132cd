"""Pins the oracle against the UNMODIFIED reference binaries in oracle/_ref on a fresh seeded workload
(CPU only; skipped when oracle/_ref has not been built -- it never reads /root/reference at run time)."""
import os, tempfile
import numpy as np
import pytest
from helpers import SynthDB, build_fmi, run_ref_kaiju, Oracle, make_params, have_ref, read_fastq_packed

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
