"""The name-reporting front-ends: `kaiju-b200 -M kaijux|kaijup` against the committed outputs of the reference's own
kaijux / kaijup binaries (tests/golden/make_golden_xp.py), byte for byte."""
import gzip, os, shutil, subprocess
import pytest
from conftest import ROOT, GOLD

pytestmark = pytest.mark.gpu
XP = {"mem_default": ["-a", "mem"], "mem_m5_noseg": ["-a", "mem", "-m", "5", "-X"], "greedy_default": ["-a", "greedy", "-e", "3", "-s", "65"],
      "greedy_e5_s40": ["-a", "greedy", "-e", "5", "-s", "40"], "greedy_e0": ["-a", "greedy", "-e", "0"]}
CLI = os.path.join(ROOT, "kaiju_b200", "kaiju-b200")


def plain(src, dst):
    with gzip.open(src, "rb") as f, open(dst, "wb") as g:
        shutil.copyfileobj(f, g)
    return dst


def expected(name):
    return gzip.open(os.path.join(GOLD, name), "rb").read().decode()


@pytest.mark.parametrize("cfg", sorted(XP))
@pytest.mark.parametrize("tag", ["se100", "pe150"])
def test_kaijux_equals_reference(built, tmp_path, cfg, tag):
    d = str(tmp_path)
    inp = ["-i", os.path.join(GOLD, "se100.fq.gz")] if tag == "se100" else ["-i", plain(GOLD + "/pe150_1.fq.gz", d + "/a.fq"), "-j", os.path.join(GOLD, "pe150_2.fq.gz")]
    out = subprocess.run([CLI, "-M", "kaijux", "-f", os.path.join(GOLD, "db.fmi")] + inp + XP[cfg], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout.decode()
    want = expected("expected_x_%s_%s.tsv.gz" % (cfg, tag))
    assert out == want, [(a, b) for a, b in zip(out.split("\n"), want.split("\n")) if a != b][:3]
    assert want.count("\nC\t") > 500


@pytest.mark.parametrize("cfg", sorted(XP))
def test_kaijup_equals_reference(built, tmp_path, cfg):
    out = subprocess.run([CLI, "-M", "kaijup", "-f", os.path.join(GOLD, "db.fmi"), "-i", os.path.join(GOLD, "prot.fa.gz"), "-o", str(tmp_path / "o.tsv")] + XP[cfg],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    got = open(str(tmp_path / "o.tsv")).read(); want = expected("expected_p_%s.tsv.gz" % cfg)
    assert got == want, [(a, b) for a, b in zip(got.split("\n"), want.split("\n")) if a != b][:3]


@pytest.mark.parametrize("cfg", ["mem_default", "mem_m5_noseg", "greedy_default", "greedy_e5_s40"])
@pytest.mark.parametrize("tag", ["se100", "pe150"])
def test_verbose_output_all_seven_columns_equal_reference(built, tmp_path, cfg, tag):
    """`kaiju-b200 -v` == `kaiju -v` (raw reference output, tests/golden/make_golden_xp.py --v7): best length / score, taxon-id set, accession
    set and the matched fragment strings of every classified read."""
    d = str(tmp_path)
    inp = ["-i", os.path.join(GOLD, "se100.fq.gz")] if tag == "se100" else ["-i", plain(GOLD + "/pe150_1.fq.gz", d + "/a.fq"), "-j", os.path.join(GOLD, "pe150_2.fq.gz")]
    out = subprocess.run([CLI, "-v", "-t", os.path.join(GOLD, "nodes.dmp"), "-f", os.path.join(GOLD, "db.fmi")] + inp + XP[cfg], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout.decode()
    want = expected("expected_v7_%s_%s.tsv.gz" % (cfg, tag))
    assert out == want, [(a, b) for a, b in zip(out.split("\n"), want.split("\n")) if a != b][:3]
