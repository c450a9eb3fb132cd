#!/usr/bin/env python
"""Golden fixtures for the name-reporting front-ends: runs the UNMODIFIED reference kaijux / kaijup (oracle/_ref, built by
oracle/Makefile) on the committed index and read sets (+ a seeded protein read set written here).  Run in the build container:

    make -C oracle ref && python tests/golden/make_golden_xp.py

Outputs: prot.fa.gz (1,200 protein reads incl. split characters), expected_x_<cfg>_<tag>.tsv.gz, expected_p_<cfg>.tsv.gz -- the reference's
output lines as they are (status, read name, best length/score, database sequence names, [empty fragment column])."""
import gzip, os, shutil, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import SynthDB, REF_DIR   # noqa: E402

XP_CONFIGS = {"mem_default": ["-a", "mem"], "mem_m5_noseg": ["-a", "mem", "-m", "5", "-X"], "greedy_default": ["-a", "greedy", "-e", "3", "-s", "65"],
              "greedy_e5_s40": ["-a", "greedy", "-e", "5", "-s", "40"], "greedy_e0": ["-a", "greedy", "-e", "0"]}


def plain(src, dst):
    with gzip.open(src, "rb") as f, open(dst, "wb") as g:
        shutil.copyfileobj(f, g)
    return dst


def main():
    tmp = "/tmp/kj_xp"; os.makedirs(tmp, exist_ok=True)
    fmi = os.path.join(HERE, "db.fmi")
    db = SynthDB(800, 3)
    s, o = db.protein_reads(42, 0, 1200, 5, 400)
    with gzip.open(os.path.join(HERE, "prot.fa.gz"), "wt") as f:
        for i in range(1200):
            f.write(">q%d extra words\n%s\n" % (i, bytes(s[int(o[i]):int(o[i + 1])]).decode()))
    inputs = {"se100": ["-i", plain(HERE + "/se100.fq.gz", tmp + "/se.fq")],
              "pe150": ["-i", plain(HERE + "/pe150_1.fq.gz", tmp + "/a.fq"), "-j", plain(HERE + "/pe150_2.fq.gz", tmp + "/b.fq")]}
    prot = plain(HERE + "/prot.fa.gz", tmp + "/prot.fa")
    for cfg, flags in XP_CONFIGS.items():
        for tag, inp in inputs.items():
            out = subprocess.run([os.path.join(REF_DIR, "kaijux"), "-f", fmi, "-z", "1"] + inp + flags, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            with gzip.open(os.path.join(HERE, "expected_x_%s_%s.tsv.gz" % (cfg, tag)), "wb") as g:
                g.write(out)
        out = subprocess.run([os.path.join(REF_DIR, "kaijup"), "-f", fmi, "-z", "1", "-i", prot] + flags, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
        with gzip.open(os.path.join(HERE, "expected_p_%s.tsv.gz" % cfg), "wb") as g:
            g.write(out)
        print(cfg, "x/p outputs written; kaijup classified", out.count(b"\nC\t") + out.startswith(b"C\t"), "of 1200")


if __name__ == "__main__" and "--v7" not in sys.argv:
    main()


def main_v7():
    """all seven columns of `kaiju -v` (raw reference output) for a few configurations: expected_v7_<cfg>_<tag>.tsv.gz"""
    tmp = "/tmp/kj_xp"; os.makedirs(tmp, exist_ok=True)
    fmi = os.path.join(HERE, "db.fmi"); nodes = os.path.join(HERE, "nodes.dmp")
    inputs = {"se100": ["-i", plain(HERE + "/se100.fq.gz", tmp + "/se.fq")],
              "pe150": ["-i", plain(HERE + "/pe150_1.fq.gz", tmp + "/a.fq"), "-j", plain(HERE + "/pe150_2.fq.gz", tmp + "/b.fq")]}
    for cfg in ("mem_default", "mem_m5_noseg", "greedy_default", "greedy_e5_s40"):
        for tag, inp in inputs.items():
            out = subprocess.run([os.path.join(REF_DIR, "kaiju"), "-t", nodes, "-f", fmi, "-z", "1", "-v"] + inp + XP_CONFIGS[cfg], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            with gzip.open(os.path.join(HERE, "expected_v7_%s_%s.tsv.gz" % (cfg, tag)), "wb") as g:
                g.write(out)
    print("7-column outputs written")


if __name__ == "__main__" and "--v7" in sys.argv:
    main_v7()
