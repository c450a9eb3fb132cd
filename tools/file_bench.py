#!/usr/bin/env python
"""Developer tool: FASTQ files in -> kaiju output file out, wall clock, kaiju-b200 (device-side parsing/formatting, kj_classify_files)
next to the unmodified reference CLI on the same box.  Start-up (index load/transcode) is excluded on both sides by the two-size
differential T(N) - T(tiny).  Usage: python tools/file_bench.py [--pairs 4000000] [--ref-pairs 400000] [--mode mem] [--gz]"""
import argparse, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nprot", type=int, default=680000); ap.add_argument("--mode", default="mem"); ap.add_argument("--pairs", type=int, default=4_000_000)
    ap.add_argument("--ref-pairs", type=int, default=400_000); ap.add_argument("--workdir", default=os.environ.get("KJ_BENCH_DIR", "/tmp/kjbench")); ap.add_argument("--gz", action="store_true")
    a = ap.parse_args()
    from helpers import SynthDB, build_fmi, REF_DIR
    os.makedirs(a.workdir, exist_ok=True)
    db = SynthDB(a.nprot, 1); fmi = os.path.join(a.workdir, "synth_%d.fmi" % a.nprot); nodes = os.path.join(a.workdir, "synth_%d_nodes.dmp" % a.nprot)
    if not os.path.exists(fmi):
        faa = os.path.join(a.workdir, "synth_%d.faa" % a.nprot); db.write(faa, nodes); build_fmi(faa, os.path.join(a.workdir, "synth_%d" % a.nprot), threads=min(32, os.cpu_count()))
    def fq(tag, n):
        p1, p2 = os.path.join(a.workdir, "fb_%s_1.fq" % tag), os.path.join(a.workdir, "fb_%s_2.fq" % tag)
        if not os.path.exists(p1): db.write_fastq(7, 0, n, 150, True, p1, p2)
        if a.gz:
            for p in (p1, p2):
                if not os.path.exists(p + ".gz"): subprocess.check_call("gzip -1 -k -f " + p, shell=True)
            return p1 + ".gz", p2 + ".gz"
        return p1, p2
    big, tiny, refset = fq("big%d" % a.pairs, a.pairs), fq("tiny", 64), fq("ref%d" % a.ref_pairs, a.ref_pairs)
    for f in big: open(f, "rb").read()                                                     # page cache warm on both sides
    def run(cmd, files, out):
        t = time.time(); r = subprocess.run(cmd + ["-i", files[0], "-j", files[1], "-o", out], stderr=subprocess.PIPE, text=True, check=True, env=dict(os.environ, KJ_CLI_TIMING="1"))
        run.inner = [float(l.split(" classified, ")[1].split(" s")[0]) for l in r.stderr.splitlines() if " classified, " in l]
        for l in r.stderr.splitlines():
            if "KJ_FILES_TRACE" in l: print(l, file=sys.stderr)
        return time.time() - t
    ours = [os.path.join(ROOT, "kaiju_b200", "kaiju-b200"), "-t", nodes, "-f", fmi, "-a", a.mode]
    ref = [os.path.join(REF_DIR, "kaiju"), "-t", nodes, "-f", fmi, "-a", a.mode, "-z", str(os.cpu_count())]
    o = os.path.join(a.workdir, "fb_out.tsv")
    t_tiny = min(run(ours, tiny, o) for _ in range(2)); inner = []
    t_big = 1e9
    for _ in range(3): t_big = min(t_big, run(ours, big, o)); inner += run.inner
    nlines = sum(1 for _ in open(o))
    r_tiny = run(ref, tiny, o + ".ref"); r_big = run(ref, refset, o + ".ref")
    ours_rate = a.pairs / min(inner)           # timed inside the CLI around kj_classify_files: first byte read -> last byte written
    ref_rate = a.ref_pairs / max(1e-3, r_big - r_tiny)
    print(json.dumps({"workload": "PE150 FASTQ files%s -> output file, %s" % (" (gzip -1)" if a.gz else "", a.mode), "pairs": a.pairs, "lines_written": nlines,
                      "b200_wall_s": t_big, "b200_startup_s": t_tiny, "b200_files_s": min(inner), "b200_pairs_per_s": ours_rate, "input_GB": sum(os.path.getsize(f) for f in big) / 1e9,
                      "ref_pairs": a.ref_pairs, "ref_wall_s": r_big, "ref_startup_s": r_tiny, "ref_pairs_per_s": ref_rate, "ref_cores": os.cpu_count(), "speedup_files": ours_rate / ref_rate}))


if __name__ == "__main__":
    main()
