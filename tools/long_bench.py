#!/usr/bin/env python
"""Developer tool: kernel-only throughput of the classification path on read lengths other than the headline PE150
(PE250/PE350 pairs, 1-16 kb single reads, protein input), next to the unmodified reference CPU on a sample.
Usage: python tools/long_bench.py [--nprot 680000] [--mode mem|greedy] [--cpu]   (index is built/cached like bench.py's)"""
import argparse, json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nprot", type=int, default=680000); ap.add_argument("--mode", default="mem"); ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--workdir", default=os.environ.get("KJ_BENCH_DIR", "/tmp/kjbench")); ap.add_argument("--only", default="")
    a = ap.parse_args()
    import torch, kaiju_b200 as kb
    from helpers import SynthDB, build_fmi, REF_DIR
    os.makedirs(a.workdir, exist_ok=True)
    db = SynthDB(a.nprot, 1); fmi = os.path.join(a.workdir, "synth_%d.fmi" % a.nprot); nodes = os.path.join(a.workdir, "synth_%d_nodes.dmp" % a.nprot)
    if not os.path.exists(fmi):
        faa = os.path.join(a.workdir, "synth_%d.faa" % a.nprot); db.write(faa, nodes)
        os.replace(build_fmi(faa, os.path.join(a.workdir, "synth_%d" % a.nprot), threads=min(32, os.cpu_count())), fmi) if not os.path.exists(fmi) else None
    work = {"pe150": lambda: db.reads(7, 0, 2_000_000, 150, True), "pe250": lambda: db.reads(7, 0, 1_500_000, 250, True),
            "pe350": lambda: db.reads(7, 0, 1_000_000, 350, True),
            "se_0.4-1.2kb": lambda: db.long_reads(8, 0, 300_000, 400, 1200) + (None, None),
            "se_1-16kb": lambda: db.long_reads(9, 0, 60_000, 1000, 16383) + (None, None),
            "protein_30-1500aa": lambda: db.protein_reads(10, 0, 600_000, 30, 1500) + (None, None)}
    for name, gen in work.items():
        if a.only and a.only not in name: continue
        prot = name.startswith("protein")
        clf = kb.Classifier(fmi, nodes, device=0, params=kb.make_params(a.mode, protein=prot))
        s1, o1, s2, o2 = gen(); n = len(o1) - 1; units = int(o1[-1]) + (int(o2[-1]) if s2 is not None else 0)
        d = [torch.from_numpy(x.view(np.int64) if x.dtype == np.uint64 else x).cuda() if x is not None else None for x in (s1, o1, s2, o2)]
        p = [t.data_ptr() if t is not None else None for t in d]
        dt = torch.zeros(n, dtype=torch.int64, device="cuda"); ms = []
        for it in range(4):
            clf.classify_device(p[0], p[1], p[2], p[3], n, dt.data_ptr()); torch.cuda.synchronize(); clf.check_errors(); ms.append(clf.last_kernel_ms)
        best = min(ms[1:]); tax = dt.cpu().numpy()
        line = {"workload": name, "mode": a.mode, "reads": n, "classified": float((tax != 0).mean()), "kernel_ms": best, "reads_per_s": n / best * 1e3,
                "Mbases_per_s" if not prot else "Mresidues_per_s": units / best / 1e3, "launch": clf.launch_geometry}
        if a.cpu:
            ns = min(n, max(2000, int(3e8 / max(1, units / n) / (1 if a.mode == "mem" else 2))))
            d_ = tempfile.mkdtemp(prefix="kjlb_", dir=a.workdir)
            def fasta(path, s, o, k):
                with open(path, "w") as f:
                    for i in range(k): f.write(">r%d\n%s\n" % (i, s[int(o[i]):int(o[i + 1])].tobytes().decode()))
            fasta(d_ + "/a.fa", s1, o1, ns); tiny = d_ + "/t.fa"; fasta(tiny, s1, o1, 8)
            cmd = [os.path.join(REF_DIR, "kaiju"), "-t", nodes, "-f", fmi, "-a", a.mode, "-z", str(os.cpu_count()), "-o", "/dev/null"] + (["-p"] if prot else [])
            extra, extra_t = ["-i", d_ + "/a.fa"], ["-i", tiny]
            if s2 is not None: fasta(d_ + "/b.fa", s2, o2, ns); fasta(d_ + "/tb.fa", s2, o2, 8); extra += ["-j", d_ + "/b.fa"]; extra_t += ["-j", d_ + "/tb.fa"]
            t0 = time.time(); subprocess.check_call(cmd + extra_t, stderr=subprocess.DEVNULL); t1 = time.time(); subprocess.check_call(cmd + extra, stderr=subprocess.DEVNULL); t2 = time.time()
            cpu = ns / max(1e-3, (t2 - t1) - (t1 - t0)); line["cpu_reads_per_s"] = cpu; line["cpu_cores"] = os.cpu_count(); line["cpu_sample"] = ns; line["speedup_kernel_vs_cpu"] = line["reads_per_s"] / cpu
        print(json.dumps(line), flush=True); clf.close()


if __name__ == "__main__":
    main()
