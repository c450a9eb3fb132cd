#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Kaiju classification path.

Metric (BASELINE.json): read items/s (one item = one output line; a pair counts once) for 150 bp paired reads, MEM mode
-m 11 (default SEG on), against a viruses-scale .fmi, at N = 1/2/4/8 B200 -- next to the reference CPU `kaiju -z <all cores>`.

Workload (configs[1]): 10 M synthetic PE150 pairs per GPU (SURVEY.md 8d recipe, tools/kjgen.c, seeded) against the
"synth-viruses" stand-in (680 k proteins, ~2e8 letters; the real kaiju_db_viruses.fmi needs a download), whose .fmi is built
here with the reference's own kaiju-mkbwt/kaiju-mkfmi (-e 3).  A "step" is one pass of the hot path over the rank's batch.

  value      : whole-job items/s with reads resident in HBM (kj_classify_device), CUDA-event timed, max over ranks
  e2e        : same metric through kj_classify() with pinned HOST buffers (H2D of bases+offsets and D2H of results inside)
  roofline   : algorithmic bytes of the REFERENCE algorithm per launch (instrumented oracle on a sample of the same reads,
               SURVEY.md 8d) / kernel time, against the measured HBM peak in MEASURED_PEAKS.json
  cpu_baseline / --impl reference : the unmodified reference binary (oracle/_ref/kaiju -z <cores>) on a bounded sample.

Multi-GPU: one process per GPU (torchrun), index replicated, reads sharded (weak scaling), one NCCL all-gather of the
per-read taxon array at the end of every step (inside the timed region).
"""
import argparse, json, os, subprocess, sys, tempfile, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="mem", choices=["mem", "greedy"])
    ap.add_argument("--nprot", type=int, default=680000, help="synthetic DB size (680000 = viruses scale)")
    ap.add_argument("--reads", type=int, default=10_000_000, help="read pairs per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs for the CPU baseline sample (0 = auto)")
    ap.add_argument("--workdir", default=os.environ.get("KJ_BENCH_DIR", "/tmp/kjbench"))
    ap.add_argument("--skip-cpu", action="store_true")
    return ap.parse_args()


def build_workload(args, rank):
    """DB + index (rank 0 builds, others wait); returns (SynthDB, fmi, nodes)."""
    from helpers import SynthDB, build_fmi
    d = os.path.join(args.workdir, "db_%d" % args.nprot); os.makedirs(d, exist_ok=True)
    fmi, nodes, done = d + "/db.fmi", d + "/nodes.dmp", d + "/.done"
    db = SynthDB(args.nprot, 1)
    if rank == 0 and not os.path.exists(done):
        db.write(d + "/db.faa", nodes)
        build_fmi(d + "/db.faa", d + "/db", threads=min(64, os.cpu_count() or 8))
        os.remove(d + "/db.faa")
        open(done, "w").write("ok")
    while not os.path.exists(done):
        time.sleep(0.5)
    return db, fmi, nodes


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True); self.index = index; self.rows = []; self.stop_flag = False

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(float(r[0])) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(len(r) > col and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(float(self.rows[0][1])) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def ref_cpu_run(db, fmi, nodes, args, n_sample, seed, first, cores):
    """Time the unmodified reference on a bounded sample; returns (items/s, seconds)."""
    from helpers import REF_DIR
    d = tempfile.mkdtemp(prefix="kjcpu_", dir=args.workdir)
    fq1, fq2 = d + "/r1.fq", d + "/r2.fq"
    db.write_fastq(seed, first, n_sample, 150, True, fq1, fq2)
    def cmd(a, b):
        c = [os.path.join(REF_DIR, "kaiju"), "-t", nodes, "-f", fmi, "-i", a, "-j", b, "-a", args.mode, "-m", "11", "-z", str(cores), "-o", d + "/out.tsv"]
        return c + (["-e", "3", "-s", "65"] if args.mode == "greedy" else [])
    # index load is excluded with the two-size differential T(n) - T(tiny)  
    tiny1, tiny2 = d + "/t1.fq", d + "/t2.fq"
    db.write_fastq(seed, first, 16, 150, True, tiny1, tiny2)
    t = time.time(); subprocess.check_call(cmd(tiny1, tiny2), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL); t_load = time.time() - t
    t = time.time(); subprocess.check_call(cmd(fq1, fq2), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL); t_all = time.time() - t
    for f in (fq1, fq2, tiny1, tiny2, d + "/out.tsv"):
        try:
            os.remove(f)
        except OSError:
            pass
    dt = max(t_all - t_load, 1e-3)
    return n_sample / dt, dt


def emit(line):
    """The ONE JSON line of the contract goes to the real stdout; everything else printed by libraries (NCCL's version banner,
    build output) was diverted to stderr in main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1); os.dup2(2, 1)          # fd 1 -> stderr for the rest of the process; emit() writes to the saved descriptor
    args = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    cores = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        db, fmi, nodes = build_workload(args, 0)
        n_sample = args.cpu_sample or max(20000, min(3000000, 20000 * cores))
        vals = []
        for s in range(args.warmup + args.steps):
            v, dt = ref_cpu_run(db, fmi, nodes, args, n_sample, 1000 + s, 0, cores)
            if s >= args.warmup:
                vals.append((v, dt))
        v = sum(n_sample for _ in vals) / sum(d for _, d in vals)
        line = {"impl": "reference", "metric": "reads/sec (150 bp paired, %s mode, synth-viruses .fmi)" % args.mode.upper(), "value": v, "unit": "read pairs/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * sum(d for _, d in vals) / len(vals),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
                "config": {"workload": "configs[1]: %s -m 11, synthetic PE150 pairs vs synth-viruses .fmi (%d proteins); bounded sample of %d pairs per step" % (args.mode.upper(), args.nprot, n_sample)},
                "cpu_baseline": {"value": v, "unit": "read pairs/s", "cores": cores, "kind": "reference", "sample": "%d pairs per step, kaiju -z %d, index load excluded by differential" % (n_sample, cores)},
                "e2e": {"value": v, "unit": "read pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line); return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import kaiju_b200 as kb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    db, fmi, nodes = build_workload(args, rank)
    if dist:
        dist.barrier()
    params = kb.make_params(args.mode, m=11)
    clf = kb.Classifier(fmi, nodes, device=local, params=params)
    n = args.reads
    # this rank's shard of the job: items [rank*n, (rank+1)*n)
    s1, o1, s2, o2 = db.reads(7, rank * n, n, 150, True)
    # pinned host copies (e2e) and device-resident copies (value)
    def pin(a):
        t = torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a); return t.pin_memory()
    h = [pin(x) for x in (s1, o1, s2, o2)]
    d = [x.cuda(non_blocking=True) for x in h]
    d_tax = torch.zeros(n, dtype=torch.int64, device="cuda"); d_best = torch.zeros(n, dtype=torch.int32, device="cuda")
    h_tax = torch.zeros(n, dtype=torch.int64).pin_memory(); h_best = torch.zeros(n, dtype=torch.int32).pin_memory()
    gathered = torch.zeros(n * world, dtype=torch.int64, device="cuda") if dist else None
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()

    def step_device():
        clf.classify_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), n, d_tax.data_ptr(), d_best.data_ptr(), 150, 150, stream.cuda_stream)
        if dist:
            dist.all_gather_into_tensor(gathered, d_tax)

    def step_host():
        clf.classify_ptrs(h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), h[3].data_ptr(), n, h_tax.data_ptr(), h_best.data_ptr())
        if dist:
            d_tax.copy_(h_tax, non_blocking=True); dist.all_gather_into_tensor(gathered, d_tax)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        l0 = clf.kernel_launches; kms = []
        t0 = time.perf_counter(); ev0.record(stream)
        for _ in range(steps):
            fn()
            if fn is step_device:
                pass
        ev1.record(stream); torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if dist:
            dist.barrier()
        ms = ev0.elapsed_time(ev1)
        if fn is step_host:
            ms = wall * 1000.0           # host-buffer path runs on the library's own streams: wall clock bracketed by synchronize
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), clf.kernel_launches - l0

    sampler = ClockSampler(local); sampler.start()
    ms_dev, launches = timed(step_device, args.steps, max(3, args.warmup))
    kernel_ms = clf.last_kernel_ms                                   # CUDA events around the last classify kernel, on its launch stream
    sampler.stop_flag = True; sampler.join(timeout=2)
    ms_host, _ = timed(step_host, max(2, args.steps // 2), 1)
    clf.check_errors()
    assert torch.equal(h_tax, d_tax.cpu()), "host-buffer and device-buffer entry points disagree"
    # untimed: the kaiju2table-style summary -- per-taxon read counts in HBM, one all-reduce across the ranks
    clf.counts_reset(); clf.counts_add_device(d_tax.data_ptr(), n); torch.cuda.synchronize()
    if dist:
        from kaiju_b200.sharding import all_reduce_counts
        all_reduce_counts(None, dist, clf); torch.cuda.synchronize()
    ids_c, cnt_c = clf.counts()
    assert int(cnt_c.sum()) == n * world, "per-taxon counts do not add up to the number of reads"
    total = n * world
    value = total * args.steps / (ms_dev / 1000.0)
    e2e_steps = max(2, args.steps // 2)
    e2e = total * e2e_steps / (ms_host / 1000.0)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        line = {"metric": "reads/sec (150 bp paired, %s mode, synth-viruses .fmi)" % args.mode.upper(), "value": value, "unit": "read pairs/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int64", "data": "synthetic",
                "config": {"workload": "configs[1]: %s -m 11 (SEG on), %d synthetic PE150 pairs per GPU per step vs synth-viruses .fmi (%d proteins, bwtlen %d)" % (args.mode.upper(), n, args.nprot, clf.bwtlen),
                           "parallelism": "read-sharded x%d, index replicated, NCCL all-gather of taxon ids per step" % world if world > 1 else "single GPU",
                           "l2": "inputs (%.1f GB/step) and index (%.2f GB) exceed the 126 MB L2" % ((s1.nbytes + s2.nbytes + o1.nbytes + o2.nbytes) / 1e9, clf.index_bytes / 1e9),
                           "launch": dict(zip(("grid", "block", "dyn_smem"), clf.launch_geometry)),
                           "per_taxon_counts": "%d taxa with reads, counts sum to %d reads (untimed; all-reduce over %d rank(s))" % (len(ids_c) - 1, int(cnt_c.sum()), world)},
                "e2e": {"value": e2e, "unit": "read pairs/s", "h2d_bytes_per_step": int(s1.nbytes + s2.nbytes + o1.nbytes + o2.nbytes), "d2h_bytes_per_step": int(n * 12),
                        "note": "kj_classify() with pinned host buffers, chunked H2D/kernel/D2H pipeline inside"},
                "gpu_launches": int(launches), "clocks": sampler.summary(), "kernel_ms": kernel_ms}
        # ---- cpu baseline + roofline numerator on a bounded sample of the same workload (rank 0, N=1 only)
        if world == 1 and not args.skip_cpu:
            from helpers import Oracle, make_params, KoCounters
            n_or = 20000
            ctr = KoCounters(); orc = Oracle(fmi, nodes)
            t = time.time(); otax, _ = orc.classify_batch(make_params(args.mode), s1[:int(o1[n_or])], o1[:n_or + 1], s2[:int(o2[n_or])], o2[:n_or + 1], ctr); t_or = time.time() - t
            assert np.array_equal(otax, h_tax.numpy()[:n_or].view(np.uint64)), "GPU result differs from the oracle on the bench workload"
            c = ctr.as_dict()
            alg = (c["fmindex"] * 10 + c["scanned_bytes"] + c["lf_steps"] * 11 + c["get_suffix"] * 6 + c["bases"] + 8 * n_or) / n_or   # SURVEY.md 8d definition
            ach = alg * n / (kernel_ms / 1000.0) / 1e9
            traffic = None
            try:   # dram__bytes_read.sum + dram__bytes_write.sum per item from the latest `ncu --set full` capture (profiles/traffic.json)
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[args.mode]
                traffic = tj["dram_bytes_per_item"] * n
            except Exception:
                pass
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                                "algorithmic_bytes_per_item": alg, "peak_source": peak_src,
                                "note": "numerator = reference algorithm's bytes (instrumented oracle, %d-pair sample) x items per launch; kernel time from CUDA events on the launch stream" % n_or}
            n_cpu = args.cpu_sample or max(20000, min(3000000, 20000 * cores))
            v, dt = ref_cpu_run(db, fmi, nodes, args, n_cpu, 7, 0, cores)
            line["cpu_baseline"] = {"value": v, "unit": "read pairs/s", "cores": cores, "kind": "reference",
                                    "sample": "first %d pairs of the same workload, oracle/_ref/kaiju -z %d, %.1f s, index load excluded by differential" % (n_cpu, cores, dt),
                                    "oracle_port_value": n_or / t_or}
        emit(line)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
