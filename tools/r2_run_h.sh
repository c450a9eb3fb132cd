#!/bin/bash
# developer tool (round 2): GPU call -- smoke, GPU tests, A/B, ncu capture (Greedy), launch list of the file path, the full bench line
export KJ_NO_BUILD=1
o=gpurun_out; mkdir -p $o; tag=${1:-r2h}
timeout 300 python __graft_entry__.py smoke > $o/smoke_$tag.log 2>&1 || echo "smoke FAILED" | tee -a $o/smoke_$tag.log
tail -2 $o/smoke_$tag.log
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $o/pytest_gpu_$tag.log
ab() { for m in mem greedy; do r=5000000; [ $m = greedy ] && r=3000000
  python bench.py --mode $m --steps 3 --warmup 3 --skip-cpu --headline-only --reads $r 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 mode=$m reads=$r value=%.2fM e2e=%.2fM kernel_ms=%.1f build_ms=%.0f'%(d['value']/1e6, d['e2e']['value']/1e6, d['kernel_ms'], d['config']['index_build_ms']), d['config']['launch'])"; done; }
(ab "default"; for v in gc su; do KJ_B200_LIB=$PWD/kaiju_b200/libkaijub200_v$v.so ab "$v"; done; ab "default-again") > $o/ab_$tag.txt 2>&1; cat $o/ab_$tag.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 3 -c 1 -f -o $o/prof_greedy_$tag python bench.py --mode greedy --reads 1000000 --steps 1 --warmup 3 --skip-cpu --headline-only > $o/ncu_greedy_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 3 -c 1 -f -o $o/prof_mem_$tag python bench.py --reads 2000000 --steps 1 --warmup 3 --skip-cpu --headline-only > $o/ncu_mem_$tag.log 2>&1
timeout 1500 python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err; tail -c 1200 $o/bench_$tag.json; tail -5 $o/bench_$tag.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $o/bench_${tag}_reference.json 2> $o/bench_${tag}_reference.err; tail -c 600 $o/bench_${tag}_reference.json
KJ_FILES_TRACE=1 timeout 600 python tools/file_bench.py --pairs 12000000 --ref-pairs 100000 > $o/file_bench_$tag.json 2> $o/file_trace_$tag.txt; cut -c1-400 $o/file_bench_$tag.json; grep KJ_FILES $o/file_trace_$tag.txt | tail -3
d=/tmp/kjbench
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $o/file_launches_$tag.csv kaiju_b200/kaiju-b200 -t $d/synth_680000_nodes.dmp -f $d/synth_680000.fmi -a mem -i $d/fb_big12000000_1.fq -j $d/fb_big12000000_2.fq -o $d/ncu_out.tsv > $o/file_ncu_$tag.log 2>&1
python - <<'PY' | tee $o/file_launches_${tag}_summary.txt
import csv, collections, sys
rows = [r for r in csv.reader(open("gpurun_out/file_launches_%s.csv" % (sys.argv[1] if len(sys.argv) > 1 else "r2h"), errors="replace")) if len(r) > 10]
if rows:
    h = rows[0]; kn = h.index("Kernel Name"); mv = h.index("Metric Value"); agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try: v = float(r[mv].replace(",", ""))
        except Exception: continue
        a = agg[r[kn].split("(")[0][:60]]; a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]: print("%-62s %6d launches %10.3f ms %5.1f %%" % (k, a[0], a[1] / 1e6, 100 * a[1] / tot))
PY
ls -la $o | tail -5
