#!/bin/bash
# developer tool (round 2): GPU call -- smoke, GPU tests, A/B of the two-kernel Greedy path, ncu capture of the search kernel, file path
export KJ_NO_BUILD=1
o=gpurun_out; mkdir -p $o; tag=${1:-r2i}
timeout 300 python __graft_entry__.py smoke > $o/smoke_$tag.log 2>&1 || echo "smoke FAILED" | tee -a $o/smoke_$tag.log
tail -2 $o/smoke_$tag.log
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $o/pytest_gpu_$tag.log
ab() { for m in ${AB_MODES:-mem greedy}; do r=5000000; [ $m = greedy ] && r=3000000
  python bench.py --mode $m --steps 3 --warmup 3 --skip-cpu --headline-only --reads $r 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 mode=$m reads=$r value=%.2fM e2e=%.2fM kernel_ms=%.1f launches=%d'%(d['value']/1e6, d['e2e']['value']/1e6, d['kernel_ms'], d['gpu_launches']), d['config']['launch'])"; done; }
(ab "default"; AB_MODES=greedy; KJ_NO_SPLIT=1 ab "one-kernel"; KJ_SPLIT_SUB=750000 ab "sub750k"; ab "default-again") > $o/ab_$tag.txt 2>&1; cat $o/ab_$tag.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 9 -c 1 -f -o $o/prof_greedy_$tag python bench.py --mode greedy --reads 1000000 --steps 1 --warmup 3 --skip-cpu --headline-only > $o/ncu_greedy_$tag.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $o/greedy_launches_$tag.csv python bench.py --mode greedy --reads 3000000 --steps 1 --warmup 3 --skip-cpu --headline-only > /dev/null 2>&1; grep -c kj_classify $o/greedy_launches_$tag.csv
timeout 1500 python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err; tail -c 1000 $o/bench_$tag.json; tail -3 $o/bench_$tag.err
for m in mem greedy; do KJ_FILES_TRACE=1 timeout 600 python tools/file_bench.py --pairs 12000000 --ref-pairs 100000 --mode $m > $o/file_bench_${m}_$tag.json 2> $o/file_trace_${m}_$tag.txt; cut -c1-330 $o/file_bench_${m}_$tag.json; grep KJ_FILES $o/file_trace_${m}_$tag.txt | tail -2; done
ls -la $o | tail -4
