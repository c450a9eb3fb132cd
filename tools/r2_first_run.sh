#!/bin/bash
# developer tool (round 2): GPU call -- smoke, GPU parity tests, A/B matrix, ncu captures, the full bench line, file pipeline trace
export KJ_NO_BUILD=1
o=gpurun_out; mkdir -p $o; tag=${1:-r2d}
timeout 300 python __graft_entry__.py smoke > $o/smoke_$tag.log 2>&1 || echo "smoke FAILED" | tee -a $o/smoke_$tag.log
tail -2 $o/smoke_$tag.log
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $o/pytest_gpu_$tag.log
ab() { for m in mem greedy; do r=5000000; [ $m = greedy ] && r=3000000
  python bench.py --mode $m --steps 3 --warmup 3 --skip-cpu --headline-only --reads $r 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 mode=$m reads=$r value=%.2fM e2e=%.2fM kernel_ms=%.1f build_ms=%.0f'%(d['value']/1e6, d['e2e']['value']/1e6, d['kernel_ms'], d['config']['index_build_ms']), d['config']['launch'])"; done; }
(ab "default"; KJ_NO_FIXED=1 ab "generic-layout"; KJ_KMER_K=5 ab "kmer5"; ab "default-again") > $o/ab_$tag.txt 2>&1; cat $o/ab_$tag.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 3 -c 1 -f -o $o/prof_greedy_$tag python bench.py --mode greedy --reads 1000000 --steps 1 --warmup 3 --skip-cpu --headline-only > $o/ncu_greedy_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 3 -c 1 -f -o $o/prof_mem_$tag python bench.py --reads 2000000 --steps 1 --warmup 3 --skip-cpu --headline-only > $o/ncu_mem_$tag.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_$tag.csv python bench.py --steps 2 --warmup 3 --reads 2000000 --skip-cpu --headline-only > $o/ncu_launch_$tag.log 2>&1
timeout 1500 python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err; tail -c 1500 $o/bench_$tag.json; tail -5 $o/bench_$tag.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $o/bench_${tag}_ref.json 2>/dev/null; tail -c 900 $o/bench_${tag}_ref.json
KJ_FILES_TRACE=1 python tools/file_bench.py --pairs 12000000 > $o/file_bench_${tag}_mem.json 2> $o/file_bench_$tag.err; tail -c 500 $o/file_bench_${tag}_mem.json; grep KJ_FILES_TRACE $o/file_bench_$tag.err | tail -4
ls -la $o | tail -6
