#!/bin/bash
# developer tool (round 2): GPU call -- smoke, GPU parity tests, A/B matrix (staging, layout, occupancy, probes), ncu captures, the full bench line
export KJ_NO_BUILD=1
o=gpurun_out; mkdir -p $o; tag=${1:-r2b}
if ! timeout 300 python __graft_entry__.py smoke > $o/smoke_$tag.log 2>&1; then
  echo "smoke FAILED with staging; retrying with KJ_NO_STAGE=1" | tee -a $o/smoke_$tag.log
  export KJ_NO_STAGE=1
  timeout 300 python __graft_entry__.py smoke >> $o/smoke_$tag.log 2>&1 || echo "smoke FAILED without staging too" | tee -a $o/smoke_$tag.log
fi
tail -3 $o/smoke_$tag.log
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $o/pytest_gpu_$tag.log
ab() { for m in mem greedy; do r=5000000; [ $m = greedy ] && r=3000000
  python bench.py --mode $m --steps 3 --warmup 3 --skip-cpu --headline-only --reads $r 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 mode=$m reads=$r value=%.2fM e2e=%.2fM kernel_ms=%.1f build_ms=%.0f'%(d['value']/1e6, d['e2e']['value']/1e6, d['kernel_ms'], d['config']['index_build_ms']), d['config']['launch'])"; done; }
(ab "default"; KJ_NOMONO=1 ab "nomono(groups-of-8)"; KJ_B200_LIB=$PWD/kaiju_b200/libkaijub200_vprobe.so ab "probe"
 KJ_B200_LIB=$PWD/kaiju_b200/libkaijub200_vg4.so ab "greedy-4ctas"; KJ_B200_LIB=$PWD/kaiju_b200/libkaijub200_vstage.so ab "stage"; KJ_KMER_K=6 ab "kmer6"; KJ_KMER_K=4 ab "kmer4") > $o/ab_$tag.txt 2>&1; cat $o/ab_$tag.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 3 -c 1 -f -o $o/prof_greedy_$tag python bench.py --mode greedy --reads 1000000 --steps 1 --warmup 3 --skip-cpu --headline-only > $o/ncu_greedy_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 3 -c 1 -f -o $o/prof_mem_$tag python bench.py --reads 2000000 --steps 1 --warmup 3 --skip-cpu --headline-only > $o/ncu_mem_$tag.log 2>&1
timeout 1500 python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err; tail -c 3500 $o/bench_$tag.json; tail -5 $o/bench_$tag.err
python tools/file_bench.py --pairs 12000000 > $o/file_bench_${tag}_mem.json 2> $o/file_bench_$tag.err; tail -c 600 $o/file_bench_${tag}_mem.json; tail -3 $o/file_bench_$tag.err
ls -la $o | tail -8
