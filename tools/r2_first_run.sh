#!/bin/bash
# developer tool (round 2): GPU call -- smoke, GPU parity tests, A/B matrix, ncu captures, the full bench line, file pipeline variants
export KJ_NO_BUILD=1
o=gpurun_out; mkdir -p $o; tag=${1:-r2e}
timeout 300 python __graft_entry__.py smoke > $o/smoke_$tag.log 2>&1 || echo "smoke FAILED" | tee -a $o/smoke_$tag.log
tail -2 $o/smoke_$tag.log
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $o/pytest_gpu_$tag.log
ab() { for m in mem greedy; do r=5000000; [ $m = greedy ] && r=3000000
  python bench.py --mode $m --steps 3 --warmup 3 --skip-cpu --headline-only --reads $r 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 mode=$m reads=$r value=%.2fM e2e=%.2fM kernel_ms=%.1f build_ms=%.0f'%(d['value']/1e6, d['e2e']['value']/1e6, d['kernel_ms'], d['config']['index_build_ms']), d['config']['launch'])"; done; }
(ab "default"; for v in pa8 pa10 gl16; do KJ_B200_LIB=$PWD/kaiju_b200/libkaijub200_v$v.so ab "$v"; done; ab "default-again") > $o/ab_$tag.txt 2>&1; cat $o/ab_$tag.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 3 -c 1 -f -o $o/prof_greedy_$tag python bench.py --mode greedy --reads 1000000 --steps 1 --warmup 3 --skip-cpu --headline-only > $o/ncu_greedy_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 3 -c 1 -f -o $o/prof_mem_$tag python bench.py --reads 2000000 --steps 1 --warmup 3 --skip-cpu --headline-only > $o/ncu_mem_$tag.log 2>&1
timeout 1500 python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err; tail -c 1200 $o/bench_$tag.json; tail -5 $o/bench_$tag.err
for cfg in "default:" "chunk32:KJ_INGEST_CHUNK=33554432" "io4:KJ_IO_THREADS=4" "chunk32io4:KJ_INGEST_CHUNK=33554432 KJ_IO_THREADS=4" "chunk128:KJ_INGEST_CHUNK=134217728"; do n=${cfg%%:*}; e=${cfg#*:}
  env $e python tools/file_bench.py --pairs 12000000 --ref-pairs 100000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('files $n: %.2f M pairs/s (%.3f s)' % (d['b200_pairs_per_s']/1e6, d['b200_files_s']))"; done | tee $o/file_variants_$tag.txt
python tools/long_bench.py --mode mem > $o/long_bench_${tag}_mem.jsonl 2>/dev/null; python tools/long_bench.py --mode greedy > $o/long_bench_${tag}_greedy.jsonl 2>/dev/null; tail -n 3 $o/long_bench_${tag}_mem.jsonl | cut -c1-300
ls -la $o | tail -6
