#!/bin/bash
# developer tool (round 2): GPU call -- GPU tests, file pipeline variants, small A/B
export KJ_NO_BUILD=1
o=gpurun_out; mkdir -p $o; tag=${1:-r2g}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $o/pytest_gpu_$tag.log
M=1048576
for cfg in "default:" "batch64:KJ_INGEST_BATCH=$((64*M))" "batch256:KJ_INGEST_BATCH=$((256*M))" "chunk8:KJ_INGEST_CHUNK=$((8*M)) KJ_INGEST_BATCH=$((128*M))" "chunk32:KJ_INGEST_CHUNK=$((32*M)) KJ_INGEST_BATCH=$((128*M))" "io16:KJ_IO_THREADS=16"; do n=${cfg%%:*}; e=${cfg#*:}
  env $e KJ_FILES_TRACE=1 timeout 600 python tools/file_bench.py --pairs 12000000 --ref-pairs 100000 2> $o/ft.tmp | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('files $n: %.2f M pairs/s (%.3f s; wall %.2f s, start-up %.2f s)' % (d['b200_pairs_per_s']/1e6, d['b200_files_s'], d['b200_wall_s'], d['b200_startup_s']))"
  grep KJ_FILES_TRACE $o/ft.tmp | tail -2; done 2>&1 | tee $o/file_variants_$tag.txt
timeout 600 python tools/file_bench.py --pairs 12000000 --ref-pairs 100000 --mode greedy 2>/dev/null | cut -c1-400 | tee $o/file_bench_greedy_$tag.json
ab() { for m in mem greedy; do r=5000000; [ $m = greedy ] && r=3000000
  python bench.py --mode $m --steps 3 --warmup 3 --skip-cpu --headline-only --reads $r 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 mode=$m reads=$r value=%.2fM e2e=%.2fM kernel_ms=%.1f build_ms=%.0f'%(d['value']/1e6, d['e2e']['value']/1e6, d['kernel_ms'], d['config']['index_build_ms']), d['config']['launch'])"; done; }
(ab "default"; KJ_NO_KMER7=1 ab "no-kmer7"; KJ_B200_LIB=$PWD/kaiju_b200/libkaijub200_vsr.so ab "greedy-split-rolled"; ab "default-again") > $o/ab_$tag.txt 2>&1; cat $o/ab_$tag.txt
