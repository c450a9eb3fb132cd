#!/bin/bash
# developer tool (round 2): GPU call -- tests + A/B of the front end running beside the search
export KJ_NO_BUILD=1
o=gpurun_out; mkdir -p $o; tag=${1:-r2l}
timeout 300 python __graft_entry__.py smoke > $o/smoke_$tag.log 2>&1 || echo "smoke FAILED" | tee -a $o/smoke_$tag.log
tail -1 $o/smoke_$tag.log
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $o/pytest_gpu_$tag.log
ab() { for m in ${AB_MODES:-mem greedy}; do r=5000000; [ $m = greedy ] && r=${AB_READS:-3000000}
  python bench.py --mode $m --steps 3 --warmup 3 --skip-cpu --headline-only --reads $r 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 mode=$m reads=$r value=%.2fM e2e=%.2fM kernel_ms=%.1f launches=%d'%(d['value']/1e6, d['e2e']['value']/1e6, d['kernel_ms'], d['gpu_launches']), d['config']['launch'])"; done; }
(AB_MODES=greedy; ab "default"; KJ_SPLIT_FULL_GRID=1 ab "search-on-all-slots"; KJ_B200_LIB=$PWD/kaiju_b200/libkaijub200_vfseg.so ab "front-seg"; KJ_SPLIT_SUB=750000 ab "sub750k"; AB_READS=10000000 ab "default-10M"; ab "default-again") > $o/ab_$tag.txt 2>&1; cat $o/ab_$tag.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $o/greedy_launches_$tag.csv python bench.py --mode greedy --reads 3000000 --steps 1 --warmup 3 --skip-cpu --headline-only > /dev/null 2>&1
for m in greedy mem; do KJ_FILES_TRACE=1 timeout 600 python tools/file_bench.py --pairs 12000000 --ref-pairs 100000 --mode $m > $o/file_bench_${m}_$tag.json 2> $o/file_trace_${m}_$tag.txt; cut -c1-330 $o/file_bench_${m}_$tag.json; grep KJ_FILES $o/file_trace_${m}_$tag.txt | tail -1; done
