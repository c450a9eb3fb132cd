#!/bin/bash
# developer tool (round 2): the final GPU call of the round -- what profiles/ cites for the committed build
export KJ_NO_BUILD=1
o=gpurun_out; mkdir -p $o; tag=${1:-r2z}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,driver_version --format=csv > $o/box_$tag.txt; nproc >> $o/box_$tag.txt
timeout 300 python __graft_entry__.py smoke > $o/smoke_$tag.log 2>&1 || echo "smoke FAILED" | tee -a $o/smoke_$tag.log
tail -1 $o/smoke_$tag.log
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $o/pytest_gpu_$tag.log
timeout 1500 python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err; tail -c 800 $o/bench_$tag.json; tail -3 $o/bench_$tag.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $o/bench_${tag}_reference.json 2> $o/bench_${tag}_reference.err; tail -c 400 $o/bench_${tag}_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_$tag.csv python bench.py --steps 2 --warmup 1 --skip-cpu --headline-only > $o/launches_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 3 -c 1 -f -o $o/prof_mem_$tag python bench.py --reads 2000000 --steps 1 --warmup 3 --skip-cpu --headline-only > $o/ncu_mem_$tag.log 2>&1
# Greedy, 1 M pairs = one sub-batch: launches alternate front end / search; 3 warm-up steps = 6 launches, then F (index 6) and S (index 7) of the timed step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 6 -c 2 -f -o $o/prof_greedy_$tag python bench.py --mode greedy --reads 1000000 --steps 1 --warmup 3 --skip-cpu --headline-only > $o/ncu_greedy_$tag.log 2>&1
for m in mem greedy; do KJ_FILES_TRACE=1 timeout 600 python tools/file_bench.py --pairs 12000000 --ref-pairs 400000 --mode $m > $o/file_bench_${m}_$tag.json 2> $o/file_trace_${m}_$tag.txt; cut -c1-330 $o/file_bench_${m}_$tag.json; grep KJ_FILES $o/file_trace_${m}_$tag.txt | tail -1; done
timeout 600 python tools/file_bench.py --pairs 12000000 --ref-pairs 100000 --gz > $o/file_bench_gz_$tag.json 2>/dev/null; cut -c1-330 $o/file_bench_gz_$tag.json
python tools/long_bench.py --mode mem > $o/long_bench_${tag}_mem.jsonl 2>/dev/null; python tools/long_bench.py --mode greedy > $o/long_bench_${tag}_greedy.jsonl 2>/dev/null; tail -n 2 $o/long_bench_${tag}_mem.jsonl | cut -c1-200
ls -la $o | tail -4
