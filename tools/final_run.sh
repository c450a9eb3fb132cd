#!/bin/bash
# developer tool: the end-of-round GPU run (one gpurun call): parity tests, benches, launch list, ncu captures, side benchmarks.
# usage: tools/final_run.sh <tag>      outputs under gpurun_out/
tag=${1:-r1g}; o=gpurun_out; mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $o/pytest_gpu_$tag.log
python bench.py > $o/bench_${tag}_mem.json 2> $o/bench_${tag}_mem.err; tail -c 2500 $o/bench_${tag}_mem.json
python bench.py --mode greedy --steps 3 > $o/bench_${tag}_greedy.json 2> $o/bench_${tag}_greedy.err; tail -c 600 $o/bench_${tag}_greedy.json
python bench.py --impl reference --steps 2 --warmup 1 > $o/bench_${tag}_ref.json 2>/dev/null; tail -c 700 $o/bench_${tag}_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_$tag.csv python bench.py --steps 2 --warmup 3 --reads 2000000 --skip-cpu > $o/ncu_launch_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 3 -c 1 -f -o $o/prof_mem_$tag python bench.py --reads 2000000 --steps 1 --warmup 3 --skip-cpu > $o/ncu_mem_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 3 -c 1 -f -o $o/prof_greedy_$tag python bench.py --mode greedy --reads 1000000 --steps 1 --warmup 3 --skip-cpu > $o/ncu_greedy_$tag.log 2>&1
python tools/long_bench.py --mode mem > $o/long_bench_${tag}_mem.jsonl 2>/dev/null; python tools/long_bench.py --mode greedy > $o/long_bench_${tag}_greedy.jsonl 2>/dev/null
python tools/file_bench.py --pairs 12000000 > $o/file_bench_${tag}_mem.json 2>/dev/null; tail -c 500 $o/file_bench_${tag}_mem.json
ls -la $o | tail -15
