#!/bin/bash
# developer tool: shorter end-of-round refresh (tests + the two benches + file benchmark); outputs under gpurun_out/
tag=${1:-r1h}; o=gpurun_out; mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $o/pytest_gpu_$tag.log
python bench.py > $o/bench_${tag}_mem.json 2> $o/bench_${tag}_mem.err; python -c "import json;d=json.load(open('$o/bench_${tag}_mem.json'));print('MEM value %.2fM e2e %.2fM frac %.3f cpu %.3fM'%(d['value']/1e6,d['e2e']['value']/1e6,d['roofline']['frac'],d['cpu_baseline']['value']/1e6), d['config'].get('per_taxon_counts'))"
python bench.py --mode greedy --steps 3 > $o/bench_${tag}_greedy.json 2> $o/bench_${tag}_greedy.err; python -c "import json;d=json.load(open('$o/bench_${tag}_greedy.json'));print('GREEDY value %.2fM e2e %.2fM frac %.3f cpu %.3fM'%(d['value']/1e6,d['e2e']['value']/1e6,d['roofline']['frac'],d['cpu_baseline']['value']/1e6))"
python tools/file_bench.py --pairs 12000000 > $o/file_bench_${tag}_mem.json 2>/dev/null; tail -c 400 $o/file_bench_${tag}_mem.json
