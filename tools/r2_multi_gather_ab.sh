#!/bin/bash
# developer tool (round 2): short 2-GPU call -- MEM weak/strong line with the gather on high-priority streams; variant with 3 buffers in flight
export KJ_NO_BUILD=1
o=gpurun_out; mkdir -p $o; tag=${1:-r2n}; N=${2:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 > $o/bench_${tag}_n$N.json 2> $o/bench_${tag}_n$N.err
python -c "import json; d=json.loads(open('$o/bench_${tag}_n$N.json').read().strip().splitlines()[-1]); print('N=$N value %.2fM e2e %.2fM ms/step %.1f kernel_ms %.1f strong %.2fM' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['kernel_ms'], d['strong_scaling']['value']/1e6))"
KJ_BENCH_GATHER_MAIN=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 4 --warmup 3 > $o/bench_${tag}_main_n$N.json 2> $o/bench_${tag}_main_n$N.err
python -c "import json; d=json.loads(open('$o/bench_${tag}_main_n$N.json').read().strip().splitlines()[-1]); print('gather on the main stream: N=$N value %.2fM e2e %.2fM ms/step %.1f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']))"
timeout 300 python bench.py --steps 4 --warmup 3 --skip-cpu --headline-only 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 value %.2fM e2e %.2fM ms/step %.1f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']))"
