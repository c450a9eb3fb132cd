#!/bin/bash
# developer tool (round 2): GPU call -- box facts, smoke, GPU parity tests, A/B (bulk-copy staging on/off), ncu captures, the full bench line
export KJ_NO_BUILD=1
o=gpurun_out; mkdir -p $o; tag=${1:-r2a}
(nproc; free -g; df -h /tmp /dev/shm . ; nvidia-smi --query-gpu=name,memory.total --format=csv; lscpu | head -20) > $o/boxinfo.txt 2>&1
if ! timeout 300 python __graft_entry__.py smoke > $o/smoke_$tag.log 2>&1; then
  echo "smoke FAILED with staging; retrying with KJ_NO_STAGE=1" | tee -a $o/smoke_$tag.log
  export KJ_NO_STAGE=1
  timeout 300 python __graft_entry__.py smoke >> $o/smoke_$tag.log 2>&1 || echo "smoke FAILED without staging too" | tee -a $o/smoke_$tag.log
fi
tail -3 $o/smoke_$tag.log
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $o/pytest_gpu_$tag.log
ab() { for m in mem greedy; do r=5000000; [ $m = greedy ] && r=3000000
  python bench.py --mode $m --steps 3 --warmup 3 --skip-cpu --headline-only --reads $r 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 mode=$m reads=$r value=%.2fM e2e=%.2fM kernel_ms=%.1f build_ms=%.0f'%(d['value']/1e6, d['e2e']['value']/1e6, d['kernel_ms'], d['config']['index_build_ms']), d['config']['launch'])"; done; }
(ab "stage=${KJ_NO_STAGE:+off}"; KJ_NO_STAGE=1 ab "stage=off"; KJ_HOST_BUILD=1 ab "hostbuild") > $o/ab_$tag.txt 2>&1; cat $o/ab_$tag.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 3 -c 1 -f -o $o/prof_greedy_$tag python bench.py --mode greedy --reads 1000000 --steps 1 --warmup 3 --skip-cpu --headline-only > $o/ncu_greedy_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 3 -c 1 -f -o $o/prof_mem_$tag python bench.py --reads 2000000 --steps 1 --warmup 3 --skip-cpu --headline-only > $o/ncu_mem_$tag.log 2>&1
timeout 1500 python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err; tail -c 3000 $o/bench_$tag.json; tail -5 $o/bench_$tag.err
timeout 600 python - <<'PY' > $o/mkbwt_timing.txt 2>&1
import sys, time, os
sys.path.insert(0, 'tests')
from helpers import SynthDB, build_fmi
for nprot in (2000000,):
    d = '/tmp/kjt_%d' % nprot; os.makedirs(d, exist_ok=True)
    t = time.time(); db = SynthDB(nprot, 1); db.write(d + '/db.faa', d + '/nodes.dmp'); t1 = time.time()
    build_fmi(d + '/db.faa', d + '/db', threads=64); t2 = time.time()
    print(nprot, 'gen %.1fs build %.1fs fmi bytes %d' % (t1 - t, t2 - t1, os.path.getsize(d + '/db.fmi')), flush=True)
PY
cat $o/boxinfo.txt $o/mkbwt_timing.txt; ls -la $o | tail -8
