#!/bin/bash
# developer tool (round 2): first GPU call -- box facts, GPU parity tests, A/B of builds, ncu captures
export KJ_NO_BUILD=1
o=gpurun_out; mkdir -p $o; tag=${1:-r2a}
(nproc; free -g; df -h /tmp /dev/shm . ; nvidia-smi --query-gpu=name,memory.total --format=csv; lscpu | head -20) > $o/boxinfo.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $o/pytest_gpu_$tag.log
BENCH=tools/_bench_frozen.py bash tools/ab_bench.sh mem 5000000 "" > $o/ab_$tag.txt 2>&1
BENCH=tools/_bench_frozen.py bash tools/ab_bench.sh greedy 3000000 "" >> $o/ab_$tag.txt 2>&1
cat $o/ab_$tag.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 3 -c 1 -f -o $o/prof_greedy_$tag python tools/_bench_frozen.py --mode greedy --reads 1000000 --steps 1 --warmup 3 --skip-cpu > $o/ncu_greedy_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kj_classify -s 3 -c 1 -f -o $o/prof_mem_$tag python tools/_bench_frozen.py --reads 2000000 --steps 1 --warmup 3 --skip-cpu > $o/ncu_mem_$tag.log 2>&1
timeout 900 python - <<'PY' > $o/mkbwt_timing.txt 2>&1
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
