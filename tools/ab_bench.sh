#!/bin/bash
# developer tool: A/B several builds of libkaijub200 on the same box (usage: tools/ab_bench.sh <mode> <reads> <variant> ...; "" = default lib)
mode=$1; reads=$2; shift; shift
for v in "$@"; do
  lib=$PWD/kaiju_b200/libkaijub200$v.so
  KJ_B200_LIB=$lib python ${BENCH:-bench.py} --mode $mode --steps 3 --warmup 3 --skip-cpu --reads $reads 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant=[$v] mode=$mode reads=$reads value=%.2fM e2e=%.2fM kernel_ms=%.1f'%(d['value']/1e6, d['e2e']['value']/1e6, d['kernel_ms']), d['config']['launch'])"
done
