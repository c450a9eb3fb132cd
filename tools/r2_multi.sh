#!/bin/bash
# developer tool (round 2): 2-GPU call -- bench.py under torchrun exactly as the driver launches it (weak + strong scaling lines), the reference arm
# under torchrun (rank 0 only), kj_classify_multi / CLI -d all on two real devices, the N=1 line next to it on the same box
export KJ_NO_BUILD=1
o=gpurun_out; mkdir -p $o; tag=${1:-r2m}; N=${2:-2}
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $o/multi_box_$tag.txt; nvidia-smi topo -m >> $o/multi_box_$tag.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 > $o/bench_${tag}_n$N.json 2> $o/bench_${tag}_n$N.err
echo "N=$N exit $?"; tail -c 1500 $o/bench_${tag}_n$N.json; tail -4 $o/bench_${tag}_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 4 --warmup 3 --mode greedy --reads 5000000 > $o/bench_${tag}_greedy_n$N.json 2> $o/bench_${tag}_greedy_n$N.err
echo "N=$N greedy exit $?"; tail -c 700 $o/bench_${tag}_greedy_n$N.json
timeout 600 python bench.py --steps 4 --warmup 3 --skip-cpu --headline-only > $o/bench_${tag}_n1.json 2> $o/bench_${tag}_n1.err; tail -c 600 $o/bench_${tag}_n1.json
timeout 600 python -m pytest tests/test_gpu_build.py -q -k "multi" 2>&1 | tail -3 | tee $o/pytest_multi_$tag.log
