#!/bin/bash
# developer tool: smoke + the whole GPU suite on the committed tree (what the driver runs at round end)
export KJ_NO_BUILD=1
o=gpurun_out; mkdir -p $o; tag=${1:-r2w}
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1 | tee $o/smoke_$tag.log
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $o/pytest_gpu_$tag.log
