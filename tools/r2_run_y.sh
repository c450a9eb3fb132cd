#!/bin/bash
# developer tool (round 2): short GPU call -- file-path tests and file bench after the FASTQ line-kernel change
export KJ_NO_BUILD=1
o=gpurun_out; mkdir -p $o; tag=${1:-r2y}
timeout 900 python -m pytest tests -m gpu -x -q -k "file or blank or cli or per_taxon or frontends or binding" 2>&1 | tail -4 | tee $o/pytest_files_$tag.log
for m in mem greedy; do KJ_FILES_TRACE=1 timeout 600 python tools/file_bench.py --pairs 12000000 --ref-pairs 100000 --mode $m > $o/file_bench_${m}_$tag.json 2> $o/file_trace_${m}_$tag.txt; cut -c1-330 $o/file_bench_${m}_$tag.json; grep KJ_FILES $o/file_trace_${m}_$tag.txt | tail -2; done

