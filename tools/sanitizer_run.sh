#!/bin/bash
# compute-sanitizer passes over the small end-to-end target.  Usage on the GPU box, from the repo root:
#   bash tools/sanitizer_run.sh <tag>        -> gpurun_out/sanitizer_{memcheck,racecheck}_<tag>.txt
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 compute-sanitizer --tool memcheck --print-limit 30 --log-file $OUT/sanitizer_memcheck_$TAG.txt python tools/sanitize_target.py 2 > $OUT/sanitizer_memcheck_$TAG.out 2>&1
echo "memcheck rc=$?" >> $OUT/sanitizer_memcheck_$TAG.txt
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 --log-file $OUT/sanitizer_racecheck_$TAG.txt python tools/sanitize_target.py 1 > $OUT/sanitizer_racecheck_$TAG.out 2>&1
echo "racecheck rc=$?" >> $OUT/sanitizer_racecheck_$TAG.txt
tail -5 $OUT/sanitizer_memcheck_$TAG.txt $OUT/sanitizer_racecheck_$TAG.txt
