#!/bin/bash
# Evidence run on the GPU box (one GPU): ncu launch list of a bench step + `ncu --set full` captures of the
# dominant kernels.  Usage (from the repo root):   bash tools/make_profiles.sh <tag>
# Outputs land in gpurun_out/ (scratch); tools/ncu_summary.py + tools/collect_profiles.py turn them into profiles/.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-sub --no-gpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"blend_bwd2?_kernel" -s 2 -c 1 -o $OUT/prof_blend_bwd_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-sub --no-gpu-baseline > $OUT/ncu_blend_bwd_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"blend_fwd2?_kernel" -s 6 -c 1 -o $OUT/prof_blend_fwd_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-sub --no-gpu-baseline > $OUT/ncu_blend_fwd_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"dec2_|dec_fold_kernel|dec_bwd_fold_kernel|dec_offsets" \
    -s 44 -c 13 -o $OUT/prof_decode_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu --no-sub --no-gpu-baseline > $OUT/ncu_decode_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"tile_sort_radix|tile_scatter|tile_hist|preprocess_(fwd|bwd)_kernel|visible_filter|l1_ssim" \
    -s 40 -c 12 -o $OUT/prof_rest_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu --no-sub --no-gpu-baseline > $OUT/ncu_rest_$TAG.log 2>&1
ls -la $OUT | grep $TAG
ncu --set full --clock-control none --import-source on -k regex:"adam_kernel|tv_add_grad_kernel|mvc_(fwd|bwd)_kernel" -s 2 -c 2 -o $OUT/prof_optim_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-sub --no-gpu-baseline > $OUT/ncu_optim_$TAG.log 2>&1
