#!/bin/bash
# The three ncu captures behind profiles/round2_* (ONE of them per gpurun call: scripts/dev_profile_all.sh list|full|hbm)
set -e
mkdir -p gpurun_out
case "$1" in
  list)  # every launch of one eager step with its device time
    CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
    $CMD > gpurun_out/plain_list.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1 ;;
  full)  # the tensor-core conv kernels of one step, full metric set
    CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
    $CMD > gpurun_out/plain_full.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:"igemm_conv3x2|igemm_conv3r|igemm_wgrad9x2" -s 144 -c 48 \
        -o /tmp/r2_conv_full $CMD > gpurun_out/ncu_full.log 2>&1
    # the report itself (70 MB) stays on the box: gpurun copies back at most 64 MiB; the raw page carries every metric
    ncu -i /tmp/r2_conv_full.ncu-rep --page raw --csv > gpurun_out/r2_conv_full_raw.csv ;;
  hbm)   # the HBM-bound statistics / loss / BatchNorm kernels at the benchmark sizes
    CMD="python scripts/dev_ncu_targets.py confusion argmax_confusion head_argmax ce_kd_loss maxpool_bwd im2col_stem head_loss bn_bwd bn_reduce bn_apply bn_apply_pool adam"
    $CMD > gpurun_out/plain_hbm.log 2>&1 &&
    ncu --set full --clock-control none -k regex:"confusion_kernel|head_argmax_kernel|ce_kd_loss_kernel|maxpool_bwd_add_reduce_kernel|im2col3x3_stem3_kernel|head_loss_kernel|bn_relu_bwd_apply_kernel|bn_bwd_reduce_kernel|bn_apply_kernel|bn_apply_pool_kernel|adam_kernel" \
        -o /tmp/r2_hbm_full $CMD > gpurun_out/ncu_hbm.log 2>&1
    ncu -i /tmp/r2_hbm_full.ncu-rep --page raw --csv > gpurun_out/r2_hbm_full_raw.csv ;;
esac
