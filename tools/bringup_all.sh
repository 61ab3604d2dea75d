#!/bin/bash
# Runs every GEMM bring-up case in its own process under `timeout`; log -> gpurun_out/bringup.log
mkdir -p gpurun_out
LOG=gpurun_out/bringup.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader >> $LOG 2>&1
CASES="fprop2d fprop2d_small fprop2d_tiny fprop_conv fprop_conv7 fprop_conv14 fprop_conv_odd dgrad2d dgrad2d_tiny dgrad_conv dgrad_conv7 wgrad2d wgrad_conv wgrad_conv7 wgrad_conv14 fprop_conv_big"
for c in $CASES; do
  timeout 90 python tools/bringup_gemm.py $c default >> $LOG 2>&1
  rc=$?
  echo "== $c default rc=$rc" >> $LOG
  if [ $rc -ne 0 ]; then
    case $c in
      dgrad*|wgrad*)
        for v in swap sbo1024; do
          timeout 90 python tools/bringup_gemm.py $c $v >> $LOG 2>&1
          echo "== $c $v rc=$?" >> $LOG
        done;;
    esac
  fi
done
tail -60 $LOG
