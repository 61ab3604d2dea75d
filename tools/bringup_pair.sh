#!/bin/bash
# GPU bring-up of the cta_group::2 GEMM form: every bring-up case in its own process under a
# timeout (a protocol bug hangs the kernel), forced pair mode (LOFT_2CTA=2) vs 1-CTA (LOFT_2CTA=0).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/bringup_pair.txt
: > $out
for mode in 2 0; do
  for c in fprop2d fprop2d_wide fprop2d_big fprop2d_tiny fprop_conv7 fprop_conv14 fprop_conv_big fprop_conv_odd \
           dgrad2d dgrad2d_tiny dgrad_conv7 dgrad_conv wgrad2d wgrad2d_wide wgrad_conv7 wgrad_conv14 wgrad_conv; do
    echo "== LOFT_2CTA=$mode $c" >> $out
    LOFT_2CTA=$mode timeout -s KILL 90 python tools/bringup_gemm.py $c >> $out 2>&1
    echo "rc=$?" >> $out
  done
done
cat $out
