#!/bin/bash
# One `ncu --set full` capture per hot kernel (second launch of each), exported to small CSVs under gpurun_out/.
# Run under gpurun from the repo root: bash tools/profile_all.sh
set -u
prof() {  # workload  kernel-regex  skip
  local wl=$1 rx=$2 skip=$3 tag=$4
  python tools/profile_all.py $wl > /dev/null 2>&1 || { echo "plain run failed: $wl"; return; }
  timeout 300 ncu --set full --clock-control none -k regex:$rx -s $skip -c 1 -f -o /tmp/prof_$tag python tools/profile_all.py $wl > /tmp/ncu_$tag.log 2>&1
  ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_r01_$tag.csv 2>/dev/null
  echo "$tag: $(wc -c < gpurun_out/ncu_r01_$tag.csv) bytes"
}
prof cfft1024 cfft_kernel 1 cfft1024
prof rfft4096 rfft_fwd_reg 1 rfft4096_fwd
prof rfft4096 rfft_inv_reg 1 rfft4096_inv
prof rfft65536 large_cols 1 rfft65536_cols
prof rfft65536 large_rows 1 rfft65536_rows
# (pconv kernels unchanged since their captures)
prof dconv4 dconv_fir 1 dconv4
