#!/bin/bash
# Round 2, last third: launch list of the bench command and full captures of the kernels added after profiles/r02_ncu_new_kernels.md
# (run under gpurun from the repo root: bash tools/profile_r02c.sh)
set -u
CMD="python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline --no-single-process"
$CMD > gpurun_out/plain_r02c.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/plain_r02c.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_bench.csv $CMD > /tmp/ncu_l.log 2>&1
bash tools/profile_kernel.sh r02c_fft_sm_r2c_pdl fft_sm_kernel 3 python tools/fft_sm_probe.py --modes 1 --iters 2 --skip-check --kinds r2c
bash tools/profile_kernel.sh r02c_push_ir_reg pconv_push_ir_reg 1 python tools/profile_all.py push_ir
bash tools/profile_kernel.sh r02c_pconv_mono_deep pconv_step_kernel 2 python tools/profile_all.py pconv_mono_deep
for k in pconv_frames_reg pconv_mac_tma pconv_mac_sum pconv_inverse_ola; do
  bash tools/profile_kernel.sh r02c_general_$k $k 2 python tools/profile_all.py pconv_general_mono
done
