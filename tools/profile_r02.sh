#!/bin/bash
# Round-2 evidence for profiles/ (run under gpurun from the repo root: bash tools/profile_r02.sh):
#   launch list of the bench command (shares of the step), one `ncu --set full` capture of each one-SM FFT kernel
#   (forward real = BASELINE config 5a, complex, inverse real) exported to CSV, plus the push_ir and TMA step kernels.
set -u
CMD="python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline --no-single-process"
$CMD > gpurun_out/plain_r02.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/plain_r02.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > /tmp/ncu_l.log 2>&1
prof() {  # probe-kind  kernel-regex  skip  tag
  python tools/fft_sm_probe.py --modes 1 --iters 2 --skip-check --kinds $1 > /dev/null 2>&1 || { echo "plain run failed: $1"; return; }
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o /tmp/prof_r02_$4 python tools/fft_sm_probe.py --modes 1 --iters 2 --skip-check --kinds $1 > /tmp/ncu_$4.log 2>&1
  ncu -i /tmp/prof_r02_$4.ncu-rep --page raw --csv > gpurun_out/ncu_r02_$4.csv 2>/dev/null
  ncu -i /tmp/prof_r02_$4.ncu-rep --page source --csv > gpurun_out/ncu_r02_$4_source.csv 2>/dev/null
  echo "$4: $(wc -c < gpurun_out/ncu_r02_$4.csv) bytes"
}
prof r2c fft_sm_kernel 3 fft_sm_r2c
prof c2c fft_sm_kernel 3 fft_sm_c2c
prof c2r fft_sm_kernel 3 fft_sm_c2r
