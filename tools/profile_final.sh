#!/bin/bash
# Final-state evidence for profiles/: launch list of the bench command, and a full capture of the headline kernel.
set -u
CMD="python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline"
$CMD > gpurun_out/plain_final.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_final.csv $CMD > /tmp/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:pconv_step -s 3 -c 1 -f -o /tmp/prof_pconv_final $CMD > /tmp/ncu_f.log 2>&1
ncu -i /tmp/prof_pconv_final.ncu-rep --page raw --csv > gpurun_out/ncu_r01_pconv_step_final.csv 2>/dev/null
# the TMA-fed variants
python tools/profile_all.py pconv_general > /dev/null 2>&1 && timeout 300 ncu --set full --clock-control none -k regex:pconv_mac_tma -s 1 -c 1 -f -o /tmp/prof_mactma python tools/profile_all.py pconv_general > /tmp/ncu_m.log 2>&1
ncu -i /tmp/prof_mactma.ncu-rep --page raw --csv > gpurun_out/ncu_r01_pconv_general_mac_tma.csv 2>/dev/null
ls -la gpurun_out/*final* gpurun_out/ncu_r01_pconv_general_mac_tma.csv
