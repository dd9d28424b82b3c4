#!/bin/bash
# One `ncu --set full` capture of one kernel of a command that has already run clean, exported to CSV under gpurun_out/
# (the .ncu-rep itself stays in /tmp: with sources imported it exceeds what gpurun copies back).
#   usage: bash tools/profile_kernel.sh <tag> <kernel-name-regex> <launches-to-skip> <command...>
set -u
TAG=$1; RE=$2; SKIP=$3; shift 3
"$@" > /tmp/plain_$TAG.log 2>&1 || { echo "plain run failed: $TAG"; tail -5 /tmp/plain_$TAG.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c 1 -f -o /tmp/prof_$TAG "$@" > /tmp/ncu_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_$TAG.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > gpurun_out/ncu_${TAG}_source.csv 2>/dev/null
echo "$TAG: $(wc -c < gpurun_out/ncu_$TAG.csv) bytes raw, $(wc -c < gpurun_out/ncu_${TAG}_source.csv) bytes source"
