#!/bin/bash
# usage: scripts/ncu_capture.sh <tag> <kernel-regex> <count> <command...>
# Runs the command plainly (must exit 0), then under `ncu --set full`; the .ncu-rep stays in /tmp on the
# GPU box (too big for gpurun_out/), its raw / source pages are exported as CSV into gpurun_out/.
set -u
tag=$1; regex=$2; count=$3; shift 3
mkdir -p gpurun_out
"$@" > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:${regex} -c ${count} -f -o /tmp/${tag} "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i /tmp/${tag}.ncu-rep --page source --print-source cuda --csv > gpurun_out/${tag}_source_cuda.csv 2>/dev/null
ncu -i /tmp/${tag}.ncu-rep --page source --print-source sass --csv > gpurun_out/${tag}_source.csv 2>/dev/null
ncu -i /tmp/${tag}.ncu-rep --page details > gpurun_out/${tag}_details.txt 2>/dev/null
ls -la /tmp/${tag}.ncu-rep gpurun_out/${tag}_*
sz=$(stat -c %s /tmp/${tag}.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -lt 4000000 ]; then cp /tmp/${tag}.ncu-rep gpurun_out/; fi
