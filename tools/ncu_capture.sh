#!/bin/bash
# ncu --set full capture of one kernel, summarised ON the GPU box (reports are tens of MB; gpurun brings back 64 MiB):
#   tools/ncu_capture.sh <tag> <kernel-regex> <skip> <prof_run.py args...>
# writes gpurun_out/<tag>_raw.csv (ncu --page raw), gpurun_out/<tag>_summary.txt, gpurun_out/<tag>_by_function.txt
tag=$1; kernel=$2; skip=$3; shift 3
rep=/tmp/$tag.ncu-rep
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$kernel" -s "$skip" -c 1 -o /tmp/$tag -f \
    python tools/prof_run.py "$@" > gpurun_out/${tag}_ncu.log 2>&1
if [ -f "$rep" ]; then
  ncu -i "$rep" --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/${tag}_raw.csv > gpurun_out/${tag}_summary.txt 2>&1
  python tools/ncu_by_function.py "$rep" "${BYFN_NAME:-$kernel}" > gpurun_out/${tag}_by_function.txt 2>&1
  ls -la "$rep" | awk '{print $5}' >> gpurun_out/${tag}_summary.txt
else
  echo "no report" > gpurun_out/${tag}_summary.txt; tail -n 5 gpurun_out/${tag}_ncu.log >> gpurun_out/${tag}_summary.txt
fi
