#!/bin/bash
# r02 (second pass): ncu captures of the kernels that changed after r02_capture.sh
set -u
O=gpurun_out
run() {  # name, kernel regex, skip, command...
  name=$1; k=$2; skip=$3; shift 3
  "$@" > $O/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o $O/prof_r02b_$name "$@" > $O/ncu_$name.log 2>&1
  echo "$name rc=$?"
}
run sweep sweep_kernel 1 python profiles/run_sweep.py --reps 2
run td td_persist_kernel 1 python profiles/run_td.py --n 4 --games 4096 --warm 600 --steps 64
run td5 td_persist_kernel 1 python profiles/run_td.py --n 5 --games 65536 --warm 100 --steps 16
