#!/bin/bash
# r02 ncu captures (one gpurun call): each kernel only after the same command has exited 0 without ncu
set -u
O=gpurun_out
run() {  # name, kernel regex, skip, command...
  name=$1; k=$2; skip=$3; shift 3
  "$@" > $O/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o $O/prof_r02_$name "$@" > $O/ncu_$name.log 2>&1
  echo "$name rc=$?"
}
run sweep sweep_kernel 1 python profiles/run_sweep.py --reps 2
run greedy_spec greedy_spec_kernel 1 python profiles/run_greedy.py --n 4 --games 1000 --pretrain 3000 --reps 2
run greedy6p greedy_play_kernel 1 python profiles/run_greedy.py --n 6 --games 131072 --pretrain 3000 --reps 2
run td td_persist_kernel 1 python profiles/run_td.py --n 4 --games 4096 --warm 600 --steps 64
run td5 td_persist_kernel 1 python profiles/run_td.py --n 5 --games 65536 --warm 100 --steps 16
python bench.py --steps 2 --warmup 1 --no-extras > $O/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > $O/ncu_bench.log 2>&1
echo "launches rc=$?"
