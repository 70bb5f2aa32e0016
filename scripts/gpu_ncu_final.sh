#!/bin/bash
# final profiles of the round: launch list of the default bench command + ncu --set full of every kernel family, summarised on the box
# (only text comes back).  usage: scripts/gpu_ncu_final.sh <tag>
tag=$1; out=gpurun_out; mkdir -p $out
cap() { # cap <name> <kernel regex> <mangled substring of the instantiation> <skip> <cmd...>
  name=$1; re=$2; sub=$3; skip=$4; shift 4
  timeout -k 5 240 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c 1 -f -o /tmp/${tag}_$name "$@" > $out/${tag}_ncu_$name.log 2>&1
  python scripts/ncu_summary.py /tmp/${tag}_$name.ncu-rep > $out/${tag}_${name}_ncu_raw_summary.txt 2>&1
  python scripts/ncu_lines.py /tmp/${tag}_$name.ncu-rep $sub 40 > $out/${tag}_${name}_source_lines.txt 2>&1
  grep -E "Kernel Name|gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|gpu__dram_throughput|smsp__issue_active|registers_per_thread" $out/${tag}_${name}_ncu_raw_summary.txt | head -8
}
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --iters 40 --therm 60 --no-cpu-baseline > $out/${tag}_ncu_b.log 2>&1
PIMC_CHAIN_MAJOR=1 cap k_chain k_chain k_chainILi0ELi4E 3 python bench.py --steps 1 --warmup 3 --iters 12 --therm 100 --no-cpu-baseline
cap k_sweep k_sweep k_sweepILi0ELi4ELb0E 200 python scripts/probe_ncu.py 4096 260
cap k_measure k_measure k_measureILi0ELi4E 60 python scripts/probe_ncu.py 4096 260
cap k_swap_iter k_swap_iter k_swap_iter 100 python bench.py --workload c2s --steps 1 --warmup 3 --iters 60 --therm 60 --no-cpu-baseline
cap k_run_cells k_run_cells k_run_cells 1 python bench.py --workload c4i --sched sweep --steps 1 --warmup 1 --iters 10 --therm 10 --no-cpu-baseline
cap k_paircorr k_paircorr k_paircorr 2 python scripts/probe_estim.py 1024
cap k_winding k_winding k_winding 2 python scripts/probe_estim.py 1024
cap k_init_world k_init_world k_init_world 0 python scripts/probe_estim.py 1024
ls $out | grep $tag | wc -l
