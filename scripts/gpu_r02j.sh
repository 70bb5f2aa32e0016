#!/bin/bash
tag=r02j; out=gpurun_out; mkdir -p $out
cap() { name=$1; re=$2; sub=$3; skip=$4; shift 4
  timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c 1 -f -o /tmp/${tag}_$name "$@" > $out/${tag}_ncu_$name.log 2>&1
  python scripts/ncu_summary.py /tmp/${tag}_$name.ncu-rep > $out/${tag}_${name}_ncu_raw_summary.txt 2>&1
  python scripts/ncu_lines.py /tmp/${tag}_$name.ncu-rep $sub 60 > $out/${tag}_${name}_source_lines.txt 2>&1
  head -3 $out/${tag}_${name}_source_lines.txt | cut -c1-150; tail -1 $out/${tag}_${name}_source_lines.txt | cut -c1-300; }
cap k_measure k_measure k_measureILi0ELi4E 60 python scripts/probe_ncu.py 4096 260
cap k_sweep k_sweep k_sweepILi0ELi4ELb0E 200 python scripts/probe_ncu.py 4096 260
