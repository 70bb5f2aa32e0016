#!/usr/bin/env python
"""one line per captured launch of an ncu_summary.py text file: kernel, time, registers, issue, fp64 pipe, DRAM bytes and %, top stalls"""
import re, sys
for f in sys.argv[1:]:
    print(f"== {f}")
    blocks = open(f).read().split("-" * 60)
    for b in blocks:
        kv = {}
        for line in b.splitlines():
            if line.startswith("Kernel Name"):
                kv["name"] = re.sub(r"\(.*", "", line[len("Kernel Name"):].strip()).replace("void ", "")
            else:
                parts = line.split()
                try:
                    kv[parts[0]] = float(parts[1])
                except (ValueError, IndexError):
                    pass
        if "name" not in kv:
            continue
        g = lambda k: kv.get(k, float("nan"))
        st = {k.split("stalled_")[1].split("_per_")[0]: v for k, v in kv.items() if "issue_stalled" in k}
        top = ", ".join(f"{k} {v:.2f}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
        print(f"{kv['name']:22s} {g('gpu__time_duration.sum'):9.1f} us  regs {g('launch__registers_per_thread'):3.0f}  issue {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} %  "
              f"fp64 {g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):5.1f} %  dram {g('dram__bytes_read.sum'):7.1f} + {g('dram__bytes_write.sum'):7.1f} MB "
              f"({g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):4.1f} %)  lanes/inst {g('smsp__thread_inst_executed_per_inst_executed.ratio'):4.1f}  inst {g('smsp__inst_executed.sum'):.3e}  stalls: {top}")
