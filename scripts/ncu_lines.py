#!/usr/bin/env python
"""Join an ncu SASS source page with nvdisasm line info: instructions executed / stall samples per source line.
usage: ncu_lines.py report.ncu-rep mangled_kernel_substring [top]
The substring must select ONE instantiation (e.g. k_measureILi0ELi4E): sections of several instantiations share their address offsets."""
import csv, subprocess, sys, re, os, collections
rep, ksub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pimc_jl_b200", "libpimc_b200.so")
os.makedirs("/tmp/cub", exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd="/tmp/cub", capture_output=True)
import glob
dis = []
for cub in sorted(glob.glob("/tmp/cub/*.cubin")):    # one cubin per translation unit of the library
    dis += subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout.splitlines()
line_of, cur, infn = {}, None, False
for l in dis:
    if l.startswith("\t.section\t.text."):
        infn = ksub in l
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*);', l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[hi]
ai, ci, si, ti = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
srci = hdr.index("Source") if "Source" in hdr else None      # the report's own SASS text (reliable for opcodes)
base = int(rows[hi + 1][ai], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
ops = collections.defaultdict(int)
tot = [0, 0, 0]
for r in rows[hi + 1:]:
    if len(r) <= ci or not r[ci].isdigit():
        continue
    off = int(r[ai], 16) - base
    key, sass = line_of.get(off, (("?", 0), "?"))
    v = (int(r[ci]), int(r[si]), int(r[ti]))
    for k in range(3):
        agg[key][k] += v[k]
        tot[k] += v[k]
    if srci is not None and len(r) > srci and r[srci].strip():
        sass = r[srci].strip()
    toks = [t for t in sass.split() if not t.startswith("@")]
    ops[toks[0].split(".")[0] if toks and sass != "?" else "?"] += v[0]
print(f"total warp-inst {tot[0]}  thread-inst {tot[2]}  samples {tot[1]}")
src_cache = {}
def src(key):
    f, n = key
    for d in ("pimc_jl_b200/csrc", "include"):
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][n - 1].strip()[:100] if 0 < n <= len(src_cache[p]) else ""
    return ""
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/tot[0]:5.1f}% inst {100*v[1]/max(1,tot[1]):5.1f}% stall  lanes {v[2]/max(1,v[0]):4.1f}  {key[0]}:{key[1]:<4d} {src(key)}")
print("opcodes:", ", ".join(f"{k} {100*v/tot[0]:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:18]))
