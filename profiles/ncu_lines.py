#!/usr/bin/env python
"""Attribute the executed instructions / stall samples of one kernel in an .ncu-rep to CUDA source lines.

usage: python profiles/ncu_lines.py <report.ncu-rep> <kernel regex> <lib.so> [top]
Joins `ncu --page source --print-source sass` (per-instruction counts, absolute addresses) with
`nvdisasm --print-line-info` of the library (offset -> file:line; needs -lineinfo at compile time).
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, kre, lib = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name": cur = {"name": r[1], "hdr": None, "rows": []}; blocks.append(cur)
    elif r[0] == "Address": cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None: cur["rows"].append(r)
cands = [b for b in blocks if re.search(kre, b["name"])]
if os.environ.get("NCU_PICK") == "max":   # the matching launch that executed the most instructions
    def _n(b):
        k = b["hdr"].index("Instructions Executed")
        return sum(int(r[k] or 0) for r in b["rows"])
    blk = max(cands, key=_n)
else:
    blk = cands[int(os.environ.get("NCU_INDEX", "0"))]
h = blk["hdr"]; ia, ii, isamp, ithr = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
base = int(blk["rows"][0][ia], 16)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# find the function whose demangled-ish name matches: use the mangled fragment from the kernel name
name = blk["name"]
m = re.search(r"k_\w+", name); short = m.group(0)
targs = re.findall(r"\((?:int|bool)\)(\d+)", name.split("(const")[0].split("(cmx::")[0])
line_of, cur_line, infun, off_re = {}, None, False, re.compile(r"/\*([0-9a-f]{4,})\*/\s+(\S.*?);")
for ln in dis.splitlines():
    if ln.startswith("//--------------------- .text."):
        fn = ln.split(".text.")[1].split()[0]
        want = short in fn and all(("Li%sE" % t in fn) or ("Lb%sE" % t in fn) for t in targs)
        infun = want
        continue
    if not infun: continue
    mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if mm: cur_line = (os.path.basename(mm.group(1)), int(mm.group(2))); continue
    mo = off_re.search(ln)
    if mo: line_of[int(mo.group(1), 16)] = cur_line
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in blk["rows"]:
    off = int(r[ia], 16) - base
    key = line_of.get(off, ("?", 0))
    vals = [int(r[ii] or 0), int(r[isamp] or 0), int(r[ithr] or 0)]
    for k in range(3): agg[key][k] += vals[k]; tot[k] += vals[k]
src = {}
print(f"kernel: {name[:90]}\ninstructions executed: {tot[0]}  samples: {tot[1]}  avg active threads: {tot[2]/max(tot[0],1):.1f}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, l = key
    text = ""
    for d in (os.path.dirname(os.path.abspath(lib)) + "/csrc", "."):
        p = os.path.join(d, f)
        if os.path.exists(p):
            src.setdefault(p, open(p).read().splitlines()); text = src[p][l - 1].strip()[:100] if 0 < l <= len(src[p]) else ""; break
    print(f"{100*v[0]/tot[0]:5.1f}% inst {100*v[1]/max(tot[1],1):5.1f}% samp  thr/inst {v[2]/max(v[0],1):4.1f}  {f}:{l}  {text}")
