"""Aggregate ncu warp-stall samples of step_kernel by code region (function / phase).
usage: ncu_regions.py <src_page.csv> <nvdisasm -g -c listing>"""
import csv, re, sys, collections
src_csv, sass = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "step_kernelILb0"
lines, cur, inside = [], None, False
for ln in open(sass):
    if ln.startswith(".text.") or ".section" in ln and ".text." in ln:
        inside = kern in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    if inside and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
# region table from the sources: function starts
import os
ROOT = os.environ.get("BW_SRC_ROOT", "/root/repo/bridges-with-reinforcement-learning_b200/csrc/")
def regions_of(fname):
    out = []
    for i, t in enumerate(open(ROOT + fname).read().splitlines(), 1):
        m = re.match(r"\s*(?:template.*)?__device__.*?\b(\w+)\(", t)
        if m and "{" in t or (m and not t.strip().endswith(";")):
            out.append((i, m.group(1)))
        if re.match(r"\s*step_kernel\(", t):                      # the kernel's own prologue: state loads, CTA order
            out.append((i, "kernel prologue (state loads)"))
        m2 = re.match(r"\s*// -+ (phase \d[^\n]*|posed bodies[^\n]*|collision flags[^\n]*|raster update[^\n]*)", t)
        if m2:
            out.append((i, m2.group(1)[:40]))
    return out
regs = {f: regions_of(f) for f in ("bw_solver.cuh", "bw_lp.cuh", "bw_step.cu", "bw_common.cuh")}
def region(key):
    if key is None: return "?"
    f, l = key
    if f not in regs: return f
    name = f + ":top"
    for start, n in regs[f]:
        if l >= start: name = f.split(".")[0][3:] + ":" + n
    return name
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg, instr = collections.Counter(), collections.Counter()
st = collections.defaultdict(collections.Counter)
k = 0
for r in rows:
    if not r or r[0] == "Kernel Name": continue
    if r[0] == "Address": k = 0; continue
    if len(r) != len(hdr): continue
    key = region(lines[k] if k < len(lines) else None)
    agg[key] += float(r[si] or 0); instr[key] += float(r[ii] or 0)
    for c in stall_cols: st[key][hdr[c]] += float(r[c] or 0)
    k += 1
tot = sum(agg.values())
print("samples", tot, "instructions", sum(instr.values()))
for key, v in agg.most_common(30):
    top = ", ".join(f"{n[6:]} {c / max(v,1):.0%}" for n, c in st[key].most_common(3))
    print(f"{v / tot:6.1%}  instr {instr[key] / sum(instr.values()):6.1%}  {key:38s} {top}")
