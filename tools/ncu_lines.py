"""Attribute ncu warp-stall samples (SASS page csv) to CUDA source lines using nvdisasm -g output.
usage: ncu_lines.py <src_page.csv> <nvdisasm -g -c listing> [kernel substring]"""
import csv, re, sys, collections
src_csv, sass, kern = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "step_kernelILb0")
# line table: ordered list of (line) per instruction of the kernel
lines, cur, inside = [], None, False
for ln in open(sass):
    if ln.startswith(".text.") or ".section" in ln and ".text." in ln:
        inside = kern in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    if inside and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
per_kernel = len(lines)
agg, instr = collections.Counter(), collections.Counter()
k = 0
n_rows = 0
for r in rows:
    if not r or r[0] == "Kernel Name":
        continue
    if r[0] == "Address":
        k = 0
        continue
    if len(r) != len(hdr):
        continue
    key = lines[k] if k < per_kernel else None
    agg[key] += float(r[si] or 0)
    instr[key] += float(r[ii] or 0)
    k += 1
    n_rows += 1
data = range(n_rows)
tot = sum(agg.values())
print("instructions in kernel:", per_kernel, "rows:", len(data), "samples:", tot)
srcs = {}
for key, v in agg.most_common(40):
    if key is None:
        continue
    f, l = key
    if f not in srcs:
        import glob
        cand = glob.glob("/root/repo/bridges-with-reinforcement-learning_b200/csrc/" + f)
        srcs[f] = open(cand[0]).read().splitlines() if cand else []
    text = srcs[f][l - 1].strip() if srcs[f] and l <= len(srcs[f]) else ""
    print(f"{v / tot:6.1%} instr={instr[key]:>11.0f} {f}:{l:<4d} {text[:100]}")
