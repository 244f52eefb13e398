"""Attribute ncu samples / executed instructions of one kernel to source-line REGIONS of its .cu file, following
inlined code back to the line of the kernel it was inlined at (nvdisasm -gi).
usage: ncu_phases.py <src_page.csv> <nvdisasm -gi -c listing> <kernel substring> <file.cu> [skip=lo:hi] name:first_line ...
(regions in ascending order; a region ends where the next begins; an instruction belongs to the innermost frame of its
inline chain that lies in <file.cu> outside the lines lo..hi -- a helper lambda that is inlined into several regions)"""
import csv, re, sys, collections
src_csv, sass, kern, cu = sys.argv[1:5]
rest = sys.argv[5:]
skip = (0, -1)
if rest and rest[0].startswith("skip="):
    skip = tuple(int(v) for v in rest[0][5:].split(":"))
    rest = rest[1:]
regions = [(a.split(":")[0], int(a.split(":")[1])) for a in rest]
lines, inside = [], False
frames = []          # frames named by the //## File comments since the last instruction, innermost first
for ln in open(sass):
    if ln.startswith(".text.") or (".section" in ln and ".text." in ln):
        inside = kern in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        frames.append((m.group(1).split("/")[-1], int(m.group(2))))
        frames += [(a.split("/")[-1], int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))]
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        if inside:
            own = [l for f, l in frames if f == cu and not (skip[0] <= l <= skip[1])]
            lines.append(own[0] if own else (lines[-1] if lines else None))
        frames = []
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
agg, ins, tins = collections.Counter(), collections.Counter(), collections.Counter()
k = launches = 0
for r in rows:
    if not r or r[0] == "Kernel Name":
        continue
    if r[0] == "Address":
        k = 0
        launches += 1
        continue
    if len(r) != len(hdr):
        continue
    L = lines[k] if k < len(lines) else None
    k += 1
    name = "?"
    if L is not None:
        for nm, first in regions:
            if L >= first:
                name = nm
    agg[name] += float(r[si] or 0); ins[name] += float(r[ii] or 0); tins[name] += float(r[ti] or 0)
tot, itot = sum(agg.values()), sum(ins.values())
print(f"kernel sections in the page: {launches}; warp instructions {itot:.0f}; samples {tot:.0f}")
for nm in [n for n, _ in regions] + ["?"]:
    if ins[nm] or agg[nm]:
        print(f"{nm:12s} samples {agg[nm] / tot:6.1%}  warp instructions {ins[nm] / itot:6.1%}  active lanes {tins[nm] / max(ins[nm], 1):5.1f}")
