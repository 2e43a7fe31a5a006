"""Attribute the per-SASS-instruction samples of an ncu source-page CSV to the phases of the area kernel (line ranges of
maf_element.cuh at the end of round 1). usage: ncu_by_phase.py <sass csv> <nvdisasm -g -c listing of the kernel> <numel>"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]; col = {n: i for i, n in enumerate(H)}; ins = rows[hdr + 1:]
lines = []; cur = ("?", 0)
for l in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l): lines.append(cur)
numel = float(sys.argv[3])
R = [("gather", 240, 388), ("interp", 389, 429), ("gauss_item", 430, 662), ("residual", 663, 714), ("blk_short", 715, 772), ("blk_tr", 773, 820),
     ("blk_mesh", 821, 883), ("blk_fused", 884, 960), ("scatter", 961, 1017), ("task_setup", 1018, 1067), ("phase_tangent", 1068, 1120)]
def rng(f, ln):
    if f != "maf_element.cuh": return None
    for n, lo, hi in R:
        if lo <= ln <= hi: return n
    return None
# attribute helper lines (<234) to nearest following ranged line
attr = [None]*len(lines)
for i,(f,ln) in enumerate(lines):
    attr[i] = rng(f, ln)
last=None
for i in range(len(lines)-1,-1,-1):
    if attr[i] is None: attr[i] = last if lines[i][0]=="maf_element.cuh" else lines[i][0]
    else: last = attr[i]
S = collections.defaultdict(lambda:[0,0,0,0])
for a, r in zip(attr, ins):
    wf = int(r[col["L1 Wavefronts Shared"]] or 0); n=int(r[col["Instructions Executed"]]); s=int(r[col["# Samples"]])
    S[a][0]+=wf; S[a][1]+=n; S[a][2]+=s
    if "DFMA" in r[1] or "DMUL" in r[1] or "DADD" in r[1]: S[a][3]+=n
ts=sum(v[2] for v in S.values())
for k,(wf,n,s,fp) in sorted(S.items(), key=lambda kv:-kv[1][0]):
    print(f"{str(k):14s} wf/el {wf/numel:7.0f}  inst/el {n/numel:7.0f} fp64 inst/el {fp/numel:6.0f} samp {100*s/ts:5.1f}%")
