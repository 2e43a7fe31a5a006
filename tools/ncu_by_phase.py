"""Attribute the per-SASS-instruction samples of an ncu source-page CSV to the phases of the area kernel (line ranges of
maf_element.cuh, located by their function names). usage: ncu_by_phase.py <sass csv> <nvdisasm -g -c listing of the kernel> <numel>"""
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
# phase = the line range between two markers of maf_element.cuh (found by text, so that edits do not shift them)
MARK = [("gather", "MAF_HD void build_basis_block"), ("interp", "MAF_HD void phase_interp"),
        ("gauss_item", "MAF_HD int a_index"), ("residual", "MAF_HD void phase_residual"),
        ("blk_short", "MAF_HD int ch_fo"), ("blk_tr", "MAF_HD void block_accumulate_tr"),
        ("blk_mesh", "MAF_HD void block_accumulate_mesh"), ("blk_fused", "MAF_HD void block_accumulate_fused"),
        ("scatter", "struct KSink"), ("task_setup", "template <int MOTION> struct SmallUnroll"),
        ("phase_tangent", "MAF_HD void phase_tangent(")]
import os
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "membranealefem.jl_b200", "csrc",
                        "maf_element.cuh")).read().splitlines()
starts = []
for name, text in MARK:
    ln = next(i + 1 for i, l in enumerate(src) if text in l)
    starts.append((name, ln - (1 if name in ("blk_tr", "blk_mesh") else 0)))
R = [(n, lo, (starts[i + 1][1] - 1) if i + 1 < len(starts) else len(src)) for i, (n, lo) in enumerate(starts)]
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
