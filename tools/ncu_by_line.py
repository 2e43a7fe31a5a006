"""Attribute the per-SASS-instruction samples of an ncu source-page CSV to source lines / functions.
usage: ncu_by_line.py <sass csv from `ncu -i rep --page source --csv --print-source sass`> <nvdisasm -g -c listing of the kernel>
       [numel]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
col = {n: i for i, n in enumerate(H)}
ins = rows[hdr + 1:]
lines = []
cur = ("?", 0)
inl = None
for l in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l):
        lines.append(cur)
assert len(lines) == len(ins), (len(lines), len(ins))
numel = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
S = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_s = tot_i = 0
stalls = [n for n in H if n.startswith("stall_") and "Not Issued" not in n]
for (f, ln), r in zip(lines, ins):
    s = int(r[col["# Samples"]]); n = int(r[col["Instructions Executed"]])
    key = (f, ln)
    S[key][0] += s; S[key][1] += n
    for st in stalls:
        v = int(r[col[st]])
        if v: S[key][2][st[6:]] += v
    tot_s += s; tot_i += n
print(f"total samples {tot_s}  warp instructions {tot_i}  per element {tot_i / numel:.0f}")
mode = sys.argv[4] if len(sys.argv) > 4 else "lines"
if mode == "lines":
    for key, (s, n, c) in sorted(S.items(), key=lambda kv: -kv[1][0])[:70]:
        top = ", ".join(f"{k}:{100 * v // max(s, 1)}%" for k, v in c.most_common(3))
        print(f"{key[0]}:{key[1]:<5d} samp {100 * s / tot_s:5.2f}%  inst/el {n / numel:7.1f}  {top}")
else:   # ranges of lines given as name=file:lo-hi,...
    pass
