"""Static instruction mix of one kernel of libmembrane_b200.so (cuobjdump -sass): mnemonic histogram, the evidence for
"FP64 CUDA cores + shared memory + cp.async, no tensor cores, no TMA" (DESIGN.md section 4).

usage: python tools/sass_histogram.py <lib.so> <kernel name regex> <out.txt>"""
import collections
import re
import subprocess
import sys


def main():
    lib, pat, out = sys.argv[1], re.compile(sys.argv[2]), sys.argv[3]
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, body = None, collections.defaultdict(list)
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", line)
        if m and cur:
            body[cur].append(m.group(1).strip())
    names = [n for n in body if pat.search(n)]
    with open(out, "w") as f:
        for n in names:
            ins = body[n]
            ops = collections.Counter()
            for i in ins:
                i = re.sub(r"^@!?U?P\d+\s+", "", i)
                ops[i.split()[0].split(".")[0] + ("." + i.split()[0].split(".")[1] if i.split()[0].startswith(("LDS", "STS", "LDG", "STG", "LDL", "STL")) and "." in i.split()[0] else "")] += 1
            fam = collections.Counter()
            for k, v in ops.items():
                base = k.split(".")[0]
                fam[base] += v
            f.write(f"# {n}\n# {len(ins)} SASS instructions ({len(ins) * 16 / 1024:.1f} KB)\n")
            groups = [("FP64 arithmetic", ("DFMA", "DMUL", "DADD", "DSETP", "MUFU")),
                      ("shared memory", ("LDS", "STS", "LDSM", "ATOMS")),
                      ("asynchronous global->shared copies (cp.async)", ("LDGSTS", "LDGDEPBAR", "DEPBAR")),
                      ("global memory", ("LDG", "STG", "REDG", "RED", "ATOMG", "ATOM")),
                      ("local memory (spills)", ("LDL", "STL")),
                      ("tensor cores / TMA / TMEM (must be absent)", ("HMMA", "DMMA", "IMMA", "UTCHMMA", "UTCQMMA",
                                                                       "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM")),
                      ("barriers", ("BAR", "WARPSYNC", "BSYNC", "BSSY"))]
            for title, keys in groups:
                tot = sum(fam.get(k, 0) for k in keys)
                detail = ", ".join(f"{k} {fam[k]}" for k in keys if fam.get(k, 0))
                f.write(f"{title:52s} {tot:6d}   {detail}\n")
            f.write("top mnemonics: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(24)) + "\n\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
