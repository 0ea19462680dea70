"""SASS evidence per kernel of the built library (run after __graft_entry__.build()):
   python profiles/sass_summary.py > profiles/sass_r02.txt
Counts the mnemonics that show which hardware paths a kernel uses: UBLKCP (cp.async.bulk = the TMA engine, 1-D),
UBLKPF (bulk L2 prefetch), SYNCS (mbarrier), ATOMS / ATOMG / REDG / RED (shared / global atomics, fire-and-forget
reductions), LDG / STG widths, BAR (CTA barriers), SHFL, MATCH."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "lua-mapreduce_b200", "lib", "libmrhbm.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
    usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
pats = ["UBLKCP", "UBLKPF", "SYNCS", "ATOMS", "ATOMG", "REDG", "RED.", "LDG.E.128", "LDG.E.64", "LDG.E ", "STG.E.128", "STG.E.64",
        "STG.E ", "LDS", "STS", "BAR.SYNC", "SHFL", "MATCH", "IMAD", "LOP3"]
print("kernel | sass instructions | regs stack smem(static) | " + " ".join(p.strip() for p in pats))
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n", 1)[0].strip()
    ops = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", part)
    c = collections.Counter()
    for o in ops:
        for p in pats:
            q = p.strip()
            if q.startswith(("LDG", "STG")):  # width classes: .128 / .64 / 32-bit, whatever cache modifiers sit in between
                width = "128" if ".128" in o else "64" if ".64" in o else ""
                want = "128" if q.endswith("128") else "64" if q.endswith("64") else ""
                if o.startswith(q[:3]) and width == want:
                    c[p] += 1
            elif o.startswith(q):
                c[p] += 1
    d = demangle(name)
    d = re.sub(r"\(.*", "", d).replace("void mrhbm::", "").replace("(int)", "").replace("(bool)", "")
    u = usage.get(name, ("?", "?", "?"))
    print("%s | %d | %s %s %s | %s" % (d, len(ops), u[0], u[1], u[2], " ".join(str(c[p]) for p in pats)))
