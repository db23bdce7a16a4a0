"""Per-kernel SASS opcode counts of libsnb200.so (cuobjdump -sass): which kernels carry tcgen05 (UTC*MMA), TMEM loads (LDTM),
TMA (UTMALDG / UTMASTG / UBLKCP), legacy warp-level MMA (HMMA) ...   python profiles/sass_opcodes.py > profiles/r2_sass_opcodes.md"""
import collections
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(REPO, "structurednets_b200", "libsnb200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)")
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = pat.match(line)
    if m and cur:
        counts[cur][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
keys = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "HMMA", "LDGSTS", "SYNCS", "ELECT", "RED", "ATOM", "FFMA", "DFMA", "SHFL"]
print("# SASS opcode counts per kernel (libsnb200.so, sm_100a, `cuobjdump -sass`)\n")
print("`tcgen05.mma` shows as UTC*MMA, `tcgen05.ld` as LDTM, TMA as UTMALDG / UTMASTG / UTMAPF (prefetch) / UBLKCP (bulk copy), `mma.sync` as HMMA, "
      "`cp.async` as LDGSTS, mbarrier operations as SYNCS, `elect.sync` as ELECT (B200_PROFILING.md).  Static instruction counts.\n")
print("| kernel | SASS instr | " + " | ".join(keys) + " |")
print("|---|---|" + "---|" * len(keys))
rows = []
for f, name in zip(counts, names):
    c = counts[f]
    name = name.replace("(anonymous namespace)::", "").replace("void ", "")
    name = re.sub(r"\(.*", "", name).replace("snb::t3::", "").replace("snb::tc::", "").replace("snb::", "")
    rows.append((name, sum(c.values()), [sum(v for k, v in c.items() if k.startswith(key)) for key in keys]))
for name, tot, vals in sorted(rows):
    print("| `%s` | %d | " % (name[:70], tot) + " | ".join(str(v) if v else "" for v in vals) + " |")
