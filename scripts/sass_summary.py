#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of liborbmatch_b200.so (cuobjdump -sass), trimmed to what the profiling recipe asks for:
the tensor-core / TMEM / TMA mnemonics (UTCQMMA, LDTM, UTMALDG, UBLKCP, UTCBAR ...), the integer pipe (POPC, LOP3, IADD3), shared /
global memory and barrier instructions.  Output: profiles/rNN_sass_summary.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "orb_slam3_comments_ghr_b200/liborbmatch_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEEP = ("UTCQMMA", "UTCHMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "POPC", "LOP3", "IADD3", "IMAD",
        "HMNMX2", "HMMA", "IMMA", "LDS", "STS", "LDG", "STG", "LDGSTS", "ATOMS", "ATOMG", "RED", "BAR", "MEMBAR", "FENCE", "REDUX", "SHFL", "VOTE", "MATCH",
        "MUFU", "DMUL", "DADD", "DSETP", "F2F", "CCTL", "ERRBAR", "NANOSLEEP", "ELECT", "UCGABAR", "ACQBULK", "USETMAXREG")
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").replace("void ", "")
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        op = m.group(1)
        mods = m.group(2)
        hist[kern]["_total"] += 1
        if op.startswith(KEEP):
            key = op + (mods if op.startswith(("UTC", "LDTM", "UTMA", "UBLKCP", "SYNCS", "HMMA", "IMMA")) else "")
            hist[kern][key] += 1
print(f"# SASS opcode summary of {so} (sm_100a); per kernel: total instructions, then the kept mnemonics by count")
for k, h in hist.items():
    tot = h.pop("_total", 0)
    if tot == 0:
        continue
    items = ", ".join(f"{op} {n}" for op, n in sorted(h.items(), key=lambda x: -x[1]))
    print(f"\n{k}  [{tot} instructions]\n    {items}")
