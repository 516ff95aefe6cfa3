#!/usr/bin/env python
"""SASS opcode evidence per kernel of libvqcpc_b200.so (B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
UTMALDG/UTMASTG = TMA, HMMA = legacy mma.sync, FFMA2 = packed fp32 FMA, STAS = st.async (DSMEM), SYNCS = mbarrier).
    python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "vectorquantizedcpc_b200", "libvqcpc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "HMMA", "FFMA2", "FFMA", "STAS", "SYNCS", "REDUX", "CREDUX",
       "MUFU", "SHFL", "BAR", "FMNMX", "FMNMX3", "LOP3", "LDS", "STS", "LDG", "STG", "LD", "ST", "ATOM", "RED", "LDL", "STL"]
kern, hist = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        kern = kern.replace("vqcpc::", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
print(f"# SASS opcode counts per kernel of {os.path.basename(lib)} (cuobjdump -sass; static instruction counts, sm_100a)")
print("# kernel | total | " + " ".join(KEY))
for k in sorted(hist, key=lambda k: -sum(hist[k].values())):
    h = hist[k]
    cells = " ".join(f"{key}={h[key]}" for key in KEY if h[key])
    print(f"{k[:100]:100s} | {sum(h.values()):6d} | {cells}")
