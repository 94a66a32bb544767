"""SASS mnemonic histogram per kernel of the built library (evidence for tcgen05 / TMA use): python scripts/sass_hist.py > profiles/sass_rNN.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "vjf_b200", "lib", "libvjf_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEEP = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "HMMA", "DFMA", "LDGSTS", "ELECT", "R2UR", "LDS", "STS", "LDL", "STL", "SHFL", "MUFU", "FFMA", "BAR", "ATOMG", "LDG", "STG"]
print("# SASS mnemonic histogram per kernel of vjf_b200/lib/libvjf_b200.so (cuobjdump -sass; scripts/sass_hist.py)")
print("# tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk(.tensor) -> UBLKCP / UTMALDG, mma.sync -> HMMA, mbarrier -> SYNCS\n")
fn, hist, n = None, None, 0
def flush():
    if fn:
        print(f"{fn}: {n} instructions")
        print("   " + "  ".join(f"{k}={hist[k]}" for k in KEEP if hist.get(k)))
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush(); fn, hist, n = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and fn:
        n += 1
        hist[m.group(1).split(".")[0]] += 1
flush()
