"""SASS opcode histogram per kernel of lightgbm_b200/lib/liblgbm_b200.so (cuobjdump -sass), the evidence that the shipped
library uses TMA / mbarrier / native shared-memory integer atomics and no CAS loops.  Writes profiles/<name>."""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "lightgbm_b200/lib/liblgbm_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
keys = ("UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "ATOMS", "ATOMG", "REDG", "RED.", "LDS", "STS", "CCTL", "BAR", "SHFL", "DFMA", "MUFU")
for k, h in hist.items():
    tot = sum(h.values())
    sel = {op: n for op, n in sorted(h.items()) if any(op.startswith(p) for p in keys)}
    print(f"{k}: {tot} instructions")
    for op, n in sel.items():
        print(f"    {op:40s} {n}")
