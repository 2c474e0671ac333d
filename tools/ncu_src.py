"""Per-instruction stall summary of one launch of an .ncu-rep source page (csv exported with --kernel-id :::N)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[ix['# Samples']].isdigit()]
tot = sum(int(r[ix['# Samples']]) for r in data)
print("total samples", tot, "warp instructions", sum(int(r[ix['Instructions Executed']]) for r in data))
stall_cols = [h for h in hdr if h.startswith('stall_') and '(' not in h]
agg = {}
for r in data:
    for s in stall_cols:
        v = r[ix[s]]
        if v not in ('', '0'): agg[s] = agg.get(s, 0) + int(v)
print({k.replace('stall_', ''): v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:12]})
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']]))[:top_n]
for i in sorted(top):
    r = data[i]
    st = {s.replace('stall_', ''): int(r[ix[s]]) for s in stall_cols if r[ix[s]] not in ('', '0')}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print(f"{i:4d} {r[ix['Source']].strip()[:58]:58s} {r[ix['# Samples']]:>6s} x{r[ix['Instructions Executed']]:>8s} {st}")
