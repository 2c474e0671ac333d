"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv)
into per-kernel time shares and, for the histogram kernel, DRAM bytes per launch (-> profiles/hist_traffic.json)."""
import collections, csv, json, sys

UNIT = {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0, "nsecond": 1e-9,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, out_json=None, workload="", n_gpus=1):
    rows = list(csv.reader(open(path, errors="replace")))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H = rows[h]
    ki, mi, ui, vi, idi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Unit"), H.index("Metric Value"), H.index("ID")
    per = collections.defaultdict(lambda: collections.defaultdict(dict))      # kernel -> launch id -> metric -> value (SI)
    for r in rows[h + 1:]:
        if len(r) <= vi or not r[vi]:
            continue
        name = r[ki].split("(")[0].replace("void ", "").split("<")[0]
        per[name][r[idi]][r[mi]] = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
    # complete trees = the launches between the first and the last k_prep of the capture
    preps = sorted(int(i) for i in per.get("k_prep", {}))
    trees = 1
    if len(preps) >= 2:
        lo, hi = preps[0], preps[-1]
        trees = len(preps) - 1
        per = {k: {i: m for i, m in v.items() if lo <= int(i) < hi} for k, v in per.items()}
        per = {k: v for k, v in per.items() if v}
        print(f"# restricted to {trees} complete tree(s): launch ids [{lo}, {hi})")
    tot = sum(m.get("gpu__time_duration.sum", 0.0) for k in per.values() for m in k.values())
    print(f"{'kernel':18s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'median us':>10s}")
    for name, ls in sorted(per.items(), key=lambda kv: -sum(m.get('gpu__time_duration.sum', 0) for m in kv[1].values())):
        t = sorted(m.get("gpu__time_duration.sum", 0.0) for m in ls.values())
        print(f"{name:18s} {len(t):8d} {sum(t) * 1e3:10.3f} {sum(t) / tot:7.3f} {t[len(t) // 2] * 1e6:10.1f}")
    hist = {k: v for k, v in per.items() if k.startswith("k_hist") and k != "k_hist_signal"}
    if out_json and hist:
        ls = [m for v in hist.values() for m in v.values()]
        rd = sum(m.get("dram__bytes_read.sum", 0.0) for m in ls); wr = sum(m.get("dram__bytes_write.sum", 0.0) for m in ls)
        n_a = max(1, len(hist.get("k_hist_a", hist.get("k_hist_q", ls))))      # one k_hist_reduce follows every k_hist_a: count the pair once
        json.dump({"kernel": "k_hist_a" if "k_hist_a" in hist else "k_hist_q", "includes": sorted(hist), "workload": workload, "n_gpus": int(n_gpus), "launches": n_a, "trees": trees,
                   "dram_bytes_per_launch": (rd + wr) / n_a, "dram_read_bytes_total": rd, "dram_write_bytes_total": wr,
                   "source": f"{path} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none)",
                   "hist_time_share_under_ncu": sum(m.get("gpu__time_duration.sum", 0.0) for m in ls) / tot}, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else "",
         sys.argv[4] if len(sys.argv) > 4 else 1)
