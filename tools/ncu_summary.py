"""Summarise an .ncu-rep (raw + source pages) into a small text file for profiles/."""
import csv, subprocess, sys, io

def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))

def main(rep, top=25):
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[-1]
    keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
            "sm__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed")
    for i, h in enumerate(hdr):
        if h == "Kernel Name": print("kernel:", vals[i])
        if h in keep or (h.startswith("smsp__average_warps_issue_stalled") and vals[i] not in ("0", "")):
            print(f"{h} [{units[i]}] = {vals[i]}")
    rows = page(rep, "source")
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
    tot = sum(int(r[ix['# Samples']]) for r in data)
    print(f"\n# top stalled SASS instructions ({tot} warp samples)")
    for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:top]:
        st = {s.replace('stall_', ''): int(r[ix[s]]) for s in hdr if s.startswith('stall_') and '(' not in s and r[ix[s]] not in ('', '0')}
        print(f"{data.index(r):4d} {r[ix['Source']].strip():50s} {r[ix['# Samples']]:>6s} x{r[ix['Instructions Executed']]:>8s} {st}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
