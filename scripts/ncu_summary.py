"""Turn an .ncu-rep into the per-kernel summary CSV kept under profiles/ (first launch of each distinct kernel)."""
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
           "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
           "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic"]


def main(rep, out, every=False):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{m} [{units[idx[m]]}]" for m in METRICS] + ["dram GB/s"])
        for r in data:
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("pdephi::", "")
            key = (name, r[idx["launch__grid_size"]], round(float(r[idx["dram__bytes_read.sum"]]), 1))
            if not every and key in seen:
                continue
            seen.add(key)
            ms = float(r[idx["gpu__time_duration.sum"]])
            gb = float(r[idx["dram__bytes_read.sum"]]) + float(r[idx["dram__bytes_write.sum"]])
            w.writerow([name] + [r[idx[m]] for m in METRICS] + [round(gb / ms * 1e3, 0)])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], "--every" in sys.argv)
