"""Selected metrics of an `ncu --set full` report, one block per captured launch (reads `ncu -i X.ncu-rep --page raw --csv`).

    python scripts/ncu_full_summary.py gpurun_out/prof_skinny.ncu-rep > profiles/rNN/ncu_full_<kernel>.txt
"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration (cold cache, serialised, under the profiler)"),
    ("launch__grid_size", "CTAs"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__cycles_active.avg", "SM active cycles"),
    ("sm__cycles_elapsed.max", "elapsed cycles"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: ncu --set full --clock-control none, {len(rows) - 2} launches")
    for r in rows[2:]:
        print(f"\n{r[hdr.index('Kernel Name')][:110]}")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print(f"  {label:<58} {r[i]:>14} {units[i]}")
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and float(r[i] or 0) > 0.5:
                print(f"  stall {h.split('stalled_')[1].split('_per_issue')[0]:<52} {float(r[i]):>14.2f} warps/issue")


if __name__ == "__main__":
    main(sys.argv[1])
