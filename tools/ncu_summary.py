"""Summarise a .ncu-rep (read on the CPU box): key raw metrics + top stall instructions."""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== {name}")
        for i, h in enumerate(hdr):
            if h in KEYS:
                print(f"  {h:75s} {r[i]:>16s} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    body = [r for r in rows[2:] if len(r) > 5 and r[2].isdigit()]
    tot = sum(int(r[2]) for r in body) or 1
    print(f"-- top warp-stall sampling sites (of {tot} samples)")
    for r in sorted(body, key=lambda r: -int(r[2]))[:14]:
        print(f"  {100 * int(r[2]) / tot:5.1f}%  {r[1].strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1])
