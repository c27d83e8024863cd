#!/usr/bin/env python
"""Condense `ncu -i <file>.ncu-rep --page raw --csv` (stdin) into the per-launch summary committed under profiles/.
One block per captured launch: `<metric> <value> <unit>` lines (the format bench.py's traffic_from_profile() parses).
    ncu -i gpurun_out/x.ncu-rep --page raw --csv | python tools/ncu_summary.py "header line" > profiles/x_summary.txt"""
import csv
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "nvlrx__bytes.sum", "nvltx__bytes.sum", "nvlrx__bytes.sum.per_second", "nvltx__bytes.sum.per_second",
        "nvlrx__bytes_data_user.sum", "nvltx__bytes_data_user.sum", "nvlrx__bytes_data_protocol.sum", "nvltx__bytes_data_protocol.sum",
        "lts__t_sectors_srcunit_ltcfabric.sum", "lts__t_bytes_srcunit_ltcfabric.sum", "pcie__read_bytes.sum", "pcie__write_bytes.sum"]


def main():
    rows = list(csv.reader(sys.stdin))
    if len(sys.argv) > 1:
        print(sys.argv[1])
        print()
    if len(rows) < 3:
        print("(no launches in the report)")
        return
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        for name in WANT:
            for h, u, v in zip(hdr, units, vals):
                if h == name:
                    print("%-82s %20s %s" % (h, v, u))
        print()


if __name__ == "__main__":
    main()
