"""Condense an `ncu --page raw --csv` export into the handful of metrics the profiles/ summaries quote, one block per launch,
plus the time-weighted tensor-pipe activity of the whole list.
    ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv [title]"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.avg.per_second", "smsp__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic"]


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    if len(sys.argv) > 2:
        print(sys.argv[2])
    tw_num = tw_den = 0.0
    for k, r in enumerate(rows[2:]):
        print(f"\n---- launch {k}: {r[idx['Kernel Name']][:110]}")
        for w in WANT:
            if w in idx:
                print(f"{w}: {r[idx[w]]} {units[idx[w]]}")
        t = num(r[idx["gpu__time_duration.sum"]])
        u = units[idx["gpu__time_duration.sum"]]
        t_us = t * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        p = num(r[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]) or 0.0
        tw_num += t_us * p
        tw_den += t_us
    if tw_den:
        print(f"\ntime-weighted sm__pipe_tensor_cycles_active over these {len(rows) - 2} launches ({tw_den:.0f} us): {tw_num / tw_den:.1f} %")


if __name__ == "__main__":
    main()
