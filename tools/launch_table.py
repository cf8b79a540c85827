"""Per-launch table from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --csv` log: one line per launch longer than `min_us`.
    python tools/launch_table.py file.csv [first_launch_id] [min_us]"""
import csv
import sys


def main():
    f = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    min_us = float(sys.argv[3]) if len(sys.argv) > 3 else 5.0
    rows = list(csv.reader(l for l in open(f) if l.startswith('"')))
    h = rows[0]
    ix = {k: i for i, k in enumerate(h)}
    launches, order = {}, []
    for r in rows[1:]:
        i = int(r[ix["ID"]])
        if i not in launches:
            launches[i] = {"k": r[ix["Kernel Name"]].split("(")[0].replace("sdg::", "").replace("void ", "").replace("<unnamed>::", "")[:44]}
            order.append(i)
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u, m = r[ix["Metric Unit"]], r[ix["Metric Name"]]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
        if m.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1) / 1e6
        launches[i][m] = v
    total = 0.0
    print("| id | kernel | us | DRAM read MB | DRAM written MB | tensor pipe active % |")
    print("|---:|---|---:|---:|---:|---:|")
    for i in order:
        if i < skip:
            continue
        d = launches[i]
        t = d.get("gpu__time_duration.sum", 0.0)
        total += t
        if t >= min_us:
            print(f"| {i} | `{d['k']}` | {t:.1f} | {d.get('dram__bytes_read.sum', 0):.1f} | {d.get('dram__bytes_write.sum', 0):.1f} | "
                  f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):.1f} |")
    print(f"\ntotal of the listed range: {total:.0f} us")


if __name__ == "__main__":
    main()
