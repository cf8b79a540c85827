"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of device time per kernel
and the kernel sequence of one chunk of the recording pass.  Usage: python tools/summarize_launches.py file.csv"""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    seq = [(re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", ""), float(r["Metric Value"].replace(",", ""))) for r in rows]
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for k, v in seq:
        tot[k] = tot.get(k, 0.0) + v
        cnt[k] += 1
    T = sum(tot.values())
    print(f"# {path}: {len(seq)} launches, {T / 1e6:.3f} ms of device time (cold-cache, serialised: compare shares)")
    print("| share | total ms | launches | avg us | kernel |")
    print("|---:|---:|---:|---:|---|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:14]:
        print(f"| {100 * v / T:.1f}% | {v / 1e6:.3f} | {cnt[k]} | {v / cnt[k] / 1e3:.1f} | `{k[:80]}` |")
    starts = [i for i, (n, _) in enumerate(seq) if "first_conv" in n or "stage_first_conv_kernel" in n]
    if len(starts) > 2:
        print("\nkernel sequence of one chunk (us):")
        for n, v in seq[starts[1]:starts[2]]:
            print(f"  {v / 1e3:9.1f}  {n[:80]}")


if __name__ == "__main__":
    main(sys.argv[1])
