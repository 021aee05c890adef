#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`): per kernel count, total,
share of the listed time, average; with --per-launch also the largest launches and the first 30 of each kernel.
    python tools/launch_summary.py gpurun_out/r2f_launches_c3_1000x8192.csv [--second-half] [--per-launch]
--second-half: the command ran warm-up + 1 timed step of equal shape; keep the timed half of every kernel's launches."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    second_half = "--second-half" in sys.argv
    per_launch = "--per-launch" in sys.argv
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", "")
        per.setdefault(name, []).append(float(row["Metric Value"].replace(",", "")) / 1e3)   # us
    if second_half:
        per = {n: v[len(v) // 2:] for n, v in per.items()}
    total = sum(sum(v) for v in per.values())
    print(f"# {path}{' (timed half)' if second_half else ''}: {sum(len(v) for v in per.values())} launches, "
          f"{total / 1e3:.1f} ms listed (ncu serialises launches and runs them cold: shares, not absolutes)")
    print(f"{'kernel':44s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>10s} {'max us':>10s}")
    for n, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        print(f"{n:44s} {len(v):8d} {sum(v) / 1e3:10.2f} {100 * sum(v) / total:6.1f}% {sum(v) / len(v):10.1f} {max(v):10.1f}")
    if per_launch:
        for n, v in per.items():
            s = sorted(v, reverse=True)
            print(f"\n{n}: top 10 us {[round(x) for x in s[:10]]}  median {s[len(s) // 2]:.1f}")
            print(f"   first 30 launches: {[round(x) for x in v[:30]]}")


if __name__ == "__main__":
    main()
