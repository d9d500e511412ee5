"""Aggregates an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel: launches, total and mean device time,
share of the listed time.  Launch lists are cold-cache and serialised: compare SHARES, not absolutes.

    python profiles/summarise_launches.py profiles/X_launches.csv [--ours] > profiles/X_launches.summary.txt
"""
import csv
import sys


def main():
    path = sys.argv[1]
    only_ours = "--ours" in sys.argv
    with open(path, newline="") as f:
        rows = [r for r in csv.reader(f) if len(r) > 10]
    header = rows[0]
    ki, vi = header.index("Kernel Name"), header.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        name = r[ki]
        if only_ours and "ctr::" not in name:
            continue
        short = name.replace("void ", "").replace("ctr::", "")
        short = short.split("(")[0][:70]
        n, t = agg.get(short, (0, 0.0))
        agg[short] = (n + 1, t + float(r[vi].replace(",", "")))
    total = sum(t for _, t in agg.values())
    print(f"# {path}: {sum(n for n, _ in agg.values())} launches, {total / 1e3:.1f} us listed")
    print("# kernel | launches | total us | mean us | share %")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k} | {n} | {t / 1e3:.1f} | {t / n / 1e3:.2f} | {100 * t / total:.1f}")


if __name__ == "__main__":
    main()
