"""Per-instruction stall sampling of a kernel from an ncu source page:

    ncu -i X.ncu-rep --page source --csv > X.source.csv
    python profiles/summarise_stalls.py X.source.csv [top] > profiles/X_stalls.txt

The page holds one block per profiled kernel: a header row (with a "Source" column, a "Warp Stall Sampling (All Samples)"
column and one "stall_*" column per reason) followed by one row per SASS instruction.  Prints, per block, the total samples, the
samples per stall reason and the `top` instructions by samples with their dominant reasons.
"""
import csv
import sys


def fnum(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return 0.0


def flush(title, header, rows, top):
    if not rows:
        return
    src = header.index("Source")
    samp = next((i for i, h in enumerate(header) if h.startswith("Warp Stall Sampling (All")), None)
    if samp is None:
        samp = next((i for i, h in enumerate(header) if h.startswith("# Samples") or h == "Samples"), None)
    stalls = [(i, h) for i, h in enumerate(header) if h.startswith("stall_") and "(" not in h]      # (all samples, not the "(Not Issued)" twins)
    total = sum(fnum(r[samp]) for r in rows) if samp is not None else 0.0
    print(f"## {title}: total samples {total:.0f}, instructions {len(rows)}")
    by_reason = sorted(((sum(fnum(r[i]) for r in rows), h) for i, h in stalls), reverse=True)
    print("   " + ", ".join(f"{h} {v:.0f}" for v, h in by_reason if v > 0))
    if samp is None:
        return
    order = sorted(range(len(rows)), key=lambda k: -fnum(rows[k][samp]))[:top]
    for k in order:
        r = rows[k]
        reasons = sorted(((fnum(r[i]), h[6:]) for i, h in stalls if fnum(r[i]) > 0), reverse=True)[:3]
        print(f"   {k:5d} {fnum(r[samp]):7.0f}  {r[src].strip()[:70]:70s} " + " ".join(f"{h}:{v:.0f}" for v, h in reasons))


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    header, rows, title, n = None, [], "", 0
    with open(path, newline="") as f:
        for r in csv.reader(f):
            if not r:
                continue
            if "Source" in r and any(h.startswith("stall_") or h.startswith("Warp Stall") for h in r):
                if header is not None:
                    flush(title, header, rows, top)
                header, rows, n = r, [], n + 1
                title = f"kernel {n}"
            elif header is not None and len(r) == len(header):
                rows.append(r)
            elif header is None and len(r) == 1:
                title = r[0]
    if header is not None:
        flush(title, header, rows, top)


if __name__ == "__main__":
    main()
