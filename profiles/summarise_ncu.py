"""Turns an `ncu --set full` report into the per-kernel summary table committed under profiles/ and into
profiles/r2_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch), which bench.py reads for
`roofline.traffic`.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv --print-units base > gpurun_out/X.raw.csv
    python profiles/summarise_ncu.py gpurun_out/X.raw.csv profiles/X.summary.txt [--traffic profiles/r2_traffic.json --source "<what was captured>"]

The raw page has one row per profiled launch, a header row of metric names and a second row of units.
"""
import csv
import json
import sys

COLS = [
    ("duration us", "gpu__time_duration.sum", 1e-3),
    ("dram read MB", "dram__bytes_read.sum", 1e-6),
    ("dram write MB", "dram__bytes_write.sum", 1e-6),
    ("dram % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    ("issue active %", "sm__inst_issued.avg.pct_of_peak_sustained_active", 1.0),
    ("L2 hit %", "lts__t_sector_hit_rate.pct", 1.0),
    ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
    ("regs", "launch__registers_per_thread", 1.0),
    ("grid", "launch__grid_size", 1.0),
]

# kernel-name prefix -> key under which bench.py looks the traffic up
TRAFFIC_KEYS = [
    ("emb_bwd_sweep_l1_kernel<4, 1", "emb_bwd_apply_fused"),
    ("emb_bwd_sweep", "emb_bwd_apply"),
    ("emb_pool_fwd_l1_kernel<4, 4, 0, 1", "emb_pool_fwd_fused"),
    ("emb_pool_fwd", "emb_pool_fwd"),
]


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except (ValueError, AttributeError):
        return float("nan")


def main():
    raw, out = sys.argv[1], sys.argv[2]
    traffic_path = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
    source = sys.argv[sys.argv.index("--source") + 1] if "--source" in sys.argv else raw
    with open(raw, newline="") as f:
        rows = [r for r in csv.reader(f) if r and r[0] != "" and not r[0].startswith("==")]
    header = rows[0]
    body = [r for r in rows[2:] if len(r) == len(header)]
    name_i = header.index("Kernel Name")
    idx = {}
    for title, metric, _ in COLS:
        cands = [i for i, h in enumerate(header) if h == metric]
        idx[title] = cands[0] if cands else None
    lines = ["# " + source,
             "# kernel | " + " | ".join(t for t, _, _ in COLS)]
    traffic = {}
    for r in body:
        vals = []
        for title, _, scale in COLS:
            i = idx[title]
            vals.append(fnum(r[i]) * scale if i is not None else float("nan"))
        name = r[name_i]
        short = name.replace("ctr::", "").replace("void ", "")[:46]
        lines.append(short + " | " + " | ".join(f"{v:.1f}" for v in vals))
        for prefix, key in TRAFFIC_KEYS:
            if prefix in name and key not in traffic:
                traffic[key] = {"kernel": name[:120], "dram_bytes": (vals[1] + vals[2]) * 1e6, "duration_us": vals[0]}
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    if traffic_path:
        with open(traffic_path, "w") as f:
            json.dump({"source": source, "kernels": traffic}, f, indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
