"""Turn an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python
tools/profile_step.py`) into the per-kernel share tables kept under profiles/.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rN_launches_summary.md

Segments: everything before the first `embed_rows_kernel` launch is weight fill + CLIP + prefill; every decode step starts with one
`embed_rows_kernel`, so the last step is the tail from the last such launch.  Per-launch times under ncu are cold-cache and
serialised (no PDL overlap, no graph): the SHARES are the evidence, not the absolute times.
"""
import csv
import re
import sys
from collections import OrderedDict


def load(path):
    rows, header = [], None
    for rec in csv.reader(open(path, newline="")):
        if header is None:
            if rec and rec[0] == "ID":
                header = rec
            continue
        if len(rec) != len(header):
            continue
        d = dict(zip(header, rec))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        unit = d.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        name = re.sub(r"\(.*$", "", d["Kernel Name"]).strip()
        name = re.sub(r"<.*$", "", name)
        name = re.sub(r"^void\s+", "", name)
        if name.startswith("at::native::") or name.startswith("at::"):
            name = "at:: (torch elementwise / copy / fill)"
        rows.append((name, us))
    return rows


def table(rows, title):
    agg = OrderedDict()
    for n, us in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(v[1] for v in agg.values()) or 1.0
    out = [f"## {title}: {len(rows)} launches, {tot / 1e3:.2f} ms (serialised, cold cache)", "",
           "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{n}` | {c} | {t:.1f} | {t / c:.1f} | {100 * t / tot:.1f}% |")
    return "\n".join(out) + "\n"


def main():
    rows = load(sys.argv[1])
    starts = [i for i, (n, _) in enumerate(rows) if n == "embed_rows_kernel"]
    print(f"# ncu launch list summary ({sys.argv[1]}): {len(rows)} launches\n")
    if not starts:
        print(table(rows, "all launches"))
        return
    print(table(rows[starts[-1]:], f"one decode step (the last of {len(starts)})"))
    print(table(rows[:starts[0]], "weight fill + CLIP + prefill"))


if __name__ == "__main__":
    main()
