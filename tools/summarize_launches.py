"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) by kernel name.

    python tools/summarize_launches.py gpurun_out/launches.csv profiles/r01_launches_c2.md "title"

ncu's per-launch times are cold-cache and serialised: read the SHARES, not the absolutes.
"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("plmc::", "")
    m = re.match(r"(?:void )?([A-Za-z0-9_:]+)(<.*>)?", name)
    base = m.group(1) if m else name
    tmpl = m.group(2) or "" if m else ""
    return base + tmpl


def main():
    src, dst = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else src
    rows = []
    with open(src, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0,
                 "second": 1e3}.get(unit, 1e-6)
        rows.append((short(r["Kernel Name"]), val * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for k, ms in rows:
        agg[k][0] += 1
        agg[k][1] += ms
    total = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n")
        f.write(f"source: `{src}` ({len(rows)} launches, {total:.1f} ms summed kernel time under ncu; "
                "cold-cache, serialised -- compare shares)\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {cnt} | {ms:.2f} | {100 * ms / total:.2f} % |\n")
    print(open(dst).read())


if __name__ == "__main__":
    main()
