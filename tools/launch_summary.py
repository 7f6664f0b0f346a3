"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/launch_summary.py <launches.csv> [--last-step-from KERNEL_SUBSTRING]
Without the option: per-kernel-name count / total / average.  With it: the launches from the LAST launch whose name
contains the substring to the end of the list, in order (one step of the layer-wise path: pass k_lw_pe)."""
import csv
import sys
from collections import OrderedDict


def load(path):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
        rows.append((row["Kernel Name"], v, row["Grid Size"], row["Block Size"]))
    return rows


def main():
    rows = load(sys.argv[1])
    if len(sys.argv) > 3 and sys.argv[2] == "--last-step-from":
        key = sys.argv[3]
        idx = [i for i, r in enumerate(rows) if key in r[0] and "bwd" not in r[0]]
        tail = rows[idx[-1]:]
        print(f"# {sys.argv[1]}: {len(tail)} launches from the last '{key}', sum {sum(r[1] for r in tail):.1f} us "
              "(serialised, cold caches: shares, not absolutes)")
        for name, us, grid, block in tail:
            print(f"{name[:72]:72s} {us:8.1f} us  grid {grid} block {block}")
        return
    agg = OrderedDict()
    for name, us, _, _ in rows:
        k = name.split("(")[0][:80]
        c = agg.setdefault(k, [0, 0.0])
        c[0] += 1; c[1] += us
    print(f"# {sys.argv[1]}: {len(rows)} launches")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:80s} n={n:4d} total {t:10.1f} us  avg {t / n:8.1f} us")


if __name__ == "__main__":
    main()
