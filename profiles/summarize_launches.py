"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel markdown table."""
import collections
import csv
import re
import sys


def main(path, title, cmd):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("mmdyn::<unnamed>::", "")
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {title}\n\nCommand: `{cmd}`\n")
    print(f"{sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time "
          "(ncu per-launch times are cold-cache and serialised: compare SHARES, not absolutes).\n")
    print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:70]}` | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
