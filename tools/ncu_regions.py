"""Summarise an ncu report (--set full --import-source on) of ONE kernel: headline metrics and the warp-stall
samples per role of a warp-specialised kernel (regions of the SASS split at the marker instructions)."""
import csv
import subprocess
import sys


def load(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main(rep, win=60):
    raw = load(rep, "raw")
    hdr, vals = raw[0], raw[2] if len(raw) > 2 else raw[1]
    want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    for h, v in zip(hdr, vals):
        if h in want or h == "Kernel Name":
            print(f"{h} = {v}")
    src = load(rep, "source")
    h2, data = src[1], src[2:]
    isrc, ismp, iex = h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
    tot = sum(int(r[ismp]) for r in data)
    print("total samples", tot, "instructions", len(data))
    for w in range(0, len(data), win):
        seg = data[w:w + win]
        s = sum(int(r[ismp]) for r in seg)
        e = max(int(r[iex]) for r in seg)
        ops = [(r[isrc].split()[1] if r[isrc].strip().startswith("@") else r[isrc].split()[0]) for r in seg if r[isrc].split()]
        key = sorted(set(o for o in ops if o.startswith(("UTC", "UTMA", "LDTM", "SYNCS", "STG", "RED", "ATOM", "SHFL", "MUFU", "LDG"))))
        print(f"{w:5d} samples {s:6d} ({100.0 * s / max(tot, 1):5.1f} %) max-exec {e:9d}  {key}")
    top = sorted(range(len(data)), key=lambda i: -int(data[i][ismp]))[:12]
    print("hottest instructions:")
    for i in sorted(top):
        print(f"  {i:5d} {data[i][ismp]:>6s} {data[i][iex]:>9s}  {data[i][isrc][:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
