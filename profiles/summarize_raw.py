"""Per-kernel table from an `ncu --page raw --csv` export: duration, DRAM bytes, tensor-pipe and DRAM
utilisation, occupancy."""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rd MB"), ("dram__bytes_write.sum", "wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("launch__registers_per_thread", "regs"), ("launch__waves_per_multiprocessor", "waves")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    name_i, grid_i = hdr.index("Kernel Name"), hdr.index("Grid Size")
    print("| kernel | grid | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|---|" + "---:|" * len(COLS))
    for r in rows[2:]:
        name = r[name_i].replace("void ", "").replace("mmdyn::<unnamed>::", "")
        name = name.split("(")[0][:44]
        vals = []
        for key, _ in COLS:
            if key in hdr:
                i = hdr.index(key)
                v = r[i].replace(",", "")
                try:
                    f = float(v)
                    u = units[i]
                    if key.startswith("dram__bytes") and u == "byte":
                        f /= 1e6
                    elif key.startswith("dram__bytes") and u == "Kbyte":
                        f /= 1e3
                    elif key.startswith("dram__bytes") and u == "Gbyte":
                        f *= 1e3
                    if key == "gpu__time_duration.sum":
                        f = f / 1e3 if u == "ns" else (f * 1e3 if u == "ms" else f)
                    vals.append(f"{f:.1f}")
                except ValueError:
                    vals.append(v)
            else:
                vals.append("-")
        print(f"| `{name}` | {r[grid_i]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
