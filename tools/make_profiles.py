"""Turns the files a tools/profile_pass_r2.sh run left in gpurun_out/ into the tracked summaries under profiles/."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def run(args, out):
    with open(out, "w") as f:
        f.write(subprocess.run(args, capture_output=True, text=True, cwd=ROOT).stdout)


def family(name):
    name = name.replace("void ", "").replace("mmdyn::<unnamed>::", "").replace("unnamed>::", "")
    m = re.match(r"([A-Za-z0-9_]+)", name)
    fam = m.group(1) if m else name
    return "igemm_tma_kernel" if fam == "igemm_pair_kernel" else fam  # one family in bench.py (ops.kernel_family)


def traffic(raw_csv, T):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units = rows[0], rows[1]
    ni = hdr.index("Kernel Name")
    ri, wi, ti = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows[2:]:
        b = float(r[ri].replace(",", "")) * mult[units[ri]] + float(r[wi].replace(",", "")) * mult[units[wi]]
        t = float(r[ti].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[ti], 1.0)
        a = agg[family(r[ni])]
        a[0] += 1
        a[1] += b
        a[2] += t
    return {k: {"launches_captured": v[0], "dram_bytes_per_launch": v[1] / v[0], "us_per_launch_under_ncu": v[2] / v[0]} for k, v in agg.items()}


def main(T):
    cmd = "python bench.py --batch 1024 --steps 1 --warmup 3 --no-graph --no-cpu --no-sweep --no-ref-cuda --no-sustained"
    run([sys.executable, "profiles/summarize_launches.py", f"gpurun_out/{T}_launches.csv",
         "Round 2 — ncu launch list of ONE steady-state training step, per-GPU batch 1024",
         f"ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 260 --csv {cmd}"], os.path.join(P, "r2_launches_b1024.md"))
    for what, title in (("gemm", "tcgen05 implicit-GEMM / patch / wgrad kernels"), ("bn", "BatchNorm streaming kernels, Adam")):
        out = os.path.join(P, f"r2_ncu_{what}_kernels.md")
        body = subprocess.run([sys.executable, "profiles/summarize_raw.py", f"gpurun_out/{T}_{what}_raw.csv"], capture_output=True,
                              text=True, cwd=ROOT).stdout
        with open(out, "w") as f:
            f.write(f"# Round 2 — ncu --set full, {title} of one training step, per-GPU batch 1024\n\n"
                    f"Command: `ncu --set full --clock-control none -k regex:... -s 100 -c N {cmd}`, exported with "
                    "`ncu -i ... --page raw --csv` (tools/profile_pass_r2.sh).\n\n" + body)
    tr = {"batch": 1024, "source": f"ncu --set full capture of this build (tools/profile_pass_r2.sh, {T}): mean of "
                                   "dram__bytes_read.sum + dram__bytes_write.sum over the captured launches of each kernel template"}
    fam = {}
    for what in ("gemm", "bn"):
        fam.update(traffic(os.path.join(G, f"{T}_{what}_raw.csv"), T))
    names = {"igemm_tma_kernel": "igemm_tma_kernel<*> + igemm_pair_kernel<*> (conv / deconv / linear forward + dgrad)",
             "igemm_patch_kernel": "igemm_patch_kernel<*> (merged 3x3-tap layers: deconv2/3/4 forward, conv2/3 dgrad)",
             "wgrad_tma_kernel": "wgrad_tma_kernel<*> (weight gradients)"}
    for k, v in fam.items():
        v["source"] = tr["source"]
        tr[names.get(k, k)] = v
    json.dump(tr, open(os.path.join(P, "r2_traffic.json"), "w"), indent=1)
    for rep, name in ((f"{T}_src_deconv1fwd", "r2_stalls_igemm_pair_128_deconv1fwd.txt"), (f"{T}_src_deconv3fwd", "r2_stalls_igemm_patch_128_deconv3fwd.txt")):
        if os.path.exists(os.path.join(G, rep + ".ncu-rep")):
            run([sys.executable, "tools/ncu_regions.py", f"gpurun_out/{rep}.ncu-rep", "80"], os.path.join(P, name))
    for src, dst in ((f"{T}_bench.json", "r2_bench_default.json"), (f"{T}_bench_ref.json", "r2_bench_reference_arm.json"),
                     (f"{T}_events.json", "r2_kernel_events_b1024.json"), (f"{T}_layers.txt", "r2_layers_microbench.txt"),
                     (f"{T}_bn.txt", "r2_bn_microbench.txt")):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
    keep = re.compile(r"rel|err|diff|step +\d|displacement|flip|passed|failed|trajectory|measured|B=|poisoned|noise floor")
    with open(os.path.join(G, f"{T}_gpu_tests.log")) as f, open(os.path.join(P, "r2_parity_measured.txt"), "w") as o:
        o.write("# measured parity numbers printed by `python -m pytest tests -m gpu -q -s` on a B200 (round 2)\n")
        for line in f:
            line = line.lstrip(".sF")
            if keep.search(line) and len(line) < 400:
                o.write(line)
    print("profiles written")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r2v")
