"""SASS evidence for the tcgen05 / TMA kernels of libmmdyn_b200.so (runs without a GPU: cuobjdump -sass):
per kernel the counts of the Blackwell-native mnemonics (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM,
TMA -> UTMALDG / UBLKCP, griddepcontrol -> ACQBULK/PDL barriers), and the MMA issue loop of three kernels verbatim."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "multimodal-dynamics_b200", "libmmdyn_b200.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "UTMAPF", "STG.E.ENL2.256", "SYNCS", "ELECT", "HMMA", "REDG", "ATOMG"]


def main(out):
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = cur.replace("void ", "").replace("mmdyn::(anonymous namespace)::", "")
            cur = re.sub(r"\(.*", "", cur)
            while cur in funcs:
                cur += "'"
            funcs[cur] = []
        elif cur is not None and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            funcs[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line.rstrip()))
    with open(out, "w") as f:
        f.write("# SASS of libmmdyn_b200.so (sm_100a), cuobjdump -sass — mnemonic counts per kernel\n")
        f.write("# UTCHMMA = tcgen05.mma.kind::f16, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk,\n")
        f.write("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops; HMMA (legacy mma.sync) must be 0 everywhere\n\n")
        f.write(f"{'kernel':70s} {'instr':>6s} " + " ".join(f"{m[:8]:>8s}" for m in MNEMONICS) + "\n")
        for name, body in funcs.items():
            cnt = {m: sum(1 for l in body if re.search(r"\b" + re.escape(m), l)) for m in MNEMONICS}
            if cnt["UTCHMMA"] or cnt["UTMALDG"] or cnt["UBLKCP"] or cnt["LDTM"]:
                f.write(f"{name[:70]:70s} {len(body):6d} " + " ".join(f"{cnt[m]:8d}" for m in MNEMONICS) + "\n")
        for pat, title in (("igemm_tma_kernel<256, 0", "generic TMA-fed implicit GEMM, N = 256: MMA issuer (4 x UTCHMMA per 64-wide k-block, elected thread)"),
                           ("igemm_pair_kernel<128, 0>", "tile-pair kernel (deconv1.fwd, deconv2.dgrad): 8 x UTCHMMA per k-block, two accumulators, one weight descriptor"),
                           ("igemm_patch_kernel<128, 0, 7, 64, 4", "patch kernel (deconv3.fwd / conv2.dgrad): compile-time tap schedule, 44 UTCHMMA per tile, straight-line"),
                           ("wgrad_tma_kernel<256, 0, 2>", "weight-gradient kernel, both operands MN-major by TMA, two column blocks per CTA")):
            for name, body in funcs.items():
                if name.startswith(pat):
                    idx = [i for i, l in enumerate(body) if "UTCHMMA" in l]
                    if not idx:
                        continue
                    lo, hi = max(0, idx[0] - 25), min(len(body), idx[min(len(idx) - 1, 7)] + 12)
                    f.write(f"\n\n## {title}\n## {name}: instructions {lo}..{hi} of {len(body)}\n")
                    f.write("\n".join(body[lo:hi]) + "\n")
                    break
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_excerpt.txt"))
