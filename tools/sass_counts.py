"""SASS opcode counts per kernel of the built library -> profiles/r02_sass_counts.md (static evidence of the hardware path:
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = bulk copy, LDGSTS = cp.async, HMMA = mma.sync).
usage: python tools/sass_counts.py [out.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "shot-vae_b200", "libshotvae.so")
PATS = collections.OrderedDict([("UTCHMMA", r"\bUTCHMMA"), ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"), ("LDTM (tcgen05.ld)", r"\bLDTM"),
                                ("UTMALDG (TMA tensor load)", r"\bUTMALDG"), ("UBLKCP (bulk copy)", r"\bUBLKCP"), ("LDGSTS (cp.async)", r"\bLDGSTS"),
                                ("SYNCS (mbarrier)", r"\bSYNCS"), ("HMMA (mma.sync)", r"\bHMMA"), ("STG", r"\bSTG"), ("LDG", r"\bLDG")])
SHOW = ("igemm_halo_kernel<9, 2", "igemm_halo_kernel<9, 4", "igemm_halo_kernel<1, 1, false", "igemm_fprop_tc", "wgrad_tc", "wgrad_halo_kernel<9>",
        "igemm_fprop_mma_kernel<64, 32>", "igemm_wgrad_mma_kernel<128>", "elbo_rec_vec_kernel<true, 3>", "mixup_image_vec_kernel<3>", "augment", "sgd_kernel",
        "bn_bwd_apply_kernel<1>", "bn_finalize_act_fwd")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", re.sub(r"\(anonymous namespace\)::", "", kern))
            counts[kern] = collections.Counter()
        elif kern:
            for name, pat in PATS.items():
                if re.search(pat, line):
                    counts[kern][name] += 1
    lines = ["# Round 2 — SASS opcode counts of the shipped kernels (`cuobjdump -sass shot-vae_b200/libshotvae.so`)", "",
             "Static instruction counts per kernel (not executions): which hardware path each kernel is on — `UTCHMMA` = tcgen05.mma, `LDTM` =",
             "tcgen05.ld (TMEM → registers), `UTMALDG` = TMA tensor-map load, `UBLKCP` = 1-D bulk copy, `LDGSTS` = cp.async, `SYNCS` = mbarrier,",
             "`HMMA` = legacy mma.sync.  Regenerate with `python tools/sass_counts.py`.", "",
             "| kernel | " + " | ".join(PATS) + " |", "|---|" + "---|" * len(PATS)]
    for k in counts:
        if any(s in k for s in SHOW):
            lines.append("| `%s` | " % k + " | ".join(str(counts[k][n]) for n in PATS) + " |")
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    lines += ["", "Whole library (%d kernels): " % len(counts) + ", ".join("%s %d" % (n, tot[n]) for n in PATS)]
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_counts.md")
    open(path, "w").write("\n".join(lines) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
