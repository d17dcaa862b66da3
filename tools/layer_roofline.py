"""Per-launch roofline table from a `bench.py --dump-kernels` file: for every instrumented launch of the step the algorithmic
FLOPs and bytes, the two roofline times t_tensor = FLOP / sustained bf16 peak and t_hbm = bytes / HBM copy peak
(MEASURED_PEAKS.json), the bound max(t_tensor, t_hbm) and the fraction of it the measured CUDA-event time reaches.
usage: python tools/layer_roofline.py kernels.json [title] > profiles/<name>.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    d = json.load(open(sys.argv[1]))
    title = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        src = "MEASURED_PEAKS.json"
    except Exception:
        pk = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
        src = "fallback of B200_PROFILING.md"
    tf, bw = pk["bf16_tflops_sustained"] * 1e12, pk["hbm_gbs"] * 1e9
    print("# %s\n" % title)
    print("Peaks: %.1f TFLOP/s sustained bf16, %.1f GB/s HBM copy (%s).  `us` = CUDA events around the launch in an instrumented eager\n"
          "replay (includes ~3 us of launch gap); algorithmic FLOPs / bytes per SURVEY section 8(d).  frac = max(t_tensor, t_hbm) / us.\n" %
          (pk["bf16_tflops_sustained"], pk["hbm_gbs"], src))
    fam = {}
    rows = []
    for k, v in d.items():
        kind, key = k.split(":", 1)
        n = v["n"]
        us = v["us"]
        tt, th = v["flops"] / tf * 1e6, v["bytes"] / bw * 1e6
        bound = max(tt, th)
        rows.append((kind, key, n // 3, us, v["flops"] / 1e9, v["bytes"] / 1e6, tt, th, bound / us if us > 0 else 0.0))
        f = fam.setdefault(kind, [0.0, 0.0, 0.0, 0])
        f[0] += us * n / 3; f[1] += bound * n / 3; f[2] += v["flops"] * n / 3; f[3] += n // 3
    print("| family | launches/step | ms/step | roofline ms/step | frac | TFLOP/s |\n|---|---|---|---|---|---|")
    tot = [0.0, 0.0]
    for kind, f in sorted(fam.items(), key=lambda kv: -kv[1][0]):
        print("| %s | %d | %.3f | %.3f | %.2f | %.0f |" % (kind, f[3], f[0] / 1e3, f[1] / 1e3, f[1] / f[0], f[2] / (f[0] * 1e-6) / 1e12))
        tot[0] += f[0]; tot[1] += f[1]
    print("| all instrumented launches | | %.3f | %.3f | %.2f | |\n" % (tot[0] / 1e3, tot[1] / 1e3, tot[1] / tot[0]))
    print("| launch | x/step | us | GFLOP | MB | t_tensor us | t_hbm us | bound | frac |\n|---|---|---|---|---|---|---|---|---|")
    for kind, key, n, us, gf, mb, tt, th, fr in sorted(rows, key=lambda r: -r[3] * r[2]):
        print("| %s %s | %d | %.1f | %.2f | %.1f | %.1f | %.1f | %s | %.2f |" % (kind, key, n, us, gf, mb, tt, th, "tensor" if tt >= th else "hbm", fr))


if __name__ == "__main__":
    main()
