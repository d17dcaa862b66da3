"""Summarise an ncu launch list (--csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]).
usage: python tools/launch_summary.py launches.csv [launches_per_step]"""
import collections
import csv
import re
import sys


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "")
    name = re.sub(r"\(.*$", "", name)
    return name[:64]


def family(name):
    if name.startswith("igemm_halo_kernel") or name.startswith("igemm_fprop_tc") or name.startswith("igemm_fprop_mma_kernel"):
        return "conv fprop/dgrad (sv_igemm_fprop)"
    if name.startswith("wgrad_halo_kernel") or name.startswith("wgrad_tc_kernel") or name.startswith("igemm_wgrad_mma_kernel"):
        return "conv wgrad (sv_igemm_wgrad)"
    if name.startswith("wgrad_reduce"):
        return "wgrad reduce"
    if name.startswith("bn_bwd"):
        return "BatchNorm backward"
    if name.startswith("bn_"):
        return "BatchNorm forward"
    if name.startswith("linear") or name.startswith("log_softmax"):
        return "heads (linear)"
    if name.startswith("at::") or name.startswith("void at::"):
        return "torch fills / RNG / copies"
    return "losses, sampling, mixup, SGD, packing"


def main():
    rows = list(csv.reader(open(sys.argv[1], newline="")))
    per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 375
    i0 = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[i0]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = collections.OrderedDict()
    for r in rows[i0 + 1:]:
        if len(r) < len(hdr) or not r[0].isdigit():
            continue
        d = launches.setdefault(r[ix["ID"]], dict(name=short(r[ix["Kernel Name"]])))
        val = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        m = r[ix["Metric Name"]]
        if m.startswith("gpu__time_duration"):
            d["us"] = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
        else:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            d[m.split(".")[0]] = val * scale
    L = list(launches.values())
    total = sum(d.get("us", 0) for d in L)
    steps = len(L) / per_step
    print("# %d launches = %.2f steps of %d launches; serialised, cold-cache kernel time %.1f us = %.2f ms/step" %
          (len(L), steps, per_step, total, total / steps / 1e3))
    for title, keyf in (("family", lambda d: family(d["name"])), ("kernel", lambda d: d["name"])):
        agg = collections.OrderedDict()
        for d in L:
            a = agg.setdefault(keyf(d), dict(n=0, us=0.0, rd=0.0, wr=0.0))
            a["n"] += 1; a["us"] += d.get("us", 0); a["rd"] += d.get("dram__bytes_read", 0); a["wr"] += d.get("dram__bytes_write", 0)
        print("\n| %s | launches/step | ms/step | share | avg us | DRAM read+write MB/launch |\n|---|---|---|---|---|---|" % title)
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"])[:40]:
            print("| %s | %.1f | %.3f | %.1f %% | %.1f | %.2f |" % (k, a["n"] / steps, a["us"] / steps / 1e3, 100 * a["us"] / total, a["us"] / a["n"],
                                                                 (a["rd"] + a["wr"]) / a["n"] / 1e6))


if __name__ == "__main__":
    main()
