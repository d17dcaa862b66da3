"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def summarize(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ns = v * 1000 if unit.startswith("us") else (v if unit.startswith("ns") else v * 1e6)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"void |<unnamed>::|\(anonymous namespace\)::", "", name)
        agg[name][0] += 1
        agg[name][1] += ns
        tot += ns
    out = ["total %.1f us over %d launches" % (tot / 1e3, sum(v[0] for v in agg.values()))]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append("%-64s n=%4d total=%8.1f us avg=%7.2f us %5.1f%%" % (k[:64], v[0], v[1] / 1e3, v[1] / 1e3 / v[0], 100 * v[1] / tot))
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30))
