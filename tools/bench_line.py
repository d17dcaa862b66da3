"""one-line summary of a bench.py JSON line: python tools/bench_line.py tag file.json"""
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    l = json.loads(open(path).read().strip().splitlines()[-1])
    r = l.get("roofline") or {}
    print(tag, "ms/step", round(l["ms_per_step"], 3), "img/s", round(l["value"], 1), "e2e", round(l["e2e"]["value"], 1), "launches", l["details"]["launches_per_step"],
          "frac", round(r.get("frac", 0), 3), {k: (round(x["ms_per_step"], 2), round(x["tflops"])) for k, x in (r.get("step_share") or {}).items()})
except Exception as e:
    print(tag, "FAILED", repr(e))
