"""Prints the headline numbers of a bench.py JSON line (default gpurun_out/enc2.json)."""
import json, sys
f = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/enc2.json"
d = json.loads(open(f).read().strip().splitlines()[-1])
print(f, "value %.3f M tok/s" % (d["value"] / 1e6), "ms/step %.4f" % d["ms_per_step"], "e2e %.3f M" % (d["e2e"]["value"] / 1e6),
      {k: round(v["ms"], 4) for k, v in d.get("stages", {}).items()})
