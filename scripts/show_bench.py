"""one-line summary of a bench.py JSON line (used by the gpurun sweep scripts)"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print(round(d["value"]), "Mbases/s", round(d["ms_per_step"], 1), "ms | e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 1),
      "ms |", {k: round(v, 1) for k, v in r["stage_ms_per_step"].items()}, "| kernel#2", round(r["achieved"]), "GB/s frac", round(r["frac"], 3),
      "| launches", d["gpu_launches"],
      "| alone:", {k: round(v, 1) for k, v in (r.get("standalone") or {}).get("stage_ms_per_step", {}).items()},
      round((r.get("standalone") or {}).get("frac") or 0, 3), round((r.get("standalone") or {}).get("ms_per_step_serial_schedule") or 0, 1))
