"""Print the headline fields of a bench.py JSON line."""
import json, sys
j = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
sk = j.get("sinkhorn_onchip") or {}
print("steps/s %.2f  ms/step %.3f  e2e %.2f  launches %d  spmm %.3f ms/launch (frac %.3f)  sinkhorn %.3f ms/solve  parity %s"
      % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["gpu_launches"], j["roofline"]["avg_launch_ms"],
         j["roofline"]["frac"], sk.get("ms_per_solve", float("nan")),
         {k: ("%.1e" % v) for k, v in (j.get("parity") or {}).items() if isinstance(v, float)}))
