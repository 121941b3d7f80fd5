"""Print the headline fields of a bench.py JSON line."""
import json, sys
j = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("steps/s %.2f  ms/step %.2f  e2e %.2f  launches %d  sinkhorn ms %.2f (share %.2f)  spmm frac %.3f  cpu %.3f steps/s"
      % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["gpu_launches"], j["roofline"]["avg_launch_ms"],
         j["roofline"]["share_of_step"], j["roofline_spmm"]["frac"], (j.get("cpu_baseline") or {}).get("value", float("nan"))))
