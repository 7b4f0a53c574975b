"""Print the key numbers of a bench.py JSON line (last line of the given file)."""
import json
import sys

j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = j["roofline"]
print("score  %.3f M cand/s (e2e %.3f)  trmm frac %.3f  xcov %.2f ms  trmm %.2f ms  whole-step frac %.3f" % (
    j["value"] / 1e6, j["e2e"]["value"] / 1e6, r["frac"], r["xcov_ms_per_step"], r["trmm_ms_per_step"], r["whole_step_frac"]))
ll = j["loglik"]
print("loglik %.0f evals/s  %.2f ms/step  frac %.3f  (single stream %.2f ms)" % (
    ll["value"], ll["ms_per_step"], ll["frac_of_peak"], ll.get("single_stream_ms_per_step", float("nan"))))
print("loglik+grad %.0f evals/s  frac %.3f   fit %.2f ms" % (
    j["loglik_grad"]["value"], j["loglik_grad"]["frac_of_peak"], j["fit"]["ms_median"]))
for c in j.get("configs", []):
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in c.items()
           if "frac" in k or "ms_per" in k or "relerr" in k or k in ("config", "kernel")})
if "single_point" in j:
    print("single point: value %.0f us/call, value+grad %.0f us/call" % (j["single_point"]["value_call_us"], j["single_point"]["value_grad_call_us"]))
if "cpu_baseline" in j:
    print("cpu", j["cpu_baseline"].get("value"), j["cpu_baseline"].get("cores"))
