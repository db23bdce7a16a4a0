"""Prints a compact table from a bench.py JSON line (file argument)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("C5 N=%d: %.2f M/s  %.4f ms  step frac %.3f  e2e %.2f M/s  clocks %s" % (d["n_gpus"], d["value"] / 1e6, d["ms_per_step"], d["roofline"]["frac"],
                                                                              d["e2e"]["value"] / 1e6, d["clocks"]))
if "sustained" in d:
    s = d["sustained"]
    print("  sustained %.2f M/s (%.4f ms, %.1f s, %s)" % (s["value"] / 1e6, s["ms_per_step"], s["seconds"], s["clocks"]))
if d.get("dp_check") is not None:
    print("  dp_check", d["dp_check"], d.get("dp_check_detail"))
print("  kernels (us/step):", {k.replace("sss_tc_", "").replace("_kernel", ""): round(v["avg_ms"] * 1000 * v["launches"] / d["steps"], 1) for k, v in d["roofline"]["kernels"].items()})
print("  cpu", (d.get("cpu_baseline") or {}).get("value"), " gpu_aten", (d.get("gpu_aten_baseline") or {}).get("value"))
for k, w in d.get("workloads", {}).items():
    if "error" in w:
        print(k, w)
        continue
    r = w["roofline"]
    steps = 20
    print("%-16s %8.3f M/s  %.4f ms  hbm frac %.4f  tensor %s  top %s (%.0f%%)  cpu %s  aten-gpu %s" % (
        k, w["value"] / 1e6, w["ms_per_step"], r["frac"], ("%.3f" % r["tensor_frac"]) if "tensor_frac" in r else "-", r["kernel"].replace("_kernel", ""),
        100 * r["kernel_share"], round((w.get("cpu_baseline") or {}).get("value", 0)), round((w.get("gpu_aten_baseline") or {}).get("value", 0))))
    if "resident" in w:
        print("                 resident %.3f M/s %.4f ms" % (w["resident"]["value"] / 1e6, w["resident"]["ms_per_step"]))
