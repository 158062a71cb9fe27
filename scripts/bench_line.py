"""one-line summary of a bench.py JSON file"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f)); r = d["roofline"]
        print("%-44s B=%-6d S=%-5d ms/step %8.2f value %8.2fM e2e %8.2fM frac %.4f fwd %.2f bwd %.2f ms [%s | %s]" % (
            f.split("/")[-1], d["config"]["paths_per_gpu"], d["config"]["euler_steps"], d["ms_per_step"], d["value"] / 1e6,
            d["e2e"]["value"] / 1e6, r["frac"], r.get("fwd_kernel_ms", 0), r["kernel_ms"], r.get("fwd_kernel", ""), r.get("kernel", "")))
    except Exception as e:
        print(f, "fail", e)
