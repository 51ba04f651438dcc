import json, sys
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print("e2e", d["e2e"], "inference", d["inference"]["value"], d["inference"]["value_b32"], "cpu", d.get("cpu_baseline"))
r = d["roofline"]
print({k: r[k] for k in ("achieved", "frac", "share_of_step", "step_tflops", "step_frac_of_peak")})
for k, v in r["per_op"].items():
    tf = f"{v['tflops']:.0f} TF" if v["tflops"] else ""
    gb = f"{v['gbs']:.0f} GB/s" if v["gbs"] else ""
    print(f"  {k:28s} {v['ms_per_step']:.3f} ms  {tf}{gb}  n={v['launches']}")
