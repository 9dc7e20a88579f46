import json, sys
d = json.load(open(sys.argv[1]))
print("ms_per_step", round(d["ms_per_step"], 4), "value", f'{d["value"]:.3e}', "e2e ms", round(d["e2e"]["ms_per_step"], 3), "launches", d["gpu_launches"], "clocks", d["clocks"])
print({k: round(v, 4) for k, v in d["stage_ms"].items()})
print("roofline", d["roofline"])
if d.get("cpu_baseline"): print("cpu", {k: v for k, v in d["cpu_baseline"].items() if k != "sample"})
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
    print(f'  {k:60s} n/step={v["launches_per_step"]:5.1f} us/step={v["ms_per_step"]*1e3:9.1f}  us/launch={v["ms_per_launch"]*1e3:8.1f}')
if "sharded" in d: print("sharded", d["sharded"])
