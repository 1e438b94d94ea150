"""Pretty-print a bench.py JSON line: python tools/show_bench.py gpurun_out/.../bench.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"], 1), d["unit"], "| ms/step", round(d["ms_per_step"], 3), "| e2e", d.get("e2e") and round(d["e2e"]["value"], 1),
      "| n_gpus", d["n_gpus"], "| launches/step", d.get("gpu_launches_per_step"))
print("parity", d.get("parity_check"))
print("clocks", d.get("clocks"))
r = d.get("roofline")
if r:
    print({k: v for k, v in r.items() if k not in ("by_kernel", "peak_source", "traffic_note")})
    for k, v in r["by_kernel"].items():
        print(f"  {k:34s} calls {v['calls_per_step']:.0f} ms/call {v['ms_per_call']:.3f} share {v['share_of_step']:.3f} "
              f"tf {v['tflops'] and round(v['tflops'], 1)} gbs {v['gbs'] and round(v['gbs'])} {v.get('bound')} frac {v.get('frac') and round(v['frac'], 3)}")
for tag, b in (d.get("c3_bf16") or {}).items():
    print("c3", tag, round(b["value"], 1), "samples/s", round(b["ms_per_step"], 3), "ms | batch/gpu", b["config"]["batch_per_gpu"], "x",
          b["config"]["micro_batches_per_step"], "| e2e", b.get("e2e") and round(b["e2e"]["value"], 1), "| parity", b.get("parity_check"))
    if "roofline" in b:
        for k, v in b["roofline"]["by_kernel"].items():
            print(f"     {k:34s} calls {v['calls_per_step']:.0f} ms/call {v['ms_per_call']:.3f} share {v['share_of_step']:.3f} "
                  f"tf {v['tflops'] and round(v['tflops'], 1)} gbs {v['gbs'] and round(v['gbs'])} {v.get('bound')} frac {v.get('frac') and round(v['frac'], 3)}")
if "sweep" in d:
    for k, v in d["sweep"].items():
        print("  batch", k, v)
print("cpu_baseline", d.get("cpu_baseline"))
