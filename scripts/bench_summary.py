#!/usr/bin/env python
"""Prints the numbers of one bench.py JSON line the way DESIGN.md quotes them."""
import json
import sys

d = json.load(open(sys.argv[1]))
r = d["roofline"]
print(f"N={d['n_gpus']} headline: value {d['value']:.0f} GB/s  ms/step {d['ms_per_step']:.4f}  kernel_ms {r['kernel_ms']:.4f}  "
      f"achieved {r['achieved']:.0f}  frac {r['frac']:.3f}  value/achieved {d['value'] / (r['achieved'] * d['n_gpus']):.3f}  "
      f"matches {d['matches']}  launches {d['gpu_launches']}")
if d["n_gpus"] > 1:
    print("  kernel_ms/rank", [round(x, 4) for x in d["kernel_ms_per_rank"]], " step/rank", [round(x, 4) for x in d["ms_per_step_per_rank"]],
          " exchange_ms/rank", [round(x, 4) for x in d["exchange_ms_per_rank"]], " rank0 host ms", round(d["rank0_host_ms_per_step"], 4))
if "e2e" in d:
    e = d["e2e"]
    print(f"  e2e {e['value']:.1f} GB/s  {e['ms_per_step']:.1f} ms  agrees {e['agrees_with_device_path']}  pinned alloc {e.get('pinned_alloc_s', 0):.1f} s")
if "cpu_baseline" in d:
    c = d["cpu_baseline"]
    print("  cpu", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in c.items() if k in ("value", "cores", "kind")},
          "t1", c.get("single_thread", {}).get("value"), "in_process", c.get("in_process", {}).get("value"))
print("  clocks", d.get("clocks"))
for name, w in d.get("workloads", {}).items():
    if "error" in w:
        print(f"  {name}: ERROR {w['error']}")
        continue
    rr = w["roofline"]
    cpu = w.get("cpu_baseline", {})
    print(f"  {name}: value {w['value']:.0f} GB/s  ms/step {w['ms_per_step']:.4f}  kernel_ms {rr['kernel_ms']:.4f}  frac {rr['frac']:.3f}  "
          f"matches {w['matches']}  exch {w['exchange_ms']:.4f}  cpu {cpu.get('value')}  ({w['filter']})")
