#!/usr/bin/env python
"""Pretty-print bench.py JSON lines: python scripts/show_bench.py gpurun_out/bench_*.json"""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable:", e)
        continue
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    cb = d.get("cpu_baseline") or {}
    print(f"== {path}: n_gpus={d.get('n_gpus')} value={d['value'] / 1e6:.3f} M {d['unit']}  {d['ms_per_step']:.3f} ms/step  "
          f"e2e={e.get('value', 0) / 1e6:.3f} M (h2d {e.get('h2d_bytes_per_step')} B, passes {e.get('passes_ms')})  "
          f"launches={d.get('gpu_launches')} graphs={d.get('cuda_graphs')} replica_diff={d.get('replica_max_abs_diff')}")
    if r:
        print(f"   roofline: {r['kernel']} {r['achieved']:.1f} {r['unit']} / {r['peak']:.0f} = {r['frac']:.3f}  "
              f"avg launch {r['avg_launch_ms']:.3f} ms x {r['launches_per_step']:.1f}/step")
    if cb:
        print(f"   cpu_baseline: {cb['value']:.1f} {cb['unit']} on {cb['cores']} cores ({cb['kind']})")
    for k, v in (d.get("kernels") or {}).items():
        print(f"   {k:30s} {v['launches_per_step']:6.1f} x  {v['ms_per_step']:8.3f} ms  {v['share_of_step'] * 100:5.1f}%")
