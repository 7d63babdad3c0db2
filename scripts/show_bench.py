#!/usr/bin/env python3
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    r = d["roofline"]
    print(f"{f}: {d['value']:.1f} Msamples/s  {d['mrays_per_s']:.0f} Mrays/s  {d['ms_per_step']:.1f} ms/step  e2e {d['e2e']['value']:.1f}  launches {d['gpu_launches']}")
    print("   kernel ms/step:", {k: round(v, 1) for k, v in r["kernel_ms_per_step"].items()}, " clocks:", d["clocks"])
    print(f"   roofline: {r['achieved']:.0f} GB/s = {r['frac']:.3f} of {r['peak']}; B/item {r['bytes_per_item']:.0f} (n_node {r['n_node']:.2f}, n_tri {r['n_tri']:.2f}); launch {r['launch_ms']:.3f} ms x {r['launches_per_step']}; share {r['share_of_step']:.2f}")
    if d.get("cpu_baseline"): print("   cpu:", d["cpu_baseline"])
