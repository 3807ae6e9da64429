#!/usr/bin/env python
"""Print the interesting fields of bench JSON lines: python tools/show_bench.py file.json [...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, "ERR", ex)
        try: print(open(f.replace(".json", ".err")).read()[-1500:])
        except Exception: pass
        continue
    e = d.get("e2e") or {}
    print(f"{f}: {d['ms_per_step']:.4f} ms  {d['value']:.0f} f/s  n_gpus={d.get('n_gpus')} e2e {e.get('ms_per_step')} lat {e.get('latency_ms')} inst {d.get('tile_instances')} sort_ms {d.get('sort_ms')}")
    km = d.get("kernel_ms") or {}
    print("    " + "  ".join("%s=%.4f" % (k.replace("k_radix_", "").replace("k_", ""), v) for k, v in km.items()))
