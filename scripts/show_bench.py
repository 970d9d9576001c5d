"""one-line digest of a bench.py JSON line: python scripts/show_bench.py <file> [--kernels]"""
import json, sys
d = json.load(open(sys.argv[1]))
r = d.get("roofline") or {}
print(f"{d['config'].get('workload')} N={d['n_gpus']} {d['value']:.1f} {d['unit']} {d['ms_per_step']:.4f} ms/step e2e {d['e2e']['value']:.1f} "
      f"graph={d['config'].get('cuda_graph')} clocks={d.get('clocks')} roofline {r.get('achieved', 0):.1f} {r.get('unit')} "
      f"burst {r.get('frac_of_burst_peak', 0):.3f} sustained {r.get('frac_of_sustained_peak', 0):.3f} launches {d.get('gpu_launches')}")
if "--kernels" in sys.argv:
    n = int(sys.argv[sys.argv.index("--kernels") + 1]) if sys.argv[-1] != "--kernels" else 99
    for k, v in list(d["kernel_breakdown_ms"].items())[:n]:
        print(f"  {k:32s} {v['launches']:4d} {v['ms']:.4f}")
