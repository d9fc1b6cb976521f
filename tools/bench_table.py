"""Print the per-entry-point table of a bench.py JSON line.  python tools/bench_table.py gpurun_out/bench_x.json"""
import json, sys
d = json.load(open(sys.argv[1]))
print(f"{d['value']:.0f} {d['unit']}  {d['ms_per_step']:.2f} ms/step  host enqueue {d.get('host_enqueue_ms_per_step', 0):.2f} ms  launches {d.get('gpu_launches')}  e2e {d['e2e']['value']:.0f}")
for k in d.get("kernels", []):
    extra = f"{k.get('achieved_gbs', 0):7.0f} GB/s" if k.get("bound") == "hbm" else (f"{k.get('achieved_tflops', 0):7.1f} TF/s" if k.get("bound") == "tensor" else "")
    print(f"  {k['entry']:30s} {k['launches_per_step']:4d}x {k['ms_per_step']:7.3f} ms {100 * k['share_of_step']:5.1f}%  frac {k.get('frac', 0):.3f} {extra}")
r = d.get("roofline", {})
print("roofline:", r.get("kernel"), f"achieved {r.get('achieved', 0):.1f} {r.get('unit')} frac {r.get('frac', 0):.3f}")
for g in r.get("gemm_shapes", []):
    print(f"  gemm M={g['M']:7d} N={g['N']:5d} K={g['K']:7d} t={g['trans']} {g['launches_per_step']:3d}x {g['ms_per_step']:7.3f} ms {g['tflops']:6.1f} TF/s")
for g in r.get("absmax_passes", []):
    print(f"  absmax {g['elements'] * 4 / 1e6:8.1f} MB {g['launches_per_step']:3d}x {g['ms_per_step']:7.3f} ms")
