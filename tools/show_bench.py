"""Pretty-print a bench.py JSON line (main record + sub-records).   python tools/show_bench.py gpurun_out/bench.json"""
import json
import sys


def show(name, r):
    if "error" in r:
        print(f"{name}: ERROR {r['error']}")
        return
    print(f"{name}: {r['value']:.0f} formulas/s (e2e {r['e2e']['value']:.0f}), {r['ms_per_step']:.1f} ms/step x {r['steps']}, launches {r['gpu_launches']}, "
          f"n_gpus {r['n_gpus']}, scaling {r['scaling']}, batch/GPU {r['config']['batch_per_gpu']} {r['config']['image']}")
    rf = r.get("roofline")
    if rf:
        print(f"   dominant conv: {rf['achieved']:.0f} TFLOP/s = {rf['frac']:.3f} of {rf['peak']:.0f} ({rf['launch_ms']:.3f} ms/launch); "
              f"encode {rf['encode_ms']:.1f} ms, decode {rf['decode_ms']:.1f} ms")
    rd = r.get("roofline_decode")
    if rd:
        for k, v in rd.items():
            if isinstance(v, dict):
                print(f"   {k} ({rd['rows']} rows): {v['achieved']:.0f} GB/s = {v['frac']:.2f} of HBM peak, {v['avg_launch_us']:.1f} us/launch, "
                      f"{v['algorithmic_mb_per_launch']:.1f} MB/launch, n={v['launches_timed']}")


d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
show("main", d)
for k, v in d.get("records", {}).items():
    show(k, v)
print("clocks:", d.get("clocks"))
print("cpu_baseline:", d.get("cpu_baseline"))
