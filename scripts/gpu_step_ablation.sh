#!/bin/bash
# Full pre-training step at B=256: per-shape GEMM table (bench.py's profiled steps, XLX_GEMM_LOG) with and without the
# epilogues' global traffic (XLX_GEMM_DEBUG=1: results wrong, timing only) → which GEMMs are epilogue-exposed?
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.build()" > /dev/null 2>&1
for d in 0 1; do
  XLX_GEMM_DEBUG=$d XLX_GEMM_LOG=gpurun_out/step_gemm_debug$d.csv timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --no-extra 2>/dev/null | tail -1 | cut -c1-200
done
python - <<'P'
import csv, collections
def agg(path):
    a = collections.OrderedDict()
    for r in csv.DictReader(open(path)):
        k = (int(r["M"]), int(r["N"]), int(r["K"]), int(r["a_mn"]), int(r["b_mn"]), int(r["epi"]))
        a.setdefault(k, [0, 0.0]); a[k][0] += 1; a[k][1] += float(r["us"])
    return a
a0, a1 = agg("gpurun_out/step_gemm_debug0.csv"), agg("gpurun_out/step_gemm_debug1.csv")
rows = sorted(a0.items(), key=lambda kv: -(kv[1][1] - a1.get(kv[0], [0, 0.0])[1]))
print("M,N,K,a_mn,b_mn,epi_flags,launches(3 steps),us_total,us_no_epilogue_traffic,exposed_us,tflops_algorithmic")
tot0 = tot1 = 0.0
for k, (n, us) in rows:
    us1 = a1.get(k, [0, 0.0])[1]
    tot0 += us; tot1 += us1
    print(",".join(map(str, k)) + f",{n},{us:.0f},{us1:.0f},{us - us1:.0f},{2.0 * k[0] * k[1] * k[2] * n / us / 1e6:.0f}")
print(f"# total {tot0 / 3e3:.2f} ms per step, {tot1 / 3e3:.2f} without epilogue traffic")
P
