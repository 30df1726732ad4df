#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2 3 4 7; do
  QCM_FORCE_VARIANT=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_v$v.json"))
    print("variant $v: %.2f TF/s  %.2f ms  phases %s" % (d["value"], d["ms_per_step"], {k: round(x, 2) for k, x in d["roofline"]["phase_ms"].items()}))
except Exception as e:
    print("variant $v failed", e, open("gpurun_out/bench_v$v.err").read()[-500:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -s 10 -c 1 -o gpurun_out/prof_gemm_step3b -f python tools/profile_sigma.py cfg2_10e26o_su2u1_M1000 1 > gpurun_out/ncu_step3b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -s 5 -c 1 -o gpurun_out/prof_wgemm64 -f python tools/profile_sigma.py cfg2_10e26o_su2u1_M1000 1 > gpurun_out/ncu_wgemm64.log 2>&1
tail -2 gpurun_out/ncu_step3b.log gpurun_out/ncu_wgemm64.log
