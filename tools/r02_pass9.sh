#!/bin/bash
# round 2, pass 9 (one GPU): GPU tests (FCI known answers, noise descriptor entry points), default bench line (boundary steps queued instead of
# awaited, boundary allocations in size classes, bra == ket flattened once), cfg5's MPO at reduced M with oracle parity (8 time-sliced shards)
mkdir -p gpurun_out
T=${1:-r02n}
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -6 gpurun_out/${T}_pytest_gpu.log
( time QCM_DEBUG=1 timeout 1200 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err
grep "bench rank 0\|rror\|real\|split seconds" gpurun_out/${T}_bench_cfg3.err | tail -12
( time timeout 900 python bench.py --config cfg5_54e54o_su2u1_M3000 --M 200 --slices 8 --steps 3 --warmup 3 --no-config-sweep --no-sweep ) > gpurun_out/${T}_bench_cfg5_M200.json 2> gpurun_out/${T}_bench_cfg5_M200.err
grep "bench rank 0\|rror\|real" gpurun_out/${T}_bench_cfg5_M200.err | tail -8
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_cfg3.json"))
    print("value %.2f TF/s  %.2f ms  e2e %.2f (engine %.2f) TF/s  phases %s parity %s frac %.3f roof %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["engine_mirror"]["value"], {k: round(v, 2) for k, v in d["roofline"]["phase_ms"].items()}, d.get("parity_rel_err_vs_oracle"), d["frac_of_fp64_peak"], d["roofline"]["frac"]))
    s = d.get("config_sweep", {})
    print("   sweep", {k: v for k, v in s.items() if k != "energies"})
    for w in d.get("sweep", []): print("   ", {k: w[k] for k in ("workload", "gpu_seconds_per_sweep", "cpu_seconds_per_sweep", "max_abs_energy_diff_vs_oracle") if k in w})
except Exception as e:
    print("failed", e)
try:
    d = json.load(open("gpurun_out/${T}_bench_cfg5_M200.json"))
    print("cfg5 M=200: value %.2f TF/s %.2f ms parity %s  cpu %s" % (d["value"], d["ms_per_step"], d.get("parity_rel_err_vs_oracle"), d.get("cpu_baseline")))
except Exception as e:
    print("cfg5 M200 failed", e)
PY
