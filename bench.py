#!/usr/bin/env python
"""bench.py -- sigma-vector (site_hamil2) FP64 throughput on B200.

Step = one application of the effective Hamiltonian to the two-site tensor of a fabricated mid-chain site
problem (true MPO of a synthetic FCIDUMP, synthetic sector lists truncated to total bond dimension M, random
boundaries; see qcmaquis_b200/csrc/qcm/scenarios.hpp).  Default workload = the configuration BASELINE.json's
metric is quoted on, configs[2]: 24e/30o SU2U1 M=2000 two-site (one site problem needs 60 GB: it fits one B200).

  value     whole-job TFLOP/s = schedule-derived algorithmic FLOPs of one sigma / device time (CUDA events on
            the library stream, max over ranks), psi/sigma and boundaries resident in HBM
  e2e       same metric through the C ABI call qcm_site_hamil2 with pinned HOST buffers for psi and sigma
            (H2D + kernels + allreduce + D2H inside the timed region); boundaries and MPO stay resident, as
            they do for the whole Davidson solve at a site
  --impl reference   the CPU oracle (restatement of the reference algorithm, OpenMP over the MPO bond index,
            OpenBLAS dgemm) on the host cores, same instance, same metric

N>1: one process per GPU (torchrun); the edges of the MPO bond graph are sharded across ranks by their step-1 bond
index and the partial sigma vectors are summed with one NCCL allreduce per step inside the library (strong scaling:
the problem is fixed; value = FLOPs of the whole problem / max-over-ranks device time).
"""
import argparse, ctypes, json, os, subprocess, sys, tempfile, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (norb, nelec, symm, M)
    "cfg1_8e8o_su2u1_M256": (8, 8, "su2u1", 256),
    "cfg2_10e26o_su2u1_M1000": (26, 10, "su2u1", 1000),
    "cfg3_24e30o_su2u1_M2000": (30, 24, "su2u1", 2000),
    "cfg4_24e30o_2u1_M4000": (30, 24, "2u1", 4000),
    "cfg5_54e54o_su2u1_M3000": (54, 54, "su2u1", 3000),
}


def errbuf():
    return ctypes.create_string_buffer(1024)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = max(mx, float(s[1]))
                for i, n in enumerate(names):
                    if s[3 + i].lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def make_fcidump(norb, nelec):
    from qcmaquis_b200.fcidump import make_fcidump as mk
    d = tempfile.mkdtemp(prefix="qcm_bench_")
    path = os.path.join(d, "synthetic_%do%de.fcidump" % (norb, nelec))
    mk(path, norb, nelec)
    return path


def run_reference(args, cfg_name, norb, nelec, symm, M, site):
    """CPU arm: the oracle on the host cores (the upstream tree needs Boost/GSL/HDF5 and cannot be built here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from qcmaquis_b200 import build
    lib = ctypes.CDLL(build.build_oracle())
    lib.orc_create.restype = ctypes.c_void_p
    lib.orc_set_threads(len(os.sched_getaffinity(0)))
    e = errbuf()
    path = make_fcidump(norb, nelec)
    h = lib.orc_create(path.encode(), symm.encode(), norb, nelec, e, 1024)
    if not h:
        raise RuntimeError(e.value.decode())
    h = ctypes.c_void_p(h)
    pe = ctypes.c_double()
    if lib.orc_setup_site(h, site, 1, M, args.seed, ctypes.byref(pe), e, 1024):
        raise RuntimeError(e.value.decode())
    flops = reference_flops(path, symm, norb, nelec, site, M, args.seed)
    sec, ov, se = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    budget_s = 150.0
    t_begin = time.time()
    warm = min(args.warmup, 1)
    for _ in range(warm):
        lib.orc_sigma(h, 1, ctypes.byref(sec), ctypes.byref(ov), ctypes.byref(se), e, 1024)
    per = sec.value if warm else None
    steps = args.steps
    if per:
        steps = max(1, min(args.steps, int((budget_s - (time.time() - t_begin)) / per)))
    total = 0.0
    done = 0
    for _ in range(steps):
        if lib.orc_sigma(h, 1, ctypes.byref(sec), ctypes.byref(ov), ctypes.byref(se), e, 1024):
            raise RuntimeError(e.value.decode())
        total += sec.value
        done += 1
        if time.time() - t_begin > budget_s:
            break
    ms = total / done * 1e3
    val = flops / (ms * 1e-3) / 1e12
    cores = lib.orc_threads()
    sample = "%d full sigma evaluation(s) of the same instance (requested %d; capped to keep the run within minutes)" % (done, args.steps)
    line = {"impl": "reference", "metric": "sigma_vector_fp64_tflops", "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": done,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": cfg_name, "site": site, "twosite": True, "M": M, "symmetry": symm},
            "steps_note": "one step of this arm is one full sigma of the same instance (25-35 s on the host cores): steps and warm-up are capped so that the run ends within minutes",
            "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def reference_flops(path, symm, norb, nelec, site, M, seed):
    """Algorithmic FLOPs of one sigma from the schedule builder (host only, no GPU needed)."""
    from qcmaquis_b200 import build
    lib = ctypes.CDLL(build.build_host())
    lib.qcmd_create.restype = ctypes.c_void_p
    lib.qcmd_plan_flops.restype = ctypes.c_double
    e = errbuf()
    h = lib.qcmd_create(path.encode(), symm.encode(), norb, nelec, e, 1024)
    if not h:
        raise RuntimeError(e.value.decode())
    f = lib.qcmd_plan_flops(ctypes.c_void_p(h), site, 1, M, seed, e, 1024)
    if f < 0:
        raise RuntimeError(e.value.decode())
    lib.qcmd_destroy(ctypes.c_void_p(h))
    return f


def sweep_measurement(host, build, device, with_cpu):
    """Sweep level: DMRG sweeps (qcm/sweep.hpp, qcm/twosite.hpp) on the B200 engine and, same driver, on the CPU oracle.
    The first workload is BASELINE configs[0] as written: 8e/8o SU2U1, M=256, two-site, 4 sweeps."""
    results = []
    workloads = (("cfg1_8e8o_su2u1_M256_twosite_4sweeps", 8, 8, "su2u1", "ts", 16, 256, 4),
                 ("12e12o_su2u1_M300_twosite_1sweep", 12, 12, "su2u1", "ts", 60, 300, 1),
                 ("12e12o_su2u1_M300_singlesite_1sweep", 12, 12, "su2u1", "ss", 300, 300, 1))
    olib = None
    if with_cpu:
        olib = ctypes.CDLL(build.build_oracle())
        olib.orc_create.restype = ctypes.c_void_p
        olib.orc_set_threads(len(os.sched_getaffinity(0)))
    for name, norb, nelec, symm, kind, M0, M, nsweeps in workloads:
        path = make_fcidump(norb, nelec)
        e = errbuf()
        h = host.qcmd_create(path.encode(), symm.encode(), norb, nelec, e, 1024)
        if not h:
            raise RuntimeError(e.value.decode())
        h = ctypes.c_void_p(h)
        en = (ctypes.c_double * 4096)(); n = ctypes.c_int(); info = (ctypes.c_double * 8)()
        if kind == "ts":
            rc = host.qcmd_ts_sweeps(h, M0, M, nsweeps, 42, device, en, 4096, ctypes.byref(n), info, e, 1024)
        else:
            rc = host.qcmd_ss_sweeps(h, M, nsweeps, 42, device, en, 4096, ctypes.byref(n), info, e, 1024)
        if rc:
            raise RuntimeError(e.value.decode())
        host.qcmd_destroy(h)
        out = {"workload": name, "driver": "two-site (TwoSiteTensor + SVD truncation)" if kind == "ts" else "single-site (QR shift)",
               "eigensolver": "Jacobi-Davidson, <= 10 sigma per site, tol 1e-8", "start": "random MPS, M0 = %d" % M0,
               "gpu_seconds_per_sweep": info[1] / nsweeps, "sigma_evaluations": int(info[0]), "micro_iterations": n.value, "final_energy": info[2]}
        if kind == "ts":
            out["largest_bond_dimension"] = int(info[3])
        if olib is not None:
            oh = olib.orc_create(path.encode(), symm.encode(), norb, nelec, e, 1024)
            if not oh:
                raise RuntimeError(e.value.decode())
            oh = ctypes.c_void_p(oh)
            eo = (ctypes.c_double * 4096)(); no = ctypes.c_int(); io = (ctypes.c_double * 8)()
            if kind == "ts":
                rc = olib.orc_ts_sweeps(oh, M0, M, nsweeps, 42, eo, 4096, ctypes.byref(no), io, e, 1024)
            else:
                rc = olib.orc_ss_sweeps(oh, M, nsweeps, 42, eo, 4096, ctypes.byref(no), io, e, 1024)
            if rc:
                raise RuntimeError(e.value.decode())
            olib.orc_destroy(oh)
            out["cpu_seconds_per_sweep"] = io[1] / nsweeps
            out["cpu_cores"] = olib.orc_threads()
            out["max_abs_energy_diff_vs_oracle"] = (max(abs(en[i] - eo[i]) for i in range(n.value)) if n.value == no.value else "length mismatch")
            if os.environ.get("QCM_DEBUG"):
                sys.stderr.write("sweep %s: gpu/cpu energies %s\n" % (name, ["%.17g / %.17g" % (en[i], eo[i]) for i in (0, min(5, n.value - 1), n.value - 1)]))
        results.append(out)
    return results


def config_sweep(host, h, args, M, device, rank, world, max_over_ranks, sum_over_ranks, peak_tflops):
    """One two-site sweep (ts_optimize.hpp:60-270 semantics: Jacobi-Davidson <= 10 sigma per site, tol 1e-8, SVD truncation
    to M) of the benchmarked configuration on N GPUs.  Wall-clock of the sweep loop (the analogue of the reference's sweep
    timer, ts_optimize.hpp:267-269), max over ranks; the engine's host-side seconds are reported next to it."""
    e = errbuf()
    en = (ctypes.c_double * 8192)(); n = ctypes.c_int(); info = (ctypes.c_double * 32)()
    host.qcmd_ts_sweeps_synth.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_double, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
    if host.qcmd_ts_sweeps_synth(h, M, 1, args.seed, device, rank, world, 0, args.sweep_budget, en, 8192, ctypes.byref(n), info, e, 1024):
        raise RuntimeError(e.value.decode())
    norb = CONFIGS[args.config][0]
    sweep_s = max_over_ranks(info[1])
    f_sigma, f_bnd = sum_over_ranks([info[16], info[17]])
    names = ["two_site_tensor", "two_site_mpo", "eigensolver", "split", "boundary_step"]
    eng_names = ["planning", "plan_upload", "sigma_calls", "boundary_calls", "flatten"]
    drv = {k: info[4 + i] for i, k in enumerate(names)}
    engs = {k: info[9 + i] for i, k in enumerate(eng_names)}
    # host-exposed time: everything of the sweep loop that is not a device call (solver incl. its sigma evaluations, boundary-step execution)
    device_s = engs["sigma_calls"] + engs["boundary_calls"]
    out = {"workload": "%s two-site sweep from a random MPS on the synthetic sectors (M = %d at every bond)" % (args.config, M),
           "micro_iterations": int(info[19]), "micro_iterations_full_sweep": 2 * norb - 2, "complete": int(info[19]) == 2 * norb - 2,
           "sweep_s": sweep_s, "init_s": info[18], "sigma_evaluations": int(info[0]), "largest_bond_dimension": int(info[3]),
           "final_energy": info[2], "energies": [en[i] for i in range(n.value)],
           "driver_s": drv, "engine_s": engs, "device_call_s": device_s, "host_exposed_s": max(0.0, info[1] - device_s),
           "host_exposed_frac": max(0.0, info[1] - device_s) / info[1] if info[1] else None,
           "plan_cache": {"hits": int(info[14]), "misses": int(info[15])},
           "sum_flops": f_sigma + f_bnd, "sum_sigma_flops": f_sigma, "sum_boundary_flops": f_bnd,
           "tflops_over_sweep": (f_sigma + f_bnd) / sweep_s / 1e12 if sweep_s else None,
           "frac_of_fp64_peak": (f_sigma + f_bnd) / sweep_s / 1e12 / (peak_tflops * world) if sweep_s and peak_tflops else None}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="cfg3_24e30o_su2u1_M2000", choices=sorted(CONFIGS))
    ap.add_argument("--M", type=int, default=0)
    ap.add_argument("--site", type=int, default=-1)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the sweep-level measurement (single-site DMRG sweeps of the 8e/8o system)")
    ap.add_argument("--parity", action="store_true", help="(kept for compatibility: the oracle parity of the sigma vector is reported at every N unless --no-cpu-baseline)")
    ap.add_argument("--slices", type=int, default=0, help="N=1: run the site problem as this many time-sliced shards on the one GPU (0: automatic -- 2 for cfg4, whose resident step-1 products are 145 GB; 1 otherwise)")
    ap.add_argument("--no-config-sweep", action="store_true", help="skip the two-site DMRG sweep of the benchmarked configuration itself")
    ap.add_argument("--sweep-budget", type=float, default=300.0, help="wall-clock budget (s) of the configuration sweep; it stops at a site boundary when exceeded")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    norb, nelec, symm, M = CONFIGS[args.config]
    if args.M:
        M = args.M
    site = args.site if args.site >= 0 else norb // 2 - 1
    slices = args.slices if args.slices > 0 else (2 if args.config.startswith("cfg4") and int(os.environ.get("WORLD_SIZE", "1")) == 1 else 1)
    if slices > 1:
        os.environ["QCM_SLICES"] = str(slices)
    if args.impl == "reference":
        run_reference(args, args.config, norb, nelec, symm, M, site)
        return

    # stdout carries exactly one JSON line: whatever libraries print while the run is set up (NCCL's version banner,
    # for one) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    # torchrun exports OMP_NUM_THREADS=1; the host-side schedule builder is OpenMP-parallel: give every rank its share of
    # the host cores (must be set before the OpenMP runtime is loaded)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os.environ["OMP_NUM_THREADS"] = str(max(1, len(os.sched_getaffinity(0)) // int(os.environ.get("LOCAL_WORLD_SIZE", os.environ["WORLD_SIZE"]))))
    import torch
    import torch.distributed as dist
    from qcmaquis_b200 import build
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback for the native arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cu = ctypes.CDLL(build.build_cuda(), mode=ctypes.RTLD_GLOBAL)
    host = ctypes.CDLL(build.build_host())
    cu.qcm_last_error.restype = ctypes.c_char_p
    cu.qcm_stream.restype = ctypes.c_void_p
    cu.qcm_launch_count.restype = ctypes.c_int64
    host.qcmd_create.restype = ctypes.c_void_p
    if cu.qcm_init(local):
        raise RuntimeError(cu.qcm_last_error().decode())
    if world > 1:
        idbuf = ctypes.create_string_buffer(128)
        if rank == 0 and cu.qcm_comm_unique_id(idbuf):
            raise RuntimeError(cu.qcm_last_error().decode())
        t = torch.tensor(list(idbuf.raw), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        idbuf = ctypes.create_string_buffer(bytes(t.cpu().tolist()), 128)
        if cu.qcm_comm_init(rank, world, idbuf):
            raise RuntimeError(cu.qcm_last_error().decode())

    t_start = time.time()

    def progress(what):
        sys.stderr.write("[bench rank %d +%.1fs] %s\n" % (rank, time.time() - t_start, what)); sys.stderr.flush()

    e = errbuf()
    path = make_fcidump(norb, nelec)
    h = host.qcmd_create(path.encode(), symm.encode(), norb, nelec, e, 1024)
    if not h:
        raise RuntimeError(e.value.decode())
    h = ctypes.c_void_p(h)
    progress("model + MPO built")
    info = (ctypes.c_double * 32)()
    if host.qcmd_setup_site(h, site, 1, M, args.seed, local, rank, world, info, e, 1024):
        raise RuntimeError(e.value.decode())
    psi_n, sig_n = int(info[5]), int(info[6])
    progress("site problem planned (%.2f s) and uploaded" % info[13])
    stream = torch.cuda.ExternalStream(cu.qcm_stream())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        cu.qcm_sync()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(xs):
        if world == 1:
            return list(xs)
        t = torch.tensor(list(xs), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.cpu()]

    # algorithmic FLOPs of the whole problem: every rank books its share of the reference schedule exactly once
    # (plan.hpp filter_owned), so the sum over ranks is the unsharded schedule's count whatever N is
    flops, f_t, f_w, f_c, x_w, x_c = sum_over_ranks([info[0], info[1], info[2], info[3], info[24], info[25]])
    bytes_alg = info[4]

    # ---- device-resident timing (value) ------------------------------------------------------------------
    if host.qcmd_sigma_dev(h, args.warmup, e, 1024):
        raise RuntimeError(e.value.decode())
    barrier()
    progress("warm-up done")
    sampler = ClockSampler(local); sampler.start()
    l0 = cu.qcm_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    if host.qcmd_sigma_dev(h, args.steps, e, 1024):
        raise RuntimeError(e.value.decode())
    ev1.record(stream)
    barrier()
    launches = cu.qcm_launch_count() - l0
    ms_dev = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    clocks = sampler.finish()
    progress("timed region done: %.2f ms per sigma" % ms_dev)

    # ---- per-phase kernel times (same steps, CUDA events between phases) ---------------------------------
    cu.qcm_set_timing(1)
    phases = [0.0] * 6
    nph = 3
    for _ in range(nph):
        host.qcmd_sigma_dev(h, 1, e, 1024)
        cu.qcm_sync()
        ms = (ctypes.c_double * 6)(); cu.qcm_last_timing(ms)
        phases = [a + b / nph for a, b in zip(phases, ms)]
    cu.qcm_set_timing(0)

    # ---- end to end through the C ABI with pinned host buffers -------------------------------------------
    psi = torch.empty(psi_n, dtype=torch.float64).pin_memory()
    sig = torch.empty(sig_n, dtype=torch.float64).pin_memory()
    host.qcmd_get_psi(h, ctypes.c_void_p(psi.data_ptr()))
    for _ in range(args.warmup):
        if host.qcmd_sigma_host(h, ctypes.c_void_p(psi.data_ptr()), ctypes.c_void_p(sig.data_ptr()), e, 1024):
            raise RuntimeError(e.value.decode())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host.qcmd_sigma_host(h, ctypes.c_void_p(psi.data_ptr()), ctypes.c_void_p(sig.data_ptr()), e, 1024)
    barrier()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    # the same call one level up: qcm::GpuEngine::site_hamil2 on host MPSTensor objects (block flatten / unflatten included),
    # i.e. what a sweep driver that keeps its solver vectors on the host pays per sigma
    ov = ctypes.c_double()
    n_eng = min(3, args.steps)
    host.qcmd_sigma_engine(h, ctypes.byref(ov), e, 1024)
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_eng):
        if host.qcmd_sigma_engine(h, ctypes.byref(ov), e, 1024):
            raise RuntimeError(e.value.decode())
    barrier()
    ms_eng = max_over_ranks((time.perf_counter() - t0) * 1e3 / n_eng)
    progress("phases + e2e done")

    peak = ctypes.c_double(); peak_fma = ctypes.c_double()
    cu.qcm_measure_fp64_dmma_peak(ctypes.byref(peak))
    cu.qcm_measure_fp64_fma_peak(ctypes.byref(peak_fma))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)

    # dominant kernel by device time: the grouped GEMM k_gemm_ws (steps 1 and 3 are launches of the same kernel) or the
    # W application (k_wgemm_ws + k_wstream).  The GEMM is bound by the FP64 tensor pipe: achieved = FLOPs the schedule
    # hands to it (step-1 products + closing products after panel routing) / its device time; the W application is
    # bound by HBM on all but its 64-destination class: achieved = panel bytes read + written / its device time.
    t_gemm, t_w = max(phases[1] + phases[3], 1e-9), phases[2]
    if t_w > t_gemm:
        b = 8.0 * (info[21] + info[22])
        roof = {"bound": "hbm", "kernel": "k_wgemm_ws + k_wstream (W application)", "achieved": b / (t_w * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "traffic": None, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback"}
    else:
        roof = {"bound": "tensor", "kernel": "k_gemm_ws (grouped FP64 GEMM: step 1 T = L^T psi and step 3 sigma += Y R)",
                "achieved": (info[1] + info[25]) / (t_gemm * 1e-3) / 1e12, "peak": peak.value, "unit": "TFLOP/s", "traffic": None,
                "flops_per_step": info[1] + info[25], "ms_per_step": t_gemm,
                "peak_source": "FP64 DMMA chain probe measured live on this GPU (qcm_measure_fp64_dmma_peak); MEASURED_PEAKS.json has no FP64 entry"}
    # DRAM traffic of that kernel family over one sigma, from the committed ncu capture of the same workload
    # (dram__bytes_read.sum + dram__bytes_write.sum, profiles/*_traffic_*.json); null when no capture matches
    try:
        import glob
        cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic_*.json")))
        for c in reversed(cands):
            tj = json.load(open(c))
            if tj.get("workload") == args.config and world == 1:
                fam = "k_gemm_ws" if roof["bound"] == "tensor" else None
                d = tj["dram_bytes_per_sigma"]
                roof["traffic"] = d["k_gemm_ws"] if fam else d.get("k_wgemm_ws", 0) + d.get("k_wstream", 0)
                roof["traffic_unit"] = "bytes per sigma over all launches of the kernel (ncu, %s)" % os.path.basename(c)
                break
    except Exception:
        pass
    roof["frac"] = roof["achieved"] / roof["peak"] if roof["peak"] else None
    roof["phase_ms"] = {"reshape": phases[0], "step1_gemm": phases[1], "w_apply": phases[2], "step3_gemm": phases[3], "allreduce": phases[4]}

    line = {"metric": "sigma_vector_fp64_tflops", "value": flops / (ms_dev * 1e-3) / 1e12, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": args.config, "site": site, "twosite": True, "M": M, "symmetry": symm},
            "details": {"parallelism": "mpo-bond-sharded x%d" % world if slices <= 1 else "one GPU, %d time-sliced shards of the MPO bond graph" % slices,
                       "l2": "inputs (boundaries %.2f GB + workspaces %.2f GB) exceed L2" % ((info[7] + info[8]) * 8 / 1e9, info[12] / 1e9),
                       "mpo": "%dx%d nnz %d" % (info[16], info[17], info[18]), "sectors": int(info[14]), "largest_sector": int(info[15]),
                       "flops_per_step": flops, "flops_split": {"step1": f_t, "w_apply": f_w, "step3": f_c},
                       "executed_flops": {"step1": f_t, "w_apply": x_w, "step3": x_c},
                       "algorithmic_bytes": bytes_alg, "plan_seconds": info[13],
                       "w_apply_bytes": 8.0 * (info[21] + info[22]), "w_groups": int(info[23])},
            "fp64_peak_tflops": peak.value, "fp64_fma_peak_tflops": peak_fma.value, "frac_of_fp64_peak": flops / (ms_dev * 1e-3) / 1e12 / (peak.value * world) if peak.value else None,
            "roofline": roof, "clocks": clocks,
            "e2e": {"value": flops / (ms_e2e * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": psi_n * 8, "d2h_bytes_per_step": sig_n * 8,
                    "through": "qcm_site_hamil2 (C ABI, pinned host psi / sigma)",
                    "engine_mirror": {"value": flops / (ms_eng * 1e-3) / 1e12, "ms_per_step": ms_eng, "through": "qcm::GpuEngine::site_hamil2 (host MPSTensor in, host MPSTensor out: block flatten / unflatten included)"}},
            "gpu_launches": int(launches)}

    # ---- CPU baseline + full-size parity on the same instance (rank 0, N=1 only) -------------------------
    if rank == 0 and not args.no_cpu_baseline:
        olib = ctypes.CDLL(build.build_oracle())
        olib.orc_create.restype = ctypes.c_void_p
        omp_before = olib.orc_threads()          # the oracle shares this process's OpenMP runtime: give the cores back afterwards
        olib.orc_set_threads(len(os.sched_getaffinity(0)))
        oh = olib.orc_create(path.encode(), symm.encode(), norb, nelec, e, 1024)
        if not oh:
            raise RuntimeError(e.value.decode())
        oh = ctypes.c_void_p(oh)
        pe = ctypes.c_double()
        if olib.orc_setup_site(oh, site, 1, M, args.seed, ctypes.byref(pe), e, 1024):
            raise RuntimeError(e.value.decode())
        sec, ov, se = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        if olib.orc_sigma(oh, 1, ctypes.byref(sec), ctypes.byref(ov), ctypes.byref(se), e, 1024):
            raise RuntimeError(e.value.decode())
        line["cpu_baseline"] = {"value": flops / sec.value / 1e12, "unit": "TFLOP/s", "cores": olib.orc_threads(), "kind": "port",
                                "sample": "1 full sigma evaluation of the same instance (%.1f s)" % sec.value}
        if int(se.value) == sig_n:
            ref = torch.empty(sig_n, dtype=torch.float64)
            olib.orc_get_sigma(oh, ctypes.c_void_p(ref.data_ptr()))
            line["parity_rel_err_vs_oracle"] = float((sig - ref).norm() / ref.norm())
        else:
            line["parity_rel_err_vs_oracle"] = "structure mismatch: %d vs %d elements" % (int(se.value), sig_n)
        olib.orc_destroy(oh)
        olib.orc_set_threads(omp_before)
    progress("cpu baseline / parity done")
    barrier()   # the other ranks' host threads stay idle while rank 0 times the CPU oracle
    # ---- sweep level, the benchmarked configuration itself (every N): one two-site DMRG sweep from a random MPS on the
    # synthetic sector lists (total bond dimension M at every bond), boundaries resident in HBM, engine calls collective
    host.qcmd_release_site(h)
    if not args.no_config_sweep:
        try:
            line["config_sweep"] = config_sweep(host, h, args, M, local, rank, world, max_over_ranks, sum_over_ranks, peak.value)
        except Exception as ex:
            line["config_sweep"] = {"error": str(ex)}
        progress("configuration sweep done")
    # ---- sweep level (N=1): single-site DMRG sweeps of BASELINE config 1's system (8e/8o SU2U1, M=256) through the
    # engine, boundaries resident in HBM, against the same driver on the CPU oracle: per-micro-iteration energies
    if world == 1 and not args.no_sweep:
        try:
            line["sweep"] = sweep_measurement(host, build, local, not args.no_cpu_baseline)
        except Exception as ex:     # the sigma line above stays valid
            line["sweep"] = {"error": str(ex)}
    if rank == 0:
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    host.qcmd_destroy(h)
    if world > 1:
        cu.qcm_comm_destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
