// Host driver (C++ above the C ABI) exposed to the measurement harness (bench.py, smoke) through ctypes.
// It builds the Hamiltonian MPO from a FCIDUMP, fabricates a mid-chain site problem (qcm/scenarios.hpp),
// keeps its boundaries resident in HBM through qcm::GpuEngine and runs the hot path exactly the way a sweep
// driver does: plan once per site, then one qcm_site_hamil2 per eigensolver iteration
// (SweepBasedEnergyMinimization.h:85-100 -> jacobi.h:397 -> ietl::mult -> Engine::site_hamil2).
// Nothing here touches oracle/.
#include "qcm/engine_gpu.hpp"
#include "qcm/scenarios.hpp"
#include "qcm/sweep.hpp"
#include "qcm/twosite.hpp"
#include <cstdio>
#include <cstring>

using namespace qcm;

namespace {
struct Driver
{
    Problem P;
    SyntheticSite S;
    std::unique_ptr<GpuEngine> eng;
    std::shared_ptr<DeviceBoundary> dl, dr;
    std::shared_ptr<CompiledPlan> plan;
    qcm_array_t d_psi = nullptr, d_sigma = nullptr;
    std::vector<double> psi_flat;
    double plan_seconds = 0;
    ~Driver() { if (d_psi) qcm_array_free(d_psi); if (d_sigma) qcm_array_free(d_sigma); }
};
void set_err(char* err, int errlen, std::string const& s) { if (err && errlen > 0) snprintf(err, errlen, "%s", s.c_str()); }
}

extern "C" void* qcmd_create(const char* fcidump, const char* symm, int L, int nelec, char* err, int errlen)
{
    try {
        std::unique_ptr<Driver> D(new Driver());
        Problem& P = D->P;
        P.params.symm = symm_from_string(symm);
        P.params.integrals = read_fcidump(fcidump);
        P.params.L = L; P.params.site_types.assign(L, 0);
        P.params.nelec = nelec; P.params.spin = 0; P.params.nup = nelec / 2; P.params.ndown = nelec - nelec / 2;
        P.build_model();
        P.build_mpo();
        return D.release();
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return nullptr; }
}
extern "C" void qcmd_destroy(void* h) { delete static_cast<Driver*>(h); }
// drops the synthetic site problem (boundaries, plan, solver vectors) of qcmd_setup_site; the MPO and the model stay
extern "C" void qcmd_release_site(void* h)
{
    Driver* D = static_cast<Driver*>(h);
    D->plan.reset(); D->dl.reset(); D->dr.reset(); D->eng.reset();
    if (D->d_psi) { qcm_array_free(D->d_psi); D->d_psi = nullptr; }
    if (D->d_sigma) { qcm_array_free(D->d_sigma); D->d_sigma = nullptr; }
    D->S = SyntheticSite(); D->psi_flat = std::vector<double>();
}

extern "C" int qcmd_mpo_dims(void* h, int* dims, int* pairs)
{
    Driver* D = static_cast<Driver*>(h);
    for (size_t p = 0; p < D->P.mpo.size(); ++p) { dims[p] = (int)D->P.mpo[p].col_dim(); pairs[p] = (int)D->P.mpo.herm_pairs[p]; }
    return 0;
}

// info: [0] flops [1] flops_t [2] flops_w [3] flops_close [4] algorithmic bytes [5] psi elems [6] sigma elems
//       [7] left elems [8] right elems [9] gemm tasks [10] axpy tasks [11] waves [12] workspace bytes
//       [13] plan seconds [14] sectors on the left bond [15] largest sector [16] mpo rows [17] mpo cols [18] mpo nnz
//       [19] setup seconds [20] kernel launches per sigma [21] W panel elements read [22] written [23] W groups
//       [24] executed W flops [25] executed closing flops [26] panel elements consumed directly from T [27] formed by the W kernels
//       [28] panel elements skipped (no closing product)
extern "C" int qcmd_setup_site(void* h, int site, int twosite, int M, unsigned seed, int device, int rank, int world, double* info, char* err, int errlen)
{
    try {
        Driver* D = static_cast<Driver*>(h);
        D->S = make_synthetic_site(D->P, site, twosite != 0, (size_t)M, seed);
        D->eng.reset(new GpuEngine(D->P.symm(), device, rank, world));
        D->dl = D->eng->mirror(D->S.left);
        D->dr = D->eng->mirror(D->S.right);
        auto t0 = std::chrono::steady_clock::now();
        D->plan = D->eng->sigma_plan(D->S.psi, D->dl, D->dr, *D->S.mpo, true);
        D->plan_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        D->S.psi.make_left_paired();
        D->psi_flat = GpuEngine::flatten(D->S.psi.data(), D->plan->ket_elems);
        if (D->d_psi) { qcm_array_free(D->d_psi); D->d_psi = nullptr; }
        if (D->d_sigma) { qcm_array_free(D->d_sigma); D->d_sigma = nullptr; }
        qcm_check(qcm_array_alloc(D->plan->ket_elems, &D->d_psi), "alloc psi");
        qcm_check(qcm_array_alloc(D->plan->out_elems, &D->d_sigma), "alloc sigma");
        qcm_check(qcm_array_upload(D->d_psi, 0, D->psi_flat.data(), D->plan->ket_elems), "upload psi");
        int64_t nl = 0, wsb = 0;
        qcm_plan_stats(D->plan->handle, nullptr, nullptr, &nl, &wsb);
        CompiledPlan const& cp = *D->plan;
        size_t mx = 0;
        for (auto const& e : D->S.psi.row_dim()) mx = std::max(mx, e.second);
        double v[29] = {cp.flops, cp.flops_t, cp.flops_w, cp.flops_close, (double)cp.bytes, (double)cp.ket_elems, (double)cp.out_elems,
                        (double)D->dl->layout.total, (double)D->dr->layout.total, (double)cp.n_gemm_tasks, (double)cp.n_axpy_tasks, (double)cp.n_waves,
                        (double)wsb, D->plan_seconds, (double)D->S.psi.row_dim().size(), (double)mx, (double)D->S.mpo->row_dim(),
                        (double)D->S.mpo->col_dim(), (double)D->S.mpo->nnz(), D->S.setup_seconds, (double)nl,
                        (double)cp.w_elems_read, (double)cp.w_elems_written, (double)cp.w_groups,
                        cp.exec_w, cp.exec_close, (double)cp.direct_panel_elems, (double)cp.w_panel_elems, (double)cp.skipped_panel_elems};
        for (int i = 0; i < 29; ++i) info[i] = v[i];
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// copy of the site tensor (left-paired blocks back to back) for callers that own the host buffers
extern "C" int qcmd_get_psi(void* h, double* psi)
{
    Driver* D = static_cast<Driver*>(h);
    std::memcpy(psi, D->psi_flat.data(), D->psi_flat.size() * sizeof(double));
    return 0;
}

// one sigma through the C ABI with HOST buffers (H2D of psi and D2H of sigma inside the call)
extern "C" int qcmd_sigma_host(void* h, const double* psi, double* sigma, char* err, int errlen)
{
    Driver* D = static_cast<Driver*>(h);
    try { D->eng->run_sigma_host(*D->plan, *D->dl, *D->dr, psi, sigma); } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
    return 0;
}

// n sigma evaluations on device-resident vectors; the caller brackets the call with its own timing
extern "C" int qcmd_sigma_dev(void* h, int n, char* err, int errlen)
{
    Driver* D = static_cast<Driver*>(h);
    try { for (int i = 0; i < n; ++i) D->eng->run_sigma_dev(*D->plan, *D->dl, *D->dr, D->d_psi, D->d_sigma); } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
    return 0;
}
extern "C" int qcmd_get_sigma_dev(void* h, double* sigma)
{
    Driver* D = static_cast<Driver*>(h);
    return qcm_array_download(D->d_sigma, 0, sigma, D->plan->out_elems);
}

// the drop-in call itself: Engine::site_hamil2 on host MPSTensor objects (includes block flatten/unflatten)
extern "C" int qcmd_sigma_engine(void* h, double* overlap, char* err, int errlen)
{
    try {
        Driver* D = static_cast<Driver*>(h);
        MPSTensor s = D->eng->site_hamil2(D->S.psi, D->S.left, D->S.right, *D->S.mpo);
        if (overlap) *overlap = s.scalar_overlap(D->S.psi);
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// one boundary propagation with the site tensor as bra and ket (single-site problems only): returns device ms via wall clock
extern "C" int qcmd_boundary_step(void* h, int direction, double* seconds, double* out_elems, char* err, int errlen)
{
    try {
        Driver* D = static_cast<Driver*>(h);
        if (D->S.twosite) throw std::runtime_error("boundary step needs a single-site problem");
        auto t0 = std::chrono::steady_clock::now();
        Boundary b = direction == 0 ? D->eng->overlap_mpo_left_step(D->S.psi, D->S.psi, D->S.left, *D->S.mpo)
                                    : D->eng->overlap_mpo_right_step(D->S.psi, D->S.psi, D->S.right, *D->S.mpo);
        qcm_sync();
        *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        *out_elems = (double)D->eng->last_plan()->out_elems;
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// schedule-derived algorithmic FLOPs of one sigma for a synthetic site (host only; used by the CPU arm of bench.py)
extern "C" double qcmd_plan_flops(void* h, int site, int twosite, int M, unsigned seed, char* err, int errlen)
{
    try {
        Driver* D = static_cast<Driver*>(h);
        SyntheticSite S = make_synthetic_site(D->P, site, twosite != 0, (size_t)M, seed);
        S.psi.make_left_paired();
        std::vector<DualIndex> lb(S.left.aux_dim()), rb(S.right.aux_dim());
        for (size_t k = 0; k < lb.size(); ++k) lb[k] = S.left[k].basis();
        for (size_t k = 0; k < rb.size(); ++k) rb[k] = S.right[k].basis();
        plan::BoundaryLayout ll, rl; ll.assign(lb); rl.assign(rb);
        plan::Planner pl(D->P.symm(), *S.mpo, true);
        plan::Plan pp = pl.plan_sigma(GpuEngine::desc_of(S.psi), ll, rl);
        return pp.flops();
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return -1.; }
}

// Single-site DMRG sweeps on the B200 engine (qcm/sweep.hpp: the reference's ss_optimize loop and Jacobi-Davidson above
// Engine::site_hamil2 / overlap_mpo_*_step; boundaries stay in HBM between sites).
// energies: theta + core energy per micro-iteration; info: [0] sigma evaluations [1] seconds of all sweeps [2] last energy
extern "C" int qcmd_ss_sweeps(void* h, int Mmax, int nsweeps, unsigned seed, int device, double* energies, int n_max, int* n_out, double* info, char* err, int errlen)
{
    try {
        Driver* D = static_cast<Driver*>(h);
        D->P.init_mps((size_t)Mmax, true, 0., seed);
        GpuEngine eng(D->P.symm(), device, 0, 1);
        eng.set_cache_capacity(4 * D->P.mpo.size() + 8);
        sweep::SweepLog log = sweep::ss_sweeps(eng, D->P.mpo, D->P.mps, nsweeps);
        int n = (int)std::min<size_t>(log.energies.size(), (size_t)n_max);
        for (int i = 0; i < n; ++i) energies[i] = log.energies[i];
        *n_out = n;
        double secs = 0; for (double s : log.sweep_seconds) secs += s;
        info[0] = (double)log.total_sigma; info[1] = secs; info[2] = log.energies.back();
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

extern "C" int qcmd_ts_sweeps_ranked(void* h, int M0, int Mmax, int nsweeps, unsigned seed, int device, int rank, int world, double* energies, int n_max,
                                     int* n_out, double* info, char* err, int errlen);
extern "C" int qcmd_ts_sweeps(void* h, int M0, int Mmax, int nsweeps, unsigned seed, int device, double* energies, int n_max, int* n_out, double* info,
                              char* err, int errlen)
{
    return qcmd_ts_sweeps_ranked(h, M0, Mmax, nsweeps, seed, device, 0, 1, energies, n_max, n_out, info, err, errlen);
}
// Two-site DMRG sweeps on the B200 engine (qcm/twosite.hpp: ts_optimize loop, TwoSiteTensor, SVD truncation to Mmax) from
// a random MPS of bond dimension M0.  info: [0] sigma evaluations [1] seconds of all sweeps [2] last energy [3] largest bond dimension
extern "C" int qcmd_ts_sweeps_ranked(void* h, int M0, int Mmax, int nsweeps, unsigned seed, int device, int rank, int world, double* energies, int n_max,
                                     int* n_out, double* info, char* err, int errlen)
{
    try {
        Driver* D = static_cast<Driver*>(h);
        D->P.init_mps((size_t)M0, true, 0., seed);
        // world > 1: every rank runs the same (deterministic) host driver; the engine shards each contraction and the
        // library's allreduce hands every rank the complete sigma / boundary (qcm_comm_init must have been called)
        GpuEngine eng(D->P.symm(), device, rank, world);
        eng.set_cache_capacity(4 * D->P.mpo.size() + 8);
        ts::TsParams prm; prm.Mmax = (size_t)Mmax;
        std::vector<size_t> dims;
        sweep::SweepLog log = ts::ts_sweeps(D->P.symm(), eng, D->P.mpo, [&](int p) -> MPOTensor const& { return D->P.twosite_mpo(p); }, D->P.mps, nsweeps, prm, &dims);
        int n = (int)std::min<size_t>(log.energies.size(), (size_t)n_max);
        for (int i = 0; i < n; ++i) energies[i] = log.energies[i];
        *n_out = n;
        double secs = 0; for (double s : log.sweep_seconds) secs += s;
        info[0] = (double)log.total_sigma; info[1] = secs; info[2] = log.energies.back();
        info[3] = dims.empty() ? 0. : (double)*std::max_element(dims.begin(), dims.end());
        if (getenv("QCM_DEBUG"))
            fprintf(stderr, "two-site sweeps, host seconds: tensor %.2f | two-site MPO %.2f | eigensolver %.2f | split %.2f | boundary step %.2f ;  engine: planning %.2f | "
                            "plan upload %.2f | sigma calls %.2f | boundary calls %.2f | flatten %.2f | sigma-plan cache hits %zu misses %zu\n", log.phase_seconds[0], log.phase_seconds[1], log.phase_seconds[2],
                    log.phase_seconds[3], log.phase_seconds[4], eng.seconds[0], eng.seconds[1], eng.seconds[2], eng.seconds[3], eng.seconds[4], eng.cache_hits, eng.cache_misses);
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// BASELINE.json's sweep-level metric at the large configurations: two-site DMRG sweeps (ts_optimize.hpp:60-270, svd
// truncation to M, Jacobi-Davidson with the reference's defaults) started from a random MPS on the synthetic sector
// lists (total bond dimension M at every bond, scenarios.hpp make_synthetic_mps), boundaries resident in HBM, stale
// boundaries dropped.  world > 1: every rank drives the same deterministic host loop, the engine calls are collective.
// max_micro > 0 bounds the LAST sweep to that many micro-iterations (for bounded measurement runs).
// info: [0] sigma evaluations [1] seconds of all sweeps [2] last energy [3] largest bond dimension kept
//       [4..8] driver seconds: two-site tensor, two-site MPO, eigensolver, split, boundary step
//       [9..13] engine seconds: planning, plan upload, sigma calls (device solver incl.), boundary calls, flatten/unflatten
//       [14] plan-cache hits [15] misses [16] sum of sigma FLOPs (this rank's share) [17] sum of boundary-step FLOPs
//       [18] seconds before the first sweep (canonisation + initial right boundaries) [19] micro-iterations
//       [20] seconds of the last sweep [21] boundaries evicted to pinned host memory [22] prefetched back (QCM_SPILL)
extern "C" int qcmd_ts_sweeps_synth(void* h, int M, int nsweeps, unsigned seed, int device, int rank, int world, int max_micro, double budget_seconds,
                                    double* energies, int n_max, int* n_out, double* info, char* err, int errlen)
{
    try {
        Driver* D = static_cast<Driver*>(h);
        auto t0 = std::chrono::steady_clock::now();
        // one process per GPU: the host BLAS gets the same share of the cores as OpenMP (its default is all of them, per process)
        sweep::scipy_openblas_set_num_threads(omp_get_max_threads());
        D->P.mps = make_synthetic_mps(D->P, (size_t)M, seed);
        GpuEngine eng(D->P.symm(), device, rank, world);
        eng.set_cache_capacity(4);
        ts::TsParams prm; prm.Mmax = (size_t)M; prm.drop_stale = true; prm.max_micro_iterations = max_micro;
        prm.spill = getenv("QCM_SPILL") != nullptr;      // boundaries the sweep has left behind move to pinned host memory
        // wall-clock budget: checked at site boundaries; with several ranks the flags are summed so that all ranks stop together
        qcm_array_t flag = nullptr;
        if (budget_seconds > 0) {
            if (world > 1) qcm_check(qcm_array_alloc(1, &flag), "qcm_array_alloc");
            prm.should_stop = [&]() {
                double over = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > budget_seconds ? 1. : 0.;
                if (world > 1) {
                    qcm_check(qcm_array_upload(flag, 0, &over, 1), "qcm_array_upload");
                    qcm_check(qcm_comm_allreduce(flag, 1), "qcm_comm_allreduce");
                    qcm_check(qcm_array_download(flag, 0, &over, 1), "qcm_array_download");
                }
                return over > 0.;
            };
        }
        std::vector<size_t> dims;
        double init_s = 0;
        sweep::SweepLog log = ts::ts_sweeps(D->P.symm(), eng, D->P.mpo, [&](int p) -> MPOTensor const& { return D->P.twosite_mpo(p); }, D->P.mps, nsweeps, prm, &dims, &init_s);
        if (flag) qcm_array_free(flag);
        int n = (int)std::min<size_t>(log.energies.size(), (size_t)n_max);
        for (int i = 0; i < n; ++i) energies[i] = log.energies[i];
        *n_out = n;
        double secs = 0; for (double s : log.sweep_seconds) secs += s;
        info[0] = (double)log.total_sigma; info[1] = secs; info[2] = log.energies.empty() ? 0. : log.energies.back();
        info[3] = dims.empty() ? 0. : (double)*std::max_element(dims.begin(), dims.end());
        for (int i = 0; i < 5; ++i) { info[4 + i] = log.phase_seconds[i]; info[9 + i] = eng.seconds[i]; }
        info[14] = (double)eng.cache_hits; info[15] = (double)eng.cache_misses; info[16] = eng.sigma_flops; info[17] = eng.boundary_flops;
        info[18] = init_s; info[19] = (double)log.energies.size(); info[20] = log.sweep_seconds.empty() ? 0. : log.sweep_seconds.back();
        info[21] = (double)eng.n_evicted; info[22] = (double)eng.n_prefetched;
        if (getenv("QCM_DEBUG")) {
            double* ss = ts::split_seconds();
            fprintf(stderr, "[rank %d] split seconds: block SVDs %.2f | combination across ranks %.2f | truncation %.2f | whole split incl. the former (reshapes, recoupling) %.2f | "
                            "normalisation + shift %.2f ; device solver host syncs %zu ; boundary steps: allocation %.2f flatten %.2f\n", rank, ss[0], ss[1], ss[2], ss[3], ss[4], eng.host_syncs,
                    eng.detail_seconds[0], eng.detail_seconds[1]);
        }
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}
