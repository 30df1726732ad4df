// TEST INFRASTRUCTURE -- C entry points (ctypes) that run the same scenario through the CPU oracle and through
// an engine under test and report structure equality and relative differences.
//   engine 0: plan interpreter (CPU; checks the schedule builder, runs without a GPU)
//   engine 1: qcm::GpuEngine (B200; goes through the C ABI of include/qcm_b200.h)
// Scenarios mirror the reference's own hot-path tests: boundary chains and the sigma-energy identity of
// dmrg/tests/test_mps_mpo_ops/test_siteproblem.cpp:38-95, BoundaryPropagatorElectronic.cpp:39-66.
#include "oracle_engine.hpp"
#include "qcm/scenarios.hpp"
#include "qcm/sweep.hpp"
#include "qcm/twosite.hpp"
#include "qcm/overlap.hpp"
#include "plan_interp.hpp"
#include "flatten_desc.hpp"
#ifdef QCMT_WITH_GPU
#include "qcm/engine_gpu.hpp"
#endif
#include <cstdio>
#include <cstring>

using namespace qcm;

static void set_err(char* err, int errlen, std::string const& s) { if (err && errlen > 0) { snprintf(err, errlen, "%s", s.c_str()); } }

static double rel_diff(DiffReport const& r) { return r.ref_norm > 0 ? std::sqrt(r.diff_norm / r.ref_norm) : std::sqrt(r.diff_norm); }

static Problem make_problem(const char* fcidump, const char* symm, int L, int nelec)
{
    Problem P;
    P.params.symm = symm_from_string(symm);
    P.params.integrals = read_fcidump(fcidump);
    P.params.L = L; P.params.site_types.assign(L, 0);
    P.params.nelec = nelec; P.params.spin = 0; P.params.nup = nelec / 2; P.params.ndown = nelec - nelec / 2;
    P.build_model();
    P.build_mpo();
    return P;
}

extern "C" int qcmt_gpu_available()
{
#ifdef QCMT_WITH_GPU
    int n = 0;
    if (qcm_device_count(&n) != 0) return 0;
    return n;
#else
    return 0;
#endif
}

// out[0] n boundaries compared   out[1] all boundary structures equal   out[2] max rel diff over boundaries
// out[3] n single-site sigma     out[4] structures equal                out[5] max rel diff
// out[6] n two-site sigma        out[7] structures equal                out[8] max rel diff
// out[9] max |<psi|sigma> - chain energy| (engine)   out[10] chain energy engine   out[11] chain energy oracle
// out[12] total flops of the last sigma plan (two-site)  out[13] max rel diff oracle lbtm vs rbtm (SU2 only)
extern "C" int qcmt_chain_parity(const char* fcidump, const char* symm, int L, int nelec, int Mmax, unsigned seed, int engine_kind, int world,
                                 long long budget, double* out, int nout, char* err, int errlen)
{
    try {
        for (int i = 0; i < nout; ++i) out[i] = 0;
        Problem P = make_problem(fcidump, symm, L, nelec);
        P.init_mps((size_t)Mmax, true, 0., seed);
        oracle::OracleEngine orc(P.params.symm);
        std::unique_ptr<EngineIface> eng;
        qcmtest::InterpEngine* interp = nullptr;
        if (engine_kind == 0) { interp = new qcmtest::InterpEngine(P.params.symm, world, budget); eng.reset(interp); }
        else {
#ifdef QCMT_WITH_GPU
            eng.reset(new GpuEngine(P.params.symm, 0, 0, 1, budget));
#else
            throw std::runtime_error("harness built without GPU support");
#endif
        }
        // boundaries through both engines
        Problem Po = P;
        Po.build_boundaries(orc);
        P.build_boundaries(*eng);
#ifdef QCMT_WITH_GPU
        if (engine_kind == 1) {
            GpuEngine* g = static_cast<GpuEngine*>(eng.get());
            for (auto& b : P.left) g->download(b);
            for (auto& b : P.right) g->download(b);
        }
#endif
        double bmax = 0; int bstruct = 1, nb = 0;
        for (int p = 0; p <= L; ++p) {
            DiffReport a = compare(P.left[p], Po.left[p]), b = compare(P.right[p], Po.right[p]);
            bmax = std::max(bmax, std::max(rel_diff(a), rel_diff(b)));
            bstruct &= a.structure_equal & b.structure_equal;
            nb += 2;
        }
        out[0] = nb; out[1] = bstruct; out[2] = bmax;
        out[10] = P.left[L][0].trace(); out[11] = Po.left[L][0].trace();
        // single-site sigma (engine uses its own boundaries)
        double smax = 0, emax = 0; int sstruct = 1, ns = 0;
        for (int p = 0; p < L; ++p) {
            MPSTensor so = orc.site_hamil2(Po.mps[p], Po.left[p], Po.right[p + 1], Po.mpo[p]);
            MPSTensor se = eng->site_hamil2(P.mps[p], P.left[p], P.right[p + 1], P.mpo[p]);
            DiffReport d = compare(se.data(), so.data());
            smax = std::max(smax, rel_diff(d)); sstruct &= d.structure_equal; ns++;
            emax = std::max(emax, std::abs(se.scalar_overlap(P.mps[p]) - out[10]));
        }
        out[3] = ns; out[4] = sstruct; out[5] = smax; out[9] = emax;
        // two-site sigma on every bond with a random two-site tensor
        double tmax = 0, lrmax = 0; int tstruct = 1, nt = 0;
        UniformGen gen(seed + 7);
        for (int p = 0; p + 1 < L; ++p) {
            MPOTensor const& ts = P.twosite_mpo(p);
            MPSTensor x = make_twosite_tensor(P.phys(p), P.phys(p + 1), P.mps[p].row_dim(), P.mps[p + 1].col_dim(), [&]() { return gen() - 0.5; });
            MPSTensor so = orc.site_hamil2(x, Po.left[p], Po.right[p + 2], ts);
            MPSTensor se = eng->site_hamil2(x, P.left[p], P.right[p + 2], ts);
            DiffReport d = compare(se.data(), so.data());
            tmax = std::max(tmax, rel_diff(d)); tstruct &= d.structure_equal; nt++;
            if (is_su2(P.params.symm)) {
                MPSTensor a = orc.site_hamil_lbtm(x, x, Po.left[p], Po.right[p + 2], ts, true);
                MPSTensor b = orc.site_hamil_rbtm(x, x, Po.left[p], Po.right[p + 2], ts, true);
                lrmax = std::max(lrmax, rel_diff(compare(a.data(), b.data())));
            }
            if (interp) out[12] = interp->last_flops;
#ifdef QCMT_WITH_GPU
            if (engine_kind == 1) out[12] = static_cast<GpuEngine*>(eng.get())->last_plan()->flops;
#endif
        }
        out[6] = nt; out[7] = tstruct; out[8] = tmax; out[13] = lrmax;
        return 0;
    } catch (std::exception const& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}

// MPO bond dimensions / Hermitian pairs ("MPO Bond p: dim/pairs" lines of the reference's stdout) and the
// exact two-site ground-state energy of tiny systems (dense effective Hamiltonian on complete bases).
extern "C" int qcmt_mpo_dims(const char* fcidump, const char* symm, int L, int nelec, int* dims, int* pairs, double* core, char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        for (int p = 0; p < L; ++p) { dims[p] = (int)P.mpo[p].col_dim(); pairs[p] = (int)P.mpo.herm_pairs[p]; }
        *core = P.mpo.core_energy;
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

extern "C" int qcmt_exact_energy(const char* fcidump, const char* symm, int L, int nelec, int engine_kind, double* energy, double* asym, char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        std::unique_ptr<EngineIface> eng;
        if (engine_kind == 0) eng.reset(new qcmtest::InterpEngine(P.params.symm));
        else if (engine_kind == 2) eng.reset(new oracle::OracleEngine(P.params.symm));
        else {
#ifdef QCMT_WITH_GPU
            eng.reset(new GpuEngine(P.params.symm));
#else
            throw std::runtime_error("harness built without GPU support");
#endif
        }
        std::vector<Index> allowed = allowed_sectors(P.params.symm, P.site_types(), P.model->phys_indices, P.model->total_charge, 1000);
        MPS ex;
        for (int p = 0; p < L; ++p) ex.push_back(MPSTensor(P.phys(p), allowed[p], allowed[p + 1], []() { return 1.0; }));
        P.mps = ex;
        int c = L / 2 - 1;
        P.build_boundaries(*eng, c, c + 2);
        MPOTensor const& ts = P.twosite_mpo(c);
        MPSTensor templ = make_twosite_tensor(P.phys(c), P.phys(c + 1), allowed[c], allowed[c + 2], []() { return 1.0; });
        std::vector<double> w = dense_heff_spectrum(*eng, templ, P.left[c], P.right[c + 2], ts, asym);
        *energy = w[0] + P.mpo.core_energy;
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// "The hamiltonian will contain N terms" (models/chem/*/model.hpp) -- a golden of the term makers
extern "C" int qcmt_num_terms(const char* fcidump, const char* symm, int L, int nelec, int* nterms, char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        *nterms = (int)P.model->terms.size();
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// Racah-sum restatement of gsl_sf_coupling_9j (arguments are 2*j), pinned by dmrg/tests/test_wigner.cpp:21-46
extern "C" double qcmt_wigner9j(int a, int b, int c, int d, int e, int f, int g, int h, int i) { return su2::wigner9j(a, b, c, d, e, f, g, h, i); }
extern "C" double qcmt_wigner6j(int a, int b, int c, int d, int e, int f) { return su2::wigner6j(a, b, c, d, e, f); }

// Synthetic mid-chain site problem (the generator behind bench.py) through the oracle and an engine under test.
// out[0] sigma structure equal  out[1] sigma rel diff  out[2] n sigma elements  out[3] <psi|sigma> oracle
// single-site problems additionally: out[4]/out[5] left-step structure/rel diff, out[6]/out[7] right step
// out[8] plan flops  out[9] waves
extern "C" int qcmt_synth_parity(const char* fcidump, const char* symm, int L, int nelec, int site, int twosite, int M, unsigned seed, int engine_kind,
                                 int world, long long budget, double* out, int nout, char* err, int errlen)
{
    try {
        for (int i = 0; i < nout; ++i) out[i] = 0;
        Problem P = make_problem(fcidump, symm, L, nelec);
        SyntheticSite S = make_synthetic_site(P, site, twosite != 0, (size_t)M, seed);
        oracle::OracleEngine orc(P.params.symm);
        std::unique_ptr<EngineIface> eng;
        qcmtest::InterpEngine* interp = nullptr;
        if (engine_kind == 0) { interp = new qcmtest::InterpEngine(P.params.symm, world, budget); eng.reset(interp); }
        else {
#ifdef QCMT_WITH_GPU
            eng.reset(new GpuEngine(P.params.symm, 0, 0, 1, budget));
#else
            throw std::runtime_error("harness built without GPU support");
#endif
        }
        MPSTensor so = orc.site_hamil2(S.psi, S.left, S.right, *S.mpo);
        MPSTensor se = eng->site_hamil2(S.psi, S.left, S.right, *S.mpo);
        DiffReport d = compare(se.data(), so.data());
        out[0] = d.structure_equal; out[1] = rel_diff(d); out[2] = (double)so.data().num_elements(); out[3] = so.scalar_overlap(S.psi);
        if (interp) { out[8] = interp->last_flops; out[9] = (double)interp->last_waves; out[10] = (double)interp->last_exchange_elems; out[11] = interp->last_exec_close; }
#ifdef QCMT_WITH_GPU
        if (engine_kind == 1) { out[8] = static_cast<GpuEngine*>(eng.get())->last_plan()->flops; out[9] = (double)static_cast<GpuEngine*>(eng.get())->last_plan()->n_waves; }
#endif
        if (!twosite) {
            Boundary lo = orc.overlap_mpo_left_step(S.psi, S.psi, S.left, *S.mpo), le = eng->overlap_mpo_left_step(S.psi, S.psi, S.left, *S.mpo);
            Boundary ro = orc.overlap_mpo_right_step(S.psi, S.psi, S.right, *S.mpo), re = eng->overlap_mpo_right_step(S.psi, S.psi, S.right, *S.mpo);
#ifdef QCMT_WITH_GPU
            if (engine_kind == 1) { static_cast<GpuEngine*>(eng.get())->download(le); static_cast<GpuEngine*>(eng.get())->download(re); }
#endif
            DiffReport a = compare(le, lo), b = compare(re, ro);
            out[4] = a.structure_equal; out[5] = rel_diff(a); out[6] = b.structure_equal; out[7] = rel_diff(b);
        }
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// One rank's share of a sharded sigma (plan interpreter; what one GPU computes before the allreduce) and the
// oracle's full sigma, both as flat left-paired block buffers.  Used by the world_size-2 gloo test.
// mode 0: number of sigma elements -> *n_out;  mode 1: rank share -> buf;  mode 2: oracle -> buf
extern "C" int qcmt_rank_sigma(const char* fcidump, const char* symm, int L, int nelec, int site, int twosite, int M, unsigned seed, int rank, int world,
                               int mode, double* buf, long long* n_out, char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        SyntheticSite S = make_synthetic_site(P, site, twosite != 0, (size_t)M, seed);
        S.psi.make_left_paired();
        if (mode == 2) {
            oracle::OracleEngine orc(P.params.symm);
            MPSTensor so = orc.site_hamil2(S.psi, S.left, S.right, *S.mpo);
            so.make_left_paired();
            size_t o = 0;
            for (size_t k = 0; k < so.data().n_blocks(); ++k) { auto const& v = so.data()[k].v; std::memcpy(buf + o, v.data(), v.size() * 8); o += v.size(); }
            *n_out = (long long)o;
            return 0;
        }
        plan::BoundaryLayout ll = qcmtest::InterpEngine::layout_of(S.left), rl = qcmtest::InterpEngine::layout_of(S.right);
        plan::Planner pl(P.params.symm, *S.mpo, true, rank, world, (int64_t)1 << 28);
        if (mode == 0) pl.structure_only = true;
        plan::Plan pp = pl.plan_sigma(qcmtest::InterpEngine::desc_of(S.psi), ll, rl);
        *n_out = pp.out_tensor.total;
        if (mode == 0) return 0;
        qcmtest::Bufs B;
        B.b[plan::BUF_KET_LP] = qcmtest::InterpEngine::flat(S.psi.data()); B.b[plan::BUF_LEFT] = qcmtest::InterpEngine::flat(S.left);
        B.b[plan::BUF_RIGHT] = qcmtest::InterpEngine::flat(S.right);
        B.b[plan::BUF_OUT].assign((size_t)pp.out_tensor.total, 0.);
        qcmtest::run_plan(pp, B);
        std::memcpy(buf, B.b[plan::BUF_OUT].data(), (size_t)pp.out_tensor.total * 8);
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// The same in the phases a rank goes through when the plan has an exchange wave (world > 1): the caller plays NCCL.
//   phase 0: plan, reshapes, resident step-1 products, W pass of the exchange wave; *n_out = elements of the exchange region
//            (0: no exchange), *n_sigma = sigma elements
//   phase 1: copy the exchange region (this rank's partial sums) into buf
//   phase 2: take the reduce-scattered region from buf (only chunk `rank` is kept, the rest is poisoned)
//   phase 3: local waves + closing products of the exchange chunk; the rank's share of sigma -> buf
namespace { struct RankState { Problem P; SyntheticSite S; plan::Plan plan; qcmtest::Bufs B; }; std::unique_ptr<RankState> g_rank_state; }
extern "C" int qcmt_rank_sigma_phase(int phase, const char* fcidump, const char* symm, int L, int nelec, int site, int twosite, int M, unsigned seed, int rank,
                                     int world, double* buf, long long* n_out, long long* n_sigma, char* err, int errlen)
{
    try {
        if (phase == 0) {
            g_rank_state.reset(new RankState());
            RankState& R = *g_rank_state;
            R.P = make_problem(fcidump, symm, L, nelec);
            R.S = make_synthetic_site(R.P, site, twosite != 0, (size_t)M, seed);
            R.S.psi.make_left_paired();
            plan::BoundaryLayout ll = qcmtest::InterpEngine::layout_of(R.S.left), rl = qcmtest::InterpEngine::layout_of(R.S.right);
            plan::Planner pl(R.P.params.symm, *R.S.mpo, true, rank, world, (int64_t)1 << 28);
            R.plan = pl.plan_sigma(qcmtest::InterpEngine::desc_of(R.S.psi), ll, rl);
            R.B.b[plan::BUF_KET_LP] = qcmtest::InterpEngine::flat(R.S.psi.data()); R.B.b[plan::BUF_LEFT] = qcmtest::InterpEngine::flat(R.S.left);
            R.B.b[plan::BUF_RIGHT] = qcmtest::InterpEngine::flat(R.S.right);
            R.B.b[plan::BUF_OUT].assign((size_t)R.plan.out_tensor.total, 0.);
            qcmtest::run_pre(R.plan, R.B);
            if (qcmtest::exchange_elems(R.plan) > 0) qcmtest::run_exchange_w(R.plan, R.B);
            *n_out = qcmtest::exchange_elems(R.plan); *n_sigma = R.plan.out_tensor.total;
            return 0;
        }
        if (!g_rank_state) throw std::runtime_error("qcmt_rank_sigma_phase: phase 0 has not run");
        RankState& R = *g_rank_state;
        const long long xe = qcmtest::exchange_elems(R.plan);
        if (phase == 1) { std::memcpy(buf, R.B.b[plan::BUF_Y].data(), (size_t)xe * 8); return 0; }
        if (phase == 2) {
            const long long C = R.plan.waves[0].x_chunk;
            for (long long i = 0; i < xe; ++i) R.B.b[plan::BUF_Y][(size_t)i] = (i / C == rank) ? buf[i] : std::nan("");
            return 0;
        }
        qcmtest::run_local(R.plan, R.B);
        std::memcpy(buf, R.B.b[plan::BUF_OUT].data(), (size_t)R.plan.out_tensor.total * 8);
        g_rank_state.reset();
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// Single-site DMRG sweeps (qcm/sweep.hpp: ss_optimize loop + Jacobi-Davidson) through an engine.
//   engine_kind -1: CPU oracle, 0: plan interpreter, 1: qcm::GpuEngine
// energies[0 .. *n_out): theta + core energy of every micro-iteration; info[0] sigma evaluations, [1] seconds of all
// sweeps, [2] last energy, [3] micro-iterations per sweep
extern "C" int qcmt_ss_dmrg(const char* fcidump, const char* symm, int L, int nelec, int Mmax, int nsweeps, unsigned seed, int engine_kind,
                            double* energies, int n_max, int* n_out, double* info, char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        P.init_mps((size_t)Mmax, true, 0., seed);
        std::unique_ptr<EngineIface> eng;
        if (engine_kind < 0) eng.reset(new oracle::OracleEngine(P.params.symm));
        else if (engine_kind == 0) eng.reset(new qcmtest::InterpEngine(P.params.symm, 1, (long long)1 << 40));
        else {
#ifdef QCMT_WITH_GPU
            eng.reset(new GpuEngine(P.params.symm, 0, 0, 1));
#else
            throw std::runtime_error("harness built without GPU support");
#endif
        }
        sweep::SweepLog log = sweep::ss_sweeps(*eng, P.mpo, P.mps, nsweeps);
        int n = (int)std::min<size_t>(log.energies.size(), (size_t)n_max);
        for (int i = 0; i < n; ++i) energies[i] = log.energies[i];
        *n_out = n;
        double secs = 0; for (double s : log.sweep_seconds) secs += s;
        info[0] = (double)log.total_sigma; info[1] = secs; info[2] = log.energies.back(); info[3] = 2.0 * L;
        return 0;
    } catch (std::exception const& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}

// diagonal_hamiltonian (abelian/h_diag.hpp, non-abelian/h_diag.hpp) on every site (single-site) and every bond (two-site)
// of a random MPS: engine under test against the oracle's literal restatement.
// out[0] cases  out[1] all structures equal  out[2] max rel diff  out[3] max |diag| (non-trivial check)
extern "C" int qcmt_hdiag_parity(const char* fcidump, const char* symm, int L, int nelec, int Mmax, unsigned seed, int engine_kind, double* out, char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        P.init_mps((size_t)Mmax, true, 0., seed);
        oracle::OracleEngine orc(P.params.symm);
        P.build_boundaries(orc);
        std::unique_ptr<qcmtest::InterpEngine> interp;
#ifdef QCMT_WITH_GPU
        std::unique_ptr<GpuEngine> gpu;
        if (engine_kind == 1) gpu.reset(new GpuEngine(P.params.symm, 0, 0, 1));
#else
        if (engine_kind == 1) throw std::runtime_error("harness built without GPU support");
#endif
        if (engine_kind == 0) interp.reset(new qcmtest::InterpEngine(P.params.symm, 1, (long long)1 << 40));
        auto run = [&](Boundary const& l, Boundary const& r, MPOTensor const& w, MPSTensor const& x) {
#ifdef QCMT_WITH_GPU
            if (gpu) return gpu->diagonal_hamiltonian(l, r, w, x);
#endif
            return interp->diagonal_hamiltonian(l, r, w, x);
        };
        double mx = 0, amax = 0; int st = 1, n = 0;
        for (int p = 0; p < L; ++p) {
            block_matrix a = oracle::hdiag::diagonal_hamiltonian(P.params.symm, P.left[p], P.right[p + 1], P.mpo[p], P.mps[p]);
            block_matrix b = run(P.left[p], P.right[p + 1], P.mpo[p], P.mps[p]);
            DiffReport d = compare(b, a);
            mx = std::max(mx, rel_diff(d)); st &= d.structure_equal; ++n;
            amax = std::max(amax, std::sqrt(a.norm_square()));
        }
        for (int p = 0; p + 1 < L; ++p) {
            MPOTensor const& ts = P.twosite_mpo(p);
            MPSTensor x = make_twosite_tensor(P.phys(p), P.phys(p + 1), P.mps[p].row_dim(), P.mps[p + 1].col_dim(), []() { return 1.0; });
            block_matrix a = oracle::hdiag::diagonal_hamiltonian(P.params.symm, P.left[p], P.right[p + 2], ts, x);
            block_matrix b = run(P.left[p], P.right[p + 2], ts, x);
            DiffReport d = compare(b, a);
            mx = std::max(mx, rel_diff(d)); st &= d.structure_equal; ++n;
        }
        out[0] = n; out[1] = st; out[2] = mx; out[3] = amax;
        return 0;
    } catch (std::exception const& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}

// Two-site DMRG sweeps (qcm/twosite.hpp: ts_optimize loop, TwoSiteTensor, SVD truncation to Mmax) through an engine.
//   engine_kind -1: CPU oracle, 0: plan interpreter, 1: qcm::GpuEngine;  M0: bond dimension of the random start
// energies[0 .. *n_out): per micro-iteration; info[0] sigma evaluations, [1] seconds, [2] last energy, [3] largest bond dimension kept
extern "C" int qcmt_ts_dmrg(const char* fcidump, const char* symm, int L, int nelec, int M0, int Mmax, int nsweeps, unsigned seed, int engine_kind,
                            double* energies, int n_max, int* n_out, double* info, char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        P.init_mps((size_t)M0, true, 0., seed);
        std::unique_ptr<EngineIface> eng;
        if (engine_kind < 0) eng.reset(new oracle::OracleEngine(P.params.symm));
        else if (engine_kind == 0) eng.reset(new qcmtest::InterpEngine(P.params.symm, 1, (long long)1 << 40));
        else {
#ifdef QCMT_WITH_GPU
            eng.reset(new GpuEngine(P.params.symm, 0, 0, 1));
#else
            throw std::runtime_error("harness built without GPU support");
#endif
        }
        ts::TsParams prm; prm.Mmax = (size_t)Mmax;
        std::vector<size_t> dims;
        sweep::SweepLog log = ts::ts_sweeps(P.params.symm, *eng, P.mpo, [&](int p) -> MPOTensor const& { return P.twosite_mpo(p); }, P.mps, nsweeps, prm, &dims);
        int n = (int)std::min<size_t>(log.energies.size(), (size_t)n_max);
        for (int i = 0; i < n; ++i) energies[i] = log.energies[i];
        *n_out = n;
        double secs = 0; for (double s : log.sweep_seconds) secs += s;
        info[0] = (double)log.total_sigma; info[1] = secs; info[2] = log.energies.back();
        info[3] = dims.empty() ? 0. : (double)*std::max_element(dims.begin(), dims.end());
        return 0;
    } catch (std::exception const& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}

// Round trip of the two-site data formats on every bond of a random MPS (no engine involved):
//   T = A[p] A[p+1] (both-paired)  ->  make_mps (right-paired; SU2: spin-coupled)  ->  operator<< (SU2: uncoupled again)
//   ->  split without truncation  ->  product of the two new site tensors must equal T;  the norm is preserved by make_mps
// (the 6j recoupling is orthogonal), U has orthonormal columns.
// out[0] bonds  out[1] max rel |T' - T|  out[2] max rel | |make_mps| - |T| |  out[3] max |U^T U - 1|
extern "C" int qcmt_twosite_roundtrip(const char* fcidump, const char* symm, int L, int nelec, int Mmax, unsigned seed, double* out, char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        P.init_mps((size_t)Mmax, true, 0., seed);
        double d_prod = 0, d_norm = 0, d_orth = 0; int n = 0;
        for (int p = 0; p + 1 < L; ++p) {
            MPSTensor a = P.mps[p], b = P.mps[p + 1];
            a.make_left_paired(); b.make_right_paired();
            block_matrix T; sweep::gemm(a.data(), b.data(), T);
            ts::TwoSiteTensor tst(P.params.symm, P.mps[p], P.mps[p + 1]);
            MPSTensor twin = tst.make_mps();
            d_norm = std::max(d_norm, std::abs(twin.scalar_norm() - std::sqrt(T.norm_square())) / std::sqrt(T.norm_square()));
            tst << twin;
            MPSTensor t1, t2; ts::Truncation tr;
            tst.split_mps_l2r(100000, 0., t1, t2, tr);
            t1.make_left_paired(); t2.make_right_paired();
            block_matrix T2; sweep::gemm(t1.data(), t2.data(), T2);
            d_prod = std::max(d_prod, rel_diff(compare(T2, T)));
            block_matrix utu; { block_matrix ut; for (size_t k = 0; k < t1.data().n_blocks(); ++k) { Matrix const& m = t1.data()[k]; Matrix mt(m.cols, m.rows); for (size_t i = 0; i < m.rows; ++i) for (size_t j = 0; j < m.cols; ++j) mt(j, i) = m(i, j); ut.insert_block(mt, t1.data().basis()[k].rc, t1.data().basis()[k].lc); } sweep::gemm(ut, t1.data(), utu); }
            for (size_t k = 0; k < utu.n_blocks(); ++k) for (size_t i = 0; i < utu[k].rows; ++i) for (size_t j = 0; j < utu[k].cols; ++j) d_orth = std::max(d_orth, std::abs(utu[k](i, j) - (i == j ? 1. : 0.)));
            ++n;
        }
        out[0] = n; out[1] = d_prod; out[2] = d_norm; out[3] = d_orth;
        return 0;
    } catch (std::exception const& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}

// Sharding quality of the sigma plan (plan.hpp shard_sources): per-rank FLOPs for `world` ranks on a synthetic two-site problem.
// out[0] step-1 FLOPs summed over ranks / step-1 FLOPs of the unsharded plan (1 = nothing computed twice)
// out[1] max over ranks of executed FLOPs / (executed FLOPs of the unsharded plan / world)   (1 = perfect balance, no replication)
// out[2] algorithmic FLOPs booked over ranks / unsharded (must be exactly 1)
extern "C" int qcmt_shard_stats(const char* fcidump, const char* symm, int L, int nelec, int site, int M, unsigned seed, int world, double* out, char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        SyntheticSite S = make_synthetic_site(P, site, true, (size_t)M, seed);
        S.psi.make_left_paired();
        std::vector<DualIndex> lb(S.left.aux_dim()), rb(S.right.aux_dim());
        for (size_t k = 0; k < lb.size(); ++k) lb[k] = S.left[k].basis();
        for (size_t k = 0; k < rb.size(); ++k) rb[k] = S.right[k].basis();
        plan::BoundaryLayout ll, rl; ll.assign(lb); rl.assign(rb);
        plan::TensorDesc td{S.psi.site_dim(), S.psi.row_dim(), S.psi.col_dim(), S.psi.data().basis()};
        plan::Planner p1(P.params.symm, *S.mpo, true, 0, 1, (int64_t)1 << 40);
        plan::Plan full = p1.plan_sigma(td, ll, rl);
        double t_sum = 0, mx = 0, alg = 0, close_sum = 0;
        for (int r = 0; r < world; ++r) {
            plan::Planner pr(P.params.symm, *S.mpo, true, r, world, (int64_t)1 << 40);
            plan::Plan q = pr.plan_sigma(td, ll, rl);
            t_sum += q.flops_t; alg += q.flops(); close_sum += q.exec_close;
            mx = std::max(mx, q.flops_t + q.exec_w + q.exec_close);
        }
        out[0] = t_sum / full.flops_t;
        out[1] = mx / ((full.flops_t + full.exec_w + full.exec_close) / world);
        out[2] = alg / full.flops();
        out[3] = close_sum / full.exec_close;      // executed closing FLOPs over all ranks / unsharded plan
        return 0;
    } catch (std::exception const& e) {
        set_err(err, errlen, e.what());
        return 1;
    }
}

// The two-site split with its block SVDs divided among `world` ranks (twosite.hpp svd_truncate, EngineIface::allreduce_sum),
// emulated in one process: pass 1 collects every rank's zero-padded buffer, pass 2 hands each rank the sum, as the allreduce
// does.  out[0] = bonds checked, out[1] = 1 if every rank ends up with bit-identical factors, out[2] = largest relative
// difference of U diag(S) V to the single-rank split, out[3] = 1 if the kept bond structure equals the single-rank one.
namespace {
struct FakeCommEngine : public qcm::EngineIface
{
    int r, w; std::vector<double>* slot; std::vector<double> const* total;
    FakeCommEngine(int r_, int w_, std::vector<double>* s, std::vector<double> const* t) : r(r_), w(w_), slot(s), total(t) {}
    qcm::MPSTensor site_hamil2(qcm::MPSTensor, qcm::Boundary const&, qcm::Boundary const&, qcm::MPOTensor const&, bool) override { throw std::runtime_error("not used"); }
    qcm::Boundary overlap_mpo_left_step(qcm::MPSTensor const&, qcm::MPSTensor const&, qcm::Boundary const&, qcm::MPOTensor const&, bool) override { throw std::runtime_error("not used"); }
    qcm::Boundary overlap_mpo_right_step(qcm::MPSTensor const&, qcm::MPSTensor const&, qcm::Boundary const&, qcm::MPOTensor const&, bool) override { throw std::runtime_error("not used"); }
    int comm_rank() const override { return r; }
    int comm_world() const override { return w; }
    void allreduce_sum(double* buf, size_t n) override
    {
        if (total) { if (total->size() != n) throw std::runtime_error("ranks disagree on the buffer size"); std::copy(total->begin(), total->end(), buf); }
        else slot->assign(buf, buf + n);
    }
};
}
extern "C" int qcmt_sharded_split(const char* fcidump, const char* symm, int L, int nelec, int Mmax, unsigned seed, int world, double* out, char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        P.init_mps((size_t)Mmax, true, 0., seed);
        int n = 0; double identical = 1, same_struct = 1, worst = 0;
        for (int p = 0; p + 1 < L; ++p) {
            ts::TwoSiteTensor tst(P.params.symm, P.mps[p], P.mps[p + 1]);
            MPSTensor twin = tst.make_mps();
            tst << twin;
            const size_t keep = (size_t)std::max(1, Mmax / 2);
            MPSTensor a1, b1; ts::Truncation t1;
            { ts::TwoSiteTensor c = tst; c.split_mps_l2r(keep, 1e-16, a1, b1, t1); }
            std::vector<std::vector<double>> parts((size_t)world);
            for (int r = 0; r < world; ++r) {
                FakeCommEngine e(r, world, &parts[(size_t)r], nullptr);
                ts::TwoSiteTensor c = tst; MPSTensor a, b; ts::Truncation t; c.split_mps_l2r(keep, 1e-16, a, b, t, &e);
            }
            std::vector<double> total(parts[0].size(), 0.);
            for (auto const& v : parts) { if (v.size() != total.size()) throw std::runtime_error("ranks disagree on the buffer size"); for (size_t i = 0; i < v.size(); ++i) total[i] += v[i]; }
            std::vector<MPSTensor> as, bs;
            for (int r = 0; r < world; ++r) {
                FakeCommEngine e(r, world, nullptr, &total);
                ts::TwoSiteTensor c = tst; MPSTensor a, b; ts::Truncation t; c.split_mps_l2r(keep, 1e-16, a, b, t, &e);
                a.make_left_paired(); b.make_right_paired();
                as.push_back(a); bs.push_back(b);
                if (t.bond_dimension != t1.bond_dimension) same_struct = 0;
            }
            for (int r = 1; r < world; ++r) {
                if (!(as[(size_t)r].data().basis() == as[0].data().basis()) || !(bs[(size_t)r].data().basis() == bs[0].data().basis())) { identical = 0; continue; }
                for (size_t k = 0; k < as[0].data().n_blocks(); ++k) if (as[(size_t)r].data()[k].v != as[0].data()[k].v) identical = 0;
                for (size_t k = 0; k < bs[0].data().n_blocks(); ++k) if (bs[(size_t)r].data()[k].v != bs[0].data()[k].v) identical = 0;
            }
            a1.make_left_paired(); b1.make_right_paired();
            if (!(as[0].data().basis() == a1.data().basis())) same_struct = 0;
            block_matrix T1, T2; sweep::gemm(a1.data(), b1.data(), T1); sweep::gemm(as[0].data(), bs[0].data(), T2);
            worst = std::max(worst, rel_diff(compare(T2, T1)));
            ++n;
        }
        out[0] = n; out[1] = identical; out[2] = worst; out[3] = same_struct;
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// The descriptor entry points of the C ABI (qcm_mpo_upload, qcm_plan_sigma / left_step / right_step, qcm_plan_out_*): the
// problem is flattened into plain arrays (tests/harness/flatten_desc.hpp -- the binding a QCMaquis maintainer would write),
// planned INSIDE the library and executed; no C++ object of this repository crosses the boundary.  Compared with the oracle.
// out[0] sigma structure equal, [1] sigma rel. error, [2..3] left step, [4..5] right step, [6] noise plans with equal structure (of 2), [7] noise rel. error (single-site problems)
extern "C" int qcmt_desc_parity(const char* fcidump, const char* symm, int L, int nelec, int site, int twosite, int M, unsigned seed, double* out, char* err, int errlen)
{
    try {
#ifdef QCMT_WITH_GPU
        for (int i = 0; i < 8; ++i) out[i] = 0;
        Problem P = make_problem(fcidump, symm, L, nelec);
        SyntheticSite S = make_synthetic_site(P, site, twosite != 0, (size_t)M, seed);
        oracle::OracleEngine orc(P.params.symm);
        auto ck = [](int rc, const char* what) { if (rc != 0) throw std::runtime_error(std::string(what) + ": " + qcm_last_error()); };
        ck(qcm_init(0), "qcm_init");
        qcmflat::FlatMPO fm(P.params.symm, *S.mpo);
        qcmflat::FlatTensor ft(S.psi);
        qcmflat::FlatBoundary fl(S.left), fr(S.right);
        qcm_mpo_t m = nullptr; ck(qcm_mpo_upload(&fm.d, &m), "qcm_mpo_upload");
        qcm_array_t aL = nullptr, aR = nullptr;
        ck(qcm_array_alloc((int64_t)fl.data.size(), &aL), "alloc"); ck(qcm_array_upload(aL, 0, fl.data.data(), (int64_t)fl.data.size()), "upload");
        ck(qcm_array_alloc((int64_t)fr.data.size(), &aR), "alloc"); ck(qcm_array_upload(aR, 0, fr.data.data(), (int64_t)fr.data.size()), "upload");
        auto fetch = [&](qcm_plan_t p, std::vector<int64_t>& ptr, std::vector<qcm_block>& blocks, std::vector<int64_t>& off, int64_t& n_elems) {
            int64_t aux = 0, nb = 0; ck(qcm_plan_out_size(p, &aux, &nb, &n_elems), "qcm_plan_out_size");
            ptr.assign((size_t)aux + 1, 0); blocks.resize((size_t)nb); off.resize((size_t)nb);
            ck(qcm_plan_out_blocks(p, ptr.data(), blocks.data(), off.data()), "qcm_plan_out_blocks");
        };
        {
            qcm_plan_t plan = nullptr; ck(qcm_plan_sigma(m, &ft.d, &fl.d, &fr.d, 0, 1, 0, &plan), "qcm_plan_sigma");
            std::vector<int64_t> ptr, off; std::vector<qcm_block> blocks; int64_t n = 0;
            fetch(plan, ptr, blocks, off, n);
            std::vector<double> sigma((size_t)n);
            ck(qcm_site_hamil2(plan, aL, aR, ft.data.data(), sigma.data()), "qcm_site_hamil2");
            block_matrix got = qcmflat::unflatten(blocks, off, ptr[0], ptr[1], sigma.data());
            MPSTensor so = orc.site_hamil2(S.psi, S.left, S.right, *S.mpo); so.make_left_paired();
            DiffReport d = compare(got, so.data());
            out[0] = d.structure_equal; out[1] = rel_diff(d);
            qcm_plan_destroy(plan);
        }
        if (!twosite)
            for (int dir = 0; dir < 2; ++dir) {
                qcm_plan_t plan = nullptr;
                if (dir == 0) ck(qcm_plan_left_step(m, &ft.d, &ft.d, &fl.d, 0, 1, 0, &plan), "qcm_plan_left_step");
                else ck(qcm_plan_right_step(m, &ft.d, &ft.d, &fr.d, 0, 1, 0, &plan), "qcm_plan_right_step");
                std::vector<int64_t> ptr, off; std::vector<qcm_block> blocks; int64_t n = 0;
                fetch(plan, ptr, blocks, off, n);
                qcm_array_t aO = nullptr; ck(qcm_array_alloc(n, &aO), "alloc");
                ck(qcm_boundary_step(plan, dir == 0 ? aL : aR, ft.data.data(), ft.data.data(), aO), "qcm_boundary_step");
                std::vector<double> flat((size_t)n); ck(qcm_array_download(aO, 0, flat.data(), n), "download");
                Boundary got; got.resize(ptr.size() - 1);
                for (size_t b = 0; b + 1 < ptr.size(); ++b) got[b] = qcmflat::unflatten(blocks, off, ptr[b], ptr[b + 1], flat.data());
                Boundary ref = dir == 0 ? orc.overlap_mpo_left_step(S.psi, S.psi, S.left, *S.mpo) : orc.overlap_mpo_right_step(S.psi, S.psi, S.right, *S.mpo);
                DiffReport d = compare(got, ref);
                out[2 + 2 * dir] = d.structure_equal; out[3 + 2 * dir] = rel_diff(d);
                qcm_array_free(aO); qcm_plan_destroy(plan);
            }
        if (!twosite)      // noise term through the descriptor entry points (kept blocks = those of the tensor's own structure)
            for (int dir = 0; dir < 2; ++dir) {
                qcm_plan_t plan = nullptr;
                if (dir == 0) ck(qcm_plan_noise_left(m, &ft.d, &fl.d, 0, &plan), "qcm_plan_noise_left");
                else ck(qcm_plan_noise_right(m, &ft.d, &fr.d, 0, &plan), "qcm_plan_noise_right");
                std::vector<int64_t> ptr, off; std::vector<qcm_block> blocks; int64_t n = 0;
                fetch(plan, ptr, blocks, off, n);
                if (ptr.size() != 2) throw std::runtime_error("noise plan: expected one bond entry");
                qcm_array_t aO = nullptr; ck(qcm_array_alloc(n, &aO), "alloc");
                ck(qcm_boundary_step(plan, dir == 0 ? aL : aR, ft.data.data(), ft.data.data(), aO), "qcm_boundary_step (noise)");
                std::vector<double> flat((size_t)n); ck(qcm_array_download(aO, 0, flat.data(), n), "download");
                block_matrix got = qcmflat::unflatten(blocks, off, ptr[0], ptr[1], flat.data());
                block_matrix ref = dir == 0 ? orc.noise_left(S.psi, S.left, *S.mpo) : orc.noise_right(S.psi, S.right, *S.mpo);
                if (dir == 0) S.psi.make_left_paired(); else S.psi.make_right_paired();
                DualIndex keep = S.psi.data().basis();
                DiffReport d = compare(ts::noise_kept(got, keep), ts::noise_kept(ref, keep));
                out[6] += d.structure_equal; out[7] = std::max(out[7], rel_diff(d));
                qcm_array_free(aO); qcm_plan_destroy(plan);
            }
        qcm_array_free(aL); qcm_array_free(aR); qcm_mpo_free(m);
        return 0;
#else
        (void)fcidump; (void)symm; (void)L; (void)nelec; (void)site; (void)twosite; (void)M; (void)seed; (void)out;
        throw std::runtime_error("harness built without GPU support");
#endif
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// Single-site sweeps from the reference's `init_type = const` state (mps_initializers.h:85-97: every allowed sector capped at
// init_bond_dimension, all entries 1, every site tensor normalised): RNG free, so the reference's own printed energies are
// a known answer for solver + sigma + boundary chain.  n_sigma[i] = Jacobi-Davidson iterations of micro-iteration i.
extern "C" int qcmt_ss_dmrg_const(const char* fcidump, const char* symm, int L, int nelec, int init_bond_dimension, int nsweeps, int engine_kind,
                                  double* energies, int* n_sigma, int n_max, int* n_out, char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        P.init_mps((size_t)init_bond_dimension, false, 1., 0);
        std::unique_ptr<EngineIface> eng;
        if (engine_kind < 0) eng.reset(new oracle::OracleEngine(P.params.symm));
        else if (engine_kind == 0) eng.reset(new qcmtest::InterpEngine(P.params.symm, 1, (long long)1 << 40));
        else {
#ifdef QCMT_WITH_GPU
            eng.reset(new GpuEngine(P.params.symm, 0, 0, 1));
#else
            throw std::runtime_error("harness built without GPU support");
#endif
        }
        sweep::SweepLog log = sweep::ss_sweeps(*eng, P.mpo, P.mps, nsweeps);
        int n = (int)std::min<size_t>(log.energies.size(), (size_t)n_max);
        for (int i = 0; i < n; ++i) { energies[i] = log.energies[i]; n_sigma[i] = log.n_sigma[i]; }
        *n_out = n;
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

static std::unique_ptr<EngineIface> make_engine(SymmKind symm, int engine_kind, long long budget = (long long)1 << 40)
{
    std::unique_ptr<EngineIface> eng;
    if (engine_kind < 0) eng.reset(new oracle::OracleEngine(symm));
    else if (engine_kind == 0) eng.reset(new qcmtest::InterpEngine(symm, 1, budget));
    else {
#ifdef QCMT_WITH_GPU
        eng.reset(new GpuEngine(symm, 0, 0, 1, budget));
#else
        throw std::runtime_error("harness built without GPU support");
#endif
    }
    return eng;
}

// Noise term of the perturbed density matrix (EngineIface::noise_left / noise_right; prediction.hpp:34-47,101-114,
// twositetensor.hpp:192-219,260-287) on every site of a random MPS and on every bond with a random two-site tensor seen as
// a fat single-site tensor: engine under test against the oracle's restatement of left/right_boundary_tensor_mpo.
// out[0] cases  out[1] all structures equal  out[2] max rel diff  out[3] max |noise| (non-trivial check)
extern "C" int qcmt_noise_parity(const char* fcidump, const char* symm, int L, int nelec, int Mmax, unsigned seed, int engine_kind, long long budget,
                                 double* out, char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        P.init_mps((size_t)Mmax, true, 0., seed);
        oracle::OracleEngine orc(P.params.symm);
        P.build_boundaries(orc);
        std::unique_ptr<EngineIface> eng = make_engine(P.params.symm, engine_kind, budget);
        double dmax = 0, nmax = 0; int st = 1, n = 0;
        auto check = [&](block_matrix const& a_all, block_matrix const& b_all, DualIndex const& keep) {
            block_matrix a = ts::noise_kept(a_all, keep), b = ts::noise_kept(b_all, keep);
            DiffReport d = compare(a, b);
            dmax = std::max(dmax, rel_diff(d)); st &= d.structure_equal; nmax = std::max(nmax, std::sqrt(d.ref_norm)); ++n;
        };
        for (int p = 0; p < L; ++p) {
            P.mps[p].make_left_paired(); DualIndex lb = P.mps[p].data().basis();
            P.mps[p].make_right_paired(); DualIndex rb = P.mps[p].data().basis();
            check(eng->noise_left(P.mps[p], P.left[p], P.mpo[p]), orc.noise_left(P.mps[p], P.left[p], P.mpo[p]), lb);
            check(eng->noise_right(P.mps[p], P.right[p + 1], P.mpo[p]), orc.noise_right(P.mps[p], P.right[p + 1], P.mpo[p]), rb);
        }
        for (int p = 0; p + 1 < L; ++p) {
            // the operands of TwoSiteTensor::predict_split_l2r / r2l: both-paired two-site data as a site tensor with a fat index
            ts::TwoSiteTensor tst(P.params.symm, P.mps[p], P.mps[p + 1]);
            MPSTensor fl = tst.fat_right_tensor(), fr = tst.fat_left_tensor();
            DualIndex both = tst.both_paired_basis();
            check(eng->noise_left(fl, P.left[p], P.mpo[p]), orc.noise_left(fl, P.left[p], P.mpo[p]), both);
            check(eng->noise_right(fr, P.right[p + 2], P.mpo[p + 1]), orc.noise_right(fr, P.right[p + 2], P.mpo[p + 1]), both);
        }
        out[0] = n; out[1] = st; out[2] = dmax; out[3] = nmax;
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// The reference's own run H2_2e4o.TI.SS (single-site, init_type = const, alpha_initial = 1e-10, truncation 1e-50, M = 100):
// single-site sweeps with the noise-perturbed subspace expansion (ts::NoiseGrow = MPS::grow_l2r_sweep / grow_r2l_sweep).
// bond_before/after[i]: "Bond dimension before / after truncation" of micro-iteration i (0 where the sweep turns).
extern "C" int qcmt_ss_dmrg_const_noise(const char* fcidump, const char* symm, int L, int nelec, int init_bond_dimension, int nsweeps, int engine_kind,
                                        double alpha, double cutoff, int Mmax, double* energies, int* n_sigma, int* bond_after, int n_max, int* n_out,
                                        char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        P.init_mps((size_t)init_bond_dimension, false, 1., 0);
        std::unique_ptr<EngineIface> eng = make_engine(P.params.symm, engine_kind);
        std::vector<ts::Truncation> trs;
        sweep::SweepLog log = sweep::ss_sweeps(*eng, P.mpo, P.mps, nsweeps, 10, 1e-8, ts::NoiseGrow{*eng, alpha, cutoff, (size_t)Mmax, &trs});
        int n = (int)std::min<size_t>(log.energies.size(), (size_t)n_max);
        // grow is skipped at the turning points (site L-1 forward, site 0 backward): map the truncation log onto micro-iterations
        size_t q = 0;
        for (int i = 0; i < n; ++i) {
            energies[i] = log.energies[i]; n_sigma[i] = log.n_sigma[i];
            const int w = i % (2 * L);
            const bool turning = (w == L - 1) || (w == 2 * L - 1);
            bond_after[i] = turning || q >= trs.size() ? 0 : (int)trs[q++].bond_dimension;
        }
        *n_out = n;
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// two-site sweeps with the noise-perturbed split (ts_optimize.hpp:198-215 predict_split_l2r / r2l), the reference's default path
extern "C" int qcmt_ts_dmrg_noise(const char* fcidump, const char* symm, int L, int nelec, int M0, int Mmax, int nsweeps, unsigned seed, int engine_kind,
                                  double alpha, double cutoff, double* energies, int n_max, int* n_out, double* info, char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        P.init_mps((size_t)M0, true, 0., seed);
        std::unique_ptr<EngineIface> eng = make_engine(P.params.symm, engine_kind);
        ts::TsParams prm; prm.Mmax = (size_t)Mmax; prm.alpha = alpha; prm.cutoff = cutoff;
        std::vector<size_t> dims;
        sweep::SweepLog log = ts::ts_sweeps(P.params.symm, *eng, P.mpo, [&](int p) -> MPOTensor const& { return P.twosite_mpo(p); }, P.mps, nsweeps, prm, &dims);
        int n = (int)std::min<size_t>(log.energies.size(), (size_t)n_max);
        for (int i = 0; i < n; ++i) energies[i] = log.energies[i];
        *n_out = n;
        double secs = 0; for (double s : log.sweep_seconds) secs += s;
        info[0] = (double)log.total_sigma; info[1] = secs; info[2] = log.energies.back();
        info[3] = dims.empty() ? 0. : (double)*std::max_element(dims.begin(), dims.end());
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// MPS-MPS overlap steps (contraction::Engine::overlap_left_step / overlap_right_step, move_boundary.hpp:21-64) along the whole
// chain for two different random states: the engine under test (boundary step with the identity MPO tensor, qcm/overlap.hpp)
// against the oracle's literal restatement.
// out[0] steps compared  out[1] structures equal  out[2] max rel diff  out[3] <bra|ket> from the left chain (engine)
// out[4] the same from the right chain (engine)  out[5] <bra|ket> oracle  out[6] norm^2 of the canonised ket through the engine
extern "C" int qcmt_overlap_parity(const char* fcidump, const char* symm, int L, int nelec, int Mbra, int Mket, unsigned seed, int engine_kind,
                                   double* out, char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        Problem Q = P;
        P.init_mps((size_t)Mket, true, 0., seed);
        Q.init_mps((size_t)Mbra, true, 0., seed + 101);
        const bool su2 = is_su2(P.params.symm);
        oracle::OracleEngine orc(P.params.symm);
        std::unique_ptr<EngineIface> eng = make_engine(P.params.symm, engine_kind);
        double dmax = 0; int st = 1, n = 0;
        auto check = [&](block_matrix const& a, block_matrix const& b) { DiffReport d = compare(a, b); dmax = std::max(dmax, rel_diff(d)); st &= d.structure_equal; ++n; };
        block_matrix le = Q.mps.left_boundary()[0], lo = le;
        for (int p = 0; p < L; ++p) {
            le = overlap_left_step(*eng, su2, Q.mps[p], P.mps[p], le);
            lo = orc.overlap_left_step(Q.mps[p], P.mps[p], lo);
            check(le, lo);
        }
        block_matrix re = Q.mps.right_boundary()[0], ro = re;
        for (int p = L - 1; p >= 0; --p) {
            re = overlap_right_step(*eng, su2, Q.mps[p], P.mps[p], re);
            ro = orc.overlap_right_step(Q.mps[p], P.mps[p], ro);
            check(re, ro);
        }
        out[0] = n; out[1] = st; out[2] = dmax; out[3] = le.trace(); out[4] = re.trace(); out[5] = lo.trace();
        sweep::canonize_to_first(P.mps);
        out[6] = overlap(*eng, su2, P.mps, P.mps);
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// single-site sweeps from a random MPS of bond dimension M0 with the noise-perturbed subspace expansion (alpha > 0) or the plain
// QR normalisation (alpha = 0); info as qcmt_ss_dmrg
extern "C" int qcmt_ss_dmrg_noise(const char* fcidump, const char* symm, int L, int nelec, int M0, int Mmax, int nsweeps, unsigned seed, int engine_kind,
                                  double alpha, double cutoff, double* energies, int n_max, int* n_out, double* info, char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        P.init_mps((size_t)M0, true, 0., seed);
        std::unique_ptr<EngineIface> eng = make_engine(P.params.symm, engine_kind);
        sweep::SweepLog log = alpha != 0. ? sweep::ss_sweeps(*eng, P.mpo, P.mps, nsweeps, 10, 1e-8, ts::NoiseGrow{*eng, alpha, cutoff, (size_t)Mmax, nullptr})
                                          : sweep::ss_sweeps(*eng, P.mpo, P.mps, nsweeps);
        int n = (int)std::min<size_t>(log.energies.size(), (size_t)n_max);
        for (int i = 0; i < n; ++i) energies[i] = log.energies[i];
        *n_out = n;
        double secs = 0; for (double s : log.sweep_seconds) secs += s;
        info[0] = (double)log.total_sigma; info[1] = secs; info[2] = log.energies.back(); info[3] = 2.0 * L;
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// A hand-made plan at the task-array level of the C ABI (qcm_plan_create): one STORE-mode step-1 output whose K-segment list is
// longer than 96 K-chunks (K = 1600 + 300 + 200: the situation of an SU2 site whose leading sector passes ~1.5k rows; store-mode
// outputs must run their whole K list in one work item) feeding one ACCUMULATING closing output whose list (K = 1000 + 400 + 300) is
// split and combined with FP64 atomics.  sigma = (A^T psi) R, checked against host dgemm.
// out[0] max rel error of sigma  out[1] |sigma|_max
extern "C" int qcmt_long_k_plan(double* out, char* err, int errlen)
{
    try {
#ifdef QCMT_WITH_GPU
        auto ck = [](int rc, const char* what) { if (rc != 0) throw std::runtime_error(std::string(what) + ": " + qcm_last_error()); };
        ck(qcm_init(0), "qcm_init");
        const int m = 96, n = 1700, n2 = 64, ks[3] = {1600, 300, 200}, K = 2100, cs[3] = {1000, 400, 300};
        UniformGen gen(123);
        std::vector<double> A((size_t)K * m), psi((size_t)K * n), R((size_t)n * n2);
        for (double& x : A) x = gen() - 0.5; for (double& x : psi) x = gen() - 0.5; for (double& x : R) x = gen() - 0.5;
        // step 1 (persistent group, QCM_BUF_TP): T(m x n) = sum over K slices of A[k0:k0+k, :]^T psi[k0:k0+k, :]
        std::vector<qcm_gemm_seg> ps, cseg; std::vector<qcm_gemm_out> po(1), co(1);
        int k0 = 0;
        for (int s = 0; s < 3; ++s) { ps.push_back(qcm_gemm_seg{qcm_ref{QCM_BUF_LEFT, 0, k0}, qcm_ref{QCM_BUF_KET_LP, 0, k0}, K, K, m, n, ks[s], 1, 0, 0, 1.0}); k0 += ks[s]; }
        po[0] = qcm_gemm_out{qcm_ref{QCM_BUF_TP, 0, 0}, m, m, n, 0, 3, 0};
        // step 3: sigma(m x n2) += T[:, c0:c0+c] R[c0:c0+c, :]
        int c0 = 0;
        for (int s = 0; s < 3; ++s) { cseg.push_back(qcm_gemm_seg{qcm_ref{QCM_BUF_TP, 0, (int64_t)c0 * m}, qcm_ref{QCM_BUF_RIGHT, 0, c0}, m, n, m, n2, cs[s], 0, 0, 0, 1.0}); c0 += cs[s]; }
        co[0] = qcm_gemm_out{qcm_ref{QCM_BUF_OUT, 0, 0}, m, m, n2, 0, 3, 0};
        qcm_wave_desc w; std::memset(&w, 0, sizeof(w));
        w.c_outs = co.data(); w.n_c_outs = 1; w.c_segs = cseg.data(); w.n_c_segs = 3;
        qcm_plan_desc d; std::memset(&d, 0, sizeof(d));
        d.kind = 0; d.n_waves = 1; d.waves = &w; d.p_outs = po.data(); d.n_p_outs = 1; d.p_segs = ps.data(); d.n_p_segs = 3;
        d.elems[QCM_BUF_KET_LP] = (int64_t)K * n; d.elems[QCM_BUF_LEFT] = (int64_t)K * m; d.elems[QCM_BUF_RIGHT] = (int64_t)n * n2;
        d.elems[QCM_BUF_TP] = (int64_t)m * n; d.elems[QCM_BUF_OUT] = (int64_t)m * n2;
        d.flops = 2.0 * m * n * K + 2.0 * m * n2 * n; d.world = 1;
        qcm_plan_t plan = nullptr; ck(qcm_plan_create(&d, &plan), "qcm_plan_create");
        qcm_array_t aL = nullptr, aR = nullptr;
        ck(qcm_array_alloc((int64_t)A.size(), &aL), "alloc"); ck(qcm_array_upload(aL, 0, A.data(), (int64_t)A.size()), "upload");
        ck(qcm_array_alloc((int64_t)R.size(), &aR), "alloc"); ck(qcm_array_upload(aR, 0, R.data(), (int64_t)R.size()), "upload");
        std::vector<double> sigma((size_t)m * n2, 0.);
        for (int rep = 0; rep < 2; ++rep) ck(qcm_site_hamil2(plan, aL, aR, psi.data(), sigma.data()), "qcm_site_hamil2");     // twice: the output is re-zeroed per call
        std::vector<double> T((size_t)m * n), ref((size_t)m * n2);
        dgemm(MatRef{A.data(), (size_t)m, (size_t)K, (size_t)K, true}, MatRef{psi.data(), (size_t)K, (size_t)n, (size_t)K, false}, 1.0, 0.0, T.data(), (size_t)m);
        dgemm(MatRef{T.data(), (size_t)m, (size_t)n, (size_t)m, false}, MatRef{R.data(), (size_t)n, (size_t)n2, (size_t)n, false}, 1.0, 0.0, ref.data(), (size_t)m);
        double dmax = 0, rmax = 0;
        for (size_t i = 0; i < ref.size(); ++i) { dmax = std::max(dmax, std::abs(sigma[i] - ref[i])); rmax = std::max(rmax, std::abs(ref[i])); }
        out[0] = dmax / rmax; out[1] = rmax;
        qcm_plan_destroy(plan); qcm_array_free(aL); qcm_array_free(aR);
        return 0;
#else
        (void)out;
        throw std::runtime_error("harness built without GPU support");
#endif
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

extern "C" int qcmt_excited_states_driver(const char* fcidump, const char* symm, int L, int nelec, int Mmax, int nsweeps, int nstates, int engine_kind, int twosite,
                                          double* energies, double* overlaps, char* err, int errlen);
// Excited states by orthogonal-state projection (optimize.h ortho_mps; ss_optimize.hpp:107-111; ietl/jacobi.h:378,393,432): state 0 is
// optimised first, state k is then optimised orthogonal to states 0 .. k-1.  Single-site sweeps from random MPS of the full bond
// dimension with the noise-perturbed subspace expansion.  energies[k] = last energy of state k; overlaps[k] = |<state k | state 0>|.
extern "C" int qcmt_excited_states(const char* fcidump, const char* symm, int L, int nelec, int Mmax, int nsweeps, int nstates, int engine_kind,
                                   double* energies, double* overlaps, char* err, int errlen)
{
    return qcmt_excited_states_driver(fcidump, symm, L, nelec, Mmax, nsweeps, nstates, engine_kind, 0, energies, overlaps, err, errlen);
}
// twosite != 0: two-site sweeps (ts_optimize.hpp:120-128: the orthogonal states enter as two-site tensors) from M0 = 4 random states
extern "C" int qcmt_excited_states_driver(const char* fcidump, const char* symm, int L, int nelec, int Mmax, int nsweeps, int nstates, int engine_kind, int twosite,
                                          double* energies, double* overlaps, char* err, int errlen)
{
    try {
        Problem P = make_problem(fcidump, symm, L, nelec);
        std::unique_ptr<EngineIface> eng = make_engine(P.params.symm, engine_kind);
        const bool su2 = is_su2(P.params.symm);
        sweep::OrthoStates found; found.su2 = su2;
        for (int k = 0; k < nstates; ++k) {
            P.init_mps(twosite ? (size_t)4 : (size_t)Mmax, true, 0., 42u + 17u * (unsigned)k);
            sweep::SweepLog log;
            if (twosite) {
                ts::TsParams prm; prm.Mmax = (size_t)Mmax; prm.ortho = &found;
                log = ts::ts_sweeps(P.params.symm, *eng, P.mpo, [&](int p) -> MPOTensor const& { return P.twosite_mpo(p); }, P.mps, nsweeps, prm);
            } else
                log = sweep::ss_sweeps(*eng, P.mpo, P.mps, nsweeps, 10, 1e-8, ts::NoiseGrow{*eng, 1e-6, 1e-14, (size_t)Mmax, nullptr}, &found);
            energies[k] = log.energies.back();
            sweep::canonize_to_first(P.mps);
            overlaps[k] = k == 0 ? 1. : std::abs(overlap(*eng, su2, found.states[0], P.mps));
            found.states.push_back(P.mps);
        }
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}
