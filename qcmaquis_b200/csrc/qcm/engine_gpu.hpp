// qcm::GpuEngine -- the B200 implementation of the reference's contraction::Engine calls on the sweep hot path
// (abelian/engine.hpp:102-122,196-209; non-abelian/engine.hpp:197-207).  Same call signatures and value
// semantics as the reference (ket by value, boundaries by const reference, results returned by value), so the
// sweep drivers (BoundaryPropagator.h:98-122, SiteProblem / ietl::mult, ietl_lanczos_solver.h:108-115) keep
// working unchanged.  What differs is where the work happens:
//   * the block loops of site_hamil2 / overlap_mpo_{left,right}_step are walked once by plan::Planner and the
//     resulting task arrays are handed to the C ABI (include/qcm_b200.h); the plan is cached per
//     (MPO tensor, boundaries, tensor structure) and reused by every Davidson iteration at that site,
//   * boundaries returned by the boundary steps stay in HBM (Boundary::device_mirror); the host object carries
//     the symmetry-block structure only until download() is called,
//   * with a communicator (one process per GPU) every rank executes its share of the edges of the MPO bond graph
//     (sharded by their step-1 index, plan.hpp shard_sources) and the library sums the partial results,
//   * the Jacobi-Davidson solve of a site can run inside the engine with the solver vectors resident in HBM.
// There is no CPU fallback: every failure of the device layer surfaces as std::runtime_error.
#pragma once
#include "../../../include/qcm_b200.h"
#include "engine_iface.hpp"
#include "plan.hpp"
#include "plan_to_desc.hpp"
#include <list>

extern "C" void scipy_dsyev_(const char* jobz, const char* uplo, const int* n, double* a, const int* lda, double* w, double* work, const int* lwork, int* info);

namespace qcm {

inline void qcm_check(int status, const char* what)
{
    if (status != 0) throw std::runtime_error(std::string(what) + ": " + qcm_last_error());
}

// pinned host buffers of the spill tier: page-locking gigabytes takes longer than copying them, so buffers are recycled
struct SpillPool
{
    std::vector<std::pair<int64_t, double*>> free_list;
    double* take(int64_t n, int64_t& cap)
    {
        size_t best = free_list.size();
        for (size_t i = 0; i < free_list.size(); ++i)
            if (free_list[i].first >= n && (best == free_list.size() || free_list[i].first < free_list[best].first)) best = i;
        if (best != free_list.size()) { double* p = free_list[best].second; cap = free_list[best].first; free_list.erase(free_list.begin() + best); return p; }
        double* p = nullptr; cap = n + n / 8;
        qcm_check_status(qcm_pinned_alloc(cap, &p));
        return p;
    }
    void give(double* p, int64_t cap) { if (p) free_list.push_back(std::make_pair(cap, p)); }
    static void qcm_check_status(int rc) { if (rc != 0) throw std::runtime_error(std::string("qcm_pinned_alloc: ") + qcm_last_error()); }
    ~SpillPool() { for (auto& e : free_list) qcm_pinned_free(e.second); }
};

// Device arrays of boundaries the sweep has dropped, kept for the boundaries it creates next.  A sweep frees one boundary and
// allocates one of nearly the same size at every site (the stale one at a bond makes room for the new one at the same bond);
// asking the device pool for gigabytes of a slightly different size each time cost 50-100 ms per site
// (profiles/r02n: 15 s of a 128 s cfg3 sweep).  Arrays are requested in size classes of 1/8 octave so that the one just given
// back fits; all device work is ordered on the library's stream, so a recycled array can be handed out at once.
struct ArrayRecycler
{
    std::vector<std::pair<int64_t, qcm_array_t>> kept;     // (capacity in elements, array)
    size_t max_kept = 3;
    static int64_t size_class(int64_t n)
    {
        if (n < ((int64_t)1 << 23)) return n;               // below 64 MB: exact
        int64_t gran = (int64_t)1 << 20;
        while ((gran << 4) <= n) gran <<= 1;                // 1/8 .. 1/16 of the size
        return (n + gran - 1) / gran * gran;
    }
    qcm_array_t take(int64_t n)
    {
        size_t best = kept.size();
        for (size_t i = 0; i < kept.size(); ++i)
            if (kept[i].first >= n && kept[i].first <= n + n / 4 + 1024 && (best == kept.size() || kept[i].first < kept[best].first)) best = i;
        if (best != kept.size()) { qcm_array_t a = kept[best].second; kept.erase(kept.begin() + (long)best); ++hits; return a; }
        ++misses;
        qcm_array_t a = nullptr;
        if (qcm_array_alloc(size_class(n), &a) != 0) {
            clear();                                        // memory is short: give everything back and ask for the exact size
            if (qcm_array_alloc(n, &a) != 0) throw std::runtime_error(std::string("qcm_array_alloc: ") + qcm_last_error());
        }
        return a;
    }
    void give(qcm_array_t a)
    {
        int64_t cap = 0;
        int resident = 0;
        if (!a) return;
        if (qcm_array_size(a, &cap) != 0 || qcm_array_resident(a, &resident) != 0 || !resident || cap < ((int64_t)1 << 23)) { qcm_array_free(a); return; }
        kept.push_back(std::make_pair(cap, a));
        while (kept.size() > max_kept) {                    // the smallest goes
            size_t m = 0;
            for (size_t i = 1; i < kept.size(); ++i) if (kept[i].first < kept[m].first) m = i;
            qcm_array_free(kept[m].second); kept.erase(kept.begin() + (long)m);
        }
    }
    void clear() { for (auto& e : kept) qcm_array_free(e.second); kept.clear(); }
    size_t hits = 0, misses = 0;
    ~ArrayRecycler() { clear(); }
};

struct DeviceBoundary
{
    qcm_array_t arr = nullptr;
    plan::BoundaryLayout layout;
    double* spill = nullptr; int64_t spill_cap = 0;     // pinned host copy while the boundary is evicted from HBM
    std::shared_ptr<SpillPool> pool;
    std::shared_ptr<ArrayRecycler> recycler;            // set for boundaries a GpuEngine produced: the array goes back there
    bool evicted = false;
    ~DeviceBoundary()
    {
        if (arr) { if (recycler && !evicted) recycler->give(arr); else qcm_array_free(arr); }
        if (spill) { if (evicted) qcm_sync(); if (pool) pool->give(spill, spill_cap); else qcm_pinned_free(spill); }
    }
    // 64-bit hash of the block structure (charges and sizes of every block of every bond entry): a plan depends on the
    // structure of its boundaries only, never on their values, so plans are keyed by it and survive from sweep to sweep
    uint64_t structure_hash() const
    {
        if (hash_) return hash_;
        uint64_t h = 1469598103934665603ull;
        auto mix = [&](uint64_t x) { h ^= x; h *= 1099511628211ull; };
        mix(layout.b.size());
        for (auto const& l : layout.b) {
            mix(l.basis.size() + 0x9E37u);
            for (auto const& q : l.basis) { mix((uint32_t)q.lc[0]); mix((uint32_t)q.lc[1]); mix((uint32_t)q.lc[2]); mix((uint32_t)q.rc[0]); mix((uint32_t)q.rc[1]); mix((uint32_t)q.rc[2]); mix(q.ls); mix(q.rs); }
        }
        mix((uint64_t)layout.total);
        hash_ = h ? h : 1;
        return hash_;
    }
private:
    mutable uint64_t hash_ = 0;
};

struct CompiledPlan
{
    qcm_plan_t handle = nullptr;
    std::vector<qcm_plan_t> slices;      // time-sliced shards of one sigma contraction (GpuEngine::slices > 1); handle == slices[0]
    plan::Layout out_tensor;
    plan::BoundaryLayout out_boundary;
    int64_t ket_elems = 0, bra_elems = 0, out_elems = 0;
    double flops = 0, flops_t = 0, flops_w = 0, flops_close = 0;
    int64_t bytes = 0;
    size_t n_gemm_tasks = 0, n_axpy_tasks = 0, n_waves = 0;
    int64_t w_elems_read = 0, w_elems_written = 0, w_groups = 0;
    double exec_w = 0, exec_close = 0;
    int64_t direct_panel_elems = 0, w_panel_elems = 0, skipped_panel_elems = 0;
    int64_t workspace_elems = 0;
    ~CompiledPlan() { if (slices.empty()) { if (handle) qcm_plan_destroy(handle); } else for (qcm_plan_t p : slices) qcm_plan_destroy(p); }
};

class GpuEngine : public EngineIface
{
public:
    // Workspace budget (elements of T + Y per wave).  An output index whose step-1 products and W panels do not fit is pushed into
    // the next wave -- but destinations in different waves cannot share one read of their sources, so a tight budget inflates the
    // W pass's HBM traffic (cfg3: 163 GB per sigma in two waves at 2^31 elements, 109 GB in one wave).  2^32 elements = 34 GB.
    static constexpr int64_t kDefaultBudget = (int64_t)1 << 32;
    explicit GpuEngine(SymmKind s, int device = 0, int rank_ = 0, int world_ = 1, int64_t ws_budget_elems = kDefaultBudget)
        : symm(s), rank(rank_), world(world_), budget(ws_budget_elems)
    {
        qcm_check(qcm_init(device), "qcm_init");
        if (const char* e = getenv("QCM_SLICES")) slices = std::max(1, atoi(e));
        if (const char* e = getenv("QCM_WS_BUDGET_GB")) budget = std::max<int64_t>(1, (int64_t)(atof(e) * 1e9 / 8));
    }
    ~GpuEngine() { for (auto a : vec_pool) qcm_array_free(a); if (host_xfer) qcm_array_free(host_xfer); }

    // ---- Engine::site_hamil2 ----------------------------------------------------------------------------
    MPSTensor site_hamil2(MPSTensor ket_tensor, Boundary const& left, Boundary const& right, MPOTensor const& mpo,
                          bool isHermitian = true) override
    {
        ket_tensor.make_left_paired();
        std::shared_ptr<DeviceBoundary> dl = mirror(left), dr = mirror(right);
        std::shared_ptr<CompiledPlan> cp = sigma_plan(ket_tensor, dl, dr, mpo, isHermitian);
        Clock c0;
        std::vector<double> psi = flatten(ket_tensor.data(), cp->ket_elems), sigma((size_t)cp->out_elems);
        seconds[4] += c0.lap();
        run_sigma_host(*cp, *dl, *dr, psi.data(), sigma.data());
        seconds[2] += c0.lap();
        MPSTensor r(ket_tensor.site_dim(), ket_tensor.row_dim(), ket_tensor.col_dim(), unflatten(cp->out_tensor, sigma), LeftPaired, true);
        seconds[4] += c0.lap();
        return r;
    }

    std::shared_ptr<CompiledPlan> sigma_plan(MPSTensor const& ket_tensor, std::shared_ptr<DeviceBoundary> const& dl,
                                             std::shared_ptr<DeviceBoundary> const& dr, MPOTensor const& mpo, bool isHermitian = true)
    {
        ket_tensor.make_left_paired();
        PlanKey key{mpo.uid(), dl->structure_hash(), dr->structure_hash(), structure_hash(ket_tensor), isHermitian ? 0 : 3};
        if (std::shared_ptr<CompiledPlan> hit = lookup(key, [&](Witness const& w) { return w.matches(dl->layout, dr->layout, ket_tensor, ket_tensor); })) { ++cache_hits; return hit; }
        ++cache_misses;
        make_room();
        Clock c0;
        std::shared_ptr<CompiledPlan> cp;
        if (slices > 1 && world == 1) {
            // the site problem as `slices` shards, run one after another on this device (qcm_site_hamil2_sliced)
            for (int v = 0; v < slices; ++v) {
                plan::Planner planner(symm, mpo, isHermitian, v, slices, budget);
                plan::Plan P = planner.plan_sigma(desc_of(ket_tensor), dl->layout, dr->layout);
                seconds[0] += c0.lap();
                std::shared_ptr<CompiledPlan> part = compile(P, dl->layout.total, dr->layout.total);
                seconds[1] += c0.lap();
                if (v == 0) cp = part;
                else {
                    cp->flops += part->flops; cp->flops_t += part->flops_t; cp->flops_w += part->flops_w; cp->flops_close += part->flops_close;
                    cp->exec_w += part->exec_w; cp->exec_close += part->exec_close; cp->n_gemm_tasks += part->n_gemm_tasks; cp->n_axpy_tasks += part->n_axpy_tasks;
                    cp->w_elems_read += part->w_elems_read; cp->w_elems_written += part->w_elems_written; cp->w_groups += part->w_groups;
                    cp->direct_panel_elems += part->direct_panel_elems; cp->w_panel_elems += part->w_panel_elems; cp->skipped_panel_elems += part->skipped_panel_elems;
                    cp->workspace_elems = std::max(cp->workspace_elems, part->workspace_elems);
                }
                cp->slices.push_back(part->handle);
                if (v > 0) part->handle = nullptr;       // owned by cp->slices now
            }
            last = cp;
        } else {
            plan::Planner planner(symm, mpo, isHermitian, rank, world, budget);
            plan::Plan P = planner.plan_sigma(desc_of(ket_tensor), dl->layout, dr->layout);
            seconds[0] += c0.lap();
            cp = compile(P, dl->layout.total, dr->layout.total);
            seconds[1] += c0.lap();
        }
        remember(key, Witness(dl->layout, dr->layout, ket_tensor, ket_tensor), cp);
        return cp;
    }

    // one sigma evaluation of a compiled plan (single plan or time-sliced shards)
    void run_sigma_dev(CompiledPlan const& cp, DeviceBoundary const& dl, DeviceBoundary const& dr, qcm_array_t psi, qcm_array_t sigma)
    {
        if (cp.slices.empty()) qcm_check(qcm_site_hamil2_dev(cp.handle, dl.arr, dr.arr, psi, sigma), "qcm_site_hamil2_dev");
        else qcm_check(qcm_site_hamil2_sliced_dev(cp.slices.data(), (int)cp.slices.size(), dl.arr, dr.arr, psi, sigma), "qcm_site_hamil2_sliced_dev");
        sigma_flops += cp.flops; ++n_sigma_calls;
    }
    void run_sigma_host(CompiledPlan const& cp, DeviceBoundary const& dl, DeviceBoundary const& dr, const double* psi, double* sigma)
    {
        if (cp.slices.empty()) qcm_check(qcm_site_hamil2(cp.handle, dl.arr, dr.arr, psi, sigma), "qcm_site_hamil2");
        else qcm_check(qcm_site_hamil2_sliced(cp.slices.data(), (int)cp.slices.size(), dl.arr, dr.arr, psi, sigma), "qcm_site_hamil2_sliced");
        sigma_flops += cp.flops; ++n_sigma_calls;
    }
    // > 1: every sigma plan is built as that many shards executed one after another (for site problems whose resident step-1
    // products exceed the device: per-shard TP is 1/slices); ignored when the engine is one rank of several
    int slices = 1;

    // ---- Jacobi-Davidson with device-resident vectors (ietl/jacobi.h:361-451, ietl_jcd_gmres = 0) ---------------------
    // Same recurrence as the host solver (qcm/sweep.hpp) on flat device arrays: one qcm_site_hamil2_dev per iteration, the small
    // projected eigenproblem on the host, correction t = -r + (r.u / u.u) u.  psi goes up once and the eigenvector comes down
    // once per site.  Scalars reach the host in THREE batched reads per iteration (qcm_vec_dots):
    //   (1) after sigma: the new column of the projected matrix, V_i . (H v);
    //   (2) after the Ritz pair: r.r, r.u, u.u and V_i . r, V_i . u -- from these the Gram-Schmidt coefficients of the
    //       correction vector follow without touching it, so the first orthogonalisation pass is ONE fused linear combination
    //       t' = -r + a u - sum_i c_i V_i;
    //   (3) the second pass ("twice is enough"): V_i . t' and t'.t', applied together with the normalisation in one more
    //       linear combination.
    // The host solver orthogonalises by modified Gram-Schmidt with conditional refinement (jacobi.h:166-186); the two-pass
    // classical scheme spans the same Krylov space and is as stable, energies agree to rounding (tests: <= 1e-8 Eh per
    // micro-iteration).  Needs sigma and psi in ONE block layout (true for a consistent site problem); otherwise the host
    // solver is used.
    bool jacobi_davidson(MPSTensor const& x0, Boundary const& left, Boundary const& right, MPOTensor const& mpo, int max_iter, double tol,
                         EigenResult& res) override
    {
        if (!device_solver || max_iter < 1 || 2 * max_iter + 3 > QCM_MAX_DOTS) return false;
        x0.make_left_paired();
        std::shared_ptr<DeviceBoundary> dl = mirror(left), dr = mirror(right);
        std::shared_ptr<CompiledPlan> cp = sigma_plan(x0, dl, dr, mpo, true);
        if (!(cp->out_tensor.basis == x0.data().basis()) || cp->out_elems != cp->ket_elems) return false;
        Clock c0;
        const int64_t n = cp->ket_elems;
        const size_t need = 2 * (size_t)max_iter + 4;
        // every pool entry holds vec_pool_n elements: a larger site rebuilds the pool, a smaller one reuses it
        if (vec_pool_n < n) { for (auto a : vec_pool) qcm_array_free(a); vec_pool.clear(); vec_pool_n = n; }
        while (vec_pool.size() < need) { qcm_array_t a = nullptr; qcm_check(qcm_array_alloc(vec_pool_n, &a), "qcm_array_alloc"); vec_pool.push_back(a); }
        std::vector<qcm_array_t> V(vec_pool.begin(), vec_pool.begin() + max_iter + 1), VA(vec_pool.begin() + max_iter + 1, vec_pool.begin() + 2 * max_iter + 1);
        qcm_array_t u = vec_pool[2 * (size_t)max_iter + 1], r = vec_pool[2 * (size_t)max_iter + 2], w = vec_pool[2 * (size_t)max_iter + 3];
        std::vector<qcm_array_t> xs, ys; std::vector<double> d, cf;
        // sharded runs: every rank computes the scalars from its own copy of the (bit-identical) vectors; rank 0's values are
        // handed to all ranks anyway, so that no rank can ever take a different convergence decision
        auto dots = [&]() {
            d.assign(xs.size(), 0.); qcm_check(qcm_vec_dots(xs.data(), ys.data(), (int)xs.size(), n, d.data()), "qcm_vec_dots"); ++host_syncs;
            if (world > 1) { if (rank != 0) std::fill(d.begin(), d.end(), 0.); allreduce_sum(d.data(), d.size()); }
        };
        auto lincomb = [&](qcm_array_t out) { qcm_check(qcm_vec_lincomb(xs.data(), cf.data(), (int)xs.size(), out, n), "qcm_vec_lincomb"); };
        {
            std::vector<double> psi = flatten(x0.data(), n);
            qcm_check(qcm_array_upload(w, 0, psi.data(), n), "qcm_array_upload");
            xs = {w}; ys = {w}; dots();
            xs = {w}; cf = {1. / std::sqrt(d[0])}; lincomb(V[0]);
        }
        std::vector<double> M((size_t)max_iter * max_iter, 0.);
        res = EigenResult();
        int it = 0;
        for (;;) {
            run_sigma_dev(*cp, *dl, *dr, V[it], VA[it]);
            res.n_sigma++;
            xs.assign(V.begin(), V.begin() + it + 1); ys.assign((size_t)it + 1, VA[it]); dots();                            // (1)
            for (int i = 0; i <= it; ++i) M[(size_t)i + (size_t)it * max_iter] = d[(size_t)i];
            const int dim = it + 1;
            std::vector<double> A((size_t)dim * dim), ev(dim), work(std::max(1, 3 * dim));
            for (int c = 0; c < dim; ++c) for (int q = 0; q <= c; ++q) A[(size_t)q + (size_t)c * dim] = M[(size_t)q + (size_t)c * max_iter];
            int lwork = (int)work.size(), info = 0;
            scipy_dsyev_("V", "U", &dim, A.data(), &dim, ev.data(), work.data(), &lwork, &info);
            if (info) throw std::runtime_error("dsyev failed in the Jacobi-Davidson subspace problem");
            const double theta = ev[0];
            const double* sv = A.data();
            // u = sum_j s_j V_j ;  r = sum_j s_j (H V_j) - theta u
            xs.assign(V.begin(), V.begin() + dim); cf.assign(sv, sv + dim); lincomb(u);
            xs.assign(VA.begin(), VA.begin() + dim); xs.insert(xs.end(), V.begin(), V.begin() + dim);
            cf.assign(sv, sv + dim); for (int j = 0; j < dim; ++j) cf.push_back(-theta * sv[j]);
            lincomb(r);
            xs = {r, r, u}; ys = {r, u, u};
            for (int i = 0; i < dim; ++i) { xs.push_back(V[i]); ys.push_back(r); }
            for (int i = 0; i < dim; ++i) { xs.push_back(V[i]); ys.push_back(u); }
            dots();                                                                                                          // (2)
            ++it;
            const double rn = std::sqrt(d[0]);
            res.theta = theta; res.resid = rn;
            if (rn <= tol * std::abs(theta) || rn <= tol || it >= max_iter) break;
            // correction without GMRES steps, t = -r + (r.u / u.u) u, orthogonalised against V_0 .. V_{it-1} in two passes
            const double alpha = d[1] / d[2];
            xs = {r, u}; cf = {-1., alpha};
            for (int i = 0; i < dim; ++i) { xs.push_back(V[i]); cf.push_back(-(-d[3 + (size_t)i] + alpha * d[3 + (size_t)dim + (size_t)i])); }
            lincomb(w);
            xs.assign(V.begin(), V.begin() + dim); ys.assign((size_t)dim, w); xs.push_back(w); ys.push_back(w); dots();      // (3)
            double nrm2 = d[(size_t)dim];
            for (int i = 0; i < dim; ++i) nrm2 -= d[(size_t)i] * d[(size_t)i];
            const double inv = 1. / std::sqrt(nrm2 > 0 ? nrm2 : d[(size_t)dim]);
            xs = {w}; cf = {inv};
            for (int i = 0; i < dim; ++i) { xs.push_back(V[i]); cf.push_back(-d[(size_t)i] * inv); }
            lincomb(V[it]);
        }
        std::vector<double> out((size_t)n);
        if (world > 1) {       // every rank continues from rank 0's eigenvector
            if (rank != 0) qcm_check(qcm_array_zero(u), "qcm_array_zero");
            qcm_check(qcm_comm_allreduce(u, n), "qcm_comm_allreduce");
        }
        qcm_check(qcm_array_download(u, 0, out.data(), n), "qcm_array_download");
        res.vec = MPSTensor(x0.site_dim(), x0.row_dim(), x0.col_dim(), unflatten(cp->out_tensor, out), LeftPaired, true);
        seconds[2] += c0.lap();
        return true;
    }
    size_t host_syncs = 0;           // batched scalar reads of the device solver (three per Jacobi-Davidson iteration)
    // ---- spill tier: the storage::disk protocol of the reference's sweep drivers (utils/storage.h:113-185: prefetch / evict /
    // drop, called around every site by optimize.h:119-165) with pinned host memory as the backing store.  evict() queues a
    // device-to-host copy on the library's copy stream and gives the HBM back; prefetch() queues the upload; both return at
    // once.  A boundary that is used while evicted is fetched on the spot.
    void evict(Boundary const& b) override
    {
        std::shared_ptr<DeviceBoundary> d = std::static_pointer_cast<DeviceBoundary>(b.device_mirror);
        if (!d || d->evicted || d->layout.total == 0) return;
        int64_t held = 0;                       // a recycled array may be larger than the boundary: the copy moves all of it
        qcm_check(qcm_array_size(d->arr, &held), "qcm_array_size");
        if (!d->spill) { d->pool = spill_pool; d->spill = spill_pool->take(std::max(held, d->layout.total), d->spill_cap); }
        qcm_check(qcm_array_evict(d->arr, d->spill), "qcm_array_evict");
        d->evicted = true; ++n_evicted;
    }
    void prefetch(Boundary const& b) override
    {
        std::shared_ptr<DeviceBoundary> d = std::static_pointer_cast<DeviceBoundary>(b.device_mirror);
        if (!d || !d->evicted) return;
        qcm_check(qcm_array_prefetch(d->arr, d->spill), "qcm_array_prefetch");
        d->evicted = false; ++n_prefetched;
    }
    size_t n_evicted = 0, n_prefetched = 0;
    std::shared_ptr<SpillPool> spill_pool{new SpillPool()};
    std::shared_ptr<ArrayRecycler> recycler{new ArrayRecycler()};

    // ---- host-side collectives of a sharded sweep (EngineIface) --------------------------------------------------------
    int comm_rank() const override { return rank; }
    int comm_world() const override { return world; }
    void allreduce_sum(double* buf, size_t n) override
    {
        if (world <= 1 || n == 0) return;
        int64_t have = 0;
        if (host_xfer) qcm_array_size(host_xfer, &have);
        if (have < (int64_t)n) { if (host_xfer) qcm_array_free(host_xfer); host_xfer = nullptr; qcm_check(qcm_array_alloc((int64_t)n + (int64_t)n / 4, &host_xfer), "qcm_array_alloc"); }
        qcm_check(qcm_array_upload(host_xfer, 0, buf, (int64_t)n), "qcm_array_upload");
        qcm_check(qcm_comm_allreduce(host_xfer, (int64_t)n), "qcm_comm_allreduce");
        qcm_check(qcm_array_download(host_xfer, 0, buf, (int64_t)n), "qcm_array_download");
    }
    void assert_consistent(uint64_t fp, const char* what) override
    {
        if (world <= 1) return;
        // 16-bit pieces h_q of the fingerprint: all ranks agree  <=>  world * sum(h_q^2) == (sum h_q)^2 for every piece
        double v[8];
        for (int q = 0; q < 4; ++q) { double h = (double)((fp >> (16 * q)) & 0xffffu); v[q] = h; v[4 + q] = h * h; }
        allreduce_sum(v, 8);
        for (int q = 0; q < 4; ++q)
            if ((double)world * v[4 + q] != v[q] * v[q])
                throw std::runtime_error(std::string("the ranks of the sharded sweep have diverged (") + what + "): the host-side state must be bit-identical on all ranks");
    }
    bool device_solver = true;       // false: always leave the eigensolver to the caller (host vectors)

    // ---- Engine::overlap_mpo_left_step / overlap_mpo_right_step -------------------------------------------
    Boundary overlap_mpo_left_step(MPSTensor const& bra_tensor, MPSTensor const& ket_tensor, Boundary const& left, MPOTensor const& mpo,
                                   bool isHermitian = true) override
    {
        return boundary_step(1, bra_tensor, ket_tensor, left, mpo, isHermitian);
    }
    Boundary overlap_mpo_right_step(MPSTensor const& bra_tensor, MPSTensor const& ket_tensor, Boundary const& right, MPOTensor const& mpo,
                                    bool isHermitian = true) override
    {
        return boundary_step(2, bra_tensor, ket_tensor, right, mpo, isHermitian);
    }

    // ---- noise term of the perturbed density matrix (EngineIface::noise_left / noise_right) ----------------------------
    // Planned like a boundary step whose closing products are panel x panel^T tiles (plan_noise_left / plan_noise_right) and
    // executed by qcm_boundary_step into a one-entry "boundary" = the density-matrix blocks.  Quadratic in the W-applied
    // product, hence never sharded: with several ranks every rank computes the whole (cheap: two boundary steps' worth) term.
    block_matrix noise_left(MPSTensor const& mps, Boundary const& left, MPOTensor const& mpo) override { return noise(true, mps, left, mpo); }
    block_matrix noise_right(MPSTensor const& mps, Boundary const& right, MPOTensor const& mpo) override { return noise(false, mps, right, mpo); }
    block_matrix noise(bool left_side, MPSTensor const& mps, Boundary const& in, MPOTensor const& mpo)
    {
        mps.make_left_paired();
        std::shared_ptr<DeviceBoundary> din = mirror(in);
        Clock c0;
        make_room();
        plan::Planner planner(symm, mpo, true, 0, 1, budget);
        plan::Plan P = left_side ? planner.plan_noise_left(desc_of(mps), din->layout) : planner.plan_noise_right(desc_of(mps), din->layout);
        seconds[0] += c0.lap();
        std::shared_ptr<CompiledPlan> cp = compile(P, left_side ? din->layout.total : 0, left_side ? 0 : din->layout.total);
        seconds[1] += c0.lap();
        qcm_array_t out = nullptr;
        qcm_check(qcm_array_alloc(cp->out_boundary.total, &out), "qcm_array_alloc");
        std::vector<double> ket = flatten(mps.data(), cp->ket_elems);
        int rc = qcm_boundary_step(cp->handle, din->arr, ket.data(), ket.data(), out);
        std::vector<double> flat((size_t)cp->out_boundary.total);
        if (rc == 0) rc = qcm_array_download(out, 0, flat.data(), cp->out_boundary.total);
        qcm_array_free(out);
        qcm_check(rc, "noise term (qcm_boundary_step)");
        boundary_flops += cp->flops; ++n_boundary_calls;
        seconds[3] += c0.lap();
        return unflatten(cp->out_boundary.b[0], flat);
    }

    // ---- Engine::diagonal_hamiltonian (abelian/engine.hpp:222-227) --------------------------------------------
    block_matrix diagonal_hamiltonian(Boundary const& left, Boundary const& right, MPOTensor const& mpo, MPSTensor const& x)
    {
        std::shared_ptr<DeviceBoundary> dl = mirror(left), dr = mirror(right);
        plan::Planner planner(symm, mpo, true, 0, 1, budget);
        plan::Plan P = planner.plan_hdiag(desc_of(x), dl->layout, dr->layout);
        std::shared_ptr<CompiledPlan> cp = compile(P, dl->layout.total, dr->layout.total);
        std::vector<double> diag((size_t)cp->out_elems);
        qcm_check(qcm_hdiag(cp->handle, dl->arr, dr->arr, diag.data()), "qcm_hdiag");
        return unflatten(cp->out_tensor, diag);
    }

    // ---- HBM-resident boundary store ----------------------------------------------------------------------
    // device mirror of a boundary (uploaded on first use; boundaries produced by this engine already have one)
    std::shared_ptr<DeviceBoundary> mirror(Boundary const& b)
    {
        if (b.device_mirror) { prefetch(b); return std::static_pointer_cast<DeviceBoundary>(b.device_mirror); }
        if (!b.host_valid) throw std::runtime_error("GpuEngine: boundary has neither host data nor a device mirror");
        std::shared_ptr<DeviceBoundary> d(new DeviceBoundary());
        std::vector<DualIndex> bases(b.aux_dim());
        for (size_t k = 0; k < b.aux_dim(); ++k) bases[k] = b[k].basis();
        d->layout.assign(bases);
        qcm_check(qcm_array_alloc(d->layout.total, &d->arr), "qcm_array_alloc");
        std::vector<double> flat((size_t)d->layout.total);
        for (size_t k = 0; k < b.aux_dim(); ++k)
            for (size_t j = 0; j < b[k].n_blocks(); ++j)
                std::copy(b[k][j].v.begin(), b[k][j].v.end(), flat.begin() + d->layout.b[k].off[j]);
        qcm_check(qcm_array_upload(d->arr, 0, flat.data(), d->layout.total), "qcm_array_upload");
        b.device_mirror = d;
        return d;
    }
    // fetch the dense blocks of a device-resident boundary into the host object
    void download(Boundary& b)
    {
        if (b.host_valid) return;
        std::shared_ptr<DeviceBoundary> d = std::static_pointer_cast<DeviceBoundary>(b.device_mirror);
        if (!d) throw std::runtime_error("GpuEngine::download: no device mirror");
        prefetch(b);
        std::vector<double> flat((size_t)d->layout.total);
        qcm_check(qcm_array_download(d->arr, 0, flat.data(), d->layout.total), "qcm_array_download");
        for (size_t k = 0; k < b.aux_dim(); ++k)
            for (size_t j = 0; j < b.raw(k).n_blocks(); ++j) {
                Matrix& m = b.raw(k)[j];
                m.v.assign(flat.begin() + d->layout.b[k].off[j], flat.begin() + d->layout.b[k].off[j] + m.rows * m.cols);
            }
        b.host_valid = true;
    }
    void fetch(Boundary& b) override { download(b); }
    void clear_cache() { cache.clear(); }
    // plans kept (least recently created dropped first).  A sweep driver sets this to a few times the chain length: once
    // the bond dimensions have settled every (site, direction) finds its plan from the previous sweep.
    void set_cache_capacity(size_t n) { cache_capacity = std::max<size_t>(1, n); while (cache.size() > cache_capacity) cache.pop_back(); }
    size_t cache_hits = 0, cache_misses = 0;
    // host-side time spent in this engine, by kind (seconds): [0] planning (Planner) [1] plan upload (qcm_plan_create)
    // [2] sigma calls (H2D + kernels + D2H) [3] boundary-step calls [4] flatten / unflatten
    double seconds[5] = {0, 0, 0, 0, 0};
    // development aid (printed by the drivers under QCM_DEBUG): [0] allocation of the new boundary [1] flattening of bra / ket in the boundary steps
    double detail_seconds[4] = {0, 0, 0, 0};
    // algorithmic FLOPs (schedule-derived, this rank's share) of the sigma evaluations / boundary steps executed so far
    double sigma_flops = 0, boundary_flops = 0;
    size_t n_sigma_calls = 0, n_boundary_calls = 0;
    std::shared_ptr<CompiledPlan> last_plan() const { return last; }

    // blocks <-> one flat array (blocks back to back in DualIndex order); the copies run block-parallel on the host cores
    static std::vector<double> flatten(block_matrix const& m, int64_t expect)
    {
        std::vector<size_t> off(m.n_blocks() + 1, 0);
        for (size_t k = 0; k < m.n_blocks(); ++k) off[k + 1] = off[k] + m[k].v.size();
        if ((int64_t)off.back() != expect) throw std::runtime_error("GpuEngine: tensor data does not match the planned structure");
        std::vector<double> flat(off.back());
        double* p = flat.data();
#pragma omp parallel for schedule(dynamic, 1) if (off.back() > (size_t)1 << 18)
        for (long k = 0; k < (long)m.n_blocks(); ++k) std::memcpy(p + off[(size_t)k], m[(size_t)k].v.data(), m[(size_t)k].v.size() * sizeof(double));
        return flat;
    }
    static block_matrix unflatten(plan::Layout const& L, std::vector<double> const& flat)
    {
        const size_t nb = L.basis.size();
        std::vector<Matrix> ms(nb);
#pragma omp parallel for schedule(dynamic, 1) if (flat.size() > (size_t)1 << 18)
        for (long k = 0; k < (long)nb; ++k) {
            Matrix m = Matrix::shell(L.basis[(size_t)k].ls, L.basis[(size_t)k].rs);
            m.v.assign(flat.begin() + L.off[(size_t)k], flat.begin() + L.off[(size_t)k] + (size_t)m.rows * m.cols);
            ms[(size_t)k] = std::move(m);
        }
        block_matrix r;
        for (size_t k = 0; k < nb; ++k) r.insert_block(std::move(ms[k]), L.basis[k].lc, L.basis[k].rc);
        return r;
    }
    static plan::TensorDesc desc_of(MPSTensor const& t)
    {
        t.make_left_paired();
        return plan::TensorDesc{t.site_dim(), t.row_dim(), t.col_dim(), t.data().basis()};
    }

private:
    struct Clock
    {
        std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
        double lap() { auto n = std::chrono::steady_clock::now(); double s = std::chrono::duration<double>(n - t).count(); t = n; return s; }
    };
    // (MPO content id, structure of the boundaries, structure of the site tensor(s), kind): everything a plan depends on.
    // The hashes only select the candidate; a hit is confirmed by comparing the full block structures (Witness).
    struct PlanKey
    {
        uint64_t mpo; uint64_t a, b, h; int kind;
        bool operator==(PlanKey const& o) const { return mpo == o.mpo && a == o.a && b == o.b && h == o.h && kind == o.kind; }
    };
    struct Witness
    {
        std::vector<DualIndex> b0, b1;     // block structure of every bond entry of the boundary operand(s)
        Index phys[2], li[2], ri[2]; DualIndex t[2];
        Witness() {}
        Witness(plan::BoundaryLayout const& x, plan::BoundaryLayout const& y, MPSTensor const& bra, MPSTensor const& ket)
        {
            for (auto const& l : x.b) b0.push_back(l.basis);
            for (auto const& l : y.b) b1.push_back(l.basis);
            set(0, bra); set(1, ket);
        }
        void set(int i, MPSTensor const& m) { phys[i] = m.site_dim(); li[i] = m.row_dim(); ri[i] = m.col_dim(); t[i] = m.data().basis(); }
        static bool same(std::vector<DualIndex> const& v, plan::BoundaryLayout const& l)
        {
            if (v.size() != l.b.size()) return false;
            for (size_t k = 0; k < v.size(); ++k) if (!(v[k] == l.b[k].basis)) return false;
            return true;
        }
        bool same_tensor(int i, MPSTensor const& m) const { return phys[i] == m.site_dim() && li[i] == m.row_dim() && ri[i] == m.col_dim() && t[i] == m.data().basis(); }
        bool matches(plan::BoundaryLayout const& x, plan::BoundaryLayout const& y, MPSTensor const& bra, MPSTensor const& ket) const
        {
            return same(b0, x) && same(b1, y) && same_tensor(0, bra) && same_tensor(1, ket);
        }
    };

    Boundary boundary_step(int kind, MPSTensor const& bra_tensor, MPSTensor const& ket_tensor, Boundary const& in, MPOTensor const& mpo, bool isHermitian)
    {
        bra_tensor.make_left_paired(); ket_tensor.make_left_paired();
        std::shared_ptr<DeviceBoundary> din = mirror(in);
        Clock c0;
        PlanKey key{mpo.uid(), din->structure_hash(), structure_hash(bra_tensor), structure_hash(ket_tensor), (isHermitian ? 0 : 3) + kind};
        static const plan::BoundaryLayout no_boundary;
        std::shared_ptr<CompiledPlan> cp = lookup(key, [&](Witness const& w) { return w.matches(din->layout, no_boundary, bra_tensor, ket_tensor); });
        if (!cp) {
            make_room();
            plan::Planner planner(symm, mpo, isHermitian, rank, world, budget);
            plan::Plan P = kind == 1 ? planner.plan_left_step(desc_of(bra_tensor), desc_of(ket_tensor), din->layout)
                                     : planner.plan_right_step(desc_of(bra_tensor), desc_of(ket_tensor), din->layout);
            seconds[0] += c0.lap();
            cp = compile(P, kind == 1 ? din->layout.total : 0, kind == 2 ? din->layout.total : 0);
            seconds[1] += c0.lap();
            remember(key, Witness(din->layout, no_boundary, bra_tensor, ket_tensor), cp);
        }
        last = cp;
        std::shared_ptr<DeviceBoundary> dout(new DeviceBoundary());
        dout->layout = cp->out_boundary;
        dout->recycler = recycler;
        dout->arr = recycler->take(dout->layout.total);
        { double s = c0.lap(); seconds[4] += s; detail_seconds[0] += s; }
        // the sweep drivers move the boundary with bra == ket (one tensor): flattened once
        const bool same = &bra_tensor == &ket_tensor;
        std::vector<double> ket = flatten(ket_tensor.data(), cp->ket_elems), bra;
        if (!same) bra = flatten(bra_tensor.data(), cp->bra_elems);
        { double s = c0.lap(); seconds[4] += s; detail_seconds[1] += s; }
        qcm_check(qcm_boundary_step(cp->handle, din->arr, same ? ket.data() : bra.data(), ket.data(), dout->arr), "qcm_boundary_step");
        boundary_flops += cp->flops; ++n_boundary_calls;
        seconds[3] += c0.lap();
        Boundary ret; ret.resize(dout->layout.aux_dim());
        for (size_t b = 0; b < ret.aux_dim(); ++b) {
            DualIndex const& basis = dout->layout.b[b].basis;
            for (size_t k = 0; k < basis.size(); ++k) ret.raw(b).insert_block(Matrix::shell(basis[k].ls, basis[k].rs), basis[k].lc, basis[k].rc);
        }
        ret.device_mirror = dout;
        ret.host_valid = false;
        return ret;
    }

    std::shared_ptr<CompiledPlan> compile(plan::Plan const& P, int64_t left_elems, int64_t right_elems)
    {
        PlanDescHolder H;
        H.fill(P, left_elems, right_elems);
        qcm_plan_desc& d = H.d;
        const int64_t out_elems = PlanDescHolder::out_elems(P);
        std::shared_ptr<CompiledPlan> cp(new CompiledPlan());
        qcm_check(qcm_plan_create(&d, &cp->handle), "qcm_plan_create");
        cp->out_tensor = P.out_tensor; cp->out_boundary = P.out_boundary;
        cp->ket_elems = P.ket_lp_elems; cp->bra_elems = P.bra_lp_elems; cp->out_elems = out_elems;
        cp->flops = P.flops(); cp->flops_t = P.flops_t; cp->flops_w = P.flops_w; cp->flops_close = P.flops_close; cp->bytes = P.bytes_algorithmic;
        cp->w_elems_read = P.w_elems_read; cp->w_elems_written = P.w_elems_written; cp->w_groups = P.w_groups;
        cp->exec_w = P.exec_w; cp->exec_close = P.exec_close;
        cp->direct_panel_elems = P.direct_panel_elems; cp->w_panel_elems = P.w_panel_elems; cp->skipped_panel_elems = P.skipped_panel_elems;
        cp->n_gemm_tasks = P.n_gemm_tasks; cp->n_axpy_tasks = P.n_axpy_tasks; cp->n_waves = P.waves.size();
        cp->workspace_elems = P.ket_rp_elems + P.t_elems_max + P.tp_elems + P.y_elems_max + P.bra_rp_elems;
        last = cp;
        return cp;
    }

    static uint64_t structure_hash(MPSTensor const& t)
    {
        uint64_t h = 1469598103934665603ull;
        auto mix = [&](uint64_t x) { h ^= x; h *= 1099511628211ull; };
        auto mix_index = [&](Index const& ix) { for (auto const& e : ix) { mix((uint32_t)e.first[0]); mix((uint32_t)e.first[1]); mix((uint32_t)e.first[2]); mix(e.second); } mix(0xABCDu); };
        mix_index(t.site_dim()); mix_index(t.row_dim()); mix_index(t.col_dim());
        for (auto const& b : t.data().basis()) { mix((uint32_t)b.lc[0]); mix((uint32_t)b.lc[1]); mix((uint32_t)b.lc[2]); mix((uint32_t)b.rc[0]); mix((uint32_t)b.rc[1]); mix((uint32_t)b.rc[2]); mix(b.ls); mix(b.rs); }
        return h;
    }
    struct CacheEntry { PlanKey key; Witness witness; std::shared_ptr<CompiledPlan> plan; };
    template <class Confirm> std::shared_ptr<CompiledPlan> lookup(PlanKey const& k, Confirm confirm)
    {
        for (auto& e : cache) if (e.key == k && confirm(e.witness)) return e.plan;
        return std::shared_ptr<CompiledPlan>();
    }
    // the oldest plan leaves the cache before a new one is built: its task arrays (0.5 GB at cfg3) go back to the device
    // pool and the new plan's arrays take their place instead of growing the pool
    void make_room() { while (cache.size() + 1 > cache_capacity && !cache.empty()) cache.pop_back(); }
    void remember(PlanKey const& k, Witness w, std::shared_ptr<CompiledPlan> const& cp)
    {
        cache.push_front(CacheEntry{k, std::move(w), cp});
        while (cache.size() > cache_capacity) cache.pop_back();
    }

    SymmKind symm;
    int rank, world;
    int64_t budget;
    std::list<CacheEntry> cache;
    size_t cache_capacity = 4;
    std::vector<qcm_array_t> vec_pool; int64_t vec_pool_n = 0;      // solver vectors, reused from site to site
    qcm_array_t host_xfer = nullptr;                                // staging array of allreduce_sum
    std::shared_ptr<CompiledPlan> last;
};

} // namespace qcm
