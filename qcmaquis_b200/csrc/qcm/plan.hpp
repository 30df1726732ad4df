// Flattened contraction schedule ("plan") for the B200 kernels.
//
// The reference rebuilds its task structure inside every site_hamil2 call (per-b OpenMP loop bodies,
// SU2::task_capsule / micro_task lists: contractions/non-abelian/apply_op.hpp:103-209, site_hamil.hpp:181).
// Here the same loops are walked ONCE per (site, direction) on the host and emitted as three flat task
// families that the device executes for every Davidson iteration:
//   * panel copies      (left<->right pairing reshapes, reshapes.h:177-223,289-335)
//   * grouped GEMMs     (one output block = a list of K-segments; block_matrix_algorithms.h:48-162,
//                        non-abelian/gemm.hpp:48-204, charge_gemm apply_op.hpp:255-266)
//   * gather-axpy panels (the W application: alps_detail.hpp:189-224, micro_kernels.hpp:19-198), with the
//                        SU2 couplings (gsl_coupling.h:177-204) and Hermitian phases folded into scalars.
// Block structure rules (which output blocks exist, their sizes, the descending charge order) follow the
// reference loops literally so that the result has the same DualIndex as the CPU implementation.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <tuple>
#include "mpo.hpp"
#include "mps.hpp"
#include <cstdint>
#include <map>
#include <numeric>
#include <unordered_map>
#ifdef _OPENMP
#include <parallel/algorithm>
#endif

namespace qcm { namespace plan {

enum Buf : int { BUF_KET_LP = 0, BUF_KET_RP, BUF_LEFT, BUF_RIGHT, BUF_T, BUF_TP, BUF_Y, BUF_OUT, BUF_BRA_LP, BUF_BRA_RP, BUF_COUNT };

struct Ref { int32_t buf; int64_t off; };
// wall-clock per planning phase, printed when QCM_PLAN_TIMING is set (development aid)
struct PhaseTimer
{
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    bool on = getenv("QCM_PLAN_TIMING") != nullptr;
    void lap(const char* what)
    {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "  [plan] %-28s %8.3f s\n", what, std::chrono::duration<double>(n - t).count());
        t = n;
    }
};
struct CopyTask { Ref src, dst; int32_t rows, cols, lds, ldd; };
struct Seg { Ref A, B; int32_t lda, ldb, m, n, k, ta, tb; double alpha; };   // op(A) is m x k, op(B) is k x n
struct Out { Ref C; int32_t ldc, m, n, seg_begin, seg_end; };
struct AxpySrc { Ref src; int32_t lds; double coef; };
struct AxpyDst { Ref dst; int32_t ldd, rows, cols; };       // its sources: AxpyList::lists at the same index

struct GemmList { std::vector<Out> outs; std::vector<Seg> segs; };
// destination panels with their source lists (each list sorted by (buf, off), equal sources merged); the lists are
// moved in from the panel collection, never copied -- at cfg3 they hold more than 10^8 (source, destination) pairs
struct AxpyList { std::vector<AxpyDst> dsts; std::vector<std::vector<AxpySrc>> lists; };

// The W application in the form the device executes: destination panels that are fed by the SAME set of source
// panels (typical for quantum-chemistry MPOs: all bond terms that differ only by an integral value) are grouped,
// and a group is evaluated as one small dense product  dst[e, d] = sum_u src_u[e] * coef[u][d]  over the panel
// elements e.  Every source panel is then read once per group instead of once per (source, destination) pair.
struct WSrc { Ref src; int32_t lds; };
struct WDst { Ref dst; int32_t ldd; };
struct WGroup { int32_t rows, cols, n_src, n_dst, ng, cls, src_begin, dst_begin; int64_t coef_begin; };   // coef[coef_begin + u*ng + d]; cls 0: DMMA product (u padded to 8), 1: FMA stream
// allocator whose resize() leaves new elements uninitialised: the coefficient tables (10^8 doubles at cfg3) are zeroed
// and filled group by group in parallel instead of by one serial value-initialisation
template <class T> struct DefaultInitAlloc : std::allocator<T>
{
    template <class U> struct rebind { typedef DefaultInitAlloc<U> other; };
    DefaultInitAlloc() = default;
    template <class U> DefaultInitAlloc(DefaultInitAlloc<U> const&) {}
    template <class U> void construct(U* p) { ::new ((void*)p) U; }
    template <class U, class... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(std::forward<A>(a)...); }
};
struct WList
{
    std::vector<WGroup> groups; std::vector<WSrc> srcs; std::vector<WDst> dsts; std::vector<double, DefaultInitAlloc<double>> coefs;
    int64_t elems_read = 0, elems_written = 0;   // panel elements moved by the grouped form
};

struct Wave
{
    GemmList t_gemm;      // step 1 products needed by this wave (single-use bonds) -> BUF_T
    AxpyList w_apply;     // step 2 -> BUF_Y (planner-internal list, cleared once grouped)
    WList w_groups;       // step 2 as executed
    GemmList close_gemm;  // step 3 -> BUF_OUT (accumulating for sigma, plain for boundary steps)
    int64_t y_elems = 0;  // BUF_Y elements this wave uses (compact multi-source panels; all written by the W pass before they are read)
    int64_t t_elems = 0;
    // Exchange wave (world > 1, always waves[0]): its W pass writes this rank's PARTIAL sums of the destination panels whose
    // sources are spread over several ranks into BUF_Y[0, world * x_chunk) -- the same layout on every rank -- the region is
    // reduce-scattered (rank r receives the complete sums of chunk r) while the other waves run, and the closing products
    // of chunk `rank` (close_gemm of this wave) are executed at the end of the plan.
    int64_t x_chunk = 0;
    bool x_zero = false;  // some exchanged panels get no contribution from this rank: the region is zeroed first
};

struct Layout   // block offsets of one block matrix inside a flat buffer
{
    DualIndex basis;
    std::vector<int64_t> off;
    int64_t total = 0;
    void assign(DualIndex const& b, int64_t base = 0)
    {
        basis = b; off.resize(b.size()); total = 0;
        for (size_t k = 0; k < b.size(); ++k) { off[k] = base + total; total += (int64_t)b[k].ls * (int64_t)b[k].rs; }
    }
};

struct BoundaryLayout
{
    std::vector<Layout> b;
    int64_t total = 0;
    void assign(std::vector<DualIndex> const& bases)
    {
        b.resize(bases.size()); total = 0;
        for (size_t i = 0; i < bases.size(); ++i) { b[i].assign(bases[i], total); total += b[i].total; }
    }
    size_t aux_dim() const { return b.size(); }
};

struct Plan
{
    int kind = 0;                          // 0 sigma, 1 left step, 2 right step, 3 diagonal_hamiltonian
    int rank = 0, world = 1;               // the shard this plan computes (world 1: the whole contraction)
    std::vector<CopyTask> pre_copies;      // pairing reshapes of ket / bra
    GemmList persistent_t;                 // step 1 products of multi-use bonds -> BUF_TP
    std::vector<Wave> waves;
    Layout out_tensor;                     // sigma (left paired)
    BoundaryLayout out_boundary;           // boundary steps
    bool accumulate_out = false;
    int64_t ket_lp_elems = 0, ket_rp_elems = 0, bra_lp_elems = 0, bra_rp_elems = 0;
    int64_t tp_elems = 0, t_elems_max = 0, y_elems_max = 0;
    double flops_t = 0, flops_w = 0, flops_close = 0;          // algorithmic (reference schedule)
    double exec_w = 0, exec_close = 0;                         // what the device executes after panel routing (before tile padding)
    int64_t direct_panel_elems = 0, w_panel_elems = 0, skipped_panel_elems = 0;
    int64_t w_elems_read = 0, w_elems_written = 0, w_groups = 0;   // traffic of the grouped W application (panel elements)
    int64_t bytes_algorithmic = 0;         // 8*(sum|L_b| + sum|R_b| + 2|psi|) resp. boundary-step analogue
    size_t n_gemm_tasks = 0, n_axpy_tasks = 0;
    double flops() const { return flops_t + flops_w + flops_close; }
};

// a (possibly transposed / scaled) view of one stored boundary entry
struct VBlock { Charge lc, rc; int32_t ls, rs; int64_t off; int32_t ld; int32_t trans; double scale; };
struct VView
{
    std::vector<VBlock> blocks;   // sorted descending by (lc, rc)
    DualIndex basis;
    void build(Layout const& L, bool transposed, std::vector<double> const& scales)
    {
        blocks.clear(); basis = DualIndex();
        for (size_t k = 0; k < L.basis.size(); ++k) {
            QnBlock const& q = L.basis[k];
            VBlock v;
            double sc = scales.empty() ? 1. : scales[k];
            if (!transposed) v = VBlock{q.lc, q.rc, (int32_t)q.ls, (int32_t)q.rs, L.off[k], (int32_t)q.ls, 0, sc};
            else v = VBlock{q.rc, q.lc, (int32_t)q.rs, (int32_t)q.ls, L.off[k], (int32_t)q.ls, 1, sc};
            size_t i = basis.insert(QnBlock(v.lc, v.rc, v.ls, v.rs));
            blocks.insert(blocks.begin() + i, v);
        }
    }
};

struct TensorDesc   // what the planner needs to know about an MPS tensor (no data)
{
    Index phys_i, left_i, right_i;
    DualIndex lp_basis;   // left-paired block structure of the data the caller will pass
};

class Planner
{
public:
    Planner(SymmKind s, MPOTensor const& mpo_, bool isHermitian_, int rank_ = 0, int world_ = 1, int64_t ws_budget_elems = (int64_t)1 << 30)
        : symm(s), su2_(is_su2(s)), mpo(mpo_), isHermitian(isHermitian_), rank(rank_), world(world_), budget(ws_budget_elems) {}
    // only the output block structure is wanted (Plan::out_tensor / out_boundary); no tasks are emitted
    bool structure_only = false;

    // -----------------------------------------------------------------------------------------------------
    // sigma = H_eff psi    (abelian/site_hamil.hpp:23-90, non-abelian/site_hamil.hpp:57-147)
    Plan plan_sigma(TensorDesc const& ket, BoundaryLayout const& left, BoundaryLayout const& right)
    {
        Plan P; P.kind = 0; P.accumulate_out = true; P.rank = rank; P.world = world;
        TensorDesc const& bra = ket;
        Layout ket_lp; ket_lp.assign(ket.lp_basis);
        Layout ket_rp = plan_left_to_right(ket, ket_lp, BUF_KET_LP, BUF_KET_RP, P.pre_copies);
        P.ket_lp_elems = ket_lp.total; P.ket_rp_elems = ket_rp.total;

        Index const& physical_i = ket.phys_i;
        Index const& left_i = bra.left_i;
        Index right_i = ket.right_i;
        Index out_left_i = physical_i * left_i;
        if (su2_) { Index right_i_bra = bra.right_i; common_subset(out_left_i, right_i_bra); }
        else common_subset(out_left_i, right_i);                       // abelian trims BOTH (site_hamil.hpp:42)
        ProductBasis out_left_pb(physical_i, left_i);
        ProductBasis in_right_pb(physical_i, right_i, true);
        Index indexForTrim = ket_rp.basis.left_basis();                // bra == ket, right paired

        PhaseTimer pt;
        setup_t_left(P, left, ket_rp, indexForTrim);
        if (su2_) build_lbtm_tables(ket_rp, physical_i, right_i, out_left_i, in_right_pb, out_left_pb);
        pt.lap("setup_t_left");

        // output structure: emulate the per-b2 products and their match_and_add_block reduction
        block_struct sigma_struct;
        struct Pending { size_t b2; Layout y; std::vector<size_t> t_rows; std::vector<std::vector<uint32_t>> block_bonds; };
        std::vector<Pending> pend(mpo.col_dim());
        int n_b2 = (int)mpo.col_dim();
        // pass A: structure of every Y[b2], the bonds that contribute to it and (sharded plans) the bonds behind every block
        auto walk = [&](size_t b2, DualIndex& ybasis, std::vector<YTask>& tasks, std::vector<size_t>& rows, YOpts const& o) {
            if (su2_) y_struct_su2_lbtm(b2, ket_rp, right_i, out_left_i, in_right_pb, out_left_pb, ybasis, tasks, rows, o);
            else y_struct_abelian_lbtm(b2, ket_rp.basis, ket_rp.basis, right_i, out_left_i, in_right_pb, out_left_pb, ybasis, tasks, rows, o);
        };
#pragma omp parallel for schedule(dynamic, 4)
        for (int b2 = 0; b2 < n_b2; ++b2) {
            Pending& pd = pend[b2]; pd.b2 = (size_t)b2;
            DualIndex ybasis; std::vector<YTask> none;
            YOpts o; o.emit_tasks = false; o.block_bonds = world > 1 ? &pd.block_bonds : nullptr;
            walk((size_t)b2, ybasis, none, pd.t_rows, o);
            pd.y.assign(ybasis);
        }
        pt.lap("y_struct (all b2)");
        for (size_t b2 = 0; b2 < mpo.col_dim(); ++b2) {
            DualIndex const& ybasis = pend[b2].y.basis;
            // closing product structure (decides which sigma blocks exist, on every rank identically)
            VView rv = right_view(right, b2);
            for (size_t k = 0; k < ybasis.size(); ++k) {
                if (su2_) {
                    Charge al = ybasis[k].lc, ar = ybasis[k].rc;
                    size_t mb = rv.basis.position(ar, al);
                    if (mb == rv.basis.size()) continue;
                    if (!out_left_i.has(al)) continue;   // the num_ops>3 post filter is implied for the plain kernel too
                    sigma_struct.add(al, al, ybasis[k].ls, rv.blocks[mb].rs);
                } else {
                    Charge ar = ybasis[k].rc;
                    for (auto it = rv.basis.left_lower_bound(ar); it != rv.basis.end() && it->lc == ar; ++it)
                        sigma_struct.add(ybasis[k].lc, it->rc, ybasis[k].ls, it->rs);
                }
            }
        }
        P.out_tensor.assign(sigma_struct.basis);
        if (structure_only) return P;
        pt.lap("sigma structure");

        // emit waves over this rank's share of b2
        std::vector<char> own = shard_sources(P, pend.size(), [&](size_t i) { return 2.0 * estimate_cost(pend[i].y, right, pend[i].b2); },
                                              [&](size_t i) -> std::vector<size_t> const& { return pend[i].t_rows; });
        // Y blocks of output i that take part in a closing product (same test as the emission below)
        auto sigma_match = [&](DualIndex const& ybasis, VView const& rv, size_t k, std::vector<size_t>& m) {
            QnBlock const& yb = ybasis[k];
            if (su2_) {
                size_t mb = rv.basis.position(yb.rc, yb.lc);
                if (mb == rv.basis.size() || !out_left_i.has(yb.lc)) return;
                m.push_back(mb);
            } else
                for (auto it = rv.basis.left_lower_bound(yb.rc); it != rv.basis.end() && it->lc == yb.rc; ++it) m.push_back(it - rv.basis.begin());
        };
        // row units of Y block o of output i: one panel per (physical state, left sector) -- the same panels the W tasks address
        auto units_of = [&](size_t i, size_t o, std::vector<XPanel>& out) {
            QnBlock const& yb = pend[i].y.basis[o];
            for (size_t p = 0; p < physical_i.size(); ++p) {
                size_t l = left_i.position(fuse(yb.lc, -physical_i[p].first));
                if (l == left_i.size() || !out_left_pb.has(physical_i[p].first, left_i[l].first)) continue;
                const int32_t off = (int32_t)out_left_pb(physical_i[p].first, left_i[l].first), ls = (int32_t)left_i[l].second;
                for (size_t c = 0; c < physical_i[p].second; ++c) out.push_back(XPanel{(uint32_t)i, (uint32_t)o, off + (int32_t)c * ls, 0, ls, (int32_t)yb.rs, -1, 0});
            }
        };
        Exchange xch = plan_exchange(pend.size(), [&](size_t i) -> std::vector<std::vector<uint32_t>> const& { return pend[i].block_bonds; },
            [&](size_t i) {
                VView rv = right_view(right, pend[i].b2);
                std::vector<char> u(pend[i].y.basis.size(), 0);
                for (size_t k = 0; k < u.size(); ++k) { std::vector<size_t> m; sigma_match(pend[i].y.basis, rv, k, m); u[k] = !m.empty(); }
                return u;
            }, units_of);
        auto closes_some = [&](size_t i) { if (!xch.active()) return false; for (uint32_t q : xch.by_output[i]) if (xch.xp[q].chunk == rank) return true; return false; };
        std::vector<char> books(pend.size(), 1);
        for (size_t i = 0; i < pend.size(); ++i) books[i] = filter_owned(own, pend[i].t_rows);
        pt.lap("sharding + persistent T");
        std::vector<char> idle(pend.size(), 0);
        for (size_t i = 0; i < pend.size(); ++i) idle[i] = world > 1 && pend[i].t_rows.empty() && !books[i] && !closes_some(i);
        // pass B, output by output: the tasks of the bonds this rank owns, into the calling thread's scratch buffer
        YOpts emit_own; emit_own.own = world > 1 ? &own : nullptr;
        auto gen_tasks = [&](size_t j, std::vector<YTask>& scratch) -> std::vector<YTask> const& {
            scratch.clear();
            DualIndex yb; std::vector<size_t> rows;
            walk(pend[j].b2, yb, scratch, rows, emit_own);
            return scratch;
        };
        PanelCache pcache;
        Wave xw;
        const int64_t y0 = xch.active() ? xch.chunk * world : 0;      // BUF_Y[0, y0) is the exchange region
        Wave cur; int64_t cur_y = y0, cur_t = 0;
        auto flush = [&]() {
            if (cur.w_apply.dsts.empty() && cur.close_gemm.outs.empty() && cur.t_gemm.outs.empty()) return;
            cur.y_elems = cur_y; cur.t_elems = cur_t;
            P.y_elems_max = std::max(P.y_elems_max, cur_y); P.t_elems_max = std::max(P.t_elems_max, cur_t);
            pt.lap("emission (panels, closing)");
            merge_outputs(cur.close_gemm); merge_outputs(cur.t_gemm);
            pt.lap("merge_outputs");
            group_axpy(P, cur.w_apply, cur.w_groups);
            pt.lap("group_axpy");
            P.waves.push_back(std::move(cur));
            cur = Wave(); cur_y = y0; cur_t = 0;
        };
        for (size_t i = 0; i < pend.size(); ++i) {
            Pending& pd = pend[i];
            if (idle[i]) continue;
            int64_t need_t = 0;
            for (size_t b1 : pd.t_rows) if (!t_persistent[b1]) need_t += t_layout_size(b1);
            if ((cur_y - y0 + cur_t) > 0 && cur_y - y0 + cur_t + pd.y.total + need_t > budget) flush();
            // step 1 for single-use rows consumed here
            const int64_t t_begin = cur_t;
            std::map<size_t, Layout> tl;
            for (size_t b1 : pd.t_rows) {
                if (t_persistent[b1]) { tl[b1] = tp_layout[b1]; continue; }
                Layout L; L.assign(t_basis[b1], cur_t);
                emit_t_gemm(P, cur.t_gemm, b1, L, BUF_T, ket_rp);
                cur_t += L.total; tl[b1] = L;
            }
            // steps 2 + 3, panel by panel: a destination panel of Y[b2] (rows (phys_out, lc, col) of one Y block) is
            // either a single scaled T panel -- then the closing product reads T directly and Y is never materialised --
            // or a sum over several T panels, formed by the W kernel in a compact Y region.  Each row unit of a sigma
            // block is its own output with its own K-segment list over (b2, panel).
            DualIndex const& ybasis = pd.y.basis;
            VView rv = right_view(right, pd.b2);
            std::vector<std::vector<size_t>> match(ybasis.size());
            for (size_t k = 0; k < ybasis.size(); ++k) {
                QnBlock const& yb = ybasis[k];
                sigma_match(ybasis, rv, k, match[k]);
                for (size_t mb : match[k])
                    if (books[i] && P.out_tensor.basis.has(yb.lc, su2_ ? yb.lc : rv.blocks[mb].rc)) { P.flops_close += 2.0 * yb.ls * rv.blocks[mb].rs * yb.rs; P.n_gemm_tasks++; }
            }
            // exchanged panels of this output whose complete sums arrive on this rank: closed from the exchange region
            if (xch.active())
                for (uint32_t q : xch.by_output[i]) {
                    XPanel const& x = xch.xp[q];
                    if (x.chunk != rank) continue;
                    QnBlock const& yb = ybasis[x.o];
                    for (size_t mb : match[x.o])
                        emit_close(P, xw.close_gemm, P.out_tensor, yb.lc, su2_ ? yb.lc : rv.blocks[mb].rc, Ref{BUF_Y, x.off}, x.rows, 0, x.rows, x.cols, 1., rv.blocks[mb], BUF_RIGHT,
                                   x.dst_row, 0);
                }
            PanelCounts pcnt;
            std::vector<Panel> panels = cached_panels(pcache, pend.size(), i, t_begin,
                [&](size_t j) -> std::vector<size_t> const& { return pend[j].t_rows; }, gen_tasks,
                [&](size_t j) { return idle[j] != 0; }, tl, pcnt);
            P.flops_w += pcnt.flops_w; P.n_axpy_tasks += pcnt.n_axpy;
            for (Panel& pn : panels) {
                if (match[pn.o].empty()) { P.skipped_panel_elems += (int64_t)pn.rows * pn.cols; continue; }
                const long xi = xch.find(i, pn.o, pn.dst_row, pn.dst_col);
                if (xi >= 0) { if (!pn.srcs.empty()) place_exchanged(P, xw, pn, xch.xp[(size_t)xi]); continue; }
                PanelRef pr;
                if (!place_panel(P, cur.w_apply, pn, cur_y, pr)) continue;
                QnBlock const& yb = ybasis[pn.o];
                for (size_t mb : match[pn.o])
                    emit_close(P, cur.close_gemm, P.out_tensor, yb.lc, su2_ ? yb.lc : rv.blocks[mb].rc, pr.A, pr.lda, 0, pn.rows, pn.cols, pr.alpha, rv.blocks[mb], BUF_RIGHT,
                               pn.dst_row, 0);
            }
        }
        flush();
        finish_exchange(P, xw, xch);
        merge_outputs(P.persistent_t);
        P.bytes_algorithmic = 8 * (left.total + right.total + 2 * ket_lp.total);
        return P;
    }

    // -----------------------------------------------------------------------------------------------------
    // L'[b2] = Y[b2]^T conj(bra)      (common/move_boundary.hpp:128-187)
    Plan plan_left_step(TensorDesc const& bra, TensorDesc const& ket, BoundaryLayout const& left)
    {
        Plan P; P.kind = 1; P.accumulate_out = false; P.rank = rank; P.world = world;
        Layout ket_lp; ket_lp.assign(ket.lp_basis);
        Layout ket_rp = plan_left_to_right(ket, ket_lp, BUF_KET_LP, BUF_KET_RP, P.pre_copies);
        Layout bra_lp; bra_lp.assign(bra.lp_basis);
        std::vector<CopyTask> dummy;
        Layout bra_rp = plan_left_to_right(bra, bra_lp, BUF_BRA_LP, BUF_BRA_RP, dummy);   // structure only
        P.ket_lp_elems = ket_lp.total; P.ket_rp_elems = ket_rp.total; P.bra_lp_elems = bra_lp.total;

        Index braBasis = bra_rp.basis.left_basis();
        Index const& left_i = bra.left_i;
        Index right_i = ket.right_i;
        Index bra_right_i = bra.right_i;
        Index out_left_i = bra.phys_i * left_i;
        common_subset(out_left_i, bra_right_i);
        ProductBasis out_left_pb(bra.phys_i, left_i);
        ProductBasis in_right_pb(ket.phys_i, right_i, true);

        setup_t_left(P, left, ket_rp, braBasis);
        if (su2_) build_lbtm_tables(ket_rp, ket.phys_i, right_i, out_left_i, in_right_pb, out_left_pb);

        size_t loop_max = mpo.col_dim();
        struct Pending { size_t b2; DualIndex y; std::vector<YTask> ytasks; std::vector<size_t> t_rows; DualIndex out; };
        std::vector<Pending> pend(loop_max);
        std::vector<DualIndex> out_bases(loop_max);
#pragma omp parallel for schedule(dynamic, 4)
        for (long b2l = 0; b2l < (long)loop_max; ++b2l) {
            size_t b2 = (size_t)b2l;
            Pending& pd = pend[b2]; pd.b2 = b2;
            if (mpo.herm_info.right_skip(b2) && isHermitian) continue;
            if (su2_) y_struct_su2_lbtm(b2, ket_rp, right_i, out_left_i, in_right_pb, out_left_pb, pd.y, pd.ytasks, pd.t_rows);
            else y_struct_abelian_lbtm(b2, ket_rp.basis, bra_rp.basis, right_i, out_left_i, in_right_pb, out_left_pb, pd.y, pd.ytasks, pd.t_rows);
            // gemm(transpose(Y), bra_lp, ret[b2], spin)
            block_struct os;
            int spin_f = su2_ ? mpo.right_spin(b2).get() : -1;
            DualIndex yt = pd.y.transposed();
            for (size_t k = 0; k < yt.size(); ++k)
                for (auto it = bra_lp.basis.left_lower_bound(yt[k].rc); it != bra_lp.basis.end() && it->lc == yt[k].rc; ++it) {
                    if (spin_f != -1 && !su2::triangle(spin(yt[k].lc), spin_f, spin(it->rc))) continue;
                    os.add(yt[k].lc, it->rc, yt[k].ls, it->rs);
                }
            out_bases[b2] = os.basis;
        }
        P.out_boundary.assign(out_bases);
        if (structure_only) return P;

        static const std::vector<size_t> no_rows;
        std::vector<char> own = shard_sources(P, loop_max, [&](size_t b2) {
                double c = 0; for (size_t k = 0; k < pend[b2].y.size(); ++k) c += 2.0 * (double)pend[b2].y[k].ls * pend[b2].y[k].rs * pend[b2].y[k].rs; return c; },
            [&](size_t b2) -> std::vector<size_t> const& { return (mpo.herm_info.right_skip(b2) && isHermitian) ? no_rows : pend[b2].t_rows; });
        std::vector<char> books(pend.size(), 1);
        for (size_t i = 0; i < pend.size(); ++i) books[i] = filter_owned(own, pend[i].ytasks, pend[i].t_rows);
        PanelCache pcache;
        Wave cur; int64_t cur_y = 0, cur_t = 0;
        auto flush = [&]() {
            if (cur.w_apply.dsts.empty() && cur.close_gemm.outs.empty() && cur.t_gemm.outs.empty()) return;
            cur.y_elems = cur_y; cur.t_elems = cur_t;
            P.y_elems_max = std::max(P.y_elems_max, cur_y); P.t_elems_max = std::max(P.t_elems_max, cur_t);
            merge_outputs(cur.close_gemm); merge_outputs(cur.t_gemm);
            group_axpy(P, cur.w_apply, cur.w_groups);
            P.waves.push_back(std::move(cur));
            cur = Wave(); cur_y = 0; cur_t = 0;
        };
        for (size_t b2 = 0; b2 < loop_max; ++b2) {
            Pending& pd = pend[b2];
            if ((world > 1 && pd.ytasks.empty() && !books[b2]) || (mpo.herm_info.right_skip(b2) && isHermitian)) continue;
            Layout ytmp; ytmp.assign(pd.y);
            int64_t need_t = 0;
            for (size_t b1 : pd.t_rows) if (!t_persistent[b1]) need_t += t_layout_size(b1);
            if ((cur_y + cur_t) > 0 && cur_y + cur_t + ytmp.total + need_t > budget) flush();
            const int64_t t_begin = cur_t;
            std::map<size_t, Layout> tl;
            for (size_t b1 : pd.t_rows) {
                if (t_persistent[b1]) { tl[b1] = tp_layout[b1]; continue; }
                Layout L; L.assign(t_basis[b1], cur_t);
                emit_t_gemm(P, cur.t_gemm, b1, L, BUF_T, ket_rp);
                cur_t += L.total; tl[b1] = L;
            }
            int spin_f = su2_ ? mpo.right_spin(b2).get() : -1;
            Layout const& ol = P.out_boundary.b[b2];
            // gemm(transpose(Y), bra_lp): the panels of a Y block are row ranges = K ranges of the closing product
            std::vector<std::vector<size_t>> match(pd.y.size());
            for (size_t k = 0; k < pd.y.size(); ++k) {
                QnBlock const& yb = pd.y[k];
                for (auto it = bra_lp.basis.left_lower_bound(yb.lc); it != bra_lp.basis.end() && it->lc == yb.lc; ++it) {
                    if (spin_f != -1 && !su2::triangle(spin(yb.rc), spin_f, spin(it->rc))) continue;
                    match[k].push_back(it - bra_lp.basis.begin());
                    if (books[b2]) { P.flops_close += 2.0 * yb.rs * it->rs * yb.ls; P.n_gemm_tasks++; }
                }
            }
            PanelCounts pcnt;
            std::vector<Panel> panels = cached_panels(pcache, pend.size(), b2, t_begin,
                [&](size_t j) -> std::vector<size_t> const& { return pend[j].t_rows; },
                [&](size_t j, std::vector<YTask>&) -> std::vector<YTask> const& { return pend[j].ytasks; },
                [&](size_t j) { return (world > 1 && pend[j].ytasks.empty() && !books[j]) || (mpo.herm_info.right_skip(j) && isHermitian); }, tl, pcnt);
            P.flops_w += pcnt.flops_w; P.n_axpy_tasks += pcnt.n_axpy;
            for (Panel& pn : panels) {
                if (match[pn.o].empty()) { P.skipped_panel_elems += (int64_t)pn.rows * pn.cols; continue; }
                PanelRef pr;
                if (!place_panel(P, cur.w_apply, pn, cur_y, pr)) continue;
                QnBlock const& yb = pd.y[pn.o];
                for (size_t mb : match[pn.o]) {
                    QnBlock const& q = bra_lp.basis[mb];
                    VBlock bv{q.lc, q.rc, (int32_t)q.ls, (int32_t)q.rs, bra_lp.off[mb], (int32_t)q.ls, 0, 1.};
                    emit_close(P, cur.close_gemm, ol, yb.rc, q.rc, pr.A, pr.lda, 1, pn.cols, pn.rows, pr.alpha, bv, BUF_BRA_LP, 0, pn.dst_row);
                }
            }
        }
        flush();
        merge_outputs(P.persistent_t);
        P.bytes_algorithmic = 8 * (left.total + ket_lp.total + bra_lp.total + P.out_boundary.total);
        return P;
    }

    // -----------------------------------------------------------------------------------------------------
    // R'[b1] = Y'[b1] conj(bra)^T     (common/move_boundary.hpp:189-229)
    Plan plan_right_step(TensorDesc const& bra, TensorDesc const& ket, BoundaryLayout const& right)
    {
        Plan P; P.kind = 2; P.accumulate_out = false; P.rank = rank; P.world = world;
        Layout ket_lp; ket_lp.assign(ket.lp_basis);
        Layout bra_lp; bra_lp.assign(bra.lp_basis);
        Layout bra_rp = plan_left_to_right(bra, bra_lp, BUF_BRA_LP, BUF_BRA_RP, P.pre_copies);
        P.ket_lp_elems = ket_lp.total; P.bra_lp_elems = bra_lp.total; P.bra_rp_elems = bra_rp.total;

        Index const& physical_i = ket.phys_i;
        Index right_i = bra.right_i;
        Index left_i = ket.left_i, out_right_i = adjoin(physical_i) * right_i, bra_left_i = bra.left_i;
        Index indexForTrim = bra_lp.basis.right_basis();
        common_subset(out_right_i, bra_left_i);
        ProductBasis in_left_pb(physical_i, left_i);
        ProductBasis out_right_pb(physical_i, right_i, true);

        setup_t_right(P, right, ket_lp, indexForTrim);

        size_t loop_max = mpo.row_dim();
        struct Pending { DualIndex y; std::vector<YTask> ytasks; std::vector<size_t> t_cols; };
        std::vector<Pending> pend(loop_max);
        std::vector<DualIndex> out_bases(loop_max);
        DualIndex bra_rp_t = bra_rp.basis.transposed();
#pragma omp parallel for schedule(dynamic, 4)
        for (long b1l = 0; b1l < (long)loop_max; ++b1l) {
            size_t b1 = (size_t)b1l;
            Pending& pd = pend[b1];
            if (mpo.herm_info.left_skip(b1) && isHermitian) continue;
            if (su2_) y_struct_su2_rbtm(b1, ket_lp.basis, left_i, out_right_i, in_left_pb, out_right_pb, pd.y, pd.ytasks, pd.t_cols);
            else y_struct_abelian_rbtm(b1, left_i, out_right_i, in_left_pb, out_right_pb, pd.y, pd.ytasks, pd.t_cols);
            block_struct os;
            int spin_f = su2_ ? mpo.left_spin(b1).get() : -1;
            for (size_t k = 0; k < pd.y.size(); ++k)
                for (auto it = bra_rp_t.left_lower_bound(pd.y[k].rc); it != bra_rp_t.end() && it->lc == pd.y[k].rc; ++it) {
                    if (spin_f != -1 && !su2::triangle(spin(pd.y[k].lc), spin_f, spin(it->rc))) continue;
                    os.add(pd.y[k].lc, it->rc, pd.y[k].ls, it->rs);
                }
            out_bases[b1] = os.basis;
        }
        P.out_boundary.assign(out_bases);
        if (structure_only) return P;

        static const std::vector<size_t> no_cols;
        std::vector<char> own = shard_sources(P, loop_max, [&](size_t b1) {
                double c = 0; for (size_t k = 0; k < pend[b1].y.size(); ++k) c += 2.0 * (double)pend[b1].y[k].ls * pend[b1].y[k].ls * pend[b1].y[k].rs; return c; },
            [&](size_t b1) -> std::vector<size_t> const& { return (mpo.herm_info.left_skip(b1) && isHermitian) ? no_cols : pend[b1].t_cols; });
        std::vector<char> books(pend.size(), 1);
        for (size_t i = 0; i < pend.size(); ++i) books[i] = filter_owned(own, pend[i].ytasks, pend[i].t_cols);
        PanelCache pcache;
        Wave cur; int64_t cur_y = 0, cur_t = 0;
        auto flush = [&]() {
            if (cur.w_apply.dsts.empty() && cur.close_gemm.outs.empty() && cur.t_gemm.outs.empty()) return;
            cur.y_elems = cur_y; cur.t_elems = cur_t;
            P.y_elems_max = std::max(P.y_elems_max, cur_y); P.t_elems_max = std::max(P.t_elems_max, cur_t);
            merge_outputs(cur.close_gemm); merge_outputs(cur.t_gemm);
            group_axpy(P, cur.w_apply, cur.w_groups);
            P.waves.push_back(std::move(cur));
            cur = Wave(); cur_y = 0; cur_t = 0;
        };
        // bra right paired, transposed view: block (rc, lc) of the stored (lc, rc)
        VView brt; { Layout l = bra_rp; brt.build(l, true, std::vector<double>()); }
        for (size_t b1 = 0; b1 < loop_max; ++b1) {
            Pending& pd = pend[b1];
            if ((world > 1 && pd.ytasks.empty() && !books[b1]) || (mpo.herm_info.left_skip(b1) && isHermitian)) continue;
            Layout ytmp; ytmp.assign(pd.y);
            int64_t need_t = 0;
            for (size_t b2 : pd.t_cols) if (!t_persistent[b2]) need_t += t_layout_size(b2);
            if ((cur_y + cur_t) > 0 && cur_y + cur_t + ytmp.total + need_t > budget) flush();
            const int64_t t_begin = cur_t;
            std::map<size_t, Layout> tl;
            for (size_t b2 : pd.t_cols) {
                if (t_persistent[b2]) { tl[b2] = tp_layout[b2]; continue; }
                Layout L; L.assign(t_basis[b2], cur_t);
                emit_t_gemm_right(P, cur.t_gemm, b2, L, BUF_T, ket_lp);
                cur_t += L.total; tl[b2] = L;
            }
            int spin_f = su2_ ? mpo.left_spin(b1).get() : -1;
            Layout const& ol = P.out_boundary.b[b1];
            // gemm(Y, transpose(bra_rp)): the panels of a Y block are column ranges = K ranges of the closing product
            std::vector<std::vector<size_t>> match(pd.y.size());
            for (size_t k = 0; k < pd.y.size(); ++k) {
                QnBlock const& yb = pd.y[k];
                for (auto it = brt.basis.left_lower_bound(yb.rc); it != brt.basis.end() && it->lc == yb.rc; ++it) {
                    if (spin_f != -1 && !su2::triangle(spin(yb.lc), spin_f, spin(it->rc))) continue;
                    match[k].push_back(it - brt.basis.begin());
                    if (books[b1]) { P.flops_close += 2.0 * yb.ls * it->rs * yb.rs; P.n_gemm_tasks++; }
                }
            }
            PanelCounts pcnt;
            std::vector<Panel> panels = cached_panels(pcache, pend.size(), b1, t_begin,
                [&](size_t j) -> std::vector<size_t> const& { return pend[j].t_cols; },
                [&](size_t j, std::vector<YTask>&) -> std::vector<YTask> const& { return pend[j].ytasks; },
                [&](size_t j) { return (world > 1 && pend[j].ytasks.empty() && !books[j]) || (mpo.herm_info.left_skip(j) && isHermitian); }, tl, pcnt);
            P.flops_w += pcnt.flops_w; P.n_axpy_tasks += pcnt.n_axpy;
            for (Panel& pn : panels) {
                if (match[pn.o].empty()) { P.skipped_panel_elems += (int64_t)pn.rows * pn.cols; continue; }
                PanelRef pr;
                if (!place_panel(P, cur.w_apply, pn, cur_y, pr)) continue;
                QnBlock const& yb = pd.y[pn.o];
                for (size_t mb : match[pn.o])
                    emit_close(P, cur.close_gemm, ol, yb.lc, brt.blocks[mb].rc, pr.A, pr.lda, 0, pn.rows, pn.cols, pr.alpha, brt.blocks[mb], BUF_BRA_RP, 0, pn.dst_col);
            }
        }
        flush();
        merge_outputs(P.persistent_t);
        P.bytes_algorithmic = 8 * (right.total + ket_lp.total + bra_lp.total + P.out_boundary.total);
        return P;
    }

    // -----------------------------------------------------------------------------------------------------
    // diag(H_eff) in the left-paired layout of x (abelian/h_diag.hpp:41-168, non-abelian/h_diag.hpp:19-155):
    //   out[(sigma, l, i), c] = sum over (b1, b2, op, diagonal W entries) alfa * L[b1](l,l)_ii * R[b2](c,c)_cc
    // The reference forms, per b2, a matrix with identical columns and scales its columns by the right diagonal.  Here
    // the sum over b1 is a vector per (block, b2) -- V[:, b2], formed by the W kernels from strided views of the stored
    // left blocks (a diagonal is a 1 x m "panel" with leading dimension ld + 1) -- the right diagonals are gathered
    // into DR[b2, :] by the panel-copy kernel, and the sum over b2 is ONE dense product V * DR per output block on the
    // grouped GEMM.  Same quirks as the reference: abelian uses op(0) without the entry's scale; only stored boundary
    // blocks are read (Hermitian-skipped bonds contribute nothing).  Every rank computes the whole (cheap) result.
    Plan plan_hdiag(TensorDesc const& x, BoundaryLayout const& left, BoundaryLayout const& right)
    {
        Plan P; P.kind = 3; P.accumulate_out = true;
        Index const& physical_i = x.phys_i;
        Index right_i = x.right_i, out_left_i = physical_i * x.left_i;
        common_subset(out_left_i, right_i);
        ProductBasis out_left_pb(physical_i, x.left_i);
        Index const& left_i = x.left_i;
        // output structure: block (c, c) for every right charge with some non-empty column b2 whose right boundary has (c, c)
        DualIndex ob;
        std::vector<std::vector<size_t>> slots(right_i.size());        // per right_i block: the b2 that contribute
        std::vector<std::vector<size_t>> slot_rblock(right_i.size());
        for (size_t b2 = 0; b2 < right.aux_dim() && b2 < mpo.col_dim(); ++b2) {
            if (mpo.col_begin(b2) == mpo.col_end(b2)) continue;
            for (size_t block = 0; block < right_i.size(); ++block) {
                Charge c = right_i[block].first;
                size_t rb = right.b[b2].basis.position(c, c);
                if (rb == right.b[b2].basis.size()) continue;
                if (!ob.has(c, c)) ob.insert(QnBlock(c, c, out_left_i[block].second, right_i[block].second));
                slots[block].push_back(b2); slot_rblock[block].push_back(rb);
            }
        }
        P.out_tensor.assign(ob);
        // V and DR workspaces
        std::vector<int64_t> voff(right_i.size(), 0), droff(right_i.size(), 0);
        int64_t vtot = 0, drtot = 0;
        for (size_t block = 0; block < right_i.size(); ++block) {
            voff[block] = vtot; droff[block] = drtot;
            vtot += (int64_t)out_left_i[block].second * (int64_t)slots[block].size();
            drtot += (int64_t)right_i[block].second * (int64_t)slots[block].size();
        }
        Wave wave;
        std::map<std::tuple<size_t, size_t, int64_t>, size_t> dst_index;      // (block, slot, row offset) -> destination
        std::vector<std::vector<AxpySrc>> dst_srcs;
        std::vector<AxpyDst> dsts;
        for (size_t block = 0; block < right_i.size(); ++block) {
            Charge in_charge = right_i[block].first;
            const int64_t rows = (int64_t)out_left_i[block].second, cols = (int64_t)right_i[block].second;
            for (size_t kk = 0; kk < slots[block].size(); ++kk) {
                const size_t b2 = slots[block][kk];
                // right diagonal -> DR[kk, :]
                {
                    Layout const& rl = right.b[b2];
                    size_t rb = slot_rblock[block][kk];
                    P.pre_copies.push_back(CopyTask{Ref{BUF_RIGHT, rl.off[rb]}, Ref{BUF_TP, droff[block] + (int64_t)kk * cols}, 1, (int32_t)std::min<int64_t>(cols, rl.basis[rb].ls),
                                                    (int32_t)rl.basis[rb].ls + 1, 1});
                }
                for (size_t e = mpo.col_begin(b2); e < mpo.col_end(b2); ++e) {
                    size_t b1 = mpo.row_of(e);
                    if (b1 >= left.aux_dim()) continue;
                    auto const& access = mpo.at_entry(e);
                    size_t n_ops = su2_ ? access.size() : 1;
                    for (size_t op_index = 0; op_index < n_ops; ++op_index) {
                        SiteOperator const& W = mpo.op(access[op_index].first);
                        int a = 0, k = 0, ap = 0;
                        if (su2_) { a = mpo.left_spin(b1).get(); k = W.spin().get(); ap = mpo.right_spin(b2).get(); }
                        for (size_t s = 0; s < physical_i.size(); ++s) {
                            Charge phys_charge = physical_i[s].first;
                            size_t l = left_i.position(fuse(in_charge, -phys_charge));
                            if (l == left_i.size()) continue;
                            Charge lc = left_i[l].first;
                            Layout const& ll = left.b[b1];
                            size_t l_block = ll.basis.position(lc, lc);
                            if (l_block == ll.basis.size()) continue;
                            const int64_t m_l = (int64_t)left_i[l].second;
                            const int64_t left_offset = (int64_t)out_left_pb(phys_charge, lc);
                            for (size_t w_block = 0; w_block < W.basis().size(); ++w_block) {
                                Charge phys_in = W.basis().left_charge(w_block), phys_out = W.basis().right_charge(w_block);
                                if (!(phys_charge == phys_in) || !(phys_in == phys_out)) continue;
                                double couplings[2] = {1., 1.};
                                if (su2_) {
                                    int i = spin(lc), ip = spin(in_charge), j = spin(lc), jp = spin(in_charge);
                                    int two_sp = std::abs(i - ip), two_s = std::abs(j - jp);
                                    double prefactor = std::sqrt((ip + 1.) * (j + 1.) / ((i + 1.) * (jp + 1.))) * access[op_index].second;
                                    couplings[0] = prefactor * su2::mod_coupling(j, two_s, jp, a, k, ap, i, two_sp, ip);
                                    couplings[1] = prefactor * su2::mod_coupling(j, 2, jp, a, k, ap, i, 2, ip);
                                }
                                for (int sp = W.sparse_ptr[w_block]; sp < W.sparse_ptr[w_block + 1]; ++sp) {
                                    SparseEntry const& en = W.sparse[sp];
                                    if (en.row != en.col) continue;
                                    const double alfa = su2_ ? en.coefficient * couplings[en.row_spin == 2 ? 1 : 0] : en.coefficient;
                                    const int64_t row_off = left_offset + (int64_t)en.row * m_l;
                                    auto key = std::make_tuple(block, kk, row_off);
                                    auto it = dst_index.find(key);
                                    if (it == dst_index.end()) {
                                        it = dst_index.emplace(key, dsts.size()).first;
                                        AxpyDst d; d.dst = Ref{BUF_Y, voff[block] + (int64_t)kk * rows + row_off}; d.ldd = 1; d.rows = 1; d.cols = (int32_t)m_l;
                                        dsts.push_back(d); dst_srcs.emplace_back();
                                    }
                                    // the diagonal of the stored block: element i at off + i * (ld + 1)
                                    dst_srcs[it->second].push_back(AxpySrc{Ref{BUF_LEFT, ll.off[l_block]}, (int32_t)ll.basis[l_block].ls + 1, alfa});
                                    P.flops_w += 2.0 * (double)m_l * (double)cols;     // the reference updates every column
                                    P.n_axpy_tasks++;
                                }
                            }
                        }
                    }
                }
            }
            // closing product of this block: out = V (rows x K) * DR (K x cols)
            if (!slots[block].empty() && ob.has(in_charge, in_charge)) {
                size_t o = P.out_tensor.basis.position(in_charge, in_charge);
                Out out; out.C = Ref{BUF_OUT, P.out_tensor.off[o]}; out.ldc = (int32_t)rows; out.m = (int32_t)rows; out.n = (int32_t)cols;
                out.seg_begin = (int32_t)wave.close_gemm.segs.size();
                wave.close_gemm.segs.push_back(Seg{Ref{BUF_Y, voff[block]}, Ref{BUF_TP, droff[block]}, (int32_t)rows, (int32_t)cols, (int32_t)rows, (int32_t)cols,
                                                   (int32_t)slots[block].size(), 0, 1, 1.0});
                out.seg_end = (int32_t)wave.close_gemm.segs.size();
                wave.close_gemm.outs.push_back(out);
                P.flops_close += (double)rows * cols * (double)slots[block].size();          // the reference's column scaling + add
                P.exec_close += 2.0 * (double)rows * cols * (double)slots[block].size();
            }
        }
        // sources of equal (block, offset) inside one destination are merged; destinations -> W groups
        for (size_t d = 0; d < dsts.size(); ++d) {
            auto& v = dst_srcs[d];
            std::stable_sort(v.begin(), v.end(), [](AxpySrc const& x1, AxpySrc const& x2) { return std::tie(x1.src.buf, x1.src.off) < std::tie(x2.src.buf, x2.src.off); });
            size_t o = 0;
            for (size_t i = 0; i < v.size(); ++i) {
                if (o && v[o - 1].src.buf == v[i].src.buf && v[o - 1].src.off == v[i].src.off && v[o - 1].lds == v[i].lds) v[o - 1].coef += v[i].coef;
                else v[o++] = v[i];
            }
            v.resize(o);
            P.exec_w += 2.0 * dsts[d].cols * (double)v.size();
            wave.w_apply.dsts.push_back(dsts[d]);
            wave.w_apply.lists.push_back(std::move(v));
        }
        wave.y_elems = vtot; wave.t_elems = 0;
        P.y_elems_max = vtot; P.tp_elems = drtot;
        group_axpy(P, wave.w_apply, wave.w_groups);
        P.waves.push_back(std::move(wave));
        P.bytes_algorithmic = 8 * (left.total + right.total);
        return P;
    }


    // -----------------------------------------------------------------------------------------------------
    // Noise term of the perturbed density matrix (C/common/prediction.hpp:34-47, twositetensor.hpp:192-219):
    //   left:  sum over b2 of Y[b2] Y[b2]^T   with Y = left_boundary_tensor_mpo  (C/common/move_boundary.hpp:68-95): blocks (lc, lc)
    //   right: sum over b1 of Y'[b1]^T Y'[b1] with Y' = right_boundary_tensor_mpo (:97-126):                          blocks (rc, rc)
    // Y is the W-applied product of the boundary steps WITHOUT the closing bra: same step-1 products, same W passes; the closing
    // products become panel x panel^T tiles of the density-matrix blocks.  The result is quadratic in Y, so the plan is never
    // sharded (every rank computes the whole term); it runs through qcm_boundary_step as a boundary with a single entry.
    Plan plan_noise_left(TensorDesc const& ket, BoundaryLayout const& left)
    {
        if (world > 1) throw std::runtime_error("plan_noise_left: the noise term is not sharded; plan it with world = 1");
        Plan P; P.kind = 1; P.accumulate_out = true;
        Layout ket_lp; ket_lp.assign(ket.lp_basis);
        Layout ket_rp = plan_left_to_right(ket, ket_lp, BUF_KET_LP, BUF_KET_RP, P.pre_copies);
        P.ket_lp_elems = ket_lp.total; P.ket_rp_elems = ket_rp.total; P.bra_lp_elems = 0;
        Index const& physical_i = ket.phys_i;
        Index const& left_i = ket.left_i;
        Index right_i = ket.right_i;
        Index out_left_i = physical_i * left_i;                         // not trimmed (move_boundary.hpp:82-83)
        ProductBasis out_left_pb(physical_i, left_i);
        ProductBasis in_right_pb(physical_i, right_i, true);
        setup_t_left(P, left, ket_rp, left_i);
        if (su2_) build_lbtm_tables(ket_rp, physical_i, right_i, out_left_i, in_right_pb, out_left_pb);
        const size_t loop_max = mpo.col_dim();
        struct Pending { DualIndex y; std::vector<YTask> ytasks; std::vector<size_t> t_rows; };
        std::vector<Pending> pend(loop_max);
#pragma omp parallel for schedule(dynamic, 4)
        for (long b2l = 0; b2l < (long)loop_max; ++b2l) {
            Pending& pd = pend[(size_t)b2l];
            if (su2_) y_struct_su2_lbtm((size_t)b2l, ket_rp, right_i, out_left_i, in_right_pb, out_left_pb, pd.y, pd.ytasks, pd.t_rows);
            else y_struct_abelian_lbtm((size_t)b2l, ket_rp.basis, ket_rp.basis, right_i, out_left_i, in_right_pb, out_left_pb, pd.y, pd.ytasks, pd.t_rows);
        }
        block_struct os;
        for (size_t b2 = 0; b2 < loop_max; ++b2) for (size_t k = 0; k < pend[b2].y.size(); ++k) os.add(pend[b2].y[k].lc, pend[b2].y[k].lc, pend[b2].y[k].ls, pend[b2].y[k].ls);
        P.out_boundary.assign(std::vector<DualIndex>(1, os.basis));
        if (structure_only) return P;
        noise_emit(P, pend.size(), [&](size_t i) -> std::vector<size_t> const& { return pend[i].t_rows; },
                   [&](size_t i) -> std::vector<YTask> const& { return pend[i].ytasks; }, [&](size_t i) -> DualIndex const& { return pend[i].y; },
                   [&](Plan& PP, GemmList& gl, size_t b, Layout const& tl, int buf) { emit_t_gemm(PP, gl, b, tl, buf, ket_rp); }, true);
        P.bytes_algorithmic = 8 * (left.total + ket_lp.total + P.out_boundary.total);
        return P;
    }
    Plan plan_noise_right(TensorDesc const& ket, BoundaryLayout const& right)
    {
        if (world > 1) throw std::runtime_error("plan_noise_right: the noise term is not sharded; plan it with world = 1");
        Plan P; P.kind = 2; P.accumulate_out = true;
        Layout ket_lp; ket_lp.assign(ket.lp_basis);
        P.ket_lp_elems = ket_lp.total; P.bra_lp_elems = 0; P.bra_rp_elems = 0;
        Index const& physical_i = ket.phys_i;
        Index right_i = ket.right_i;
        Index left_i = ket.left_i, out_right_i = adjoin(physical_i) * right_i;     // not trimmed (move_boundary.hpp:110-111)
        ProductBasis in_left_pb(physical_i, left_i);
        ProductBasis out_right_pb(physical_i, right_i, true);
        // MPSBoundaryProduct(mps, right, mpo) "without index" trims by the tensor's ROW index (boundary_times_mps.hpp:258-261)
        setup_t_right(P, right, ket_lp, left_i);
        const size_t loop_max = mpo.row_dim();
        struct Pending { DualIndex y; std::vector<YTask> ytasks; std::vector<size_t> t_cols; };
        std::vector<Pending> pend(loop_max);
#pragma omp parallel for schedule(dynamic, 4)
        for (long b1l = 0; b1l < (long)loop_max; ++b1l) {
            Pending& pd = pend[(size_t)b1l];
            if (su2_) y_struct_su2_rbtm((size_t)b1l, ket_lp.basis, left_i, out_right_i, in_left_pb, out_right_pb, pd.y, pd.ytasks, pd.t_cols);
            else y_struct_abelian_rbtm((size_t)b1l, left_i, out_right_i, in_left_pb, out_right_pb, pd.y, pd.ytasks, pd.t_cols);
        }
        block_struct os;
        for (size_t b1 = 0; b1 < loop_max; ++b1) for (size_t k = 0; k < pend[b1].y.size(); ++k) os.add(pend[b1].y[k].rc, pend[b1].y[k].rc, pend[b1].y[k].rs, pend[b1].y[k].rs);
        P.out_boundary.assign(std::vector<DualIndex>(1, os.basis));
        if (structure_only) return P;
        noise_emit(P, pend.size(), [&](size_t i) -> std::vector<size_t> const& { return pend[i].t_cols; },
                   [&](size_t i) -> std::vector<YTask> const& { return pend[i].ytasks; }, [&](size_t i) -> DualIndex const& { return pend[i].y; },
                   [&](Plan& PP, GemmList& gl, size_t b, Layout const& tl, int buf) { emit_t_gemm_right(PP, gl, b, tl, buf, ket_lp); }, false);
        P.bytes_algorithmic = 8 * (right.total + ket_lp.total + P.out_boundary.total);
        return P;
    }

private:
    struct block_struct   // accumulates an output block structure with match_and_add_block growth semantics
    {
        DualIndex basis;
        void add(Charge const& lc, Charge const& rc, size_t ls, size_t rs)
        {
            size_t p = basis.position(lc, rc);
            if (p == basis.size()) basis.insert(QnBlock(lc, rc, ls, rs));
            else { basis[p].ls = std::max(basis[p].ls, ls); basis[p].rs = std::max(basis[p].rs, rs); }
        }
    };
    // What a walk over the W-application loops of one output index is asked to produce.  The block structure of Y (and the
    // order in which its blocks are created) always comes out.  The planner walks every output twice: once for the structure,
    // the bonds that really contribute and -- for sharded plans -- which bonds touch which Y block (pass A, cheap), and once,
    // output by output inside the emission, for the tasks of the bonds this rank owns (pass B); the task list of an output
    // lives in a per-thread scratch buffer only until its panels are collected.
    struct YOpts
    {
        bool emit_tasks;                                     // false: structure and used bonds only
        std::vector<char> const* own;                        // tasks only for these bonds
        std::vector<std::vector<uint32_t>>* block_bonds;     // per Y block (final position): the bonds that touch it, ascending
        YOpts() : emit_tasks(true), own(nullptr), block_bonds(nullptr) {}
    };
    // one W-application contribution: dst panel in Y block `o`, source panel in T[bt] block `t_block`
    struct YTask { size_t o; int32_t dst_row, dst_col, rows, cols; size_t bt, t_block; int32_t src_row, src_col; double coef; };

    // ---- reshape structure [(phys,left),right] -> [left,(-phys,right)] (reshapes.h:177-223) as panel copies
    Layout plan_left_to_right(TensorDesc const& t, Layout const& lp, int src_buf, int dst_buf, std::vector<CopyTask>& copies)
    {
        ProductBasis in_left(t.phys_i, t.left_i);
        ProductBasis out_right(t.phys_i, t.right_i, true);
        DualIndex rb;
        struct C { size_t block; Charge ol, orc; int32_t in_off, out_off, sdim, ldim, rdim; };
        std::vector<C> cs;
        for (size_t block = 0; block < lp.basis.size(); ++block) {
            size_t r = t.right_i.position(lp.basis[block].rc);
            if (r == t.right_i.size()) continue;
            Charge in_r = t.right_i[r].first;
            for (size_t s = 0; s < t.phys_i.size(); ++s) {
                size_t l = t.left_i.position(fuse(lp.basis[block].lc, -t.phys_i[s].first));
                if (l == t.left_i.size()) continue;
                Charge ol = t.left_i[l].first, orc = fuse(-t.phys_i[s].first, in_r);
                if (!rb.has(ol, orc)) rb.insert(QnBlock(ol, orc, t.left_i[l].second, out_right.size(orc)));
                cs.push_back(C{block, ol, orc, (int32_t)in_left(t.phys_i[s].first, t.left_i[l].first), (int32_t)out_right(t.phys_i[s].first, in_r),
                               (int32_t)t.phys_i[s].second, (int32_t)t.left_i[l].second, (int32_t)t.right_i[r].second});
            }
        }
        Layout rp; rp.assign(rb);
        for (auto const& c : cs) {
            size_t ob = rp.basis.position(c.ol, c.orc);
            int32_t ld_in = (int32_t)lp.basis[c.block].ls, ld_out = (int32_t)rp.basis[ob].ls;
            for (int32_t ss = 0; ss < c.sdim; ++ss)
                copies.push_back(CopyTask{Ref{src_buf, lp.off[c.block] + c.in_off + ss * c.ldim}, Ref{dst_buf, rp.off[ob] + (int64_t)(c.out_off + ss * c.rdim) * ld_out},
                                          c.ldim, c.rdim, ld_in, ld_out});
        }
        return rp;
    }

    // ---- Hermitian-aware views of boundary entries (common/boundary_times_mps.hpp:19-43,163-209,266-361)
    std::vector<double> conj_phases(DualIndex const& b, size_t k, bool left, bool forward) const
    {
        if (!su2_) return std::vector<double>();
        int S = left ? mpo.left_spin(k).get() : mpo.right_spin(k).get();
        std::vector<double> ret(b.size());
        for (size_t i = 0; i < b.size(); ++i) {
            double scale = su2::conjugate_correction(spin(b[i].lc), spin(b[i].rc), S);
            if (forward) scale *= left ? mpo.herm_info.left_phase(mpo.herm_info.left_conj(k)) : mpo.herm_info.right_phase(mpo.herm_info.right_conj(k));
            else scale *= left ? mpo.herm_info.left_phase(k) : mpo.herm_info.right_phase(k);
            ret[i] = scale;
        }
        return ret;
    }
    // A operand of step 1 (left): transpose(left[b1]) or conjugate(left[conj]) with conjugate phases
    VView left_view(BoundaryLayout const& left, size_t b1) const
    {
        VView v;
        if (mpo.herm_info.left_skip(b1) && isHermitian) {
            Layout const& src = left.b[mpo.herm_info.left_conj(b1)];
            v.build(src, false, conj_phases(src.basis, b1, true, false));
        } else
            v.build(left.b[b1], true, std::vector<double>());
        return v;
    }
    // B operand for right-side products: right[b2] or adjoint(right[conj]) with phases computed on the adjoint view
    VView right_view(BoundaryLayout const& right, size_t b2) const
    {
        VView v;
        if (mpo.herm_info.right_skip(b2) && isHermitian) {
            Layout const& src = right.b[mpo.herm_info.right_conj(b2)];
            v.build(src, true, std::vector<double>());
            std::vector<double> ph = conj_phases(v.basis, b2, false, true);
            for (size_t i = 0; i < v.blocks.size(); ++i) v.blocks[i].scale = ph.empty() ? 1. : ph[i];
        } else
            v.build(right.b[b2], false, std::vector<double>());
        return v;
    }

    // ---- step 1 structure, left: T[b1] = op(L[b1]) * psi_rp  (gemm_trim_left)
    void setup_t_left(Plan& P, BoundaryLayout const& left, Layout const& ket_rp, Index const& ref_left_basis)
    {
        size_t B = left.aux_dim();
        t_basis.assign(B, DualIndex()); t_views.assign(B, VView()); t_persistent.assign(B, 0); tp_layout.assign(B, Layout());
        Index B_left_basis = ket_rp.basis.left_basis();
        for (size_t b1 = 0; b1 < B; ++b1) {
            t_views[b1] = left_view(left, b1);
            VView const& A = t_views[b1];
            DualIndex tb;
            for (size_t k = 0; k < A.blocks.size(); ++k) {
                if (!su2_) {
                    size_t mb = B_left_basis.position(A.blocks[k].rc);
                    if (mb == ket_rp.basis.size()) continue;
                    if (!ref_left_basis.has(A.blocks[k].lc)) continue;
                    tb.insert(QnBlock(A.blocks[k].lc, ket_rp.basis[mb].rc, A.blocks[k].ls, ket_rp.basis[mb].rs));
                } else {
                    if (!ref_left_basis.has(A.blocks[k].lc)) continue;
                    for (auto it = ket_rp.basis.left_lower_bound(A.blocks[k].rc); it != ket_rp.basis.end() && it->lc == A.blocks[k].rc; ++it)
                        if (!tb.has(A.blocks[k].lc, it->rc)) tb.insert(QnBlock(A.blocks[k].lc, it->rc, A.blocks[k].ls, it->rs));
                }
            }
            t_basis[b1] = tb;
        }
        // multi-use bonds are computed once up front and stay resident for all waves (emit_persistent, after sharding)
        for (size_t b1 = 0; b1 < B; ++b1) {
            if (b1 < mpo.row_dim() && mpo.num_row_non_zeros(b1) == 1) continue;
            if (b1 >= mpo.row_dim() || mpo.num_row_non_zeros(b1) == 0) continue;
            t_persistent[b1] = 1;
        }
        t_is_left = true; t_ket = ket_rp;
    }
    // step-1 products of the multi-use bonds this rank's share of the output index needs: laid out in BUF_TP and emitted
    void emit_persistent(Plan& P, std::vector<char> const& needed)
    {
        int64_t off = 0;
        for (size_t b = 0; b < t_basis.size(); ++b) {
            if (!t_persistent[b] || !needed[b]) continue;
            tp_layout[b].assign(t_basis[b], off);
            off += tp_layout[b].total;
            if (t_is_left) emit_t_gemm(P, P.persistent_t, b, tp_layout[b], BUF_TP, t_ket);
            else emit_t_gemm_right(P, P.persistent_t, b, tp_layout[b], BUF_TP, t_ket);
        }
        P.tp_elems = off;
    }
    // FLOPs of the step-1 product of bond b (dry run of the emission)
    double t_cost(size_t b)
    {
        Plan tmp; GemmList gl; Layout L; L.assign(t_basis[b], 0);
        if (t_is_left) emit_t_gemm(tmp, gl, b, L, BUF_T, t_ket); else emit_t_gemm_right(tmp, gl, b, L, BUF_T, t_ket);
        return tmp.flops_t;
    }
    void emit_t_gemm(Plan& P, GemmList& gl, size_t b1, Layout const& tl, int buf, Layout const& ket_rp)
    {
        VView const& A = t_views[b1];
        Index B_left_basis = ket_rp.basis.left_basis();
        for (size_t k = 0; k < A.blocks.size(); ++k) {
            VBlock const& a = A.blocks[k];
            auto one = [&](size_t mb) {
                size_t cb = tl.basis.position(a.lc, ket_rp.basis[mb].rc);
                if (cb == tl.basis.size()) return;
                Out o; o.C = Ref{buf, tl.off[cb]}; o.ldc = (int32_t)tl.basis[cb].ls; o.m = a.ls; o.n = (int32_t)ket_rp.basis[mb].rs;
                o.seg_begin = (int32_t)gl.segs.size();
                gl.segs.push_back(Seg{Ref{BUF_LEFT, a.off}, Ref{BUF_KET_RP, ket_rp.off[mb]}, a.ld, (int32_t)ket_rp.basis[mb].ls, o.m, o.n, a.rs, a.trans, 0, a.scale});
                o.seg_end = (int32_t)gl.segs.size();
                gl.outs.push_back(o);
                P.flops_t += 2.0 * o.m * o.n * a.rs; P.n_gemm_tasks++;
            };
            if (!tl.basis.left_has(a.lc)) continue;
            if (!su2_) { size_t mb = B_left_basis.position(a.rc); if (mb != ket_rp.basis.size()) one(mb); }
            else for (auto it = ket_rp.basis.left_lower_bound(a.rc); it != ket_rp.basis.end() && it->lc == a.rc; ++it) one(it - ket_rp.basis.begin());
        }
    }

    // ---- step 1 structure, right: T[b2] = psi_lp * op(R[b2])  (gemm_trim_right)
    void setup_t_right(Plan& P, BoundaryLayout const& right, Layout const& ket_lp, Index const& ref_right_basis)
    {
        size_t B = right.aux_dim();
        t_basis.assign(B, DualIndex()); t_views.assign(B, VView()); t_persistent.assign(B, 0); tp_layout.assign(B, Layout());
        Index A_right_basis = ket_lp.basis.right_basis();
        for (size_t b2 = 0; b2 < B; ++b2) {
            t_views[b2] = right_view(right, b2);
            VView const& Bv = t_views[b2];
            DualIndex tb;
            if (!su2_) {
                for (size_t k = 0; k < Bv.blocks.size(); ++k) {
                    size_t mb = A_right_basis.position(Bv.blocks[k].lc);
                    if (mb == ket_lp.basis.size()) continue;
                    if (!ref_right_basis.has(Bv.blocks[k].rc)) continue;
                    tb.insert(QnBlock(ket_lp.basis[mb].lc, Bv.blocks[k].rc, ket_lp.basis[mb].ls, Bv.blocks[k].rs));
                }
            } else {
                for (size_t a = 0; a < ket_lp.basis.size(); ++a)
                    for (auto it = Bv.basis.left_lower_bound(ket_lp.basis[a].rc); it != Bv.basis.end() && it->lc == ket_lp.basis[a].rc; ++it) {
                        if (!ref_right_basis.has(it->rc)) continue;
                        if (!tb.has(ket_lp.basis[a].lc, it->rc)) tb.insert(QnBlock(ket_lp.basis[a].lc, it->rc, ket_lp.basis[a].ls, it->rs));
                    }
            }
            t_basis[b2] = tb;
        }
        for (size_t b2 = 0; b2 < B; ++b2) {
            if (b2 >= mpo.col_dim() || mpo.num_col_non_zeros(b2) <= 1) continue;
            t_persistent[b2] = 1;
        }
        t_is_left = false; t_ket = ket_lp;
        ket_lp_for_right = ket_lp;
    }
    void emit_t_gemm_right(Plan& P, GemmList& gl, size_t b2, Layout const& tl, int buf, Layout const& ket_lp)
    {
        VView const& Bv = t_views[b2];
        Index A_right_basis = ket_lp.basis.right_basis();
        auto one = [&](size_t a, size_t k) {
            VBlock const& b = Bv.blocks[k];
            size_t cb = tl.basis.position(ket_lp.basis[a].lc, b.rc);
            if (cb == tl.basis.size()) return;
            Out o; o.C = Ref{buf, tl.off[cb]}; o.ldc = (int32_t)tl.basis[cb].ls; o.m = (int32_t)ket_lp.basis[a].ls; o.n = b.rs;
            o.seg_begin = (int32_t)gl.segs.size();
            gl.segs.push_back(Seg{Ref{BUF_KET_LP, ket_lp.off[a]}, Ref{BUF_RIGHT, b.off}, (int32_t)ket_lp.basis[a].ls, b.ld, o.m, o.n, (int32_t)ket_lp.basis[a].rs, 0, b.trans, b.scale});
            o.seg_end = (int32_t)gl.segs.size();
            gl.outs.push_back(o);
            P.flops_t += 2.0 * o.m * o.n * ket_lp.basis[a].rs; P.n_gemm_tasks++;
        };
        if (!su2_) {
            for (size_t k = 0; k < Bv.blocks.size(); ++k) {
                size_t mb = A_right_basis.position(Bv.blocks[k].lc);
                if (mb == ket_lp.basis.size()) continue;
                one(mb, k);
            }
        } else
            for (size_t a = 0; a < ket_lp.basis.size(); ++a)
                for (auto it = Bv.basis.left_lower_bound(ket_lp.basis[a].rc); it != Bv.basis.end() && it->lc == ket_lp.basis[a].rc; ++it)
                    one(a, it - Bv.basis.begin());
    }
    int64_t t_layout_size(size_t b) const
    {
        int64_t s = 0;
        for (size_t k = 0; k < t_basis[b].size(); ++k) s += (int64_t)t_basis[b][k].ls * (int64_t)t_basis[b][k].rs;
        return s;
    }

    static Charge delta_of(DualIndex const& b) { return fuse(b.right_charge(0), -b.left_charge(0)); }

    // ---- step 2 structure, abelian lbtm (abelian/apply_op.hpp:23-139)
    void y_struct_abelian_lbtm(size_t b2, DualIndex const& /*ket_basis*/, DualIndex const& /*bra_basis*/, Index const& right_i, Index const& out_left_i,
                               ProductBasis const& in_right_pb, ProductBasis const& out_left_pb,
                               DualIndex& ret, std::vector<YTask>& tasks, std::vector<size_t>& t_rows, YOpts const& opts = YOpts())
    {
        std::vector<std::pair<std::pair<Charge, Charge>, uint32_t>> touches;      // (block, bond), only when asked for
        for (size_t e = mpo.col_begin(b2); e < mpo.col_end(b2); ++e) {   // allocate
            size_t b1 = mpo.row_of(e);
            DualIndex const& T = t_basis[b1];
            if (T.size() == 0) continue;
            for (auto const& term : mpo.at_entry(e)) {
                SiteOperator const& W = mpo.op(term.first);
                if (W.n_blocks() == 0) continue;
                Charge total_delta = fuse(delta_of(W.basis()), -delta_of(T));
                for (size_t r = 0; r < right_i.size(); ++r) {
                    Charge out_r = right_i[r].first, out_l = fuse(out_r, total_delta);
                    if (!out_left_i.has(out_l)) continue;
                    if (!ret.has(out_l, out_r)) ret.insert(QnBlock(out_l, out_r, out_left_i.size_of_block(out_l), right_i[r].second));
                    if (opts.block_bonds) touches.push_back(std::make_pair(std::make_pair(out_l, out_r), (uint32_t)b1));
                }
            }
        }
        if (opts.block_bonds) {
            opts.block_bonds->assign(ret.size(), std::vector<uint32_t>());
            for (auto const& t : touches) { auto& v = (*opts.block_bonds)[ret.position(t.first.first, t.first.second)]; if (v.empty() || v.back() != t.second) v.push_back(t.second); }
        }
        if (structure_only) return;
        for (size_t e = mpo.col_begin(b2); e < mpo.col_end(b2); ++e) {   // execute
            size_t b1 = mpo.row_of(e);
            DualIndex const& T = t_basis[b1];
            if (T.size() == 0) continue;
            const bool emit = opts.emit_tasks && (!opts.own || (*opts.own)[b1]);
            if (!emit && opts.emit_tasks) continue;                      // pass B: somebody else's bond
            bool used = false;
            for (auto const& term : mpo.at_entry(e)) {
                SiteOperator const& W = mpo.op(term.first);
                if (W.n_blocks() == 0) continue;
                Charge T_delta = delta_of(T);
                Charge total_delta = fuse(delta_of(W.basis()), -T_delta);
                for (size_t r = 0; r < right_i.size() && (emit || !used); ++r) {
                    Charge out_r = right_i[r].first, out_l = fuse(out_r, total_delta);
                    if (!out_left_i.has(out_l)) continue;
                    int32_t r_size = (int32_t)right_i[r].second;
                    size_t o = ret.position(out_l, out_r);
                    for (size_t w = 0; w < W.n_blocks(); ++w) {
                        Charge c1 = W.basis().left_charge(w), c2 = W.basis().right_charge(w);
                        Charge in_r = fuse(out_r, -c1), in_l = fuse(in_r, -T_delta);
                        size_t tb = T.position(in_l, in_r);
                        if (tb == T.size()) continue;
                        int32_t in_off = (int32_t)in_right_pb(c1, out_r), out_off = (int32_t)out_left_pb(c2, in_l);
                        int32_t ldim = (int32_t)T[tb].ls;
                        for (size_t s1 = 0; s1 < W[w].rows; ++s1)
                            for (size_t s2 = 0; s2 < W[w].cols; ++s2) {
                                double alfa = W[w](s1, s2) * term.second;
                                if (alfa == 0.0) continue;   // the reference adds 0*x here
                                if (emit) tasks.push_back(YTask{o, out_off + (int32_t)s2 * ldim, 0, ldim, r_size, b1, tb, 0, in_off + (int32_t)s1 * r_size, alfa});
                                used = true;
                            }
                    }
                }
            }
            if (used) t_rows.push_back(b1);
        }
    }

    // ---- step 2 structure, SU2 lbtm (non-abelian/apply_op.hpp:23-101)
    // Everything the reference's inner loop looks up per (T block, W block) -- the position of the output sector in
    // right_i / out_left_i, the product-basis offsets, the spins -- depends on the T block and ONE physical charge only.
    // It is tabulated once per distinct T block structure (the bonds of one operator type share it) and per operator
    // (positions of its block charges in the physical index), so that the loop over (bond, term, T block, W block) does
    // two table reads instead of six hash / binary searches.
    struct TbPhys { int32_t rb, r_size, in_off, ol_pos, out_off; int16_t jp, ip; };   // rb.., jp: phys as phys_in; ol_pos.., ip: phys as phys_out
    struct TbInfo { int32_t l_size; int16_t i_spin, j_spin; std::vector<TbPhys> ph; };
    void build_lbtm_tables(Layout const& ket_rp, Index const& physical_i, Index const& right_i, Index const& out_left_i,
                           ProductBasis const& in_right_pb, ProductBasis const& out_left_pb)
    {
        const size_t B = t_basis.size(), np = physical_i.size();
        // distinct T block structures
        t_basis_id_.assign(B, 0);
        std::vector<size_t> reps;
        {
            std::unordered_map<uint64_t, std::vector<uint32_t>> by_hash;
            for (size_t b = 0; b < B; ++b) {
                uint64_t h = 1469598103934665603ull;
                for (auto const& q : t_basis[b]) for (uint64_t x : {(uint64_t)(uint32_t)q.lc[0], (uint64_t)(uint32_t)q.lc[1], (uint64_t)(uint32_t)q.lc[2], (uint64_t)(uint32_t)q.rc[0],
                                                                    (uint64_t)(uint32_t)q.rc[1], (uint64_t)(uint32_t)q.rc[2], (uint64_t)q.ls, (uint64_t)q.rs}) { h ^= x; h *= 1099511628211ull; }
                auto& cands = by_hash[h];
                uint32_t id = (uint32_t)-1;
                for (uint32_t c : cands) if (t_basis[reps[c]] == t_basis[b]) { id = c; break; }
                if (id == (uint32_t)-1) { id = (uint32_t)reps.size(); reps.push_back(b); cands.push_back(id); }
                t_basis_id_[b] = id;
            }
        }
        tbinfo_.assign(reps.size(), std::vector<TbInfo>());
#pragma omp parallel for schedule(dynamic, 4)
        for (long q = 0; q < (long)reps.size(); ++q) {
            DualIndex const& T = t_basis[reps[(size_t)q]];
            std::vector<TbInfo>& out = tbinfo_[(size_t)q];
            out.resize(T.size());
            for (size_t tb = 0; tb < T.size(); ++tb) {
                Charge lc = T[tb].lc, rc = T[tb].rc;
                Charge mc = rc;   // ket right-paired blocks are charge-diagonal: mc == rc (site_hamil.hpp:85-89, apply_op.hpp:62-63)
                { auto it = ket_rp.basis.left_lower_bound(rc); if (it != ket_rp.basis.end()) mc = it->lc; }
                TbInfo& ti = out[tb];
                ti.l_size = (int32_t)T[tb].ls; ti.i_spin = (int16_t)spin(lc); ti.j_spin = (int16_t)spin(mc);
                ti.ph.resize(np);
                for (size_t p = 0; p < np; ++p) {
                    Charge ph = physical_i[p].first;
                    TbPhys& x = ti.ph[p];
                    Charge out_r = fuse(rc, ph), out_l = fuse(lc, ph);
                    size_t rb = right_i.position(out_r);
                    x.rb = rb == right_i.size() ? -1 : (int32_t)rb;
                    x.r_size = x.rb < 0 ? 0 : (int32_t)right_i[rb].second;
                    x.in_off = x.rb < 0 ? 0 : (int32_t)in_right_pb(ph, out_r);
                    x.jp = (int16_t)spin(out_r);
                    size_t ol = out_left_i.position(out_l);
                    x.ol_pos = ol == out_left_i.size() ? -1 : (int32_t)ol;
                    x.out_off = out_left_pb.has(ph, lc) ? (int32_t)out_left_pb(ph, lc) : 0;
                    x.ip = (int16_t)spin(out_l);
                }
            }
        }
        // positions of every operator's block charges in the physical index
        op_phys_.clear();
        auto tbl = mpo.get_operator_table();
        op_phys_.resize(tbl ? tbl->size() : 0);
        std::vector<char> seen_tag(op_phys_.size(), 0);
        for (size_t b2 = 0; b2 < mpo.col_dim(); ++b2)
            for (size_t e = mpo.col_begin(b2); e < mpo.col_end(b2); ++e)
                for (auto const& term : mpo.at_entry(e)) {
                    if (seen_tag[term.first]) continue;
                    seen_tag[term.first] = 1;
                    SiteOperator const& W = mpo.op(term.first);
                    auto& v = op_phys_[term.first];
                    v.resize(W.basis().size());
                    for (size_t w = 0; w < W.basis().size(); ++w) {
                        size_t pi = physical_i.position(W.basis().left_charge(w)), po = physical_i.position(W.basis().right_charge(w));
                        if (pi == physical_i.size() || po == physical_i.size()) throw std::runtime_error("plan: operator block outside the physical index of the site");
                        v[w] = std::make_pair((int16_t)pi, (int16_t)po);
                    }
                }
        lbtm_nr_ = right_i.size(); lbtm_nol_ = out_left_i.size();
    }
    void y_struct_su2_lbtm(size_t b2, Layout const& /*ket_rp*/, Index const& right_i, Index const& out_left_i,
                           ProductBasis const& /*in_right_pb*/, ProductBasis const& /*out_left_pb*/,
                           DualIndex& ret, std::vector<YTask>& tasks, std::vector<size_t>& t_rows, YOpts const& opts = YOpts())
    {
        std::vector<std::vector<uint32_t>> touched;           // per created block: bonds that touch it
        // Blocks are created lazily, in loop order, but sorted on insertion: a task first carries the creation number
        // of its block and gets the block's final position once the structure is complete.
        std::vector<int32_t> cid_of(lbtm_nol_ * lbtm_nr_, -1);
        std::vector<std::pair<int32_t, int32_t>> created;      // (position in out_left_i, position in right_i)
        const size_t first_task = tasks.size();
        const int ap = mpo.right_spin(b2).get();
        for (size_t e = mpo.col_begin(b2); e < mpo.col_end(b2); ++e) {
            size_t b1 = mpo.row_of(e);
            std::vector<TbInfo> const& TI = tbinfo_[t_basis_id_[b1]];
            const int a = mpo.left_spin(b1).get();
            bool used = false;
            const bool emit = !structure_only && opts.emit_tasks && (!opts.own || (*opts.own)[b1]);
            const bool probe = !structure_only && !opts.emit_tasks;      // pass A: does this bond contribute at all?
            for (auto const& term : mpo.at_entry(e)) {
                SiteOperator const& W = mpo.op(term.first);
                const int k = W.spin().get();
                std::vector<std::pair<int16_t, int16_t>> const& oph = op_phys_[term.first];
                const size_t nw = oph.size();
                for (size_t tb = 0; tb < TI.size(); ++tb) {
                    TbInfo const& ti = TI[tb];
                    for (size_t w = 0; w < nw; ++w) {
                        TbPhys const& qi = ti.ph[oph[w].first];
                        if (qi.rb < 0) continue;
                        TbPhys const& qo = ti.ph[oph[w].second];
                        if (!su2::triangle(qi.jp, ap, qo.ip)) continue;
                        if (qo.ol_pos < 0) continue;
                        int32_t& cslot = cid_of[(size_t)qo.ol_pos * lbtm_nr_ + qi.rb];
                        if (cslot < 0) {
                            cslot = (int32_t)created.size();
                            created.push_back(std::make_pair(qo.ol_pos, qi.rb));
                            Charge out_l = out_left_i[qo.ol_pos].first, out_r = right_i[qi.rb].first;
                            if (!ret.has(out_l, out_r)) ret.insert(QnBlock(out_l, out_r, out_left_i[qo.ol_pos].second, (size_t)qi.r_size));
                            if (opts.block_bonds) touched.emplace_back();
                        }
                        if (opts.block_bonds) { auto& v = touched[(size_t)cslot]; if (v.empty() || v.back() != (uint32_t)b1) v.push_back((uint32_t)b1); }
                        if (!emit && !(probe && !used)) continue;
                        const size_t cid = (size_t)cslot;
                        const int i = ti.i_spin, ip = qo.ip, j = ti.j_spin, jp = qi.jp;
                        const int two_sp = std::abs(i - ip), two_s = std::abs(j - jp);
                        double couplings[4];
                        if (const double* cv = su2::coupling_table(j, jp, i, ip, a, k, ap)) { for (int q = 0; q < 4; ++q) couplings[q] = cv[q] * term.second; }
                        else su2::set_coupling(j, two_s, jp, a, k, ap, i, two_sp, ip, term.second, couplings);
                        const int32_t in_off = qi.in_off, out_off = qo.out_off, l_size = ti.l_size, r_size = qi.r_size;
                        for (int s = W.sparse_ptr[w]; s < W.sparse_ptr[w + 1]; ++s) {
                            SparseEntry const& en = W.sparse[s];
                            int cn = 0;
                            if (en.row_spin == 2 && en.col_spin == 2) cn = 3; else if (en.row_spin == 2) cn = 1; else if (en.col_spin == 2) cn = 2;
                            double alfa = en.coefficient * couplings[cn];
                            if (alfa == 0.0) continue;
                            used = true;
                            if (!emit) break;
                            tasks.push_back(YTask{cid, out_off + (int32_t)en.col * l_size, 0, l_size, r_size, b1, tb, 0, in_off + (int32_t)en.row * r_size, alfa});
                        }
                    }
                }
            }
            if (used) t_rows.push_back(b1);
        }
        std::vector<size_t> final_pos(created.size());
        for (size_t c = 0; c < created.size(); ++c) final_pos[c] = ret.position(out_left_i[created[c].first].first, right_i[created[c].second].first);
        for (size_t t = first_task; t < tasks.size(); ++t) tasks[t].o = final_pos[tasks[t].o];
        if (opts.block_bonds) {
            opts.block_bonds->assign(ret.size(), std::vector<uint32_t>());
            for (size_t c = 0; c < created.size(); ++c) (*opts.block_bonds)[final_pos[c]] = std::move(touched[c]);
        }
    }

    // ---- step 2 structure, abelian rbtm (abelian/apply_op.hpp:141-250)
    void y_struct_abelian_rbtm(size_t b1, Index const& left_i, Index const& out_right_i, ProductBasis const& in_left_pb, ProductBasis const& out_right_pb,
                               DualIndex& ret, std::vector<YTask>& tasks, std::vector<size_t>& t_cols)
    {
        for (size_t b2 : mpo.row(b1)) {
            DualIndex const& T = t_basis[b2];
            if (T.size() == 0) continue;
            for (auto const& term : mpo.at(b1, b2)) {
                SiteOperator const& W = mpo.op(term.first);
                if (W.n_blocks() == 0) continue;
                Charge total_delta = fuse(delta_of(W.basis()), -delta_of(T));
                for (size_t l = 0; l < left_i.size(); ++l) {
                    Charge out_l = left_i[l].first, out_r = fuse(out_l, -total_delta);
                    if (!out_right_i.has(out_r)) continue;
                    if (!ret.has(out_l, out_r)) ret.insert(QnBlock(out_l, out_r, left_i[l].second, out_right_i.size_of_block(out_r)));
                }
            }
        }
        if (structure_only) return;
        for (size_t b2 : mpo.row(b1)) {
            DualIndex const& T = t_basis[b2];
            if (T.size() == 0) continue;
            bool used = false;
            for (auto const& term : mpo.at(b1, b2)) {
                SiteOperator const& W = mpo.op(term.first);
                if (W.n_blocks() == 0) continue;
                Charge T_delta = delta_of(T);
                Charge total_delta = fuse(delta_of(W.basis()), -T_delta);
                for (size_t l = 0; l < left_i.size(); ++l) {
                    Charge out_l = left_i[l].first, out_r = fuse(out_l, -total_delta);
                    if (!out_right_i.has(out_r)) continue;
                    int32_t l_size = (int32_t)left_i[l].second;
                    size_t o = ret.position(out_l, out_r);
                    for (size_t w = 0; w < W.n_blocks(); ++w) {
                        Charge c1 = W.basis().left_charge(w), c2 = W.basis().right_charge(w);
                        Charge in_l = fuse(out_l, c1), in_r = fuse(in_l, T_delta);
                        size_t tb = T.position(in_l, in_r);
                        if (tb == T.size()) continue;
                        int32_t in_off = (int32_t)in_left_pb(c1, out_l), out_off = (int32_t)out_right_pb(c2, in_r);
                        int32_t rdim = (int32_t)T[tb].rs;
                        for (size_t s1 = 0; s1 < W[w].rows; ++s1)
                            for (size_t s2 = 0; s2 < W[w].cols; ++s2) {
                                double alfa = W[w](s1, s2) * term.second;
                                if (alfa == 0.0) continue;
                                tasks.push_back(YTask{o, 0, out_off + (int32_t)s2 * rdim, l_size, rdim, b2, tb, in_off + (int32_t)s1 * l_size, 0, alfa});
                                used = true;
                            }
                    }
                }
            }
            if (used) t_cols.push_back(b2);
        }
    }

    // ---- step 2 structure, SU2 rbtm (non-abelian/apply_op.hpp:114-232)
    void y_struct_su2_rbtm(size_t b1, DualIndex const& ket_basis, Index const& left_i, Index const& out_right_i,
                           ProductBasis const& in_left_pb, ProductBasis const& out_right_pb,
                           DualIndex& ret, std::vector<YTask>& tasks, std::vector<size_t>& t_cols)
    {
        // task_capsule map order decides block creation; value: (row size of the first task, creation number).  A task
        // first carries the creation number of its block and gets the block's final position at the end.
        std::map<std::pair<Charge, Charge>, std::pair<int32_t, size_t>> first_l_size;
        const size_t first_task = tasks.size();
        for (size_t b2 : mpo.row(b1)) {
            DualIndex const& T = t_basis[b2];
            bool used = false;
            for (auto const& term : mpo.at(b1, b2)) {
                SiteOperator const& W = mpo.op(term.first);
                int a = mpo.left_spin(b1).get(), k = W.spin().get(), ap = mpo.right_spin(b2).get();
                for (size_t tb = 0; tb < T.size(); ++tb) {
                    Charge lc = T[tb].lc, rc = T[tb].rc;
                    Charge mc = lc;
                    { auto it = ket_basis.left_lower_bound(lc); if (it != ket_basis.end()) mc = it->rc; }
                    for (size_t w = 0; w < W.basis().size(); ++w) {
                        Charge phys_in = W.basis().left_charge(w), phys_out = W.basis().right_charge(w);
                        Charge out_l = fuse(lc, -phys_in);
                        size_t lb = left_i.position(out_l);
                        if (lb == left_i.size()) continue;
                        Charge out_r = fuse(rc, -phys_out);
                        if (!su2::triangle(spin(out_l), a, spin(out_r))) continue;
                        if (!out_right_i.has(out_r)) continue;
                        int32_t l_size = (int32_t)left_i[lb].second;
                        if (structure_only) {
                            auto it = first_l_size.emplace(std::make_pair(out_l, out_r), std::make_pair((int32_t)-1, first_l_size.size())).first;
                            if (W.sparse_ptr[w + 1] > W.sparse_ptr[w] && it->second.first < 0) it->second.first = l_size;
                            continue;
                        }
                        int i = spin(out_r), ip = spin(rc), j = spin(out_l), jp = spin(mc);
                        int two_sp = std::abs(i - ip), two_s = std::abs(j - jp);
                        double couplings[4];
                        su2::set_coupling(j, two_s, jp, a, k, ap, i, two_sp, ip, term.second, couplings);
                        int32_t in_off = (int32_t)in_left_pb(phys_in, out_l), out_off = (int32_t)out_right_pb(phys_out, rc);
                        int32_t r_size = (int32_t)T[tb].rs;
                        // the reference creates the map entry before looking at the operator entries (apply_op.hpp:170)
                        auto it = first_l_size.emplace(std::make_pair(out_l, out_r), std::make_pair((int32_t)-1, first_l_size.size())).first;
                        for (int s = W.sparse_ptr[w]; s < W.sparse_ptr[w + 1]; ++s) {
                            SparseEntry const& en = W.sparse[s];
                            int cn = 0;
                            if (en.row_spin == 2 && en.col_spin == 2) cn = 3; else if (en.row_spin == 2) cn = 1; else if (en.col_spin == 2) cn = 2;
                            double alfa = en.coefficient * couplings[cn];
                            if (it->second.first < 0) it->second.first = l_size;
                            if (alfa != 0.0)
                                tasks.push_back(YTask{it->second.second, 0, out_off + (int32_t)en.col * r_size, l_size, r_size, b2, tb, in_off + (int32_t)en.row * l_size, 0, alfa});
                            used = true;
                        }
                    }
                }
            }
            if (used) t_cols.push_back(b2);
        }
        for (auto const& kv : first_l_size) {
            if (kv.second.first < 0) continue;   // "if (otasks.size() == 0) continue"
            ret.insert(QnBlock(kv.first.first, kv.first.second, (size_t)kv.second.first, out_right_i.size_of_block(kv.first.second)));
        }
        std::vector<size_t> final_pos(first_l_size.size(), 0);
        for (auto const& kv : first_l_size) if (kv.second.first >= 0) final_pos[kv.second.second] = ret.position(kv.first.first, kv.first.second);
        for (size_t t = first_task; t < tasks.size(); ++t) tasks[t].o = final_pos[tasks[t].o];
    }

    // W-application contributions grouped by destination panel.  Sources that reach the same panel through several
    // MPO terms are merged (coefficients add up).
    typedef AxpySrc PanelSrc;
    struct Panel { size_t o; int32_t dst_row, dst_col, rows, cols; std::vector<PanelSrc> srcs; };
    struct PanelRef { Ref A; int32_t lda; double alpha; };
    struct PanelCounts { double flops_w = 0; size_t n_axpy = 0; };
    // Tasks are grouped by destination panel.  A panel is identified by (Y block, first row, first column); in both
    // directions one of the two offsets is zero, so (block, row + column) addresses a flat table directly -- no hashing,
    // no sorting of the task list (the lists of the high fan-in outputs hold millions of entries).  Sources are counted
    // first and every panel's list is allocated once.  The panels are then put in (block, column, row) order.
    std::vector<Panel> collect_panels(PanelCounts& cnt, std::vector<YTask> const& tasks, std::map<size_t, Layout> const& tl) const
    {
        std::vector<Panel> panels;
        if (tasks.empty()) return panels;
        size_t n_o = 0;
        for (YTask const& t : tasks) n_o = std::max(n_o, t.o + 1);
        std::vector<int64_t> base(n_o + 1, 0);
        for (YTask const& t : tasks) base[t.o + 1] = std::max(base[t.o + 1], (int64_t)t.dst_row + t.dst_col + 1);
        for (size_t o = 0; o < n_o; ++o) base[o + 1] += base[o];
        std::vector<int32_t> slot_panel((size_t)base[n_o], -1);
        std::vector<uint32_t> pid(tasks.size()), count;
        for (size_t i = 0; i < tasks.size(); ++i) {
            YTask const& t = tasks[i];
            int32_t& sp = slot_panel[(size_t)(base[t.o] + t.dst_row + t.dst_col)];
            if (sp < 0) { sp = (int32_t)panels.size(); panels.push_back(Panel{t.o, t.dst_row, t.dst_col, t.rows, t.cols, {}}); count.push_back(0); }
            else {
                Panel const& pn = panels[(size_t)sp];
                if (pn.dst_row != t.dst_row || pn.dst_col != t.dst_col || pn.rows != t.rows || pn.cols != t.cols)
                    throw std::runtime_error("plan: two destination panels of different shape start at the same element");
            }
            pid[i] = (uint32_t)sp; count[(size_t)sp]++;
        }
        for (size_t p = 0; p < panels.size(); ++p) panels[p].srcs.reserve(count[p]);
        size_t last_bt = (size_t)-1; Layout const* L = nullptr; int srcbuf = 0;
        for (size_t i = 0; i < tasks.size(); ++i) {
            YTask const& t = tasks[i];
            if (t.bt != last_bt) { L = &tl.at(t.bt); srcbuf = t_persistent[t.bt] ? BUF_TP : BUF_T; last_bt = t.bt; }
            int32_t lds = (int32_t)L->basis[t.t_block].ls;
            panels[pid[i]].srcs.push_back(PanelSrc{Ref{srcbuf, L->off[t.t_block] + t.src_row + (int64_t)t.src_col * lds}, lds, t.coef});
            cnt.flops_w += 2.0 * t.rows * t.cols;
        }
        cnt.n_axpy += tasks.size();
        auto src_less = [](PanelSrc const& a, PanelSrc const& b) { return std::tie(a.src.buf, a.src.off) < std::tie(b.src.buf, b.src.off); };
        for (Panel& pn : panels) {
            // the sources of a panel arrive in bond order, i.e. mostly sorted already
            if (!std::is_sorted(pn.srcs.begin(), pn.srcs.end(), src_less)) std::stable_sort(pn.srcs.begin(), pn.srcs.end(), src_less);
            size_t o = 0;
            for (size_t i = 0; i < pn.srcs.size(); ++i) {
                if (o && pn.srcs[o - 1].src.buf == pn.srcs[i].src.buf && pn.srcs[o - 1].src.off == pn.srcs[i].src.off) pn.srcs[o - 1].coef += pn.srcs[i].coef;
                else pn.srcs[o++] = pn.srcs[i];
            }
            pn.srcs.resize(o);
        }
        std::sort(panels.begin(), panels.end(), [](Panel const& x, Panel const& y) {
            return std::tie(x.o, x.dst_col, x.dst_row, x.rows, x.cols) < std::tie(y.o, y.dst_col, y.dst_row, y.rows, y.cols);
        });
        return panels;
    }
    // layouts of the step-1 products one output index needs: multi-use ones where they live in BUF_TP, single-use ones
    // back to back in BUF_T from offset t on; returns the offset after them
    int64_t layouts_for(std::vector<size_t> const& rows, int64_t t, std::map<size_t, Layout>& tl) const
    {
        for (size_t b : rows) {
            if (t_persistent[b]) { tl[b] = tp_layout[b]; continue; }
            Layout L; L.assign(t_basis[b], t);
            t += L.total; tl[b] = L;
        }
        return t;
    }
    // The panels of output i.  They are computed for a batch of outputs at a time in parallel, under the assumption that
    // no wave is flushed inside the batch (the offsets of single-use step-1 products in BUF_T then follow from a prefix
    // sum); when the assumption fails for an output its panels are recomputed and the rest of the batch is discarded.
    struct PanelCache { std::vector<std::vector<Panel>> panels; std::vector<PanelCounts> cnt; std::vector<int64_t> t0; size_t batch = 256; };
    // tasks_of(j, scratch) returns the task list of output j -- either a stored one or one it generates into `scratch`
    static std::vector<YTask>& task_scratch() { static thread_local std::vector<YTask> v; return v; }
    template <class RowsOf, class TasksOf, class Skip>
    std::vector<Panel> cached_panels(PanelCache& pc, size_t n, size_t i, int64_t t_begin, RowsOf rows_of, TasksOf tasks_of, Skip skip,
                                     std::map<size_t, Layout> const& tl, PanelCounts& cnt) const
    {
        if (pc.t0.empty()) { pc.t0.assign(n, -1); pc.panels.resize(n); pc.cnt.resize(n); }
        if (pc.t0[i] < 0) {
            const size_t hi = std::min(n, i + pc.batch);
            int64_t t = t_begin;
            std::vector<size_t> todo;
            for (size_t j = i; j < hi; ++j) {
                pc.t0[j] = -1;
                if (skip(j)) continue;
                pc.t0[j] = t; todo.push_back(j);
                for (size_t b : rows_of(j)) if (!t_persistent[b]) t += t_layout_size(b);
            }
            // the outputs fed by many bonds first: they take a thousand times longer than the rest
            std::stable_sort(todo.begin(), todo.end(), [&](size_t a, size_t b) { return rows_of(a).size() > rows_of(b).size(); });
#pragma omp parallel for schedule(dynamic, 1)
            for (long q = 0; q < (long)todo.size(); ++q) {
                const size_t j = todo[(size_t)q];
                std::map<size_t, Layout> tlj;
                layouts_for(rows_of(j), pc.t0[j], tlj);
                pc.cnt[j] = PanelCounts();
                pc.panels[j] = collect_panels(pc.cnt[j], tasks_of(j, task_scratch()), tlj);
            }
        }
        if (pc.t0[i] == t_begin) {
            cnt = pc.cnt[i];
            std::vector<Panel> r = std::move(pc.panels[i]);
            pc.panels[i] = std::vector<Panel>();
            return r;
        }
        for (size_t j = i + 1; j < std::min(n, i + pc.batch); ++j) { pc.t0[j] = -1; pc.panels[j] = std::vector<Panel>(); }   // stale: offsets moved
        cnt = PanelCounts();
        return collect_panels(cnt, tasks_of(i, task_scratch()), tl);
    }
    // Where the closing product finds a panel: the T panel itself (one source; its coefficient becomes the alpha of
    // the K-segment) or a compact region of BUF_Y filled by the W kernel.  false: the panel is identically zero.
    bool place_panel(Plan& P, AxpyList& al, Panel& pn, int64_t& cur_y, PanelRef& pr)
    {
        if (pn.srcs.empty()) return false;
        int64_t el = (int64_t)pn.rows * pn.cols;
        if (pn.srcs.size() == 1) {
            if (pn.srcs[0].coef == 0.) return false;
            pr = PanelRef{pn.srcs[0].src, pn.srcs[0].lds, pn.srcs[0].coef};
            P.direct_panel_elems += el;
            return true;
        }
        cur_y = (cur_y + 1) & ~(int64_t)1;      // panels start 16-byte aligned
        AxpyDst d; d.dst = Ref{BUF_Y, cur_y}; d.ldd = pn.rows; d.rows = pn.rows; d.cols = pn.cols;
        P.w_panel_elems += el; P.exec_w += 2.0 * el * (double)pn.srcs.size();
        al.dsts.push_back(d);
        al.lists.push_back(std::move(pn.srcs));
        cur_y += el;
        pr = PanelRef{d.dst, pn.rows, 1.};
        return true;
    }

    // Multi-source destination panels, grouped for the device:
    //   cls 1 ("stream"): at most 4 sources; destinations with IDENTICAL source sets are put together (at most 4).
    //                     Executed by a plain FMA streaming kernel -- pure HBM traffic, every source read once.
    //   cls 0 ("gemm"):   more than 4 sources -- the integral-weighted sums over many bond terms.  Destinations that
    //                     share at least half of their sources with the group leader are put together (at most 64);
    //                     the group is evaluated as one dense product  dst[e, d] = sum_u src_u[e] * coef[u][d]  over
    //                     the panel elements e on DMMA tiles (absent pairs get a zero coefficient), so every source
    //                     panel is read once per group instead of once per destination.
    // Candidates are bucketed by two min-hashes of their source sets.
    static void group_axpy(Plan& P, AxpyList& al, WList& wl)
    {
        size_t nd = al.dsts.size();
        struct Key { uint64_t h1, h2; int32_t rows, cols, n; size_t idx; };
        std::vector<Key> keys(nd);
        // packed source reference (buffer, offset): the lists are sorted by it
        auto ref_of = [](AxpySrc const& a) { return ((int64_t)a.src.buf << 56) | a.src.off; };
        typedef std::vector<AxpySrc> SrcVec;
        std::vector<SrcVec>& sorted = al.lists;
        auto mixh = [](uint64_t x, uint64_t seed) { x ^= seed; x *= 0x9E3779B97F4A7C15ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 29; return x; };
#pragma omp parallel for schedule(dynamic, 256)
        for (long il = 0; il < (long)nd; ++il) {
            size_t i = (size_t)il;
            AxpyDst const& d = al.dsts[i];
            SrcVec const& v = sorted[i];
            uint64_t h1 = ~0ull, h2 = ~0ull;
            for (auto const& e : v) { uint64_t r = (uint64_t)ref_of(e); h1 = std::min(h1, mixh(r, 0x1234567ull)); h2 = std::min(h2, mixh(r, 0xABCDEF01ull)); }
            keys[i] = Key{h1, h2, d.rows, d.cols, (int32_t)v.size(), i};
        }
        auto key_less = [](Key const& a, Key const& b) {
            return std::tie(a.rows, a.cols, a.h1, a.h2, a.n, a.idx) < std::tie(b.rows, b.cols, b.h1, b.h2, b.n, b.idx);
        };
        // a total order (idx breaks ties): the parallel sort gives the same sequence as the serial one
#ifdef _OPENMP
        __gnu_parallel::sort(keys.begin(), keys.end(), key_less);
#else
        std::sort(keys.begin(), keys.end(), key_less);
#endif
        auto overlap = [&](SrcVec const& a, SrcVec const& b) {
            size_t i = 0, j = 0, c = 0;
            while (i < a.size() && j < b.size()) { int64_t x = ref_of(a[i]), y = ref_of(b[j]); if (x == y) { ++c; ++i; ++j; } else if (x < y) ++i; else ++j; }
            return c;
        };
        // A group never crosses a change of (rows, cols, h1): the runs of equal (rows, cols, h1) are cut into groups
        // independently (in parallel, greedily from the front of each run, as a serial pass over all keys would), the
        // groups get their places in the source / destination / coefficient arrays by a prefix sum, and are written
        // out in parallel.  The result does not depend on the number of threads.
        typedef std::vector<std::pair<int64_t, int32_t>> UniVec;   // (packed src ref, leading dimension) of a group's sources
        struct GInfo { size_t q; int32_t g; bool stream; UniVec uni; std::vector<uint32_t> extra; bool absorbed; };    // extra: destinations taken over from absorbed groups
        std::vector<size_t> run_begin;
        for (size_t q = 0; q < nd; ++q)
            if (q == 0 || keys[q].rows != keys[q - 1].rows || keys[q].cols != keys[q - 1].cols || keys[q].h1 != keys[q - 1].h1) run_begin.push_back(q);
        run_begin.push_back(nd);
        const long n_runs = (long)run_begin.size() - 1;
        std::vector<std::vector<GInfo>> per_run((size_t)std::max(n_runs, 0L));
#pragma omp parallel for schedule(dynamic, 8)
        for (long r = 0; r < n_runs; ++r) {
            UniVec uni, tmp;
            const size_t q_end = run_begin[(size_t)r + 1];
            for (size_t q = run_begin[(size_t)r]; q < q_end;) {
                size_t lead = keys[q].idx;
                size_t q2 = q + 1;
                bool stream = sorted[lead].size() <= 4;
                size_t cap = stream ? 4 : 64;
                while (q2 < q_end && q2 - q < cap) {
                    SrcVec const& cand = sorted[keys[q2].idx];
                    size_t c = overlap(sorted[lead], cand);
                    if (stream) { if (c != sorted[lead].size() || c != cand.size()) break; }
                    else if (cand.size() <= 4 || 2 * c < std::max(sorted[lead].size(), cand.size())) break;
                    ++q2;
                }
                int32_t g = (int32_t)(q2 - q);
                uni.clear();
                for (auto const& e : sorted[lead]) uni.push_back(std::make_pair(ref_of(e), e.lds));
                for (int32_t d = 1; d < g; ++d) {
                    tmp.clear();
                    SrcVec const& m = sorted[keys[q + d].idx];
                    size_t i = 0, j = 0;
                    while (i < uni.size() || j < m.size()) {
                        if (j == m.size() || (i < uni.size() && uni[i].first < ref_of(m[j]))) tmp.push_back(uni[i++]);
                        else if (i == uni.size() || ref_of(m[j]) < uni[i].first) { tmp.push_back(std::make_pair(ref_of(m[j]), m[j].lds)); ++j; }
                        else { tmp.push_back(uni[i]); ++i; ++j; }
                    }
                    uni.swap(tmp);
                }
                per_run[(size_t)r].push_back(GInfo{q, g, stream, uni, std::vector<uint32_t>(), false});
                q = q2;
            }
        }
        std::vector<GInfo*> flat;
        for (auto& v : per_run) for (auto& gi : v) flat.push_back(&gi);
        // Piggy-back pass.  The DMMA product of a group always computes ng = 8/16/32/64 destination columns; a group of 55
        // destinations leaves 9 columns idle.  Small groups with many sources (a single column of the MPO with its own row
        // set: 1-2 destinations, hundreds of sources -- a pure streaming read) whose sources are (mostly) read by a large
        // group of the same panel shape anyway move into those idle columns: their sources are then read once instead of
        // twice.  At cfg3 this removes a quarter of the W pass's HBM traffic.  Serial, deterministic, cheap (a few thousand
        // candidate pairs).
        if (!getenv("QCM_NO_PIGGYBACK")) {
            auto ng_of = [](int32_t g) { return g <= 8 ? 8 : g <= 16 ? 16 : g <= 32 ? 32 : 64; };
            std::map<std::pair<int32_t, int32_t>, std::vector<size_t>> big;       // panel shape -> groups of the 32 / 64 classes
            for (size_t f = 0; f < flat.size(); ++f) if (!flat[f]->stream && flat[f]->g > 16) big[std::make_pair(keys[flat[f]->q].rows, keys[flat[f]->q].cols)].push_back(f);
            std::vector<size_t> small;
            for (size_t f = 0; f < flat.size(); ++f) if (!flat[f]->stream && flat[f]->g <= 16 && flat[f]->uni.size() >= 32) small.push_back(f);
            std::stable_sort(small.begin(), small.end(), [&](size_t a, size_t b) { return flat[a]->uni.size() > flat[b]->uni.size(); });
            UniVec tmp;
            for (size_t sf : small) {
                GInfo& sm = *flat[sf];
                auto it = big.find(std::make_pair(keys[sm.q].rows, keys[sm.q].cols));
                if (it == big.end()) continue;
                size_t best = (size_t)-1, best_c = 0;
                for (size_t bf : it->second) {
                    GInfo const& bg = *flat[bf];
                    const int32_t have = bg.g + (int32_t)bg.extra.size();
                    if (have + sm.g > ng_of(bg.g)) continue;                  // idle columns only: the product does not grow
                    size_t i = 0, j = 0, c = 0;
                    while (i < sm.uni.size() && j < bg.uni.size()) { if (sm.uni[i].first == bg.uni[j].first) { ++c; ++i; ++j; } else if (sm.uni[i].first < bg.uni[j].first) ++i; else ++j; }
                    if (c > best_c) { best_c = c; best = bf; }
                }
                if (best == (size_t)-1 || 4 * best_c < 3 * sm.uni.size()) continue;     // at least three quarters of the sources are shared
                GInfo& bg = *flat[best];
                for (int32_t d = 0; d < sm.g; ++d) bg.extra.push_back((uint32_t)keys[sm.q + d].idx);
                tmp.clear();
                size_t i = 0, j = 0;
                while (i < bg.uni.size() || j < sm.uni.size()) {
                    if (j == sm.uni.size() || (i < bg.uni.size() && bg.uni[i].first < sm.uni[j].first)) tmp.push_back(bg.uni[i++]);
                    else if (i == bg.uni.size() || sm.uni[j].first < bg.uni[i].first) tmp.push_back(sm.uni[j++]);
                    else { tmp.push_back(bg.uni[i]); ++i; ++j; }
                }
                bg.uni.swap(tmp);
                sm.absorbed = true;
            }
            std::vector<GInfo*> kept;
            for (GInfo* gi : flat) if (!gi->absorbed) kept.push_back(gi);
            flat.swap(kept);
        }
        auto member = [&](GInfo const& gi, int32_t d) -> size_t { return d < gi.g ? keys[gi.q + d].idx : (size_t)gi.extra[(size_t)(d - gi.g)]; };
        const size_t g_base = wl.groups.size();
        size_t n_srcs = wl.srcs.size(), n_dsts = wl.dsts.size(), n_coefs = wl.coefs.size();
        wl.groups.resize(g_base + flat.size());
        for (size_t f = 0; f < flat.size(); ++f) {
            GInfo const& gi = *flat[f];
            int32_t ns = (int32_t)gi.uni.size(), g = gi.g + (int32_t)gi.extra.size();
            WGroup G; G.rows = keys[gi.q].rows; G.cols = keys[gi.q].cols; G.n_src = ns; G.n_dst = g; G.cls = gi.stream ? 1 : 0;
            G.ng = gi.stream ? 4 : (g <= 8 ? 8 : g <= 16 ? 16 : g <= 32 ? 32 : 64);
            G.src_begin = (int32_t)n_srcs; G.dst_begin = (int32_t)n_dsts; G.coef_begin = (int64_t)n_coefs;
            int32_t ns_pad = gi.stream ? ns : (ns + 15) / 16 * 16;   // the DMMA kernel stages sources sixteen at a time
            n_srcs += (size_t)ns; n_dsts += (size_t)g; n_coefs += (size_t)ns_pad * G.ng;
            wl.groups[g_base + f] = G;
            wl.elems_read += (int64_t)ns * G.rows * G.cols;
            wl.elems_written += (int64_t)g * G.rows * G.cols;
        }
        wl.srcs.resize(n_srcs); wl.dsts.resize(n_dsts); wl.coefs.resize(n_coefs);   // coefficients: zeroed per group below
#pragma omp parallel for schedule(dynamic, 64)
        for (long fl = 0; fl < (long)flat.size(); ++fl) {
            GInfo const& gi = *flat[(size_t)fl];
            WGroup const& G = wl.groups[g_base + (size_t)fl];
            const size_t ns_pad = gi.stream ? gi.uni.size() : (gi.uni.size() + 15) / 16 * 16;
            std::fill(wl.coefs.begin() + G.coef_begin, wl.coefs.begin() + G.coef_begin + (int64_t)(ns_pad * (size_t)G.ng), 0.);
            for (size_t u = 0; u < gi.uni.size(); ++u)
                wl.srcs[(size_t)G.src_begin + u] = WSrc{Ref{(int32_t)(gi.uni[u].first >> 56), gi.uni[u].first & (((int64_t)1 << 56) - 1)}, gi.uni[u].second};
            for (int32_t d = 0; d < G.n_dst; ++d) {
                size_t di = member(gi, d);
                wl.dsts[(size_t)G.dst_begin + d] = WDst{al.dsts[di].dst, al.dsts[di].ldd};
                SrcVec const& m = sorted[di];
                size_t u = 0;
                for (auto const& e : m) {
                    while (gi.uni[u].first != ref_of(e)) ++u;
                    wl.coefs[(size_t)G.coef_begin + u * G.ng + d] = e.coef;
                }
            }
        }
        P.w_elems_read += wl.elems_read; P.w_elems_written += wl.elems_written; P.w_groups += (int64_t)wl.groups.size();
        AxpyList().dsts.swap(al.dsts); AxpyList().lists.swap(al.lists);
    }


    // waves, panels and closing tiles of a noise plan.  left: a Y block (lc, rc) is made of row units (panels p, q) and adds
    // p q^T to tile (row of p, row of q) of density-matrix block (lc, lc); right: column units, p^T q into block (rc, rc).
    template <class RowsOf, class TasksOf, class BasisOf, class EmitT>
    void noise_emit(Plan& P, size_t n_out, RowsOf rows_of, TasksOf tasks_of, BasisOf basis_of, EmitT emit_t, bool left_side)
    {
        std::vector<char> own = shard_sources(P, n_out, [&](size_t) { return 1.0; }, rows_of);
        (void)own;
        Layout const& ol = P.out_boundary.b[0];
        PanelCache pcache;
        Wave cur; int64_t cur_y = 0, cur_t = 0;
        auto flush = [&]() {
            if (cur.w_apply.dsts.empty() && cur.close_gemm.outs.empty() && cur.t_gemm.outs.empty()) return;
            cur.y_elems = cur_y; cur.t_elems = cur_t;
            P.y_elems_max = std::max(P.y_elems_max, cur_y); P.t_elems_max = std::max(P.t_elems_max, cur_t);
            merge_outputs(cur.close_gemm); merge_outputs(cur.t_gemm);
            group_axpy(P, cur.w_apply, cur.w_groups);
            P.waves.push_back(std::move(cur));
            cur = Wave(); cur_y = 0; cur_t = 0;
        };
        for (size_t i = 0; i < n_out; ++i) {
            DualIndex const& ybasis = basis_of(i);
            if (ybasis.size() == 0) continue;
            Layout ytmp; ytmp.assign(ybasis);
            int64_t need_t = 0;
            for (size_t b : rows_of(i)) if (!t_persistent[b]) need_t += t_layout_size(b);
            if ((cur_y + cur_t) > 0 && cur_y + cur_t + ytmp.total + need_t > budget) flush();
            const int64_t t_begin = cur_t;
            std::map<size_t, Layout> tl;
            for (size_t b : rows_of(i)) {
                if (t_persistent[b]) { tl[b] = tp_layout[b]; continue; }
                Layout L; L.assign(t_basis[b], cur_t);
                emit_t(P, cur.t_gemm, b, L, BUF_T);
                cur_t += L.total; tl[b] = L;
            }
            for (size_t k = 0; k < ybasis.size(); ++k) {
                QnBlock const& yb = ybasis[k];
                P.flops_close += left_side ? 2.0 * yb.ls * yb.ls * yb.rs : 2.0 * yb.rs * yb.rs * yb.ls; P.n_gemm_tasks++;
            }
            PanelCounts pcnt;
            std::vector<Panel> panels = cached_panels(pcache, n_out, i, t_begin, rows_of,
                [&](size_t j, std::vector<YTask>&) -> std::vector<YTask> const& { return tasks_of(j); }, [&](size_t j) { return basis_of(j).size() == 0; }, tl, pcnt);
            P.flops_w += pcnt.flops_w; P.n_axpy_tasks += pcnt.n_axpy;
            struct Placed { PanelRef pr; Panel const* pn; };
            std::vector<std::vector<Placed>> per_block(ybasis.size());
            for (Panel& pn : panels) {
                PanelRef pr;
                if (!place_panel(P, cur.w_apply, pn, cur_y, pr)) continue;
                per_block[pn.o].push_back(Placed{pr, &pn});
            }
            for (size_t k = 0; k < ybasis.size(); ++k) {
                QnBlock const& yb = ybasis[k];
                Charge const& c = left_side ? yb.lc : yb.rc;
                for (Placed const& p : per_block[k])
                    for (Placed const& q : per_block[k]) {
                        if (left_side) {
                            // tile (p rows, q rows) += (alpha_p P)(rows_p x cols) * (alpha_q Q)^T (cols x rows_q)
                            VBlock bq{c, c, p.pn->cols, q.pn->rows, q.pr.A.off, q.pr.lda, 1, q.pr.alpha};
                            emit_close(P, cur.close_gemm, ol, c, c, p.pr.A, p.pr.lda, 0, p.pn->rows, p.pn->cols, p.pr.alpha, bq, q.pr.A.buf, p.pn->dst_row, 0, q.pn->dst_row);
                        } else {
                            // tile (p cols, q cols) += (alpha_p P)^T (cols_p x rows) * (alpha_q Q)(rows x cols_q)
                            VBlock bq{c, c, p.pn->rows, q.pn->cols, q.pr.A.off, q.pr.lda, 0, q.pr.alpha};
                            emit_close(P, cur.close_gemm, ol, c, c, p.pr.A, p.pr.lda, 1, p.pn->cols, p.pn->rows, p.pr.alpha, bq, q.pr.A.buf, p.pn->dst_col, 0, q.pn->dst_col);
                        }
                    }
            }
        }
        flush();
        merge_outputs(P.persistent_t);
    }

    // step 3: one K-segment alpha * op(A)(m x k) * op(B)(k x n) into rows [c_row, c_row + m) of output block (lc, rc);
    // op(B) starts at row b_k of the stored operand (a panel covers a K sub-range of the reference's product)
    void emit_close(Plan& P, GemmList& gl, Layout const& ol, Charge const& lc, Charge const& rc, Ref A, int32_t lda, int32_t ta, int32_t m, int32_t k,
                    double alpha, VBlock const& b, int bbuf, int32_t c_row, int32_t b_k, int32_t c_col = 0)
    {
        size_t cb = ol.basis.position(lc, rc);
        if (cb == ol.basis.size()) return;
        Out o; o.C = Ref{BUF_OUT, ol.off[cb] + c_row + (int64_t)c_col * (int64_t)ol.basis[cb].ls}; o.ldc = (int32_t)ol.basis[cb].ls; o.m = m; o.n = b.rs;
        o.seg_begin = (int32_t)gl.segs.size();
        int64_t boff = b.off + (b.trans ? (int64_t)b_k * b.ld : (int64_t)b_k);
        gl.segs.push_back(Seg{A, Ref{bbuf, boff}, lda, b.ld, m, o.n, k, ta, b.trans, alpha * b.scale});
        o.seg_end = (int32_t)gl.segs.size();
        gl.outs.push_back(o);
        P.exec_close += 2.0 * m * o.n * k;
    }
    // outputs that target the same block become ONE output with a longer segment list (segments keep their own m, n)
    static void merge_outputs(GemmList& gl)
    {
        std::vector<size_t> order(gl.outs.size());
        std::iota(order.begin(), order.end(), 0);
        auto by_target = [&](size_t a, size_t b) { return gl.outs[a].C.off < gl.outs[b].C.off; };
#ifdef _OPENMP
        __gnu_parallel::stable_sort(order.begin(), order.end(), by_target);
#else
        std::stable_sort(order.begin(), order.end(), by_target);
#endif
        // runs of equal target -> one output each; the segments are copied run by run at prefix-summed offsets (in parallel)
        std::vector<size_t> run_begin;
        for (size_t q = 0; q < order.size(); ++q)
            if (q == 0 || gl.outs[order[q]].C.off != gl.outs[order[q - 1]].C.off) run_begin.push_back(q);
        const size_t n_runs = run_begin.size();
        run_begin.push_back(order.size());
        std::vector<int64_t> seg_off(n_runs + 1, 0);
#pragma omp parallel for schedule(static) if (n_runs > 4096)
        for (long r_ = 0; r_ < (long)n_runs; ++r_) {
            int64_t c = 0;
            for (size_t q = run_begin[(size_t)r_]; q < run_begin[(size_t)r_ + 1]; ++q) c += gl.outs[order[q]].seg_end - gl.outs[order[q]].seg_begin;
            seg_off[(size_t)r_ + 1] = c;
        }
        for (size_t r_ = 0; r_ < n_runs; ++r_) seg_off[r_ + 1] += seg_off[r_];
        GemmList r;
        r.outs.resize(n_runs); r.segs.resize((size_t)seg_off[n_runs]);
#pragma omp parallel for schedule(static) if (n_runs > 4096)
        for (long r_ = 0; r_ < (long)n_runs; ++r_) {
            Out o = gl.outs[order[run_begin[(size_t)r_]]];
            int64_t w = seg_off[(size_t)r_];
            for (size_t q = run_begin[(size_t)r_]; q < run_begin[(size_t)r_ + 1]; ++q) {
                Out const& x = gl.outs[order[q]];
                o.m = std::max(o.m, x.m); o.n = std::max(o.n, x.n);
                for (int32_t s_ = x.seg_begin; s_ < x.seg_end; ++s_) r.segs[(size_t)w++] = gl.segs[(size_t)s_];
            }
            o.seg_begin = (int32_t)seg_off[(size_t)r_]; o.seg_end = (int32_t)seg_off[(size_t)r_ + 1];
            r.outs[(size_t)r_] = o;
        }
        gl = std::move(r);
    }

    // ---- sharding across ranks.  sigma = sum over the EDGES (b1, b2) of the MPO bond graph of W(b1,b2) (*) T[b1] . R[b2] is
    // bilinear, and every result is combined by an allreduce anyway, so the edges can go to any rank.  They are sharded
    // by the step-1 index (b1 for sigma and the left step, b2 for the right step): a rank owns a set of bonds, computes
    // their step-1 products exactly once, and handles every edge that leaves them -- partial W sums for all output
    // indices they feed and the closing products of those partial sums.  In the quantum-chemical MPO most output
    // indices have a single source (their whole cost follows that source), while the few high fan-in ones (the
    // integral-weighted sums over all operator pairs) are summed in slices, one per rank; sharding by the OUTPUT index
    // instead makes every rank recompute nearly all step-1 products, because those few outputs need all
    // of them.  Bonds are taken in order of decreasing cost and go to the least loaded rank; every rank runs the same
    // deterministic assignment.  Returns own[b]; emits the step-1 products of the owned multi-use bonds.
    template <class Cost, class Rows> std::vector<char> shard_sources(Plan& P, size_t n_out, Cost out_cost, Rows rows)
    {
        const size_t B = t_basis.size();
        std::vector<char> own(B, 0);
        owner_.assign(B, -1);
        if (world <= 1) {
            for (size_t i = 0; i < n_out; ++i) for (size_t b : rows(i)) { own[b] = 1; owner_[b] = 0; }
            emit_persistent(P, own);
            return own;
        }
        std::vector<double> c(B, 0.);
        std::vector<char> used(B, 0);
        for (size_t i = 0; i < n_out; ++i) {
            auto const& r = rows(i);
            if (r.empty()) continue;
            const double share = out_cost(i) / (double)r.size();     // closing + W cost of output i, spread over its sources
            for (size_t b : r) { c[b] += share; used[b] = 1; }
        }
        std::vector<std::pair<double, size_t>> order;
        for (size_t b = 0; b < B; ++b) if (used[b]) order.push_back(std::make_pair(c[b] + t_cost(b), b));
        std::stable_sort(order.begin(), order.end(), [](auto const& x, auto const& y) { return x.first > y.first; });
        std::vector<double> load(world, 0.);
        for (auto const& e : order) {
            int r = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            load[r] += e.first;
            own[e.second] = (r == rank);
            owner_[e.second] = r;
        }
        // the step-1 products behind an output whose sources sit on several ranks are read by the exchange wave, which runs
        // before all others: they are kept resident like the multi-use ones
        for (size_t i = 0; i < n_out; ++i) {
            auto const& r = rows(i);
            bool spread = false;
            for (size_t b : r) if (owner_[b] != owner_[r[0]]) { spread = true; break; }
            if (spread) for (size_t b : r) t_persistent[b] = 1;
        }
        emit_persistent(P, own);
        return own;
    }

    // ---- exchange of partial W sums (world > 1).  With the edges of the MPO bond graph sharded by their step-1 index, an
    // output fed by bonds of several ranks has its destination panels summed in slices, one per rank.  Closing every slice
    // on its rank repeats the closing product world times (46 % extra closing FLOPs at cfg3 on 8 ranks).  Instead
    // the slices of such panels are written into one region of BUF_Y that has the same layout on every rank, cut into
    // `world` chunks of equal size; one reduce-scatter hands rank r the complete sums of chunk r, and rank r alone closes
    // them.  (The reference's dead Ambient code did the same for its "exception" columns, detail/ambient.hpp:127-214.)
    // Every rank computes the same assignment from the unsharded task lists.
    struct XPanel { uint32_t out, o; int32_t dst_row, dst_col, rows, cols; int32_t chunk; int64_t off; };
    struct Exchange
    {
        std::vector<XPanel> xp;
        std::vector<std::unordered_map<uint64_t, uint32_t>> index;      // per output: (Y block, first row + first column) -> xp
        std::vector<std::vector<uint32_t>> by_output;
        int64_t chunk = 0;
        bool active() const { return chunk > 0; }
        static uint64_t key(size_t o, int32_t dst_row, int32_t dst_col) { return ((uint64_t)o << 32) | (uint32_t)(dst_row + dst_col); }
        long find(size_t out, size_t o, int32_t dst_row, int32_t dst_col) const
        {
            if (!active() || index[out].empty()) return -1;
            auto it = index[out].find(key(o, dst_row, dst_col));
            return it == index[out].end() ? -1 : (long)it->second;
        }
    };
    // block_bonds(i)[o]: the bonds that touch Y block o of output i; used_blocks(i)[o]: the block takes part in a closing
    // product; units_of(i, o, out): the row (column) units of the block -- the panels the W tasks address.  A block fed by
    // bonds of two or more ranks is exchanged as a whole.
    template <class BlockBonds, class Used, class Units>
    Exchange plan_exchange(size_t n_out, BlockBonds block_bonds, Used used_blocks, Units units_of) const
    {
        Exchange X;
        if (world <= 1 || getenv("QCM_NO_EXCHANGE")) return X;
        X.index.resize(n_out); X.by_output.resize(n_out);
        std::vector<std::vector<XPanel>> per_out(n_out);
#pragma omp parallel for schedule(dynamic, 8)
        for (long il = 0; il < (long)n_out; ++il) {
            const size_t i = (size_t)il;
            auto const& bb = block_bonds(i);
            std::vector<char> used;
            for (size_t o = 0; o < bb.size(); ++o) {
                if (bb[o].empty()) continue;
                bool spread = false;
                for (uint32_t b : bb[o]) if (owner_[b] != owner_[bb[o][0]]) { spread = true; break; }
                if (!spread) continue;
                if (used.empty()) used = used_blocks(i);
                if (used[o]) units_of(i, o, per_out[i]);
            }
            std::sort(per_out[i].begin(), per_out[i].end(), [](XPanel const& x, XPanel const& y) {
                return std::tie(x.o, x.dst_col, x.dst_row) < std::tie(y.o, y.dst_col, y.dst_row); });
        }
        for (size_t i = 0; i < n_out; ++i) for (XPanel const& p : per_out[i]) X.xp.push_back(p);
        if (X.xp.empty()) return X;
        // chunks of (nearly) equal size: largest panels first, each to the least filled chunk; panels start 16-byte aligned
        std::vector<uint32_t> order(X.xp.size());
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return (int64_t)X.xp[a].rows * X.xp[a].cols > (int64_t)X.xp[b].rows * X.xp[b].cols; });
        std::vector<int64_t> fill((size_t)world, 0);
        for (uint32_t q : order) {
            int c = (int)(std::min_element(fill.begin(), fill.end()) - fill.begin());
            X.xp[q].chunk = c; X.xp[q].off = fill[c];
            fill[c] += ((int64_t)X.xp[q].rows * X.xp[q].cols + 1) & ~(int64_t)1;
        }
        X.chunk = *std::max_element(fill.begin(), fill.end());
        for (size_t q = 0; q < X.xp.size(); ++q) {
            XPanel& p = X.xp[q];
            p.off += (int64_t)p.chunk * X.chunk;
            X.index[p.out].emplace(Exchange::key(p.o, p.dst_row, p.dst_col), (uint32_t)q);
            X.by_output[p.out].push_back((uint32_t)q);
        }
        return X;
    }
    // this rank's partial sum of an exchanged panel goes to the panel's slot of the exchange region
    void place_exchanged(Plan& P, Wave& xw, Panel& pn, XPanel const& x)
    {
        if (pn.rows != x.rows || pn.cols != x.cols) throw std::runtime_error("plan: exchanged panel changed shape between ranks");
        AxpyDst d; d.dst = Ref{BUF_Y, x.off}; d.ldd = pn.rows; d.rows = pn.rows; d.cols = pn.cols;
        P.w_panel_elems += (int64_t)pn.rows * pn.cols; P.exec_w += 2.0 * pn.rows * pn.cols * (double)pn.srcs.size();
        xw.w_apply.dsts.push_back(d);
        xw.w_apply.lists.push_back(std::move(pn.srcs));
    }
    // the exchange wave becomes waves[0]
    void finish_exchange(Plan& P, Wave& xw, Exchange const& X)
    {
        if (!X.active()) return;
        xw.x_chunk = X.chunk; xw.y_elems = X.chunk * world; xw.t_elems = 0;
        xw.x_zero = xw.w_apply.dsts.size() < X.xp.size();
        merge_outputs(xw.close_gemm);
        group_axpy(P, xw.w_apply, xw.w_groups);
        P.y_elems_max = std::max(P.y_elems_max, xw.y_elems);
        P.waves.insert(P.waves.begin(), std::move(xw));
    }
    // the tasks / step-1 rows of one output index that belong to this rank
    // returns whether this rank books the reference's closing FLOPs of the output (it owns the first source): the
    // algorithmic FLOP counts of the ranks then add up to exactly the unsharded schedule's
    bool filter_owned(std::vector<char> const& own, std::vector<size_t>& rows) const { std::vector<YTask> none; return filter_owned(own, none, rows); }
    bool filter_owned(std::vector<char> const& own, std::vector<YTask>& tasks, std::vector<size_t>& rows) const
    {
        if (world <= 1) return true;
        const bool books = rows.empty() ? rank == 0 : own[*std::min_element(rows.begin(), rows.end())] != 0;
        size_t o = 0;
        for (size_t i = 0; i < tasks.size(); ++i) if (own[tasks[i].bt]) tasks[o++] = tasks[i];
        tasks.resize(o);
        o = 0;
        for (size_t i = 0; i < rows.size(); ++i) if (own[rows[i]]) rows[o++] = rows[i];
        rows.resize(o);
        return books;
    }
    double estimate_cost(Layout const& y, BoundaryLayout const&, size_t) const
    {
        double c = 0;
        for (size_t k = 0; k < y.basis.size(); ++k) c += (double)y.basis[k].ls * (double)y.basis[k].rs * (double)y.basis[k].rs;
        return c;
    }

    SymmKind symm; bool su2_;
    MPOTensor const& mpo;
    bool isHermitian;
    int rank, world;
    int64_t budget;
    std::vector<DualIndex> t_basis;
    std::vector<VView> t_views;
    std::vector<char> t_persistent;
    std::vector<Layout> tp_layout;
    bool t_is_left = true; Layout t_ket;
    Layout ket_lp_for_right;
    // tables of the SU2 lbtm structure pass (build_lbtm_tables)
    std::vector<int> owner_;            // rank that computes the step-1 product of every bond (shard_sources)
    std::vector<uint32_t> t_basis_id_;
    std::vector<std::vector<TbInfo>> tbinfo_;
    std::vector<std::vector<std::pair<int16_t, int16_t>>> op_phys_;
    size_t lbtm_nr_ = 0, lbtm_nol_ = 0;
};

}} // namespace qcm::plan
