// plan_capi.cpp -- the descriptor entry points of include/qcm_b200.h: qcm_mpo_upload, qcm_plan_sigma / left_step / right_step,
// qcm_plan_out_*.  They rebuild the planner's view of the problem (MPOTensor, TensorDesc, BoundaryLayout) from the plain
// arrays a QCMaquis-side binding passes, run plan::Planner and hand the task arrays to qcm_plan_create -- the caller never
// sees a C++ type.  Host code (g++, OpenMP), linked into libqcm_b200.so next to the kernels.
#include "qcm/plan_to_desc.hpp"
#include <mutex>
#include <string>
#include <unordered_map>

using namespace qcm;

extern "C" int qcm_internal_fail(const char* msg);     // sets qcm_last_error (qcm_b200.cu); returns 1

struct qcm_mpo_s { SymmKind symm; MPOTensor mpo; };

namespace {
struct OutStructure { int kind; plan::Layout tensor; plan::BoundaryLayout boundary; };
std::mutex g_mutex;
std::unordered_map<qcm_plan_t, OutStructure> g_out;

int fail(std::string const& s) { return qcm_internal_fail(s.c_str()); }
Charge charge_of(qcm_charge const& q) { Charge c; c[0] = q.c[0]; c[1] = q.c[1]; c[2] = q.c[2]; return c; }
qcm_charge charge_to(Charge const& c) { qcm_charge q; q.c[0] = c[0]; q.c[1] = c[1]; q.c[2] = c[2]; return q; }
SymmKind symm_of(int s)
{
    switch (s) {
        case QCM_SYMM_2U1: return symm_from_string("2u1");
        case QCM_SYMM_2U1PG: return symm_from_string("2u1pg");
        case QCM_SYMM_SU2U1: return symm_from_string("su2u1");
        case QCM_SYMM_SU2U1PG: return symm_from_string("su2u1pg");
    }
    throw std::runtime_error("unknown symmetry code");
}
Index index_of(const qcm_sector* s, int32_t n)
{
    Index ix;
    for (int32_t i = 0; i < n; ++i) ix.insert(std::make_pair(charge_of(s[i].q), (size_t)s[i].size));
    return ix;
}
DualIndex dual_of(const qcm_block* b, int64_t n)
{
    DualIndex d;
    for (int64_t i = 0; i < n; ++i) {
        QnBlock q(charge_of(b[i].lc), charge_of(b[i].rc), (size_t)b[i].ls, (size_t)b[i].rs);
        size_t pos = d.insert(q);
        if (pos != (size_t)i) throw std::runtime_error("block list is not in DualIndex (descending charge) order");
    }
    return d;
}
plan::TensorDesc tensor_of(const qcm_tensor_desc* t)
{
    if (!t) throw std::runtime_error("null tensor descriptor");
    return plan::TensorDesc{index_of(t->phys, t->n_phys), index_of(t->left, t->n_left), index_of(t->right, t->n_right), dual_of(t->blocks, t->n_blocks)};
}
plan::BoundaryLayout boundary_of(const qcm_boundary_desc* b)
{
    if (!b) throw std::runtime_error("null boundary descriptor");
    std::vector<DualIndex> bases((size_t)b->aux_dim);
    for (int64_t k = 0; k < b->aux_dim; ++k) bases[(size_t)k] = dual_of(b->blocks + b->block_ptr[k], b->block_ptr[k + 1] - b->block_ptr[k]);
    plan::BoundaryLayout L; L.assign(bases);
    return L;
}
int finish(plan::Plan const& P, int64_t left_elems, int64_t right_elems, qcm_plan_t* out)
{
    PlanDescHolder H;
    H.fill(P, left_elems, right_elems);
    if (qcm_plan_create(&H.d, out)) return 1;
    std::lock_guard<std::mutex> lock(g_mutex);
    g_out[*out] = OutStructure{P.kind, P.out_tensor, P.out_boundary};
    return 0;
}
}

extern "C" void qcm_internal_forget_plan(qcm_plan_t p)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    g_out.erase(p);
}

extern "C" int qcm_mpo_upload(const qcm_mpo_desc* d, qcm_mpo_t* out)
{
    try {
        if (!d || !out) return fail("qcm_mpo_upload: null argument");
        const SymmKind symm = symm_of(d->symm);
        const bool su2 = is_su2(symm);
        std::shared_ptr<OPTable> table(new OPTable());
        for (int32_t t = 0; t < d->n_ops; ++t) {
            qcm_site_op_desc const& od = d->ops[t];
            SiteOperator op;
            op.spin_ = SpinDescriptor(od.spin_twoS, od.spin_in, od.spin_out);
            for (int32_t b = 0; b < od.n_blocks; ++b) {
                qcm_block const& q = od.blocks[b];
                Matrix m((size_t)q.ls, (size_t)q.rs, 0.);
                Charge lc = charge_of(q.lc), rc = charge_of(q.rc);
                std::vector<int> ls((size_t)q.ls, std::abs(spin(lc))), rs((size_t)q.rs, std::abs(spin(rc)));
                for (int32_t e = od.entry_ptr[b]; e < od.entry_ptr[b + 1]; ++e) {
                    if (od.row[e] < 0 || od.row[e] >= q.ls || od.col[e] < 0 || od.col[e] >= q.rs) return fail("qcm_mpo_upload: operator entry outside its block");
                    m((size_t)od.row[e], (size_t)od.col[e]) = od.coef[e];
                    if (su2 && od.row_spin) ls[(size_t)od.row[e]] = od.row_spin[e];
                    if (su2 && od.col_spin) rs[(size_t)od.col[e]] = od.col_spin[e];
                }
                size_t pos = op.bm.insert_block(m, lc, rc);
                if (pos != (size_t)b) return fail("qcm_mpo_upload: operator blocks are not in DualIndex order");
                if (su2) op.spin_basis[std::make_pair(lc, rc)] = std::make_pair(ls, rs);
            }
            table->register_op(op);
        }
        // the constructor prepends the terms of an entry one by one (mpotensor.hpp:24-34): feed them last to first
        std::vector<PreTerm> tags;
        for (int64_t b2 = 0; b2 < d->col_dim; ++b2)
            for (int64_t e = d->col_ptr[b2]; e < d->col_ptr[b2 + 1]; ++e) {
                if (d->row_idx[e] < 0 || d->row_idx[e] >= d->row_dim) return fail("qcm_mpo_upload: row index outside the MPO tensor");
                for (int64_t t = d->term_ptr[e + 1] - 1; t >= d->term_ptr[e]; --t) {
                    if (d->term_op[t] < 0 || d->term_op[t] >= d->n_ops) return fail("qcm_mpo_upload: term refers to an operator outside the table");
                    tags.push_back(PreTerm{(size_t)d->row_idx[e], (size_t)b2, (tag_type)d->term_op[t], d->term_scale[t]});
                }
            }
        std::vector<SpinDescriptor> ls((size_t)d->row_dim), rs((size_t)d->col_dim);
        if (su2) {
            if (!d->left_spin || !d->right_spin) return fail("qcm_mpo_upload: SU2 groups need the bond spins");
            for (int64_t b = 0; b < d->row_dim; ++b) ls[(size_t)b] = SpinDescriptor(d->left_spin[b], 0, 0);
            for (int64_t b = 0; b < d->col_dim; ++b) rs[(size_t)b] = SpinDescriptor(d->right_spin[b], 0, 0);
        }
        Hermitian h((size_t)d->row_dim, (size_t)d->col_dim);
        if (d->left_herm && d->right_herm) {
            std::vector<size_t> lh((size_t)d->row_dim), rh((size_t)d->col_dim);
            std::vector<int> lp((size_t)d->row_dim, 1), rp((size_t)d->col_dim, 1);
            for (int64_t b = 0; b < d->row_dim; ++b) { lh[(size_t)b] = (size_t)d->left_herm[b]; if (d->left_phase) lp[(size_t)b] = d->left_phase[b]; }
            for (int64_t b = 0; b < d->col_dim; ++b) { rh[(size_t)b] = (size_t)d->right_herm[b]; if (d->right_phase) rp[(size_t)b] = d->right_phase[b]; }
            h = Hermitian(lh, rh, lp, rp);
        }
        qcm_mpo_s* M = new qcm_mpo_s{symm, MPOTensor((size_t)d->row_dim, (size_t)d->col_dim, tags, table, h, ls, rs, su2)};
        *out = M;
        return 0;
    } catch (std::exception const& e) { return fail(std::string("qcm_mpo_upload: ") + e.what()); }
}
extern "C" int qcm_mpo_free(qcm_mpo_t m) { delete m; return 0; }

extern "C" int qcm_plan_sigma(qcm_mpo_t m, const qcm_tensor_desc* ket, const qcm_boundary_desc* left, const qcm_boundary_desc* right,
                              int rank, int world, int64_t budget, qcm_plan_t* out)
{
    try {
        if (!m || !out) return fail("qcm_plan_sigma: null argument");
        plan::BoundaryLayout ll = boundary_of(left), rl = boundary_of(right);
        plan::Planner pl(m->symm, m->mpo, true, world > 1 ? rank : 0, world > 1 ? world : 1, budget > 0 ? budget : (int64_t)1 << 32);
        plan::Plan P = pl.plan_sigma(tensor_of(ket), ll, rl);
        return finish(P, ll.total, rl.total, out);
    } catch (std::exception const& e) { return fail(std::string("qcm_plan_sigma: ") + e.what()); }
}
extern "C" int qcm_plan_left_step(qcm_mpo_t m, const qcm_tensor_desc* bra, const qcm_tensor_desc* ket, const qcm_boundary_desc* left,
                                  int rank, int world, int64_t budget, qcm_plan_t* out)
{
    try {
        if (!m || !out) return fail("qcm_plan_left_step: null argument");
        plan::BoundaryLayout ll = boundary_of(left);
        plan::Planner pl(m->symm, m->mpo, true, world > 1 ? rank : 0, world > 1 ? world : 1, budget > 0 ? budget : (int64_t)1 << 32);
        plan::Plan P = pl.plan_left_step(tensor_of(bra), tensor_of(ket), ll);
        return finish(P, ll.total, 0, out);
    } catch (std::exception const& e) { return fail(std::string("qcm_plan_left_step: ") + e.what()); }
}
extern "C" int qcm_plan_right_step(qcm_mpo_t m, const qcm_tensor_desc* bra, const qcm_tensor_desc* ket, const qcm_boundary_desc* right,
                                   int rank, int world, int64_t budget, qcm_plan_t* out)
{
    try {
        if (!m || !out) return fail("qcm_plan_right_step: null argument");
        plan::BoundaryLayout rl = boundary_of(right);
        plan::Planner pl(m->symm, m->mpo, true, world > 1 ? rank : 0, world > 1 ? world : 1, budget > 0 ? budget : (int64_t)1 << 32);
        plan::Plan P = pl.plan_right_step(tensor_of(bra), tensor_of(ket), rl);
        return finish(P, 0, rl.total, out);
    } catch (std::exception const& e) { return fail(std::string("qcm_plan_right_step: ") + e.what()); }
}

// noise term of the perturbed density matrix (C/common/move_boundary.hpp:68-126 + prediction.hpp:34-47,101-114): the plan runs
// through qcm_boundary_step(plan, boundary, ket, ket, out); out = the density-matrix blocks (one bond entry).  Never sharded.
extern "C" int qcm_plan_noise_left(qcm_mpo_t m, const qcm_tensor_desc* ket, const qcm_boundary_desc* left, int64_t budget, qcm_plan_t* out)
{
    try {
        if (!m || !out) return fail("qcm_plan_noise_left: null argument");
        plan::BoundaryLayout ll = boundary_of(left);
        plan::Planner pl(m->symm, m->mpo, true, 0, 1, budget > 0 ? budget : (int64_t)1 << 32);
        plan::Plan P = pl.plan_noise_left(tensor_of(ket), ll);
        return finish(P, ll.total, 0, out);
    } catch (std::exception const& e) { return fail(std::string("qcm_plan_noise_left: ") + e.what()); }
}
extern "C" int qcm_plan_noise_right(qcm_mpo_t m, const qcm_tensor_desc* ket, const qcm_boundary_desc* right, int64_t budget, qcm_plan_t* out)
{
    try {
        if (!m || !out) return fail("qcm_plan_noise_right: null argument");
        plan::BoundaryLayout rl = boundary_of(right);
        plan::Planner pl(m->symm, m->mpo, true, 0, 1, budget > 0 ? budget : (int64_t)1 << 32);
        plan::Plan P = pl.plan_noise_right(tensor_of(ket), rl);
        return finish(P, 0, rl.total, out);
    } catch (std::exception const& e) { return fail(std::string("qcm_plan_noise_right: ") + e.what()); }
}

extern "C" int qcm_plan_out_size(qcm_plan_t p, int64_t* aux_dim, int64_t* n_blocks, int64_t* n_elems)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_out.find(p);
    if (it == g_out.end()) return fail("qcm_plan_out_size: the plan was not made by qcm_plan_sigma / qcm_plan_left_step / qcm_plan_right_step");
    OutStructure const& O = it->second;
    if (O.kind == 0 || O.kind == 3) { if (aux_dim) *aux_dim = 1; if (n_blocks) *n_blocks = (int64_t)O.tensor.basis.size(); if (n_elems) *n_elems = O.tensor.total; }
    else {
        int64_t nb = 0; for (auto const& l : O.boundary.b) nb += (int64_t)l.basis.size();
        if (aux_dim) *aux_dim = (int64_t)O.boundary.b.size(); if (n_blocks) *n_blocks = nb; if (n_elems) *n_elems = O.boundary.total;
    }
    return 0;
}
extern "C" int qcm_plan_out_blocks(qcm_plan_t p, int64_t* block_ptr, qcm_block* blocks, int64_t* elem_off)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_out.find(p);
    if (it == g_out.end()) return fail("qcm_plan_out_blocks: the plan was not made by qcm_plan_sigma / qcm_plan_left_step / qcm_plan_right_step");
    OutStructure const& O = it->second;
    auto put = [&](plan::Layout const& L, int64_t& n) {
        for (size_t k = 0; k < L.basis.size(); ++k, ++n) {
            if (blocks) blocks[n] = qcm_block{charge_to(L.basis[k].lc), charge_to(L.basis[k].rc), (int64_t)L.basis[k].ls, (int64_t)L.basis[k].rs};
            if (elem_off) elem_off[n] = L.off[k];
        }
    };
    int64_t n = 0;
    if (O.kind == 0 || O.kind == 3) { if (block_ptr) block_ptr[0] = 0; put(O.tensor, n); if (block_ptr) block_ptr[1] = n; }
    else for (size_t b = 0; b < O.boundary.b.size(); ++b) { if (block_ptr) block_ptr[b] = n; put(O.boundary.b[b], n); if (block_ptr) block_ptr[b + 1] = n; }
    return 0;
}
