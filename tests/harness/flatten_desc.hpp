// The QCMaquis-side half of the descriptor ABI, written against this repository's mirror classes (same member names as the
// reference's MPOTensor / SiteOperator / MPSTensor / Boundary): flatten the problem into the plain arrays of
// include/qcm_b200.h.  This is the code INTEGRATION.md shows as the binding a maintainer adds to
// contractions/engine.hpp -- ~100 lines, no planner, no C++ type crosses the boundary.  Test infrastructure here.
#pragma once
#include "../../include/qcm_b200.h"
#include "qcm/mpo.hpp"
#include "qcm/mps.hpp"

namespace qcmflat {
using namespace qcm;

inline qcm_charge charge(Charge const& c) { qcm_charge q; q.c[0] = c[0]; q.c[1] = c[1]; q.c[2] = c[2]; return q; }
inline int symm_code(SymmKind s) { return is_su2(s) ? (has_pg(s) ? QCM_SYMM_SU2U1PG : QCM_SYMM_SU2U1) : (has_pg(s) ? QCM_SYMM_2U1PG : QCM_SYMM_2U1); }

struct FlatMPO
{
    std::vector<qcm_site_op_desc> ops;
    std::vector<std::vector<qcm_block>> op_blocks;
    std::vector<std::vector<int32_t>> op_ptr, op_row, op_col, op_rs, op_cs;
    std::vector<std::vector<double>> op_coef;
    std::vector<int64_t> col_ptr, row_idx, term_ptr, left_herm, right_herm;
    std::vector<int32_t> term_op, left_spin, right_spin, left_phase, right_phase;
    std::vector<double> term_scale;
    qcm_mpo_desc d;

    FlatMPO(SymmKind symm, MPOTensor const& mpo)
    {
        OPTable const& tbl = *mpo.get_operator_table();
        const size_t nt = tbl.size();
        ops.resize(nt); op_blocks.resize(nt); op_ptr.resize(nt); op_row.resize(nt); op_col.resize(nt); op_rs.resize(nt); op_cs.resize(nt); op_coef.resize(nt);
        for (size_t t = 0; t < nt; ++t) {
            SiteOperator const& op = tbl[t];
            for (size_t b = 0; b < op.n_blocks(); ++b) {
                QnBlock const& q = op.basis()[b];
                op_blocks[t].push_back(qcm_block{charge(q.lc), charge(q.rc), (int64_t)q.ls, (int64_t)q.rs});
                op_ptr[t].push_back((int32_t)op_row[t].size());
                for (int s = op.sparse_ptr[b]; s < op.sparse_ptr[b + 1]; ++s) {      // op.get_sparse() in the reference
                    SparseEntry const& e = op.sparse[s];
                    op_row[t].push_back((int32_t)e.row); op_col[t].push_back((int32_t)e.col);
                    op_rs[t].push_back(e.row_spin); op_cs[t].push_back(e.col_spin); op_coef[t].push_back(e.coefficient);
                }
            }
            op_ptr[t].push_back((int32_t)op_row[t].size());
            ops[t] = qcm_site_op_desc{op.spin().get(), 0, op.spin().action(), (int32_t)op.n_blocks(), op_blocks[t].data(), op_ptr[t].data(), op_row[t].data(), op_col[t].data(),
                                      op_rs[t].data(), op_cs[t].data(), op_coef[t].data()};
        }
        col_ptr.push_back(0); term_ptr.push_back(0);
        for (size_t b2 = 0; b2 < mpo.col_dim(); ++b2) {
            for (size_t e = mpo.col_begin(b2); e < mpo.col_end(b2); ++e) {
                row_idx.push_back((int64_t)mpo.row_of(e));
                for (auto const& term : mpo.at_entry(e)) { term_op.push_back((int32_t)term.first); term_scale.push_back(term.second); }
                term_ptr.push_back((int64_t)term_op.size());
            }
            col_ptr.push_back((int64_t)row_idx.size());
        }
        for (size_t b = 0; b < mpo.row_dim(); ++b) { left_spin.push_back(mpo.left_spin(b).get()); left_herm.push_back((int64_t)mpo.herm_info.left_conj(b)); left_phase.push_back(mpo.herm_info.left_phase(b)); }
        for (size_t b = 0; b < mpo.col_dim(); ++b) { right_spin.push_back(mpo.right_spin(b).get()); right_herm.push_back((int64_t)mpo.herm_info.right_conj(b)); right_phase.push_back(mpo.herm_info.right_phase(b)); }
        d = qcm_mpo_desc{symm_code(symm), (int32_t)nt, ops.data(), (int64_t)mpo.row_dim(), (int64_t)mpo.col_dim(), (int64_t)row_idx.size(), col_ptr.data(), row_idx.data(),
                         term_ptr.data(), term_op.data(), term_scale.data(), left_spin.data(), right_spin.data(), left_herm.data(), right_herm.data(), left_phase.data(), right_phase.data()};
    }
};

struct FlatTensor
{
    std::vector<qcm_sector> phys, left, right; std::vector<qcm_block> blocks; std::vector<double> data;
    qcm_tensor_desc d;
    explicit FlatTensor(MPSTensor const& t)
    {
        t.make_left_paired();
        for (auto const& e : t.site_dim()) phys.push_back(qcm_sector{charge(e.first), (int64_t)e.second});
        for (auto const& e : t.row_dim()) left.push_back(qcm_sector{charge(e.first), (int64_t)e.second});
        for (auto const& e : t.col_dim()) right.push_back(qcm_sector{charge(e.first), (int64_t)e.second});
        for (size_t k = 0; k < t.data().n_blocks(); ++k) {
            QnBlock const& q = t.data().basis()[k];
            blocks.push_back(qcm_block{charge(q.lc), charge(q.rc), (int64_t)q.ls, (int64_t)q.rs});
            data.insert(data.end(), t.data()[k].v.begin(), t.data()[k].v.end());
        }
        d = qcm_tensor_desc{(int32_t)phys.size(), (int32_t)left.size(), (int32_t)right.size(), (int32_t)blocks.size(), phys.data(), left.data(), right.data(), blocks.data()};
    }
};

struct FlatBoundary
{
    std::vector<int64_t> ptr; std::vector<qcm_block> blocks; std::vector<double> data;
    qcm_boundary_desc d;
    explicit FlatBoundary(Boundary const& b)
    {
        ptr.push_back(0);
        for (size_t k = 0; k < b.aux_dim(); ++k) {
            for (size_t j = 0; j < b[k].n_blocks(); ++j) {
                QnBlock const& q = b[k].basis()[j];
                blocks.push_back(qcm_block{charge(q.lc), charge(q.rc), (int64_t)q.ls, (int64_t)q.rs});
                data.insert(data.end(), b[k][j].v.begin(), b[k][j].v.end());
            }
            ptr.push_back((int64_t)blocks.size());
        }
        d = qcm_boundary_desc{(int64_t)b.aux_dim(), ptr.data(), blocks.data()};
    }
};

// result blocks of a plan made through the descriptor entry points -> block_matrix (sigma) / Boundary
inline block_matrix unflatten(std::vector<qcm_block> const& blocks, std::vector<int64_t> const& off, int64_t b0, int64_t b1, const double* data)
{
    block_matrix r;
    for (int64_t k = b0; k < b1; ++k) {
        Matrix m((size_t)blocks[k].ls, (size_t)blocks[k].rs);
        std::copy(data + off[k], data + off[k] + m.v.size(), m.v.begin());
        Charge lc, rc; for (int q = 0; q < 3; ++q) { lc[q] = blocks[k].lc.c[q]; rc[q] = blocks[k].rc.c[q]; }
        r.insert_block(std::move(m), lc, rc);
    }
    return r;
}

} // namespace qcmflat
