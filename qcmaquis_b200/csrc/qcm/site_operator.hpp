// SiteOperator (+ SpinDescriptor, sparse entry list) and the tagged operator table.
//   SpinDescriptor   dmrg/block_matrix/symmetry/spin_descriptor.h:43-117
//   SiteOperator     dmrg/block_matrix/site_operator.h:25-154, site_operator.hpp:300-341
//   SparseOperator   dmrg/block_matrix/sparse_operator.h:16-168
//   gemm / op_kron   dmrg/block_matrix/site_operator_algorithms.h:24-67,168-321
//   TagHandler/OPTable  dmrg/models/OperatorHandlers/TagHandler.hpp, OpTable.hpp; tag_detail.h:56-134
#pragma once
#include "block_matrix.hpp"
#include "wigner.hpp"
#include <map>
#include <memory>

namespace qcm {

struct SpinDescriptor
{
    int twoS = 0, diff = 0;
    SpinDescriptor() {}
    SpinDescriptor(int twoS_, int in, int out) : twoS(twoS_), diff(out - in) {}
    int get() const { return twoS; }
    int action() const { return diff; }
    bool operator==(SpinDescriptor const& o) const { return twoS == o.twoS && diff == o.diff; }
};
// "Attention: not symmetric" (spin_descriptor.h:86-89): a.twoS += b.action()
inline SpinDescriptor couple(SpinDescriptor a, SpinDescriptor const& b) { a.twoS += b.action(); return a; }

// spin_descriptor.h:107-117 (spin-1/2 sites only)
inline int product_spin(Charge const& a, Charge const& b)
{
    int sa = spin(a), sb = spin(b);
    if (sa == -1 && sb == 1) return 2;
    return std::abs(sa + sb);
}

struct SparseEntry { unsigned block, row, col; int row_spin, col_spin; double coefficient; };

class SiteOperator
{
public:
    typedef std::map<std::pair<Charge, Charge>, std::pair<std::vector<int>, std::vector<int>>> spin_basis_type;

    block_matrix bm;
    SpinDescriptor spin_;          // SU2 groups; stays (0,0) for abelian groups, as the reference's empty tag
    spin_basis_type spin_basis;    // per-block row/column two-site spins J, J' (set by the SU2 op_kron only)
    std::vector<SparseEntry> sparse;
    std::vector<int> sparse_ptr;   // block b owns sparse[sparse_ptr[b] .. sparse_ptr[b+1])

    SpinDescriptor& spin() { return spin_; }
    SpinDescriptor const& spin() const { return spin_; }
    DualIndex const& basis() const { return bm.basis(); }
    size_t n_blocks() const { return bm.n_blocks(); }
    Matrix const& operator[](size_t k) const { return bm[k]; }
    Matrix& operator[](size_t k) { return bm[k]; }
    bool has_block(Charge const& a, Charge const& b) const { return bm.has_block(a, b); }
    void insert_block(double v, Charge const& a, Charge const& b) { bm.insert_block(Matrix(1, 1, v), a, b); }

    // site_operator.hpp:57-101: += merges blocks and extends the spin basis with the non-zero labels of rhs
    SiteOperator& operator+=(SiteOperator const& rhs)
    {
        if (n_blocks() == 0) spin_ = rhs.spin_;
        bm += rhs.bm;
        for (auto const& kv : rhs.spin_basis) {
            auto& sb = spin_basis[kv.first];
            sb.first.resize(std::max(sb.first.size(), kv.second.first.size()));
            sb.second.resize(std::max(sb.second.size(), kv.second.second.size()));
            for (size_t i = 0; i < std::min(sb.first.size(), kv.second.first.size()); ++i)
                if (kv.second.first[i] != 0) sb.first[i] = kv.second.first[i];
            for (size_t i = 0; i < std::min(sb.second.size(), kv.second.second.size()); ++i)
                if (kv.second.second[i] != 0) sb.second[i] = kv.second.second[i];
        }
        return *this;
    }
    SiteOperator& operator*=(double a) { bm *= a; return *this; }

    // site_operator.hpp:300-341 + sparse_operator.h:128-158
    void update_sparse(bool su2)
    {
        if (su2 && spin_basis.empty())
            for (size_t b = 0; b < bm.n_blocks(); ++b) {
                auto key = std::make_pair(bm.basis().left_charge(b), bm.basis().right_charge(b));
                spin_basis[key] = std::make_pair(std::vector<int>(bm[b].rows, std::abs(qcm::spin(key.first))),
                                                 std::vector<int>(bm[b].cols, std::abs(qcm::spin(key.second))));
            }
        sparse.clear();
        sparse_ptr.assign(bm.n_blocks() + 1, 0);
        for (size_t b = 0; b < bm.n_blocks(); ++b) {
            sparse_ptr[b] = (int)sparse.size();
            std::vector<int> const* ls = nullptr; std::vector<int> const* rs = nullptr;
            if (su2) {
                auto const& sb = spin_basis.at(std::make_pair(bm.basis().left_charge(b), bm.basis().right_charge(b)));
                ls = &sb.first; rs = &sb.second;
            }
            for (size_t s1 = 0; s1 < bm[b].rows; ++s1)
                for (size_t s2 = 0; s2 < bm[b].cols; ++s2)
                    if (bm[b](s1, s2) != 0.0)
                        sparse.push_back(SparseEntry{(unsigned)b, (unsigned)s1, (unsigned)s2,
                                                     su2 ? (*ls)[s1] : 0, su2 ? (*rs)[s2] : 0, bm[b](s1, s2)});
        }
        sparse_ptr[bm.n_blocks()] = (int)sparse.size();
    }
};

// site_operator_algorithms.h:24-67 (C = A*B over matching sectors, accumulate)
inline void gemm(SiteOperator const& A, SiteOperator const& B, SiteOperator& C)
{
    C = SiteOperator();
    for (size_t k = 0; k < A.n_blocks(); ++k) {
        Charge ar = A.basis().right_charge(k);
        for (auto it = B.basis().left_lower_bound(ar); it != B.basis().end() && it->lc == ar; ++it) {
            size_t mb = it - B.basis().begin();
            Matrix tmp(A[k].rows, it->rs);
            dgemm(A.bm.block(k), B.bm.block(mb), 1.0, 0.0, tmp.data(), tmp.rows);
            C.bm.match_and_add_block(tmp, A.basis().left_charge(k), it->rc);
        }
    }
}

// tag_detail.h:20-45
inline void remove_empty_blocks(SiteOperator& op)
{
    for (size_t b = 0; b < op.n_blocks(); ++b) {
        bool only_zero = true;
        for (double x : op[b].v) if (x != 0.0) { only_zero = false; break; }
        if (only_zero) { op.bm.remove_block(b); --b; }
    }
}

// tag_detail.h:56-134: equality modulo a scale factor (shape AND spin descriptor must match)
inline std::pair<bool, double> op_equal(SiteOperator const& ref, SiteOperator const& sample)
{
    if (!(ref.basis() == sample.basis() && ref.spin() == sample.spin())) return std::make_pair(false, 0.);
    if (sample.n_blocks() == 0) return std::make_pair(true, 1.0);
    double inv1 = 0, inv2 = 0;
    auto first_nz = [](Matrix const& m, double& inv) {
        for (size_t i = 0; i < m.rows; ++i)
            for (size_t j = 0; j < m.cols; ++j) {
                if (std::abs(m(i, j)) > 1.e-50) { inv = 1. / m(i, j); return true; }
                if (i == m.rows - 1 && j == m.cols - 1) return false;
            }
        return false;
    };
    if (!first_nz(ref[0], inv1)) return std::make_pair(false, 0.);
    if (!first_nz(sample[0], inv2)) return std::make_pair(false, 0.);
    for (size_t b = 0; b < ref.n_blocks(); ++b)
        for (size_t i = 0; i < ref[b].v.size(); ++i)
            if (std::abs(ref[b].v[i] * inv1 - sample[b].v[i] * inv2) > 1e-12) return std::make_pair(false, 0.);
    return std::make_pair(true, inv1 / inv2);
}

typedef unsigned tag_type;

class OPTable : public std::vector<SiteOperator>
{
public:
    tag_type register_op(SiteOperator const& op) { push_back(op); return (tag_type)size() - 1; }
    std::pair<tag_type, double> checked_register(SiteOperator const& sample)
    {
        for (size_t i = 0; i < size(); ++i) {
            auto cmp = op_equal((*this)[i], sample);
            if (cmp.first) return std::make_pair((tag_type)i, cmp.second);
        }
        return std::make_pair(register_op(sample), 1.0);
    }
};

class TagHandler
{
public:
    TagHandler() : table(new OPTable()) {}
    std::shared_ptr<OPTable> table;
    std::vector<char> sign_table;        // 1 = fermionic
    std::vector<tag_type> hermitian;
    std::map<std::pair<tag_type, tag_type>, std::pair<tag_type, double>> product_tags;

    tag_type size() const { return (tag_type)table->size(); }
    bool is_fermionic(tag_type t) const { return sign_table[t]; }
    tag_type herm_conj(tag_type t) const { return hermitian[t]; }
    SiteOperator const& get_op(tag_type t) const { return (*table)[t]; }
    tag_type register_op(SiteOperator const& op, bool fermionic)
    {
        sign_table.push_back(fermionic);
        tag_type r = table->register_op(op);
        hermitian.push_back(r);
        return r;
    }
    std::pair<tag_type, double> checked_register(SiteOperator const& op, bool fermionic)
    {
        auto r = table->checked_register(op);
        if (sign_table.size() < table->size()) { sign_table.push_back(fermionic); hermitian.push_back(r.first); }
        return r;
    }
    void hermitian_pair(tag_type a, tag_type b)
    {
        if (hermitian[a] == b && hermitian[b] == a) return;
        std::swap(hermitian[a], hermitian[b]);
    }
    // TagHandler.hpp:134-171
    std::pair<tag_type, double> get_product_tag(tag_type t1, tag_type t2)
    {
        auto it = product_tags.find(std::make_pair(t1, t2));
        if (it != product_tags.end()) return it->second;
        SiteOperator product;
        gemm((*table)[t1], (*table)[t2], product);
        bool kind = sign_table[t1] != sign_table[t2];
        product.spin() = couple(get_op(t2).spin(), get_op(t1).spin());
        auto r = checked_register(product, kind);
        product_tags[std::make_pair(t1, t2)] = r;
        return r;
    }
    std::pair<std::vector<tag_type>, std::vector<double>> get_product_tags(std::vector<tag_type> const& a, std::vector<tag_type> const& b)
    {
        std::pair<std::vector<tag_type>, std::vector<double>> ret;
        for (size_t s = 0; s < a.size(); ++s) {
            auto p = get_product_tag(a[s], b[s]);
            ret.first.push_back(p.first); ret.second.push_back(p.second);
        }
        return ret;
    }
};

// ts_ops.h:21-34
inline std::vector<int> allowed_spins(int left, int right, int k1, int k2)
{
    std::vector<int> r;
    for (int s = std::abs(k1 - k2); s <= std::abs(k1 + k2); s += 2)
        if (right >= std::abs(s - left) && right <= std::abs(s + left)) r.push_back(s);
    return r;
}

// detail::op_kron (block_matrix/detail/alps_detail.hpp:63-75)
inline void kron_fill(Matrix& out, Matrix const& in, Matrix const& alfa, size_t oy, size_t ox,
                      size_t ldim1, size_t ldim2, size_t rdim1, size_t rdim2)
{
    for (size_t l1 = 0; l1 < ldim1; ++l1)
        for (size_t r1 = 0; r1 < rdim1; ++r1)
            for (size_t l2 = 0; l2 < ldim2; ++l2)
                for (size_t r2 = 0; r2 < rdim2; ++r2)
                    out(oy + l1 * ldim2 + l2, ox + r1 * rdim2 + r2) = in(l2, r2) * alfa(l1, r1);
}

// abelian Kronecker product of two site operators (site_operator_algorithms.h:168-207)
inline void op_kron_abelian(Index const& phys_A, Index const& phys_B, SiteOperator const& A, SiteOperator const& B, SiteOperator& C)
{
    C = SiteOperator();
    ProductBasis pb(phys_A, phys_B);
    for (size_t i = 0; i < A.n_blocks(); ++i)
        for (size_t j = 0; j < B.n_blocks(); ++j) {
            Charge nl = fuse(A.basis().left_charge(i), B.basis().left_charge(j));
            Charge nr = fuse(A.basis().right_charge(i), B.basis().right_charge(j));
            Matrix tmp(pb.size(A.basis().left_charge(i), B.basis().left_charge(j)),
                       pb.size(A.basis().right_charge(i), B.basis().right_charge(j)), 0.);
            kron_fill(tmp, B[j], A[i], pb(A.basis().left_charge(i), B.basis().left_charge(j)),
                      pb(A.basis().right_charge(i), B.basis().right_charge(j)),
                      A.basis().left_size(i), B.basis().left_size(j), A.basis().right_size(i), B.basis().right_size(j));
            C.bm.match_and_add_block(tmp, nl, nr);
        }
}

// SU2 Kronecker product with 9j recoupling (site_operator_algorithms.h:215-321)
inline void op_kron_su2(Index const& phys_A, Index const& phys_B, SiteOperator const& Ao, SiteOperator const& Bo, SiteOperator& C,
                        SpinDescriptor lspin, SpinDescriptor mspin, SpinDescriptor rspin, int target_spin)
{
    ProductBasis pb(phys_A, phys_B);
    SiteOperator A = Ao, B = Bo;
    // expand the small identity to the full one ("Hack", :238-257)
    if (A.spin().get() > 0 && B.spin().get() == 0) {
        Charge cb = phys_B[1].first, cc = phys_B[2].first;
        if (!B.has_block(cb, cc)) { B.insert_block(1., cb, cc); B.insert_block(1., cc, cb); }
    }
    if (A.spin().get() == 0 && B.spin().get() > 0) {
        Charge cb = phys_A[1].first, cc = phys_A[2].first;
        if (!A.has_block(cb, cc)) { A.insert_block(1., cb, cc); A.insert_block(1., cc, cb); }
    }
    int k1 = A.spin().get(), k2 = B.spin().get();
    int j = lspin.get(), jpp = mspin.get(), jp = rspin.get();
    std::vector<int> ps = allowed_spins(j, jp, k1, k2);
    int k = (target_spin > -1) ? target_spin : ps[0];

    SiteOperator::spin_basis_type basis_spins;
    block_matrix blocks;
    for (size_t i = 0; i < A.n_blocks(); ++i)
        for (size_t jb = 0; jb < B.n_blocks(); ++jb) {
            Charge inA = A.basis().left_charge(i), outA = A.basis().right_charge(i);
            Charge inB = B.basis().left_charge(jb), outB = B.basis().right_charge(jb);
            Charge nl = fuse(inA, inB), nr = fuse(outA, outB);
            Matrix tmp(pb.size(inA, inB), pb.size(outA, outB), 0.);
            size_t in_offset = pb(inA, inB), out_offset = pb(outA, outB);
            kron_fill(tmp, B[jb], A[i], in_offset, out_offset,
                      A.basis().left_size(i), B.basis().left_size(jb), A.basis().right_size(i), B.basis().right_size(jb));
            int j1 = std::abs(spin(inA)), j2 = std::abs(spin(inB)), J = product_spin(inA, inB);
            int j1p = std::abs(spin(outA)), j2p = std::abs(spin(outB)), Jp = product_spin(outA, outB);
            tmp *= su2::mod_coupling(j1, j2, J, k1, k2, k, j1p, j2p, Jp);
            blocks.match_and_add_block(tmp, nl, nr);
            auto& bs = basis_spins[std::make_pair(nl, nr)];
            bs.first.resize(tmp.rows); bs.first[in_offset] = J;
            bs.second.resize(tmp.cols); bs.second[out_offset] = Jp;
        }
    double coupling = std::sqrt((jpp + 1.) * (k + 1.)) * su2::wigner6j(j, jp, k, k2, k1, jpp);
    coupling = (((j + jp + k1 + k2) / 2) % 2) ? -coupling : coupling;
    blocks *= coupling;
    C = SiteOperator();
    C.bm = blocks;
    C.spin_basis = basis_spins;
    C.spin() = SpinDescriptor(k, j, jp);
}

} // namespace qcm
