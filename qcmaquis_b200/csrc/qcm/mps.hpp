// MPSTensor (paired storage + reshapes), Boundary, MPS sector bookkeeping and initial states.
//   MPSTensor            dmrg/mp_tensors/mpstensor.h:29-142, mpstensor.hpp:24-183,346-395
//   reshapes             dmrg/mp_tensors/reshapes.h:177-223,289-335; block_matrix/detail/alps_detail.hpp:128-152
//   Boundary             dmrg/mp_tensors/boundary.h:20-145
//   allowed_sectors      dmrg/mp_tensors/mps_sectors.h:26-111, charge_detail.h
//   default/const init   dmrg/mp_tensors/mps_initializers.h:35-95
//   left/right_boundary  dmrg/mp_tensors/mps.hpp:244-268
#pragma once
#include "block_matrix.hpp"
#include <memory>
#include <random>

namespace qcm {

enum MPSStorageLayout { LeftPaired, RightPaired };

// [(phys, left), right] --> [left, (-phys, right)]   (reshapes.h:177-223)
inline void reshape_left_to_right_new(Index const& physical_i, Index const& left_i, Index const& right_i,
                                      block_matrix const& m1, block_matrix& m2)
{
    m2 = block_matrix();
    ProductBasis in_left(physical_i, left_i);
    ProductBasis out_right(physical_i, right_i, true);
    for (size_t block = 0; block < m1.n_blocks(); ++block) {
        size_t r = right_i.position(m1.basis().right_charge(block));
        if (r == right_i.size()) continue;
        Charge in_r_charge = right_i[r].first;
        for (size_t s = 0; s < physical_i.size(); ++s) {
            size_t l = left_i.position(fuse(m1.basis().left_charge(block), -physical_i[s].first));
            if (l == left_i.size()) continue;
            Charge out_l_charge = left_i[l].first;
            Charge out_r_charge = fuse(-physical_i[s].first, in_r_charge);
            if (!m2.has_block(out_l_charge, out_r_charge))
                m2.insert_block(Matrix(left_i[l].second, out_right.size(out_r_charge), 0.), out_l_charge, out_r_charge);
            size_t in_left_offset = in_left(physical_i[s].first, left_i[l].first);
            size_t out_right_offset = out_right(physical_i[s].first, in_r_charge);
            Matrix const& in_block = m1[block];
            Matrix& out_block = m2(out_l_charge, out_r_charge);
            size_t sdim = physical_i[s].second, ldim = left_i[l].second, rdim = right_i[r].second;
            for (size_t ss = 0; ss < sdim; ++ss)
                for (size_t rr = 0; rr < rdim; ++rr)
                    std::memcpy(&out_block(0, out_right_offset + ss * rdim + rr), &in_block(in_left_offset + ss * ldim, rr), sizeof(double) * ldim);
        }
    }
}

// left-paired m1 over (in_left_i, in_right_i) copied into the (larger) left-paired structure of m2 over (out_left_i, out_right_i),
// everything else of m2 zeroed   (reshapes.h:228-281 reshape_and_pad_left; used by site_ortho_boundaries)
inline void reshape_and_pad_left(Index const& physical_i, Index const& in_left_i, Index const& in_right_i, Index const& out_left_i,
                                 Index const& /*out_right_i*/, block_matrix const& m1, block_matrix& m2)
{
    m2 *= 0.;
    ProductBasis in_left(physical_i, in_left_i);
    ProductBasis out_left(physical_i, out_left_i);
    for (size_t block = 0; block < m1.n_blocks(); ++block)
        for (size_t s = 0; s < physical_i.size(); ++s) {
            size_t r = in_right_i.position(m1.basis().right_charge(block));
            if (r == in_right_i.size()) continue;
            size_t l = in_left_i.position(fuse(m1.basis().left_charge(block), -physical_i[s].first));
            if (l == in_left_i.size()) continue;
            Charge l_charge = fuse(physical_i[s].first, in_left_i[l].first), r_charge = in_right_i[r].first;
            if (!out_left_i.has(in_left_i[l].first)) continue;
            if (!m1.has_block(l_charge, r_charge) || !m2.has_block(l_charge, r_charge)) continue;
            size_t in_left_offset = in_left(physical_i[s].first, in_left_i[l].first);
            size_t out_left_offset = out_left(physical_i[s].first, in_left_i[l].first);
            Matrix const& in_block = m1(l_charge, r_charge);
            Matrix& out_block = m2(l_charge, r_charge);
            const size_t ldim = in_left_i[l].second;
            for (size_t ss = 0; ss < physical_i[s].second; ++ss)
                for (size_t rr = 0; rr < in_right_i[r].second; ++rr)
                    for (size_t ll = 0; ll < ldim; ++ll) out_block(out_left_offset + ss * ldim + ll, rr) = in_block(in_left_offset + ss * ldim + ll, rr);
        }
}

// [left, (-phys, right)] --> [(phys, left), right]   (reshapes.h:289-335)
inline void reshape_right_to_left_new(Index const& physical_i, Index const& left_i, Index const& right_i,
                                      block_matrix const& m1, block_matrix& m2)
{
    m2 = block_matrix();
    ProductBasis in_right(physical_i, right_i, true);
    ProductBasis out_left(physical_i, left_i);
    for (size_t block = 0; block < m1.n_blocks(); ++block) {
        size_t l = left_i.position(m1.basis().left_charge(block));
        if (l == left_i.size()) continue;
        Charge in_l_charge = left_i[l].first;
        for (size_t s = 0; s < physical_i.size(); ++s) {
            size_t r = right_i.position(fuse(m1.basis().right_charge(block), physical_i[s].first));
            if (r == right_i.size()) continue;
            Charge out_l_charge = fuse(physical_i[s].first, in_l_charge);
            Charge out_r_charge = right_i[r].first;
            if (!m2.has_block(out_l_charge, out_r_charge))
                m2.insert_block(Matrix(out_left.size(physical_i[s].first, in_l_charge), right_i[r].second, 0.), out_l_charge, out_r_charge);
            size_t in_right_offset = in_right(physical_i[s].first, out_r_charge);
            size_t out_left_offset = out_left(physical_i[s].first, in_l_charge);
            Matrix const& in_block = m1[block];
            Matrix& out_block = m2(out_l_charge, out_r_charge);
            size_t sdim = physical_i[s].second, ldim = left_i[l].second, rdim = right_i[r].second;
            for (size_t ss = 0; ss < sdim; ++ss)
                for (size_t rr = 0; rr < rdim; ++rr)
                    std::memcpy(&out_block(out_left_offset + ss * ldim, rr), &in_block(0, in_right_offset + ss * rdim + rr), sizeof(double) * ldim);
        }
    }
}

class MPSTensor
{
public:
    Index phys_i, left_i, right_i;

    MPSTensor() {}
    // mpstensor.hpp:24-60
    MPSTensor(Index const& sd, Index const& ld, Index const& rd, std::function<double()> gen)
        : phys_i(sd), left_i(ld), right_i(rd), cur_storage(LeftPaired)
    {
        Index lb = sd * ld, rb = rd;
        common_subset(lb, rb);
        right_i = rb;
        Index possible_rp = adjoin(phys_i) * right_i, ltemp = ld;
        common_subset(ltemp, possible_rp);
        left_i = ltemp;
        lb.sort(); rb.sort(); left_i.sort(); right_i.sort();
        data_ = block_matrix(lb, rb);
        data_.generate(gen);
    }
    // mpstensor.hpp:62-98
    MPSTensor(Index const& sd, Index const& ld, Index const& rd, block_matrix const& block, MPSStorageLayout layout)
        : phys_i(sd), left_i(ld), right_i(rd), data_(block), cur_storage(layout)
    {
        if (cur_storage == LeftPaired) {
            Index new_right_i = data_.right_basis();
            Index possible_left_i = adjoin(phys_i) * new_right_i;
            Index old_left_i = left_i;
            common_subset(old_left_i, possible_left_i);
            std::swap(right_i, new_right_i);
            std::swap(left_i, old_left_i);
        } else {
            Index new_left_i = data_.left_basis();
            Index possible_right_i = phys_i * new_left_i;
            Index old_right_i = right_i;
            common_subset(old_right_i, possible_right_i);
            std::swap(left_i, new_left_i);
            std::swap(right_i, old_right_i);
        }
    }
    // member-wise assembly, as site_hamil2 does for its result (abelian/site_hamil.hpp:48-51): indices untouched
    MPSTensor(Index const& sd, Index const& ld, Index const& rd, block_matrix const& block, MPSStorageLayout layout, bool /*raw*/)
        : phys_i(sd), left_i(ld), right_i(rd), data_(block), cur_storage(layout) {}
    Index const& site_dim() const { return phys_i; }
    Index const& row_dim() const { return left_i; }
    Index const& col_dim() const { return right_i; }
    block_matrix& data() { return data_; }
    block_matrix const& data() const { return data_; }
    MPSStorageLayout storage() const { return cur_storage; }

    void make_left_paired() const
    {
        if (cur_storage == LeftPaired) return;
        block_matrix tmp;
        reshape_right_to_left_new(phys_i, left_i, right_i, data_, tmp);
        cur_storage = LeftPaired;
        std::swap(data_, tmp);
    }
    void make_right_paired() const
    {
        if (cur_storage == RightPaired) return;
        block_matrix tmp;
        reshape_left_to_right_new(phys_i, left_i, right_i, data_, tmp);
        cur_storage = RightPaired;
        std::swap(data_, tmp);
    }
    double scalar_norm() const { return std::sqrt(data_.norm_square()); }
    double scalar_overlap(MPSTensor const& rhs) const
    {
        make_left_paired(); rhs.make_left_paired();
        return data_.scalar_overlap(rhs.data_);
    }
    void multiply_by_scalar(double a) { data_ *= a; }
    void divide_by_scalar(double a) { data_ *= (1. / a); }
    // replace_left_paired / replace_right_paired (mpstensor.hpp:100-134)
    void replace_left_paired(block_matrix const& rhs)
    {
        make_left_paired();
        Index new_right_i = rhs.right_basis();
        Index possible_left_i = adjoin(phys_i) * new_right_i;
        Index old_left_i = left_i;
        common_subset(old_left_i, possible_left_i);
        std::swap(right_i, new_right_i);
        std::swap(left_i, old_left_i);
        data_ = rhs;
    }
    void replace_right_paired(block_matrix const& rhs)
    {
        make_right_paired();
        Index new_left_i = rhs.left_basis();
        Index possible_right_i = phys_i * new_left_i;
        Index old_right_i = right_i;
        common_subset(old_right_i, possible_right_i);
        std::swap(left_i, new_left_i);
        std::swap(right_i, old_right_i);
        data_ = rhs;
    }

private:
    mutable block_matrix data_;
    mutable MPSStorageLayout cur_storage = LeftPaired;
};

// A Boundary may live in HBM: `device_mirror` is an opaque handle owned by the GPU engine (qcm::GpuEngine),
// `host_valid` tells whether the dense blocks on the host hold the data (false: structure-only shells whose
// contents are fetched with GpuEngine::download).  Mutable access drops the mirror, as the reference's
// storage layer invalidates a boundary that is assigned to or dropped (utils/storage.h:176-181,363-365).
class Boundary
{
public:
    Boundary() {}
    Boundary(Index const& ud, Index const& ld, size_t ad = 1) : data_(ad, block_matrix(ud, ld)) {}
    size_t aux_dim() const { return data_.size(); }
    void resize(size_t n) { data_.resize(n); device_mirror.reset(); }
    block_matrix& operator[](size_t k) { device_mirror.reset(); return data_[k]; }
    block_matrix const& operator[](size_t k) const { return data_[k]; }
    block_matrix& raw(size_t k) { return data_[k]; }   // engine-internal: does not invalidate the mirror
    std::vector<double> traces() const { std::vector<double> r; for (auto const& b : data_) r.push_back(b.trace()); return r; }
    size_t num_elements() const { size_t r = 0; for (auto const& b : data_) r += b.num_elements(); return r; }

    mutable std::shared_ptr<void> device_mirror;
    bool host_valid = true;

private:
    std::vector<block_matrix> data_;
};

struct MPS : public std::vector<MPSTensor>
{
    size_t length() const { return size(); }
    // mps.hpp:244-268: one entry, all-ones blocks on the diagonal charges
    Boundary left_boundary() const
    {
        Index i = (*this)[0].row_dim();
        Boundary ret(i, i, 1);
        for (size_t k = 0; k < ret[0].n_blocks(); ++k) for (auto& x : ret[0][k].v) x = 1.;
        return ret;
    }
    Boundary right_boundary() const
    {
        Index i = (*this)[size() - 1].col_dim();
        Boundary ret(i, i, 1);
        for (size_t k = 0; k < ret[0].n_blocks(); ++k) for (auto& x : ret[0][k].v) x = 1.;
        return ret;
    }
};

// charge_detail.h
inline bool charge_physical(SymmKind k, Charge const& c)
{
    if (k == TWOU1PG) return c[0] >= 0 && c[1] >= 0;
    if (k == TWOU1) return c[0] >= 0 && c[1] >= 0;    // NU1_template<N> specialisation
    return spin(c) >= 0;
}
inline bool charge_has_less_particles(SymmKind k, Charge const& a, Charge const& ref)
{
    if (k == TWOU1PG || k == TWOU1) return a[0] <= ref[0] && a[1] <= ref[1];
    return true;
}

// mps_sectors.h:26-111
inline std::vector<Index> allowed_sectors(SymmKind symm, std::vector<int> const& site_type, std::vector<Index> const& phys_dims,
                                          Charge right_end, size_t Mmax)
{
    size_t L = site_type.size();
    std::vector<Charge> maximum_charges(phys_dims.size()), minimum_charges(phys_dims.size());
    for (size_t type = 0; type < phys_dims.size(); ++type) {
        Index physc = phys_dims[type];
        physc.sort();
        maximum_charges[type] = physc.begin()->first;
        minimum_charges[type] = (physc.end() - 1)->first;
        if (minimum_charges[type] > maximum_charges[type]) std::swap(maximum_charges[type], minimum_charges[type]);
    }
    Charge maximum_total_charge, minimum_total_charge;
    for (size_t i = 0; i < L; ++i) {
        maximum_total_charge = fuse(maximum_total_charge, maximum_charges[site_type[i]]);
        minimum_total_charge = fuse(minimum_total_charge, minimum_charges[site_type[i]]);
    }
    Index l_triv, r_triv;
    l_triv.insert(std::make_pair(Charge(), (size_t)1));
    r_triv.insert(std::make_pair(right_end, (size_t)1));
    std::vector<Index> left_allowed(L + 1), right_allowed(L + 1), allowed(L + 1);
    left_allowed[0] = l_triv;
    right_allowed[L] = r_triv;
    Charge cmaxi = maximum_total_charge, cmini = minimum_total_charge;
    for (size_t i = 1; i < L + 1; ++i) {
        left_allowed[i] = phys_dims[site_type[i - 1]] * left_allowed[i - 1];
        cmaxi = fuse(cmaxi, -maximum_charges[site_type[i - 1]]);
        cmini = fuse(cmini, -minimum_charges[site_type[i - 1]]);
        auto it = left_allowed[i].begin();
        while (it != left_allowed[i].end()) {
            if (fuse(it->first, cmaxi) < right_end) it = left_allowed[i].erase(it);
            else if (fuse(it->first, cmini) > right_end) it = left_allowed[i].erase(it);
            else if (!charge_physical(symm, it->first) && charge_has_less_particles(symm, it->first, right_end)) it = left_allowed[i].erase(it);
            else { it->second = std::min(Mmax, it->second); ++it; }
        }
    }
    cmaxi = maximum_total_charge; cmini = minimum_total_charge;
    for (int i = (int)L - 1; i >= 0; --i) {
        right_allowed[i] = adjoin(phys_dims[site_type[i]]) * right_allowed[i + 1];
        cmaxi = fuse(cmaxi, -maximum_charges[site_type[i]]);
        cmini = fuse(cmini, -minimum_charges[site_type[i]]);
        auto it = right_allowed[i].begin();
        while (it != right_allowed[i].end()) {
            if (fuse(it->first, -cmaxi) > Charge()) it = right_allowed[i].erase(it);
            else if (fuse(it->first, -cmini) < Charge()) it = right_allowed[i].erase(it);
            else if (!charge_physical(symm, it->first) && charge_has_less_particles(symm, it->first, right_end)) it = right_allowed[i].erase(it);
            else { it->second = std::min(Mmax, it->second); ++it; }
        }
        // extract_common_subset (indexing_stable.hpp:519-531)
        common_subset(left_allowed[i], right_allowed[i]);
    }
    for (size_t i = 0; i < L + 1; ++i) {
        allowed[i] = common_subset(left_allowed[i], right_allowed[i]);
        for (auto it = allowed[i].begin(); it != allowed[i].end(); ++it)
            it->second = std::min(Mmax, std::min(left_allowed[i].size_of_block(it->first), right_allowed[i].size_of_block(it->first)));
    }
    return allowed;
}

// uniform [0,1) stream in the spirit of dmrg_random::uniform (mt19937 + uniform_real, utils/random.hpp:14-29)
struct UniformGen
{
    std::mt19937 eng;
    explicit UniformGen(unsigned seed = 42) : eng(seed) {}
    double operator()() { return (double)eng() / 4294967296.0; }
};

// default_mps_init::init_sectors (mps_initializers.h:62-77): fillrand = random, otherwise constant `val`
inline MPS make_mps(SymmKind symm, std::vector<int> const& site_type, std::vector<Index> const& phys_dims, Charge right_end,
                    size_t Mmax, bool fillrand, double val, unsigned seed = 42)
{
    std::vector<Index> allowed = allowed_sectors(symm, site_type, phys_dims, right_end, Mmax);
    UniformGen gen(seed);
    MPS mps;
    for (size_t i = 0; i < site_type.size(); ++i) {
        std::function<double()> g;
        if (fillrand) g = [&gen]() { return gen(); }; else g = [val]() { return val; };
        MPSTensor t(phys_dims[site_type[i]], allowed[i], allowed[i + 1], g);
        t.divide_by_scalar(t.scalar_norm());
        mps.push_back(t);
    }
    return mps;
}

} // namespace qcm
