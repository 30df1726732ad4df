// Index / DualIndex / ProductBasis: the quantum-number bookkeeping of symmetry-blocked tensors.
//
// Semantics follow the reference's containers so that block k of every block matrix is the same
// (lc, rc) sector as in QCMaquis:
//   Index         dmrg/block_matrix/indexing_stable.hpp:81-300   (sorted DESCENDING, lower_bound lookup)
//   DualIndex     dmrg/block_matrix/dual_index.h:122-338         (sorted descending by (lc, rc))
//   ProductBasis  dmrg/block_matrix/indexing_stable.hpp:304-375  (offsets = running sums in loop order a, b)
//   operator*, adjoin, common_subset   indexing_stable.hpp:458-547
#pragma once
#include "symmetry.hpp"
#include <algorithm>
#include <cassert>
#include <unordered_map>
#include <utility>
#include <vector>

namespace qcm {

class Index
{
public:
    typedef std::pair<Charge, size_t> value_type;
    typedef std::vector<value_type>::iterator iterator;
    typedef std::vector<value_type>::const_iterator const_iterator;

    Index() {}
    explicit Index(size_t n) : data_(n) {}

    size_t position(Charge const& c) const
    {
        auto it = std::lower_bound(data_.begin(), data_.end(), c,
                                   [](value_type const& a, Charge const& b) { return a.first > b; });
        if (it != data_.end() && it->first != c) it = data_.end();
        return it - data_.begin();
    }
    bool has(Charge const& c) const { return position(c) != data_.size(); }
    size_t size_of_block(Charge const& c) const
    {
        size_t p = position(c);
        if (p == data_.size()) throw std::runtime_error("Index::size_of_block: charge not present");
        return data_[p].second;
    }
    size_t size_of_block(Charge const& c, bool) const
    {
        size_t p = position(c);
        return p == data_.size() ? 0 : data_[p].second;
    }
    void sort()
    {
        std::sort(data_.begin(), data_.end(), [](value_type const& a, value_type const& b) { return a.first > b.first; });
    }
    // sorted insert: in front of the first element that is smaller (indexing_stable.hpp:141-151,270-276)
    size_t insert(value_type const& x)
    {
        size_t d = std::find_if(data_.begin(), data_.end(), [&](value_type const& a) { return a.first < x.first; }) - data_.begin();
        data_.insert(data_.begin() + d, x);
        return d;
    }
    size_t sum_of_sizes() const { size_t r = 0; for (auto const& e : data_) r += e.second; return r; }
    bool operator==(Index const& o) const { return data_ == o.data_; }
    bool operator!=(Index const& o) const { return !(*this == o); }

    iterator begin() { return data_.begin(); }
    iterator end() { return data_.end(); }
    const_iterator begin() const { return data_.begin(); }
    const_iterator end() const { return data_.end(); }
    value_type& operator[](size_t p) { return data_[p]; }
    value_type const& operator[](size_t p) const { return data_[p]; }
    size_t size() const { return data_.size(); }
    iterator erase(iterator p) { return data_.erase(p); }
    iterator erase(iterator a, iterator b) { return data_.erase(a, b); }

private:
    std::vector<value_type> data_;
};

// indexing_stable.hpp:497-517
inline Index operator*(Index const& i1, Index const& i2)
{
    Index ret;
    for (auto const& a : i1)
        for (auto const& b : i2) {
            Charge pdc = fuse(a.first, b.first);
            size_t ps = a.second * b.second;
            size_t match = ret.position(pdc);
            if (match < ret.size()) ret[match].second += ps;
            else ret.insert(std::make_pair(pdc, ps));
        }
    ret.sort();
    return ret;
}

// indexing_stable.hpp:458-480: negate charges, keep sizes, sorted descending
inline Index adjoin(Index const& inp)
{
    Index ret;
    for (auto const& e : inp) ret.insert(std::make_pair(-e.first, e.second));
    ret.sort();
    return ret;
}

// indexing_stable.hpp:533-547: NOTE both arguments are trimmed
inline Index common_subset(Index& a, Index& b)
{
    a.erase(std::remove_if(a.begin(), a.end(), [&](Index::value_type const& e) { return !b.has(e.first); }), a.end());
    b.erase(std::remove_if(b.begin(), b.end(), [&](Index::value_type const& e) { return !a.has(e.first); }), b.end());
    return a;
}

struct QnBlock
{
    Charge lc, rc;
    size_t ls, rs;
    QnBlock() : ls(0), rs(0) {}
    QnBlock(Charge l, Charge r, size_t a, size_t b) : lc(l), rc(r), ls(a), rs(b) {}
    bool operator==(QnBlock const& o) const { return lc == o.lc && rc == o.rc && ls == o.ls && rs == o.rs; }
};

class DualIndex
{
public:
    typedef QnBlock value_type;
    typedef std::vector<QnBlock>::const_iterator const_iterator;

    static bool gt(QnBlock const& a, QnBlock const& b)
    {
        if (a.lc > b.lc) return true;
        if (a.lc < b.lc) return false;
        return a.rc > b.rc;
    }
    size_t position(Charge const& row, Charge const& col) const
    {
        QnBlock probe(row, col, 0, 0);
        auto it = std::lower_bound(data_.begin(), data_.end(), probe, gt);
        if (it != data_.end() && (it->lc != row || it->rc != col)) it = data_.end();
        return it - data_.begin();
    }
    bool has(Charge const& row, Charge const& col) const { return position(row, col) != data_.size(); }
    const_iterator left_lower_bound(Charge const& row) const
    {
        return std::lower_bound(data_.begin(), data_.end(), row, [](QnBlock const& a, Charge const& r) { return a.lc > r; });
    }
    bool left_has(Charge const& row) const
    {
        auto it = left_lower_bound(row);
        return it != data_.end() && it->lc == row;
    }
    // sorted insert (dual_index.h:199-209,311-317): in front of the first element that is smaller
    size_t insert(QnBlock const& x)
    {
        size_t d = std::find_if(data_.begin(), data_.end(), [&](QnBlock const& a) {
                       if (a.lc < x.lc) return true;
                       if (a.lc > x.lc) return false;
                       return a.rc < x.rc;
                   }) - data_.begin();
        data_.insert(data_.begin() + d, x);
        return d;
    }
    void push_back_unsorted(QnBlock const& x) { data_.push_back(x); }
    bool operator==(DualIndex const& o) const { return data_ == o.data_; }
    bool operator!=(DualIndex const& o) const { return !(*this == o); }

    Charge const& left_charge(size_t k) const { return data_[k].lc; }
    Charge const& right_charge(size_t k) const { return data_[k].rc; }
    size_t left_size(size_t k) const { return data_[k].ls; }
    size_t right_size(size_t k) const { return data_[k].rs; }
    QnBlock& operator[](size_t k) { return data_[k]; }
    QnBlock const& operator[](size_t k) const { return data_[k]; }
    size_t size() const { return data_.size(); }
    const_iterator begin() const { return data_.begin(); }
    const_iterator end() const { return data_.end(); }
    void erase(size_t k) { data_.erase(data_.begin() + k); }
    void clear() { data_.clear(); }

    Index left_basis() const
    {
        Index r(data_.size());
        for (size_t s = 0; s < data_.size(); ++s) r[s] = std::make_pair(data_[s].lc, data_[s].ls);
        return r;
    }
    Index right_basis() const
    {
        Index r(data_.size());
        for (size_t s = 0; s < data_.size(); ++s) r[s] = std::make_pair(data_[s].rc, data_[s].rs);
        return r;
    }
    DualIndex transposed() const
    {
        DualIndex r;
        for (auto const& b : data_) r.insert(QnBlock(b.rc, b.lc, b.rs, b.ls));
        return r;
    }

private:
    std::vector<QnBlock> data_;
};

// ProductBasis(a, b [, right-pairing fusion]) -- offsets are running sums per fused charge in the
// nested loop order "for a: for b:" (indexing_stable.hpp:326-342). `minus_first` selects the fusion
// f(s, r) = fuse(-s, r) used for right pairing (e.g. contractions/abelian/site_hamil.hpp:44-46).
class ProductBasis
{
public:
    ProductBasis() {}
    ProductBasis(Index const& a, Index const& b, bool minus_first = false) : minus_first_(minus_first)
    {
        for (auto const& x : a)
            for (auto const& y : b) {
                Charge pc = fused(x.first, y.first);
                size_t& s = size_[pc];
                keys_vals_[std::make_pair(x.first, y.first)] = s;
                s += x.second * y.second;
            }
    }
    size_t operator()(Charge const& a, Charge const& b) const
    {
        auto it = keys_vals_.find(std::make_pair(a, b));
        if (it == keys_vals_.end()) throw std::runtime_error("ProductBasis: pair not found");
        return it->second;
    }
    bool has(Charge const& a, Charge const& b) const { return keys_vals_.count(std::make_pair(a, b)) > 0; }
    size_t size(Charge const& pc) const
    {
        auto it = size_.find(pc);
        if (it == size_.end()) throw std::runtime_error("ProductBasis: fused charge not found");
        return it->second;
    }
    // size of the fused sector containing (a, b) under the plain fusion (indexing_stable.hpp:355-367)
    size_t size(Charge const& a, Charge const& b) const { return size(fuse(a, b)); }

private:
    Charge fused(Charge const& a, Charge const& b) const { return minus_first_ ? fuse(-a, b) : fuse(a, b); }
    bool minus_first_ = false;
    std::unordered_map<Charge, size_t, ChargeHash> size_;
    std::unordered_map<std::pair<Charge, Charge>, size_t, ChargePairHash> keys_vals_;
};

} // namespace qcm
