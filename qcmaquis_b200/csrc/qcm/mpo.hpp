// MPOTensor / MPO containers, the Hermitian-pair map, the tagged MPO builder and the two-site MPO fusion.
//   MPOTensor         dmrg/mp_tensors/mpotensor.h:23-107, mpotensor.hpp:10-65
//   Hermitian         dmrg/mp_tensors/mpotensor_detail.h:118-168
//   TaggedMPOMaker    dmrg/models/generate_mpo/tagged_mpo_maker_optim.hpp:30-739
//   make_twosite_mpo  dmrg/mp_tensors/ts_ops.h:36-126
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "site_operator.hpp"
#include <array>
#include <atomic>
#include <numeric>
#include <set>

namespace qcm {

struct Hermitian
{
    std::vector<size_t> LeftHerm, RightHerm;
    std::vector<int> LeftPhase, RightPhase;
    Hermitian(size_t ld = 1, size_t rd = 1) : LeftHerm(ld), RightHerm(rd), LeftPhase(ld, 1), RightPhase(rd, 1)
    {
        std::iota(LeftHerm.begin(), LeftHerm.end(), 0);
        std::iota(RightHerm.begin(), RightHerm.end(), 0);
    }
    Hermitian(std::vector<size_t> lh, std::vector<size_t> rh, std::vector<int> lp, std::vector<int> rp)
        : LeftHerm(std::move(lh)), RightHerm(std::move(rh)), LeftPhase(std::move(lp)), RightPhase(std::move(rp)) {}
    bool left_skip(size_t b) const { return LeftHerm[b] < b; }
    bool right_skip(size_t b) const { return RightHerm[b] < b; }
    size_t left_conj(size_t b) const { return LeftHerm[b]; }
    size_t right_conj(size_t b) const { return RightHerm[b]; }
    int left_phase(size_t b) const { return LeftPhase[b]; }
    int right_phase(size_t b) const { return RightPhase[b]; }
    size_t left_size() const { return LeftHerm.size(); }
    size_t right_size() const { return RightHerm.size(); }
};
inline Hermitian operator*(Hermitian const& a, Hermitian const& b)
{
    return Hermitian(a.LeftHerm, b.RightHerm, a.LeftPhase, b.RightPhase);
}

struct PreTerm { size_t b1, b2; tag_type tag; double scale; };   // boost::tuple<size_t,size_t,tag,scale>

class MPOTensor
{
public:
    typedef std::vector<std::pair<tag_type, double>> terms_type;

    MPOTensor() : herm_info(1, 1), uid_(next_uid()) {}
    // mpotensor.hpp:10-65
    MPOTensor(size_t ld, size_t rd, std::vector<PreTerm> tags, std::shared_ptr<OPTable> tbl, Hermitian h,
              std::vector<SpinDescriptor> lspins, std::vector<SpinDescriptor> rspins, bool su2)
        : herm_info(ld, rd), left_i(ld), right_i(rd), left_spins(std::move(lspins)), right_spins(std::move(rspins)),
          operator_table(std::move(tbl)), uid_(next_uid())
    {
        row_index.resize(ld);
        col_ptr.assign(rd + 1, 0);
        // CSC order; terms sharing (b1,b2) are prepended one by one as in the reference
        std::stable_sort(tags.begin(), tags.end(), [](PreTerm const& i, PreTerm const& j) {
            if (i.b2 != j.b2) return i.b2 < j.b2;
            return i.b1 < j.b1;
        });
        for (auto const& t : tags) {
            if (row_idx.empty() || entry_col.back() != t.b2 || row_idx.back() != t.b1) {
                row_idx.push_back(t.b1); entry_col.push_back(t.b2);
                terms.push_back(terms_type(1, std::make_pair(t.tag, t.scale)));
                row_index[t.b1].insert(t.b2);
            } else
                terms.back().insert(terms.back().begin(), std::make_pair(t.tag, t.scale));
        }
        for (size_t e = 0; e < entry_col.size(); ++e) col_ptr[entry_col[e] + 1]++;
        for (size_t c = 0; c < rd; ++c) col_ptr[c + 1] += col_ptr[c];
        if (operator_table)
            for (auto& op : *operator_table) op.update_sparse(su2);
        else
            operator_table.reset(new OPTable());
        row_non_zeros.assign(ld, 0); col_non_zeros.assign(rd, 0);
        for (size_t e = 0; e < row_idx.size(); ++e) { row_non_zeros[row_idx[e]]++; col_non_zeros[entry_col[e]]++; }
        num_one_rows_ = std::count(row_non_zeros.begin(), row_non_zeros.end(), (size_t)1);
        num_one_cols_ = std::count(col_non_zeros.begin(), col_non_zeros.end(), (size_t)1);
        if (h.left_size() == left_i && h.right_size() == right_i) herm_info = h;
    }

    size_t row_dim() const { return left_i; }
    size_t col_dim() const { return right_i; }
    // column(b2): entries [col_begin(b2), col_end(b2)) with row index row_of(e), ascending b1
    size_t col_begin(size_t b2) const { return col_ptr[b2]; }
    size_t col_end(size_t b2) const { return col_ptr[b2 + 1]; }
    size_t row_of(size_t e) const { return row_idx[e]; }
    std::set<size_t> const& row(size_t b1) const { return row_index[b1]; }
    size_t entry(size_t b1, size_t b2) const
    {
        auto b = row_idx.begin() + col_ptr[b2], e = row_idx.begin() + col_ptr[b2 + 1];
        auto it = std::lower_bound(b, e, b1);
        if (it == e || *it != b1) return (size_t)-1;
        return it - row_idx.begin();
    }
    bool has(size_t b1, size_t b2) const { return entry(b1, b2) != (size_t)-1; }
    terms_type const& at(size_t b1, size_t b2) const { return terms[entry(b1, b2)]; }
    terms_type const& at_entry(size_t e) const { return terms[e]; }
    SiteOperator const& op(tag_type t) const { return (*operator_table)[t]; }
    std::shared_ptr<OPTable> get_operator_table() const { return operator_table; }
    SpinDescriptor left_spin(size_t b) const { return left_spins[b]; }
    SpinDescriptor right_spin(size_t b) const { return right_spins[b]; }
    std::vector<SpinDescriptor> const& row_spin_dim() const { return left_spins; }
    std::vector<SpinDescriptor> const& col_spin_dim() const { return right_spins; }
    size_t num_row_non_zeros(size_t b) const { return row_non_zeros[b]; }
    size_t num_col_non_zeros(size_t b) const { return col_non_zeros[b]; }
    size_t num_one_rows() const { return num_one_rows_; }
    size_t num_one_cols() const { return num_one_cols_; }
    size_t nnz() const { return row_idx.size(); }
    Hermitian herm_info;
    // identity of the tensor's CONTENT: every constructed tensor gets a fresh number, copies keep it.  Plans bake the MPO
    // coefficients in, so plan caches key on this, never on the object's address.
    uint64_t uid() const { return uid_; }

private:
    size_t left_i = 1, right_i = 1;
    std::vector<SpinDescriptor> left_spins, right_spins;
    std::vector<size_t> row_non_zeros, col_non_zeros;
    size_t num_one_rows_ = 0, num_one_cols_ = 0;
    std::vector<size_t> col_ptr, row_idx, entry_col;
    std::vector<terms_type> terms;
    std::vector<std::set<size_t>> row_index;
    std::shared_ptr<OPTable> operator_table;
    uint64_t uid_ = 0;
    static uint64_t next_uid() { static std::atomic<uint64_t> n{1}; return n.fetch_add(1); }
};

struct MPO : public std::vector<MPOTensor>
{
    double core_energy = 0.;
    std::vector<size_t> herm_pairs;   // "MPO Bond p: dim/pairs" second number, per site
    double getCoreEnergy() const { return core_energy; }
};

// ---------------------------------------------------------------------------------------------------------
// model terms
struct Term : public std::vector<std::pair<int, tag_type>>
{
    double coeff = 1.;
    bool is_fermionic = false;
    int position(size_t i) const { return (*this)[i].first; }
    tag_type operator_tag(size_t i) const { return (*this)[i].second; }
    void canonical_order()
    {
        std::stable_sort(begin(), end(), [](value_type const& a, value_type const& b) { return a.first < b.first; });
    }
};

struct Lattice
{
    int L = 0;
    std::vector<int> irreps;   // "type" property per site (point-group irrep of the orbital)
    int size() const { return L; }
    int type(int p) const { return irreps[p]; }
    int max_type() const { return *std::max_element(irreps.begin(), irreps.end()) + 1; }
};

struct ModelBase
{
    SymmKind symm;
    Lattice lat;
    std::shared_ptr<TagHandler> tag_handler;
    std::vector<Index> phys_indices;    // per site type
    std::vector<Term> terms;
    std::vector<tag_type> ident, ident_full, fill;
    Charge total_charge;
    Index const& phys_dim(int type) const { return phys_indices[type]; }
};

// ---------------------------------------------------------------------------------------------------------
namespace mpo_detail {

struct Key   // prempo_key (tagged_mpo_maker_optim.hpp:33-64); pos_op never exceeds 2 entries for <= 4-operator terms
{
    enum { trivial_left = 0, bulk = 1, bulk_no_merge = 2, trivial_right = 3 };
    int kind = bulk;
    int n = 0;
    std::array<std::pair<int, tag_type>, 4> po;
    size_t offset = 0;
    Key(int k = bulk) : kind(k) {}
    void push_back(std::pair<int, tag_type> const& x) { po[n++] = x; }
    int cmp_pos(Key const& o) const
    {
        for (int i = 0; i < std::min(n, o.n); ++i) {
            if (po[i] < o.po[i]) return -1;
            if (o.po[i] < po[i]) return 1;
        }
        return n < o.n ? -1 : (n > o.n ? 1 : 0);
    }
    bool operator==(Key const& o) const
    {
        if (kind != o.kind) return false;
        if (kind == trivial_left || kind == trivial_right) return true;
        return cmp_pos(o) == 0 && offset == o.offset;
    }
    bool operator<(Key const& o) const
    {
        if (kind != o.kind) return kind < o.kind;
        int c = cmp_pos(o);
        return c == 0 ? offset < o.offset : c < 0;
    }
};
struct PairInverseLess   // compare_pair_inverse (dmrg/utils/utils.hpp:61-73): second key first
{
    bool operator()(std::pair<Key, Key> const& i, std::pair<Key, Key> const& j) const
    {
        if (i.second < j.second) return true;
        if (j.second < i.second) return false;
        return i.first < j.first;
    }
};

} // namespace mpo_detail

class TaggedMPOMaker
{
    typedef mpo_detail::Key Key;
    typedef std::pair<tag_type, double> Value;
    typedef std::multimap<std::pair<Key, Key>, Value, mpo_detail::PairInverseLess> prempo_map;
    enum merge_kind { attach, detach };

public:
    explicit TaggedMPOMaker(ModelBase const& model)
        : m(model), length(model.lat.size()), th(model.tag_handler), prempo(length), trivial_left(Key::trivial_left),
          trivial_right(Key::trivial_right), leftmost_right(length), rightmost_left(0)
    {
        for (auto const& t : model.terms) add_term(t);
    }

    void add_term(Term term)
    {
        term.canonical_order();
        switch (term.size()) {
            case 1: add_1term(term); break;
            case 2: add_2term(term); break;
            default: add_generic_term(term); break;
        }
        leftmost_right = std::min(leftmost_right, term.back().first);
        rightmost_left = std::max(rightmost_left, term.front().first);
    }

    // tagged_mpo_maker_optim.hpp:189-296
    MPO create_mpo()
    {
        if (!finalized) finalize();
        bool su2 = is_su2(m.symm);
        MPO mpo;
        std::map<Key, size_t> left;
        left[trivial_left] = 0;
        std::vector<SpinDescriptor> left_spins(1);
        std::vector<size_t> LeftHerm(1, 0);
        std::vector<int> LeftPhase(1, 1);
        for (int p = 0; p < length; ++p) {
            std::vector<PreTerm> pre_tensor; pre_tensor.reserve(prempo[p].size());
            std::map<Key, Key> HermKeyPairs;
            std::map<Key, std::pair<int, int>> HermitianPhases;
            std::map<Key, size_t> right;
            size_t r = 2;
            for (auto it = prempo[p].begin(); it != prempo[p].end(); ++it) {
                Key const& k1 = it->first.first; Key const& k2 = it->first.second;
                auto ll = left.find(k1);
                if (ll == left.end()) throw std::runtime_error("k1 not found!");
                auto rr = right.find(k2);
                if (k2 == trivial_left && rr == right.end()) rr = right.insert(std::make_pair(k2, (size_t)0)).first;
                else if (k2 == trivial_right && rr == right.end()) rr = right.insert(std::make_pair(k2, (size_t)1)).first;
                else if (rr == right.end()) rr = right.insert(std::make_pair(k2, r++)).first;
                size_t rr_dim = (p == length - 1) ? 0 : rr->second;
                pre_tensor.push_back(PreTerm{ll->second, rr_dim, it->second.first, it->second.second});
                std::pair<int, int> phase; Key ck2;
                std::tie(ck2, phase) = conjugate_key(k2, p);
                if (!(k2 == ck2)) { HermKeyPairs[k2] = ck2; HermitianPhases[k2] = phase; }
            }
            size_t ldim = 0, rdim = 0;
            for (auto const& t : pre_tensor) { ldim = std::max(ldim, t.b1 + 1); rdim = std::max(rdim, t.b2 + 1); }
            std::vector<SpinDescriptor> right_spins(rdim);
            for (auto const& t : pre_tensor) right_spins[t.b2] = couple(left_spins[t.b1], th->get_op(t.tag).spin());
            std::vector<size_t> RightHerm(rdim);
            std::vector<int> RightPhase(rdim, 1);
            size_t cnt = 0;
            std::iota(RightHerm.begin(), RightHerm.end(), 0);
            for (auto h_it = HermKeyPairs.begin(); h_it != HermKeyPairs.end(); ++h_it) {
                size_t romeo = right[h_it->first];
                size_t julia = right[h_it->second];
                if (romeo < julia) {
                    cnt++;
                    std::swap(RightHerm[romeo], RightHerm[julia]);
                    RightPhase[romeo] = HermitianPhases[h_it->first].first;
                    RightPhase[julia] = HermitianPhases[h_it->first].second;
                }
            }
            Hermitian h_(LeftHerm, RightHerm, LeftPhase, RightPhase);
            size_t ld = (p == 0) ? 1 : ldim, rd = (p == length - 1) ? 1 : rdim;
            mpo.push_back(MPOTensor(ld, rd, pre_tensor, th->table, h_, left_spins, right_spins, su2));
            std::swap(left, right);
            std::swap(left_spins, right_spins);
            std::swap(LeftHerm, RightHerm);
            std::swap(LeftPhase, RightPhase);
            mpo.herm_pairs.push_back(cnt);
        }
        mpo.core_energy = core_energy;
        return mpo;
    }

private:
    void add_1term(Term const& term)
    {
        if (term.operator_tag(0) == m.ident[m.lat.type(term.position(0))]) core_energy += term.coeff;
        else {
            SiteOperator op = th->get_op(term.operator_tag(0));
            op *= term.coeff;
            site_terms[term.position(0)] += op;
        }
    }
    void add_2term(Term const& term)
    {
        SpinDescriptor mpo_spin;
        int nferm = 0;
        for (int i = 0; i < 2; ++i) if (th->is_fermionic(term.operator_tag(i))) nferm++;
        bool trivial_fill = true;
        Key k1 = trivial_left;
        {
            mpo_spin = couple(mpo_spin, th->get_op(term.operator_tag(0)).spin());
            Key k2; k2.push_back(term[1]);
            k1 = insert_operator(term.position(0), std::make_pair(k1, k2), Value(term.operator_tag(0), term.coeff), detach);
            if (th->is_fermionic(term.operator_tag(0))) nferm--;
            trivial_fill = (nferm % 2 == 0);
        }
        insert_filling(term.position(0) + 1, term.position(1), k1, trivial_fill, mpo_spin.get() > 1);
        {
            mpo_spin = couple(mpo_spin, th->get_op(term.operator_tag(1)).spin());
            insert_operator(term.position(1), std::make_pair(k1, trivial_right), Value(term.operator_tag(1), 1.), attach);
        }
    }
    // tagged_mpo_maker_optim.hpp:447-520 (prefer_fork = true)
    void add_generic_term(Term const& term)
    {
        size_t nops = term.size();
        int nferm = 0;
        for (size_t i = 0; i < nops; ++i) if (th->is_fermionic(term.operator_tag(i))) nferm++;
        bool trivial_fill = true;
        SpinDescriptor mpo_spin;
        size_t thresh = nops / 2;
        Key k1 = trivial_left;
        Key ops_left;
        for (size_t i = 0; i < thresh; ++i) {
            mpo_spin = couple(mpo_spin, th->get_op(term.operator_tag(i)).spin());
            ops_left.push_back(term[i]);
            Key k2 = ops_left;
            k1 = insert_operator(term.position(i), std::make_pair(k1, k2), Value(term.operator_tag(i), 1.), attach);
            if (th->is_fermionic(term.operator_tag(i))) nferm--;
            trivial_fill = (nferm % 2 == 0);
            insert_filling(term.position(i) + 1, term.position(i + 1), k1, trivial_fill, mpo_spin.get() > 1);
        }
        Key k2;
        for (size_t j = thresh + 1; j < nops; j++) k2.push_back(term[j]);
        mpo_spin = couple(mpo_spin, th->get_op(term.operator_tag(thresh)).spin());
        k1 = insert_operator(term.position(thresh), std::make_pair(k1, k2), Value(term.operator_tag(thresh), term.coeff), detach);
        if (th->is_fermionic(term.operator_tag(thresh))) nferm--;
        trivial_fill = (nferm % 2 == 0);
        insert_filling(term.position(thresh) + 1, term.position(thresh + 1), k1, trivial_fill, mpo_spin.get() > 1);
        for (size_t i = thresh + 1; i < nops; i++) {
            Key k2m;
            if (i == nops - 1) k2m = trivial_right;
            else for (size_t j = i + 1; j < nops; j++) k2m.push_back(term[j]);
            mpo_spin = couple(mpo_spin, th->get_op(term.operator_tag(i)).spin());
            k1 = insert_operator(term.position(i), std::make_pair(k1, k2m), Value(term.operator_tag(i), 1.), attach);
            if (th->is_fermionic(term.operator_tag(i))) nferm--;
            if (i != nops - 1) {
                trivial_fill = (nferm % 2 == 0);
                insert_filling(term.position(i) + 1, term.position(i + 1), k1, trivial_fill, mpo_spin.get() > 1);
            }
        }
    }
    void insert_filling(int i, int j, Key const& k, bool trivial_fill, bool spin_larger_than_one)
    {
        for (; i < j; ++i) {
            int typei = m.lat.type(i);
            tag_type use_ident = spin_larger_than_one ? m.ident_full[typei] : m.ident[typei];
            tag_type op = trivial_fill ? use_ident : m.fill[typei];
            auto kk = std::make_pair(k, k);
            auto f = prempo[i].find(kk);
            if (f == prempo[i].end()) prempo[i].insert(std::make_pair(kk, Value(op, 1.)));
            else if (f->second != Value(op, 1.))
                throw std::runtime_error("Pre-existing term at site " + std::to_string(i));
        }
    }
    Key insert_operator(int p, std::pair<Key, Key> const& kk, Value const& val, merge_kind mk)
    {
        if (mk == detach) prempo[p].insert(std::make_pair(kk, val));
        else if (prempo[p].count(kk) == 0) prempo[p].insert(std::make_pair(kk, val));
        return kk.second;
    }
    void finalize()
    {
        auto kk = std::make_pair(trivial_left, trivial_right);
        for (auto const& st : site_terms) {
            tag_type site_tag = th->register_op(st.second, false);
            prempo[st.first].insert(std::make_pair(kk, Value(site_tag, 1.)));
        }
        for (int p = 0; p < rightmost_left; ++p)
            prempo[p].insert(std::make_pair(std::make_pair(trivial_left, trivial_left), Value(m.ident[m.lat.type(p)], 1.)));
        for (int p = leftmost_right + 1; p < length; ++p)
            prempo[p].insert(std::make_pair(std::make_pair(trivial_right, trivial_right), Value(m.ident[m.lat.type(p)], 1.)));
        finalized = true;
    }
    // tagged_mpo_maker_optim.hpp:674-721
    std::pair<Key, std::pair<int, int>> conjugate_key(Key const& k, int p)
    {
        auto np = [&](Charge const& c) { return particle_number(m.symm, c); };
        Key conj = k;
        for (int i = 0; i < k.n; ++i) conj.po[i].second = th->herm_conj(k.po[i].second);
        std::pair<int, int> phase(1, 1);
        if (k.n == 1) {
            SiteOperator const& op1 = th->get_op(k.po[0].second);
            if (p < k.po[0].first) {
                int pdiff = np(op1.basis().left_charge(0)) - np(op1.basis().right_charge(0));
                if (pdiff == 1) phase = std::make_pair(1, -1);
                else if (pdiff == -1) phase = std::make_pair(-1, 1);
            } else if (op1.spin().get() == 1)
                phase = std::make_pair(-1, 1);
        }
        if (k.n == 2) {
            SiteOperator const& op1 = th->get_op(k.po[0].second);
            SiteOperator const& op2 = th->get_op(k.po[1].second);
            int d1 = np(op1.basis().left_charge(0)) - np(op1.basis().right_charge(0));
            int d2 = np(op2.basis().left_charge(0)) - np(op2.basis().right_charge(0));
            if (op1.spin().get() == 1 && op2.spin().get() == 1 && op2.spin().action() == -1 && d1 == -d2) phase = std::make_pair(-1, -1);
            if (op1.spin().get() == 1 && op2.spin().get() == 1 && op2.spin().action() == 1 && d1 == d2) phase = std::make_pair(-1, -1);
        }
        return std::make_pair(conj, phase);
    }

    ModelBase const& m;
    int length;
    std::shared_ptr<TagHandler> th;
    std::vector<prempo_map> prempo;
    Key trivial_left, trivial_right;
    std::map<int, SiteOperator> site_terms;
    int leftmost_right, rightmost_left;
    bool finalized = false;
    double core_energy = 0.;
};

// ---------------------------------------------------------------------------------------------------------
// two-site MPO fusion (ts_ops.h:36-126)
inline MPOTensor make_twosite_mpo(SymmKind symm, MPOTensor const& mpo1, MPOTensor const& mpo2, Index const& phys_i1, Index const& phys_i2)
{
    bool su2 = is_su2(symm);
    std::shared_ptr<OPTable> kron_table(new OPTable());
    std::vector<PreTerm> prempo;
    // The reference loops over all (b1, b3) pairs and collects the b2 that connect them (ts_ops.h:99-116).  Same
    // products in the same order here, but found by walking row b1 of the first tensor and row b2 of the second (the
    // pairs without a connecting b2 are never visited), and the rows b1 are fused in parallel; the fused operators are
    // then registered serially in (b1, b3, spin) order so that the tags come out exactly as in the reference's loop.
    const long n1 = (long)mpo1.row_dim();
    auto tdbg0 = std::chrono::steady_clock::now();
    std::vector<std::vector<std::pair<size_t, std::map<int, SiteOperator>>>> fused((size_t)n1);
#pragma omp parallel for schedule(dynamic, 4)
    for (long b1l = 0; b1l < n1; ++b1l) {
        const size_t b1 = (size_t)b1l;
        std::map<size_t, std::map<int, SiteOperator>> acc;
        for (size_t b2 : mpo1.row(b1)) {
            auto const& p1 = mpo1.at(b1, b2);
            for (size_t b3 : mpo2.row(b2)) {
                auto const& p2 = mpo2.at(b2, b3);
                std::map<int, SiteOperator>& coupled = acc[b3];
                for (auto const& t1 : p1)
                    for (auto const& t2 : p2) {
                        SiteOperator const& o1 = mpo1.op(t1.first); SiteOperator const& o2 = mpo2.op(t2.first);
                        if (su2) {
                            std::vector<int> op_spins = allowed_spins(mpo1.left_spin(b1).get(), mpo2.right_spin(b3).get(),
                                                                      o1.spin().get(), o2.spin().get());
                            for (int s : op_spins) {
                                SiteOperator product;
                                op_kron_su2(phys_i1, phys_i2, o1, o2, product, mpo1.left_spin(b1), mpo1.right_spin(b2), mpo2.right_spin(b3), s);
                                remove_empty_blocks(product);
                                product *= t1.second * t2.second;
                                coupled[s] += product;
                            }
                        } else {
                            // abelian groups: allowed_spins(0,0,0,0) == {0}
                            SiteOperator product;
                            op_kron_abelian(phys_i1, phys_i2, o1, o2, product);
                            remove_empty_blocks(product);
                            product *= t1.second * t2.second;
                            coupled[0] += product;
                        }
                    }
            }
        }
        fused[b1].assign(std::make_move_iterator(acc.begin()), std::make_move_iterator(acc.end()));
    }
    for (size_t b1 = 0; b1 < (size_t)n1; ++b1)
        for (auto& e : fused[b1])
            for (auto& kv : e.second) {
                kron_table->push_back(std::move(kv.second));
                prempo.push_back(PreTerm{b1, e.first, (tag_type)kron_table->size() - 1, 1.0});
            }
    auto tdbg1 = std::chrono::steady_clock::now();
    MPOTensor ret(mpo1.row_dim(), mpo2.col_dim(), prempo, kron_table, mpo1.herm_info * mpo2.herm_info,
                  mpo1.row_spin_dim(), mpo2.col_spin_dim(), su2);
    if (getenv("QCM_PLAN_TIMING"))
        fprintf(stderr, "  [two-site mpo] fuse + register %.3f s, tensor construction %.3f s\n", std::chrono::duration<double>(tdbg1 - tdbg0).count(),
                std::chrono::duration<double>(std::chrono::steady_clock::now() - tdbg1).count());
    return ret;
}

} // namespace qcm
