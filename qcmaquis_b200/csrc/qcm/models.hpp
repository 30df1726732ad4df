// Quantum-chemistry models: integral parsing, elementary operators and Hamiltonian term generation.
// These run once per calculation; they are restated so that the MPO that reaches the hot path has the same
// bond indexing, operator tags, spins and Hermitian pairs as the reference's.
//   parse_integrals        dmrg/models/chem/parse_integrals.h:40-213, dmrg/utils/align.h:17-28
//   qc_su2 (SU2U1[PG])     dmrg/models/chem/su2u1/model.hpp:20-300, term_maker.h, chem_helper.h
//   qc_model (TwoU1[PG])   dmrg/models/chem/2u1/model.hpp:15-456, term_maker.h, chem_helper.h
#pragma once
#include "mpo.hpp"
#include <fstream>
#include <sstream>

namespace qcm {

struct Integral { int idx[4]; double val; };

struct ModelParams
{
    SymmKind symm = SU2U1;
    int L = 0;
    std::vector<int> site_types;        // irrep per orbital (all 0 without point group)
    std::vector<Integral> integrals;    // 1-based indices, FCIDUMP convention (0 = absent index)
    double integral_cutoff = 0.;      // DmrgParameters.h:116
    int nelec = 0, spin = 0, irrep = 0; // SU2 groups
    int nup = 0, ndown = 0;             // TwoU1 groups (u1_total_charge1/2)
};

inline std::array<int, 4> align_idx(int i, int j, int k, int l)
{
    if (i < j) std::swap(i, j);
    if (k < l) std::swap(k, l);
    if (i < k) { std::swap(i, k); std::swap(j, l); }
    if (i == k && j < l) std::swap(j, l);
    return {i, j, k, l};
}

// FCIDUMP text: 4 header lines skipped exactly as parse_integrals.h:100-102 does
inline std::vector<Integral> read_fcidump(std::string const& path)
{
    std::ifstream f(path);
    if (!f) throw std::runtime_error("integral_file " + path + " does not exist");
    std::string line;
    for (int i = 0; i < 4; ++i) std::getline(f, line);
    std::vector<Integral> r;
    Integral t;
    while (f >> t.val >> t.idx[0] >> t.idx[1] >> t.idx[2] >> t.idx[3]) r.push_back(t);
    return r;
}

typedef std::array<int, 4> IndexTuple;
typedef std::array<int, 6> SixTuple;
typedef std::array<int, 8> EightTuple;

struct ParsedIntegrals
{
    std::vector<IndexTuple> idx;     // aligned, 0-based, -1 = absent
    std::vector<double> val;
};
inline ParsedIntegrals parse_integrals(ModelParams const& p)
{
    ParsedIntegrals r;
    for (auto const& t : p.integrals)
        if (std::abs(t.val) > p.integral_cutoff) {
            r.val.push_back(t.val);
            r.idx.push_back(align_idx(t.idx[0] - 1, t.idx[1] - 1, t.idx[2] - 1, t.idx[3] - 1));
        }
    return r;
}

inline bool perm_sgn(int i, int j, int k, int l)
{
    int idx[4] = {i, j, k, l}, inv = 0;
    for (int a = 0; a < 3; ++a) for (int b = a + 1; b < 4; ++b) if (idx[a] > idx[b]) inv++;
    return inv % 2 != 0;
}

inline std::vector<SiteOperator> site_specific_ops(SymmKind symm, SiteOperator const& op, int max_irrep)
{
    std::vector<SiteOperator> ret;
    for (int sc = 0; sc <= max_irrep; ++sc) {
        SiteOperator mod;
        mod.spin() = op.spin();
        // PGDecorator (chem/pg_util.h): odd-particle charges carry the orbital irrep
        for (size_t b = 0; b < op.n_blocks(); ++b) {
            Charge l = op.basis().left_charge(b), r = op.basis().right_charge(b);
            if (has_pg(symm)) {
                if (particle_number(symm, l) % 2 != 0) l[2] = sc;
                if (particle_number(symm, r) % 2 != 0) r[2] = sc;
            }
            mod.bm.insert_block(op[b], l, r);
        }
        ret.push_back(mod);
    }
    return ret;
}

// =========================================================================================================
// SU2U1 / SU2U1PG
class ModelSU2 : public ModelBase
{
    typedef std::vector<tag_type> tag_vec;
    struct Bundle { tag_vec couple_up, couple_down, fill_couple_up, fill_couple_down, no_couple, fill_no_couple; };
    struct Ops {
        tag_vec create_fill, create, destroy_fill, destroy, create_fill_couple_down, destroy_fill_couple_down,
            create_couple_up, destroy_couple_up, create_fill_count, create_count, destroy_fill_count, destroy_count,
            count, docc, e2d, d2e, flip_S0, flip_to_S2, flip_to_S0, ident, ident_full, fill, count_fill;
    } ops;
    struct Collection { Bundle ident, ident_full, fill, create, destroy, count, flip, e2d, d2e, docc, create_count, destroy_count; } oc;

    std::map<IndexTuple, Term> two_terms;
    std::map<SixTuple, Term> three_terms;
    std::map<EightTuple, Term> four_terms;

public:
    explicit ModelSU2(ModelParams const& p)
    {
        symm = p.symm;
        lat.L = p.L; lat.irreps = p.site_types;
        if ((int)lat.irreps.size() != p.L) lat.irreps.assign(p.L, 0);
        tag_handler.reset(new TagHandler());
        total_charge = pg_charge(symm, Charge(p.nelec, p.spin, 0), p.irrep);
        int max_irrep = *std::max_element(lat.irreps.begin(), lat.irreps.end());
        Charge A(2, 0), B(1, 1), C(1, -1), D(0, 0);
        for (int irr = 0; irr <= max_irrep; ++irr) {
            Index phys;
            phys.insert(std::make_pair(A, (size_t)1));
            phys.insert(std::make_pair(pg_charge(symm, B, irr), (size_t)1));
            phys.insert(std::make_pair(pg_charge(symm, C, irr), (size_t)1));
            phys.insert(std::make_pair(D, (size_t)1));
            phys_indices.push_back(phys);
        }
        construct_operators(max_irrep);
        ident = ops.ident; ident_full = ops.ident_full; fill = ops.fill;
        create_terms(p);
    }

private:
    int ty(int p) const { return lat.type(p); }

    // su2u1/term_maker.h:139-352
    void construct_operators(int max_irrep)
    {
        Charge A(2, 0), B(1, 1), C(1, -1), D(0, 0);
        SpinDescriptor one_half_up(1, 0, 1), one_half_down(1, 1, 0), one_up(2, 0, 2), one_flat(2, 1, 1), one_down(2, 2, 0);
        const double s2 = std::sqrt(2.), s32 = std::sqrt(3. / 2.);
        SiteOperator ident_op, ident_full_op, fill_op;
        ident_op.insert_block(1, A, A); ident_op.insert_block(1, B, B); ident_op.insert_block(1, C, C); ident_op.insert_block(1, D, D);
        ident_full_op.insert_block(1, A, A); ident_full_op.insert_block(1, D, D); ident_full_op.insert_block(1, B, B);
        ident_full_op.insert_block(1, C, C); ident_full_op.insert_block(1, B, C); ident_full_op.insert_block(1, C, B);
        fill_op.insert_block(1, A, A); fill_op.insert_block(1, D, D); fill_op.insert_block(-1, B, B);
        fill_op.insert_block(-1, C, C); fill_op.insert_block(-1, B, C); fill_op.insert_block(-1, C, B);

        SiteOperator create_fill_op; create_fill_op.spin() = one_half_up;
        create_fill_op.insert_block(s2, B, A); create_fill_op.insert_block(s2, C, A);
        create_fill_op.insert_block(1, D, B); create_fill_op.insert_block(1, D, C);
        SiteOperator destroy_op; destroy_op.spin() = one_half_down;
        destroy_op.insert_block(1, A, B); destroy_op.insert_block(1, A, C);
        destroy_op.insert_block(s2, B, D); destroy_op.insert_block(s2, C, D);
        SiteOperator destroy_fill_op; destroy_fill_op.spin() = one_half_up;
        destroy_fill_op.insert_block(1, A, B); destroy_fill_op.insert_block(1, A, C);
        destroy_fill_op.insert_block(-s2, B, D); destroy_fill_op.insert_block(-s2, C, D);
        SiteOperator create_op; create_op.spin() = one_half_down;
        create_op.insert_block(s2, B, A); create_op.insert_block(s2, C, A);
        create_op.insert_block(-1, D, B); create_op.insert_block(-1, D, C);

        SiteOperator create_fill_couple_down_op = create_fill_op; create_fill_couple_down_op.spin() = one_half_down;
        SiteOperator destroy_fill_couple_down_op = destroy_fill_op; destroy_fill_couple_down_op.spin() = one_half_down;
        SiteOperator create_couple_up_op = create_op; create_couple_up_op.spin() = one_half_up;
        SiteOperator destroy_couple_up_op = destroy_op; destroy_couple_up_op.spin() = one_half_up;

        SiteOperator create_fill_count_op; create_fill_count_op.spin() = one_half_up;
        create_fill_count_op.insert_block(s2, B, A); create_fill_count_op.insert_block(s2, C, A);
        SiteOperator destroy_count_op; destroy_count_op.spin() = one_half_down;
        destroy_count_op.insert_block(1, A, B); destroy_count_op.insert_block(1, A, C);
        SiteOperator destroy_fill_count_op; destroy_fill_count_op.spin() = one_half_up;
        destroy_fill_count_op.insert_block(1, A, B); destroy_fill_count_op.insert_block(1, A, C);
        SiteOperator create_count_op; create_count_op.spin() = one_half_down;
        create_count_op.insert_block(s2, B, A); create_count_op.insert_block(s2, C, A);

        SiteOperator count_op; count_op.insert_block(2, A, A); count_op.insert_block(1, B, B); count_op.insert_block(1, C, C);
        SiteOperator docc_op; docc_op.insert_block(1, A, A);
        SiteOperator e2d_op; e2d_op.insert_block(1, D, A);
        SiteOperator d2e_op; d2e_op.insert_block(1, A, D);
        SiteOperator count_fill_op; count_fill_op.insert_block(2, A, A); count_fill_op.insert_block(-1, B, B);
        count_fill_op.insert_block(-1, C, C); count_fill_op.insert_block(-1, B, C); count_fill_op.insert_block(-1, C, B);
        SiteOperator flip_to_S2_op; flip_to_S2_op.spin() = one_up;
        flip_to_S2_op.insert_block(s32, B, B); flip_to_S2_op.insert_block(s32, C, C);
        flip_to_S2_op.insert_block(s32, B, C); flip_to_S2_op.insert_block(s32, C, B);
        SiteOperator flip_to_S0_op = flip_to_S2_op; flip_to_S0_op.spin() = one_down;
        SiteOperator flip_S0_op = flip_to_S2_op; flip_S0_op.spin() = one_flat;

        auto reg = [&](SiteOperator const& op, bool fermionic) {
            tag_vec ret;
            for (auto const& o : site_specific_ops(symm, op, max_irrep)) ret.push_back(tag_handler->checked_register(o, fermionic).first);
            return ret;
        };
        ops.ident = reg(ident_op, false); ops.ident_full = reg(ident_full_op, false); ops.fill = reg(fill_op, false);
        ops.create_fill = reg(create_fill_op, true); ops.create = reg(create_op, true);
        ops.destroy_fill = reg(destroy_fill_op, true); ops.destroy = reg(destroy_op, true);
        ops.create_fill_couple_down = reg(create_fill_couple_down_op, true);
        ops.destroy_fill_couple_down = reg(destroy_fill_couple_down_op, true);
        ops.create_couple_up = reg(create_couple_up_op, true); ops.destroy_couple_up = reg(destroy_couple_up_op, true);
        ops.create_fill_count = reg(create_fill_count_op, true); ops.create_count = reg(create_count_op, true);
        ops.destroy_fill_count = reg(destroy_fill_count_op, true); ops.destroy_count = reg(destroy_count_op, true);
        ops.count = reg(count_op, false); ops.docc = reg(docc_op, false); ops.e2d = reg(e2d_op, false); ops.d2e = reg(d2e_op, false);
        ops.flip_S0 = reg(flip_S0_op, false); ops.flip_to_S2 = reg(flip_to_S2_op, false); ops.flip_to_S0 = reg(flip_to_S0_op, false);
        ops.count_fill = reg(count_fill_op, false);

        auto herm = [&](tag_vec const& a, tag_vec const& b) { for (size_t h = 0; h < a.size(); ++h) tag_handler->hermitian_pair(a[h], b[h]); };
        herm(ops.create_fill, ops.destroy_fill); herm(ops.create, ops.destroy); herm(ops.e2d, ops.d2e);
        herm(ops.create_fill_count, ops.destroy_fill_count); herm(ops.create_count, ops.destroy_count);
        herm(ops.create_fill_couple_down, ops.destroy_fill_couple_down); herm(ops.create_couple_up, ops.destroy_couple_up);

        // construct_operator_collection (term_maker.h:354-424)
        oc.create.couple_up = ops.create_couple_up; oc.create.couple_down = ops.create;
        oc.create.fill_couple_up = ops.create_fill; oc.create.fill_couple_down = ops.create_fill_couple_down;
        oc.destroy.couple_up = ops.destroy_couple_up; oc.destroy.couple_down = ops.destroy;
        oc.destroy.fill_couple_up = ops.destroy_fill; oc.destroy.fill_couple_down = ops.destroy_fill_couple_down;
        oc.create_count.couple_down = ops.create_count; oc.create_count.fill_couple_up = ops.create_fill_count;
        oc.destroy_count.couple_down = ops.destroy_count; oc.destroy_count.fill_couple_up = ops.destroy_fill_count;
        oc.ident.no_couple = ops.ident; oc.ident_full.no_couple = ops.ident_full; oc.fill.no_couple = ops.fill;
        oc.count.no_couple = ops.count; oc.count.fill_no_couple = ops.count_fill;
        oc.e2d.no_couple = ops.e2d; oc.d2e.no_couple = ops.d2e; oc.docc.no_couple = ops.docc;
        oc.flip.no_couple = ops.flip_S0; oc.flip.couple_up = ops.flip_to_S2; oc.flip.couple_down = ops.flip_to_S0;
    }

    // ---- TermMakerSU2 (term_maker.h:426-537)
    Term tm_two_term(bool sign, double scale, int i, int j, tag_vec const& op1, tag_vec const& op2)
    {
        Term t; t.is_fermionic = sign; t.coeff = scale;
        t.push_back(std::make_pair(i, op1[ty(i)])); t.push_back(std::make_pair(j, op2[ty(j)]));
        return t;
    }
    Term tm_positional_two_term(bool sign, double scale, int i, int j, tag_vec const& op1, tag_vec const& op1_fill,
                                tag_vec const& op2, tag_vec const& op2_fill)
    {
        Term t; t.is_fermionic = sign; t.coeff = scale;
        tag_vec const& op1_use = (i < j) ? op1_fill : op2_fill;
        tag_vec const& op2_use = (i < j) ? op2 : op1;
        if (j < i && sign) t.coeff = -t.coeff;
        int start = std::min(i, j), end = std::max(i, j);
        t.push_back(std::make_pair(start, op1_use[ty(start)]));
        t.push_back(std::make_pair(end, op2_use[ty(end)]));
        return t;
    }
    Term tm_three_term(double scale, int pb, int p1, int p2, tag_vec const& boson_op, tag_vec const& boson_op_fill,
                       tag_vec const& op1, tag_vec const& op1_fill, tag_vec const& op2, tag_vec const& op2_fill)
    {
        Term t; t.is_fermionic = true; t.coeff = scale;
        tag_vec const& op1_use = (p1 < p2) ? op1_fill : op2_fill;
        tag_vec const& op2_use = (p1 < p2) ? op2 : op1;
        tag_vec const& boson_use = ((pb > p1 && pb < p2) || (pb > p2 && pb < p1)) ? boson_op_fill : boson_op;
        if (p2 < p1) t.coeff = -t.coeff;
        int start = std::min(p1, p2), end = std::max(p1, p2);
        t.push_back(std::make_pair(pb, boson_use[ty(pb)]));
        t.push_back(std::make_pair(start, op1_use[ty(start)]));
        t.push_back(std::make_pair(end, op2_use[ty(end)]));
        t.canonical_order();
        return t;
    }
    Term tm_four_term(int max_two_S, double scale, int i, int j, int k, int l, Bundle const& op_i, Bundle const& op_k)
    {
        Term t; t.is_fermionic = true; t.coeff = scale;
        if (perm_sgn(i, j, k, l)) t.coeff = -t.coeff;
        std::vector<std::pair<int, Bundle const*>> st = {{i, &op_i}, {j, &op_i}, {k, &op_k}, {l, &op_k}};
        std::stable_sort(st.begin(), st.end(), [](auto const& a, auto const& b) { return a.first < b.first; });
        if (max_two_S == 2) {
            t.push_back(std::make_pair(st[0].first, st[0].second->fill_couple_up[ty(st[0].first)]));
            t.push_back(std::make_pair(st[1].first, st[1].second->couple_up[ty(st[1].first)]));
            t.push_back(std::make_pair(st[2].first, st[2].second->fill_couple_down[ty(st[2].first)]));
            t.push_back(std::make_pair(st[3].first, st[3].second->couple_down[ty(st[3].first)]));
        } else {
            t.push_back(std::make_pair(st[0].first, st[0].second->fill_couple_up[ty(st[0].first)]));
            t.push_back(std::make_pair(st[1].first, st[1].second->couple_down[ty(st[1].first)]));
            t.push_back(std::make_pair(st[2].first, st[2].second->fill_couple_up[ty(st[2].first)]));
            t.push_back(std::make_pair(st[3].first, st[3].second->couple_down[ty(st[3].first)]));
        }
        return t;
    }

    // ---- SpinSumSU2 (term_maker.h:539-790)
    typedef std::vector<Term> term_vec;
    term_vec ss_V_term(double me, int i, int k, int l, int j)
    {
        std::vector<int> ps = {i, k, l, j};
        std::sort(ps.begin(), ps.end());
        size_t n_unique = std::unique(ps.begin(), ps.end()) - ps.begin();
        switch (n_unique) {
            case 4: return ss_four_term(me, i, k, l, j);
            case 3: return ss_three_term(me, i, k, l, j);
            case 2: return ss_two_term(me, i, k, l, j);
            case 1: { Term t; t.coeff = 2. * me; t.push_back(std::make_pair(i, oc.docc.no_couple[ty(i)])); return term_vec(1, t); }
        }
        return term_vec();
    }
    term_vec ss_two_term(double me, int i, int k, int l, int j)
    {
        term_vec ret;
        if (i == j && k == l && j != k)
            ret.push_back(tm_two_term(false, me, i, k, oc.count.no_couple, oc.count.no_couple));
        else if (i == k && j == l && j != k)
            ret.push_back(tm_two_term(false, 2.0 * me, i, j, oc.e2d.no_couple, oc.d2e.no_couple));
        else if (i == l && j == k && i != j) {
            ret.push_back(tm_positional_two_term(false, std::sqrt(3.) * me, i, j, oc.flip.couple_down, oc.flip.couple_up,
                                                 oc.flip.couple_down, oc.flip.couple_up));
            ret.push_back(tm_two_term(false, -0.5 * me, i, j, oc.count.no_couple, oc.count.no_couple));
        } else if ((i == k && k == l) || (k == l && l == j) || (i == l && l == j) || (i == k && k == j)) {
            int s, p;
            if (i == k && k == l) { s = i; p = j; }
            else if (k == l && l == j) { s = j; p = i; }
            else if (i == l && l == j) { s = i; p = k; }
            else { s = i; p = l; }
            if (i == k)
                ret.push_back(tm_positional_two_term(true, std::sqrt(2.) * me, s, p, oc.create_count.couple_down, oc.create_count.fill_couple_up,
                                                     oc.destroy.couple_down, oc.destroy.fill_couple_up));
            else
                ret.push_back(tm_positional_two_term(true, -std::sqrt(2.) * me, s, p, oc.destroy_count.couple_down, oc.destroy_count.fill_couple_up,
                                                     oc.create.couple_down, oc.create.fill_couple_up));
        } else
            throw std::runtime_error("Unexpected index arrangement for V_ijjj term");
        return ret;
    }
    term_vec ss_three_term(double me, int i, int k, int l, int j)
    {
        term_vec ret;
        if (i == j || k == l) {   // three_termA
            int same_idx = 0;
            if (i == j) same_idx = i;
            if (k == l) { same_idx = k; k = i; l = j; }
            ret.push_back(tm_three_term(std::sqrt(2.) * me, same_idx, k, l, oc.count.no_couple, oc.count.fill_no_couple,
                                        oc.create.couple_down, oc.create.fill_couple_up, oc.destroy.couple_down, oc.destroy.fill_couple_up));
            return ret;
        }
        // three_termB
        int same_idx, pos1, pos2;
        if (i == k) {
            same_idx = i; pos1 = std::min(l, j); pos2 = std::max(l, j);
            ret.push_back(tm_three_term(-std::sqrt(2.) * me, same_idx, pos1, pos2, oc.e2d.no_couple, oc.e2d.no_couple,
                                        oc.destroy.couple_down, oc.destroy.fill_couple_up, oc.destroy.couple_down, oc.destroy.fill_couple_up));
        }
        if (j == l) {
            same_idx = j; pos1 = std::min(i, k); pos2 = std::max(i, k);
            ret.push_back(tm_three_term(-std::sqrt(2.) * me, same_idx, pos1, pos2, oc.d2e.no_couple, oc.d2e.no_couple,
                                        oc.create.couple_down, oc.create.fill_couple_up, oc.create.couple_down, oc.create.fill_couple_up));
        }
        if (j == k || i == l) {
            if (j == k) { same_idx = j; pos1 = l; pos2 = i; }
            else { same_idx = i; pos1 = j; pos2 = k; }
            double phase = perm_sgn(i, k, l, j) ? -1. : 1.;
            if (same_idx < std::min(pos1, pos2)) {
                ret.push_back(tm_three_term(phase * std::sqrt(3.) * me, same_idx, pos1, pos2, oc.flip.couple_up, oc.flip.couple_up,
                                            oc.create.couple_down, oc.create.fill_couple_down, oc.destroy.couple_down, oc.destroy.fill_couple_down));
                ret.push_back(tm_three_term(-0.5 * std::sqrt(2.) * me, same_idx, pos1, pos2, oc.count.no_couple, oc.count.no_couple,
                                            oc.create.couple_down, oc.create.fill_couple_up, oc.destroy.couple_down, oc.destroy.fill_couple_up));
            } else if (same_idx > std::max(pos1, pos2)) {
                ret.push_back(tm_three_term(phase * std::sqrt(3.) * me, same_idx, pos1, pos2, oc.flip.couple_down, oc.flip.couple_down,
                                            oc.create.couple_up, oc.create.fill_couple_up, oc.destroy.couple_up, oc.destroy.fill_couple_up));
                ret.push_back(tm_three_term(-0.5 * std::sqrt(2.) * me, same_idx, pos1, pos2, oc.count.no_couple, oc.count.no_couple,
                                            oc.create.couple_down, oc.create.fill_couple_up, oc.destroy.couple_down, oc.destroy.fill_couple_up));
            } else {
                ret.push_back(tm_three_term(phase * std::sqrt(3.) * me, same_idx, pos1, pos2, oc.flip.no_couple, oc.flip.no_couple,
                                            oc.create.couple_down, oc.create.fill_couple_up, oc.destroy.couple_down, oc.destroy.fill_couple_up));
                ret.push_back(tm_three_term(-0.5 * std::sqrt(2.) * me, same_idx, pos1, pos2, oc.count.fill_no_couple, oc.count.fill_no_couple,
                                            oc.create.couple_down, oc.create.fill_couple_up, oc.destroy.couple_down, oc.destroy.fill_couple_up));
            }
        }
        return ret;
    }
    term_vec ss_four_term(double me, int i, int k, int l, int j)
    {
        term_vec ret;
        IndexTuple key = align_idx(i, j, k, l);
        int j_ = key[1], k_ = key[2], l_ = key[3];
        if (k_ > l_ && l_ > j_) {
            ret.push_back(tm_four_term(2, -std::sqrt(3.) * me, i, k, l, j, oc.create, oc.destroy));
            ret.push_back(tm_four_term(1, me, i, k, l, j, oc.create, oc.destroy));
        } else if (k_ > j_ && j_ > l_) {
            double le = perm_sgn(i, k, l, j) ? -me : me;
            ret.push_back(tm_four_term(2, std::sqrt(3.) * le, i, k, l, j, oc.create, oc.destroy));
            ret.push_back(tm_four_term(1, le, i, k, l, j, oc.create, oc.destroy));
        } else if (j_ > k_ && k_ > l_) {
            ret.push_back(tm_four_term(1, 2. * me, i, k, l, j, oc.create, oc.destroy));
        } else
            throw std::runtime_error("unexpected index arrangment in V_ijkl term");
        return ret;
    }

    // ---- ChemHelperSU2 (su2u1/chem_helper.h:58-94)
    void add_2term(Term const& t)
    {
        IndexTuple id = {t.position(0), t.position(1), (int)t.operator_tag(0), (int)t.operator_tag(1)};
        auto it = two_terms.find(id);
        if (it == two_terms.end()) two_terms[id] = t; else it->second.coeff += t.coeff;
    }
    void add_3term(Term const& t)
    {
        SixTuple id = {t.position(0), t.position(1), t.position(2), (int)t.operator_tag(0), (int)t.operator_tag(1), (int)t.operator_tag(2)};
        auto it = three_terms.find(id);
        if (it == three_terms.end()) three_terms[id] = t; else it->second.coeff += t.coeff;
    }
    void add_4term(Term const& t)
    {
        EightTuple id = {t.position(0), t.position(1), t.position(2), t.position(3),
                         (int)t.operator_tag(0), (int)t.operator_tag(1), (int)t.operator_tag(2), (int)t.operator_tag(3)};
        auto it = four_terms.find(id);
        if (it == four_terms.end()) four_terms[id] = t; else it->second.coeff += t.coeff;
    }
    static void append(term_vec& a, term_vec const& b) { a.insert(a.end(), b.begin(), b.end()); }

    // su2u1/model.hpp:48-298
    void create_terms(ModelParams const& p)
    {
        int N = p.nelec;
        ParsedIntegrals pi = parse_integrals(p);
        for (size_t mI = 0; mI < pi.val.size(); ++mI) {
            int i = pi.idx[mI][0], j = pi.idx[mI][1], k = pi.idx[mI][2], l = pi.idx[mI][3];
            double me = pi.val[mI];
            if (i == -1 && j == -1 && k == -1 && l == -1) {
                Term t; t.coeff = me; t.push_back(std::make_pair(0, ops.ident[ty(0)])); terms.push_back(t);
            } else if (i == j && k == -1 && l == -1) {
                Term t; t.coeff = me; t.push_back(std::make_pair(i, ops.count[ty(i)])); terms.push_back(t);
            } else if (k == -1 && l == -1) {
                if (N == 1) {
                    terms.push_back(tm_positional_two_term(true, std::sqrt(2.) * me, j, i, ops.create, ops.create_fill, ops.destroy, ops.destroy_fill));
                    terms.push_back(tm_positional_two_term(true, std::sqrt(2.) * me, i, j, ops.create, ops.create_fill, ops.destroy, ops.destroy_fill));
                } else {
                    term_vec tv;
                    for (int kk = 0; kk < lat.size(); ++kk) {
                        if (kk == j || kk == i) continue;
                        append(tv, ss_three_term(me * (1. / (N - 1)), i, kk, kk, j));
                        append(tv, ss_three_term(me * (1. / (N - 1)), j, kk, kk, i));
                    }
                    for (auto const& t : tv) add_3term(t);
                    tv.clear();
                    append(tv, ss_V_term(me * (1. / (N - 1)), i, i, i, j));
                    append(tv, ss_V_term(me * (1. / (N - 1)), j, i, i, i));
                    append(tv, ss_V_term(me * (1. / (N - 1)), i, j, j, j));
                    append(tv, ss_V_term(me * (1. / (N - 1)), j, j, j, i));
                    for (auto const& t : tv) add_2term(t);
                }
            } else if (i == j && j == k && k == l) {
                Term t; t.coeff = me; t.push_back(std::make_pair(i, ops.docc[ty(i)])); terms.push_back(t);
            } else if ((i == j && j == k && k != l) || (i != j && j == k && k == l)) {
                int s, pp;
                if (i == j) { s = i; pp = l; } else { s = l; pp = i; }
                term_vec tv;
                append(tv, ss_two_term(me, s, s, s, pp));
                append(tv, ss_two_term(me, s, pp, s, s));
                for (auto const& t : tv) add_2term(t);
            } else if (i == j && k == l && j != k) {
                for (auto const& t : ss_two_term(me, i, k, k, i)) add_2term(t);
            } else if (i == k && j == l && i != j) {
                term_vec tv;
                append(tv, ss_two_term(0.5 * me, i, i, j, j));
                append(tv, ss_two_term(0.5 * me, j, j, i, i));
                append(tv, ss_two_term(me, i, j, i, j));
                for (auto const& t : tv) add_2term(t);
            } else if ((i == j && j != k && k != l) || (k == l && i != j && j != k)) {
                term_vec tv;
                if (i == j) { append(tv, ss_three_term(me, i, k, l, i)); append(tv, ss_three_term(me, i, l, k, i)); }
                else { append(tv, ss_three_term(me, i, k, k, j)); append(tv, ss_three_term(me, j, k, k, i)); }
                for (auto const& t : tv) add_3term(t);
            } else if (((i == k && j != l) || j == k || (j == l && i != k)) && (i != j && k != l)) {
                term_vec tv;
                append(tv, ss_three_term(me, i, k, l, j)); append(tv, ss_three_term(me, i, l, k, j));
                append(tv, ss_three_term(me, j, k, l, i)); append(tv, ss_three_term(me, j, l, k, i));
                for (auto const& t : tv) add_3term(t);
            } else if (i != j && j != k && k != l && i != k && j != l) {
                term_vec tv;
                append(tv, ss_four_term(me, i, k, l, j)); append(tv, ss_four_term(me, i, l, k, j));
                append(tv, ss_four_term(me, j, k, l, i)); append(tv, ss_four_term(me, j, l, k, i));
                for (auto const& t : tv) add_4term(t);
            }
        }
        for (auto const& kv : two_terms) terms.push_back(kv.second);
        for (auto const& kv : three_terms) terms.push_back(kv.second);
        for (auto const& kv : four_terms) terms.push_back(kv.second);
    }
};

// =========================================================================================================
// TwoU1 / TwoU1PG
class Model2U1 : public ModelBase
{
    typedef std::vector<tag_type> tag_vec;
    tag_vec create_up, create_down, destroy_up, destroy_down, count_up, count_down, count_up_down, docc, e2d, d2e, d2u, u2d;
    std::map<IndexTuple, double> coefficients;
    std::map<IndexTuple, Term> two_terms;
    std::map<SixTuple, Term> three_terms;

public:
    explicit Model2U1(ModelParams const& p)
    {
        symm = p.symm;
        lat.L = p.L; lat.irreps = p.site_types;
        if ((int)lat.irreps.size() != p.L) lat.irreps.assign(p.L, 0);
        tag_handler.reset(new TagHandler());
        total_charge = pg_charge(symm, Charge(p.nup, p.ndown, 0), p.irrep);
        int max_irrep = *std::max_element(lat.irreps.begin(), lat.irreps.end());
        Charge A(0, 0), B(1, 0), C(0, 1), D(1, 1);
        for (int irr = 0; irr <= max_irrep; ++irr) {
            Index phys;
            phys.insert(std::make_pair(A, (size_t)1));
            phys.insert(std::make_pair(pg_charge(symm, B, irr), (size_t)1));
            phys.insert(std::make_pair(pg_charge(symm, C, irr), (size_t)1));
            phys.insert(std::make_pair(D, (size_t)1));
            phys_indices.push_back(phys);
        }
        SiteOperator create_up_op, create_down_op, destroy_up_op, destroy_down_op, count_up_op, count_down_op, count_up_down_op,
            docc_op, e2d_op, d2e_op, d2u_op, u2d_op, ident_op, fill_op;
        ident_op.insert_block(1, A, A); ident_op.insert_block(1, B, B); ident_op.insert_block(1, C, C); ident_op.insert_block(1, D, D);
        create_up_op.insert_block(1, A, B); create_up_op.insert_block(1, C, D);
        create_down_op.insert_block(1, A, C); create_down_op.insert_block(1, B, D);
        destroy_up_op.insert_block(1, B, A); destroy_up_op.insert_block(1, D, C);
        destroy_down_op.insert_block(1, C, A); destroy_down_op.insert_block(1, D, B);
        count_up_op.insert_block(1, B, B); count_up_op.insert_block(1, D, D);
        count_down_op.insert_block(1, C, C); count_down_op.insert_block(1, D, D);
        count_up_down_op.insert_block(1, B, B); count_up_down_op.insert_block(1, C, C); count_up_down_op.insert_block(2, D, D);
        docc_op.insert_block(1, D, D);
        e2d_op.insert_block(1, A, D); d2e_op.insert_block(1, D, A);
        fill_op.insert_block(1, A, A); fill_op.insert_block(-1, B, B); fill_op.insert_block(-1, C, C); fill_op.insert_block(1, D, D);
        SiteOperator tmp;
        gemm(fill_op, create_down_op, tmp); create_down_op = tmp;
        gemm(destroy_down_op, fill_op, tmp); destroy_down_op = tmp;
        gemm(destroy_down_op, create_up_op, d2u_op);
        gemm(destroy_up_op, create_down_op, u2d_op);

        auto reg = [&](SiteOperator const& op, bool fermionic) {
            tag_vec ret;
            for (auto const& o : site_specific_ops(symm, op, max_irrep)) ret.push_back(tag_handler->checked_register(o, fermionic).first);
            return ret;
        };
        ident = reg(ident_op, false); fill = reg(fill_op, false);
        create_up = reg(create_up_op, true); create_down = reg(create_down_op, true);
        destroy_up = reg(destroy_up_op, true); destroy_down = reg(destroy_down_op, true);
        count_up = reg(count_up_op, false); count_down = reg(count_down_op, false);
        e2d = reg(e2d_op, false); d2e = reg(d2e_op, false); docc = reg(docc_op, false);
        count_up_down = reg(count_up_down_op, false);
        d2u = reg(d2u_op, false); u2d = reg(u2d_op, false);
        ident_full = ident;

        auto& th = *tag_handler;
        auto cutf = th.get_product_tags(create_up, fill), cdtf = th.get_product_tags(create_down, fill);
        auto ftdu = th.get_product_tags(fill, destroy_up), ftdd = th.get_product_tags(fill, destroy_down);
        auto cund = th.get_product_tags(create_up, count_down), dund = th.get_product_tags(destroy_up, count_down);
        auto cdnu = th.get_product_tags(create_down, count_up), ddnu = th.get_product_tags(destroy_down, count_up);
        auto cundtf = th.get_product_tags(cund.first, fill), ftdund = th.get_product_tags(fill, dund.first);
        auto cdnutf = th.get_product_tags(cdnu.first, fill), ftddnu = th.get_product_tags(fill, ddnu.first);
        auto ddcu = th.get_product_tags(destroy_down, create_up), ducd = th.get_product_tags(destroy_up, create_down);
        auto herm = [&](tag_vec const& a, tag_vec const& b) { for (size_t h = 0; h < a.size(); ++h) th.hermitian_pair(a[h], b[h]); };
        herm(create_up, destroy_up); herm(create_down, destroy_down);
        herm(cutf.first, ftdu.first); herm(cdtf.first, ftdd.first); herm(e2d, d2e);
        herm(cund.first, dund.first); herm(cdnu.first, ddnu.first);
        herm(cundtf.first, ftdund.first); herm(cdnutf.first, ftddnu.first); herm(ddcu.first, ducd.first);
        create_terms(p);
    }

private:
    int ty(int p) const { return lat.type(p); }

    // ---- TermMaker (2u1/term_maker.h)
    Term tm_two_term(double scale, int i, int j, tag_vec const& op1, tag_vec const& op2)
    {
        Term t; t.coeff = scale;
        t.push_back(std::make_pair(i, op1[ty(i)])); t.push_back(std::make_pair(j, op2[ty(j)]));
        return t;
    }
    Term tm_positional_two_term(double scale, int i, int j, tag_vec const& op1, tag_vec const& op2)
    {
        Term t; t.is_fermionic = true; t.coeff = scale;
        if (i < j) {
            auto pt = tag_handler->get_product_tag(fill[ty(i)], op1[ty(i)]);
            t.push_back(std::make_pair(i, pt.first)); t.push_back(std::make_pair(j, op2[ty(j)]));
            t.coeff *= pt.second;
        } else {
            auto pt = tag_handler->get_product_tag(fill[ty(j)], op2[ty(j)]);
            t.push_back(std::make_pair(i, op1[ty(i)])); t.push_back(std::make_pair(j, pt.first));
            t.coeff *= -pt.second;
        }
        return t;
    }
    Term tm_positional_two_term3(double scale, int i, int j, tag_vec const& op1, tag_vec const& op2, tag_vec const& op3)
    {
        Term t; t.is_fermionic = true; t.coeff = scale;
        auto pre = tag_handler->get_product_tag(op1[ty(i)], op2[ty(i)]);
        if (i < j) {
            auto pt = tag_handler->get_product_tag(fill[ty(i)], pre.first);
            t.push_back(std::make_pair(i, pt.first)); t.push_back(std::make_pair(j, op3[ty(j)]));
            t.coeff *= pt.second * pre.second;
        } else {
            auto pt = tag_handler->get_product_tag(fill[ty(j)], op3[ty(j)]);
            t.push_back(std::make_pair(i, pre.first)); t.push_back(std::make_pair(j, pt.first));
            t.coeff *= -pt.second * pre.second;
        }
        return t;
    }
    Term tm_three_term(double scale, int pb, int p1, int p2, tag_vec const& opb1, tag_vec const& opb2, tag_vec const& ops1, tag_vec const& ops2)
    {
        Term t; t.is_fermionic = true; t.coeff = scale;
        tag_type boson_op, op1 = ops1[ty(p1)], op2 = ops2[ty(p2)];
        if ((pb > p1 && pb < p2) || (pb > p2 && pb < p1)) {
            auto pt1 = tag_handler->get_product_tag(fill[ty(pb)], opb2[ty(pb)]);
            t.coeff *= pt1.second;
            auto pt2 = tag_handler->get_product_tag(pt1.first, opb1[ty(pb)]);
            t.coeff *= pt2.second;
            boson_op = pt2.first;
        } else {
            auto pt1 = tag_handler->get_product_tag(opb2[ty(pb)], opb1[ty(pb)]);
            boson_op = pt1.first; t.coeff *= pt1.second;
        }
        if (p1 < p2) {
            auto pt = tag_handler->get_product_tag(fill[ty(p1)], ops1[ty(p1)]);
            op1 = pt.first; t.coeff *= pt.second;
        } else {
            auto pt = tag_handler->get_product_tag(fill[ty(p2)], ops2[ty(p2)]);
            op2 = pt.first; t.coeff *= -pt.second;
        }
        std::vector<std::pair<int, tag_type>> st = {{pb, boson_op}, {p1, op1}, {p2, op2}};
        std::stable_sort(st.begin(), st.end(), [](auto const& a, auto const& b) { return a.first < b.first; });
        for (auto const& s : st) t.push_back(s);
        return t;
    }
    Term tm_four_term(double scale, int i, int j, int k, int l, tag_vec const& op_i, tag_vec const& op_j, tag_vec const& op_k, tag_vec const& op_l)
    {
        Term t; t.is_fermionic = true; t.coeff = scale;
        bool odd = perm_sgn(i, j, k, l);
        std::vector<std::pair<int, tag_type>> st = {{i, op_i[ty(i)]}, {j, op_j[ty(j)]}, {k, op_k[ty(k)]}, {l, op_l[ty(l)]}};
        std::stable_sort(st.begin(), st.end(), [](auto const& a, auto const& b) { return a.first < b.first; });
        auto pt = tag_handler->get_product_tag(fill[ty(st[0].first)], st[0].second);
        st[0].second = pt.first; t.coeff *= pt.second;
        pt = tag_handler->get_product_tag(fill[ty(st[2].first)], st[2].second);
        st[2].second = pt.first; t.coeff *= pt.second;
        if (odd) t.coeff = -t.coeff;
        for (auto const& s : st) t.push_back(s);
        return t;
    }

    // ---- ChemHelper (2u1/chem_helper.h:57-149)
    void add_term2(double scale, int p1, int p2, tag_vec const& op_1, tag_vec const& op_2)
    {
        Term t = tm_two_term(scale, p1, p2, op_1, op_2);
        IndexTuple id = {p1, p2, (int)op_1[ty(p1)], (int)op_2[ty(p2)]};
        auto it = two_terms.find(id);
        if (it == two_terms.end()) two_terms[id] = t; else it->second.coeff += t.coeff;
    }
    void add_term2x2(double scale, int p1, int p2, tag_vec const& o1, tag_vec const& o2, tag_vec const& o3, tag_vec const& o4)
    {
        auto pt1 = tag_handler->get_product_tag(o1[ty(p1)], o2[ty(p1)]);
        auto pt2 = tag_handler->get_product_tag(o3[ty(p2)], o4[ty(p2)]);
        Term t; t.coeff = scale * pt1.second * pt2.second;
        t.push_back(std::make_pair(p1, pt1.first)); t.push_back(std::make_pair(p2, pt2.first));
        IndexTuple id = {p1, p2, (int)pt1.first, (int)pt2.first};
        auto it = two_terms.find(id);
        if (it == two_terms.end()) two_terms[id] = t; else it->second.coeff += t.coeff;
    }
    void add_term3(double scale, int s, int p1, int p2, tag_vec const& op_i, tag_vec const& op_k, tag_vec const& op_l, tag_vec const& op_j)
    {
        Term t = tm_three_term(scale, s, p1, p2, op_i, op_k, op_l, op_j);
        SixTuple id = {t.position(0), t.position(1), t.position(2), (int)t.operator_tag(0), (int)t.operator_tag(1), (int)t.operator_tag(2)};
        auto it = three_terms.find(id);
        if (it == three_terms.end()) three_terms[id] = t; else it->second.coeff += t.coeff;
    }
    double coef(IndexTuple const& t) { return coefficients[align_idx(t[0], t[1], t[2], t[3])]; }
    void add_term4(int i, int k, int l, int j, tag_vec const& op_i, tag_vec const& op_k, tag_vec const& op_l, tag_vec const& op_j)
    {
        if (op_i[0] == op_k[0] && op_j[0] == op_l[0]) {
            IndexTuple self = {i, j, k, l}, twin = {i, l, k, j};
            if (i < j) twin = IndexTuple{k, j, i, l};
            if (self > twin) {
                Term t = tm_four_term(coef({i, j, k, l}), i, k, l, j, op_i, op_k, op_l, op_j);
                Term tt = tm_four_term(coef(twin), twin[0], twin[2], twin[3], twin[1], op_i, op_k, op_l, op_j);
                t.coeff += tt.coeff;
                terms.push_back(t);
            }
        } else
            terms.push_back(tm_four_term(coef({i, j, k, l}), i, k, l, j, op_i, op_k, op_l, op_j));
    }

    // 2u1/model.hpp:173-454
    void create_terms(ModelParams const& p)
    {
        ParsedIntegrals pi = parse_integrals(p);
        for (size_t mI = 0; mI < pi.val.size(); ++mI) coefficients[pi.idx[mI]] = pi.val[mI];
        for (size_t mI = 0; mI < pi.val.size(); ++mI) {
            int i = pi.idx[mI][0], j = pi.idx[mI][1], k = pi.idx[mI][2], l = pi.idx[mI][3];
            double me = pi.val[mI];
            if (i == -1 && j == -1 && k == -1 && l == -1) {
                Term t; t.coeff = me; t.push_back(std::make_pair(0, ident[ty(0)])); terms.push_back(t);
            } else if (i == j && k == -1 && l == -1) {
                { Term t; t.coeff = me; t.push_back(std::make_pair(i, count_up[ty(i)])); terms.push_back(t); }
                { Term t; t.coeff = me; t.push_back(std::make_pair(i, count_down[ty(i)])); terms.push_back(t); }
            } else if (k == -1 && l == -1) {
                terms.push_back(tm_positional_two_term(me, i, j, create_up, destroy_up));
                terms.push_back(tm_positional_two_term(me, i, j, create_down, destroy_down));
                terms.push_back(tm_positional_two_term(me, j, i, create_up, destroy_up));
                terms.push_back(tm_positional_two_term(me, j, i, create_down, destroy_down));
            } else if (i == j && j == k && k == l) {
                Term t; t.coeff = me; t.push_back(std::make_pair(i, docc[ty(0)])); terms.push_back(t);
            } else if ((i == j && j == k && k != l) || (i != j && j == k && k == l)) {
                int same_idx, pos1;
                if (i == j) { same_idx = i; pos1 = l; } else { same_idx = l; pos1 = i; }
                terms.push_back(tm_positional_two_term3(me, same_idx, pos1, count_down, create_up, destroy_up));
                terms.push_back(tm_positional_two_term3(-me, same_idx, pos1, destroy_up, count_down, create_up));
                terms.push_back(tm_positional_two_term3(me, same_idx, pos1, count_up, create_down, destroy_down));
                terms.push_back(tm_positional_two_term3(-me, same_idx, pos1, destroy_down, count_up, create_down));
            } else if (i == j && k == l && j != k) {
                add_term2(me, i, k, count_up_down, count_up_down);
            } else if (i == k && j == l && i != j) {
                add_term2(me, i, j, e2d, d2e);
                add_term2(me, i, j, d2e, e2d);
                add_term2(-me, i, j, count_up, count_up);
                add_term2(-me, i, j, count_down, count_down);
                add_term2x2(-me, i, j, destroy_down, create_up, destroy_up, create_down);
                add_term2x2(-me, i, j, destroy_up, create_down, destroy_down, create_up);
            } else if ((i == j && j != k && k != l) || (k == l && i != j && j != k)) {
                int same_idx = 0;
                if (i == j) same_idx = i;
                if (k == l) { same_idx = k; k = i; l = j; }
                add_term3(me, same_idx, k, l, create_up, destroy_up, create_up, destroy_up);
                add_term3(me, same_idx, k, l, create_up, destroy_up, create_down, destroy_down);
                add_term3(me, same_idx, k, l, create_down, destroy_down, create_up, destroy_up);
                add_term3(me, same_idx, k, l, create_down, destroy_down, create_down, destroy_down);
                add_term3(me, same_idx, l, k, create_up, destroy_up, create_up, destroy_up);
                add_term3(me, same_idx, l, k, create_up, destroy_up, create_down, destroy_down);
                add_term3(me, same_idx, l, k, create_down, destroy_down, create_up, destroy_up);
                add_term3(me, same_idx, l, k, create_down, destroy_down, create_down, destroy_down);
            } else if (((i == k && j != l) || j == k || (j == l && i != k)) && (i != j && k != l)) {
                int same_idx = 0, pos1 = 0, pos2 = 0;
                if (i == k) { same_idx = i; pos1 = l; pos2 = j; }
                if (j == k) { same_idx = j; pos1 = l; pos2 = i; }
                if (j == l) { same_idx = j; pos1 = k; pos2 = i; }
                add_term3(me, same_idx, pos1, pos2, create_up, create_down, destroy_down, destroy_up);
                add_term3(me, same_idx, pos1, pos2, create_down, create_up, destroy_up, destroy_down);
                add_term3(me, same_idx, pos1, pos2, destroy_down, destroy_up, create_up, create_down);
                add_term3(me, same_idx, pos1, pos2, destroy_up, destroy_down, create_down, create_up);
                add_term3(-me, same_idx, pos1, pos2, create_up, destroy_up, create_up, destroy_up);
                add_term3(-me, same_idx, pos1, pos2, create_up, destroy_down, create_down, destroy_up);
                add_term3(-me, same_idx, pos1, pos2, create_down, destroy_up, create_up, destroy_down);
                add_term3(-me, same_idx, pos1, pos2, create_down, destroy_down, create_down, destroy_down);
                add_term3(-me, same_idx, pos2, pos1, create_up, destroy_up, create_up, destroy_up);
                add_term3(-me, same_idx, pos2, pos1, create_up, destroy_down, create_down, destroy_up);
                add_term3(-me, same_idx, pos2, pos1, create_down, destroy_up, create_up, destroy_down);
                add_term3(-me, same_idx, pos2, pos1, create_down, destroy_down, create_down, destroy_down);
            } else if (i != j && j != k && k != l && i != k && j != l) {
                int perms[4][4] = {{i, k, l, j}, {i, l, k, j}, {j, k, l, i}, {j, l, k, i}};
                for (auto& q : perms) {
                    add_term4(q[0], q[1], q[2], q[3], create_up, create_up, destroy_up, destroy_up);
                    add_term4(q[0], q[1], q[2], q[3], create_up, create_down, destroy_down, destroy_up);
                    add_term4(q[0], q[1], q[2], q[3], create_down, create_up, destroy_up, destroy_down);
                    add_term4(q[0], q[1], q[2], q[3], create_down, create_down, destroy_down, destroy_down);
                }
            }
        }
        for (auto const& kv : two_terms) terms.push_back(kv.second);
        for (auto const& kv : three_terms) terms.push_back(kv.second);
    }
};

inline std::shared_ptr<ModelBase> make_model(ModelParams const& p)
{
    if (is_su2(p.symm)) return std::shared_ptr<ModelBase>(new ModelSU2(p));
    return std::shared_ptr<ModelBase>(new Model2U1(p));
}

} // namespace qcm
