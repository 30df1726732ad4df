// Wigner 6j / 9j symbols and the SU(2) coupling coefficients folded into contraction tasks.
//
// The reference calls GSL (gsl_sf_coupling_6j / _9j, declared extern "C" at
// dmrg/block_matrix/symmetry/gsl_coupling.h:18-23, version unpinned, GSL itself is not vendored).
// GSL evaluates the Racah single-sum formula for 6j and the sum over products of three 6j for 9j;
// that published algorithm is restated here. Arguments are 2*j integers, as in GSL.
// Pinned by the reference's own table dmrg/tests/test_wigner.cpp:21-46 (see tests/test_wigner.py).
//   set_coupling / mod_coupling / triangle   gsl_coupling.h:25-28,166-204
//   conjugate_correction                     mp_tensors/contractions/non-abelian/gemm.hpp:17-46
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace qcm { namespace su2 {

inline bool triangle(int a, int b, int c)
{
    return ((a + b + c) % 2 == 0) && std::abs(a - b) <= c && c <= a + b;
}

namespace detail {
inline long double lfact(int n) { return lgammal((long double)n + 1.0L); }
// log of triangle coefficient Delta(a,b,c) with 2j arguments
inline long double ldelta(int a, int b, int c)
{
    return 0.5L * (lfact((a + b - c) / 2) + lfact((a - b + c) / 2) + lfact((-a + b + c) / 2) - lfact((a + b + c) / 2 + 1));
}
} // namespace detail

// {ja jb jc; jd je jf}, all arguments doubled
inline double wigner6j_uncached(int ja, int jb, int jc, int jd, int je, int jf)
{
    if (!triangle(ja, jb, jc) || !triangle(ja, je, jf) || !triangle(jd, jb, jf) || !triangle(jd, je, jc)) return 0.0;
    using detail::lfact; using detail::ldelta;
    long double pre = ldelta(ja, jb, jc) + ldelta(ja, je, jf) + ldelta(jd, jb, jf) + ldelta(jd, je, jc);
    int a1 = (ja + jb + jc) / 2, a2 = (ja + je + jf) / 2, a3 = (jd + jb + jf) / 2, a4 = (jd + je + jc) / 2;
    int b1 = (ja + jb + jd + je) / 2, b2 = (jb + jc + je + jf) / 2, b3 = (ja + jc + jd + jf) / 2;
    int tmin = std::max(std::max(a1, a2), std::max(a3, a4));
    int tmax = std::min(b1, std::min(b2, b3));
    long double sum = 0.0L;
    for (int t = tmin; t <= tmax; ++t) {
        long double term = lfact(t + 1) - lfact(t - a1) - lfact(t - a2) - lfact(t - a3) - lfact(t - a4)
                         - lfact(b1 - t) - lfact(b2 - t) - lfact(b3 - t);
        long double v = expl(term + pre);
        sum += (t % 2) ? -v : v;
    }
    return (double)sum;
}

// The two-site operator fusion and the two-site recoupling evaluate the same few hundred 6j symbols millions of times:
// per-thread table (no locking), as the reference keeps its Wigner symbols in a table (gsl_coupling.h:113-144)
inline double wigner6j(int ja, int jb, int jc, int jd, int je, int jf)
{
    const int args[6] = {ja, jb, jc, jd, je, jf};
    uint64_t key = 0;
    for (int q = 0; q < 6; ++q) { if (args[q] < 0 || args[q] > 255) return wigner6j_uncached(ja, jb, jc, jd, je, jf); key = (key << 8) | (uint64_t)args[q]; }
    static thread_local std::unordered_map<uint64_t, double> cache;
    auto it = cache.find(key);
    if (it == cache.end()) it = cache.emplace(key, wigner6j_uncached(ja, jb, jc, jd, je, jf)).first;
    return it->second;
}

// {a b c; d e f; g h i}, all arguments doubled
inline double wigner9j(int a, int b, int c, int d, int e, int f, int g, int h, int i)
{
    if (!triangle(a, b, c) || !triangle(d, e, f) || !triangle(g, h, i) ||
        !triangle(a, d, g) || !triangle(b, e, h) || !triangle(c, f, i)) return 0.0;
    int kmin = std::max(std::abs(a - i), std::max(std::abs(h - d), std::abs(b - f)));
    int kmax = std::min(a + i, std::min(h + d, b + f));
    long double sum = 0.0L;
    for (int k = kmin; k <= kmax; k += 2) {
        long double t = (long double)(k + 1) * wigner6j(a, b, c, f, i, k) * wigner6j(d, e, f, b, k, h) * wigner6j(g, h, i, k, a, d);
        sum += (k % 2) ? -t : t;
    }
    return (double)sum;
}

// gsl_coupling.h:166-175
inline double mod_coupling(int a, int b, int c, int d, int e, int f, int g, int h, int i)
{
    return std::sqrt((g + 1.) * (h + 1.) * (c + 1.) * (f + 1.)) * wigner9j(a, b, c, d, e, f, g, h, i);
}

// gsl_coupling.h:177-204: the four couplings selected per W entry by (row_spin==2, col_spin==2)
inline void set_coupling_uncached(int a, int b, int c, int d, int e, int f, int g, int h, int i, double init, double couplings[4])
{
    double prefactor = std::sqrt((i + 1.) * (a + 1.) / ((g + 1.) * (c + 1.))) * init;
    if (triangle(a, b, c)) {
        couplings[0] = prefactor * mod_coupling(a, b, c, d, e, f, g, h, i);
        couplings[2] = prefactor * mod_coupling(a, b, c, d, e, f, g, 2, i);
    } else { couplings[0] = 0.0; couplings[2] = 0.0; }
    if (triangle(a, 2, c)) {
        couplings[1] = prefactor * mod_coupling(a, 2, c, d, e, f, g, h, i);
        couplings[3] = prefactor * mod_coupling(a, 2, c, d, e, f, g, 2, i);
    } else { couplings[1] = 0.0; couplings[3] = 0.0; }
}

// The reference never evaluates a 9j twice: WignerWrapper keeps every value in a hash table filled at start-up
// (gsl_coupling.h:113-144, su2_wrapper.cpp:14-36).  Same idea here, per thread (no locking), keyed by the nine
// doubled spins; the table holds the couplings for init == 1.
inline void set_coupling(int a, int b, int c, int d, int e, int f, int g, int h, int i, double init, double couplings[4])
{
    struct Entry { double v[4]; };
    static thread_local std::unordered_map<uint64_t, Entry> cache;
    const int args[9] = {a, b, c, d, e, f, g, h, i};
    uint64_t key = 0;
    bool cacheable = true;
    for (int q = 0; q < 9; ++q) { if (args[q] < 0 || args[q] > 126) cacheable = false; key = (key << 7) | (uint64_t)(args[q] & 127); }
    if (!cacheable) { set_coupling_uncached(a, b, c, d, e, f, g, h, i, init, couplings); return; }
    auto it = cache.find(key);
    if (it == cache.end()) {
        Entry en;
        set_coupling_uncached(a, b, c, d, e, f, g, h, i, 1.0, en.v);
        it = cache.emplace(key, en).first;
    }
    for (int q = 0; q < 4; ++q) couplings[q] = it->second.v[q] * init;
}

// The planner asks for the couplings of one (T block, W block) pair millions of times per plan, for a few hundred distinct
// spin combinations: a small direct-mapped per-thread cache (it stays in L1/L2) in front of set_coupling's hash map.
// Values are those of set_coupling with init == 1; the caller multiplies by the term's scale exactly as set_coupling does.
inline const double* coupling_table(int j, int jp, int i, int ip, int a, int k, int ap)
{
    if ((unsigned)j > 255u || (unsigned)jp > 255u || (unsigned)i > 255u || (unsigned)ip > 255u || (unsigned)a > 15u || (unsigned)k > 15u || (unsigned)ap > 15u) return nullptr;
    const uint64_t key = 1 + (((((((uint64_t)j << 8 | (uint64_t)jp) << 8 | (uint64_t)i) << 8 | (uint64_t)ip) << 4 | (uint64_t)a) << 4 | (uint64_t)k) << 4 | (uint64_t)ap);
    struct Entry { uint64_t key; double v[4]; };
    constexpr size_t N = 2048;
    static thread_local Entry cache[N];
    Entry& e = cache[(key * 0x9E3779B97F4A7C15ull) >> 53];
    if (e.key != key) { set_coupling(j, std::abs(j - jp), jp, a, k, ap, i, std::abs(i - ip), ip, 1.0, e.v); e.key = key; }
    return e.v;
}

// non-abelian/gemm.hpp:17-46; lspin/rspin = SU2 spin components of the block's left/right charge
inline double conjugate_correction(int lspin, int rspin, int tensor_spin)
{
    int S = std::min(rspin, lspin);
    int spin_diff = rspin - lspin;
    if (tensor_spin == 0) return 1.;
    if (tensor_spin == 1) {
        if (spin_diff > 0) return -std::sqrt((S + 1.) / (S + 2.));
        if (spin_diff < 0) return std::sqrt((S + 2.) / (S + 1.));
        return 0.;
    }
    if (tensor_spin == 2) {
        if (spin_diff > 0) return -std::sqrt((S + 1.) / (S + 3.));
        if (spin_diff < 0) return -std::sqrt((S + 3.) / (S + 1.));
        return 1.;
    }
    throw std::runtime_error("hermitian conjugate for reduced tensor operators only implemented up to rank 1");
}

}} // namespace qcm::su2
