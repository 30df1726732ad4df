// Host-side symmetry algebra shared by the schedule builder, the Engine mirror and the oracle.
//
// One runtime-tagged charge type covers the four groups QCMaquis builds for electronic DMRG:
//   TwoU1 / TwoU1PG : (N_up, N_down [, irrep])     reference: dmrg/block_matrix/symmetry/nu1_tpl.h, nu1pg.h:26-84
//   SU2U1 / SU2U1PG : (N, 2S [, irrep])            reference: dmrg/block_matrix/symmetry/su2u1.h:18-98
// All four order charges lexicographically over every component (nu1pg.h:133-166,195-205), fuse by
// component-wise addition with the point-group irrep combined through the D2h product table, which is
// bitwise XOR (nu1pg.h:207-222,311-324), and negate only the U(1) components (irrep self-adjoint).
// Groups without point group simply keep irrep == 0, so a single 3-int representation is exact.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <string>
#include <stdexcept>
#include <tuple>

namespace qcm {

enum SymmKind : int { TWOU1 = 0, TWOU1PG = 1, SU2U1 = 2, SU2U1PG = 3 };

inline bool is_su2(SymmKind k) { return k == SU2U1 || k == SU2U1PG; }
inline bool has_pg(SymmKind k) { return k == TWOU1PG || k == SU2U1PG; }

// symmetry=... strings accepted by the reference (dmrg/block_matrix/symmetry/symmetry_traits.h:132-157)
inline SymmKind symm_from_string(std::string const& s)
{
    if (s == "2u1") return TWOU1;
    if (s == "2u1pg") return TWOU1PG;
    if (s == "su2u1") return SU2U1;
    if (s == "su2u1pg") return SU2U1PG;
    throw std::runtime_error("unknown symmetry " + s);
}

struct Charge
{
    int32_t c[3];
    Charge() : c{0, 0, 0} {}
    Charge(int a, int b, int irr = 0) : c{a, b, irr} {}
    int32_t& operator[](int i) { return c[i]; }
    int32_t const& operator[](int i) const { return c[i]; }
};

inline bool operator==(Charge const& a, Charge const& b) { return a.c[0] == b.c[0] && a.c[1] == b.c[1] && a.c[2] == b.c[2]; }
inline bool operator!=(Charge const& a, Charge const& b) { return !(a == b); }
inline bool operator<(Charge const& a, Charge const& b)
{
    return std::tie(a.c[0], a.c[1], a.c[2]) < std::tie(b.c[0], b.c[1], b.c[2]);
}
inline bool operator>(Charge const& a, Charge const& b) { return b < a; }
inline Charge operator-(Charge const& a) { return Charge(-a.c[0], -a.c[1], a.c[2]); }
inline Charge fuse(Charge const& a, Charge const& b) { return Charge(a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] ^ b.c[2]); }

inline int spin(Charge const& a) { return a.c[1]; }   // SU2 groups only (su2u1.h:31-32)
inline int particle_number(SymmKind k, Charge const& a) { return is_su2(k) ? a.c[0] : a.c[0] + a.c[1]; }

// PGCharge / PGDecorator (dmrg/models/chem/pg_util.h): set the irrep only for point-group aware groups
inline Charge pg_charge(SymmKind k, Charge a, int irr) { if (has_pg(k)) a.c[2] = irr; return a; }

struct ChargeHash
{
    size_t operator()(Charge const& a) const
    {
        uint64_t h = (uint32_t)a.c[0];
        h = h * 0x9E3779B97F4A7C15ull + (uint32_t)a.c[1];
        h = h * 0x9E3779B97F4A7C15ull + (uint32_t)a.c[2];
        return (size_t)(h ^ (h >> 29));
    }
};
struct ChargePairHash
{
    size_t operator()(std::pair<Charge, Charge> const& p) const
    {
        return ChargeHash()(p.first) * 1000003u ^ ChargeHash()(p.second);
    }
};

} // namespace qcm
