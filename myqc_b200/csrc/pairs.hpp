// Host-side shell-pair builder for the B200 ERI engine.
//
// Replaces the per-(a,b) / per-(c,d) work of the reference's quartet loop
// (src/integrals/int2e.f90:192-263: p, P, PA, PB, EIJ, getcoef, getDk) by tables built once:
// contracted shells are recovered from the reference's "set" arrays and ordered by their first
// orbital; every shell pair (A <= B) with at least one primitive pair of prefactor >= 1e-14 gets
// its primitive-pair records (exponent sum, centre, prefactor, Hermite coefficients folded with
// normalisation and contraction coefficients).  Pairs are classed by the number of SP sets
// (0: S.S, 1: S.SP, 2: SP.SP) and stored in (A,B) row-major order, so the partners (C,D) of one
// first shell C are consecutive records.  The largest prefactor of a pair, emax, bounds the
// reference's EIJ*EGH >= 1e-14 rule (int2e.f90:257) exactly: fl(x*y) is monotone, so a quartet
// of pairs with fl(emax_u*emax_v) < 1e-14 has no surviving primitive quartet.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace myqc {

constexpr int kMaxPrim = 9;  // primitive pairs per shell pair (STO-nG with n<=3)

// Pair types
enum PairType { PT_SS = 0, PT_SSP = 1, PT_SPSP = 2 };

// number of function pairs, Hermite terms and padded fields per primitive record
constexpr int pt_nf(int t) { return t == 0 ? 1 : (t == 1 ? 4 : 16); }
constexpr int pt_nterm(int t) { return t == 0 ? 1 : (t == 1 ? 7 : 46); }
// fields: p, Px, Py, Pz, E, 1/sqrt(p), coef[nterm], padded to a multiple of 2 doubles (16 B, the
// TMA bulk-copy granule)
constexpr int kCoefField = 6;
constexpr int pt_nfield(int t) { return ((kCoefField + pt_nterm(t)) + 1) / 2 * 2; }

struct Shell {
    int centre;
    int type;         // 0 = S (one s function), 1 = SP (s,px,py,pz; s may be absent)
    int fn[4];        // orbital ids of (s,px,py,pz); -1 if absent (S shells use fn[0] only)
    std::vector<int> sets;  // primitive sets (indices into set[]) in reference order
    int first_fn;     // smallest orbital id
    int end_fn;       // one past the largest orbital id
};

struct PairList {
    int type = 0;
    int n = 0;     // number of shell pairs kept (emax >= 1e-14)
    std::vector<double> emax;     // largest primitive prefactor of the pair
    std::vector<int32_t> nprim;   // primitive pairs with E >= 1e-14, sorted by E descending
    std::vector<int32_t> shA, shB;  // shell indices (A <= B) in the sorted shell order
    std::vector<double> aos;  // [n][kMaxPrim][nfield]: one contiguous block per pair (TMA source on the owner side, per-lane reads on the partner side)
};

// Returns 0 or a negative MYQC_ERR_* code; err receives a message.  Shells come out ordered by
// their first orbital; each shell's orbitals must form one contiguous range (what buildBasis
// produces, basis.f90:101-205), and the ranges must not interleave.
int build_shells(int nnuc, int nset, int setl, const int32_t* setinfo, int ops,
                 const int32_t* basinfo, std::vector<Shell>& shells, std::string& err);

// Build the three pair lists ((A,B) row-major inside each list).
int build_pairs(int nnuc, const double* xyz, const double* set, const int32_t* setinfo,
                int setl, int ops, const double* bas, const int32_t* basinfo,
                const std::vector<Shell>& shells, PairList lists[3], std::string& err);

}  // namespace myqc
