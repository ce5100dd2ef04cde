// Host-side shell-pair builder for the B200 ERI engine.
//
// Replaces the per-(a,b) / per-(c,d) work of the reference's quartet loop
// (src/integrals/int2e.f90:192-263: p, P, PA, PB, EIJ, getcoef, getDk) by tables built once:
// contracted shells are recovered from the reference's "set" arrays, every shell pair gets its
// primitive-pair records (exponent sum, centre, prefactor, Hermite coefficients folded with
// normalisation and contraction coefficients), pairs are classed by the number of SP sets
// (0: S.S, 1: S.SP, 2: SP.SP) and sorted by their largest prefactor so that the reference's
// EIJ*EGH >= 1e-14 screen becomes a prefix of the list (a Schwarz-like bound that is *exactly*
// the reference's inclusion rule).
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
    int first_fn;     // smallest orbital id (ownership of packed rows)
};

struct PairList {
    int type = 0;
    int n = 0;     // number of shell pairs kept
    int npad = 0;  // n rounded up to 32 (SoA leading dimension)
    // Order: groups of decreasing largest-prefactor (emax equal to 1e-6 relative, i.e. pairs of
    // one symmetry-equivalent kind), and inside a group the Morton order of the pair centre, so
    // that the 32 pairs a warp holds are of one kind (same primitive survival pattern) and
    // spatially close (same Boys regime against a given row).
    std::vector<double> emax;
    std::vector<int32_t> bucket;  // [n] non-decreasing bucket id
    std::vector<int32_t> nprim;
    std::vector<int32_t> pidx;    // [n][nf] packed pair index P(i,j) of each function pair; -1: not stored
                                  // (absent function, or the (j,i) duplicate of a diagonal shell pair)
    std::vector<int32_t> shA, shB;
    std::vector<int32_t> owner_fn;  // min(first_fn(A), first_fn(B)) -> shard ownership key
    // Schwarz factor of the pair: an upper bound of sqrt((ij|ij)) over its function pairs (empty: no Schwarz skip).
    // Filled by the plan from a device pass over the unscreened diagonal quartets (eri_diag_kernel).
    std::vector<double> qmax;
    // primitive records (prims sorted by E descending inside each pair, unused slots zero)
    std::vector<double> aos;  // [n][kMaxPrim][nfield]   (uniform / TMA side)
    std::vector<double> soa;  // [kMaxPrim][nfield][npad] (per-lane side)
};

struct Basis {
    int nnuc = 0, nset = 0, norb = 0;
    std::vector<Shell> shells;
    PairList lists[3];
};

// Returns 0 or a negative MYQC_ERR_* code; err receives a message.
int build_shells(int nnuc, int nset, int setl, const int32_t* setinfo, int ops,
                 const int32_t* basinfo, std::vector<Shell>& shells, std::string& err);

// Group id of a largest-prefactor value (1e-6 relative grid, emax = 1 -> group 0).
int emax_bucket(double emax);

// For every row u of `U`: the number of leading pairs of `T` that must be visited so that every
// pair v with emax_u*emax_v >= 1e-14 is included (pairs inside the prefix that fail the product
// test simply find no surviving primitive quartet).
// With Schwarz factors on both lists and tau > 0 the prefix also ends where qmax_u*qmax_v < tau for every later
// pair: |(ij|kl)| <= sqrt((ij|ij)(kl|kl)) < tau for all their integrals (SURVEY.md 7 "Parity vs. screening").
std::vector<int32_t> row_prefix(const PairList& U, const PairList& T, double tau = 0.0);

// Build the three pair lists restricted to shells for which keep_shell[s] != 0 on BOTH sides
// (keep_shell == nullptr keeps all).
int build_pairs(int nnuc, const double* xyz, const double* set, const int32_t* setinfo,
                int setl, int ops, const double* bas, const int32_t* basinfo,
                const std::vector<Shell>& shells, PairList lists[3], std::string& err);

// Select a sub-list (keeping order) of the pairs for which pred[k] != 0.
PairList sublist(const PairList& src, const std::vector<char>& pred);

}  // namespace myqc
