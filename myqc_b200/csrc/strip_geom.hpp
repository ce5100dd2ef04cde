// Index arithmetic of the packed 8-fold-unique array, shared by the device kernels (eri_kernels.cu) and the
// host-side coverage check of the planner (eri_api.cu: myqc_eri_plan_selfcheck), so that both use the same
// formulas.
//
// Layout (include/myqc_eri.h; the reference's canonical set, int2e.f90:686,692,695):
//   pair index  P(i,j) = i n - i(i-1)/2 + (j-i), i <= j;  element (P,P'), P <= P', at  P np - P(P-1)/2 + (P'-P).
// Row P = (i,j) therefore holds, in memory order, the columns (k,l) with k = i, l >= j and then k > i, l >= k:
// for a fixed first index k the columns are one contiguous run over l.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define MYQC_GEOM_HD __host__ __device__ __forceinline__
#else
#define MYQC_GEOM_HD inline
#endif

namespace myqc {

MYQC_GEOM_HD int64_t pair_index64(int64_t i, int64_t j, int64_t n) { return i * n - ((i * (i - 1)) >> 1) + (j - i); }
// packed index of the diagonal element (P,P) = start of row P
MYQC_GEOM_HD int64_t row_start64(int64_t P, int64_t np) { return P * np - ((P * (P - 1)) >> 1); }
// P(k,l) = col_base64(k,n) + l
MYQC_GEOM_HD int64_t col_base64(int64_t k, int64_t n) { return k * n - ((k * (k - 1)) >> 1) - k; }

// Columns of row (i,j) with first index k and second index l in [llo, lhi): first such l and how many.
MYQC_GEOM_HD int row_piece(int i, int j, int k, int llo, int lhi, int* l0) {
    if (k < i) { *l0 = 0; return 0; }
    const int lmin = (k == i) ? j : k;
    const int a = llo > lmin ? llo : lmin;
    *l0 = a;
    return lhi > a ? lhi - a : 0;
}

// orbitals of the row function pair f of an owner shell pair (A,B) of type UT (0: S.S, 1: S.SP, 2: SP.SP);
// fa/fb = orbital ids of the shells' slots (s,px,py,pz), -1 where absent.  slice >= 0: rows of one first
// function of an SP.SP pair (f = second slot).  Returns false when the row does not exist.
MYQC_GEOM_HD bool owner_row(int UT, int slice, bool a_is_sp, bool diagonal, const int* fa, const int* fb, int f, int* i, int* j) {
    int mu = 0, nu = 0;
    if (UT == 1) { mu = a_is_sp ? f : 0; nu = a_is_sp ? 0 : f; }
    else if (UT == 2) { mu = slice >= 0 ? slice : (f >> 2); nu = slice >= 0 ? f : (f & 3); }
    *i = fa[mu]; *j = fb[nu];
    if (*i < 0 || *j < 0) return false;
    if (diagonal && *i > *j) return false;  // the (j,i) duplicate inside a diagonal shell pair
    return true;
}

}  // namespace myqc
