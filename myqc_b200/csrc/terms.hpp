// Static Hermite-term tables shared by the host pair builder and the device kernels.
//
// A shell pair of type t (0: S.S, 1: S.SP, 2: SP.SP) has pt_nf(t) function pairs and
// pt_nterm(t) non-vanishing Hermite terms; term k belongs to function pair term_fn(t,k) and
// Hermite index term_h(t,k).  This is the compile-time form of the lists getDk builds at run
// time (src/integrals/auxilary.f90:576-631): for l <= 1 the set of (N,L,M) with a structurally
// non-zero coefficient is known in advance (1 / 7 / 46 terms).
//
// Hermite indices (N,L,M) are enumerated by total degree, then N descending, then L descending:
//   0:(000) 1:(100) 2:(010) 3:(001) 4:(200) 5:(110) 6:(101) 7:(020) 8:(011) 9:(002) ...
#pragma once

#if defined(__CUDACC__)
#define MYQC_HD __host__ __device__
#else
#define MYQC_HD
#endif

namespace myqc {

MYQC_HD constexpr int h_index(int N, int L, int M) {
    const int d = N + L + M;
    return d * (d + 1) * (d + 2) / 6 + (d - N) * (d - N + 1) / 2 + (d - N - L);
}
MYQC_HD constexpr int h_count(int deg) { return (deg + 1) * (deg + 2) * (deg + 3) / 6; }

MYQC_HD constexpr int h_N(int idx) {
    for (int d = 0; d <= 4; ++d)
        for (int N = d; N >= 0; --N)
            for (int L = d - N; L >= 0; --L)
                if (h_index(N, L, d - N - L) == idx) return N;
    return -1;
}
MYQC_HD constexpr int h_L(int idx) {
    for (int d = 0; d <= 4; ++d)
        for (int N = d; N >= 0; --N)
            for (int L = d - N; L >= 0; --L)
                if (h_index(N, L, d - N - L) == idx) return L;
    return -1;
}
MYQC_HD constexpr int h_M(int idx) {
    for (int d = 0; d <= 4; ++d)
        for (int N = d; N >= 0; --N)
            for (int L = d - N; L >= 0; --L)
                if (h_index(N, L, d - N - L) == idx) return d - N - L;
    return -1;
}
MYQC_HD constexpr int h_add(int a, int b) {
    return h_index(h_N(a) + h_N(b), h_L(a) + h_L(b), h_M(a) + h_M(b));
}
MYQC_HD constexpr int h_parity(int a) { return (h_N(a) + h_L(a) + h_M(a)) & 1; }

MYQC_HD constexpr int tt_nf(int t) { return t == 0 ? 1 : (t == 1 ? 4 : 16); }
MYQC_HD constexpr int tt_nterm(int t) { return t == 0 ? 1 : (t == 1 ? 7 : 46); }
MYQC_HD constexpr int tt_nh(int t) { return h_count(t); }  // 1, 4, 10 Hermite indices
// primitive record: p, Px, Py, Pz, E, 1/sqrt(p), coef[nterm], padded to an even number of doubles
constexpr int kRecCoef = 6;
MYQC_HD constexpr int tt_nfield(int t) { return ((kRecCoef + tt_nterm(t)) + 1) / 2 * 2; }

// unit vector index of axis w (1..3) -> Hermite index of e_w is simply w
// SP.SP function pair f = 4*mu + nu, mu,nu in {0=s,1=x,2=y,3=z}.
// number of terms of SP.SP function pair (mu,nu)
MYQC_HD constexpr int spsp_nterm(int mu, int nu) {
    return (mu == 0 && nu == 0) ? 1 : ((mu == 0 || nu == 0) ? 2 : (mu == nu ? 3 : 4));
}
// Hermite index of the j-th term of SP.SP function pair (mu,nu)
MYQC_HD constexpr int spsp_term_h(int mu, int nu, int j) {
    if (j == 0) return 0;
    if (mu == 0) return nu;                 // (s,w): e_w
    if (nu == 0) return mu;                 // (w,s): e_w
    if (mu == nu) return j == 1 ? mu : h_add(mu, mu);        // (w,w): e_w, 2e_w
    return j == 1 ? mu : (j == 2 ? nu : h_add(mu, nu));      // (w,w'): e_w, e_w', e_w+e_w'
}

MYQC_HD constexpr int term_fn(int t, int k) {
    if (t == 0) return 0;
    if (t == 1) return (k + 1) / 2;  // 0 | 1 1 | 2 2 | 3 3
    int c = 0;
    for (int f = 0; f < 16; ++f) {
        const int n = spsp_nterm(f / 4, f % 4);
        if (k < c + n) return f;
        c += n;
    }
    return -1;
}
MYQC_HD constexpr int term_h(int t, int k) {
    if (t == 0) return 0;
    if (t == 1) return (k == 0 || (k & 1)) ? 0 : k / 2;  // k: 0->0, 1->0, 2->1, 3->0, 4->2, 5->0, 6->3
    int c = 0;
    for (int f = 0; f < 16; ++f) {
        const int n = spsp_nterm(f / 4, f % 4);
        if (k < c + n) return spsp_term_h(f / 4, f % 4, k - c);
        c += n;
    }
    return -1;
}

static_assert(term_fn(2, 45) == 15 && term_h(2, 45) == h_index(0, 0, 2), "SP.SP term table");
static_assert(term_fn(1, 6) == 3 && term_h(1, 6) == 3 && term_h(1, 5) == 0, "S.SP term table");
static_assert(h_add(1, 2) == 5 && h_add(4, 9) == h_index(2, 0, 2), "Hermite index table");

}  // namespace myqc
