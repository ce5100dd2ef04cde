// Device-side Boys function and Hermite Coulomb recursion shared by the ERI class kernels
// (eri_kernels.cu) and the one-electron kernels (int1e.cu).
//
// Reference arithmetic: src/integrals/auxilary.f90:85-215 (Boys, Boys1/2/3), :265-285 (BoysG),
// :22-80 (RNLMj).  F_j(T) is obtained exactly as the reference does: 7-term Taylor expansion about the
// nearest Ftab node for T < 12, starting at order Q and recurring downwards; F0 = sqrt(pi)/2/sqrt(T) -
// exp(-T) g(T)/T with upward recursion for 12 <= T < 2Q+36; the bare asymptotic form above.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "terms.hpp"

namespace myqc {
namespace {

// 0.5 * Pi**0.5 with the reference's float32 Pi = 3.1415927410125732 (auxilary.f90:169,182,196,208)
constexpr double kHalfSqrtPi = 0.8862269377835134;

// compile-time loop: f(std::integral_constant<int,I>) for I in [0,N).  Forces every table lookup
// (term_fn, term_h, h_add ...) to be evaluated by the front end, so all accumulator indices are
// literal constants and the arrays live in registers.
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(static_cast<F&&>(f));
    }
}

// ------------------------------------------------------------------------------------------
// 1/sqrt(x) for positive normal x: MUFU.RSQ64H seed + one third-order correction (the arithmetic
// of CUDA's rsqrt() without its special-case branch).
__device__ __forceinline__ double rsqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double r = fma(x, -(y * y), 1.0);
    const double c = fma(r, 0.375, 0.5);
    return fma(c, y * r, y);
}

// exp(d) for |d| <= 0.06: 9-term series (truncation < 3e-20)
__device__ __forceinline__ double exp_small(double d) {
    double e = 1.0 / 362880.0;
    e = fma(e, d, 1.0 / 40320.0);
    e = fma(e, d, 1.0 / 5040.0);
    e = fma(e, d, 1.0 / 720.0);
    e = fma(e, d, 1.0 / 120.0);
    e = fma(e, d, 1.0 / 24.0);
    e = fma(e, d, 1.0 / 6.0);
    e = fma(e, d, 0.5);
    e = fma(e, d, 1.0);
    return fma(e, d, 1.0);
}

// NINT(x) for x >= 0: round half away from zero (auxilary.f90:152)
__device__ __forceinline__ int nint_pos(double x) {
    int k = (int)x;
    if (x - (double)k >= 0.5) ++k;
    return k;
}

// auxilary.f90:265-285; T >= 30 is undefined in the reference, we keep 0.490
__device__ __forceinline__ double boys_g(double T, double invT) {
    double c0 = 0.490, c1 = 0.0, c2 = 0.0, c3 = 0.0;
    if (T < 15.0) { c0 = 0.4999489092; c1 = -0.2473631686; c2 = 0.321180909; c3 = -0.3811559346; }
    else if (T < 18.0) { c0 = 0.4998436875; c1 = -0.24249438; c2 = 0.24642845; }
    else if (T < 24.0) { c0 = 0.499093162; c1 = -0.2152832; }
    return fma(invT, fma(invT, fma(invT, c3, c2), c1), c0);
}

// G_j = (-2 alpha)^j F_j(T) / sqrt(p+q), j = 0..LT, for the two regimes that need alpha and T
// (auxilary.f90:130-189).  s_ft row t: {Ft(t,Q+k)/k!, k=0..6 ; unused}; s_exp[k] = {exp(-k/10), k/10}.
template <int Q, int LT>
__device__ __forceinline__ void boys_near_mid(double T, double alpha, double rs, double (&G)[LT + 1],
                                              const double* __restrict__ s_ft,
                                              const double2* __restrict__ s_exp) {
    double F[LT + 1];
    const int Tk = nint_pos(T * 10.0);
    const double2 ex = s_exp[Tk];
    const double d = ex.y - T;  // Tk/10.0D0 - T
    if (T < 12.0) {
        const double2* row = reinterpret_cast<const double2*>(s_ft + Tk * 8);
        const double2 c01 = row[0], c23 = row[1], c45 = row[2], c6 = row[3];
        double f = c6.x;
        f = fma(f, d, c45.y);
        f = fma(f, d, c45.x);
        f = fma(f, d, c23.y);
        f = fma(f, d, c23.x);
        f = fma(f, d, c01.y);
        f = fma(f, d, c01.x);
        if (Q <= LT) F[Q] = f;
        if (Q > 0) {
            const double e = ex.x * exp_small(d);  // exp(-T)
            const double t2 = 2.0 * T;
#pragma unroll
            for (int j = Q - 1; j >= 0; --j) {
                f = fma(t2, f, e) * (1.0 / (2.0 * j + 1.0));
                if (j <= LT) F[j] = f;
            }
        }
    } else {
        const double rT = rsqrt_pos(T);
        const double invT = rT * rT;
        const double e = ex.x * exp_small(d);
        double f = fma(-e * boys_g(T, invT), invT, kHalfSqrtPi * rT);
        F[0] = f;
        const double h = 0.5 * invT;
#pragma unroll
        for (int j = 1; j <= LT; ++j) {
            f = h * fma((double)(2 * j - 1), f, -e);
            F[j] = f;
        }
    }
    const double m2a = -2.0 * alpha;
    double w = rs;
#pragma unroll
    for (int j = 0; j <= LT; ++j) {
        G[j] = w * F[j];
        w *= m2a;
    }
}

// ------------------------------------------------------------------------------------------
// Hermite Coulomb integrals R_{NLM} = R^{(0)}_{NLM}, auxilary.f90:22-80: N is reduced first, then
// L, then M.  In place: level j overwrites level j+1 from the highest degree downwards.
template <int LT>
__device__ __forceinline__ void build_R(const double (&G)[LT + 1], double X, double Y, double Z,
                                        double (&R)[h_count(LT)]) {
    R[0] = G[LT];
    static_for<0, LT>([&](auto jc) {
        constexpr int j = LT - 1 - decltype(jc)::value;
        constexpr int ne = h_count(LT - j);
        static_for<0, ne - 1>([&](auto ec) {
            constexpr int e = ne - 1 - decltype(ec)::value;  // ne-1 ... 1: degree descending
            constexpr int N = h_N(e), L = h_L(e), M = h_M(e);
            if constexpr (N > 0) {
                double v = X * R[h_index(N - 1, L, M)];
                if constexpr (N > 1) v = fma((double)(N - 1), R[h_index(N - 2, L, M)], v);
                R[e] = v;
            } else if constexpr (L > 0) {
                double v = Y * R[h_index(0, L - 1, M)];
                if constexpr (L > 1) v = fma((double)(L - 1), R[h_index(0, L - 2, M)], v);
                R[e] = v;
            } else {
                double v = Z * R[h_index(0, 0, M - 1)];
                if constexpr (M > 1) v = fma((double)(M - 1), R[h_index(0, 0, M - 2)], v);
                R[e] = v;
            }
        });
        R[0] = G[j];
    });
}


}  // namespace
}  // namespace myqc
