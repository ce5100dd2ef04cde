// Quartet-class ERI kernels for sm_100a (B200).
//
// What is computed (reference: src/integrals/int2e.f90:618-726 clmnew, auxilary.f90:22-215):
// for every canonical contracted shell quartet (u | v) that passes the reference's
// EIJ*EGH >= 1e-14 rule, the sum over primitive quartets of
//     ll * sum_{k,k'} (-1)^{N'+L'+M'} D_k D'_k' R_{N+N',L+L',M+M'}(alpha, P-Q)
// with Boys values F_j(T) obtained exactly as the reference does (7-term Taylor expansion
// about the nearest Ftab node for T < 12 starting at order Q = 3*(number of SP sets) and
// recurring downwards; asymptotic forms above).
//
// Mapping onto the B200:
//   * one kernel instantiation per quartet class (UT, TT) = (#SP sets in the uniform-side pair,
//     #SP sets in the lane-side pair); all loops over Hermite terms are compile-time unrolled,
//     so K-, R- and output accumulators live in registers (FP64 DFMA pipe bound).
//   * persistent CTAs; each CTA iteration owns one row u: its primitive-pair record block
//     (<= 3.7 KB) is staged into shared memory with one TMA bulk copy (cp.async.bulk +
//     mbarrier) and then read as warp-uniform broadcasts.
//   * each lane owns one lane-side pair v (coalesced SoA loads) and loops over primitive
//     pairs: for each lane-side primitive, accumulate K[f][H'] over the uniform-side primitives
//     (step A: |terms_U| x |H_T| DFMA per primitive quartet), then fold the lane-side
//     coefficients once (step B), i.e. the second half-contraction is hoisted out of the
//     primitive-quartet loop.
//   * the Boys table of the class (121 x 8 doubles, pre-divided by k!) lives in shared memory.
//   * (SP SP|SP SP) is split into four mu-slices of the uniform side so that 40 K- and 64
//     output accumulators fit; outputs of the big classes are kept in lane-private shared
//     memory between lane-side primitives.
//   * no tensor cores: this is not a dense contraction.
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "eri_kernels.cuh"
#include "terms.hpp"

namespace myqc {

namespace {

constexpr double kScreen = 1.0e-14;  // int2e.f90:257

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
// TMA bulk copy + mbarrier helpers (sm_90+ PTX; SASS: UBLKCP / SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ------------------------------------------------------------------------------------------
// Boys function, auxilary.f90:85-215.  F[0..LN] are returned; the T < 12 branch starts at order
// Q like the reference and recurs downwards through all orders (T3 in SURVEY.md).
// s_ft row t: {Ft(t,Q+k)/k!, k=0..6 ; t/10.0}.
__device__ __forceinline__ double boys_g(double T, double invT) {
    // auxilary.f90:265-285; T >= 30 is undefined in the reference, we keep 0.490 (T5)
    double c0 = 0.490, c1 = 0.0, c2 = 0.0, c3 = 0.0;
    if (T < 15.0) { c0 = 0.4999489092; c1 = -0.2473631686; c2 = 0.321180909; c3 = -0.3811559346; }
    else if (T < 18.0) { c0 = 0.4998436875; c1 = -0.24249438; c2 = 0.24642845; }
    else if (T < 24.0) { c0 = 0.499093162; c1 = -0.2152832; }
    return fma(invT, fma(invT, fma(invT, c3, c2), c1), c0);
}

template <int Q, int LN>
__device__ __forceinline__ void boys(double T, double (&F)[LN + 1], const double* __restrict__ s_ft) {
    // 0.5 * Pi**0.5 with the reference's float32 Pi (auxilary.f90:169,182)
    constexpr double kHalfSqrtPi = 0.8862269377835134;  // 0.5*sqrt(3.1415927410125732)
    if (T < 12.0) {
        // Tk = NINT(T*10): round half away from zero (T7)
        const double x = T * 10.0;
        int Tk = (int)x;
        if (x - (double)Tk >= 0.5) ++Tk;
        const double2* row = reinterpret_cast<const double2*>(s_ft + Tk * 8);
        const double2 c01 = row[0], c23 = row[1], c45 = row[2], c6t = row[3];
        const double d = c6t.y - T;  // Tk/10.0D0 - T
        double f = c6t.x;
        f = fma(f, d, c45.y);
        f = fma(f, d, c45.x);
        f = fma(f, d, c23.y);
        f = fma(f, d, c23.x);
        f = fma(f, d, c01.y);
        f = fma(f, d, c01.x);
        if (Q <= LN) F[Q] = f;
        if (Q > 0) {
            const double e = exp(-T);
            const double t2 = 2.0 * T;
#pragma unroll
            for (int j = Q - 1; j >= 0; --j) {
                f = fma(t2, f, e) * (1.0 / (2.0 * j + 1.0));
                if (j <= LN) F[j] = f;
            }
        }
    } else {
        const double r = rsqrt(T);
        const double invT = r * r;
        double f = kHalfSqrtPi * r;
        double e = 0.0;
        if (T < (double)(2 * Q + 36)) {
            e = exp(-T);
            f = fma(-e * boys_g(T, invT), invT, f);
        }
        F[0] = f;
        const double h = 0.5 * invT;
#pragma unroll
        for (int j = 1; j <= LN; ++j) {
            f = h * fma((double)(2 * j - 1), f, -e);
            F[j] = f;
        }
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

// smem layout: [ftab 121*8 doubles][u record 9*FU doubles][mbarrier 8 B (+8 pad)][out staging]
template <int UT, int TT, int USL>
static constexpr size_t smem_bytes() {
    size_t b = 121 * 8 * 8 + 9 * tt_nfield(UT) * 8 + 16;
    constexpr int nout = ((USL >= 0) ? 4 : tt_nf(UT)) * tt_nf(TT);
    if (nout > 16) b += (size_t)nout * 128 * 8;
    return b;
}

template <int UT, int TT, int USL>
__global__ void __launch_bounds__(((((USL >= 0) ? 4 : tt_nf(UT)) * tt_nf(TT)) > 16) ? 128 : 256)
    eri_class_kernel(const ClassArgs a) {
    constexpr int LT = UT + TT;
    constexpr int Q = 3 * LT;
    constexpr int NR = h_count(LT);
    constexpr int NHT = tt_nh(TT);
    constexpr int NFU = (USL >= 0) ? 4 : tt_nf(UT);
    constexpr int NFU_FULL = tt_nf(UT);
    constexpr int NFT = tt_nf(TT);
    constexpr int NTU = tt_nterm(UT);
    constexpr int NTT = tt_nterm(TT);
    constexpr int FU = tt_nfield(UT);
    constexpr int FT = tt_nfield(TT);
    constexpr int NOUT = NFU * NFT;
    constexpr bool OUT_SMEM = (NOUT > 16);
    constexpr int NTHREADS = OUT_SMEM ? 128 : 256;
    constexpr int NWARPS = NTHREADS / 32;
    constexpr uint32_t U_BYTES = 9 * FU * 8;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_ft = reinterpret_cast<double*>(smem_raw);
    double* s_u = s_ft + 121 * 8;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_u + 9 * FU);
    double* s_out = reinterpret_cast<double*>(s_bar + 2);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    for (int i = tid; i < 121 * 8; i += NTHREADS) s_ft[i] = a.ftab_q[i];
    if (tid == 0) {
        mbar_init(s_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t parity = 0;
    for (int u = blockIdx.x; u < a.nU; u += gridDim.x) {
        if (tid == 0) {
            mbar_expect_tx(s_bar, U_BYTES);
            tma_bulk_g2s(s_u, a.u_aos + (size_t)u * 9 * FU, U_BYTES, s_bar);
        }
        const int npu = a.u_nprim[u];
        const int ntv = a.u_ntv[u];
        const int v0 = a.tri ? u : 0;
        const int udiag = a.u_diag[u];
        mbar_wait(s_bar, parity);
        parity ^= 1;
        const double eu_max = s_u[4];

        for (int vb = v0 + warp * 32; vb < ntv; vb += NWARPS * 32) {
            const int v = vb + lane;
            if (v < ntv) {
                double out_r[OUT_SMEM ? 1 : NOUT];
                if (OUT_SMEM) {
#pragma unroll
                    for (int o = 0; o < NOUT; ++o) s_out[o * NTHREADS + tid] = 0.0;
                } else {
#pragma unroll
                    for (int o = 0; o < NOUT; ++o) out_r[o] = 0.0;
                }
                const int npt = a.t_nprim[v];
                for (int kt = 0; kt < npt; ++kt) {
                    const double* tp = a.t_soa + (size_t)kt * FT * a.t_npad + v;
                    const double et = tp[4 * (size_t)a.t_npad];
                    if (eu_max * et < kScreen) break;  // prims sorted by E descending
                    const double q = tp[0];
                    const double Qx = tp[(size_t)a.t_npad], Qy = tp[2 * (size_t)a.t_npad],
                                 Qz = tp[3 * (size_t)a.t_npad];
                    double K[NFU][NHT];
#pragma unroll
                    for (int f = 0; f < NFU; ++f)
#pragma unroll
                        for (int h = 0; h < NHT; ++h) K[f][h] = 0.0;

                    for (int ku = 0; ku < npu; ++ku) {
                        const double* up = s_u + ku * FU;
                        if (up[4] * et < kScreen) break;  // IF (EGH*EIJ .LT. 1.0D-14) CYCLE
                        const double p = up[0];
                        const double X = up[1] - Qx, Y = up[2] - Qy, Z = up[3] - Qz;
                        const double s = p + q;
                        const double rs = rsqrt(s);
                        const double alpha = p * q * (rs * rs);
                        const double T = alpha * (X * X + Y * Y + Z * Z);
                        double F[LT + 1];
                        boys<Q, LT>(T, F, s_ft);
                        // G_j = (-2 alpha)^j F_j / sqrt(p+q)   (R_000^j, auxilary.f90:51; ll folded)
                        double G[LT + 1];
                        {
                            const double m2a = -2.0 * alpha;
                            double w = rs;
#pragma unroll
                            for (int j = 0; j <= LT; ++j) {
                                G[j] = w * F[j];
                                w *= m2a;
                            }
                        }
                        double R[NR];
                        build_R<LT>(G, X, Y, Z, R);
                        // step A: K[f][H'] += D_k * R[H_k + H']
                        static_for<0, NTU>([&](auto kc) {
                            constexpr int k = decltype(kc)::value;
                            constexpr int f = term_fn(UT, k);
                            if constexpr (USL < 0 || f / 4 == USL) {
                                constexpr int lf = (USL >= 0) ? (f % 4) : f;
                                constexpr int hk = term_h(UT, k);
                                const double cu = up[5 + k];
                                static_for<0, NHT>([&](auto hc) {
                                    constexpr int hp = decltype(hc)::value;
                                    constexpr int ri = h_add(hk, hp);
                                    K[lf][hp] = fma(cu, R[ri], K[lf][hp]);
                                });
                            }
                        });
                    }
                    // step B: out[f][f'] += (-1)^{|H'|} D'_k' K[f][H'_k']
                    static_for<0, NTT>([&](auto kc) {
                        constexpr int k = decltype(kc)::value;
                        constexpr int fp = term_fn(TT, k);
                        constexpr int hp = term_h(TT, k);
                        double ct = tp[(size_t)(5 + k) * a.t_npad];
                        if constexpr (h_parity(hp) != 0) ct = -ct;
#pragma unroll
                        for (int f = 0; f < NFU; ++f) {
                            if constexpr (OUT_SMEM) {
                                double* o = &s_out[(f * NFT + fp) * NTHREADS + tid];
                                *o = fma(ct, K[f][hp], *o);
                            } else {
                                out_r[f * NFT + fp] = fma(ct, K[f][hp], out_r[f * NFT + fp]);
                            }
                        }
                    });
                }
                // store the distinct canonical integrals of this shell quartet
                const int tdiag = a.t_diag[v];
                const bool same_pair = a.tri && (v == u);
                const int64_t n = a.norb;
#pragma unroll
                for (int f = 0; f < NFU; ++f) {
                    const int fu = (USL >= 0) ? (4 * USL + f) : f;
                    const int i0 = a.u_fi[(size_t)u * NFU_FULL + fu];
                    const int j0 = a.u_fj[(size_t)u * NFU_FULL + fu];
                    if (i0 < 0 || (udiag && i0 > j0)) continue;
                    const int64_t i = i0 < j0 ? i0 : j0, j = i0 < j0 ? j0 : i0;
                    const int64_t P1 = i * n - i * (i - 1) / 2 + (j - i);
#pragma unroll
                    for (int fp = 0; fp < NFT; ++fp) {
                        const int g0 = a.t_fi[(size_t)v * NFT + fp];
                        const int h0 = a.t_fj[(size_t)v * NFT + fp];
                        if (g0 < 0 || (tdiag && g0 > h0)) continue;
                        const int64_t g = g0 < h0 ? g0 : h0, h = g0 < h0 ? h0 : g0;
                        const int64_t P2 = g * n - g * (g - 1) / 2 + (h - g);
                        if (same_pair && P1 > P2) continue;
                        const int64_t lo = P1 < P2 ? P1 : P2, hi = P1 < P2 ? P2 : P1;
                        const int64_t idx = lo * a.npair - lo * (lo - 1) / 2 + (hi - lo) - a.out_offset;
                        const double val = OUT_SMEM ? s_out[(f * NFT + fp) * NTHREADS + tid] : out_r[f * NFT + fp];
                        a.out[idx] = val;
                    }
                }
            }
        }
        __syncthreads();  // every warp is done with s_u before the next row's TMA overwrites it
    }
}

// ------------------------------------------------------------------------------------------
__global__ void fill_zero_kernel(double* __restrict__ out, int64_t n) {
    // 128-bit stores, grid-stride; a slice may start on an odd element
    const int64_t head = ((reinterpret_cast<uintptr_t>(out) & 15) != 0 && n > 0) ? 1 : 0;
    const int64_t n2 = (n - head) / 2;
    double2* o2 = reinterpret_cast<double2*>(out + head);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride)
        o2[i] = make_double2(0.0, 0.0);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (head) out[0] = 0.0;
        if ((n - head) & 1) out[n - 1] = 0.0;
    }
}

// dense XX(i,j,g,h) (column-major, i fastest) from the packed array: the fillsym pass of the
// reference (int2e.f90:290-304,540-554) done as a gather so that the 8n^4-byte stream is written
// once, coalesced.
__global__ void expand_dense_kernel(const double* __restrict__ packed, int norb, double* __restrict__ xx) {
    const int64_t n = norb;
    const int64_t npair = n * (n + 1) / 2;
    const int64_t total = n * n * n * n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        int64_t r = e;
        const int64_t i = r % n; r /= n;
        const int64_t j = r % n; r /= n;
        const int64_t g = r % n; r /= n;
        const int64_t h = r;
        const int64_t a = i < j ? i : j, b = i < j ? j : i;
        const int64_t c = g < h ? g : h, d = g < h ? h : g;
        const int64_t P1 = a * n - a * (a - 1) / 2 + (b - a);
        const int64_t P2 = c * n - c * (c - 1) / 2 + (d - c);
        const int64_t lo = P1 < P2 ? P1 : P2, hi = P1 < P2 ? P2 : P1;
        xx[e] = packed[lo * npair - lo * (lo - 1) / 2 + (hi - lo)];
    }
}

// ------------------------------------------------------------------------------------------
template <int UT, int TT, int USL>
static int launch_one(const ClassArgs& a, int num_sms, cudaStream_t st) {
    if (a.nU <= 0 || a.nT <= 0) return 0;
    constexpr int nout = ((USL >= 0) ? 4 : tt_nf(UT)) * tt_nf(TT);
    constexpr int nthreads = nout > 16 ? 128 : 256;
    const size_t smem = smem_bytes<UT, TT, USL>();
    auto kern = eri_class_kernel<UT, TT, USL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int occ = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nthreads, smem);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) occ = 1;
    int grid = num_sms * occ;
    if (grid > a.nU) grid = a.nU;
    kern<<<grid, nthreads, smem, st>>>(a);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// FP64 roofline denominator: 8 independent DFMA chains per thread, 256 threads, 8 CTAs per SM.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int measure_dfma_peak(int num_sms, double* tflops) {
    const int grid = num_sms * 8, iters = 4096;
    double* d = nullptr;
    cudaError_t e = cudaMalloc((void**)&d, (size_t)grid * 256 * sizeof(double));
    if (e != cudaSuccess) return (int)e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(t0);
        dfma_peak_kernel<<<grid, 256>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(t1);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, t0, t1);
        const double flops = 2.0 * 8 * 16 * (double)iters * 256.0 * grid;
        if (rep > 0) best = best > flops / (ms * 1e-3) / 1e12 ? best : flops / (ms * 1e-3) / 1e12;
    }
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    cudaFree(d);
    *tflops = best;
    return (int)e;
}

int class_nlaunch(int UT, int TT) { return (UT == 2 && TT == 2) ? 4 : 1; }

int launch_class(int UT, int TT, const ClassArgs& a, int num_sms, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (UT == 0 && TT == 0) return launch_one<0, 0, -1>(a, num_sms, st);
    if (UT == 0 && TT == 1) return launch_one<0, 1, -1>(a, num_sms, st);
    if (UT == 0 && TT == 2) return launch_one<0, 2, -1>(a, num_sms, st);
    if (UT == 1 && TT == 1) return launch_one<1, 1, -1>(a, num_sms, st);
    if (UT == 1 && TT == 2) return launch_one<1, 2, -1>(a, num_sms, st);
    if (UT == 2 && TT == 2) {
        int e = launch_one<2, 2, 0>(a, num_sms, st);
        if (!e) e = launch_one<2, 2, 1>(a, num_sms, st);
        if (!e) e = launch_one<2, 2, 2>(a, num_sms, st);
        if (!e) e = launch_one<2, 2, 3>(a, num_sms, st);
        return e;
    }
    return (int)cudaErrorInvalidValue;
}

int launch_fill_zero(double* out, int64_t n, int num_sms, void* stream) {
    if (n <= 0) return 0;
    fill_zero_kernel<<<num_sms * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n);
    return (int)cudaGetLastError();
}

int launch_expand_dense(const double* packed, int norb, double* xx, int num_sms, void* stream) {
    expand_dense_kernel<<<num_sms * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(packed, norb, xx);
    return (int)cudaGetLastError();
}

}  // namespace myqc
