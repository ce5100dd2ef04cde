// Quartet-class ERI kernels for sm_100a (B200).
//
// What is computed (reference: src/integrals/int2e.f90:618-726 clmnew, auxilary.f90:22-215):
// for every canonical contracted shell quartet (u | v) that passes the reference's
// EIJ*EGH >= 1e-14 rule, the sum over primitive quartets of
//     ll * sum_{k,k'} (-1)^{N'+L'+M'} D_k D'_k' R_{N+N',L+L',M+M'}(alpha, P-Q)
// with Boys values F_j(T) obtained exactly as the reference does: 7-term Taylor expansion about
// the nearest Ftab node for T < 12, starting at order Q = 3*(number of SP sets) and recurring
// downwards (Boys1); F0 = sqrt(pi)/2/sqrt(T) - exp(-T) g(T)/T with upward recursion for
// 12 <= T < 2Q+36 (Boys2); the bare asymptotic form above (Boys3).
//
// Mapping onto the B200:
//   * one kernel instantiation per quartet class (UT, TT) = (#SP sets in the uniform-side pair,
//     #SP sets in the lane-side pair); all loops over Hermite terms are unrolled at compile time
//     (static_for), so the K-, R- and output accumulators live in registers and the inner loop is
//     straight-line DFMA code (FP64 pipe bound; no tensor cores: this is not a dense contraction).
//   * warps are autonomous: each warp pulls rows u of the quartet space from a global counter,
//     stages the row's primitive-pair record block (<= 3.7 KB) into its own shared-memory double
//     buffer with a TMA bulk copy (cp.async.bulk + mbarrier) while it still works on the previous
//     row, and reads it back as warp-uniform broadcasts.  No CTA-wide barrier in the main loop.
//   * each lane owns one lane-side pair v (coalesced SoA loads).  Pair lists are ordered by
//     prefactor bucket and Morton code of the pair centre, so the 32 pairs of a warp are spatially
//     close: they sit in the same Boys regime and lose the same primitives to the screen.
//   * for each lane-side primitive, K[f][H'] is accumulated over the uniform-side primitives
//     (step A: |terms_U| x |H_T| DFMA per primitive quartet); the lane-side coefficients are folded
//     once afterwards (step B) -- the second half-contraction is hoisted out of the inner loop.
//   * far-field quartets (T >= 2Q+36, the bulk of a large molecule) need one reciprocal square
//     root and no exponential: G_j = sqrt(pi)/2 /sqrt(p q) (2j-1)!! (-1)^j / R^(2j+1).
//   * the Boys Taylor table of the class (121 x 8 doubles, pre-divided by k!) and an exp(-k/10)
//     table live in shared memory; exp(-T) = exp(-k/10) * exp(k/10 - T) with a 9-term series.
//   * (SP SP|SP SP) is split into four mu-slices of the uniform side so that 40 K- and 64 output
//     accumulators fit; outputs of the two big classes are kept in lane-private shared memory
//     between lane-side primitives.
#include <cuda_runtime.h>

#include <climits>
#include <cstdint>
#include <cstdlib>
#include <type_traits>

#include "eri_kernels.cuh"
#include "terms.hpp"
#include "boys.cuh"

namespace myqc {

namespace {

constexpr double kScreen = 1.0e-14;  // int2e.f90:257

// Store of one integral into the zero-filled packed array.  -DMYQC_STORE_OP=1/2/3 builds the cache-hint
// variants tools/build_variants.sh measures (st.global.cs / .cg / .wt); the default is a plain store.
#ifndef MYQC_STORE_OP
#define MYQC_STORE_OP 0
#endif
__device__ __forceinline__ void store_eri(double* p, double v) {
#if MYQC_STORE_OP == 1
    __stcs(p, v);
#elif MYQC_STORE_OP == 2
    __stcg(p, v);
#elif MYQC_STORE_OP == 3
    __stwt(p, v);
#else
    *p = v;
#endif
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

template <int UT, int TT, int USL>
struct Cfg {
    static constexpr int LT = UT + TT;
    static constexpr int Q = 3 * LT;  // Boys start order, int2e.f90:654-658,668
    static constexpr int NR = h_count(LT);
    static constexpr int NHT = tt_nh(TT);
    static constexpr int NFU = (USL >= 0) ? 4 : tt_nf(UT);
    static constexpr int NFU_FULL = tt_nf(UT);
    static constexpr int NFT = tt_nf(TT);
    static constexpr int NTU = tt_nterm(UT);
    static constexpr int NTT = tt_nterm(TT);
    static constexpr int FU = tt_nfield(UT);
    static constexpr int FT = tt_nfield(TT);
    static constexpr int NOUT = NFU * NFT;
    static constexpr bool OUT_SMEM = (NOUT > 16);
    // 128-thread CTAs for every class: (S SP|S SP) needs 140 registers, and a 256-thread CTA of it
    // would leave an SM with a single resident CTA (8 warps); four-warp CTAs pack 3 per SM
    static constexpr int NTHREADS = 128;
    static constexpr int NWARPS = NTHREADS / 32;
    // resident CTAs per SM the register allocation is held to: 72 / 80 / 128 / 140 registers for the
    // four small classes, two CTAs for the two large ones
    static constexpr int MINB = (LT == 0) ? 7 : (LT == 1) ? 6 : (UT == 0 && TT == 2) ? 4 : (LT == 2) ? 3 : 2;
    static constexpr uint32_t U_BYTES = 9 * FU * 8;
    // shared memory: Taylor table | exp table | per-warp U double buffers | mbarriers | out staging
    static constexpr size_t OFF_EXP = 121 * 8 * 8;
    static constexpr size_t OFF_U = OFF_EXP + 608 * 16;
    static constexpr size_t OFF_BAR = OFF_U + (size_t)NWARPS * 2 * U_BYTES;
    static constexpr bool LANE_AOS = (UT >= 1);  // measured: pays for {1,1},{1,2},{2,2}, costs for {0,x}
    static constexpr int TASKP = class_task_pairs(UT, TT);                 // lane-side pairs per task
    static constexpr size_t OFF_SORT = OFF_BAR + (size_t)NWARPS * 2 * 8;   // per warp: TASKP idx + 64 bins (int)
    static constexpr size_t OFF_OUT = OFF_SORT + (size_t)NWARPS * (TASKP + 64) * 4;
    static constexpr size_t SMEM = OFF_OUT + (OUT_SMEM ? (size_t)NOUT * NTHREADS * 8 : 0);
};

}  // namespace

// MULTI: the launch has several parts (a shard's own x own, own x later, later x own lists) and task.w selects one;
// with MULTI = false every reference to the part is a kernel-parameter operand at a fixed offset, as before the merge.
template <int UT, int TT, int USL, bool MULTI>
__global__ void __launch_bounds__(Cfg<UT, TT, USL>::NTHREADS, Cfg<UT, TT, USL>::MINB) eri_class_kernel(const __grid_constant__ ClassArgs a) {
    using C = Cfg<UT, TT, USL>;
    constexpr int LT = C::LT, Q = C::Q, NR = C::NR, NHT = C::NHT, NFU = C::NFU, NFT = C::NFT;
    constexpr int NTU = C::NTU, NTT = C::NTT, FU = C::FU, FT = C::FT, NOUT = C::NOUT;
    constexpr int NTHREADS = C::NTHREADS;
    constexpr bool OUT_SMEM = C::OUT_SMEM;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* s_ft = reinterpret_cast<double*>(smem_raw);
    double2* s_exp = reinterpret_cast<double2*>(smem_raw + C::OFF_EXP);
    double* s_out = reinterpret_cast<double*>(smem_raw + C::OFF_OUT);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    double* s_ubuf = reinterpret_cast<double*>(smem_raw + C::OFF_U) + (size_t)warp * 2 * 9 * FU;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw + C::OFF_BAR) + warp * 2;
    int* s_vidx = reinterpret_cast<int*>(smem_raw + C::OFF_SORT) + warp * (C::TASKP + 64);
    int* s_bin = s_vidx + C::TASKP;

    for (int i = tid; i < 121 * 8; i += NTHREADS) s_ft[i] = a.ftab_q[i];
    for (int i = tid; i < 601; i += NTHREADS) s_exp[i] = a.exptab[i];
    if (lane == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // ---- warp-autonomous task loop with a two-deep TMA prefetch ------------------------------
    // task = (row u, lane-side range [vb0, vend)): rows are cut into pieces of at most
    // kTaskPairs lane-side pairs so that no warp owns more than a sliver of the launch.
    int t = 0;
    int4 task = make_int4(0, 0, 0, 0);
    if (lane == 0) {
        t = atomicAdd(a.row_counter, 1);
        if (t < a.ntasks) {
            task = a.tasks[t];
            mbar_expect_tx(&s_bar[0], C::U_BYTES);
            tma_bulk_g2s(s_ubuf, a.part[MULTI ? task.w : 0].u_aos + (size_t)task.x * 9 * FU, C::U_BYTES, &s_bar[0]);
        }
    }
    t = __shfl_sync(0xffffffffu, t, 0);
    task.x = __shfl_sync(0xffffffffu, task.x, 0);
    task.y = __shfl_sync(0xffffffffu, task.y, 0);
    task.z = __shfl_sync(0xffffffffu, task.z, 0);
    if (MULTI) task.w = __shfl_sync(0xffffffffu, task.w, 0);
    int buf = 0;
    uint32_t parity0 = 0, parity1 = 0;
    while (t < a.ntasks) {
        int tn = 0;
        int4 taskn = make_int4(0, 0, 0, 0);
        if (lane == 0) {
            tn = atomicAdd(a.row_counter, 1);
            if (tn < a.ntasks) {
                taskn = a.tasks[tn];
                mbar_expect_tx(&s_bar[buf ^ 1], C::U_BYTES);
                tma_bulk_g2s(s_ubuf + (size_t)(buf ^ 1) * 9 * FU, a.part[MULTI ? taskn.w : 0].u_aos + (size_t)taskn.x * 9 * FU, C::U_BYTES,
                             &s_bar[buf ^ 1]);
            }
        }
        tn = __shfl_sync(0xffffffffu, tn, 0);
        taskn.x = __shfl_sync(0xffffffffu, taskn.x, 0);
        taskn.y = __shfl_sync(0xffffffffu, taskn.y, 0);
        taskn.z = __shfl_sync(0xffffffffu, taskn.z, 0);
        if (MULTI) taskn.w = __shfl_sync(0xffffffffu, taskn.w, 0);
        const PartArgs& S = a.part[MULTI ? task.w : 0];
        const int u = task.x;
        const int v0 = task.y;
        const int ntv = task.z;
        const int npu = S.u_nprim[u];
        int P1[NFU];
#pragma unroll
        for (int f = 0; f < NFU; ++f) P1[f] = S.u_pidx[(size_t)u * C::NFU_FULL + ((USL >= 0) ? (4 * USL + f) : f)];
        if (buf == 0) { mbar_wait(&s_bar[0], parity0); parity0 ^= 1; }
        else          { mbar_wait(&s_bar[1], parity1); parity1 ^= 1; }
        const double* s_u = s_ubuf + (size_t)buf * 9 * FU;
        const double eu_max = s_u[4];

        // ---- order the task's lane-side pairs by their distance from the row's pair -----------
        // The Boys regime of a primitive quartet is set by alpha*|P-Q|^2; pairs of one task are of
        // one kind (same exponents), so after a counting sort on |P-Q|^2 the 32 lanes of a chunk
        // take the same branch.  Which lane gets which pair does not affect any result.
        const int nitem = ntv - v0;
        if (C::TASKP > 32 && nitem > 32) {
            const double Px = s_u[1], Py = s_u[2], Pz = s_u[3];
            double key[C::TASKP / 32];
            double kmin = 1.0e300, kmax = 0.0;
#pragma unroll
            for (int i = 0; i < C::TASKP / 32; ++i) {
                const int v = v0 + i * 32 + lane;
                key[i] = -1.0;
                if (v < ntv) {
                    double dx, dy, dz;
                    if constexpr (C::LANE_AOS) {
                        const double* r0 = S.t_aos + (size_t)v * 9 * FT;
                        dx = Px - __ldg(r0 + 1);
                        dy = Py - __ldg(r0 + 2);
                        dz = Pz - __ldg(r0 + 3);
                    } else {
                        dx = Px - S.t_soa[(size_t)S.t_npad + v];
                        dy = Py - S.t_soa[2 * (size_t)S.t_npad + v];
                        dz = Pz - S.t_soa[3 * (size_t)S.t_npad + v];
                    }
                    key[i] = fma(dx, dx, fma(dy, dy, dz * dz));
                    kmin = fmin(kmin, key[i]);
                    kmax = fmax(kmax, key[i]);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                kmin = fmin(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
                kmax = fmax(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
            }
            const double scale = 63.999 / (kmax - kmin + 1.0e-300);
            s_bin[lane] = 0;
            s_bin[lane + 32] = 0;
            __syncwarp();
#pragma unroll
            for (int i = 0; i < C::TASKP / 32; ++i)
                if (key[i] >= 0.0) atomicAdd(&s_bin[(int)((key[i] - kmin) * scale)], 1);
            __syncwarp();
            // exclusive scan of the 64 bin counts (two bins per lane)
            const int c0 = s_bin[2 * lane], c1 = s_bin[2 * lane + 1];
            int incl = c0 + c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            __syncwarp();
            s_bin[2 * lane] = incl - c0 - c1;
            s_bin[2 * lane + 1] = incl - c1;
            __syncwarp();
#pragma unroll
            for (int i = 0; i < C::TASKP / 32; ++i)
                if (key[i] >= 0.0) s_vidx[atomicAdd(&s_bin[(int)((key[i] - kmin) * scale)], 1)] = v0 + i * 32 + lane;
            __syncwarp();
        } else {
            s_vidx[lane] = v0 + lane;
            __syncwarp();
        }

        // One lane-side primitive (record tp) against the row's primitives: K[f][H'] over the row primitives
        // (step A), then the lane-side coefficients folded in (step B); every out[f][f'] contribution goes to
        // sink(f, f', value).  Returns false when the primitive is below the screen against the whole row.
        bool any = false;
        unsigned npq = 0;  // primitive quartets this lane evaluated in this task
        const double qu = (a.tau > 0.0) ? __ldg(S.u_q + u) : 0.0;
        const size_t ld = C::LANE_AOS ? (size_t)1 : (size_t)S.t_npad;
        auto contract_prim = [&](const double* tp, bool has_next, auto&& sink) -> bool {
            double et, q, Qx, Qy, Qz, cfar;
            if constexpr (C::LANE_AOS) {
                const double2 t45 = __ldg(reinterpret_cast<const double2*>(tp) + 2);
                et = t45.x;
                if (eu_max * et < kScreen) return false;  // primitives are sorted by E, descending
                const double2 t01 = __ldg(reinterpret_cast<const double2*>(tp));
                const double2 t23 = __ldg(reinterpret_cast<const double2*>(tp) + 1);
                if (has_next) {
#pragma unroll
                    for (int b = 0; b < FT * 8; b += 128)
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(tp + FT) + b));
                }
                q = t01.x; Qx = t01.y; Qy = t23.x; Qz = t23.y;
                cfar = kHalfSqrtPi * t45.y;  // sqrt(pi)/2 / sqrt(q)
            } else {
                et = tp[4 * ld];
                if (eu_max * et < kScreen) return false;  // primitives are sorted by E, descending
                q = tp[0];
                Qx = tp[ld]; Qy = tp[2 * ld]; Qz = tp[3 * ld];
                cfar = kHalfSqrtPi * tp[5 * ld];  // sqrt(pi)/2 / sqrt(q)
            }
            double K[NFU][NHT];
#pragma unroll
            for (int f = 0; f < NFU; ++f)
#pragma unroll
                for (int h = 0; h < NHT; ++h) K[f][h] = 0.0;

            for (int ku = 0; ku < npu; ++ku) {
                const double* up = s_u + ku * FU;
                if (up[4] * et < kScreen) break;  // IF (EGH*EIJ .LT. 1.0D-14) CYCLE  (int2e.f90:257)
                any = true;
                ++npq;
                const double p = up[0];
                const double X = up[1] - Qx, Y = up[2] - Qy, Z = up[3] - Qz;
                const double R2 = fma(X, X, fma(Y, Y, Z * Z));
                const double s = p + q;
                const double pq = p * q;
                const double w = pq * R2;  // T*(p+q)
                double G[LT + 1];
                if (w >= (double)(2 * Q + 36) * s) {
                    // Boys3 (auxilary.f90:194-215): G_j = sqrt(pi)/2 /sqrt(pq) (2j-1)!! (-1)^j R^-(2j+1)
                    const double rinv = rsqrt_pos(R2);
                    const double m = -(rinv * rinv);
                    double g = cfar * up[5] * rinv;
                    G[0] = g;
#pragma unroll
                    for (int j = 1; j <= LT; ++j) {
                        g *= (double)(2 * j - 1) * m;
                        G[j] = g;
                    }
                } else {
                    const double rs = rsqrt_pos(s);
                    const double alpha = pq * (rs * rs);
                    boys_near_mid<Q, LT>(alpha * R2, alpha, rs, G, s_ft, s_exp);
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
                        const double cu = up[kRecCoef + k];
                        static_for<0, NHT>([&](auto hc) {
                            constexpr int hp = decltype(hc)::value;
                            constexpr int ri = h_add(hk, hp);
                            K[lf][hp] = fma(cu, R[ri], K[lf][hp]);
                        });
                    }
                });
            }
            // step B: out[f][f'] += (-1)^{|H'|} D'_k' K[f][H'_k'], the terms of one f' summed in registers first
            static_for<0, NFT>([&](auto fc) {
                constexpr int fp = decltype(fc)::value;
                double acc[NFU];
#pragma unroll
                for (int f = 0; f < NFU; ++f) acc[f] = 0.0;
                static_for<0, NTT>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    if constexpr (term_fn(TT, k) == fp) {
                        constexpr int hp = term_h(TT, k);
                        double ct = C::LANE_AOS ? __ldg(tp + kRecCoef + k) : tp[(size_t)(kRecCoef + k) * ld];
                        if constexpr (h_parity(hp) != 0) ct = -ct;
#pragma unroll
                        for (int f = 0; f < NFU; ++f) acc[f] = fma(ct, K[f][hp], acc[f]);
                    }
                });
                static_for<0, NFU>([&](auto ffc) { sink(ffc, fc, acc[decltype(ffc)::value]); });
            });
            return true;
        };

        for (int ib = 0; ib < nitem; ib += 32) {
            if (ib + lane < nitem) {
                const int v = s_vidx[ib + lane];
                double out_r[OUT_SMEM ? 1 : NOUT];
                if constexpr (OUT_SMEM) {
#pragma unroll
                    for (int o = 0; o < NOUT; ++o) s_out[o * NTHREADS + tid] = 0.0;
                } else {
#pragma unroll
                    for (int o = 0; o < NOUT; ++o) out_r[o] = 0.0;
                }
                any = false;
                // Schwarz skip: every integral of the quartet is below tau in magnitude (|(ij|kl)| <= sqrt((ij|ij)(kl|kl)));
                // the slice was zero filled, so the quartet is simply left out
                const bool skip = (a.tau > 0.0) && (qu * __ldg(S.t_q + v) < a.tau);
                const int npt = skip ? 0 : S.t_nprim[v];
                // Lane-side records.  Small classes read the structure-of-arrays copy (a task's pairs are
                // a contiguous range, so the 8 chunks of a task share its lines in L1).  The classes with
                // many Hermite coefficients read the lane's own contiguous [9][FT] block instead: one base
                // pointer and compile-time offsets instead of a 64-bit multiply-add per field, and the next
                // primitive's lines are prefetched into L1 while this one is contracted.
                const double* trec = C::LANE_AOS ? S.t_aos + (size_t)v * 9 * FT : S.t_soa + v;
                for (int kt = 0; kt < npt; ++kt) {
                    const double* tp = C::LANE_AOS ? trec + kt * FT : trec + (size_t)kt * FT * ld;
                    const bool live = contract_prim(tp, kt + 1 < npt, [&](auto fc, auto fpc, double val) {
                        constexpr int o = decltype(fc)::value * NFT + decltype(fpc)::value;
                        if constexpr (OUT_SMEM) s_out[o * NTHREADS + tid] += val;
                        else out_r[o] += val;
                    });
                    if (!live) break;
                }
                if (a.stage != nullptr) {
                    // compose mode: the quartet's block [f_u][f_v] of the staging array, written whole with 128-bit
                    // stores (every sector it touches is covered).  Quartets the screen or the Schwarz skip leave out
                    // write nothing: their blocks keep the zeros of plan creation -- except the one-double blocks of
                    // (S S|S S), whose sectors are shared by four quartets and are therefore always written.
                    if (any || NOUT == 1) {
                        double* dst = a.stage + (__ldg(S.stage_row + u) + (int64_t)v * (C::NFU_FULL * NFT) + ((USL >= 0) ? USL * 4 * NFT : 0));
                        if constexpr (NOUT == 1) {
                            dst[0] = out_r[0];
                        } else {
#pragma unroll
                            for (int o = 0; o < NOUT; o += 2) {
                                double2 w;
                                if constexpr (OUT_SMEM) w = make_double2(s_out[o * NTHREADS + tid], s_out[(o + 1) * NTHREADS + tid]);
                                else w = make_double2(out_r[o], out_r[o + 1]);
                                reinterpret_cast<double2*>(dst)[o >> 1] = w;
                            }
                        }
                    }
                } else
                // scatter mode: store the distinct canonical integrals of this shell quartet (the slice was zero
                // filled: quartets the screen removes entirely keep the reference's exact zeros)
                if (any) {
                    const bool same_pair = S.tri && (v == u);
                    const int64_t np = a.npair;
                    int P2[NFT];
#pragma unroll
                    for (int fp = 0; fp < NFT; ++fp) P2[fp] = S.t_pidx[(size_t)v * NFT + fp];
#pragma unroll
                    for (int f = 0; f < NFU; ++f) {
                        if (P1[f] < 0) continue;
#pragma unroll
                        for (int fp = 0; fp < NFT; ++fp) {
                            if (P2[fp] < 0) continue;
                            if (same_pair && P1[f] > P2[fp]) continue;
                            const int64_t lo = P1[f] < P2[fp] ? P1[f] : P2[fp];
                            const int64_t hi = P1[f] < P2[fp] ? P2[fp] : P1[f];
                            const int64_t idx = lo * np - ((lo * (lo - 1)) >> 1) + (hi - lo) - a.out_offset;
                            store_eri(a.out + idx, OUT_SMEM ? s_out[(f * NFT + fp) * NTHREADS + tid] : out_r[f * NFT + fp]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) npq += __shfl_xor_sync(0xffffffffu, npq, o);
        if (lane == 0 && npq) atomicAdd(a.pq_counter, (unsigned long long)npq);
        __syncwarp();  // every lane is done with this buffer before the TMA two tasks ahead reuses it
        t = tn;
        task = taskn;
        buf ^= 1;
    }
}

// ------------------------------------------------------------------------------------------
// Warp-cooperative (SP SP|SP SP) kernel (MYQC_PP_KERNEL=warp).  The class kernel above gives a whole contracted
// quartet -- up to 81 primitive quartets x four mu-slices -- to ONE lane, so a launch with a few thousand quartets
// (C20H42, (H2O)_16) lasts as long as its longest lane (~0.3 ms) however empty the machine is.  Here a warp takes ONE
// contracted quartet at a time and its lanes take the PRIMITIVE quartets (row primitive ku, lane-side primitive kt)
// that pass the reference's test, 32 at a time:
//   * both pair records sit in shared memory (row: TMA bulk copy two tasks ahead as above; lane side: one coalesced
//     copy per quartet); a lane reads the coefficients of its own (ku, kt);
//   * per lane: Boys values and R_{NLM} (35 values) once, then for each mu-slice step A (K[4][10]) and step B -- the
//     full 16 x 16 contribution of the primitive quartet, 32 values at a time;
//   * the 32 values of every lane go through a [32][33] shared-memory tile and lane l adds up column l (fixed order:
//     primitive quartets in (kt, ku) order), so after eight tiles every lane holds 8 of the quartet's 256 integrals;
//   * stores: lane l owns f' = l % 8 (+8) and f = 4 mu + l / 8, so the 16 f' of one f -- four runs of four adjacent
//     packed columns -- leave in one warp instruction.
// Same screens, same Boys / R arithmetic, same term tables as the class kernel; only the order in which the
// primitive quartets of one integral are added differs (|difference| ~ 1e-16 relative).
// UT = 2: (SP SP|SP SP), four mu-slices of the row pair; UT = 1: (S SP|SP SP), one slice.  The lane side is SP.SP.
template <int UT>
struct PPW {
    static constexpr int LT = UT + 2, Q = 3 * LT, NR = h_count(LT), NHT = tt_nh(2), NT = tt_nterm(2), FU = tt_nfield(2);
    static constexpr int NTU = tt_nterm(UT), NFU = tt_nf(UT), NSL = NFU / 4, FUU = tt_nfield(UT);
    static constexpr int NTHREADS = 128, NWARPS = 4;
    static constexpr uint32_t U_BYTES = 9 * FUU * 8;
    // Primitive records sit LD doubles apart in shared memory (52 in global memory): lanes read the same field of
    // different primitives, and with a stride of 54 doubles primitives 0..7 fall into different banks (52: 0, 4, 8 collide)
    static constexpr int LD = FU + 2;
    static constexpr int LDU = (FUU % 8 == 4) ? FUU + 2 : FUU;  // 54 for SP.SP rows; S.SP rows (14 doubles) need no padding
    static constexpr int RED_LD = 33;
    static constexpr size_t OFF_EXP = 121 * 8 * 8;
    static constexpr size_t OFF_U = OFF_EXP + 608 * 16;                          // per warp: two row buffers
    static constexpr size_t OFF_V = OFF_U + (size_t)NWARPS * 2 * 9 * LDU * 8;    // per warp: the lane-side record
    static constexpr size_t OFF_RED = OFF_V + (size_t)NWARPS * 9 * LD * 8;       // per warp: [32][33] reduction tile
    static constexpr size_t OFF_BAR = OFF_RED + (size_t)NWARPS * 32 * RED_LD * 8;
    static constexpr size_t SMEM = OFF_BAR + (size_t)NWARPS * 2 * 8;
};

template <int UT, bool MULTI>
__global__ void __launch_bounds__(PPW<UT>::NTHREADS, 2) eri_ppw_kernel(const __grid_constant__ ClassArgs a) {
    using C = PPW<UT>;
    constexpr int LT = C::LT, Q = C::Q, NR = C::NR, NHT = C::NHT, NT = C::NT, FU = C::FU, LD = C::LD;
    constexpr int NTU = C::NTU, NFU = C::NFU, NSL = C::NSL, FUU = C::FUU, LDU = C::LDU;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* s_ft = reinterpret_cast<double*>(smem_raw);
    double2* s_exp = reinterpret_cast<double2*>(smem_raw + C::OFF_EXP);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    double* s_ubuf = reinterpret_cast<double*>(smem_raw + C::OFF_U) + (size_t)warp * 2 * 9 * LDU;
    double* s_v = reinterpret_cast<double*>(smem_raw + C::OFF_V) + (size_t)warp * 9 * LD;
    double* s_red = reinterpret_cast<double*>(smem_raw + C::OFF_RED) + (size_t)warp * 32 * C::RED_LD;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw + C::OFF_BAR) + warp * 2;

    for (int i = tid; i < 121 * 8; i += C::NTHREADS) s_ft[i] = a.ftab_q[i];
    for (int i = tid; i < 601; i += C::NTHREADS) s_exp[i] = a.exptab[i];
    if (lane == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    int t = 0;
    int4 task = make_int4(0, 0, 0, 0);
    if (lane == 0) {
        t = atomicAdd(a.row_counter, 1);
        if (t < a.ntasks) {
            task = a.tasks[t];
            mbar_expect_tx(&s_bar[0], C::U_BYTES);
            const double* src = a.part[MULTI ? task.w : 0].u_aos + (size_t)task.x * 9 * FUU;
            if constexpr (LDU == FUU) tma_bulk_g2s(s_ubuf, src, C::U_BYTES, &s_bar[0]);
            else {
#pragma unroll
                for (int k = 0; k < 9; ++k) tma_bulk_g2s(s_ubuf + k * LDU, src + k * FUU, FUU * 8, &s_bar[0]);
            }
        }
    }
    t = __shfl_sync(0xffffffffu, t, 0);
    task.x = __shfl_sync(0xffffffffu, task.x, 0);
    task.y = __shfl_sync(0xffffffffu, task.y, 0);
    task.z = __shfl_sync(0xffffffffu, task.z, 0);
    if (MULTI) task.w = __shfl_sync(0xffffffffu, task.w, 0);
    int buf = 0;
    uint32_t parity0 = 0, parity1 = 0;
    // this lane's 2 NSL integrals of a quartet: slot 2*mu + hh is (f, f') = (4 mu + lane / 8, 8 hh + lane % 8)
    const int fl = lane >> 3, fpl = lane & 7;
    while (t < a.ntasks) {
        int tn = 0;
        int4 taskn = make_int4(0, 0, 0, 0);
        if (lane == 0) {
            tn = atomicAdd(a.row_counter, 1);
            if (tn < a.ntasks) {
                taskn = a.tasks[tn];
                mbar_expect_tx(&s_bar[buf ^ 1], C::U_BYTES);
                const double* src = a.part[MULTI ? taskn.w : 0].u_aos + (size_t)taskn.x * 9 * FUU;
                double* dstb = s_ubuf + (size_t)(buf ^ 1) * 9 * LDU;
                if constexpr (LDU == FUU) tma_bulk_g2s(dstb, src, C::U_BYTES, &s_bar[buf ^ 1]);
                else {
#pragma unroll
                    for (int k = 0; k < 9; ++k) tma_bulk_g2s(dstb + k * LDU, src + k * FUU, FUU * 8, &s_bar[buf ^ 1]);
                }
            }
        }
        tn = __shfl_sync(0xffffffffu, tn, 0);
        taskn.x = __shfl_sync(0xffffffffu, taskn.x, 0);
        taskn.y = __shfl_sync(0xffffffffu, taskn.y, 0);
        taskn.z = __shfl_sync(0xffffffffu, taskn.z, 0);
        if (MULTI) taskn.w = __shfl_sync(0xffffffffu, taskn.w, 0);
        const PartArgs& S = a.part[MULTI ? task.w : 0];
        const int u = task.x;
        const int npu = S.u_nprim[u];
        int P1[NSL];
#pragma unroll
        for (int m = 0; m < NSL; ++m) P1[m] = S.u_pidx[(size_t)u * NFU + 4 * m + fl];
        if (buf == 0) { mbar_wait(&s_bar[0], parity0); parity0 ^= 1; }
        else          { mbar_wait(&s_bar[1], parity1); parity1 ^= 1; }
        const double* s_u = s_ubuf + (size_t)buf * 9 * LDU;
        const double eu_max = s_u[4];
        const double qu = (a.tau > 0.0) ? __ldg(S.u_q + u) : 0.0;
        unsigned long long npq_task = 0;

        for (int v = task.y; v < task.z; ++v) {
            // Schwarz skip (warp uniform): every integral of the quartet is below tau
            if (a.tau > 0.0 && qu * __ldg(S.t_q + v) < a.tau) continue;
            const int npt = S.t_nprim[v];
            __syncwarp();  // the previous quartet's readers of s_v are done
            {
                const double* src = S.t_aos + (size_t)v * 9 * FU;
                for (int i = lane; i < npt * FU; i += 32) s_v[(i / FU) * LD + i % FU] = __ldg(src + i);
            }
            __syncwarp();
            // lane kt < npt: number of row primitives that pass IF (EGH*EIJ .LT. 1.0D-14) CYCLE (int2e.f90:257) against
            // lane-side primitive kt (both lists are sorted by prefactor, so the survivors are a prefix)
            int nk = 0;
            if (lane < npt) {
                const double et = s_v[lane * LD + 4];
                if (!(eu_max * et < kScreen))
                    for (int ku = 0; ku < npu; ++ku) {
                        if (s_u[ku * LDU + 4] * et < kScreen) break;
                        ++nk;
                    }
            }
            int incl = nk;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            const int npq = __shfl_sync(0xffffffffu, incl, 15);  // at most nine lanes hold a count
            if (npq == 0) continue;
            const int excl = incl - nk;
            npq_task += (unsigned long long)npq;
            double out[2 * NSL];
#pragma unroll
            for (int o = 0; o < 2 * NSL; ++o) out[o] = 0.0;

            for (int base = 0; base < npq; base += 32) {
                const int i = base + lane;
                const bool active = i < npq;
                int kt = 0, ku = 0;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const int ek = __shfl_sync(0xffffffffu, excl, k), ik = __shfl_sync(0xffffffffu, incl, k);
                    if (i >= ek && i < ik) { kt = k; ku = i - ek; }
                }
                const int nact = min(32, npq - base);
                const int nsum = (nact + 7) & ~7;  // lanes without a primitive quartet write zeros
                const double* up = s_u + ku * LDU;
                const double* tp = s_v + kt * LD;
                double R[NR];
                if (active) {
                    const double p = up[0], q = tp[0];
                    const double X = up[1] - tp[1], Y = up[2] - tp[2], Z = up[3] - tp[3];
                    const double R2 = fma(X, X, fma(Y, Y, Z * Z));
                    const double sm = p + q;
                    const double pq = p * q;
                    const double w = pq * R2;  // T*(p+q)
                    double G[LT + 1];
                    if (w >= (double)(2 * Q + 36) * sm) {
                        const double rinv = rsqrt_pos(R2);
                        const double m = -(rinv * rinv);
                        double g = (kHalfSqrtPi * tp[5]) * up[5] * rinv;
                        G[0] = g;
#pragma unroll
                        for (int j = 1; j <= LT; ++j) {
                            g *= (double)(2 * j - 1) * m;
                            G[j] = g;
                        }
                    } else {
                        const double rs = rsqrt_pos(sm);
                        const double alpha = pq * (rs * rs);
                        boys_near_mid<Q, LT>(alpha * R2, alpha, rs, G, s_ft, s_exp);
                    }
                    build_R<LT>(G, X, Y, Z, R);
                } else {
#pragma unroll
                    for (int r = 0; r < NR; ++r) R[r] = 0.0;
                }
                // The four mu-slices run through ONE copy of the step-B / reduction code (a run-time loop; only step A,
                // whose R indices differ from slice to slice, exists four times): the first version unrolled everything
                // into 100 KB of straight-line code and spent 37 % of its stall samples waiting for instructions.
                // Lanes without a primitive quartet carry R = 0 and so produce exact zeros without a branch.
#pragma unroll 1
                for (int mu = 0; mu < NSL; ++mu) {
                    // step A: K[f][H'] = sum_k D_k R[H_k + H'] over the terms of the four function pairs of this slice
                    double K[4][NHT];
                    auto step_a = [&](auto mc) {
                        constexpr int m = decltype(mc)::value;
                        static_for<0, NTU>([&](auto kc) {
                            constexpr int k = decltype(kc)::value;
                            constexpr int f = term_fn(UT, k);
                            if constexpr (f / 4 == m) {
                                constexpr int hk = term_h(UT, k);
                                constexpr bool fst = (k == 0) || (term_fn(UT, k > 0 ? k - 1 : 0) != f);  // first term of its function pair
                                const double cu = up[kRecCoef + k];
                                static_for<0, NHT>([&](auto hc) {
                                    constexpr int hp = decltype(hc)::value;
                                    if constexpr (fst) K[f % 4][hp] = cu * R[h_add(hk, hp)];
                                    else K[f % 4][hp] = fma(cu, R[h_add(hk, hp)], K[f % 4][hp]);
                                });
                            }
                        });
                    };
                    if constexpr (NSL == 1) step_a(std::integral_constant<int, 0>{});
                    else
                        switch (mu) {
                            case 0: step_a(std::integral_constant<int, 0>{}); break;
                            case 1: step_a(std::integral_constant<int, 1>{}); break;
                            case 2: step_a(std::integral_constant<int, 2>{}); break;
                            default: step_a(std::integral_constant<int, 3>{}); break;
                        }
                    // step B, eight f' at a time: this primitive quartet's share of 32 integrals into the tile
                    static_for<0, 2>([&](auto hhc) {
                        constexpr int hh = decltype(hhc)::value;
                        static_for<0, 8>([&](auto fc) {
                            constexpr int fp = 8 * hh + decltype(fc)::value;
                            double acc[4] = {0.0, 0.0, 0.0, 0.0};
                            static_for<0, NT>([&](auto kc) {
                                constexpr int k = decltype(kc)::value;
                                if constexpr (term_fn(2, k) == fp) {
                                    constexpr int hp = term_h(2, k);
                                    double ct = tp[kRecCoef + k];
                                    if constexpr (h_parity(hp) != 0) ct = -ct;
#pragma unroll
                                    for (int f = 0; f < 4; ++f) acc[f] = fma(ct, K[f][hp], acc[f]);
                                }
                            });
#pragma unroll
                            for (int f = 0; f < 4; ++f) s_red[(f * 8 + decltype(fc)::value) * C::RED_LD + lane] = acc[f];
                        });
                        __syncwarp();
                        // lane l adds up column l = (f, f') = (l / 8, l % 8) over the primitive quartets of the batch
                        {
                            const double* col = s_red + lane * C::RED_LD;
                            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 1
                            for (int r = 0; r < nsum; r += 8) {
                                const double c0 = col[r], c1 = col[r + 1], c2 = col[r + 2], c3 = col[r + 3];
                                const double c4 = col[r + 4], c5 = col[r + 5], c6 = col[r + 6], c7 = col[r + 7];
                                s0 += c0 + c4;
                                s1 += c1 + c5;
                                s2 += c2 + c6;
                                s3 += c3 + c7;
                            }
                            const double sum = (s0 + s1) + (s2 + s3);
#pragma unroll
                            for (int m = 0; m < NSL; ++m)
                                if (mu == m) out[2 * m + hh] += sum;
                        }
                        __syncwarp();
                    });
                }
            }

            if (a.stage != nullptr) {
                // compose mode: this quartet's dense block [f][f'] of the staging array
                double* dst = a.stage + (__ldg(S.stage_row + u) + (int64_t)v * (NFU * 16));
#pragma unroll
                for (int m = 0; m < NSL; ++m)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) dst[(4 * m + fl) * 16 + 8 * hh + fpl] = out[2 * m + hh];
            } else {
                const bool same_pair = S.tri && (v == u);
                const int64_t np = a.npair;
                const int P2a = S.t_pidx[(size_t)v * 16 + fpl], P2b = S.t_pidx[(size_t)v * 16 + 8 + fpl];
#pragma unroll
                for (int m = 0; m < NSL; ++m) {
                    if (P1[m] < 0) continue;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int P2 = hh ? P2b : P2a;
                        if (P2 < 0) continue;
                        if (same_pair && P1[m] > P2) continue;
                        const int64_t lo = P1[m] < P2 ? P1[m] : P2;
                        const int64_t hi = P1[m] < P2 ? P2 : P1[m];
                        const int64_t idx = lo * np - ((lo * (lo - 1)) >> 1) + (hi - lo) - a.out_offset;
                        store_eri(a.out + idx, out[2 * m + hh]);
                    }
                }
            }
        }
        if (lane == 0 && npq_task) atomicAdd(a.pq_counter, npq_task);
        __syncwarp();  // every lane is done with this row buffer before the TMA two tasks ahead reuses it
        t = tn;
        task = taskn;
        buf ^= 1;
    }
}

// ------------------------------------------------------------------------------------------
// Schwarz factors.  (f|f) for every function pair f of the shell pairs of kind T, one thread per shell pair, over
// ALL primitive quartets of the pair with itself: no EIJ*EGH screen here, because the reference's own diagonal is
// exactly zero once E < 1e-7 and a bound built from it would not bound anything.  Same Boys / R arithmetic as the
// class kernels; only the terms of one function pair against themselves are contracted.
template <int T>
__global__ void __launch_bounds__(128) eri_diag_kernel(const double* __restrict__ aos, const int32_t* __restrict__ nprim,
                                                       const int32_t* __restrict__ pidx, int n, const double* __restrict__ ftab_q,
                                                       const double2* __restrict__ exptab, double* __restrict__ diag) {
    constexpr int LT = 2 * T, Q = 3 * LT, NR = h_count(LT), NF = tt_nf(T), NT = tt_nterm(T), FT = tt_nfield(T);
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const double* rec = aos + (size_t)u * 9 * FT;
    const int np = nprim[u];
    double acc[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] = 0.0;
    for (int k1 = 0; k1 < np; ++k1) {
        const double* r1 = rec + k1 * FT;
        for (int k2 = 0; k2 < np; ++k2) {
            const double* r2 = rec + k2 * FT;
            const double p = r1[0], q = r2[0];
            const double X = r1[1] - r2[1], Y = r1[2] - r2[2], Z = r1[3] - r2[3];
            const double R2 = fma(X, X, fma(Y, Y, Z * Z));
            const double s = p + q, pq = p * q, w = pq * R2;
            double G[LT + 1];
            if (w >= (double)(2 * Q + 36) * s) {
                const double rinv = rsqrt_pos(R2);
                const double m = -(rinv * rinv);
                double g = kHalfSqrtPi * r2[5] * r1[5] * rinv;
                G[0] = g;
#pragma unroll
                for (int j = 1; j <= LT; ++j) { g *= (double)(2 * j - 1) * m; G[j] = g; }
            } else {
                const double rs = rsqrt_pos(s);
                const double alpha = pq * (rs * rs);
                boys_near_mid<Q, LT>(alpha * R2, alpha, rs, G, ftab_q, exptab);
            }
            double R[NR];
            build_R<LT>(G, X, Y, Z, R);
            static_for<0, NT>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                static_for<0, NT>([&](auto kc2) {
                    constexpr int kk = decltype(kc2)::value;
                    if constexpr (term_fn(T, k) == term_fn(T, kk)) {
                        constexpr int ri = h_add(term_h(T, k), term_h(T, kk));
                        const double c = r1[kRecCoef + k] * r2[kRecCoef + kk];
                        if constexpr (h_parity(term_h(T, kk)) != 0) acc[term_fn(T, k)] = fma(-c, R[ri], acc[term_fn(T, k)]);
                        else acc[term_fn(T, k)] = fma(c, R[ri], acc[term_fn(T, k)]);
                    }
                });
            });
        }
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        const int P = pidx[(size_t)u * NF + f];
        if (P >= 0) diag[P] = acc[f];
    }
}

int launch_diag(int T, const double* aos, const int32_t* nprim, const int32_t* pidx, int n, const double* ftab_q,
                const double2* exptab, double* diag, void* stream) {
    if (n <= 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = (n + 127) / 128;
    if (T == 0) eri_diag_kernel<0><<<grid, 128, 0, st>>>(aos, nprim, pidx, n, ftab_q, exptab, diag);
    else if (T == 1) eri_diag_kernel<1><<<grid, 128, 0, st>>>(aos, nprim, pidx, n, ftab_q, exptab, diag);
    else eri_diag_kernel<2><<<grid, 128, 0, st>>>(aos, nprim, pidx, n, ftab_q, exptab, diag);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
__global__ void fill_zero_kernel(double* __restrict__ out, int64_t n, int* __restrict__ counters, int ncounters) {
    // 128-bit stores, grid-stride; a slice may start on an odd element.  Also resets the per-launch
    // row counters of the class kernels that follow on the same stream.
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < ncounters; c += blockDim.x) counters[c] = 0;
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

// ------------------------------------------------------------------------------------------
// Screened zero fill.  An element (P,P') of the packed array is written by a class kernel iff its
// two shell pairs pass the reference's test on their largest prefactors,
// fl(emax_P * emax_P') >= 1e-14 (int2e.f90:257); the product is monotone in each factor, so with
// rk[] = rank of emax in the descending list of all shell-pair prefactors and cut[P] = number of
// list entries whose product with emax_P passes, the test is the integer compare rk[P'] < cut[P].
// This kernel writes the zeros of exactly the other elements: every element of the slice is
// written once, by one kernel, and the fill needs no ordering against the FP64 kernels -- it runs
// next to them (HBM-write bound next to DFMA bound) instead of in front of them.
//
// Work unit = kFillRows packed rows x kFillCols columns; thread t owns columns cb*kFillCols +
// j*256 + t and keeps their ranks in registers for all rows of the unit, so rk[] is read once per
// unit, and a warp writes 256 contiguous bytes per store.  Units are enumerated column block by
// column block (ucb[] = prefix sums of the row blocks each column block needs: only rows <= the
// block's last column exist in the upper triangle) and pulled from a global counter.
constexpr int kFillThreads = 256;
constexpr int kFillColsPerThread = kFillCols / kFillThreads;
constexpr int kFillUcbSmem = 1024;  // column-block prefix entries cached in shared memory

__global__ void __launch_bounds__(kFillThreads) fill_screened_kernel(const FillArgs a) {
    __shared__ int s_unit[2];
    __shared__ int s_ucb[kFillUcbSmem];
    const int tid = threadIdx.x;
    const int64_t np = a.npair;
    const bool ucb_smem = a.ncb + 1 <= kFillUcbSmem;
    if (ucb_smem)
        for (int i = tid; i <= a.ncb; i += kFillThreads) s_ucb[i] = a.ucb[i];
    // thread 0 claims units; with pacing it first waits until the class kernels have got far enough
    bool pace = a.nprog > 0;  // thread 0 only
    auto claim = [&]() -> int {
        if (pace) {
            unsigned long long t_wait0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_wait0));
            for (;;) {
                float p = 0.0f;
                for (int k = 0; k < a.nprog; ++k) {
                    const int n = a.prog_n[k];
                    int c = *reinterpret_cast<const volatile int*>(a.counters + a.prog_idx[k]);
                    c = c < n ? c : n;
                    p += a.prog_w[k] * (float)c / (float)n;
                }
                // a little ahead of the compute (tasks are sorted heaviest first, so the task count
                // lags the time), everything once the class kernels are done (p == 1 -> 1.2)
                const float allowed = (0.04f + 1.16f * p) * (float)a.nunits;
                const int cur = *reinterpret_cast<const volatile int*>(a.counter);
                if ((float)cur < allowed || p >= 0.999f) break;
                // safety net: never wait more than 2 ms on the class kernels (they advance every few
                // microseconds when they run; if they cannot run, throttling must not become a deadlock)
                unsigned long long t_now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
                if (t_now - t_wait0 > 2000000ull) { pace = false; break; }
                __nanosleep(2000);
            }
        }
        return atomicAdd(a.counter, 1);
    };
    if (tid == 0) s_unit[0] = claim();
    __syncthreads();
    int unit = s_unit[0];
    for (int it = 0; unit < a.nunits; ++it) {
        // the next unit is claimed while this one is written (one barrier per unit)
        if (tid == 0) s_unit[(it + 1) & 1] = claim();
        // column block of this unit: last cb with ucb[cb] <= unit
        int lo = 0, hi = a.ncb;
        if (ucb_smem) {
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_ucb[mid] <= unit) lo = mid; else hi = mid;
            }
        } else {
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(a.ucb + mid) <= unit) lo = mid; else hi = mid;
            }
        }
        const int ulo = ucb_smem ? s_ucb[lo] : __ldg(a.ucb + lo);
        const int64_t r0 = a.row_lo + (int64_t)(unit - ulo) * kFillRows;
        const int64_t c0 = (int64_t)(a.cb0 + lo) * kFillCols;
        int rk[kFillColsPerThread];
#pragma unroll
        for (int j = 0; j < kFillColsPerThread; ++j) {
            const int64_t c = c0 + j * kFillThreads + tid;
            rk[j] = c < np ? __ldg(a.rk + c) : -1;  // -1: column does not exist, never stored
        }
        int64_t rend = r0 + kFillRows;
        if (rend > a.row_hi) rend = a.row_hi;
        const bool interior = (r0 + kFillRows - 1 <= c0);  // every row of the unit is <= every column
        // element (r,c) lives at out[off(r) - out_offset + (c - r)], off(r) = r*np - r(r-1)/2
        double* o = a.out + (r0 * np - ((r0 * (r0 - 1)) >> 1) - a.out_offset - r0 + c0 + tid);
        int64_t r = r0;
        if (interior) {
            for (; r + 4 <= rend; r += 4) {
                int cut[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) cut[i] = a.all ? 0 : __ldg(a.cut + r + i);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#pragma unroll
                    for (int j = 0; j < kFillColsPerThread; ++j)
                        if (rk[j] >= cut[i]) __stcs(o + j * kFillThreads, 0.0);
                    o += np - (r + i) - 1;  // off(r+1) - (r+1) - (off(r) - r)
                }
                if (a.sleep_ns > 0) __nanosleep(a.sleep_ns);
            }
        }
        for (; r < rend; ++r) {
            const int cut = a.all ? 0 : __ldg(a.cut + r);
#pragma unroll
            for (int j = 0; j < kFillColsPerThread; ++j)
                if (rk[j] >= cut && c0 + j * kFillThreads + tid >= r) __stcs(o + j * kFillThreads, 0.0);
            o += np - r - 1;
        }
        __syncthreads();
        unit = s_unit[(it + 1) & 1];
    }
}

// ------------------------------------------------------------------------------------------
// Compose pass.  The class kernels leave every evaluated shell quartet (u|v) as a dense block [f_u][f_v] in the
// staging array (u = uniform side of its launch).  This kernel writes the packed slice once, row by row, with
// 256-byte warp stores: element (P,P') is an exact zero unless the shell pairs U of P and V of P' pass the
// reference's rule on their largest prefactors (fl(emax_U*emax_V) >= 1e-14, int2e.f90:257, as the integer compare
// rk[P'] < cut[P]) and the Schwarz test Q_U*Q_V >= tau; otherwise it is read from the quartet's block: the uniform
// side of the quartet is the pair of the lower type, or for equal types the pair that comes first in its list
// ("mine" before "later"), exactly the rule the launches were built with.
//
// Work unit = kCompRows packed rows x kCompCols columns; a thread owns kCompColsPerThread columns (stride
// kCompThreads, so a warp stores 256 contiguous bytes) and keeps their pair data in registers for all rows of the
// unit; the rows' pair data sit in shared memory.  Units are enumerated row block by row block (urb[] = prefix
// sums of the column blocks of each row block: only columns >= the block's first row exist) and pulled from a
// global counter, so that concurrently running CTAs work on neighbouring rows and columns and the 32-byte sectors
// of a quartet block that several rows share are fetched from DRAM once.
// Gathers go through cp.async (LDGSTS) into a per-thread landing area in shared memory: the packed array is ~90 %
// zeros, so the few loads a thread needs are spread over many rows; with register loads only a handful are in flight
// per warp while the stores wait on them (measured: 12 ms, long-scoreboard bound), with cp.async every needed
// element of kSub rows x 4 columns is in flight at once, two such groups deep, and costs no registers.  A thread only
// ever reads the slots it filled itself, so cp.async.wait_group is the only synchronisation.
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int kSub>
__global__ void __launch_bounds__(kCompThreads) compose_kernel(const ComposeArgs a) {
    static_assert(kSub * kCompColsPerThread <= 32, "one mask bit per landing slot");
    extern __shared__ __align__(16) double s_land[];  // [2][kSub][kCompCols]
    __shared__ int s_unit[2];
    __shared__ int s_cut[kCompRows], s_g[kCompRows], s_iun[kCompRows], s_pk[kCompRows];
    __shared__ double s_q[kCompRows];
    __shared__ uint32_t s_rel[kCompRows][8];       // rowrel of the unit's row pairs, by lane list
    __shared__ uint32_t s_crel[3][kCompCols];      // rowrel of the unit's column pairs against the own list of each type
    __shared__ long long s_lb[36];
    const int tid = threadIdx.x;
    const int64_t np = a.npair;
    const bool schwarz = a.tau > 0.0 && a.pq != nullptr;
    if (tid < 36) s_lb[tid] = a.launch_base[tid];
    if (tid == 0) s_unit[0] = atomicAdd(a.counter, 1);
    __syncthreads();
    int unit = s_unit[0];
    for (int it = 0; unit < a.nunits; ++it) {
        if (tid == 0) s_unit[(it + 1) & 1] = atomicAdd(a.counter, 1);  // the next unit is claimed while this one is written
        int lo = 0, hi = a.nrb;  // row block of this unit: last rb with urb[rb] <= unit
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(a.urb + mid) <= unit) lo = mid; else hi = mid;
        }
        const int64_t r0 = a.row_lo + (int64_t)lo * kCompRows;
        const int64_t c0 = ((r0 / kCompCols) + (unit - __ldg(a.urb + lo))) * kCompCols;
        int64_t rend = r0 + kCompRows;
        if (rend > a.row_hi) rend = a.row_hi;
        const int nrow = (int)(rend - r0);
        // rows of the unit -> shared memory.  pk = f | (2*type) << 4 | lid << 8
        if (tid < kCompRows) {
            int cut = 0, g = 0, iun = 0, pk = 0;
            double q = 0.0;
            const int info = tid < nrow ? __ldg(a.fpinfo + r0 + tid) : -1;
            if (info >= 0) {
                g = info >> 4;
                cut = a.all_zero ? 0 : __ldg(a.cut + r0 + tid);
                const int2 m = __ldg(a.pmeta + g);
                iun = m.x;
                pk = (info & 15) | ((m.y >> 1) << 5) | (m.y << 8);
                if (schwarz) q = __ldg(a.pq + g);
#pragma unroll
                for (int k = 0; k < 6; ++k) s_rel[tid][k] = __ldg(a.rowrel + (size_t)g * 6 + k);
            }
            s_cut[tid] = cut; s_g[tid] = g; s_iun[tid] = iun; s_pk[tid] = pk; s_q[tid] = q;
        }
        // columns of this thread -> registers (and their rowrel -> shared memory)
        int c_rk[kCompColsPerThread], c_g[kCompColsPerThread], c_ivn[kCompColsPerThread], c_pk[kCompColsPerThread];
        int c_imax[kCompColsPerThread];  // last row of the unit in which the column exists (column >= row)
        double c_q[kCompColsPerThread];
#pragma unroll
        for (int j = 0; j < kCompColsPerThread; ++j) {
            const int64_t c = c0 + j * kCompThreads + tid;
            c_rk[j] = INT32_MAX; c_g[j] = 0; c_ivn[j] = 0; c_pk[j] = 0; c_q[j] = 0.0;
            c_imax[j] = -1;
            if (c < np) {
                c_imax[j] = (c - r0 < nrow - 1) ? (int)(c - r0) : nrow - 1;
                const int info = __ldg(a.fpinfo + c);
                if (info >= 0) {
                    const int g = info >> 4;
                    const int2 m = __ldg(a.pmeta + g);
                    c_rk[j] = __ldg(a.rk + c);
                    c_g[j] = g;
                    c_ivn[j] = m.x;
                    c_pk[j] = (info & 15) | ((m.y >> 1) << 5) | (m.y << 8);
                    if (schwarz) c_q[j] = __ldg(a.pq + g);
#pragma unroll
                    for (int k = 0; k < 3; ++k) s_crel[k][j * kCompThreads + tid] = __ldg(a.rowrel + (size_t)g * 6 + 2 * k);
                }
            }
        }
        __syncthreads();
        // element (r,c) lives at out[off(r) - out_offset + (c - r)], off(r) = r*np - r(r-1)/2
        double* o = a.out + (r0 * np - ((r0 * (r0 - 1)) >> 1) - a.out_offset - r0 + c0 + tid);

        // issue the gathers of rows [i0, i0+kSub) into landing buffer sb; returns the mask of the slots in use
        auto issue = [&](int sb, int i0) -> uint32_t {
            uint32_t mask = 0;
            double* land = s_land + (size_t)sb * kSub * kCompCols + tid;
#pragma unroll
            for (int ii = 0; ii < kSub; ++ii) {
                const int i = i0 + ii;
                if (i < nrow) {
                    const int cut = s_cut[i];
#pragma unroll
                    for (int j = 0; j < kCompColsPerThread; ++j) {
                        if (c_rk[j] < cut) {  // skipped by the whole warp when none of its 32 columns passes the reference's rule
                            if (!(schwarz && s_q[i] * c_q[j] < a.tau)) {
                                const int pu = s_pk[i], pv = c_pk[j];
                                const int fu = pu & 15, tu2 = (pu >> 4) & 6, lidu = pu >> 8;
                                const int fv = pv & 15, tv2 = (pv >> 4) & 6, lidv = pv >> 8;
                                // the uniform side of the quartet is the pair with the smaller id (lower type; equal
                                // types: own list before later list, then list order): the rule the launches were built with
                                const bool u_uni = s_g[i] <= c_g[j];
                                const uint32_t w = u_uni ? s_rel[i][lidv] + (uint32_t)(fu << tv2) + ((uint32_t)c_ivn[j] << tu2) + (uint32_t)fv
                                                         : s_crel[tu2 >> 1][j * kCompThreads + tid] + (uint32_t)(fv << tu2) + ((uint32_t)s_iun[i] << tv2) + (uint32_t)fu;
                                const long long lb = s_lb[u_uni ? lidu * 6 + lidv : lidv * 6 + lidu];
                                if ((int)(lb >> 32) != INT32_MIN) {
                                    cp_async8(land + (ii * kCompCols + j * kCompThreads), a.stage + (lb + (long long)w));
                                    mask |= 1u << (ii * kCompColsPerThread + j);
                                }
                            }
                        }
                    }
                }
            }
            cp_async_commit();
            return mask;
        };
        const int nsub = (nrow + kSub - 1) / kSub;
        uint32_t m_cur = issue(0, 0);
        for (int sidx = 0; sidx < nsub; ++sidx) {
            uint32_t m_next = 0;
            if (sidx + 1 < nsub) { m_next = issue((sidx + 1) & 1, (sidx + 1) * kSub); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            const double* land = s_land + (size_t)(sidx & 1) * kSub * kCompCols + tid;
#pragma unroll
            for (int ii = 0; ii < kSub; ++ii) {
                const int i = sidx * kSub + ii;
#pragma unroll
                for (int j = 0; j < kCompColsPerThread; ++j) {
                    const double val = ((m_cur >> (ii * kCompColsPerThread + j)) & 1u) ? land[ii * kCompCols + j * kCompThreads] : 0.0;
                    if (i <= c_imax[j]) __stcs(o + j * kCompThreads, val);
                }
                o += np - (r0 + i) - 1;  // off(r+1) - (r+1) - (off(r) - r)
            }
            m_cur = m_next;
        }
        __syncthreads();
        unit = s_unit[(it + 1) & 1];
    }
}

int launch_compose(const ComposeArgs& a, int num_sms, void* stream) {
    if (a.nunits <= 0) return 0;
    static int ctas_per_sm = 0, sub = 0;
    if (ctas_per_sm == 0) {
        const char* e = getenv("MYQC_COMPOSE_CTAS");
        ctas_per_sm = e ? atoi(e) : 6;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        e = getenv("MYQC_COMPOSE_SUB");
        sub = e ? atoi(e) : 4;
    }
    int grid = num_sms * ctas_per_sm;
    if (grid > a.nunits) grid = a.nunits;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (sub == 4) {
        const size_t sm = 2ull * 4 * kCompCols * sizeof(double);
        compose_kernel<4><<<grid, kCompThreads, sm, st>>>(a);
    } else {
        const size_t sm = 2ull * 8 * kCompCols * sizeof(double);
        const int e0 = prepare_kernels();  // raises the dynamic shared-memory limit of compose_kernel<8> on this device
        if (e0) return e0;
        compose_kernel<8><<<grid, kCompThreads, sm, st>>>(a);
    }
    return (int)cudaGetLastError();
}

// dense XX(i,j,g,h) (column-major, i fastest) from the packed array: the fillsym pass of the
// reference (int2e.f90:290-304,540-554) done as a gather so that the 8n^4-byte stream is written
// once, coalesced.  [h0,h1) selects the slab XX(:,:,:,h0:h1-1) (one slab per device in the multi-GPU dense path).
__global__ void expand_dense_kernel(const double* __restrict__ packed, int norb, int h0, int h1, double* __restrict__ xx) {
    const int64_t n = norb;
    const int64_t npair = n * (n + 1) / 2;
    const int64_t total = n * n * n * (int64_t)(h1 - h0);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        int64_t r = e;
        const int64_t i = r % n; r /= n;
        const int64_t j = r % n; r /= n;
        const int64_t g = r % n; r /= n;
        const int64_t h = r + h0;
        const int64_t a = i < j ? i : j, b = i < j ? j : i;
        const int64_t c = g < h ? g : h, d = g < h ? h : g;
        const int64_t P1 = a * n - a * (a - 1) / 2 + (b - a);
        const int64_t P2 = c * n - c * (c - 1) / 2 + (d - c);
        const int64_t lo = P1 < P2 ? P1 : P2, hi = P1 < P2 ? P2 : P1;
        xx[e] = packed[lo * npair - lo * (lo - 1) / 2 + (hi - lo)];
    }
}

// ------------------------------------------------------------------------------------------
// Sparse transfer to the host.  The slice is cut into chunks of CH consecutive elements (CH = 32 ... 256, a multiple of
// the 32 elements = 256 bytes one warp store covers).  A warp walks groups of 256 elements (eight warp rows), so that
// eight independent loads are in flight per lane whatever the chunk size.
template <int CH>
__global__ void __launch_bounds__(256) chunk_flags_kernel(const double* __restrict__ out, int64_t n, unsigned char* __restrict__ flags) {
    constexpr int RPC = CH / 32;  // warp rows per chunk
    const int lane = threadIdx.x & 31;
    const int64_t ngroup = (n + 255) / 256;
    const int64_t nchunk = (n + CH - 1) / CH;
    const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < ngroup; g += nwarp) {
        const int64_t base = g * 256;
        long long bits[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t e = base + j * 32 + lane;
            bits[j] = e < n ? __double_as_longlong(__ldcs(out + e)) : 0ll;
        }
#pragma unroll
        for (int c = 0; c < 8 / RPC; ++c) {
            long long b = 0;
#pragma unroll
            for (int j = 0; j < RPC; ++j) b |= bits[c * RPC + j];
            const bool any = __any_sync(0xffffffffu, b != 0);
            const int64_t ci = g * (8 / RPC) + c;
            if (lane == 0 && ci < nchunk) flags[ci] = any ? 1 : 0;
        }
    }
}

template <int CH>
__global__ void __launch_bounds__(256) chunk_push_kernel(const double* __restrict__ out, int64_t n, const unsigned char* __restrict__ flags,
                                                         double* __restrict__ host) {
    constexpr int RPC = CH / 32;
    const int lane = threadIdx.x & 31;
    const int64_t ngroup = (n + 255) / 256;
    const int64_t nchunk = (n + CH - 1) / CH;
    const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < ngroup; g += nwarp) {
        const int64_t base = g * 256;
        // the group's flags: one byte per chunk, read by the first lanes and broadcast as a bit mask
        const int64_t c0 = g * (8 / RPC);
        const bool f = lane < 8 / RPC && c0 + lane < nchunk && flags[c0 + lane] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (m == 0) continue;
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t e = base + j * 32 + lane;
            v[j] = ((m >> (j / RPC)) & 1u) && e < n ? __ldcs(out + e) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t e = base + j * 32 + lane;
            if (((m >> (j / RPC)) & 1u) && e < n) host[e] = v[j];  // 256 contiguous bytes per warp store, posted over PCIe
        }
    }
}

int xfer_chunk_ok(int chunk) { return chunk == 32 || chunk == 64 || chunk == 128 || chunk == 256; }

int launch_chunk_flags(const double* out, int64_t n, int chunk, unsigned char* flags, int num_sms, void* stream) {
    if (n <= 0) return 0;
    if (!xfer_chunk_ok(chunk)) return (int)cudaErrorInvalidValue;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (chunk == 32) chunk_flags_kernel<32><<<num_sms * 8, 256, 0, st>>>(out, n, flags);
    else if (chunk == 64) chunk_flags_kernel<64><<<num_sms * 8, 256, 0, st>>>(out, n, flags);
    else if (chunk == 128) chunk_flags_kernel<128><<<num_sms * 8, 256, 0, st>>>(out, n, flags);
    else chunk_flags_kernel<256><<<num_sms * 8, 256, 0, st>>>(out, n, flags);
    return (int)cudaGetLastError();
}

int launch_chunk_push(const double* out, int64_t n, int chunk, const unsigned char* flags, double* host, int num_sms, void* stream) {
    if (n <= 0) return 0;
    if (!xfer_chunk_ok(chunk)) return (int)cudaErrorInvalidValue;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (chunk == 32) chunk_push_kernel<32><<<num_sms * 8, 256, 0, st>>>(out, n, flags, host);
    else if (chunk == 64) chunk_push_kernel<64><<<num_sms * 8, 256, 0, st>>>(out, n, flags, host);
    else if (chunk == 128) chunk_push_kernel<128><<<num_sms * 8, 256, 0, st>>>(out, n, flags, host);
    else chunk_push_kernel<256><<<num_sms * 8, 256, 0, st>>>(out, n, flags, host);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Kernels that are meant to share SMs (the class kernels and the screened fill) must ask for the
// same L1/shared split; with the default fill mode the split of every kernel is left to the driver
// (a larger L1 is worth ~2 % on the small classes).
static bool common_carveout() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MYQC_FILL_MODE");
        v = (e && e[0] == 's') ? 1 : 0;
    }
    return v == 1;
}

// Kernel attributes are set (and the kernels loaded: CUDA loads modules lazily, and a first-time load
// cannot complete while the paced fill kernel is waiting on the device for that very kernel) once
// per device, before the first launch of a plan.
template <int UT, int TT, int USL, bool MULTI>
static int prepare_one(int* occ_out) {
    using C = Cfg<UT, TT, USL>;
    auto kern = eri_class_kernel<UT, TT, USL, MULTI>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return (int)e;
    // Every kernel of a plan asks for the same (maximum) shared-memory carve-out: an SM cannot hold
    // CTAs of kernels with different L1/shared splits at the same time (measured:
    // tools/probes/concurrency_probe.cu), and the class kernels and the screened fill share SMs.
    if (common_carveout()) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return (int)e;
    }
    int occ = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NTHREADS, C::SMEM);
    if (e != cudaSuccess) return (int)e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return (int)e;
    *occ_out = occ < 1 ? 1 : occ;
    return 0;
}

template <int UT, bool MULTI>
static int prepare_ppw(int* occ_out) {
    auto kern = eri_ppw_kernel<UT, MULTI>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PPW<UT>::SMEM);
    if (e != cudaSuccess) return (int)e;
    int occ = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PPW<UT>::NTHREADS, PPW<UT>::SMEM);
    if (e != cudaSuccess) return (int)e;
    *occ_out = occ < 1 ? 1 : occ;
    return 0;
}

constexpr int kMaxDevices = 64;
static int g_occ[kMaxDevices][24];  // [slot] one-part kernels, [12 + slot] multi-part kernels; slots 9, 10: warp-cooperative kernels
static bool g_prepared[kMaxDevices];

int prepare_kernels() {
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return (int)ce;
    if (dev < 0 || dev >= kMaxDevices) return (int)cudaErrorInvalidDevice;
    if (g_prepared[dev]) return 0;
    int e = 0;
    if (!e) e = prepare_one<0, 0, -1, false>(&g_occ[dev][0]);
    if (!e) e = prepare_one<0, 0, -1, true>(&g_occ[dev][12 + 0]);
    if (!e) e = prepare_one<0, 1, -1, false>(&g_occ[dev][1]);
    if (!e) e = prepare_one<0, 1, -1, true>(&g_occ[dev][12 + 1]);
    if (!e) e = prepare_one<0, 2, -1, false>(&g_occ[dev][2]);
    if (!e) e = prepare_one<0, 2, -1, true>(&g_occ[dev][12 + 2]);
    if (!e) e = prepare_one<1, 1, -1, false>(&g_occ[dev][3]);
    if (!e) e = prepare_one<1, 1, -1, true>(&g_occ[dev][12 + 3]);
    if (!e) e = prepare_one<1, 2, -1, false>(&g_occ[dev][4]);
    if (!e) e = prepare_one<1, 2, -1, true>(&g_occ[dev][12 + 4]);
    if (!e) e = prepare_one<2, 2, 0, false>(&g_occ[dev][5]);
    if (!e) e = prepare_one<2, 2, 0, true>(&g_occ[dev][12 + 5]);
    if (!e) e = prepare_one<2, 2, 1, false>(&g_occ[dev][6]);
    if (!e) e = prepare_one<2, 2, 1, true>(&g_occ[dev][12 + 6]);
    if (!e) e = prepare_one<2, 2, 2, false>(&g_occ[dev][7]);
    if (!e) e = prepare_one<2, 2, 2, true>(&g_occ[dev][12 + 7]);
    if (!e) e = prepare_one<2, 2, 3, false>(&g_occ[dev][8]);
    if (!e) e = prepare_one<2, 2, 3, true>(&g_occ[dev][12 + 8]);
    if (!e) e = prepare_ppw<2, false>(&g_occ[dev][9]);
    if (!e) e = prepare_ppw<2, true>(&g_occ[dev][12 + 9]);
    if (!e) e = prepare_ppw<1, false>(&g_occ[dev][10]);
    if (!e) e = prepare_ppw<1, true>(&g_occ[dev][12 + 10]);
    if (e) return e;
    if (common_carveout()) {
        ce = cudaFuncSetAttribute(fill_screened_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared);
        if (ce != cudaSuccess) return (int)ce;
    }
    cudaFuncAttributes fa;
    ce = cudaFuncGetAttributes(&fa, fill_screened_kernel);
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaFuncSetAttribute(compose_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2ull * 8 * kCompCols * sizeof(double)));
    if (ce != cudaSuccess) return (int)ce;
    g_prepared[dev] = true;
    return 0;
}

template <int UT, int TT, int USL, bool MULTI>
static int launch_impl(const ClassArgs& a, int num_sms, cudaStream_t st, int slot) {
    using C = Cfg<UT, TT, USL>;
    auto kern = eri_class_kernel<UT, TT, USL, MULTI>;
    int dev = 0;
    cudaGetDevice(&dev);
    int e0 = prepare_kernels();
    if (e0) return e0;
    const int occ = g_occ[dev][slot + (MULTI ? 12 : 0)];
    int grid = num_sms * occ;
    const int need = (a.ntasks + C::NWARPS - 1) / C::NWARPS;
    if (grid > need) grid = need;
    kern<<<grid, C::NTHREADS, C::SMEM, st>>>(a);
    return (int)cudaGetLastError();
}

template <int UT, int TT, int USL>
static int launch_one(const ClassArgs& a, int num_sms, cudaStream_t st, int slot) {
    if (a.ntasks <= 0 || a.nparts <= 0) return 0;
    if (a.nparts > 1) return launch_impl<UT, TT, USL, true>(a, num_sms, st, slot);
    return launch_impl<UT, TT, USL, false>(a, num_sms, st, slot);
}

// MYQC_PP_KERNEL=warp: (SP SP|SP SP) always by the warp-cooperative kernel (one launch); =slices: always four mu-slices
// of the class kernel; unset: the plan decides per piece from its number of contracted quartets.  1 / 0 / -1.
int pp_kernel_mode() {  // read at plan creation (not cached: tests switch it between plans)
    const char* e = getenv("MYQC_PP_KERNEL");
    return !e ? -1 : (e[0] == 'w') ? 1 : (e[0] == 's') ? 0 : -1;
}
// the same for (S SP|SP SP): MYQC_SP_KERNEL = warp (1) / class (0) / unset (-1)
int sp_kernel_mode() {
    const char* e = getenv("MYQC_SP_KERNEL");
    return !e ? -1 : (e[0] == 'w') ? 1 : (e[0] == 'c') ? 0 : -1;
}

template <int UT, bool MULTI>
static int launch_ppw(const ClassArgs& a, int num_sms, cudaStream_t st) {
    int dev = 0;
    cudaGetDevice(&dev);
    int e0 = prepare_kernels();
    if (e0) return e0;
    int grid = num_sms * g_occ[dev][(UT == 2 ? 9 : 10) + (MULTI ? 12 : 0)];
    const int need = (a.ntasks + PPW<UT>::NWARPS - 1) / PPW<UT>::NWARPS;
    if (grid > need) grid = need;
    eri_ppw_kernel<UT, MULTI><<<grid, PPW<UT>::NTHREADS, PPW<UT>::SMEM, st>>>(a);
    return (int)cudaGetLastError();
}

int class_nlaunch(int UT, int TT) { return (UT == 2 && TT == 2) ? 4 : 1; }

int launch_class(int UT, int TT, int slice, const ClassArgs& a0, int num_sms, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ClassArgs a = a0;
    if (slice < 0) {  // warp-cooperative kernel of (S SP|SP SP) / (SP SP|SP SP): one launch, task counter 0
        if (TT != 2 || UT < 1) return (int)cudaErrorInvalidValue;
        if (a.ntasks <= 0 || a.nparts <= 0) return 0;
        if (UT == 2) return a.nparts > 1 ? launch_ppw<2, true>(a, num_sms, st) : launch_ppw<2, false>(a, num_sms, st);
        return a.nparts > 1 ? launch_ppw<1, true>(a, num_sms, st) : launch_ppw<1, false>(a, num_sms, st);
    }
    a.row_counter = a0.row_counter + slice;
    if (UT == 0 && TT == 0) return launch_one<0, 0, -1>(a, num_sms, st, 0);
    if (UT == 0 && TT == 1) return launch_one<0, 1, -1>(a, num_sms, st, 1);
    if (UT == 0 && TT == 2) return launch_one<0, 2, -1>(a, num_sms, st, 2);
    if (UT == 1 && TT == 1) return launch_one<1, 1, -1>(a, num_sms, st, 3);
    if (UT == 1 && TT == 2) return launch_one<1, 2, -1>(a, num_sms, st, 4);
    if (UT == 2 && TT == 2) {  // four mu-slices, each with its own task counter
        if (slice == 0) return launch_one<2, 2, 0>(a, num_sms, st, 5);
        if (slice == 1) return launch_one<2, 2, 1>(a, num_sms, st, 6);
        if (slice == 2) return launch_one<2, 2, 2>(a, num_sms, st, 7);
        if (slice == 3) return launch_one<2, 2, 3>(a, num_sms, st, 8);
    }
    return (int)cudaErrorInvalidValue;
}

int launch_fill_zero(double* out, int64_t n, int* counters, int ncounters, int num_sms, void* stream) {
    // A small footprint (default 2 CTAs of 256 threads per SM) is enough to saturate HBM with
    // 128-bit stores and leaves the SMs to the FP64 kernels that run next to a later region's fill.
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        const char* e = getenv("MYQC_FILL_CTAS");
        ctas_per_sm = e ? atoi(e) : 2;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    fill_zero_kernel<<<num_sms * ctas_per_sm, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n, counters, ncounters);
    return (int)cudaGetLastError();
}

int launch_fill_screened(const FillArgs& a, int num_sms, void* stream) {
    if (a.nunits <= 0) return 0;
    // persistent: a few CTAs per SM are enough to keep HBM busy with posted stores, and leave the
    // register file to the FP64 kernels that run next to the fill
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        const char* e = getenv("MYQC_FILL_CTAS");
        ctas_per_sm = e ? atoi(e) : 1;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    int grid = num_sms * ctas_per_sm;
    if (grid > a.nunits) grid = a.nunits;
    const int e0 = prepare_kernels();
    if (e0) return e0;
    fill_screened_kernel<<<grid, kFillThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return (int)cudaGetLastError();
}

int launch_expand_dense(const double* packed, int norb, int h0, int h1, double* xx, int num_sms, void* stream) {
    if (h1 <= h0) return 0;
    expand_dense_kernel<<<num_sms * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(packed, norb, h0, h1, xx);
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
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    cudaFree(d);
    *tflops = best;
    return (int)e;
}

}  // namespace myqc
