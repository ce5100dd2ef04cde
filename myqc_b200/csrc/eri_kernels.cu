// Owner-row strip kernels for the two-electron integrals on sm_100a (B200).
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
// How the work follows the output (the reference writes every element of XX once, int2e.f90:290-307):
//   * In the packed 8-fold-unique array the integrals of a shell quartet (AB|CD) all lie in the rows
//     of the shell pair whose first shell is smaller.  That pair is the quartet's OWNER u; the other
//     pair v = (C,D) is its partner.  (If A = C both pairs own some of the elements; the quartet is
//     then evaluated from both sides and each side keeps the elements of its own rows.)
//   * A task = (owner pair u; a range of partner first shells C).  For a fixed first index k the
//     columns (k,l) of a packed row are one contiguous run over l, so the task's part of u's rows is
//     a few contiguous runs per row.  One warp runs the task: it walks C and, for each C, the
//     partner second shells D >= C in blocks of 32 (one D per lane), keeps the (C,D) that pass the
//     pair-level bound emax_u*emax_v >= 1e-14, and evaluates them 32 at a time per partner kind
//     (D an S shell / D an SP shell), one quartet per lane.
//   * Results are parked in a per-warp shared-memory buffer.  When the buffer is full (or the task
//     ends) the warp FLUSHES: it writes zeros over the whole span of the packed rows it has walked
//     since the last flush (128-bit coalesced stores) and then stores the parked integrals into it.
//     Both happen within microseconds, so the 8-byte stores merge in L2 into sectors that are
//     already resident and fully written; every sector reaches DRAM once, whole.  There is no
//     separate zero-fill pass and no read-modify-write of evicted sectors.
//   * Kernel variants by (owner kind UT, partner-first-shell kind TC) keep the register budget of a
//     launch at what its two quartet classes need; tasks of different launches write disjoint
//     runs of the array, so the launches are independent.
//   * Per quartet (unchanged arithmetic): the owner's primitive records are staged into shared memory
//     by TMA bulk copy (cp.async.bulk + mbarrier) and read as warp-uniform broadcasts; K[f][H'] is
//     accumulated over the owner primitives (step A), the partner coefficients are folded once per
//     partner primitive (step B); far-field quartets (T >= 2Q+36) need one reciprocal square root
//     and no exponential; the Boys Taylor tables and an exp(-k/10) table live in shared memory.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <type_traits>

#include "eri_kernels.cuh"
#include "strip_geom.hpp"
#include "terms.hpp"
#include "boys.cuh"

namespace myqc {

namespace {

constexpr double kScreen = 1.0e-14;  // int2e.f90:257

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
template <int UT, int TC, int USL>
struct SCfg {
    static constexpr int NFU = (USL >= 0) ? 4 : tt_nf(UT);  // packed rows of the owner this launch writes
    static constexpr int NK = TC ? 4 : 1;                   // first indices k of a partner first shell
    static constexpr int FU = tt_nfield(UT);
    static constexpr uint32_t U_BYTES = 9 * FU * 8;
    static constexpr int NOUT0 = NFU * tt_nf(TC), NOUT1 = NFU * tt_nf(TC + 1);  // integrals per quartet, TD = 0 / 1
    static constexpr bool HEAVY = (NOUT1 > 16);
    // result buffer (doubles per warp) and the most quartet chunks it is cut into
    static constexpr int RES_CAP = HEAVY ? 3584 : (NOUT1 == 16 ? 1024 : 512);
    static constexpr int MAXCH = HEAVY ? 7 : (NOUT1 == 16 ? 8 : 16);
    static constexpr int NWARPS = HEAVY ? 6 : (NOUT1 == 16 ? 6 : 8);
    static constexpr int NTHREADS = 32 * NWARPS;
    static constexpr int MINB = HEAVY ? 1 : 2;
    static constexpr int NBUF = (UT == 2) ? 1 : 2;   // owner-record buffers per warp (TMA prefetch of the next task)
    static constexpr bool FT_SMEM = (UT < 2);        // Taylor tables in shared memory (else read through L1)
    static constexpr int DESC = 36;                  // ints per chunk descriptor: 32 items, kind|count, result offset
    static constexpr int NLIST = 8;                  // pending lists: partner kind (second shell S / SP) x 4 bins of surviving primitives
    static constexpr bool BINS = !HEAVY;             // the heavy launches see too few quartets per task to afford four lists per kind
    // shared memory: [Taylor tables] | exp table | per warp: owner records, mbarriers, rows, pending, descriptors, results
    static constexpr size_t OFF_EXP = FT_SMEM ? 2 * 121 * 8 * 8 : 0;
    static constexpr size_t OFF_WARP = OFF_EXP + 608 * 16;
    static constexpr size_t W_BAR = (size_t)NBUF * U_BYTES;
    static constexpr size_t W_ROW = W_BAR + 16;                 // int rowi[16], rowj[16]; int64 rbase[16]
    static constexpr size_t W_PEND = W_ROW + 16 * 4 * 2 + 16 * 8;
    static constexpr size_t W_DESC = W_PEND + (size_t)NLIST * 64 * 4;
    static constexpr size_t W_RES = W_DESC + (size_t)MAXCH * DESC * 4;
    static constexpr size_t W_BYTES = (W_RES + (size_t)RES_CAP * 8 + 15) / 16 * 16;
    static constexpr size_t SMEM = OFF_WARP + (size_t)NWARPS * W_BYTES;
    static_assert(W_RES % 16 == 0 && W_BAR % 16 == 0 && OFF_WARP % 16 == 0, "alignment of the per-warp areas");
    static_assert(SMEM <= 232448, "shared memory per CTA");
};

// partner function pair fp -> (index of k in C's slots, index of l in D's slots)
template <int TC, int TD>
__device__ __forceinline__ constexpr int fp_kc(int fp) { return (TC + TD == 0) ? 0 : (TC + TD == 1) ? (TC ? fp : 0) : (fp >> 2); }
template <int TC, int TD>
__device__ __forceinline__ constexpr int fp_ld(int fp) { return (TC + TD == 0) ? 0 : (TC + TD == 1) ? (TC ? 0 : fp) : (fp & 3); }

__device__ __forceinline__ int slot_of(const int4& f, int s) { return s == 0 ? f.x : (s == 1 ? f.y : (s == 2 ? f.z : f.w)); }

// ------------------------------------------------------------------------------------------
// One quartet (u | v) per lane: u = the warp's owner pair (records in shared memory), v = record `v`
// of the partner kind TT.  The NOUT integrals go to res[o*32], o = f*NFT + fp.  Returns the number
// of primitive quartets that passed the reference's screen.
template <int UT, int TT, int USL>
__device__ __forceinline__ unsigned quartet_lane(const double* __restrict__ s_u, int npu, double eu_max,
                                                 const double* __restrict__ trec, int npt, const double* __restrict__ ft,
                                                 const double2* __restrict__ s_exp, double* __restrict__ res) {
    constexpr int LT = UT + TT, Q = 3 * LT;  // Boys start order, int2e.f90:654-658,668
    constexpr int NR = h_count(LT), NHT = tt_nh(TT);
    constexpr int NFU = (USL >= 0) ? 4 : tt_nf(UT), NFT = tt_nf(TT);
    constexpr int NTU = tt_nterm(UT), NTT = tt_nterm(TT), FU = tt_nfield(UT), FT = tt_nfield(TT);
    constexpr int NOUT = NFU * NFT;
    constexpr bool OUT_SMEM = (NOUT > 16);

    double out_r[OUT_SMEM ? 1 : NOUT];
    if constexpr (OUT_SMEM) {
#pragma unroll
        for (int o = 0; o < NOUT; ++o) res[o * 32] = 0.0;
    } else {
#pragma unroll
        for (int o = 0; o < NOUT; ++o) out_r[o] = 0.0;
    }
    unsigned nquart = 0;
    // Partner records: the lane reads its own contiguous [9][FT] block (one base pointer, compile-time offsets;
    // every 32-byte sector it touches is used in full), and the next primitive's lines are prefetched into L1
    // while this one is contracted.
    for (int kt = 0; kt < npt; ++kt) {
        const double* tp = trec + kt * FT;
        const double2 t45 = __ldg(reinterpret_cast<const double2*>(tp) + 2);
        const double et = t45.x;
        if (eu_max * et < kScreen) break;  // primitives are sorted by E, descending
        const double2 t01 = __ldg(reinterpret_cast<const double2*>(tp));
        const double2 t23 = __ldg(reinterpret_cast<const double2*>(tp) + 1);
        if (kt + 1 < npt) {
#pragma unroll
            for (int b = 0; b < FT * 8; b += 128)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(tp + FT) + b));
        }
        const double q = t01.x, Qx = t01.y, Qy = t23.x, Qz = t23.y;
        const double cfar = kHalfSqrtPi * t45.y;  // sqrt(pi)/2 / sqrt(q)
        double K[NFU][NHT];
#pragma unroll
        for (int f = 0; f < NFU; ++f)
#pragma unroll
            for (int h = 0; h < NHT; ++h) K[f][h] = 0.0;

        for (int ku = 0; ku < npu; ++ku) {
            const double* up = s_u + ku * FU;
            if (up[4] * et < kScreen) break;  // IF (EGH*EIJ .LT. 1.0D-14) CYCLE  (int2e.f90:257)
            ++nquart;
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
                boys_near_mid<Q, LT>(alpha * R2, alpha, rs, G, ft, s_exp);
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
                    double ct = __ldg(tp + kRecCoef + k);
                    if constexpr (h_parity(hp) != 0) ct = -ct;
#pragma unroll
                    for (int f = 0; f < NFU; ++f) acc[f] = fma(ct, K[f][hp], acc[f]);
                }
            });
#pragma unroll
            for (int f = 0; f < NFU; ++f) {
                if constexpr (OUT_SMEM) res[(f * NFT + fp) * 32] += acc[f];
                else out_r[f * NFT + fp] += acc[f];
            }
        });
    }
    if constexpr (!OUT_SMEM) {
#pragma unroll
        for (int o = 0; o < NOUT; ++o) res[o * 32] = out_r[o];
    }
    return nquart;
}

// warp-cooperative zero fill of out[0..n): 128-bit stores, the run may start on an odd element
__device__ __forceinline__ void zero_run(double* __restrict__ p, int64_t n, int lane) {
    if (n <= 0) return;
    const int64_t head = (reinterpret_cast<uintptr_t>(p) & 15) ? 1 : 0;
    if (head && lane == 0) p[0] = 0.0;
    double2* q = reinterpret_cast<double2*>(p + head);
    const int64_t n2 = (n - head) >> 1;
    for (int64_t x = lane; x < n2; x += 32) q[x] = make_double2(0.0, 0.0);
    if (((n - head) & 1) && lane == 0) p[n - 1] = 0.0;
}

// per-warp view of the shared-memory areas
template <class C>
struct WarpMem {
    double* ubuf;
    uint64_t* bar;
    int* rowi;
    int* rowj;
    long long* rbase;
    int* pend;  // [NLIST][64] items (C << 16 | D)
    int* desc;  // [MAXCH][DESC]
    double* res;
    __device__ __forceinline__ explicit WarpMem(unsigned char* w) {
        ubuf = reinterpret_cast<double*>(w);
        bar = reinterpret_cast<uint64_t*>(w + C::W_BAR);
        rowi = reinterpret_cast<int*>(w + C::W_ROW);
        rowj = rowi + 16;
        rbase = reinterpret_cast<long long*>(w + C::W_ROW + 128);
        pend = reinterpret_cast<int*>(w + C::W_PEND);
        desc = reinterpret_cast<int*>(w + C::W_DESC);
        res = reinterpret_cast<double*>(w + C::W_RES);
    }
};

// stores the parked integrals of one chunk (kind TD) into the packed rows of the owner
template <int UT, int TC, int TD, int USL>
__device__ __forceinline__ void scatter_chunk(const StripArgs& a, const WarpMem<SCfg<UT, TC, USL>>& m, const int* __restrict__ d,
                                              int n, int lane) {
    using C = SCfg<UT, TC, USL>;
    constexpr int NFU = C::NFU, NFT = tt_nf(TC + TD);
    if (lane >= n) return;
    const int item = d[lane];
    const int Cs = item >> 16, Ds = item & 0xffff;
    const int4 fc = __ldg(a.sh_fn + Cs);
    const int4 fd = __ldg(a.sh_fn + Ds);
    const double* __restrict__ col = m.res + d[33] + lane;
    long long cb[C::NK];
#pragma unroll
    for (int kc = 0; kc < C::NK; ++kc) cb[kc] = col_base64(slot_of(fc, kc), a.norb);
#pragma unroll
    for (int f = 0; f < NFU; ++f) {
        const int i = m.rowi[f];
        if (i < 0) continue;  // the row does not exist (absent function, or the duplicate half of a diagonal pair)
        const int j = m.rowj[f];
        const long long rb = m.rbase[f];
#pragma unroll
        for (int fp = 0; fp < NFT; ++fp) {
            const int kc = fp_kc<TC, TD>(fp), ldx = fp_ld<TC, TD>(fp);
            const int k = slot_of(fc, kc), l = slot_of(fd, ldx);
            // the element exists in row (i,j) iff (k,l) is an orbital pair with k <= l at or after (i,j)
            if (k < 0 || l < 0 || k > l || k < i || (k == i && l < j)) continue;
            a.out[rb + cb[kc] + l] = col[(f * NFT + fp) * 32];
        }
    }
}

}  // namespace

template <int UT, int TC, int USL>
__global__ void __launch_bounds__(SCfg<UT, TC, USL>::NTHREADS, SCfg<UT, TC, USL>::MINB) eri_strip_kernel(const __grid_constant__ StripArgs a) {
    using C = SCfg<UT, TC, USL>;
    constexpr int NFU = C::NFU, NK = C::NK, FU = C::FU;
    constexpr int CH0 = 32 * C::NOUT0, CH1 = 32 * C::NOUT1;
    constexpr int Q0 = 3 * (UT + TC), Q1 = Q0 + 3;
    (void)Q0; (void)Q1;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const double* ft0 = a.ftab_q[0];
    const double* ft1 = a.ftab_q[1];
    if constexpr (C::FT_SMEM) {
        double* s_ft = reinterpret_cast<double*>(smem_raw);
        for (int i = tid; i < 121 * 8; i += C::NTHREADS) { s_ft[i] = a.ftab_q[0][i]; s_ft[121 * 8 + i] = a.ftab_q[1][i]; }
        ft0 = s_ft;
        ft1 = s_ft + 121 * 8;
    }
    double2* s_exp = reinterpret_cast<double2*>(smem_raw + C::OFF_EXP);
    for (int i = tid; i < 601; i += C::NTHREADS) s_exp[i] = a.exptab[i];
    const WarpMem<C> m(smem_raw + C::OFF_WARP + (size_t)warp * C::W_BYTES);
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < C::NBUF; ++b) mbar_init(&m.bar[b], 1);
        fence_mbar_init();
    }
    __syncthreads();

    const int ns = a.ns, nblk = a.nblk;

    // ---- warp-autonomous task loop; the owner records of the next task arrive by TMA meanwhile ----
    auto fetch = [&](int slot, int& t, int4& task, int& urec) {
        t = 0; urec = -1; task = make_int4(0, 0, 0, 0);
        if (lane == 0) {
            t = atomicAdd(a.counter, 1);
            if (t < a.ntasks) {
                task = a.tasks[t];
                urec = __ldg(a.pair_rec + (size_t)(task.x & 0xffff) * ns + (task.x >> 16));
                if (urec >= 0) {
                    mbar_expect_tx(&m.bar[slot], C::U_BYTES);
                    tma_bulk_g2s(m.ubuf + (size_t)slot * 9 * FU, a.u_aos + (size_t)urec * 9 * FU, C::U_BYTES, &m.bar[slot]);
                }
            }
        }
        t = __shfl_sync(0xffffffffu, t, 0);
        urec = __shfl_sync(0xffffffffu, urec, 0);
        task.x = __shfl_sync(0xffffffffu, task.x, 0);
        task.y = __shfl_sync(0xffffffffu, task.y, 0);
        task.z = __shfl_sync(0xffffffffu, task.z, 0);
        task.w = __shfl_sync(0xffffffffu, task.w, 0);
    };
    int t, urec, buf = 0;
    int4 task;
    uint32_t parity0 = 0, parity1 = 0;
    fetch(0, t, task, urec);
    while (t < a.ntasks) {
        int tn = 0, urecn = -1;
        int4 taskn = make_int4(0, 0, 0, 0);
        if constexpr (C::NBUF == 2) fetch(buf ^ 1, tn, taskn, urecn);
        // task = {A | B << 16, first index into clist, end index, b0 | b1 << 16}: the span runs from block b0 of the
        // first C through block b1-1 of the last C (b0 = block of the first C and b1 = nblk for whole runs)
        const int A = task.x & 0xffff, B = task.x >> 16;
        const int c_lo = task.y, c_hi = task.z, b_lo = task.w & 0xffff, b_hi = task.w >> 16;
        // ---- the packed rows of the owner: orbitals (i,j), and the offset that turns a pair index P(k,l) into an address
        if (lane < 16) {
            int i = -1, j = -1;
            long long rb = 0;
            if (lane < NFU) {
                const int4 fa4 = __ldg(a.sh_fn + A), fb4 = __ldg(a.sh_fn + B);
                const int fa[4] = {fa4.x, fa4.y, fa4.z, fa4.w}, fb[4] = {fb4.x, fb4.y, fb4.z, fb4.w};
                const bool a_is_sp = (UT == 1) && a.sh_type[A] == 1;
                if (owner_row(UT, USL, a_is_sp, A == B, fa, fb, lane, &i, &j)) {
                    const long long P1 = pair_index64(i, j, a.norb);
                    rb = row_start64(P1, a.npair) - P1 - a.out_offset;
                } else {
                    i = -1;
                }
            }
            m.rowi[lane] = i; m.rowj[lane] = j; m.rbase[lane] = rb;
        }
        int npu = 0;
        double eu_max = 0.0;
        const double* s_u = m.ubuf + (size_t)buf * 9 * FU;
        if (urec >= 0) {
            npu = __ldg(a.u_nprim + urec);
            if (buf == 0) { mbar_wait(&m.bar[0], parity0); parity0 ^= 1; }
            else          { mbar_wait(&m.bar[C::NBUF - 1], parity1); parity1 ^= 1; }
            eu_max = s_u[4];
        }
        __syncwarp();

        // ---- the warp's state machine for one task.  Everything before (zci, zb) has been written
        // (zb < 0: from the start of that C); (pci, pb) is the last block enumerated.  Each pass of the
        // loop does one thing: flush, evaluate one chunk of 32 quartets, take up to 32 partners from the
        // current segment, or move the cursor (class -> block -> first shell).
        // Partners: for every first shell C and block b of second shells, the D >= C of one shell class that form a
        // live pair with C are listed by decreasing emax(C,D) (same exponents: by increasing distance), so the
        // partners that pass the bound emax_u*emax_v >= 1e-14 are a prefix of the segment, and the 32 quartets a
        // chunk evaluates have the same partner exponents and similar distances (same primitive survival, same
        // Boys regime).  pc = eight pending counts (two kinds x four survival bins), one byte each (always < 64).
        int zci = c_lo, zb = b_lo;
        int nch = 0, used = 0;
        unsigned long long pc = 0ull;
        unsigned tq0 = 0, tq1 = 0;
        long long span = 0;
        int ci = c_lo - 1, Cs = 0, b = 0, bfirst_c = 0, bend = 0, cls = a.ncls;
        int sbase = 0, send = 0, skind = 0;
        bool blk_live = false;
        int pci = c_lo, pb = b_lo - 1;  // last block whose partners have all been listed
        int fci = c_lo, fb = b_lo - 1;  // newest block a parked chunk holds integrals of: a flush must reach it
        bool finishing = false;
        if (urec < 0) { finishing = true; pci = c_hi - 1; pb = b_hi - 1; }  // dead owner: its rows are zeros
        for (;;) {
            int L = -1;
            {
                const unsigned long long r = pc & 0xE0E0E0E0E0E0E0E0ull;  // lists with >= 32 pending
                if (r) L = (__ffsll((long long)r) - 1) >> 3;
                else if (finishing && pc) L = (__ffsll((long long)pc) - 1) >> 3;
            }
            const int kind = L >> 2;  // 0: D is an S shell, 1: D is an SP shell (meaningful when L >= 0)
            bool do_flush = false, last = false;
            if (L >= 0) do_flush = (used + (kind ? CH1 : CH0) > C::RES_CAP) || (nch == C::MAXCH);
            else if (finishing) { do_flush = true; last = true; }
            // bound the span one flush zero-fills, so that the integrals stored right after it meet their sectors in L2
            else if (span > (256 << 10) && nch > 0) do_flush = true;
            if (do_flush) {
                // ---- zeros over the span from (zci,zb) through (tci,tb), then the parked integrals into it.  The span
                // ends at the newest block the parked chunks touch (blocks after it are zero-filled together with
                // their own integrals, by a later flush); the last flush of a task runs to the task's end.
                const int tci = last ? pci : fci, tb = last ? pb : fb;
                for (int f = 0; f < NFU; ++f) {
                    const int i = m.rowi[f];
                    if (i < 0) continue;
                    const int j = m.rowj[f];
                    const long long rb = m.rbase[f];
                    long long pa = 0, pcount = 0;
                    for (int c2 = zci; c2 <= tci; ++c2) {
                        const int C2 = __ldg(a.clist + c2);
                        const int bfirst = C2 / kBlockShells;
                        const int blo = (c2 == zci && zb >= 0) ? zb : bfirst;
                        const int bhi = (c2 == tci) ? tb : nblk - 1;
                        if (blo > bhi) continue;
                        const int llo = (blo == bfirst) ? 0 : __ldg(a.sh_first + blo * kBlockShells);
                        const int lhi = (bhi == nblk - 1) ? a.norb : __ldg(a.sh_first + (bhi + 1) * kBlockShells);
                        const int4 fc = __ldg(a.sh_fn + C2);
#pragma unroll
                        for (int kc = 0; kc < NK; ++kc) {
                            const int k = slot_of(fc, kc);
                            if (k < 0) continue;
                            int l0;
                            const int n = row_piece(i, j, k, llo, lhi, &l0);
                            if (n == 0) continue;
                            const long long addr = rb + col_base64(k, a.norb) + l0;
                            if (addr == pa + pcount) { pcount += n; }
                            else { zero_run(a.out + pa, pcount, lane); pa = addr; pcount = n; }
                        }
                    }
                    zero_run(a.out + pa, pcount, lane);
                }
                __syncwarp();  // orders the zeros before the integrals stored below by other lanes
                for (int ch = 0; ch < nch; ++ch) {
                    const int* d = m.desc + ch * C::DESC;
                    const int hdr = d[32];
                    if ((hdr & 1) == 0) scatter_chunk<UT, TC, 0, USL>(a, m, d, hdr >> 8, lane);
                    else scatter_chunk<UT, TC, 1, USL>(a, m, d, hdr >> 8, lane);
                }
                __syncwarp();
                nch = 0; used = 0; span = 0;
                if (tb >= nblk - 1) { zci = tci + 1; zb = -1; }
                else { zci = tci; zb = tb + 1; }
                if (last) break;
                continue;
            }
            if (L >= 0) {
                // ---- evaluate the first (up to) 32 pending quartets of list L and park their integrals
                int* pl = m.pend + L * 64;
                const int pn = (int)((pc >> (8 * L)) & 0xffull);
                const int n = pn < 32 ? pn : 32;
                int* d = m.desc + nch * C::DESC;
                int item = 0;
                if (lane < n) item = pl[lane];
                const int keep = (lane + 32 < pn) ? pl[lane + 32] : 0;
                __syncwarp();
                if (lane + 32 < pn) pl[lane] = keep;
                d[lane] = item;
                if (lane == 0) { d[32] = kind | (n << 8); d[33] = used; }
                unsigned nq = 0;
                if (lane < n) {
                    const int v = __ldg(a.pair_rec + (size_t)(item >> 16) * ns + (item & 0xffff));
                    if (kind == 0) {
                        constexpr int FT = tt_nfield(TC);
                        nq = quartet_lane<UT, TC, USL>(s_u, npu, eu_max, a.t_aos[0] + (size_t)v * 9 * FT, __ldg(a.t_nprim[0] + v), ft0,
                                                       s_exp, m.res + used + lane);
                    } else {
                        constexpr int FT = tt_nfield(TC + 1);
                        nq = quartet_lane<UT, TC + 1, USL>(s_u, npu, eu_max, a.t_aos[1] + (size_t)v * 9 * FT, __ldg(a.t_nprim[1] + v), ft1,
                                                           s_exp, m.res + used + lane);
                    }
                }
                if (kind) tq1 += nq; else tq0 += nq;
                // the chunk may hold partners of the block the cursor is in
                if (ci >= c_lo && ci < c_hi && b >= bfirst_c && b < bend) { fci = ci; fb = b; }
                else { fci = pci; fb = pb; }
                __syncwarp();
                pc -= (unsigned long long)n << (8 * L);
                used += kind ? CH1 : CH0;
                ++nch;
                continue;
            }
            if (sbase < send) {
                // ---- up to 32 partners of the current segment: those that pass the pair-level bound (a prefix).
                // Each goes to the pending list of its kind and of the number of its primitives that survive
                // against this owner (1-2, 3-4, 5-6, 7-9): the lanes of a chunk then run about the same number of
                // primitive quartets.  ep = {E(1) = emax, E(3), E(5), E(7)} of the partner's sorted prefactors.
                const int e = sbase + lane;
                double2 ep01 = make_double2(0.0, 0.0), ep23 = make_double2(0.0, 0.0);
                if (e < send) {
                    ep01 = __ldg(a.seg_eprof + 2 * (size_t)e);
                    ep23 = __ldg(a.seg_eprof + 2 * (size_t)e + 1);
                }
                const bool ok = eu_max * ep01.x >= kScreen;
                const int bin = C::BINS ? (eu_max * ep01.y >= kScreen ? 1 : 0) + (eu_max * ep23.x >= kScreen ? 1 : 0) + (eu_max * ep23.y >= kScreen ? 1 : 0) : 0;
                const unsigned mk = __ballot_sync(0xffffffffu, ok);
                const int cnt = __popc(mk);
                const unsigned lt = (1u << lane) - 1u;
                int item = 0;
                if (ok) item = (Cs << 16) | (int)__ldg(a.seg_d + e);
#pragma unroll
                for (int j = 0; j < (C::BINS ? 4 : 1); ++j) {
                    const unsigned mj = __ballot_sync(0xffffffffu, ok && bin == j);
                    const int L2 = skind * 4 + j;
                    const int pn = (int)((pc >> (8 * L2)) & 0xffull);
                    if (ok && bin == j) m.pend[L2 * 64 + pn + __popc(mj & lt)] = item;
                    pc += (unsigned long long)__popc(mj) << (8 * L2);
                }
                sbase = (cnt == 32) ? sbase + 32 : send;
                __syncwarp();
                continue;
            }
            // ---- move the cursor
            if (blk_live && cls + 1 < a.ncls) {
                ++cls;
                const size_t sidx = ((size_t)Cs * nblk + b) * a.ncls + cls;
                sbase = __ldg(a.seg_start + sidx);
                send = __ldg(a.seg_start + sidx + 1);
                skind = __ldg(a.cls_kind + cls);
                continue;
            }
            if (ci >= c_lo && b >= bfirst_c && b < bend) { pci = ci; pb = b; }  // block (ci,b) has been enumerated
            if (ci >= c_lo && b + 1 < bend) {
                ++b;
                blk_live = eu_max * __ldg(a.blk_emax + (size_t)Cs * nblk + b) >= kScreen;
                cls = -1;
                continue;
            }
            if (ci >= c_lo) span += (long long)NFU * NK * (a.norb - __ldg(a.sh_first + Cs)) * 8;  // C is done
            ++ci;
            if (ci >= c_hi) { finishing = true; continue; }
            Cs = __ldg(a.clist + ci);
            bend = (ci == c_hi - 1) ? b_hi : nblk;
            bfirst_c = (ci == c_lo) ? b_lo : Cs / kBlockShells;
            b = bfirst_c - 1;
            blk_live = false;
            cls = a.ncls;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            tq0 += __shfl_xor_sync(0xffffffffu, tq0, o);
            tq1 += __shfl_xor_sync(0xffffffffu, tq1, o);
        }
        if (lane == 0) {
            if (tq0) atomicAdd(a.stats, (unsigned long long)tq0);
            if (tq1) atomicAdd(a.stats + 1, (unsigned long long)tq1);
        }

        __syncwarp();  // every lane is done with the owner records before the next TMA reuses the buffer
        if constexpr (C::NBUF == 1) fetch(0, tn, taskn, urecn);
        t = tn; task = taskn; urec = urecn;
        if constexpr (C::NBUF == 2) buf ^= 1;
    }
}

// ------------------------------------------------------------------------------------------
// dense XX(i,j,g,h) (column-major, i fastest) for h in [h0,h1) from the packed array: the fillsym pass of the
// reference (int2e.f90:290-304,540-554) done as a gather so that the 8n^4-byte stream is written
// once, coalesced.
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
// Sparse transfer to the host.  A warp owns a chunk of kXferChunk consecutive elements of the slice.
__global__ void __launch_bounds__(256) chunk_flags_kernel(const double* __restrict__ out, int64_t n, unsigned char* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int64_t nchunk = (n + kXferChunk - 1) / kXferChunk;
    const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < nchunk; c += nwarp) {
        const int64_t base = c * kXferChunk;
        long long bits = 0;
#pragma unroll
        for (int j = 0; j < kXferChunk / 32; ++j) {
            const int64_t e = base + j * 32 + lane;
            if (e < n) bits |= __double_as_longlong(__ldcs(out + e));
        }
        const bool any = __any_sync(0xffffffffu, bits != 0);
        if (lane == 0) flags[c] = any ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) chunk_push_kernel(const double* __restrict__ out, int64_t n, const unsigned char* __restrict__ flags,
                                                         double* __restrict__ host) {
    const int lane = threadIdx.x & 31;
    const int64_t nchunk = (n + kXferChunk - 1) / kXferChunk;
    const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < nchunk; c += nwarp) {
        if (!flags[c]) continue;
        const int64_t base = c * kXferChunk;
        double v[kXferChunk / 32];
#pragma unroll
        for (int j = 0; j < kXferChunk / 32; ++j) {
            const int64_t e = base + j * 32 + lane;
            v[j] = e < n ? __ldcs(out + e) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < kXferChunk / 32; ++j) {
            const int64_t e = base + j * 32 + lane;
            if (e < n) host[e] = v[j];  // 256 contiguous bytes per warp store, posted over PCIe
        }
    }
}

int launch_chunk_flags(const double* out, int64_t n, unsigned char* flags, int num_sms, void* stream) {
    if (n <= 0) return 0;
    chunk_flags_kernel<<<num_sms * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n, flags);
    return (int)cudaGetLastError();
}

int launch_chunk_push(const double* out, int64_t n, const unsigned char* flags, double* host, int num_sms, void* stream) {
    if (n <= 0) return 0;
    chunk_push_kernel<<<num_sms * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n, flags, host);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Kernel attributes are set (and the kernels loaded: CUDA loads modules lazily) once per device.
constexpr int kNumVariants = 9;  // (0,0) (0,1) (1,0) (1,1) (2,0) (2,1)x4
template <int UT, int TC, int USL>
static int prepare_one(int* occ_out) {
    using C = SCfg<UT, TC, USL>;
    auto kern = eri_strip_kernel<UT, TC, USL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return (int)e;
    int occ = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NTHREADS, C::SMEM);
    if (e != cudaSuccess) return (int)e;
    *occ_out = occ < 1 ? 1 : occ;
    return 0;
}

constexpr int kMaxDevices = 64;
static int g_occ[kMaxDevices][kNumVariants];
static bool g_prepared[kMaxDevices];

int prepare_kernels() {
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return (int)ce;
    if (dev < 0 || dev >= kMaxDevices) return (int)cudaErrorInvalidDevice;
    if (g_prepared[dev]) return 0;
    int e = 0;
    if (!e) e = prepare_one<0, 0, -1>(&g_occ[dev][0]);
    if (!e) e = prepare_one<0, 1, -1>(&g_occ[dev][1]);
    if (!e) e = prepare_one<1, 0, -1>(&g_occ[dev][2]);
    if (!e) e = prepare_one<1, 1, -1>(&g_occ[dev][3]);
    if (!e) e = prepare_one<2, 0, -1>(&g_occ[dev][4]);
    if (!e) e = prepare_one<2, 1, 0>(&g_occ[dev][5]);
    if (!e) e = prepare_one<2, 1, 1>(&g_occ[dev][6]);
    if (!e) e = prepare_one<2, 1, 2>(&g_occ[dev][7]);
    if (!e) e = prepare_one<2, 1, 3>(&g_occ[dev][8]);
    if (e) return e;
    g_prepared[dev] = true;
    return 0;
}

template <int UT, int TC, int USL>
static int launch_one(const StripArgs& a, int num_sms, cudaStream_t st, int slot) {
    if (a.ntasks <= 0) return 0;
    using C = SCfg<UT, TC, USL>;
    auto kern = eri_strip_kernel<UT, TC, USL>;
    int dev = 0;
    cudaGetDevice(&dev);
    int e0 = prepare_kernels();
    if (e0) return e0;
    int grid = num_sms * g_occ[dev][slot];
    const int need = (a.ntasks + C::NWARPS - 1) / C::NWARPS;
    if (grid > need) grid = need;
    kern<<<grid, C::NTHREADS, C::SMEM, st>>>(a);
    return (int)cudaGetLastError();
}

int strip_nslices(int UT, int TC) { return (UT == 2 && TC == 1) ? 4 : 1; }

int launch_strip(int UT, int TC, int slice, const StripArgs& a, int num_sms, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (UT == 0 && TC == 0) return launch_one<0, 0, -1>(a, num_sms, st, 0);
    if (UT == 0 && TC == 1) return launch_one<0, 1, -1>(a, num_sms, st, 1);
    if (UT == 1 && TC == 0) return launch_one<1, 0, -1>(a, num_sms, st, 2);
    if (UT == 1 && TC == 1) return launch_one<1, 1, -1>(a, num_sms, st, 3);
    if (UT == 2 && TC == 0) return launch_one<2, 0, -1>(a, num_sms, st, 4);
    if (UT == 2 && TC == 1) {
        if (slice == 0) return launch_one<2, 1, 0>(a, num_sms, st, 5);
        if (slice == 1) return launch_one<2, 1, 1>(a, num_sms, st, 6);
        if (slice == 2) return launch_one<2, 1, 2>(a, num_sms, st, 7);
        if (slice == 3) return launch_one<2, 1, 3>(a, num_sms, st, 8);
    }
    return (int)cudaErrorInvalidValue;
}

int launch_expand_dense(const double* packed, int norb, int h0, int h1, double* xx_slab, int num_sms, void* stream) {
    if (h1 <= h0) return 0;
    expand_dense_kernel<<<num_sms * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(packed, norb, h0, h1, xx_slab);
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
