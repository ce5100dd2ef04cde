// AO -> MO four-index transformation from the packed unique ERI array, on the device
// (include/myqc_ao2mo.h).
//
// Reference: src/ao2mo/ao2mo.f90 -- idx1_trans..idx4_trans (:1306-1439) are four explicit O(n^5)
// loop nests over the dense XX(n,n,n,n) array read from disk; slow_ao2mo_MP2_RHF (:465-602),
// slow_ao2mo_MP2_UHF (:614-904) and slow_ao2mo_CIS_UHF (:919-1227) call them with different
// coefficient column blocks.  All of them are
//     O(p,q,r,s) = sum_{u,v,l,d} C1(u,p) C2(v,q) C3(l,r) C4(d,s) (uv|ld).
//
// Here (n = norb, NP = n(n+1)/2 pairs, X = the symmetric NP x NP matrix whose upper triangle is
// the packed array):
//   stage 1, per panel of BP packed rows P:
//     unpack   Xsq[P][l][d] = X[P, pair(l,d)]                       (gather, both triangles)
//     GEMM 1a  T[(P,l)][s]  = sum_d Xsq[(P,l)][d] C4(d,s)           M = BP n, K = n, N = n4
//     GEMM 1b  H[P][s][r]   = sum_l T[P][l][s]    C3(l,r)           batched over P
//   stage 2, per chunk of RSB columns rs = s n3 + r of H:
//     unpack   Gsq[u][v][rs] = H[pair(u,v)][rs]                     (row gather, coalesced)
//     GEMM 2a  V[u][rs][q]   = sum_v Gsq[u][v][rs] C2(v,q)          batched over u
//     GEMM 2b  O[p][(rs,q)]  = sum_u C1(u,p) V[u][(rs,q)]           written straight into Om(p,q,r,s)
// i.e. 2 NP n^2 n4 + 2 NP n n3 n4 + 2 n^2 n2 n3 n4 + 2 n n1 n2 n3 n4 flops instead of the reference's
// 2 n^4 (n1 + ...) -- the pair symmetry of (uv| and |ld) halves both halves.
//
// The GEMM: 128 x 128 x 16 CTA tiles, 8 warps of 64 x 32, FP64 tensor pipe (mma.sync m8n8k4 f64 --
// FP64 has no tcgen05 form), operands staged through padded shared memory (conflict-free fragment
// reads), next tile prefetched into registers while the current one is multiplied.  A second,
// plain-DFMA instantiation of the same tiling (MYQC_AO2MO_GEMM=simt) exists to cross-check the
// fragment layout in the tests.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/myqc_ao2mo.h"
#include "../../include/myqc_eri.h"

namespace myqc {
int fock_fail(int code, const std::string& msg);  // sets myqc_last_error (eri_api.cu)
}

namespace {

#define CUA(x)                                                                                   \
    do {                                                                                         \
        cudaError_t e_ = (x);                                                                    \
        if (e_ != cudaSuccess)                                                                   \
            return myqc::fock_fail(MYQC_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
    } while (0)

constexpr int BM = 128, BN = 128, BK = 16, GT = 256;
constexpr int PA_K = BK + 4;   // pitch of A stored [m][k]   (20 = 4 mod 16: fragment reads conflict-free)
constexpr int PA_M = BM + 4;   // pitch of A stored [k][m]   (132 = 4 mod 16)
constexpr int PB = BN + 4;     // pitch of B stored [k][n]
constexpr int A_ELEMS = (BM * PA_K > BK * PA_M) ? BM * PA_K : BK * PA_M;
constexpr int B_ELEMS = BK * PB;
constexpr size_t GEMM_SMEM = 2 * (size_t)(A_ELEMS + B_ELEMS) * sizeof(double);

struct GemmArgs {
    const double* A; int64_t a_sm, a_sk, a_batch;  // A(m,k) at m*a_sm + k*a_sk
    const double* B; int64_t ldb, b_batch;         // B(k,n) at k*ldb + n
    double* C; int64_t c_sm, c_sn, c_batch;        // C(m,n) at m*c_sm + n*c_sn
    int M, N, K;
    int g_n, g_np;  // A mode 2: A(m = pp*g_n + l, k = d) = A[pp*g_np + pair(min(l,d), max(l,d))]
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// C = A B.  AMODE 1: A is contiguous along k (loader walks k fastest, tile kept [m][k]); 0: A is
// contiguous along m (loader walks m fastest, tile kept [k][m]); 2: row m = (pp, l) of A is row l of
// the symmetric n x n matrix whose upper triangle is the packed row pp of a panel (both triangles
// gathered from the one stored, thread mapping and tile as mode 1).  MMA: tensor pipe or plain DFMA.
template <int AMODE, bool MMA>
__global__ void __launch_bounds__(GT) gemm_f64_kernel(const GemmArgs g) {
    constexpr bool A_KC = AMODE != 0;
    extern __shared__ __align__(16) double gsm[];
    constexpr int STAGE = A_ELEMS + B_ELEMS;  // stage b: A tile at gsm + b*STAGE, B tile behind it
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const double* A = g.A + (int64_t)blockIdx.z * g.a_batch;
    const double* B = g.B + (int64_t)blockIdx.z * g.b_batch;
    double* C = g.C + (int64_t)blockIdx.z * g.c_batch;
    const int M = g.M, N = g.N, K = g.K;

    auto a_at = [](int m, int k) { return A_KC ? m * PA_K + k : k * PA_M + m; };

    // Element e = tid + 256 i of a tile: the eight elements a thread moves per operand differ by a
    // uniform step, so one base pointer per operand and loop-invariant row/column masks are enough.
    //   A, k-contiguous: m = tid/16 + 16 i, k = tid%16      A, m-contiguous: m = tid%128, k = tid/128 + 2 i
    //   B:               k = tid/128 + 2 i, n = tid%128
    const int am = A_KC ? tid / BK : tid % BM, ak = A_KC ? tid % BK : tid / BM;
    const int bk = tid / BN, bn = tid % BN;
    const double* pa = A + (AMODE == 2 ? 0 : (int64_t)(m0 + am) * g.a_sm + (int64_t)ak * g.a_sk);
    const double* pb = B + (int64_t)bk * g.ldb + (n0 + bn);
    const int64_t a_step = A_KC ? 16 * g.a_sm : 2 * g.a_sk;   // between the eight elements
    const int64_t b_step = 2 * g.ldb;
    const int64_t a_tile = (int64_t)BK * g.a_sk, b_tile = (int64_t)BK * g.ldb;  // between k tiles
    unsigned a_ok = 0;  // A_KC: bit i = row m0 + am + 16 i exists; else: the one row m0 + am exists
#pragma unroll
    for (int i = 0; i < 8; ++i) a_ok |= ((m0 + am + (A_KC ? 16 * i : 0)) < M ? 1u : 0u) << i;
    const bool b_ok = n0 + bn < N;
    // mode 2: first row of this thread as (panel row pp, AO index l); the other seven follow by +16
    const int gn = g.g_n, gnp = g.g_np;
    int g_l0 = 0, g_rb0 = 0, g_w = 8;  // g_w: first i whose row has wrapped into the next panel row (n >= 128)
    if (AMODE == 2) {
        const int pp = (m0 + am) / gn;
        g_l0 = (m0 + am) - pp * gn;
        g_rb0 = pp * gnp;
        g_w = (gn - g_l0 + 15) >> 4;
    }
    double ra[8], rb[8];
    auto gload = [&](int k0) {
        if constexpr (AMODE == 2) {
            const int d = k0 + ak;
            const int td = d * gn - ((d * (d - 1)) >> 1) - d;  // pair(d, l) = td + l for l >= d
            const bool kin = d < K;
            if (gn >= 128) {  // at most one wrap inside the 8 x 16 rows of a thread: no dependent chain
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const bool wr = i >= g_w;
                    const int l = g_l0 + 16 * i - (wr ? gn : 0);
                    const int rbase = g_rb0 + (wr ? gnp : 0);
                    const int tl = l * gn - ((l * (l - 1)) >> 1) - l;  // pair(l, d) = tl + d for d >= l
                    const int off = rbase + (d >= l ? tl + d : td + l);
                    ra[i] = (kin && ((a_ok >> i) & 1u)) ? __ldg(pa + off) : 0.0;
                }
            } else {
                int l = g_l0, rbase = g_rb0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int tl = l * gn - ((l * (l - 1)) >> 1) - l;
                    const int off = rbase + (d >= l ? tl + d : td + l);
                    ra[i] = (kin && ((a_ok >> i) & 1u)) ? __ldg(pa + off) : 0.0;
                    l += 16;
                    while (l >= gn) { l -= gn; rbase += gnp; }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if constexpr (AMODE != 2) {
                const bool kin = A_KC ? (k0 + ak < K) : (k0 + ak + 2 * i < K);
                ra[i] = (kin && ((a_ok >> i) & 1u)) ? __ldg(pa + i * a_step) : 0.0;
            }
            rb[i] = (b_ok && k0 + bk + 2 * i < K) ? __ldg(pb + i * b_step) : 0.0;
        }
        if constexpr (AMODE != 2) pa += a_tile;
        pb += b_tile;
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + GT * i;
            const int m = A_KC ? e / BK : e % BM;
            const int k = A_KC ? e % BK : e / BM;
            gsm[buf * STAGE + a_at(m, k)] = ra[i];
            gsm[buf * STAGE + A_ELEMS + (e / BN) * PB + (e % BN)] = rb[i];
        }
    };

    // MMA: warp tile 64 x 32 = 8 x 4 fragments of 8 x 8, two accumulators per lane and fragment.
    // SIMT: thread tile 8 x 8, rows ty*8 + i, columns tx + 16 j.
    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
    const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;
    const int gq = lane >> 2, tq = lane & 3;  // groupID, threadID_in_group of the PTX fragment layout
    const int ty = tid >> 4, tx = tid & 15;

    gload(0);
    sstore(0);
    __syncthreads();
    const int nkt = (K + BK - 1) / BK;
    for (int kt = 0; kt < nkt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nkt) gload((kt + 1) * BK);
        const double* a_s = gsm + buf * STAGE;
        const double* b_s = a_s + A_ELEMS;
        if constexpr (MMA) {
            // a warp whose 64 x 32 tile lies entirely outside C (ragged last tiles: 320 = 2 x 128 + 64) issues
            // no DMMA and leaves the tensor pipe of its SM sub-partition to the warp that shares it
            if (m0 + wm < M && n0 + wn < N) {
#pragma unroll
            for (int kk = 0; kk < BK; kk += 4) {
                double af[8], bf[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) af[i] = a_s[a_at(wm + i * 8 + gq, kk + tq)];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = b_s[(kk + tq) * PB + wn + j * 8 + gq];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][2 * j], acc[i][2 * j + 1], af[i], bf[j]);
            }
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                double av[8], bv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) av[i] = a_s[a_at(ty * 8 + i, kk)];
#pragma unroll
                for (int j = 0; j < 8; ++j) bv[j] = b_s[kk * PB + tx + 16 * j];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
            }
        }
        if (kt + 1 < nkt) sstore(buf ^ 1);
        __syncthreads();
    }

    if constexpr (MMA) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + wm + i * 8 + gq;
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int n = n0 + wn + j * 8 + 2 * tq + h;
                    if (n < N) C[(int64_t)m * g.c_sm + (int64_t)n * g.c_sn] = acc[i][2 * j + h];
                }
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + ty * 8 + i;
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = n0 + tx + 16 * j;
                if (n < N) C[(int64_t)m * g.c_sm + (int64_t)n * g.c_sn] = acc[i][j];
            }
        }
    }
}

// FP64 tensor-pipe roofline denominator: eight independent DMMA accumulator pairs per warp
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b) {
    double c[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) c[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) dmma884(c[2 * k], c[2 * k + 1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += c[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

bool use_simt() {
    const char* e = std::getenv("MYQC_AO2MO_GEMM");
    return e && e[0] == 's';
}

template <int AMODE, bool MMA>
int gemm_launch_t(const GemmArgs& g, int batch, cudaStream_t st) {
    static bool prepared[64] = {};
    int dev = 0;
    CUA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !prepared[dev]) {
        CUA(cudaFuncSetAttribute(gemm_f64_kernel<AMODE, MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
        prepared[dev] = true;
    }
    if (g.M <= 0 || g.N <= 0 || batch <= 0) return MYQC_OK;
    // gridDim.y / .z are limited to 65535: split the batch, fold large M into x by swapping roles is
    // not needed here (M/128 <= 65535 for every panel size this file chooses)
    const int gx = (g.N + BN - 1) / BN, gy = (g.M + BM - 1) / BM;
    if (gy > 65535) return myqc::fock_fail(MYQC_ERR_BAD_ARG, "ao2mo GEMM: too many row tiles");
    for (int b0 = 0; b0 < batch; b0 += 65535) {
        const int nb = std::min(65535, batch - b0);
        GemmArgs a = g;
        a.A += (int64_t)b0 * g.a_batch;
        a.B += (int64_t)b0 * g.b_batch;
        a.C += (int64_t)b0 * g.c_batch;
        gemm_f64_kernel<AMODE, MMA><<<dim3(gx, gy, nb), GT, GEMM_SMEM, st>>>(a);
    }
    CUA(cudaGetLastError());
    return MYQC_OK;
}

int gemm_launch(int amode, const GemmArgs& g, int batch, cudaStream_t st) {
    const bool simt = use_simt();
    if (amode == 2) return simt ? gemm_launch_t<2, false>(g, batch, st) : gemm_launch_t<2, true>(g, batch, st);
    if (amode == 1) return simt ? gemm_launch_t<1, false>(g, batch, st) : gemm_launch_t<1, true>(g, batch, st);
    return simt ? gemm_launch_t<0, false>(g, batch, st) : gemm_launch_t<0, true>(g, batch, st);
}

// ---- data movement kernels ---------------------------------------------------------------------
__device__ __forceinline__ int64_t tri_off(int64_t a, int64_t b, int64_t dim) {  // a <= b
    return a * dim - ((a * (a - 1)) >> 1) + (b - a);
}

// Row panel of the symmetric pair matrix: Xrow[pp][P'] = X[P0+pp, P'] for all NP columns.
// Columns P' >= P0 (the stored part of the rows, plus the small triangle inside the panel's own
// diagonal block): one pass along the rows, coalesced.     grid (BP, chunks of 1024 columns)
__global__ void __launch_bounds__(256) panel_upper_kernel(const double* __restrict__ packed, int64_t np, int64_t p0,
                                                          double* __restrict__ xrow) {
    const int64_t P = p0 + blockIdx.x;
    double* dst = xrow + (int64_t)blockIdx.x * np;
    const int64_t c0 = p0 + (int64_t)blockIdx.y * 1024;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t Pp = c0 + j * 256 + threadIdx.x;
        if (Pp < np) {
            const int64_t lo = P < Pp ? P : Pp, hi = P < Pp ? Pp : P;
            dst[Pp] = __ldg(packed + tri_off(lo, hi, np));
        }
    }
}

// Columns P' < P0 live in the rows P' of the packed array, at columns P0 .. P0+BP-1 (contiguous):
// 32 x 32 tiles transposed through shared memory, coalesced on both sides.  grid (BP/32, P0/32), block (32, 8)
__global__ void __launch_bounds__(256) panel_lower_kernel(const double* __restrict__ packed, int64_t np, int64_t p0, int bp,
                                                          double* __restrict__ xrow) {
    __shared__ double tile[32][33];
    const int pp0 = blockIdx.x * 32;
    const int64_t q0 = (int64_t)blockIdx.y * 32;  // first P' of the tile
#pragma unroll
    for (int y = threadIdx.y; y < 32; y += 8) {
        const int64_t Pp = q0 + y;
        const int pp = pp0 + threadIdx.x;
        if (Pp < p0 && pp < bp) tile[y][threadIdx.x] = __ldg(packed + tri_off(Pp, p0 + pp, np));
    }
    __syncthreads();
#pragma unroll
    for (int y = threadIdx.y; y < 32; y += 8) {
        const int pp = pp0 + y;
        const int64_t Pp = q0 + threadIdx.x;
        if (Pp < p0 && pp < bp) xrow[(int64_t)pp * np + Pp] = tile[threadIdx.x][y];
    }
}

// Gsq[u][v][x] = H[pair(u,v)][rs0 + x], x < w   grid (n, n), threads over x
__global__ void unpack_cols_kernel(const double* __restrict__ h, int n, int64_t nrs, int64_t rs0, int w, double* __restrict__ gsq) {
    const int u = blockIdx.x, v = blockIdx.y;
    const int a = u < v ? u : v, b = u < v ? v : u;
    const double* src = h + tri_off(a, b, n) * nrs + rs0;
    double* dst = gsq + ((int64_t)u * n + v) * w;
    for (int x = threadIdx.x; x < w; x += blockDim.x) dst[x] = __ldg(src + x);
}

// Cr[u][p] = C(u,p)  (column-major block with leading dimension n -> row-major n x nk)
__global__ void coef_rowmajor_kernel(const double* __restrict__ c, int n, int nk, double* __restrict__ cr) {
    const int64_t total = (int64_t)n * nk;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int u = (int)(e / nk), p = (int)(e % nk);
        cr[e] = c[u + (int64_t)n * p];
    }
}

// MYQC_AO2MO_TRACE=1: CUDA-event time of every stage, summed over panels, on stderr (synchronises)
struct StageTrace {
    bool on = false;
    cudaStream_t st = nullptr;
    std::vector<cudaEvent_t> ev;  // pairs
    std::vector<int> stage;
    void begin(int s) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev.push_back(e);
        stage.push_back(s);
    }
    void end() {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev.push_back(e);
    }
    void report(const double* flops) {
        if (!on) return;
        cudaStreamSynchronize(st);
        static const char* names[6] = {"row panel", "GEMM 1a", "GEMM 1b", "unpack cols", "GEMM 2a", "GEMM 2b"};
        double ms[6] = {0, 0, 0, 0, 0, 0};
        for (size_t k = 0; k < stage.size(); ++k) {
            float t = 0;
            cudaEventElapsedTime(&t, ev[2 * k], ev[2 * k + 1]);
            ms[stage[k]] += t;
        }
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        for (int k = 0; k < 6; ++k) {
            if (flops[k] > 0)
                std::fprintf(stderr, "[myqc ao2mo trace] %-12s %9.3f ms  %6.2f TFLOP/s\n", names[k], ms[k],
                             flops[k] / (ms[k] * 1e-3 + 1e-30) / 1e12);
            else
                std::fprintf(stderr, "[myqc ao2mo trace] %-12s %9.3f ms\n", names[k], ms[k]);
        }
    }
};

int64_t env_i64(const char* name, int64_t dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoll(e) : dflt;
}

// Scratch layout of one transformation (doubles): [panel | T/V | H | C2r C3r C4r]
struct Sizes {
    int64_t n, np, nrs, bp, rsb, panel, t, h, c;
    int64_t total() const { return panel + t + h + c; }
};

Sizes plan_sizes(int norb, int n1, int n2, int n3, int n4) {
    Sizes z{};
    z.n = norb; z.np = z.n * (z.n + 1) / 2; z.nrs = (int64_t)n3 * n4;
    // panel sizes: about MYQC_AO2MO_SCRATCH_MB (default 1536 MB) per panel
    const int64_t budget = std::min<int64_t>(env_i64("MYQC_AO2MO_SCRATCH_MB", 1536) * 1000000 / 8, (int64_t)2000000000);  // doubles, < 2^31
    z.bp = std::max<int64_t>(1, std::min<int64_t>(z.np, budget / z.np));          // rows of Xrow[bp][np]
    z.bp = std::min<int64_t>(z.bp, (int64_t)65535 * BM / z.n);                     // row tiles of GEMM 1a
    z.rsb = std::max<int64_t>(1, std::min<int64_t>(z.nrs, budget / (z.n * z.n)));  // columns of Gsq[n][n][rsb]
    if (z.rsb < z.nrs && z.rsb > BM) z.rsb -= z.rsb % BM;                          // whole row tiles of GEMM 2a
    z.panel = std::max(z.bp * z.np, z.rsb * z.n * z.n);
    z.t = std::max(z.bp * z.n * n4, z.n * z.rsb * n2);
    z.h = z.np * z.nrs;
    z.c = z.n * ((int64_t)n2 + n3 + n4);
    (void)n1;
    return z;
}

int transform_device(const double* d_packed, int norb, const double* d_c1, int n1, const double* d_c2, int n2,
                     const double* d_c3, int n3, const double* d_c4, int n4, double* d_out, cudaStream_t st,
                     double* ws = nullptr, int64_t ws_bytes = 0) {
    if (!d_packed || !d_c1 || !d_c2 || !d_c3 || !d_c4 || !d_out) return myqc::fock_fail(MYQC_ERR_BAD_ARG, "ao2mo: null pointer");
    if (norb < 1 || n1 < 1 || n2 < 1 || n3 < 1 || n4 < 1 || n1 > norb || n2 > norb || n3 > norb || n4 > norb)
        return myqc::fock_fail(MYQC_ERR_BAD_ARG, "ao2mo: block sizes must be in 1..norb");
    const Sizes z = plan_sizes(norb, n1, n2, n3, n4);
    const int64_t n = z.n, np = z.np, nrs = z.nrs, bp = z.bp, rsb = z.rsb;
    double* buf = ws;
    const size_t bytes = sizeof(double) * (size_t)z.total();
    if (ws) {
        if ((size_t)ws_bytes < bytes) return myqc::fock_fail(MYQC_ERR_BAD_ARG, "ao2mo: workspace smaller than myqc_ao2mo_workspace_bytes()");
    } else {
        cudaError_t e = cudaMallocAsync((void**)&buf, bytes, st);
        if (e != cudaSuccess)
            return myqc::fock_fail(MYQC_ERR_NOMEM, "ao2mo scratch (" + std::to_string(bytes >> 20) + " MiB): " + cudaGetErrorString(e));
    }
    double* panel = buf;
    double* tbuf = panel + z.panel;
    double* h = tbuf + z.t;
    double* c2r = h + z.h;
    double* c3r = c2r + n * n2;
    double* c4r = c3r + n * n3;
    coef_rowmajor_kernel<<<64, 256, 0, st>>>(d_c2, norb, n2, c2r);
    coef_rowmajor_kernel<<<64, 256, 0, st>>>(d_c3, norb, n3, c3r);
    coef_rowmajor_kernel<<<64, 256, 0, st>>>(d_c4, norb, n4, c4r);
    int rc = MYQC_OK;
    StageTrace tr;
    tr.on = std::getenv("MYQC_AO2MO_TRACE") != nullptr;
    tr.st = st;

    // ---- stage 1: ket half transformation, H[P][s][r] -------------------------------------------
    for (int64_t p0 = 0; p0 < np && !rc; p0 += bp) {
        const int64_t cur = std::min(bp, np - p0);
        tr.begin(0);
        panel_upper_kernel<<<dim3((unsigned)cur, (unsigned)((np - p0 + 1023) / 1024)), 256, 0, st>>>(d_packed, np, p0, panel);
        if (p0 > 0)
            panel_lower_kernel<<<dim3((unsigned)((cur + 31) / 32), (unsigned)((p0 + 31) / 32)), dim3(32, 8), 0, st>>>(d_packed, np, p0, (int)cur, panel);
        tr.end();
        GemmArgs a{};
        a.A = panel; a.a_sm = 0; a.a_sk = 0; a.a_batch = 0; a.g_n = norb; a.g_np = (int)np;
        a.B = c4r; a.ldb = n4; a.b_batch = 0;
        a.C = tbuf; a.c_sm = n4; a.c_sn = 1; a.c_batch = 0;
        a.M = (int)(cur * n); a.N = n4; a.K = norb;
        tr.begin(1);
        rc = gemm_launch(2, a, 1, st);
        tr.end();
        if (rc) break;
        GemmArgs b{};
        b.A = tbuf; b.a_sm = 1; b.a_sk = n4; b.a_batch = n * n4;
        b.B = c3r; b.ldb = n3; b.b_batch = 0;
        b.C = h + p0 * nrs; b.c_sm = n3; b.c_sn = 1; b.c_batch = nrs;
        b.M = n4; b.N = n3; b.K = norb;
        tr.begin(2);
        rc = gemm_launch(0, b, (int)cur, st);
        tr.end();
    }
    // ---- stage 2: bra half transformation, written into Om(p,q,r,s) -----------------------------
    for (int64_t rs0 = 0; rs0 < nrs && !rc; rs0 += rsb) {
        const int64_t w = std::min(rsb, nrs - rs0);
        tr.begin(3);
        unpack_cols_kernel<<<dim3((unsigned)n, (unsigned)n), 128, 0, st>>>(h, norb, nrs, rs0, (int)w, panel);
        tr.end();
        GemmArgs a{};
        a.A = panel; a.a_sm = 1; a.a_sk = w; a.a_batch = n * w;
        a.B = c2r; a.ldb = n2; a.b_batch = 0;
        a.C = tbuf; a.c_sm = n2; a.c_sn = 1; a.c_batch = w * n2;
        a.M = (int)w; a.N = n2; a.K = norb;
        tr.begin(4);
        rc = gemm_launch(0, a, norb, st);
        tr.end();
        if (rc) break;
        GemmArgs b{};
        b.A = d_c1; b.a_sm = n; b.a_sk = 1; b.a_batch = 0;
        b.B = tbuf; b.ldb = w * n2; b.b_batch = 0;
        b.C = d_out + rs0 * n2 * n1; b.c_sm = 1; b.c_sn = n1; b.c_batch = 0;
        b.M = n1; b.N = (int)(w * n2); b.K = norb;
        tr.begin(5);
        rc = gemm_launch(1, b, 1, st);
        tr.end();
    }
    {
        const double dn = (double)n, dnp = (double)np;
        const double fl[6] = {0.0, 2.0 * dnp * dn * dn * n4, 2.0 * dnp * dn * n3 * n4, 0.0,
                              2.0 * dn * dn * n2 * (double)nrs, 2.0 * dn * n1 * n2 * (double)nrs};
        tr.report(fl);
    }
    cudaError_t le = cudaGetLastError();
    if (!ws) cudaFreeAsync(buf, st);
    if (rc) return rc;
    if (le != cudaSuccess) return myqc::fock_fail(MYQC_ERR_CUDA, std::string("ao2mo launch: ") + cudaGetErrorString(le));
    return MYQC_OK;
}

// ---- file layer of PROGRAM ao2mo ---------------------------------------------------------------
std::string joinp(const char* dir, const char* name) {
    std::string d = (dir && *dir) ? dir : ".";
    if (d.back() != '/') d += '/';
    return d + name;
}

bool read_reals(const std::string& path, std::vector<double>& v) {  // list-directed READ(u,*)
    std::ifstream f(path);
    if (!f) return false;
    std::string line;
    while (std::getline(f, line)) {
        for (char& c : line)
            if (c == ',') c = ' ';
        std::istringstream ss(line);
        std::string t;
        while (ss >> t) {
            int rep = 1;
            const size_t star = t.find('*');
            if (star != std::string::npos && star > 0) {
                rep = std::atoi(t.substr(0, star).c_str());
                t = t.substr(star + 1);
            }
            for (char& c : t)
                if (c == 'D' || c == 'd') c = 'E';
            char* end = nullptr;
            const double x = std::strtod(t.c_str(), &end);
            if (end == t.c_str()) return true;  // first non-number ends the data (comment lines)
            for (int k = 0; k < rep; ++k) v.push_back(x);
        }
    }
    return true;
}

bool write_record(FILE* f, const double* v, size_t n) {  // WRITE(u) v(0:n-1), unformatted sequential
    const int32_t len = (int32_t)(n * sizeof(double));
    return std::fwrite(&len, 4, 1, f) == 1 && (n == 0 || std::fwrite(v, sizeof(double), n, f) == n) &&
           std::fwrite(&len, 4, 1, f) == 1;
}

void touch_err(const char* dir) {
    FILE* f = std::fopen(joinp(dir, "error").c_str(), "a");
    if (f) std::fclose(f);
}

struct Job {
    const char* dir;
    int n;
    double* d_packed;
    double* d_ca;  // device copies of CmA / CmB (column-major n x n)
    double* d_cb;
    double* d_out;
    std::vector<double> host;
};

// Om(0:n1-1,0:n2-1,0:n3-1,0:n4-1) of the column blocks [o_k, o_k + n_k) of the given spin matrices
int run_transform(Job& j, const double* c1, int o1, int n1, const double* c2, int o2, int n2, const double* c3, int o3, int n3,
                  const double* c4, int o4, int n4) {
    const size_t total = (size_t)n1 * n2 * n3 * n4;
    j.host.assign(total, 0.0);
    if (total == 0) return MYQC_OK;
    const int64_t n = j.n;
    int rc = transform_device(j.d_packed, j.n, c1 + n * o1, n1, c2 + n * o2, n2, c3 + n * o3, n3, c4 + n * o4, n4, j.d_out, nullptr);
    if (rc) return rc;
    CUA(cudaMemcpy(j.host.data(), j.d_out, total * sizeof(double), cudaMemcpyDeviceToHost));
    return MYQC_OK;
}

// WRITE(u) Om(i,:,j,:) for the listed (i,j): ao2mo.f90:567-581,719-725,805-811,891-897
int write_ijab(Job& jb, const char* name, int n1, int n2, int n3, int n4, bool upper_only) {
    FILE* f = std::fopen(joinp(jb.dir, name).c_str(), "wb");
    if (!f) return myqc::fock_fail(MYQC_ERR_IO, std::string("cannot write ") + name);
    std::vector<double> rec((size_t)n2 * n4);
    bool ok = true;
    for (int i = 0; i < (upper_only ? n1 - 1 : n1) && ok; ++i)
        for (int j = (upper_only ? i + 1 : 0); j < n3 && ok; ++j) {
            for (int b = 0; b < n4; ++b)
                for (int a = 0; a < n2; ++a)
                    rec[a + (size_t)n2 * b] = jb.host[i + (size_t)n1 * (a + (size_t)n2 * (j + (size_t)n3 * b))];
            ok = write_record(f, rec.data(), rec.size());
        }
    std::fclose(f);
    return ok ? MYQC_OK : myqc::fock_fail(MYQC_ERR_IO, std::string("short write to ") + name);
}

// CIS records (ao2mo.f90:975-989,1044-1058,...): DO j, DO b: vec(i*nv + a) = Om(a,i,j,b) [ajib]
// or Om(a,b,j,i) [ajbi]; nrec_j x nrec_b records of no*nv values.
int write_cis(Job& jb, const char* name, bool ajbi, int nv, int no, int nj, int nb) {
    FILE* f = std::fopen(joinp(jb.dir, name).c_str(), "wb");
    if (!f) return myqc::fock_fail(MYQC_ERR_IO, std::string("cannot write ") + name);
    std::vector<double> vec((size_t)no * nv);
    bool ok = true;
    for (int j = 0; j < nj && ok; ++j)
        for (int b = 0; b < nb && ok; ++b) {
            size_t idx = 0;
            for (int i = 0; i < no; ++i)
                for (int a = 0; a < nv; ++a) {
                    // ajib: Om(0:nv-1,0:no-1,0:nj-1,0:nb-1)(a,i,j,b);  ajbi: Om(0:nv-1,0:nb-1,0:nj-1,0:no-1)(a,b,j,i)
                    vec[idx++] = ajbi ? jb.host[a + (size_t)nv * (b + (size_t)nb * (j + (size_t)nj * i))]
                                      : jb.host[a + (size_t)nv * (i + (size_t)no * (j + (size_t)nj * b))];
                }
            ok = write_record(f, vec.data(), vec.size());
        }
    std::fclose(f);
    return ok ? MYQC_OK : myqc::fock_fail(MYQC_ERR_IO, std::string("short write to ") + name);
}

}  // namespace

extern "C" {

int myqc_dmma_peak(int device, double* tflops) {
    if (!tflops) return myqc::fock_fail(MYQC_ERR_BAD_ARG, "null result");
    if (myqc_device_count() == 0) return myqc::fock_fail(MYQC_ERR_NO_DEVICE, "no CUDA device");
    CUA(cudaSetDevice(device));
    int sms = 0;
    CUA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int grid = sms * 4, iters = 2048;
    double* d = nullptr;
    CUA(cudaMalloc((void**)&d, (size_t)grid * 256 * sizeof(double)));
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    double best = 0.0;
    cudaError_t e = cudaSuccess;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(t0);
        dmma_peak_kernel<<<grid, 256>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(t1);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, t0, t1);
        // one m8n8k4 DMMA = 8*8*4 FMA = 512 flop per warp
        const double flops = 512.0 * 32 * (double)iters * 8.0 * grid;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    cudaFree(d);
    *tflops = best;
    if (e != cudaSuccess) return myqc::fock_fail(MYQC_ERR_CUDA, cudaGetErrorString(e));
    return MYQC_OK;
}

double myqc_ao2mo_flops(int norb, int n1, int n2, int n3, int n4) {
    const double n = norb, np = n * (n + 1) / 2;
    return 2.0 * np * n * n * n4 + 2.0 * np * n * n3 * n4 + 2.0 * n * n * n2 * n3 * n4 + 2.0 * n * n1 * n2 * n3 * n4;
}

int myqc_ao2mo_transform(const double* d_packed, int norb, const double* d_c1, int n1, const double* d_c2, int n2,
                         const double* d_c3, int n3, const double* d_c4, int n4, double* d_out, void* stream) {
    if (myqc_device_count() == 0) return myqc::fock_fail(MYQC_ERR_NO_DEVICE, "no CUDA device: ao2mo has no CPU fallback");
    return transform_device(d_packed, norb, d_c1, n1, d_c2, n2, d_c3, n3, d_c4, n4, d_out, static_cast<cudaStream_t>(stream));
}

int64_t myqc_ao2mo_workspace_bytes(int norb, int n1, int n2, int n3, int n4) {
    if (norb < 1 || n1 < 1 || n2 < 1 || n3 < 1 || n4 < 1) return 0;
    return (int64_t)sizeof(double) * plan_sizes(norb, n1, n2, n3, n4).total();
}

int myqc_ao2mo_transform_ws(const double* d_packed, int norb, const double* d_c1, int n1, const double* d_c2, int n2,
                            const double* d_c3, int n3, const double* d_c4, int n4, double* d_out, void* d_workspace,
                            int64_t workspace_bytes, void* stream) {
    if (myqc_device_count() == 0) return myqc::fock_fail(MYQC_ERR_NO_DEVICE, "no CUDA device: ao2mo has no CPU fallback");
    if (!d_workspace) return myqc::fock_fail(MYQC_ERR_BAD_ARG, "ao2mo: null workspace");
    return transform_device(d_packed, norb, d_c1, n1, d_c2, n2, d_c3, n3, d_c4, n4, d_out, static_cast<cudaStream_t>(stream),
                            static_cast<double*>(d_workspace), workspace_bytes);
}

int myqc_ao2mo_transform_host(const double* packed, int norb, const double* c1, int n1, const double* c2, int n2,
                              const double* c3, int n3, const double* c4, int n4, double* out) {
    if (!packed || !c1 || !c2 || !c3 || !c4 || !out || norb < 1) return myqc::fock_fail(MYQC_ERR_BAD_ARG, "ao2mo: null pointer or bad norb");
    if (n1 < 1 || n2 < 1 || n3 < 1 || n4 < 1 || n1 > norb || n2 > norb || n3 > norb || n4 > norb)
        return myqc::fock_fail(MYQC_ERR_BAD_ARG, "ao2mo: block sizes must be in 1..norb");
    if (myqc_device_count() == 0) return myqc::fock_fail(MYQC_ERR_NO_DEVICE, "no CUDA device: ao2mo has no CPU fallback");
    const int64_t n = norb, np = n * (n + 1) / 2, total = np * (np + 1) / 2;
    const size_t nout = (size_t)n1 * n2 * n3 * n4;
    double *d_p = nullptr, *d_c = nullptr, *d_o = nullptr;
    CUA(cudaMalloc((void**)&d_p, sizeof(double) * total));
    cudaError_t e = cudaMalloc((void**)&d_c, sizeof(double) * n * (n1 + n2 + n3 + n4));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_o, sizeof(double) * nout);
    if (e != cudaSuccess) { cudaFree(d_p); cudaFree(d_c); return myqc::fock_fail(MYQC_ERR_NOMEM, cudaGetErrorString(e)); }
    double* dc1 = d_c; double* dc2 = dc1 + n * n1; double* dc3 = dc2 + n * n2; double* dc4 = dc3 + n * n3;
    e = cudaMemcpy(d_p, packed, sizeof(double) * total, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dc1, c1, sizeof(double) * n * n1, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dc2, c2, sizeof(double) * n * n2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dc3, c3, sizeof(double) * n * n3, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dc4, c4, sizeof(double) * n * n4, cudaMemcpyHostToDevice);
    int rc = e == cudaSuccess ? MYQC_OK : myqc::fock_fail(MYQC_ERR_CUDA, cudaGetErrorString(e));
    if (!rc) rc = transform_device(d_p, norb, dc1, n1, dc2, n2, dc3, n3, dc4, n4, d_o, nullptr);
    if (!rc) {
        e = cudaMemcpy(out, d_o, sizeof(double) * nout, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = myqc::fock_fail(MYQC_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFree(d_p); cudaFree(d_c); cudaFree(d_o);
    return rc;
}

int myqc_pack_dense(const double* xx, int norb, double* packed) {
    if (!xx || !packed || norb < 1) return myqc::fock_fail(MYQC_ERR_BAD_ARG, "pack_dense: bad arguments");
    const int64_t n = norb, np = n * (n + 1) / 2;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = i; j < n; ++j) {
            const int64_t P = i * n - i * (i - 1) / 2 + (j - i);
            double* row = packed + (P * np - P * (P - 1) / 2) - P;  // row[P'] for P' >= P
            for (int64_t k = 0; k < n; ++k)
                for (int64_t l = k; l < n; ++l) {
                    const int64_t Pp = k * n - k * (k - 1) / 2 + (l - k);
                    if (Pp >= P) row[Pp] = xx[i + n * (j + n * (k + n * l))];
                }
        }
    return MYQC_OK;
}

static int ao2mo_main_impl(const char* dir);
int myqc_ao2mo_main(const char* dir) {
    const int rc = ao2mo_main_impl(dir);
    std::fflush(stdout);  // see myqc_parse_main
    return rc;
}
static int ao2mo_main_impl(const char* dir) {
    // banner: ao2mo.f90:39-45
    std::printf("\n                 STARTING AO TRANSFORM\n ------------------------------------------------------------\n"
                " ao2mo called\n\n Starting AO to MO integral transform\n");
    auto bail = [&](int rc, const char* what) {
        std::printf(" ao2mo: %s: %s\n", what, myqc_last_error());
        touch_err(dir);
        return rc;
    };
    int nnuc = 0, nA = 0, nB = 0, nopt = 0;
    double fmem = 0;
    int rc = myqc_read_env(dir, 0, 0, &nnuc, &nA, &nB, nullptr, nullptr, &fmem, &nopt, nullptr);
    if (rc) return bail(rc, "getenv");
    std::vector<int32_t> atoms(nnuc), options(nopt > 17 ? nopt : 17, 0);
    std::vector<double> xyz((size_t)3 * nnuc);
    rc = myqc_read_env(dir, nnuc, (int)options.size(), &nnuc, &nA, &nB, atoms.data(), xyz.data(), &fmem, &nopt, options.data());
    if (rc) return bail(rc, "getenv");
    {
        FILE* f = std::fopen(joinp(dir, "error").c_str(), "rb");  // INQUIRE(file='error'); IF (flag) STOP  (:49-50)
        if (f) { std::fclose(f); return MYQC_OK; }
    }
    std::vector<double> bi;
    if (!read_reals(joinp(dir, "basinfo"), bi) || bi.size() < 2) {
        myqc::fock_fail(MYQC_ERR_IO, "cannot read basinfo");
        return bail(MYQC_ERR_IO, "basinfo");
    }
    const int n = (int)bi[1];  // ntot = line(1), :53-55
    const int noccA = nA, noccB = nB, nvrtA = n - nA, nvrtB = n - nB;
    const bool mp2 = options[1] == 1, uhf = options[3] == 1, rhf = options[3] == 0;
    const bool cis = options[13] == 1 && options[1] == 0;
    // the reference's dispatch and its messages (:62-92)
    if (mp2 && !(rhf || uhf)) {
        std::printf(" Sorry, that reference not coded yet\n"); touch_err(dir); return MYQC_ERR_UNSUPPORTED;
    }
    if (!mp2 && cis && !uhf) {
        std::printf(" Sorry, only UHF CIS is coded\n"); touch_err(dir); return MYQC_ERR_UNSUPPORTED;
    }
    if (!mp2 && !cis) {
        std::printf(" Sorry, that transform type has not been coded yet\n"); touch_err(dir); return MYQC_ERR_UNSUPPORTED;
    }
    if (myqc_device_count() == 0) {
        myqc::fock_fail(MYQC_ERR_NO_DEVICE, "no CUDA device: ao2mo has no CPU fallback");
        return bail(MYQC_ERR_NO_DEVICE, "device");
    }
    if (noccA < 0 || noccB < 0 || nvrtA < 0 || nvrtB < 0 || n < 1) {
        myqc::fock_fail(MYQC_ERR_BAD_ARG, "inconsistent electron / orbital counts");
        return bail(MYQC_ERR_BAD_ARG, "envdat");
    }

    const int64_t nn = (int64_t)n * n, np = (int64_t)n * (n + 1) / 2, total = np * (np + 1) / 2;
    std::vector<double> cui;
    if (!read_reals(joinp(dir, "Cui"), cui) || (int64_t)cui.size() < (rhf && mp2 ? nn : 2 * nn)) {
        myqc::fock_fail(MYQC_ERR_IO, "Cui does not hold the MO coefficients");
        return bail(MYQC_ERR_IO, "Cui");
    }
    Job jb{};
    jb.dir = dir; jb.n = n;
    {
        std::vector<double> xx((size_t)(nn * nn)), packed((size_t)total);
        rc = myqc_read_xx(joinp(dir, "XX").c_str(), xx.data(), n);  // READ(100) Km(:,:,:,:)
        if (rc) return bail(rc, "XX");
        myqc_pack_dense(xx.data(), n, packed.data());
        cudaError_t e = cudaMalloc((void**)&jb.d_packed, sizeof(double) * total);
        if (e == cudaSuccess) e = cudaMalloc((void**)&jb.d_ca, sizeof(double) * 2 * nn);
        if (e == cudaSuccess) e = cudaMalloc((void**)&jb.d_out, sizeof(double) * (size_t)(nn * nn));
        if (e == cudaSuccess) e = cudaMemcpy(jb.d_packed, packed.data(), sizeof(double) * total, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(jb.d_ca, cui.data(), sizeof(double) * nn, cudaMemcpyHostToDevice);
        jb.d_cb = jb.d_ca + nn;
        if (e == cudaSuccess)
            e = cudaMemcpy(jb.d_cb, cui.data() + ((int64_t)cui.size() >= 2 * nn ? nn : 0), sizeof(double) * nn, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            myqc::fock_fail(MYQC_ERR_CUDA, cudaGetErrorString(e));
            cudaFree(jb.d_packed); cudaFree(jb.d_ca); cudaFree(jb.d_out);
            return bail(MYQC_ERR_CUDA, "device setup");
        }
    }
    const double *CA = jb.d_ca, *CB = jb.d_cb;
    if (mp2 && rhf) {  // slow_ao2mo_MP2_RHF, :465-602
        std::printf("\n Spin Case AB\n Transforming (uv|ld) -> <ij|ab>\n");
        rc = run_transform(jb, CA, 0, noccA, CA, noccA, nvrtA, CA, 0, noccA, CA, noccA, nvrtA);
        if (!rc) { std::printf(" Writing to ijab_AA\n"); rc = write_ijab(jb, "ijab_AA", noccA, nvrtA, noccA, nvrtA, true); }
        if (!rc) { std::printf(" Writing to ijab_AB\n"); rc = write_ijab(jb, "ijab_AB", noccA, nvrtA, noccA, nvrtA, false); }
    } else if (mp2) {  // slow_ao2mo_MP2_UHF, :614-904
        std::printf(" Spin Case AA\n");
        rc = run_transform(jb, CA, 0, noccA, CA, noccA, nvrtA, CA, 0, noccA, CA, noccA, nvrtA);
        if (!rc) { std::printf(" Writing to ijab_AA\n"); rc = write_ijab(jb, "ijab_AA", noccA, nvrtA, noccA, nvrtA, true); }
        if (!rc) {
            std::printf("\n Spin Case BB\n");
            rc = run_transform(jb, CB, 0, noccB, CB, noccB, nvrtB, CB, 0, noccB, CB, noccB, nvrtB);
        }
        if (!rc) { std::printf(" Writing to ijab_BB\n"); rc = write_ijab(jb, "ijab_BB", noccB, nvrtB, noccB, nvrtB, true); }
        if (!rc) {
            std::printf("\n Spin Case AB\n");
            rc = run_transform(jb, CA, 0, noccA, CA, noccA, nvrtA, CB, 0, noccB, CB, noccB, nvrtB);
        }
        if (!rc) { std::printf(" Writing to ijab_AB\n"); rc = write_ijab(jb, "ijab_AB", noccA, nvrtA, noccB, nvrtB, false); }
    } else {  // slow_ao2mo_CIS_UHF, :919-1227
        std::printf(" Spin Case AA\n Transforming (uv|ld) -> <aj|ib>\n");
        rc = run_transform(jb, CA, noccA, nvrtA, CA, 0, noccA, CA, 0, noccA, CA, noccA, nvrtA);
        if (!rc) { std::printf(" Writing to ajib_AA\n"); rc = write_cis(jb, "ajib_AA", false, nvrtA, noccA, noccA, nvrtA); }
        if (!rc) {
            std::printf(" Transforming (uv|ld) -> <aj|bi>\n");
            rc = run_transform(jb, CA, noccA, nvrtA, CA, noccA, nvrtA, CA, 0, noccA, CA, 0, noccA);
        }
        if (!rc) { std::printf(" Writing to ajbi_AA\n"); rc = write_cis(jb, "ajbi_AA", true, nvrtA, noccA, noccA, nvrtA); }
        if (!rc) {
            std::printf(" Spin Case AB\n Transforming (uv|ld) -> <aj|ib>\n");
            rc = run_transform(jb, CA, noccA, nvrtA, CA, 0, noccA, CB, 0, noccB, CB, noccB, nvrtB);
        }
        if (!rc) { std::printf(" Writing to ajib_AB\n"); rc = write_cis(jb, "ajib_AB", false, nvrtA, noccA, noccB, nvrtB); }
        if (!rc) {
            std::printf(" Spin case BB\n Transforming (uv|ld) -> <aj|ib>\n");
            rc = run_transform(jb, CB, noccB, nvrtB, CB, 0, noccB, CB, 0, noccB, CB, noccB, nvrtB);
        }
        if (!rc) { std::printf(" Writing to ajib_BB\n"); rc = write_cis(jb, "ajib_BB", false, nvrtB, noccB, noccB, nvrtB); }
        if (!rc) {
            std::printf(" Transforming (uv|ld) -> <aj|bi>\n");
            rc = run_transform(jb, CB, noccB, nvrtB, CB, noccB, nvrtB, CB, 0, noccB, CB, 0, noccB);
        }
        if (!rc) { std::printf(" Writing to ajbi_BB\n"); rc = write_cis(jb, "ajbi_BB", true, nvrtB, noccB, noccB, nvrtB); }
    }
    cudaFree(jb.d_packed); cudaFree(jb.d_ca); cudaFree(jb.d_out);
    if (rc) return bail(rc, "transform");
    return MYQC_OK;
}

}  // extern "C"
