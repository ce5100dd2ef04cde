// G(D) from the packed unique ERI array, on the device (include/myqc_fock.h).
//
// Reference: src/I2G/RHFI2G.f90:80-90 and src/I2G/UHFI2G.f90:80-93 -- n^4 scalar iterations over the
// dense XX array, once per SCF iteration.  Here: one streaming pass over the packed upper triangle
// (n^4/8 elements, HBM-read bound; exact zeros -- 90 % of a large molecule -- cost a load and a
// compare).  For a unique integral V = (ij|kl), i<=j, k<=l, P(i,j) <= P(k,l), the eight index
// images contribute
//     J_ij += D_kl V   (both orders of k,l)      J_kl += D_ij V
//     K_ik += D_jl w   K_jk += D_il w   K_il += D_jk w   K_jl += D_ik w    (+ transposes)
// with w = V / (s_ij s_kl s_PP'), s = 2 where the two indices coincide, so that every distinct
// ordered index quadruple is counted once.
//
// Mapping: one CTA per packed row P = (i,j) (rows pulled from a global counter, longest first).
// Rows i and j of the density live in shared memory.  Warp w takes the column blocks k = i+w,
// i+w+NW, ...; its lanes run over l (contiguous in memory), so
//   * K_ik, K_jk are lane-private sums reduced with shuffles once per k,
//   * K_il, K_jl go to a warp-private shared-memory row (lanes hold distinct l: no atomics),
//   * J_ij is a thread-private sum reduced once per row,
//   * J_kl is the only global reduction per nonzero integral (RED.ADD.F64 to consecutive addresses).
// The per-row K rows are flushed with 2n global reductions.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/myqc_eri.h"
#include "../../include/myqc_fock.h"

namespace myqc {
int fock_fail(int code, const std::string& msg);  // sets myqc_last_error (eri_api.cu)
}

namespace {

struct FockArgs {
    const double* packed;  // slice of the packed array
    int64_t out_offset;    // packed index of packed[0]
    int64_t row_lo, row_hi;
    int n;
    int64_t npair;
    const int2* ij;     // [npair] pair index -> (i,j), i <= j
    const double* dp2;  // [npair] J density, off-diagonal pairs doubled
    const double* dk0;  // [n*n] symmetric K density (RHF: D, UHF: Da)
    const double* dk1;  // [n*n] UHF: Db
    double* jp;         // [npair] J accumulator (packed, upper triangle)
    double* k0;         // [n*n] K accumulator (first four images; the finalize kernel adds the transpose)
    double* k1;
    int* counter;
    const uint32_t* mask;  // optional [npair][mask_words]: bit k of row P set iff (P | k, .) holds a nonzero
    int mask_words;
};

__global__ void pair_table_kernel(int n, int2* ij) {
    const int i = blockIdx.x;
    const int64_t base = (int64_t)i * n - (int64_t)i * (i - 1) / 2;
    for (int j = i + threadIdx.x; j < n; j += blockDim.x) ij[base + (j - i)] = make_int2(i, j);
}

// dk = (D + D^T)/2 ; dp2[P(i,j)] += (i==j ? 1 : 2) * dk(i,j)   (dp2 zeroed before; UHF adds both spins)
__global__ void density_prep_kernel(int n, const double* d, double* dk, double* dp2) {
    const int i = blockIdx.x;
    const int64_t base = (int64_t)i * n - (int64_t)i * (i - 1) / 2;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double s = 0.5 * (d[i + (size_t)n * j] + d[j + (size_t)n * i]);
        dk[i + (size_t)n * j] = s;
        if (j >= i) atomicAdd(dp2 + base + (j - i), (i == j ? 1.0 : 2.0) * s);
    }
}

template <int NSPIN, int NW>
__global__ void __launch_bounds__(NW * 32) fock_rows_kernel(const FockArgs a) {
    constexpr int kThreads = NW * 32;
    extern __shared__ __align__(16) double sm[];
    __shared__ int s_row;
    __shared__ double s_red[NW];
    const int n = a.n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARPS = NW;
    // sD[s][r][x], r = 0: row i, 1: row j ;  sK[s][w][r][x]
    double* sD = sm;
    double* sK = sm + (size_t)NSPIN * 2 * n;
    const int nk = NSPIN * NWARPS * 2 * n;
    // the warp-private K rows are zeroed once; the flush at the end of a row clears what it reads
    for (int x = tid; x < nk; x += kThreads) sK[x] = 0.0;

    for (;;) {
        if (tid == 0) {
            // with a mask, rows without a single nonzero integral are skipped by the claiming thread
            int r;
            for (;;) {
                r = atomicAdd(a.counter, 1);
                if (!a.mask || a.row_lo + r >= a.row_hi) break;
                const uint32_t* mw = a.mask + (a.row_lo + r) * a.mask_words;
                uint32_t bits = 0;
                for (int w = 0; w < a.mask_words; ++w) bits |= __ldg(mw + w);
                if (bits) break;
            }
            s_row = r;
        }
        __syncthreads();
        const int64_t P = a.row_lo + s_row;
        if (P >= a.row_hi) break;
        const uint32_t* mrow = a.mask ? a.mask + P * a.mask_words : nullptr;
        const int2 pij = a.ij[P];
        const int i = pij.x, j = pij.y;
        for (int x = tid; x < n; x += kThreads) {
            sD[x] = a.dk0[i + (size_t)n * x];
            sD[n + x] = a.dk0[j + (size_t)n * x];
            if (NSPIN == 2) {
                sD[2 * n + x] = a.dk1[i + (size_t)n * x];
                sD[3 * n + x] = a.dk1[j + (size_t)n * x];
            }
        }
        __syncthreads();
        const double dp2_P = a.dp2[P];
        const double facP = (i == j) ? 0.5 : 1.0;
        const double* row = a.packed + ((P * a.npair - ((P * (P - 1)) >> 1)) - a.out_offset) - P;  // row[P'] = (P|P')
        double jacc = 0.0;
        bool any = false;
        double* myK0 = sK + (size_t)(0 * NWARPS + warp) * 2 * n;
        double* myK1 = sK + (size_t)(1 * NWARPS + warp) * 2 * n;  // only touched when NSPIN == 2
        for (int k = i + warp; k < n; k += NWARPS) {
            if (mrow && !((__ldg(mrow + (k >> 5)) >> (k & 31)) & 1u)) continue;  // (P | k, .) is all zero
            const int64_t Pk = (int64_t)k * n - (int64_t)k * (k - 1) / 2 - k;  // P'(k,l) = Pk + l
            const int l0 = (k == i) ? j : k;
            const double d0_ik = sD[k], d0_jk = sD[n + k];
            const double d1_ik = NSPIN == 2 ? sD[2 * n + k] : 0.0, d1_jk = NSPIN == 2 ? sD[3 * n + k] : 0.0;
            double ka0_i = 0.0, ka0_j = 0.0, ka1_i = 0.0, ka1_j = 0.0;
            bool anyk = false;
            // the loads of a (k, .) block are issued eight at a time before any of them is used: the pass
            // is a stream over HBM and needs many loads in flight per lane
            for (int lb = l0 + lane; lb < n; lb += 32 * 8) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int l = lb + 32 * u;
                    v[u] = l < n ? __ldcs(row + Pk + l) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double V = v[u];
                    if (V != 0.0) {
                        const int l = lb + 32 * u;
                        const int64_t Pp = Pk + l;
                        anyk = true;
                        const double wP = (Pp == P) ? 0.5 * V : V;
                        jacc = fma(a.dp2[Pp], wP, jacc);
                        atomicAdd(a.jp + Pp, dp2_P * wP);
                        const double w = wP * facP * ((k == l) ? 0.5 : 1.0);
                        ka0_i = fma(sD[n + l], w, ka0_i);  // K_ik += D_jl w
                        ka0_j = fma(sD[l], w, ka0_j);      // K_jk += D_il w
                        myK0[l] = fma(d0_jk, w, myK0[l]);          // K_il += D_jk w
                        myK0[n + l] = fma(d0_ik, w, myK0[n + l]);  // K_jl += D_ik w
                        if (NSPIN == 2) {
                            ka1_i = fma(sD[3 * n + l], w, ka1_i);
                            ka1_j = fma(sD[2 * n + l], w, ka1_j);
                            myK1[l] = fma(d1_jk, w, myK1[l]);
                            myK1[n + l] = fma(d1_ik, w, myK1[n + l]);
                        }
                    }
                }
            }
            if (!__any_sync(0xffffffffu, anyk)) continue;  // the whole (k, .) block of this row is zero
            any = true;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ka0_i += __shfl_xor_sync(0xffffffffu, ka0_i, o);
                ka0_j += __shfl_xor_sync(0xffffffffu, ka0_j, o);
                if (NSPIN == 2) {
                    ka1_i += __shfl_xor_sync(0xffffffffu, ka1_i, o);
                    ka1_j += __shfl_xor_sync(0xffffffffu, ka1_j, o);
                }
            }
            __syncwarp();
            if (lane == 0) {
                myK0[k] += ka0_i;
                myK0[n + k] += ka0_j;
                if (NSPIN == 2) { myK1[k] += ka1_i; myK1[n + k] += ka1_j; }
            }
            __syncwarp();
        }
        // J_ij of this row
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) jacc += __shfl_xor_sync(0xffffffffu, jacc, o);
        if (lane == 0) s_red[warp] = jacc;
        // rows without a single surviving integral (most rows of a large molecule) have nothing to flush
        if (!__syncthreads_or(any ? 1 : 0)) continue;
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < NWARPS; ++w) t += s_red[w];
            if (t != 0.0) atomicAdd(a.jp + P, t);
        }
        // flush the K rows of this packed row
        for (int x = tid; x < n; x += kThreads) {
#pragma unroll
            for (int s = 0; s < NSPIN; ++s) {
                double vi = 0.0, vj = 0.0;
                for (int w = 0; w < NWARPS; ++w) {
                    double* pk = sK + (size_t)(s * NWARPS + w) * 2 * n;
                    vi += pk[x];
                    vj += pk[n + x];
                    pk[x] = 0.0;
                    pk[n + x] = 0.0;
                }
                double* kacc = s == 0 ? a.k0 : a.k1;
                if (vi != 0.0) atomicAdd(kacc + i + (size_t)n * x, vi);
                if (vj != 0.0) atomicAdd(kacc + j + (size_t)n * x, vj);
            }
        }
        __syncthreads();
    }
}

// Sparsity mask of the packed array: bit k of row P = (i,j) is set iff some (ij|kl), l >= k, is nonzero.
// Built once per integral evaluation (one streaming pass), reused by every G build of the SCF.
__global__ void __launch_bounds__(256) fock_mask_kernel(const double* __restrict__ packed, int64_t out_offset, int64_t row_lo,
                                                        int64_t row_hi, int n, int64_t npair, const int2* __restrict__ ij,
                                                        uint32_t* __restrict__ mask, int mask_words) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t P = row_lo + blockIdx.x; P < row_hi; P += gridDim.x) {
        const int2 pij = ij[P];
        const int i = pij.x, j = pij.y;
        const double* row = packed + ((P * npair - ((P * (P - 1)) >> 1)) - out_offset) - P;
        for (int k = i + warp; k < n; k += 8) {
            const int64_t Pk = (int64_t)k * n - (int64_t)k * (k - 1) / 2 - k;
            const int l0 = (k == i) ? j : k;
            bool nz = false;
            for (int lb = l0 + lane; lb < n; lb += 32 * 8) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int l = lb + 32 * u;
                    v[u] = l < n ? __ldcs(row + Pk + l) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) nz |= (v[u] != 0.0);
            }
            if (__any_sync(0xffffffffu, nz) && lane == 0) atomicOr(mask + P * mask_words + (k >> 5), 1u << (k & 31));
        }
    }
}

// G = J - fac (K + K^T)
__global__ void fock_finalize_kernel(int n, const double* jp, const double* k, double fac, double* g) {
    const int b = blockIdx.x;
    for (int a = threadIdx.x; a < n; a += blockDim.x) {
        const int lo = a < b ? a : b, hi = a < b ? b : a;
        const int64_t P = (int64_t)lo * n - (int64_t)lo * (lo - 1) / 2 + (hi - lo);
        g[a + (size_t)n * b] = jp[P] - fac * (k[a + (size_t)n * b] + k[b + (size_t)n * a]);
    }
}

#define CUF(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return myqc::fock_fail(MYQC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

int fock_build(const double* d_packed, int64_t out_offset, int64_t out_elems, int norb, const double* d_da,
               const double* d_db, double* d_ga, double* d_gb, cudaStream_t st, const uint32_t* d_mask = nullptr) {
    const bool uhf = d_db != nullptr;
    if (!d_packed || !d_da || !d_ga || norb < 1 || (uhf && !d_gb)) return myqc::fock_fail(MYQC_ERR_BAD_ARG, "null pointer or bad norb");
    const int64_t n = norb, npair = n * (n + 1) / 2;
    // the slice must be a range of whole packed rows
    auto off = [&](int64_t P) { return P * npair - P * (P - 1) / 2; };
    auto row_of = [&](int64_t o) {  // smallest P with off(P) >= o
        int64_t lo = 0, hi = npair;
        while (lo < hi) {
            const int64_t mid = (lo + hi) / 2;
            if (off(mid) < o) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    const int64_t row_lo = row_of(out_offset), row_hi = row_of(out_offset + out_elems);
    if (off(row_lo) != out_offset || off(row_hi) != out_offset + out_elems)
        return myqc::fock_fail(MYQC_ERR_BAD_ARG, "the packed slice does not consist of whole rows");

    int dev = 0, sms = 0;
    CUF(cudaGetDevice(&dev));
    CUF(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int nspin = uhf ? 2 : 1;
    // warp-private K rows: RHF 8 warps x 2 rows, UHF 4 warps x 2 spins x 2 rows (same footprint)
    constexpr int NW1 = 8, NW2 = 4;
    const int nw = uhf ? NW2 : NW1;
    const size_t smem = (size_t)(nspin * 2 + nspin * nw * 2) * n * sizeof(double);
    if (smem > 220 * 1024) return myqc::fock_fail(MYQC_ERR_UNSUPPORTED, "norb too large for the shared-memory K rows of the Fock build");

    // temporaries, stream ordered
    const size_t nn = (size_t)n * n;
    const size_t bytes = sizeof(int2) * npair + sizeof(double) * (2 * (size_t)npair + 2 * nspin * nn) + 256;
    char* buf = nullptr;
    CUF(cudaMallocAsync((void**)&buf, bytes, st));
    int2* ij = reinterpret_cast<int2*>(buf);
    double* dp2 = reinterpret_cast<double*>(buf + sizeof(int2) * npair);
    double* jp = dp2 + npair;
    double* dk0 = jp + npair;
    double* k0 = dk0 + nn;
    double* dk1 = uhf ? k0 + nn : nullptr;
    double* k1 = uhf ? dk1 + nn : nullptr;
    int* counter = reinterpret_cast<int*>(buf + bytes - 256);
    CUF(cudaMemsetAsync(dp2, 0, sizeof(double) * 2 * npair, st));  // dp2, jp
    CUF(cudaMemsetAsync(k0, 0, sizeof(double) * nn, st));
    if (uhf) CUF(cudaMemsetAsync(k1, 0, sizeof(double) * nn, st));
    CUF(cudaMemsetAsync(counter, 0, sizeof(int), st));
    pair_table_kernel<<<norb, 128, 0, st>>>(norb, ij);
    density_prep_kernel<<<norb, 128, 0, st>>>(norb, d_da, dk0, dp2);
    if (uhf) density_prep_kernel<<<norb, 128, 0, st>>>(norb, d_db, dk1, dp2);
    CUF(cudaGetLastError());

    FockArgs a;
    a.packed = d_packed; a.out_offset = out_offset; a.row_lo = row_lo; a.row_hi = row_hi;
    a.n = norb; a.npair = npair; a.ij = ij; a.dp2 = dp2; a.dk0 = dk0; a.dk1 = dk1;
    a.jp = jp; a.k0 = k0; a.k1 = k1; a.counter = counter;
    a.mask = d_mask; a.mask_words = (norb + 31) / 32;
    int occ = 1;
    if (uhf) {
        CUF(cudaFuncSetAttribute(fock_rows_kernel<2, NW2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUF(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fock_rows_kernel<2, NW2>, NW2 * 32, smem));
    } else {
        CUF(cudaFuncSetAttribute(fock_rows_kernel<1, NW1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUF(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fock_rows_kernel<1, NW1>, NW1 * 32, smem));
    }
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)sms * occ;
    if (grid > row_hi - row_lo) grid = row_hi - row_lo;
    if (grid > 0) {
        if (uhf) fock_rows_kernel<2, NW2><<<(int)grid, NW2 * 32, smem, st>>>(a);
        else fock_rows_kernel<1, NW1><<<(int)grid, NW1 * 32, smem, st>>>(a);
        CUF(cudaGetLastError());
    }
    fock_finalize_kernel<<<norb, 128, 0, st>>>(norb, jp, k0, uhf ? 1.0 : 0.5, d_ga);
    if (uhf) fock_finalize_kernel<<<norb, 128, 0, st>>>(norb, jp, k1, 1.0, d_gb);
    CUF(cudaGetLastError());
    CUF(cudaFreeAsync(buf, st));
    return MYQC_OK;
}

int fock_host(const double* packed, int norb, const double* da, const double* db, double* ga, double* gb) {
    if (!packed || !da || !ga || norb < 1) return myqc::fock_fail(MYQC_ERR_BAD_ARG, "null pointer or bad norb");
    if (myqc_device_count() == 0) return myqc::fock_fail(MYQC_ERR_NO_DEVICE, "no CUDA device: the Fock build has no CPU fallback");
    const int64_t n = norb, npair = n * (n + 1) / 2, total = npair * (npair + 1) / 2;
    const size_t nn = (size_t)n * n;
    double *d_p = nullptr, *d_m = nullptr;
    CUF(cudaMalloc((void**)&d_p, sizeof(double) * total));
    cudaError_t e = cudaMalloc((void**)&d_m, sizeof(double) * 4 * nn);
    if (e != cudaSuccess) { cudaFree(d_p); return myqc::fock_fail(MYQC_ERR_NOMEM, cudaGetErrorString(e)); }
    int rc = MYQC_OK;
    e = cudaMemcpy(d_p, packed, sizeof(double) * total, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_m, da, sizeof(double) * nn, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && db) e = cudaMemcpy(d_m + nn, db, sizeof(double) * nn, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = myqc::fock_fail(MYQC_ERR_CUDA, cudaGetErrorString(e));
    if (!rc) rc = fock_build(d_p, 0, total, norb, d_m, db ? d_m + nn : nullptr, d_m + 2 * nn, db ? d_m + 3 * nn : nullptr, nullptr);
    if (!rc) {
        e = cudaMemcpy(ga, d_m + 2 * nn, sizeof(double) * nn, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && db) e = cudaMemcpy(gb, d_m + 3 * nn, sizeof(double) * nn, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = myqc::fock_fail(MYQC_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFree(d_p);
    cudaFree(d_m);
    return rc;
}

}  // namespace

extern "C" {

int myqc_fock_rhf(const double* d_packed, int64_t out_offset, int64_t out_elems, int norb, const double* d_da,
                  double* d_g, void* stream) {
    return fock_build(d_packed, out_offset, out_elems, norb, d_da, nullptr, d_g, nullptr, static_cast<cudaStream_t>(stream));
}

int myqc_fock_uhf(const double* d_packed, int64_t out_offset, int64_t out_elems, int norb, const double* d_da,
                  const double* d_db, double* d_ga, double* d_gb, void* stream) {
    if (!d_db) return myqc::fock_fail(MYQC_ERR_BAD_ARG, "null beta density");
    return fock_build(d_packed, out_offset, out_elems, norb, d_da, d_db, d_ga, d_gb, static_cast<cudaStream_t>(stream));
}

int64_t myqc_fock_mask_words(int norb) {
    const int64_t n = norb;
    return norb < 1 ? 0 : (n * (n + 1) / 2) * ((n + 31) / 32);
}

int myqc_fock_mask_build(const double* d_packed, int64_t out_offset, int64_t out_elems, int norb, uint32_t* d_mask, void* stream) {
    if (!d_packed || !d_mask || norb < 1) return myqc::fock_fail(MYQC_ERR_BAD_ARG, "null pointer or bad norb");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t n = norb, npair = n * (n + 1) / 2;
    auto off = [&](int64_t P) { return P * npair - P * (P - 1) / 2; };
    auto row_of = [&](int64_t o) {
        int64_t lo = 0, hi = npair;
        while (lo < hi) {
            const int64_t mid = (lo + hi) / 2;
            if (off(mid) < o) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    const int64_t row_lo = row_of(out_offset), row_hi = row_of(out_offset + out_elems);
    if (off(row_lo) != out_offset || off(row_hi) != out_offset + out_elems)
        return myqc::fock_fail(MYQC_ERR_BAD_ARG, "the packed slice does not consist of whole rows");
    const int mw = (norb + 31) / 32;
    int dev = 0, sms = 0;
    CUF(cudaGetDevice(&dev));
    CUF(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int2* ij = nullptr;
    CUF(cudaMallocAsync((void**)&ij, sizeof(int2) * npair, st));
    pair_table_kernel<<<norb, 128, 0, st>>>(norb, ij);
    CUF(cudaMemsetAsync(d_mask + row_lo * mw, 0, sizeof(uint32_t) * (size_t)((row_hi - row_lo) * mw), st));
    int64_t grid = (int64_t)sms * 8;
    if (grid > row_hi - row_lo) grid = row_hi - row_lo;
    if (grid > 0) fock_mask_kernel<<<(int)grid, 256, 0, st>>>(d_packed, out_offset, row_lo, row_hi, norb, npair, ij, d_mask, mw);
    CUF(cudaGetLastError());
    CUF(cudaFreeAsync(ij, st));
    return MYQC_OK;
}

int myqc_fock_rhf_masked(const double* d_packed, int64_t out_offset, int64_t out_elems, int norb, const double* d_da,
                         const uint32_t* d_mask, double* d_g, void* stream) {
    return fock_build(d_packed, out_offset, out_elems, norb, d_da, nullptr, d_g, nullptr, static_cast<cudaStream_t>(stream), d_mask);
}

int myqc_fock_uhf_masked(const double* d_packed, int64_t out_offset, int64_t out_elems, int norb, const double* d_da,
                         const double* d_db, const uint32_t* d_mask, double* d_ga, double* d_gb, void* stream) {
    if (!d_db) return myqc::fock_fail(MYQC_ERR_BAD_ARG, "null beta density");
    return fock_build(d_packed, out_offset, out_elems, norb, d_da, d_db, d_ga, d_gb, static_cast<cudaStream_t>(stream), d_mask);
}

int myqc_fock_rhf_host(const double* packed, int norb, const double* da, double* g) {
    return fock_host(packed, norb, da, nullptr, g, nullptr);
}

int myqc_fock_uhf_host(const double* packed, int norb, const double* da, const double* db, double* ga, double* gb) {
    if (!db || !gb) return myqc::fock_fail(MYQC_ERR_BAD_ARG, "null beta density / result");
    return fock_host(packed, norb, da, db, ga, gb);
}

}  // extern "C"
