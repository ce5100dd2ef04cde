// One-electron integrals on the device: overlap S and core Hamiltonian H = T + V
// (include/myqc_int1e.h).
//
// Reference: src/integrals/int1e.f90 proc1e :132-280 (ordered loop over primitive sets a, b with the
// EIJ < 1e-14 skip :246-248), overlap :321-386, kinetic :391-474, coulomb :479-582, with getcoef /
// getDk / Boys / RNLMj of auxilary.f90.  Same conventions as the ERI kernels: float32 pi, Boys values
// from the Ftab bytes with start order Q = 3(la+lb), terms whose Hermite factor is below 1e-16 dropped.
//
// Mapping: one CTA per ordered set pair (a,b) -- the reference's loop, so both triangles of S and H
// are produced by the same arithmetic as there.  The (at most 4x4) orbital pairs of the two sets
// take the overlap and kinetic terms (closed-form Hermite coefficients up to nbar = 3 by the
// McMurchie-Davidson recurrences); all threads stride over the nuclei for the attraction term
// (Boys + R_NLM(p, P-C) per nucleus, 16 partial sums per thread), one block reduction, one atomic
// add per orbital pair.  O(nset^2 nnuc) work: microseconds next to int2e; it is on the device so that
// the whole integral stage runs without a Fortran binary, sharing boys.cuh with the ERI kernels.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/myqc_eri.h"
#include "../../include/myqc_int1e.h"
#include "boys.cuh"

namespace myqc {
int fock_fail(int code, const std::string& msg);  // sets myqc_last_error (eri_api.cu)
void build_boys_tables(const double* ftab, std::vector<double>& taylor, std::vector<double>& ex);  // eri_api.cu
}

namespace myqc {
namespace {

constexpr int kThreads1e = 128;
constexpr double kPiRefD = 3.1415927410125732;  // float32 pi widened

struct Int1eArgs {
    int nnuc, nset, setl, ops, norb;
    const double* xyz;        // [nnuc*3] Fortran layout xyz[i + nnuc*c]
    const int32_t* atoms;     // [nnuc] nuclear charges
    const double* set;        // [nset]
    const int32_t* setinfo;   // [2 + setl*nset]
    const double* bas;        // [ops*nset]
    const int32_t* basinfo;   // [2 + 5*norb]
    const double* g0;         // [nset] gtoD(0, alpha)
    const double* g1;         // [nset] gtoD(1, alpha)
    const double* taylor;     // [5][121][8] Boys Taylor tables for Q = 0,3,6,9,12
    const double2* exptab;    // [601] {exp(-k/10), k/10}
    double* S;                // [norb*norb] column-major
    double* H;
};

// Hermite expansion coefficients E_N^{i,j} for i <= 1, j <= 3 along one axis (auxilary.f90:406-536):
//   E_N^{i,j+1} = h E_{N-1}^{i,j} + PB E_N^{i,j} + (N+1) E_{N+1}^{i,j},  same with PA for i+1
struct Herm {
    double e[5][2][4];  // [N][i][j]
};
__device__ void hermite_axis(double PA, double PB, double h, Herm& t) {
    for (int N = 0; N < 5; ++N)
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 4; ++j) t.e[N][i][j] = 0.0;
    t.e[0][0][0] = 1.0;
    for (int j = 0; j < 3; ++j)
        for (int N = 0; N <= j + 1; ++N) {
            double v = PB * t.e[N][0][j];
            if (N > 0) v += h * t.e[N - 1][0][j];
            if (N + 1 <= j) v += (double)(N + 1) * t.e[N + 1][0][j];
            t.e[N][0][j + 1] = v;
        }
    for (int j = 0; j < 4; ++j)
        for (int N = 0; N <= j + 1; ++N) {
            double v = PA * t.e[N][0][j];
            if (N > 0) v += h * t.e[N - 1][0][j];
            if (N + 1 <= j) v += (double)(N + 1) * t.e[N + 1][0][j];
            t.e[N][1][j] = v;
        }
}

// R_NLM(p, P-C), N+L+M <= LT, from the reference's Boys procedure with start order Q = 3 LT
template <int LT>
__device__ void coulomb_R(double p, double X, double Y, double Z, const double* taylor, const double2* exptab,
                          double (&R)[10]) {
    constexpr int Q = 3 * LT;
    const double R2 = fma(X, X, fma(Y, Y, Z * Z));
    const double T = p * R2;
    double G[LT + 1];
    if (T >= (double)(2 * Q + 36)) {
        // Boys3, auxilary.f90:194-215: F0 = sqrt(pi)/2 / sqrt(T), F_j = (2j-1)/(2T) F_{j-1}
        const double rT = rsqrt_pos(T);
        const double h = 0.5 * rT * rT;
        double f = kHalfSqrtPi * rT, w = 1.0;
        G[0] = f;
#pragma unroll
        for (int j = 1; j <= LT; ++j) {
            f = f * (double)(2 * j - 1) * h;
            w *= -2.0 * p;
            G[j] = w * f;
        }
    } else {
        boys_near_mid<Q, LT>(T, p, 1.0, G, taylor + (size_t)LT * 121 * 8, exptab);  // G_j = (-2p)^j F_j
    }
    double Rt[h_count(LT)];
    build_R<LT>(G, X, Y, Z, Rt);
#pragma unroll
    for (int k = 0; k < h_count(LT); ++k) R[k] = Rt[k];
}

__global__ void __launch_bounds__(kThreads1e) int1e_kernel(const Int1eArgs a) {
    const int sa = blockIdx.x, sb = blockIdx.y;
    const int tid = threadIdx.x;
    const int nnuc = a.nnuc;
    const int32_t* ia = a.setinfo + 1 + sa * a.setl + 1;  // {#orbitals, max l, centre, orbital ids}
    const int32_t* ib = a.setinfo + 1 + sb * a.setl + 1;
    const double aa = a.set[sa], bb = a.set[sb];
    const int u = ia[2], v = ib[2];
    const int la = ia[1], lb = ib[1];
    const double p = aa + bb, m = aa * bb;
    double AB2 = 0.0, PP[3], PA[3], PB[3];
    for (int c = 0; c < 3; ++c) {
        const double xu = a.xyz[u + nnuc * c], xv = a.xyz[v + nnuc * c];
        const double ab = xu - xv;
        AB2 += ab * ab;
        PP[c] = (aa * xu + bb * xv) / p;
        PA[c] = PP[c] - xu;
        PB[c] = PP[c] - xv;
    }
    const double EIJ = exp(-m * AB2 / p);
    if (EIJ < 1.0e-14) return;  // int1e.f90:246-248

    __shared__ Herm sh[3];
    __shared__ double s_acc[kThreads1e / 32][16];
    if (tid < 3) hermite_axis(PA[tid], PB[tid], 1.0 / (2.0 * p), sh[tid]);
    __syncthreads();

    const int na = ia[0], nb = ib[0];
    auto lvec = [&](int orb, int* l) {  // basinfo: {n, l, ori(-1 = s, 0/1/2 = x/y/z), #prim, centre}
        l[0] = l[1] = l[2] = 0;
        const int ori = a.basinfo[1 + 5 * orb + 3];
        if (a.basinfo[1 + 5 * orb + 2] == 1 && ori >= 0) l[ori] = 1;
    };
    // ---- overlap and kinetic energy: one thread per orbital pair ------------------------------
    if (tid < na * nb) {
        const int i = tid / nb, j = tid % nb;
        const int oa = ia[3 + i], ob = ib[3 + j];
        int n1[3], n2[3];
        lvec(oa, n1);
        lvec(ob, n2);
        const double x = kPiRefD / p;
        const double pref = EIJ * (x * sqrt(x));  // EIJ (Pi/p)^(3/2)
        const double norm = a.bas[sa * a.ops + i] * a.bas[sb * a.ops + j] *
                            (a.basinfo[1 + 5 * oa + 2] == 0 ? a.g0[sa] : a.g1[sa]) *
                            (a.basinfo[1 + 5 * ob + 2] == 0 ? a.g0[sb] : a.g1[sb]);
        auto E0 = [&](int w, int ii, int jj) { return jj < 0 ? 0.0 : sh[w].e[0][ii][jj]; };
        const double s = pref * norm * E0(0, n1[0], n2[0]) * E0(1, n1[1], n2[1]) * E0(2, n1[2], n2[2]);
        atomicAdd(a.S + oa + (size_t)a.norb * ob, s);
        double val = 0.0;
        for (int w = 0; w < 3; ++w) {
            const int w1 = (w + 1) % 3, w2 = (w + 2) % 3;
            double t = (double)(n2[w] * (n2[w] - 1)) * E0(w, n1[w], n2[w] - 2);
            t = t - 2.0 * bb * (double)n2[w] * E0(w, n1[w], n2[w]);
            t = t - 2.0 * bb * (double)(n2[w] + 1) * E0(w, n1[w], n2[w]);
            t = t + 4.0 * bb * bb * E0(w, n1[w], n2[w] + 2);
            val += t * E0(w1, n1[w1], n2[w1]) * E0(w2, n1[w2], n2[w2]);
        }
        atomicAdd(a.H + oa + (size_t)a.norb * ob, val * (-0.5) * pref * norm);
    }
    // ---- nuclear attraction: threads stride over the nuclei --------------------------------------
    double acc[16];
#pragma unroll
    for (int f = 0; f < 16; ++f) acc[f] = 0.0;
    const int LT = la + lb;
    for (int c = tid; c < nnuc; c += kThreads1e) {
        const double X = PP[0] - a.xyz[c], Y = PP[1] - a.xyz[c + nnuc], Z = PP[2] - a.xyz[c + 2 * nnuc];  // P - C
        double R[10];
        if (LT == 0) coulomb_R<0>(p, X, Y, Z, a.taylor, a.exptab, R);
        else if (LT == 1) coulomb_R<1>(p, X, Y, Z, a.taylor, a.exptab, R);
        else coulomb_R<2>(p, X, Y, Z, a.taylor, a.exptab, R);
        const double zc = (double)a.atoms[c];
        for (int i = 0; i < na; ++i) {
            int n1[3];
            lvec(ia[3 + i], n1);
            for (int j = 0; j < nb; ++j) {
                int n2[3];
                lvec(ib[3 + j], n2);
                double sum = 0.0;
                for (int N = 0; N <= n1[0] + n2[0]; ++N) {
                    const double dx = sh[0].e[N][n1[0]][n2[0]];
                    if (fabs(dx) < 1.0e-16) continue;  // getDk drops terms with a vanishing factor (auxilary.f90:573)
                    for (int L = 0; L <= n1[1] + n2[1]; ++L) {
                        const double dy = sh[1].e[L][n1[1]][n2[1]];
                        if (fabs(dy) < 1.0e-16) continue;
                        for (int M = 0; M <= n1[2] + n2[2]; ++M) {
                            const double dz = sh[2].e[M][n1[2]][n2[2]];
                            if (fabs(dz) < 1.0e-16) continue;
                            sum = fma(dx * dy * dz, R[h_index(N, L, M)], sum);
                        }
                    }
                }
                acc[i * 4 + j] = fma(zc, sum, acc[i * 4 + j]);
            }
        }
    }
#pragma unroll
    for (int f = 0; f < 16; ++f) {
        double t = acc[f];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((tid & 31) == 0) s_acc[tid >> 5][f] = t;
    }
    __syncthreads();
    if (tid < na * nb) {
        const int i = tid / nb, j = tid % nb;
        const int oa = ia[3 + i], ob = ib[3 + j];
        double t = 0.0;
        for (int w = 0; w < kThreads1e / 32; ++w) t += s_acc[w][i * 4 + j];
        const double norm = a.bas[sa * a.ops + i] * a.bas[sb * a.ops + j] *
                            (a.basinfo[1 + 5 * oa + 2] == 0 ? a.g0[sa] : a.g1[sa]) *
                            (a.basinfo[1 + 5 * ob + 2] == 0 ? a.g0[sb] : a.g1[sb]);
        atomicAdd(a.H + oa + (size_t)a.norb * ob, -(2.0 * kPiRefD / p) * EIJ * norm * t);
    }
}

double gtoD_host(int l, double alpha) {  // auxilary.f90:704-726
    const double pi = (double)3.1415926535897931f;
    if (l == 0) return std::pow(2.0 * alpha / pi, 3.0 / 4.0);
    return std::pow(128.0 * std::pow(alpha, 5.0) / std::pow(pi, 3.0), 1.0 / 4.0);
}

}  // namespace
}  // namespace myqc

using namespace myqc;

extern "C" int myqc_int1e(int nnuc, const double* xyz, const int32_t* atoms, int nset, int setl, const double* set,
                          const int32_t* setinfo, int ops, const double* bas, const int32_t* basinfo,
                          const double* ftab, double* s_out, double* h_out) {
    if (!xyz || !atoms || !set || !setinfo || !bas || !basinfo || !ftab || !s_out || !h_out)
        return fock_fail(MYQC_ERR_BAD_ARG, "null pointer");
    if (nnuc < 1 || nset < 1 || setinfo[0] != nset || setinfo[1] != setl || basinfo[0] != ops || ops < 1 || ops > 4 || basinfo[1] < 1)
        return fock_fail(MYQC_ERR_BAD_ARG, "inconsistent sizes");
    if (myqc_device_count() == 0) return fock_fail(MYQC_ERR_NO_DEVICE, "no CUDA device: the integral engine has no CPU fallback");
    const int norb = basinfo[1];
    for (int o = 0; o < norb; ++o)
        if (basinfo[1 + 5 * o + 2] > 1) return fock_fail(MYQC_ERR_UNSUPPORTED, "angular momentum l > 1 is not implemented (basis.f90:170-174)");
    std::vector<double> taylor, ex, g0(nset), g1(nset);
    build_boys_tables(ftab, taylor, ex);
    for (int s = 0; s < nset; ++s) { g0[s] = gtoD_host(0, set[s]); g1[s] = gtoD_host(1, set[s]); }

    // one device buffer for everything
    const size_t nn = (size_t)norb * norb;
    const size_t n_d = (size_t)3 * nnuc + nset + (size_t)ops * nset + 2 * (size_t)nset + taylor.size() + ex.size() + 2 * nn;
    const size_t n_i = (size_t)nnuc + 2 + (size_t)setl * nset + 2 + (size_t)5 * norb;
    char* buf = nullptr;
    cudaError_t e = cudaMalloc((void**)&buf, n_d * sizeof(double) + n_i * sizeof(int32_t) + 64);
    if (e != cudaSuccess) return fock_fail(MYQC_ERR_NOMEM, cudaGetErrorString(e));
    double* dd = reinterpret_cast<double*>(buf);
    size_t off = 0;
    auto put_d = [&](const double* src, size_t n) {
        double* d = dd + off;
        if (src) cudaMemcpy(d, src, n * sizeof(double), cudaMemcpyHostToDevice);
        off += n;
        return d;
    };
    Int1eArgs a;
    a.nnuc = nnuc; a.nset = nset; a.setl = setl; a.ops = ops; a.norb = norb;
    // the Boys tables are read with 16-byte loads: they go first (cudaMalloc alignment, even sizes)
    a.taylor = put_d(taylor.data(), taylor.size());
    a.exptab = reinterpret_cast<const double2*>(put_d(ex.data(), ex.size()));
    a.xyz = put_d(xyz, (size_t)3 * nnuc);
    a.set = put_d(set, nset);
    a.bas = put_d(bas, (size_t)ops * nset);
    a.g0 = put_d(g0.data(), nset);
    a.g1 = put_d(g1.data(), nset);
    a.S = put_d(nullptr, nn);
    a.H = put_d(nullptr, nn);
    int32_t* di = reinterpret_cast<int32_t*>(dd + off);
    size_t ioff = 0;
    auto put_i = [&](const int32_t* src, size_t n) {
        int32_t* d = di + ioff;
        cudaMemcpy(d, src, n * sizeof(int32_t), cudaMemcpyHostToDevice);
        ioff += n;
        return d;
    };
    a.atoms = put_i(atoms, nnuc);
    a.setinfo = put_i(setinfo, 2 + (size_t)setl * nset);
    a.basinfo = put_i(basinfo, 2 + (size_t)5 * norb);
    cudaMemset(a.S, 0, 2 * nn * sizeof(double));
    int rc = MYQC_OK;
    if (nset > 65535) rc = fock_fail(MYQC_ERR_UNSUPPORTED, "more than 65535 primitive sets");
    if (!rc) {
        int1e_kernel<<<dim3(nset, nset), kThreads1e>>>(a);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(s_out, a.S, nn * sizeof(double), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(h_out, a.H, nn * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fock_fail(MYQC_ERR_CUDA, std::string("int1e kernel: ") + cudaGetErrorString(e));
    }
    cudaFree(buf);
    return rc;
}
