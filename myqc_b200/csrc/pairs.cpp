// Host-side shell-pair builder (see pairs.hpp).  Built with -ffp-contract=off so the
// prefactor EIJ is bit-identical to what the reference's gfortran build computes: the
// EIJ*EGH >= 1e-14 inclusion rule (int2e.f90:257) must make the same decisions.
#include "pairs.hpp"
#include "terms.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <numeric>
#include <thread>

#include "../../include/myqc_eri.h"

namespace myqc {

// REAL(KIND=8),PARAMETER :: Pi = 3.1415926535897931 -- a default-REAL literal, i.e. float32
// pi widened (int2e.f90:621, auxilary.f90:711).  SURVEY.md T1.
static const double kPiRef = (double)3.1415926535897931f;

// auxilary.f90:704-726
static double gtoD(int l, double a) {
    if (l == 0) return std::pow(2.0 * a / kPiRef, 3.0 / 4.0);
    return std::pow(128.0 * std::pow(a, 5.0) / std::pow(kPiRef, 3.0), 1.0 / 4.0);
}

int build_shells(int nnuc, int nset, int setl, const int32_t* setinfo, int ops,
                 const int32_t* basinfo, std::vector<Shell>& shells, std::string& err) {
    (void)nnuc;
    shells.clear();
    if (setl < 4 || ops < 1 || ops > 4 || setl != 3 + ops) {
        err = "setl/ops not of the form 3+OpS with OpS<=4";
        return MYQC_ERR_BAD_ARG;
    }
    const int norb = basinfo[1];
    for (int s = 0; s < nset; ++s) {
        const int32_t* si = &setinfo[1 + s * setl + 1];  // {#orbs, max l, centre, orbital ids...}
        const int no = si[0];
        if (no < 1 || no > ops) { err = "set with no orbitals or more than OpS"; return MYQC_ERR_BAD_ARG; }
        for (int k = 0; k < no; ++k)
            if (si[3 + k] < 0 || si[3 + k] >= norb) { err = "orbital id out of range"; return MYQC_ERR_BAD_ARG; }
        // look for an existing shell with the same centre and orbital list
        int found = -1;
        for (size_t h = 0; h < shells.size() && found < 0; ++h) {
            const Shell& sh = shells[h];
            if (sh.centre != si[2]) continue;
            const int32_t* s0 = &setinfo[1 + sh.sets[0] * setl + 1];
            if (s0[0] != no) continue;
            bool same = true;
            for (int k = 0; k < no; ++k) same = same && (s0[3 + k] == si[3 + k]);
            if (same) found = (int)h;
        }
        if (found >= 0) { shells[found].sets.push_back(s); continue; }
        Shell sh;
        sh.centre = si[2];
        sh.fn[0] = sh.fn[1] = sh.fn[2] = sh.fn[3] = -1;
        sh.sets.push_back(s);
        auto L = [&](int o) { return basinfo[1 + 5 * o + 2]; };
        auto ORI = [&](int o) { return basinfo[1 + 5 * o + 3]; };
        for (int k = 0; k < no; ++k)
            if (L(si[3 + k]) > 1) { err = "angular momentum l > 1 is not implemented (basis.f90:170-174)"; return MYQC_ERR_UNSUPPORTED; }
        if (no == 1 && L(si[3]) == 0 && ORI(si[3]) == -1) {
            sh.type = 0;
            sh.fn[0] = si[3];
        } else if (no == 4 && L(si[3]) == 0 && ORI(si[3]) == -1 && L(si[4]) == 1 && ORI(si[4]) == 0 &&
                   L(si[5]) == 1 && ORI(si[5]) == 1 && L(si[6]) == 1 && ORI(si[6]) == 2) {
            sh.type = 1;
            for (int k = 0; k < 4; ++k) sh.fn[k] = si[3 + k];
        } else if (no == 3 && L(si[3]) == 1 && ORI(si[3]) == 0 && L(si[4]) == 1 && ORI(si[4]) == 1 &&
                   L(si[5]) == 1 && ORI(si[5]) == 2) {
            sh.type = 1;  // p-only set: an SP shell whose s function is absent
            for (int k = 0; k < 3; ++k) sh.fn[1 + k] = si[3 + k];
        } else {
            err = "unsupported set layout (only S, SP and P sets are implemented)";
            return MYQC_ERR_UNSUPPORTED;
        }
        if (si[1] != sh.type) { err = "set max-l inconsistent with its orbitals"; return MYQC_ERR_BAD_ARG; }
        sh.first_fn = norb;
        for (int k = 0; k < 4; ++k)
            if (sh.fn[k] >= 0) sh.first_fn = std::min(sh.first_fn, sh.fn[k]);
        shells.push_back(sh);
    }
    for (const Shell& sh : shells)
        if ((int)sh.sets.size() > 3) { err = "more than 3 primitives per shell"; return MYQC_ERR_UNSUPPORTED; }
    return MYQC_OK;
}

namespace {

struct PrimRec {
    double E;
    double f[kCoefField + 46 + 2];
};

// 30-bit Morton code of a point inside the bounding box [lo, lo+ext]^3
uint32_t morton3(const double* r, const double* lo, double ext) {
    uint32_t code = 0;
    uint32_t q[3];
    for (int c = 0; c < 3; ++c) {
        double t = ext > 0 ? (r[c] - lo[c]) / ext : 0.0;
        t = t < 0 ? 0 : (t > 1 ? 1 : t);
        q[c] = (uint32_t)(t * 1023.0);
    }
    for (int b = 9; b >= 0; --b)
        for (int c = 0; c < 3; ++c) code = (code << 1) | ((q[c] >> b) & 1u);
    return code;
}

// contraction coefficient of function slot mu (0=s,1..3=p) of `shell` in primitive set `s`
double slot_coef(const Shell& sh, int s, int mu, const int32_t* setinfo, int setl, int ops,
                 const double* bas) {
    const int32_t* si = &setinfo[1 + s * setl + 1];
    if (sh.fn[mu] < 0) return 0.0;
    for (int k = 0; k < si[0]; ++k)
        if (si[3 + k] == sh.fn[mu]) return bas[s * ops + k];
    return 0.0;
}

}  // namespace

int build_pairs(int nnuc, const double* xyz, const double* set, const int32_t* setinfo, int setl,
                int ops, const double* bas, const int32_t* basinfo,
                const std::vector<Shell>& shells, PairList lists[3], std::string& err) {
    const double tol = 0.1e-15;  // auxilary.f90:573
    struct Tmp {
        int type, A, B, nprim, bucket;
        uint32_t morton;
        double emax;
        double centre[3];
        std::vector<PrimRec> prims;
    };
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < nnuc; ++i)
        for (int c = 0; c < 3; ++c) {
            lo[c] = std::min(lo[c], xyz[i + nnuc * c]);
            hi[c] = std::max(hi[c], xyz[i + nnuc * c]);
        }
    const double ext = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));
    const bool trace = std::getenv("MYQC_TRACE") != nullptr;
    auto tprev = std::chrono::steady_clock::now();
    auto stage = [&](const char* name) {
        if (!trace) return;
        const auto t = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[myqc trace]     pairs stage %-20s %.1f ms\n", name, std::chrono::duration<double, std::milli>(t - tprev).count());
        tprev = t;
    };
    const int ns = (int)shells.size();
    // normalisation constants per primitive set (gtoD, auxilary.f90:704-726) -- same values as computing
    // them per term, computed once
    int nset_all = 0;
    for (const Shell& sh : shells)
        for (int a : sh.sets) nset_all = std::max(nset_all, a + 1);
    std::vector<double> g0(nset_all), g1(nset_all);
    for (int a = 0; a < nset_all; ++a) { g0[a] = gtoD(0, set[a]); g1[a] = gtoD(1, set[a]); }
    // shell pairs (A, B >= A) are independent: rows A are dealt to a few host threads, results are
    // concatenated in A order so that the lists do not depend on the thread count
    std::vector<std::vector<Tmp>> perA(ns);
    std::vector<int> rcA(ns, MYQC_OK);
    auto do_row = [&](int A) {
        std::vector<Tmp>& outA = perA[A];
        for (int B = A; B < ns; ++B) {
            const Shell& sa = shells[A];
            const Shell& sb = shells[B];
            const int type = sa.type + sb.type;
            const int nterm = pt_nterm(type);
            Tmp t;
            t.type = type; t.A = A; t.B = B; t.emax = 0.0;
            for (size_t ia = 0; ia < sa.sets.size(); ++ia) {
                for (size_t ib = 0; ib < sb.sets.size(); ++ib) {
                    // A shell with itself: the set pairs (a,b) and (b,a) have the same p, P and E and differ only in
                    // their coefficients, which add; one record with the summed coefficients does the work of two.
                    if (A == B && ib < ia) continue;
                    const int a = sa.sets[ia], b = sb.sets[ib];
                    // int2e.f90:215-223
                    const double aa = set[a], bb = set[b];
                    const int u = sa.centre, v = sb.centre;
                    const double p = aa + bb, mm = aa * bb;
                    double AB[3], PP[3], PA[3], PB[3];
                    for (int i = 0; i < 3; ++i) {
                        AB[i] = xyz[u + nnuc * i] - xyz[v + nnuc * i];
                        PP[i] = (aa * xyz[u + nnuc * i] + bb * xyz[v + nnuc * i]) / p;
                        PA[i] = PP[i] - xyz[u + nnuc * i];
                        PB[i] = PP[i] - xyz[v + nnuc * i];
                    }
                    const double EIJ = std::exp(-mm * (AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2]) / p);
                    if (EIJ < 1.0e-14) continue;  // EGH <= 1, so EIJ*EGH < 1e-14 for every ket (:257)
                    // closed forms of lrec/rrec for n,nbar <= 1 (auxilary.f90:443-462)
                    const double half = 1.0 / (2.0 * p);
                    double d[3][2][2][3];  // [axis][n][nbar][N]
                    for (int w = 0; w < 3; ++w) {
                        d[w][0][0][0] = 1.0;  d[w][0][0][1] = 0.0;  d[w][0][0][2] = 0.0;
                        d[w][1][0][0] = PA[w]; d[w][1][0][1] = half; d[w][1][0][2] = 0.0;
                        d[w][0][1][0] = PB[w]; d[w][0][1][1] = half; d[w][0][1][2] = 0.0;
                        d[w][1][1][0] = PB[w] * PA[w] + half;
                        d[w][1][1][1] = PA[w] / (2.0 * p) + PB[w] * half;
                        d[w][1][1][2] = half / (2.0 * p);
                    }
                    // 2 Pi^2.5 / (p q sqrt(p+q)) is split as scale(p)*scale(q)/sqrt(p+q)
                    const double scale = std::sqrt(2.0) * std::pow(kPiRef, 1.25) / p;
                    // coefficient of term (mu,nu;N,L,M) with the first function in set sa_ and the second in set sb_
                    auto term = [&](int sa_, int sb_, int mu, int nu, int N, int L, int M) -> double {
                        // mu,nu in {0=s,1=x,2=y,3=z}; getDk, auxilary.f90:610-622
                        int n[3] = {mu == 1, mu == 2, mu == 3};
                        int nb[3] = {nu == 1, nu == 2, nu == 3};
                        const double cx = d[0][n[0]][nb[0]][N], cy = d[1][n[1]][nb[1]][L], cz = d[2][n[2]][nb[2]][M];
                        if (std::fabs(cz) < tol || std::fabs(cy) < tol || std::fabs(cx) < tol) return 0.0;
                        double Dk = cx * cy * cz;
                        Dk = Dk * EIJ * (mu != 0 ? g1[sa_] : g0[sa_]) * (nu != 0 ? g1[sb_] : g0[sb_]);
                        Dk = Dk * slot_coef(sa, sa_, mu, setinfo, setl, ops, bas) * slot_coef(sb, sb_, nu, setinfo, setl, ops, bas);
                        return Dk * scale;
                    };
                    PrimRec r;
                    std::memset(&r, 0, sizeof(r));
                    r.E = EIJ;
                    r.f[0] = p; r.f[1] = PP[0]; r.f[2] = PP[1]; r.f[3] = PP[2]; r.f[4] = EIJ;
                    r.f[5] = 1.0 / std::sqrt(p);
                    double* c = &r.f[kCoefField];
                    const bool a_is_sp = (sa.type == 1);
                    for (int k = 0; k < nterm; ++k) {
                        const int f = term_fn(type, k), h = term_h(type, k);
                        int mu = 0, nu = 0;
                        if (type == PT_SSP) { mu = a_is_sp ? f : 0; nu = a_is_sp ? 0 : f; }
                        else if (type == PT_SPSP) { mu = f / 4; nu = f % 4; }
                        c[k] = term(a, b, mu, nu, h_N(h), h_L(h), h_M(h));
                        if (A == B && ia != ib) c[k] += term(b, a, mu, nu, h_N(h), h_L(h), h_M(h));  // same centre: PA = PB = 0 both ways
                    }
                    t.prims.push_back(r);
                    t.emax = std::max(t.emax, EIJ);
                }
            }
            if (t.prims.empty()) continue;
            std::stable_sort(t.prims.begin(), t.prims.end(), [](const PrimRec& x, const PrimRec& y) { return x.E > y.E; });
            t.nprim = (int)t.prims.size();
            if (t.nprim > kMaxPrim) { rcA[A] = MYQC_ERR_UNSUPPORTED; return; }
            t.bucket = emax_bucket(t.emax);
            for (int c = 0; c < 3; ++c) t.centre[c] = t.prims[0].f[1 + c];  // centre of the dominant primitive
            t.morton = morton3(t.centre, lo, ext);
            outA.push_back(std::move(t));
        }
    };
    {
        int nthr = (int)std::thread::hardware_concurrency();
        if (const char* e = std::getenv("MYQC_HOST_THREADS")) nthr = std::atoi(e);
        nthr = std::max(1, std::min(nthr, 8));
        if (ns < 64) nthr = 1;
        if (nthr == 1) {
            for (int A = 0; A < ns; ++A) do_row(A);
        } else {
            std::vector<std::thread> th;
            for (int w = 0; w < nthr; ++w)
                th.emplace_back([&, w] { for (int A = w; A < ns; A += nthr) do_row(A); });
            for (auto& x : th) x.join();
        }
    }
    stage("records");
    std::vector<Tmp> tmp[3];
    for (int A = 0; A < ns; ++A) {
        if (rcA[A] != MYQC_OK) { err = "more than 9 primitive pairs per shell pair"; return rcA[A]; }
        for (Tmp& t : perA[A]) tmp[t.type].push_back(std::move(t));
    }
    stage("merge");
    for (int type = 0; type < 3; ++type) {
        std::vector<Tmp>& v = tmp[type];
        std::stable_sort(v.begin(), v.end(), [](const Tmp& x, const Tmp& y) {
            if (x.bucket != y.bucket) return x.bucket < y.bucket;
            return x.morton < y.morton;
        });
        PairList& pl = lists[type];
        pl = PairList();
        pl.type = type;
        pl.n = (int)v.size();
        pl.npad = (pl.n + 31) / 32 * 32;
        const int nf = pt_nf(type), nfield = pt_nfield(type);
        pl.emax.resize(pl.n); pl.nprim.resize(pl.n); pl.bucket.resize(pl.n);
        pl.shA.resize(pl.n); pl.shB.resize(pl.n); pl.owner_fn.resize(pl.n);
        pl.pidx.assign((size_t)pl.n * nf, -1);
        pl.aos.assign((size_t)pl.n * kMaxPrim * nfield, 0.0);
        pl.soa.assign((size_t)kMaxPrim * nfield * pl.npad, 0.0);
        const int64_t norb = basinfo[1];
        for (int k = 0; k < pl.n; ++k) {
            const Tmp& t = v[k];
            const Shell& sa = shells[t.A];
            const Shell& sb = shells[t.B];
            pl.emax[k] = t.emax; pl.nprim[k] = t.nprim; pl.bucket[k] = t.bucket;
            pl.shA[k] = t.A; pl.shB[k] = t.B;
            pl.owner_fn[k] = std::min(sa.first_fn, sb.first_fn);
            for (int f = 0; f < nf; ++f) {
                int fi, fj;
                if (type == PT_SS) { fi = sa.fn[0]; fj = sb.fn[0]; }
                else if (type == PT_SSP) {
                    const bool a_is_sp = (sa.type == 1);
                    fi = a_is_sp ? sa.fn[f] : sa.fn[0];
                    fj = a_is_sp ? sb.fn[0] : sb.fn[f];
                } else { fi = sa.fn[f / 4]; fj = sb.fn[f % 4]; }
                if (fi < 0 || fj < 0) continue;            // absent function (p-only set)
                if (t.A == t.B && fi > fj) continue;       // (j,i) duplicate inside a diagonal shell pair
                const int64_t i = std::min(fi, fj), j = std::max(fi, fj);
                pl.pidx[(size_t)k * nf + f] = (int32_t)(i * norb - i * (i - 1) / 2 + (j - i));
            }
            for (int q = 0; q < t.nprim; ++q)
                std::memcpy(&pl.aos[((size_t)k * kMaxPrim + q) * nfield], t.prims[q].f, sizeof(double) * nfield);
        }
        // structure-of-arrays copy: tiles of 32 pairs so that both sides of the transpose stay in cache
        for (int k0 = 0; k0 < pl.n; k0 += 32) {
            const int k1 = std::min(pl.n, k0 + 32);
            for (int q = 0; q < kMaxPrim; ++q)
                for (int f = 0; f < nfield; ++f) {
                    double* dst = &pl.soa[((size_t)q * nfield + f) * pl.npad];
                    for (int k = k0; k < k1; ++k) dst[k] = pl.aos[((size_t)k * kMaxPrim + q) * nfield + f];
                }
        }
    }
    stage("sort + layout");
    return MYQC_OK;
}

PairList sublist(const PairList& src, const std::vector<char>& pred) {
    PairList d;
    d.type = src.type;
    const int nf = pt_nf(src.type), nfield = pt_nfield(src.type);
    std::vector<int> idx;
    for (int k = 0; k < src.n; ++k)
        if (pred[k]) idx.push_back(k);
    d.n = (int)idx.size();
    d.npad = (d.n + 31) / 32 * 32;
    d.emax.resize(d.n); d.nprim.resize(d.n); d.bucket.resize(d.n);
    if (!src.qmax.empty()) d.qmax.resize(d.n);
    d.shA.resize(d.n); d.shB.resize(d.n); d.owner_fn.resize(d.n);
    d.pidx.resize((size_t)d.n * nf);
    d.aos.assign((size_t)d.n * kMaxPrim * nfield, 0.0);
    d.soa.assign((size_t)kMaxPrim * nfield * d.npad, 0.0);
    for (int k = 0; k < d.n; ++k) {
        const int s = idx[k];
        d.emax[k] = src.emax[s]; d.nprim[k] = src.nprim[s]; d.bucket[k] = src.bucket[s];
        if (!src.qmax.empty()) d.qmax[k] = src.qmax[s];
        d.shA[k] = src.shA[s]; d.shB[k] = src.shB[s]; d.owner_fn[k] = src.owner_fn[s];
        for (int f = 0; f < nf; ++f) d.pidx[(size_t)k * nf + f] = src.pidx[(size_t)s * nf + f];
        std::memcpy(&d.aos[(size_t)k * kMaxPrim * nfield], &src.aos[(size_t)s * kMaxPrim * nfield],
                    sizeof(double) * kMaxPrim * nfield);
        for (int q = 0; q < kMaxPrim; ++q)
            for (int f = 0; f < nfield; ++f)
                d.soa[((size_t)q * nfield + f) * d.npad + k] = src.soa[((size_t)q * nfield + f) * src.npad + s];
    }
    return d;
}

int emax_bucket(double emax) {
    // Fine groups: symmetry-equivalent pairs share emax up to rounding noise, so a 1e-6 relative
    // grid keeps pairs of one kind (same primitive survival pattern) together while the Morton
    // order inside a group keeps them spatially close.
    if (!(emax > 0.0)) return 1 << 30;
    const double b = std::floor(-std::log(emax) * 1.0e6);
    return b < 0 ? 0 : (b > 1.0e9 ? 1000000000 : (int)b);
}

std::vector<int32_t> row_prefix(const PairList& U, const PairList& T, double tau) {
    // suffix maximum of T.emax is non-increasing: binary search for the last index that can still
    // hold a pair passing emax_u*emax_v >= 1e-14; same for the Schwarz factors when they are there
    std::vector<double> smax(T.n + 1, 0.0);
    for (int k = T.n - 1; k >= 0; --k) smax[k] = std::max(smax[k + 1], T.emax[k]);
    const bool schwarz = tau > 0.0 && (int)U.qmax.size() == U.n && (int)T.qmax.size() == T.n;
    std::vector<double> sq(T.n + 1, 0.0);
    if (schwarz)
        for (int k = T.n - 1; k >= 0; --k) sq[k] = std::max(sq[k + 1], T.qmax[k]);
    std::vector<int32_t> out(U.n, 0);
    for (int u = 0; u < U.n; ++u) {
        int lo = 0, hi = T.n;  // first index with emax_u*smax < 1e-14
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            if (U.emax[u] * smax[mid] < 1.0e-14) hi = mid; else lo = mid + 1;
        }
        if (schwarz) {
            int l2 = 0, h2 = lo;  // first index with qmax_u*sq < tau
            while (l2 < h2) {
                const int mid = (l2 + h2) / 2;
                if (U.qmax[u] * sq[mid] < tau) h2 = mid; else l2 = mid + 1;
            }
            lo = l2;
        }
        out[u] = lo;
    }
    return out;
}

}  // namespace myqc
