// C-ABI of the ERI engine (include/myqc_eri.h): planner (owner-row strips, tasks, shards), plans,
// one-shot host-buffer calls.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/myqc_eri.h"
#include "eri_kernels.cuh"
#include "pairs.hpp"
#include "strip_geom.hpp"

namespace myqc {

thread_local std::string g_last_error;
thread_local int64_t g_last_d2h_bytes = 0;

static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
int fock_fail(int code, const std::string& msg) { return fail(code, msg); }  // for fock.cu, int1e.cu

// Boys tables shared by the ERI and the one-electron kernels (boys.cuh): taylor[qi][t][k] =
// Ft(t, 3 qi + k)/k! for k < 7 (the bytes of the Ftab file), ex[k] = {exp(-k/10), k/10}
void build_boys_tables(const double* ftab, std::vector<double>& h, std::vector<double>& ex) {
    h.assign(5 * 121 * 8, 0.0);
    for (int qi = 0; qi < 5; ++qi)
        for (int t = 0; t <= 120; ++t) {
            double fact = 1.0;
            for (int k = 0; k < 7; ++k) {
                if (k > 1) fact *= k;
                h[((size_t)qi * 121 + t) * 8 + k] = ftab[t + 121 * (3 * qi + k)] / fact;
            }
        }
    ex.assign(2 * 608, 0.0);
    for (int k = 0; k <= 600; ++k) { ex[2 * k] = std::exp(-(k / 10.0)); ex[2 * k + 1] = k / 10.0; }
}
static int cuda_fail(cudaError_t e, const char* what) {
    return fail(MYQC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                              \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

// model flops per canonical primitive quartet, SURVEY.md 8d
static const double kW[6] = {60, 99, 228, 228, 693, 2691};
static int class_id(int la, int lb) {
    if (la > lb) std::swap(la, lb);
    static const int id[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    return id[la][lb];
}

// ---------------------------------------------------------------------------------------------
// The host plan: everything the kernels read, as host vectors.  Pure host arithmetic (no device): the
// shard layout and the coverage check use it without a GPU.
struct LaunchH {
    int UT = 0, TC = 0;
    std::vector<int4> tasks;
};

struct HostPlan {
    int ns = 0, norb = 0, nblk = 0;
    int64_t npair = 0;
    std::vector<Shell> shells;
    PairList lists[3];
    std::vector<int32_t> sh_fn;     // [ns][4]
    std::vector<int32_t> sh_first;  // [ns+1]
    std::vector<signed char> sh_type;
    std::vector<int32_t> pair_rec;  // [ns*ns]
    std::vector<double> pair_emax;  // [ns*ns]
    std::vector<double> blk_emax;   // [ns*nblk]
    std::vector<int32_t> clist[2];  // shells of type 0 / 1, ascending
    // shell classes (same type and exponents) and the partner segments (eri_kernels.cuh)
    int ncls = 0;
    std::vector<int32_t> sh_class, cls_kind, seg_start;
    std::vector<unsigned short> seg_d;
    std::vector<double> seg_eprof;  // [entries][4]
    std::vector<double> shell_cost; // estimated seconds of one warp for everything first shell A owns
    // shard
    int A0 = 0, A1 = 0;             // owner first shells [A0, A1)
    int64_t out_offset = 0, out_elems = 0;
    LaunchH launches[6];            // (UT,TC) = (0,0) (0,1) (1,0) (1,1) (2,0) (2,1)
};

// packed index of the first element of the first row whose leading orbital is `fn`
static int64_t packed_row_offset(int64_t fn, int64_t norb) {
    const int64_t np = norb * (norb + 1) / 2;
    if (fn >= norb) return np * (np + 1) / 2;
    const int64_t P = fn * norb - fn * (fn - 1) / 2;  // P(fn,fn)
    return P * np - P * (P - 1) / 2;
}

// Work model behind the task sizes and the shard cuts.  One warp runs a task; a task should take
// a small fraction of the launch.  Rates are per warp, for a GPU running ~2400 warps: the same
// constants on every rank (the shard layout is pure host arithmetic), and only ratios matter.
constexpr double kWarpFlops = 3.5e9;   // model flops per second and warp
constexpr double kWarpBytes = 2.0e9;   // packed-array bytes written per second and warp

struct CostModel {
    // partner kind k (0..2), first-shell type tc: emax of the pairs sorted descending, with prefix sums of
    // their primitive counts; cnt_ge[k][tc][A] = number of kind-k pairs whose first shell is >= A and of type tc
    std::vector<double> emax[3];
    std::vector<double> primsum[3];
    std::vector<int32_t> cnt_ge[3][2];
    int ntot[3] = {0, 0, 0};
    void build(const HostPlan& hp) {
        for (int k = 0; k < 3; ++k) {
            const PairList& L = hp.lists[k];
            std::vector<int> order(L.n);
            for (int i = 0; i < L.n; ++i) order[i] = i;
            std::sort(order.begin(), order.end(), [&](int x, int y) { return L.emax[x] > L.emax[y]; });
            emax[k].resize(L.n);
            primsum[k].assign(L.n + 1, 0.0);
            for (int i = 0; i < L.n; ++i) {
                emax[k][i] = L.emax[order[i]];
                primsum[k][i + 1] = primsum[k][i] + L.nprim[order[i]];
            }
            ntot[k] = L.n;
            for (int tc = 0; tc < 2; ++tc) cnt_ge[k][tc].assign(hp.ns + 1, 0);
            for (int i = 0; i < L.n; ++i) cnt_ge[k][hp.sh_type[L.shA[i]]][L.shA[i]] += 1;
            for (int tc = 0; tc < 2; ++tc)
                for (int A = hp.ns - 1; A >= 0; --A) cnt_ge[k][tc][A] += cnt_ge[k][tc][A + 1];
        }
    }
    // estimated number of kind-k partners (first-shell type tc, first shell >= A) that pass the bound against emax_u
    double partners(double emax_u, int A, int k, int tc) const {
        if (ntot[k] == 0 || !(emax_u > 0.0)) return 0.0;
        size_t lo = 0, hi = emax[k].size();
        while (lo < hi) {
            const size_t mid = (lo + hi) / 2;
            if (emax_u * emax[k][mid] < 1.0e-14) hi = mid; else lo = mid + 1;
        }
        return (double)lo * (double)cnt_ge[k][tc][A] / (double)ntot[k];
    }
    // estimated primitive quartets of owner (emax_u, nprim_u, first shell A) against the kind-k partners whose
    // first shell is of type tc: the partners that pass the bound, taken as an unbiased sample of first shells
    double quartets(double emax_u, int nprim_u, int A, int k, int tc) const {
        if (ntot[k] == 0 || !(emax_u > 0.0)) return 0.0;
        size_t lo = 0, hi = emax[k].size();
        while (lo < hi) {  // first index with emax_u*emax < 1e-14
            const size_t mid = (lo + hi) / 2;
            if (emax_u * emax[k][mid] < 1.0e-14) hi = mid; else lo = mid + 1;
        }
        return (double)nprim_u * primsum[k][lo] * (double)cnt_ge[k][tc][A] / (double)ntot[k];
    }
};

static int build_host_tables(int nnuc, const double* xyz, int nset, int setl, const double* set,
                             const int32_t* setinfo, int ops, const double* bas, const int32_t* basinfo,
                             HostPlan& hp, std::string& err) {
    int rc;
    if ((rc = build_shells(nnuc, nset, setl, setinfo, ops, basinfo, hp.shells, err))) return rc;
    hp.ns = (int)hp.shells.size();
    hp.norb = basinfo[1];
    hp.npair = (int64_t)hp.norb * (hp.norb + 1) / 2;
    if (hp.ns >= 32768 || hp.norb >= 32768) { err = "more than 32767 shells or orbitals"; return MYQC_ERR_UNSUPPORTED; }
    {
        int covered = 0;
        for (const Shell& sh : hp.shells) covered += sh.end_fn - sh.first_fn;
        if (covered != hp.norb) { err = "the sets do not cover orbitals 0..norb-1 exactly once"; return MYQC_ERR_BAD_ARG; }
    }
    if ((rc = build_pairs(nnuc, xyz, set, setinfo, setl, ops, bas, basinfo, hp.shells, hp.lists, err))) return rc;
    const int ns = hp.ns;
    hp.nblk = (ns + kBlockShells - 1) / kBlockShells;
    hp.sh_fn.assign((size_t)ns * 4, -1);
    hp.sh_first.assign(ns + 1, hp.norb);
    hp.sh_type.assign(ns, 0);
    for (int s = 0; s < ns; ++s) {
        for (int k = 0; k < 4; ++k) hp.sh_fn[(size_t)s * 4 + k] = hp.shells[s].fn[k];
        hp.sh_first[s] = hp.shells[s].first_fn;
        hp.sh_type[s] = (signed char)hp.shells[s].type;
        hp.clist[hp.shells[s].type].push_back(s);
    }
    hp.pair_rec.assign((size_t)ns * ns, -1);
    hp.pair_emax.assign((size_t)ns * ns, 0.0);
    hp.blk_emax.assign((size_t)ns * hp.nblk, 0.0);
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < hp.lists[k].n; ++i) {
            const int A = hp.lists[k].shA[i], B = hp.lists[k].shB[i];
            hp.pair_rec[(size_t)A * ns + B] = i;
            hp.pair_emax[(size_t)A * ns + B] = hp.lists[k].emax[i];
            double& bm = hp.blk_emax[(size_t)A * hp.nblk + B / kBlockShells];
            bm = std::max(bm, hp.lists[k].emax[i]);
        }
    // shell classes: one per shell type (S, SP): a segment lists the partners of one kind of a (C, block).
    // (The survival bins of the kernel, not the class, make the lanes of a chunk alike.)
    {
        hp.sh_class.assign(ns, 0);
        for (int s = 0; s < ns; ++s) hp.sh_class[s] = hp.shells[s].type;
        hp.ncls = 2;
        hp.cls_kind = {0, 1};
    }
    // partner segments: (C, block, class) -> the D >= C of that class in that block with a live pair, by decreasing emax
    {
        const int nblk = hp.nblk, ncls = hp.ncls;
        hp.seg_start.assign((size_t)ns * nblk * ncls + 1, 0);
        hp.seg_d.clear(); hp.seg_eprof.clear();
        std::vector<std::vector<std::pair<double, int>>> bucket(ncls);
        for (int C = 0; C < ns; ++C)
            for (int b = 0; b < nblk; ++b) {
                for (auto& v : bucket) v.clear();
                for (int D = std::max(C, b * kBlockShells); D < std::min(ns, (b + 1) * kBlockShells); ++D)
                    if (hp.pair_rec[(size_t)C * ns + D] >= 0) bucket[hp.sh_class[D]].push_back({hp.pair_emax[(size_t)C * ns + D], D});
                for (int c = 0; c < ncls; ++c) {
                    std::stable_sort(bucket[c].begin(), bucket[c].end(), [](const std::pair<double, int>& x, const std::pair<double, int>& y) { return x.first > y.first; });
                    hp.seg_start[((size_t)C * nblk + b) * ncls + c] = (int32_t)hp.seg_d.size();
                    for (const auto& e : bucket[c]) {
                        hp.seg_d.push_back((unsigned short)e.second);
                        // prefactors of ranks 1, 3, 5, 7 (the primitives of a record are sorted by E, descending)
                        const int kind = hp.sh_type[C] + hp.sh_type[e.second];
                        const int rec = hp.pair_rec[(size_t)C * ns + e.second];
                        const PairList& L = hp.lists[kind];
                        const int nfield = pt_nfield(kind);
                        for (int r = 0; r < 8; r += 2)
                            hp.seg_eprof.push_back(r < L.nprim[rec] ? L.aos[((size_t)rec * kMaxPrim + r) * nfield + 4] : 0.0);
                    }
                }
            }
        hp.seg_start.back() = (int32_t)hp.seg_d.size();
    }
    return MYQC_OK;
}

// seconds of one warp for owner pair (A,B) against the partner first shells of type tc (launch (UT,tc))
static double owner_cost(const HostPlan& hp, const CostModel& cm, int A, int B, int tc, double* bytes_out) {
    const int ns = hp.ns;
    const int UT = hp.sh_type[A] + hp.sh_type[B];
    const int rec = hp.pair_rec[(size_t)A * ns + B];
    // rows of the owner
    int nrows = 0;
    for (int mu = 0; mu < 4; ++mu)
        for (int nu = 0; nu < 4; ++nu) {
            const int i = hp.sh_fn[(size_t)A * 4 + mu], j = hp.sh_fn[(size_t)B * 4 + nu];
            if (i < 0 || j < 0 || (A == B && i > j)) continue;
            ++nrows;
        }
    // columns whose first index lies in a type-tc shell C >= A (upper bound: the runs l >= k)
    double cols = 0.0;
    {
        const std::vector<int32_t>& cl = hp.clist[tc];
        const size_t c0 = std::lower_bound(cl.begin(), cl.end(), A) - cl.begin();
        for (size_t c = c0; c < cl.size(); ++c)
            for (int kc = 0; kc < 4; ++kc) {
                const int k = hp.sh_fn[(size_t)cl[c] * 4 + kc];
                if (k >= 0) cols += hp.norb - k;
            }
    }
    const double bytes = 8.0 * nrows * cols;
    if (bytes_out) *bytes_out = bytes;
    double flops = 0.0;
    if (rec >= 0) {
        const PairList& L = hp.lists[UT];
        for (int td = 0; td < 2; ++td) {
            const int k = tc + td;
            flops += kW[class_id(UT, k)] * cm.quartets(L.emax[rec], L.nprim[rec], A, k, tc);
        }
    }
    return flops / kWarpFlops + bytes / kWarpBytes;
}

static void build_shell_costs(HostPlan& hp, const CostModel& cm) {
    hp.shell_cost.assign(hp.ns, 0.0);
    for (int A = 0; A < hp.ns; ++A)
        for (int B = A; B < hp.ns; ++B)
            for (int tc = 0; tc < 2; ++tc) hp.shell_cost[A] += owner_cost(hp, cm, A, B, tc, nullptr);
}

// cut [0, ns) into m contiguous ranges of first shells of ~equal cost; returns m+1 shell indices
static std::vector<int> split_shells(const std::vector<double>& w, int m) {
    const int n = (int)w.size();
    std::vector<int> bound(m + 1, n);
    bound[0] = 0;
    double tot = 0;
    for (double x : w) tot += x;
    double acc = 0;
    int sidx = 1;
    for (int c = 0; c < n && sidx < m; ++c) {
        acc += w[c];
        while (sidx < m && acc >= tot * sidx / m) bound[sidx++] = c + 1;
    }
    for (int k = 1; k <= m; ++k) bound[k] = std::max(bound[k], bound[k - 1]);
    bound[m] = n;
    return bound;
}

// tasks of the shard [A0, A1): every owner pair (dead ones included: their rows are zeros that still
// have to be written), against every partner first shell C >= A, cut into pieces of bounded cost
static void build_tasks(HostPlan& hp, const CostModel& cm) {
    const int ns = hp.ns, nblk = hp.nblk;
    // A task is one warp's unit of work.  Its quartets are evaluated 32 at a time per pending list (partner kind x
    // survival bin), and the last chunk of every list is partly empty, so a task must hold many quartets: the
    // targets below are partners per task (estimated from the sorted prefactors).  Tasks with few partners
    // are mostly zero fill and are bounded by the bytes they write instead.
    const double kItemsLight = std::getenv("MYQC_TASK_ITEMS") ? std::atof(std::getenv("MYQC_TASK_ITEMS")) : 1024.0;
    const double kItemsHeavy = std::getenv("MYQC_TASK_ITEMS_HEAVY") ? std::atof(std::getenv("MYQC_TASK_ITEMS_HEAVY")) : 320.0;
    const double kTaskBytes = 4.0e6;
    for (int UT = 0; UT < 3; ++UT)
        for (int tc = 0; tc < 2; ++tc) {
            LaunchH& L = hp.launches[UT * 2 + tc];
            L.UT = UT; L.TC = tc;
            L.tasks.clear();
        }
    for (int A = hp.A0; A < hp.A1; ++A)
        for (int B = A; B < ns; ++B) {
            const int UT = hp.sh_type[A] + hp.sh_type[B];
            for (int tc = 0; tc < 2; ++tc) {
                const std::vector<int32_t>& cl = hp.clist[tc];
                const int c0 = (int)(std::lower_bound(cl.begin(), cl.end(), A) - cl.begin());
                const int nc = (int)cl.size() - c0;
                if (nc <= 0) continue;
                std::vector<int4>& tasks = hp.launches[UT * 2 + tc].tasks;
                double bytes = 0.0;
                owner_cost(hp, cm, A, B, tc, &bytes);
                const int rec = hp.pair_rec[(size_t)A * ns + B];
                double items = 0.0;
                if (rec >= 0)
                    for (int td = 0; td < 2; ++td) items += cm.partners(hp.lists[UT].emax[rec], A, tc + td, tc);
                const bool heavy = (UT + tc >= 2) && !(UT == 0);  // launches whose second partner kind has 64 integrals per quartet
                int npiece = (int)std::max(std::floor(items / (heavy ? kItemsHeavy : kItemsLight)), std::ceil(bytes / kTaskBytes));
                npiece = std::max(1, npiece);
                const int ab = A | (B << 16);
                if (npiece <= nc) {
                    for (int p = 0; p < npiece; ++p) {
                        const int lo = c0 + (int)((int64_t)nc * p / npiece), hi = c0 + (int)((int64_t)nc * (p + 1) / npiece);
                        if (hi > lo) tasks.push_back(make_int4(ab, lo, hi, (cl[lo] / kBlockShells) | (nblk << 16)));
                    }
                } else {
                    // more pieces than first shells: cut the partner blocks of each C too
                    const int per_c = (npiece + nc - 1) / nc;
                    for (int c = c0; c < c0 + nc; ++c) {
                        const int bfirst = cl[c] / kBlockShells, nb = nblk - bfirst;
                        const int np = std::min(per_c, nb);
                        for (int p = 0; p < np; ++p) {
                            const int lo = bfirst + (int)((int64_t)nb * p / np), hi = bfirst + (int)((int64_t)nb * (p + 1) / np);
                            if (hi > lo) tasks.push_back(make_int4(ab, c, c + 1, lo | (hi << 16)));
                        }
                    }
                }
            }
        }
}

static int build_host_plan(int nnuc, const double* xyz, int nset, int setl, const double* set,
                           const int32_t* setinfo, int ops, const double* bas, const int32_t* basinfo,
                           int shard, int nshards, HostPlan& hp, std::string& err) {
    int rc = build_host_tables(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, hp, err);
    if (rc) return rc;
    CostModel cm;
    cm.build(hp);
    build_shell_costs(hp, cm);
    const std::vector<int> cuts = split_shells(hp.shell_cost, nshards);
    hp.A0 = cuts[shard];
    hp.A1 = cuts[shard + 1];
    hp.out_offset = packed_row_offset(hp.sh_first[hp.A0], hp.norb);
    hp.out_elems = packed_row_offset(hp.sh_first[hp.A1], hp.norb) - hp.out_offset;
    build_tasks(hp, cm);
    return MYQC_OK;
}

// canonical primitive-quartet statistics: unordered primitive pairs {a<=b},
// unordered pairs of pairs, kept iff EIJ*EGH >= 1e-14.
static std::vector<int32_t> prefix_counts(const std::vector<double>& eu, const std::vector<double>& et) {
    std::vector<int32_t> out(eu.size());
    for (size_t u = 0; u < eu.size(); ++u) {
        const double e = eu[u];
        size_t lo = 0, hi = et.size();
        while (lo < hi) {  // first index with e*et < 1e-14
            const size_t mid = (lo + hi) / 2;
            if (e * et[mid] < 1.0e-14) hi = mid; else lo = mid + 1;
        }
        out[u] = (int32_t)lo;
    }
    return out;
}

static void canonical_stats(int nnuc, const double* xyz, int nset, int setl, const double* set,
                            const int32_t* setinfo, int64_t nq[6], double* flops) {
    std::vector<double> E[3];
    for (int a = 0; a < nset; ++a)
        for (int b = a; b < nset; ++b) {
            const double aa = set[a], bb = set[b];
            const int u = setinfo[1 + a * setl + 3], v = setinfo[1 + b * setl + 3];
            double r2 = 0;
            for (int i = 0; i < 3; ++i) { const double d = xyz[u + nnuc * i] - xyz[v + nnuc * i]; r2 += d * d; }
            const double e = std::exp(-aa * bb * r2 / (aa + bb));
            if (e < 1.0e-14) continue;
            E[setinfo[1 + a * setl + 2] + setinfo[1 + b * setl + 2]].push_back(e);
        }
    for (int t = 0; t < 3; ++t) std::sort(E[t].begin(), E[t].end(), std::greater<double>());
    for (int c = 0; c < 6; ++c) nq[c] = 0;
    for (int ta = 0; ta < 3; ++ta)
        for (int tb = ta; tb < 3; ++tb) {
            const std::vector<int32_t> cnt = prefix_counts(E[ta], E[tb]);
            int64_t n = 0;
            if (ta == tb) {
                for (size_t u = 0; u < cnt.size(); ++u)
                    if ((int64_t)cnt[u] > (int64_t)u) n += cnt[u] - (int64_t)u;  // v >= u
            } else {
                for (size_t u = 0; u < cnt.size(); ++u) n += cnt[u];
            }
            nq[class_id(ta, tb)] += n;
        }
    *flops = 0;
    for (int c = 0; c < 6; ++c) *flops += kW[c] * (double)nq[c];
}

static int check_args(int nnuc, const double* xyz, int nset, int setl, const double* set,
                      const int32_t* setinfo, int ops, const double* bas, const int32_t* basinfo,
                      const double* ftab) {
    if (!xyz || !set || !setinfo || !bas || !basinfo || !ftab) return fail(MYQC_ERR_BAD_ARG, "null input pointer");
    if (nnuc < 1 || nset < 1) return fail(MYQC_ERR_BAD_ARG, "nnuc/nset must be positive");
    if (setinfo[0] != nset || setinfo[1] != setl) return fail(MYQC_ERR_BAD_ARG, "setinfo header does not match nset/setl");
    if (basinfo[0] != ops || basinfo[1] < 1) return fail(MYQC_ERR_BAD_ARG, "basinfo header does not match ops/norb");
    for (int s = 0; s < nset; ++s) {
        const int c = setinfo[1 + s * setl + 3];
        if (c < 0 || c >= nnuc) return fail(MYQC_ERR_BAD_ARG, "set centre out of range");
        if (!(set[s] > 0.0)) return fail(MYQC_ERR_BAD_ARG, "non-positive exponent");
    }
    return MYQC_OK;
}

// ---------------------------------------------------------------------------------------------
// Coverage check of a host plan (no device): replays what the kernels write -- the zero-filled spans of
// every task, cut at pseudo-random flush points, and the integrals of every partner that passes the
// pair-level bound -- with the same index arithmetic (strip_geom.hpp), and verifies that every element of
// the shard's slice is zero-filled exactly once, that every integral lands inside its own task's span,
// and that every integral that should exist is stored exactly once.
struct CheckResult { int64_t zero_elems = 0, value_elems = 0, expected_values = 0, errors = 0; };

static void check_plan(const HostPlan& hp, CheckResult& cr) {
    const int ns = hp.ns, nblk = hp.nblk, norb = hp.norb;
    std::vector<unsigned char> z((size_t)hp.out_elems, 0), val((size_t)hp.out_elems, 0);
    uint32_t rng = 12345u;
    auto rnd = [&]() { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
    for (int li = 0; li < 6; ++li) {
        const LaunchH& L = hp.launches[li];
        const std::vector<int32_t>& cl = hp.clist[L.TC];
        const int nslice = (L.UT == 2 && L.TC == 1) ? 4 : 1;
        for (int slice_i = 0; slice_i < nslice; ++slice_i) {
            const int USL = nslice == 4 ? slice_i : -1;
            const int NFU = USL >= 0 ? 4 : pt_nf(L.UT);
            for (const int4& t : L.tasks) {
                const int A = t.x & 0xffff, B = t.x >> 16, c_lo = t.y, c_hi = t.z, b_lo = t.w & 0xffff, b_hi = t.w >> 16;
                const int* fa = &hp.sh_fn[(size_t)A * 4];
                const int* fb = &hp.sh_fn[(size_t)B * 4];
                const bool a_is_sp = (L.UT == 1) && hp.sh_type[A] == 1;
                const double eu = hp.pair_emax[(size_t)A * ns + B];
                // spans: consecutive positions (ci, b), cut at random places like the kernel's flushes
                int zci = c_lo, zb = b_lo;
                auto zero_span = [&](int pci, int pb) {
                    for (int f = 0; f < NFU; ++f) {
                        int i, j;
                        if (!owner_row(L.UT, USL, a_is_sp, A == B, fa, fb, f, &i, &j)) continue;
                        const int64_t P1 = pair_index64(i, j, norb);
                        const int64_t rb = row_start64(P1, hp.npair) - P1 - hp.out_offset;
                        for (int c2 = zci; c2 <= pci; ++c2) {
                            const int C2 = cl[c2];
                            const int bfirst = C2 / kBlockShells;
                            const int blo = (c2 == zci && zb >= 0) ? zb : bfirst;
                            const int bhi = (c2 == pci) ? pb : nblk - 1;
                            if (blo > bhi) continue;
                            const int llo = (blo == bfirst) ? 0 : hp.sh_first[blo * kBlockShells];
                            const int lhi = (bhi == nblk - 1) ? norb : hp.sh_first[(bhi + 1) * kBlockShells];
                            for (int kc = 0; kc < (L.TC ? 4 : 1); ++kc) {
                                const int k = hp.sh_fn[(size_t)C2 * 4 + kc];
                                if (k < 0) continue;
                                int l0;
                                const int n = row_piece(i, j, k, llo, lhi, &l0);
                                const int64_t addr = rb + col_base64(k, norb) + l0;
                                for (int x = 0; x < n; ++x) {
                                    if (addr + x < 0 || addr + x >= hp.out_elems) { ++cr.errors; continue; }
                                    if (z[(size_t)(addr + x)]++) ++cr.errors;
                                    ++cr.zero_elems;
                                }
                            }
                        }
                    }
                    if (pb >= nblk - 1) { zci = pci + 1; zb = -1; } else { zci = pci; zb = pb + 1; }
                };
                for (int ci = c_lo; ci < c_hi; ++ci) {
                    const int Cs = cl[ci];
                    const int bmin = (ci == c_lo) ? b_lo : Cs / kBlockShells;
                    const int bend = (ci == c_hi - 1) ? b_hi : nblk;
                    for (int b = bmin; b < bend; ++b) {
                        // partners of this block
                        for (int Ds = std::max(b * kBlockShells, Cs); Ds < std::min(ns, (b + 1) * kBlockShells); ++Ds) {
                            if (hp.pair_rec[(size_t)Cs * ns + Ds] < 0) continue;
                            if (!(eu * hp.pair_emax[(size_t)Cs * ns + Ds] >= 1.0e-14)) continue;
                            if (eu * hp.blk_emax[(size_t)Cs * nblk + b] < 1.0e-14) ++cr.errors;  // block bound must not hide it
                            const int TD = hp.sh_type[Ds];
                            const int* fc = &hp.sh_fn[(size_t)Cs * 4];
                            const int* fd = &hp.sh_fn[(size_t)Ds * 4];
                            for (int f = 0; f < NFU; ++f) {
                                int i, j;
                                if (!owner_row(L.UT, USL, a_is_sp, A == B, fa, fb, f, &i, &j)) continue;
                                const int64_t P1 = pair_index64(i, j, norb);
                                const int64_t rb = row_start64(P1, hp.npair) - P1 - hp.out_offset;
                                for (int kc = 0; kc < (L.TC ? 4 : 1); ++kc)
                                    for (int ld = 0; ld < (TD ? 4 : 1); ++ld) {
                                        const int k = fc[kc], l = fd[ld];
                                        if (k < 0 || l < 0 || k > l || k < i || (k == i && l < j)) continue;
                                        const int64_t addr = rb + col_base64(k, norb) + l;
                                        if (addr < 0 || addr >= hp.out_elems) { ++cr.errors; continue; }
                                        if (val[(size_t)addr]++) ++cr.errors;
                                        ++cr.value_elems;
                                    }
                            }
                        }
                        if (rnd() % 5 == 0) zero_span(ci, b);
                    }
                    if (rnd() % 3 == 0 && bend == nblk) zero_span(ci, nblk - 1);
                }
                zero_span(c_hi - 1, b_hi - 1);
            }
        }
    }
    for (int64_t e = 0; e < hp.out_elems; ++e)
        if (z[(size_t)e] != 1) ++cr.errors;
    // the partner segments list every live pair (C,D), D >= C, exactly once, in its block and class, by decreasing emax
    {
        std::vector<int> seen((size_t)ns * ns, 0);
        for (int Cs = 0; Cs < ns; ++Cs)
            for (int b = 0; b < nblk; ++b)
                for (int c = 0; c < hp.ncls; ++c) {
                    const size_t si = ((size_t)Cs * nblk + b) * hp.ncls + c;
                    for (int e = hp.seg_start[si]; e < hp.seg_start[si + 1]; ++e) {
                        const int Ds = hp.seg_d[e];
                        if (Ds < Cs || Ds / kBlockShells != b || hp.sh_class[Ds] != c) ++cr.errors;
                        if (hp.seg_eprof[4 * (size_t)e] != hp.pair_emax[(size_t)Cs * ns + Ds] || hp.pair_rec[(size_t)Cs * ns + Ds] < 0) ++cr.errors;
                        if (e > hp.seg_start[si] && hp.seg_eprof[4 * (size_t)e] > hp.seg_eprof[4 * (size_t)(e - 1)]) ++cr.errors;
                        for (int r = 1; r < 4; ++r)
                            if (hp.seg_eprof[4 * (size_t)e + r] > hp.seg_eprof[4 * (size_t)e + r - 1]) ++cr.errors;
                        if (hp.cls_kind[c] != hp.sh_type[Ds]) ++cr.errors;
                        ++seen[(size_t)Cs * ns + Ds];
                    }
                }
        for (int Cs = 0; Cs < ns; ++Cs)
            for (int Ds = Cs; Ds < ns; ++Ds)
                if (seen[(size_t)Cs * ns + Ds] != (hp.pair_rec[(size_t)Cs * ns + Ds] >= 0 ? 1 : 0)) ++cr.errors;
    }
    // expected: every element (P1 <= P2) of the slice whose shell pairs pass the bound
    std::vector<int> fn_shell(norb, -1);
    for (int s = 0; s < ns; ++s)
        for (int k = 0; k < 4; ++k)
            if (hp.sh_fn[(size_t)s * 4 + k] >= 0) fn_shell[hp.sh_fn[(size_t)s * 4 + k]] = s;
    if (hp.out_elems <= (int64_t)40000000) {
        for (int i = 0; i < norb; ++i)
            for (int j = i; j < norb; ++j) {
                const int64_t P1 = pair_index64(i, j, norb);
                const int64_t row0 = row_start64(P1, hp.npair) - hp.out_offset;
                if (row0 < 0 || row0 >= hp.out_elems) continue;
                const double eu = hp.pair_emax[(size_t)fn_shell[i] * ns + fn_shell[j]];
                for (int k = i; k < norb; ++k)
                    for (int l = (k == i ? j : k); l < norb; ++l) {
                        const double ev = hp.pair_emax[(size_t)fn_shell[k] * ns + fn_shell[l]];
                        const bool want = eu > 0.0 && ev > 0.0 && eu * ev >= 1.0e-14;
                        const int64_t addr = row0 + (pair_index64(k, l, norb) - P1);
                        if (want) ++cr.expected_values;
                        if ((val[(size_t)addr] != 0) != want) ++cr.errors;
                    }
            }
    } else {
        cr.expected_values = -1;
    }
}

}  // namespace myqc

using namespace myqc;

struct LaunchD {
    int UT = 0, TC = 0, slice = 0;
    StripArgs args;
};

struct myqc_eri_plan {
    int device = 0, num_sms = 0;
    int norb = 0, nset = 0;
    int64_t npair = 0;
    int64_t out_offset = 0, out_elems = 0;
    std::vector<void*> dev_allocs;
    std::vector<LaunchD> launches;  // one per kernel launch of an execute()
    int* d_counters = nullptr;                 // one task counter per launch
    unsigned long long* d_stats = nullptr;     // two counters per launch: primitive quartets evaluated, TD = 0 / 1
    // internal streams: launches are independent (disjoint runs of the array), their tails overlap
    static constexpr int kNumCompute = 4;
    cudaStream_t s_comp[kNumCompute] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t e_start = nullptr, e_done[kNumCompute] = {nullptr, nullptr, nullptr, nullptr};
    // stats (canonical primitive-quartet counts of the whole molecule, lazily evaluated)
    int64_t nquartets[6] = {0, 0, 0, 0, 0, 0};
    double model_flops = 0.0;
    int64_t h2d_bytes = 0;  // bytes uploaded at plan creation
    int64_t ntasks = 0;
    std::vector<double> in_xyz, in_set;
    std::vector<int32_t> in_setinfo;
    int in_nnuc = 0, in_setl = 0;
    bool stats_done = false, stats_whole = false;
};

namespace myqc {

template <class T>
static int upload(myqc_eri_plan* pl, const std::vector<T>& h, T** d) {
    *d = nullptr;
    if (h.empty()) return MYQC_OK;
    CU(cudaMalloc((void**)d, h.size() * sizeof(T)));
    pl->dev_allocs.push_back(*d);
    CU(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    pl->h2d_bytes += (int64_t)(h.size() * sizeof(T));
    return MYQC_OK;
}

}  // namespace myqc

extern "C" {

const char* myqc_last_error(void) { return g_last_error.c_str(); }
int64_t myqc_eri_last_d2h_bytes(void) { return myqc::g_last_d2h_bytes; }

int myqc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int myqc_eri_plan_create(int nnuc, const double* xyz, int nset, int setl, const double* set,
                         const int32_t* setinfo, int ops, const double* bas, const int32_t* basinfo,
                         const double* ftab, int device, int shard, int nshards, myqc_eri_plan** plan) {
    if (!plan) return fail(MYQC_ERR_BAD_ARG, "plan is null");
    *plan = nullptr;
    int rc = check_args(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, ftab);
    if (rc) return rc;
    if (nshards < 1 || shard < 0 || shard >= nshards) return fail(MYQC_ERR_BAD_ARG, "bad shard/nshards");
    const int ndev = myqc_device_count();
    if (ndev == 0) return fail(MYQC_ERR_NO_DEVICE, "no CUDA device: the ERI engine has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(MYQC_ERR_BAD_ARG, "device index out of range");
    CU(cudaSetDevice(device));
    {
        const int e = prepare_kernels();
        if (e) return cuda_fail((cudaError_t)e, "kernel preparation");
    }

    const bool trace = std::getenv("MYQC_TRACE") != nullptr;
    auto tprev = std::chrono::steady_clock::now();
    auto stage = [&](const char* name) {
        if (!trace) return;
        const auto t = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[myqc trace]   plan stage %-24s %.1f ms\n", name, std::chrono::duration<double, std::milli>(t - tprev).count());
        tprev = t;
    };
    std::unique_ptr<myqc_eri_plan> pl(new myqc_eri_plan());
    pl->device = device;
    CU(cudaDeviceGetAttribute(&pl->num_sms, cudaDevAttrMultiProcessorCount, device));
    pl->norb = basinfo[1];
    pl->nset = nset;
    pl->npair = (int64_t)pl->norb * (pl->norb + 1) / 2;

    std::string err;
    HostPlan hp;
    if ((rc = build_host_plan(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, shard, nshards, hp, err))) return fail(rc, err);
    stage("host plan (pairs, tasks)");
    pl->out_offset = hp.out_offset;
    pl->out_elems = hp.out_elems;

    // ---- upload ----------------------------------------------------------------------------------
    double *d_ftab = nullptr, *d_exptab = nullptr;
    {
        std::vector<double> h, ex;
        build_boys_tables(ftab, h, ex);
        if ((rc = upload(pl.get(), h, &d_ftab))) return rc;
        if ((rc = upload(pl.get(), ex, &d_exptab))) return rc;
    }
    int32_t *d_sh_fn = nullptr, *d_sh_first = nullptr, *d_pair_rec = nullptr, *d_clist[2] = {nullptr, nullptr};
    signed char* d_sh_type = nullptr;
    double *d_pair_emax = nullptr, *d_blk_emax = nullptr;
    if ((rc = upload(pl.get(), hp.sh_fn, &d_sh_fn))) return rc;
    if ((rc = upload(pl.get(), hp.sh_first, &d_sh_first))) return rc;
    if ((rc = upload(pl.get(), hp.sh_type, &d_sh_type))) return rc;
    if ((rc = upload(pl.get(), hp.pair_rec, &d_pair_rec))) return rc;
    if ((rc = upload(pl.get(), hp.pair_emax, &d_pair_emax))) return rc;
    if ((rc = upload(pl.get(), hp.blk_emax, &d_blk_emax))) return rc;
    for (int t = 0; t < 2; ++t)
        if ((rc = upload(pl.get(), hp.clist[t], &d_clist[t]))) return rc;
    int32_t *d_seg_start = nullptr, *d_cls_kind = nullptr;
    unsigned short* d_seg_d = nullptr;
    double* d_seg_eprof = nullptr;
    if ((rc = upload(pl.get(), hp.seg_start, &d_seg_start))) return rc;
    if ((rc = upload(pl.get(), hp.seg_d, &d_seg_d))) return rc;
    if ((rc = upload(pl.get(), hp.seg_eprof, &d_seg_eprof))) return rc;
    if ((rc = upload(pl.get(), hp.cls_kind, &d_cls_kind))) return rc;
    double* d_aos[3] = {nullptr, nullptr, nullptr};
    int32_t* d_nprim[3] = {nullptr, nullptr, nullptr};
    for (int k = 0; k < 3; ++k) {
        if ((rc = upload(pl.get(), hp.lists[k].aos, &d_aos[k]))) return rc;
        if ((rc = upload(pl.get(), hp.lists[k].nprim, &d_nprim[k]))) return rc;
    }
    stage("table upload");
    constexpr int kMaxLaunch = 16;
    {
        std::vector<int> zc(kMaxLaunch, 0);
        std::vector<unsigned long long> zs(2 * kMaxLaunch, 0ull);
        if ((rc = upload(pl.get(), zc, &pl->d_counters))) return rc;
        if ((rc = upload(pl.get(), zs, &pl->d_stats))) return rc;
    }
    // launches: the long tasks first (SP.SP owners), so that the launches that end on many short tasks come last
    static const int order[6] = {5, 4, 3, 2, 1, 0};
    for (int oi = 0; oi < 6; ++oi) {
        const LaunchH& L = hp.launches[order[oi]];
        if (L.tasks.empty()) continue;
        int4* d_tasks = nullptr;
        if ((rc = upload(pl.get(), L.tasks, &d_tasks))) return rc;
        for (int slice = 0; slice < strip_nslices(L.UT, L.TC); ++slice) {
            LaunchD D;
            D.UT = L.UT; D.TC = L.TC; D.slice = slice;
            StripArgs& a = D.args;
            std::memset(&a, 0, sizeof(a));
            a.sh_fn = reinterpret_cast<const int4*>(d_sh_fn);
            a.sh_type = d_sh_type;
            a.sh_first = d_sh_first;
            a.ns = hp.ns; a.nblk = hp.nblk; a.norb = hp.norb;
            a.npair = hp.npair;
            a.pair_rec = d_pair_rec; a.pair_emax = d_pair_emax; a.blk_emax = d_blk_emax;
            a.seg_start = d_seg_start; a.seg_d = d_seg_d; a.seg_eprof = reinterpret_cast<const double2*>(d_seg_eprof);
            a.cls_kind = d_cls_kind; a.ncls = hp.ncls;
            a.clist = d_clist[L.TC];
            a.u_aos = d_aos[L.UT]; a.u_nprim = d_nprim[L.UT];
            for (int td = 0; td < 2; ++td) {
                const int k = L.TC + td;
                a.t_aos[td] = d_aos[k]; a.t_nprim[td] = d_nprim[k];
                a.ftab_q[td] = d_ftab + (size_t)(L.UT + k) * 121 * 8;
            }
            a.exptab = reinterpret_cast<const double2*>(d_exptab);
            a.tasks = d_tasks;
            a.ntasks = (int)L.tasks.size();
            const int li = (int)pl->launches.size();
            if (li >= kMaxLaunch) return fail(MYQC_ERR_UNSUPPORTED, "too many launches in one plan");
            a.counter = pl->d_counters + li;
            a.stats = pl->d_stats + 2 * li;
            a.out = nullptr;
            a.out_offset = hp.out_offset;
            pl->launches.push_back(D);
            pl->ntasks += a.ntasks;
        }
    }
    stage("task upload");
    for (auto& st : pl->s_comp) CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&pl->e_start, cudaEventDisableTiming));
    for (auto& e : pl->e_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    stage("streams + events");
    // the canonical work statistics cost a few ms of host time: evaluated when plan_stats asks for them
    pl->stats_whole = (nshards == 1);
    pl->in_nnuc = nnuc; pl->in_setl = setl;
    pl->in_xyz.assign(xyz, xyz + 3 * (size_t)nnuc);
    pl->in_set.assign(set, set + nset);
    pl->in_setinfo.assign(setinfo, setinfo + 2 + (size_t)setl * nset);
    *plan = pl.release();
    return MYQC_OK;
}

int myqc_eri_canonical_stats(int nnuc, const double* xyz, int nset, int setl, const double* set,
                              const int32_t* setinfo, int64_t* nquartets, double* model_flops) {
    if (!xyz || !set || !setinfo || !nquartets || !model_flops) return fail(MYQC_ERR_BAD_ARG, "null pointer");
    canonical_stats(nnuc, xyz, nset, setl, set, setinfo, nquartets, model_flops);
    return MYQC_OK;
}

int myqc_eri_shard_layout(int nnuc, const double* xyz, int nset, int setl, const double* set,
                          const int32_t* setinfo, int ops, const double* bas, const int32_t* basinfo,
                          int nshards, int64_t* offsets) {
    // host only: no device is touched
    static const double dummy_ft[1] = {0.0};
    int rc = check_args(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, dummy_ft);
    if (rc) return rc;
    if (nshards < 1 || !offsets) return fail(MYQC_ERR_BAD_ARG, "bad nshards/offsets");
    std::string err;
    HostPlan hp;
    if ((rc = build_host_tables(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, hp, err))) return fail(rc, err);
    CostModel cm;
    cm.build(hp);
    build_shell_costs(hp, cm);
    const std::vector<int> cuts = split_shells(hp.shell_cost, nshards);
    for (int k = 0; k <= nshards; ++k) offsets[k] = packed_row_offset(hp.sh_first[cuts[k]], hp.norb);
    return MYQC_OK;
}

int myqc_eri_plan_check(int nnuc, const double* xyz, int nset, int setl, const double* set,
                        const int32_t* setinfo, int ops, const double* bas, const int32_t* basinfo,
                        int shard, int nshards, int64_t* result) {
    // host only: builds the plan of one shard and replays what its kernels would write
    static const double dummy_ft[1] = {0.0};
    int rc = check_args(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, dummy_ft);
    if (rc) return rc;
    if (nshards < 1 || shard < 0 || shard >= nshards || !result) return fail(MYQC_ERR_BAD_ARG, "bad shard/nshards/result");
    std::string err;
    HostPlan hp;
    if ((rc = build_host_plan(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, shard, nshards, hp, err))) return fail(rc, err);
    if (hp.out_elems > (int64_t)2000000000) return fail(MYQC_ERR_UNSUPPORTED, "slice too large for the host-side coverage check");
    CheckResult cr;
    check_plan(hp, cr);
    int64_t ntasks = 0;
    for (const LaunchH& L : hp.launches) ntasks += (int64_t)L.tasks.size() * strip_nslices(L.UT, L.TC);
    result[0] = cr.errors; result[1] = hp.out_elems; result[2] = cr.zero_elems; result[3] = cr.value_elems;
    result[4] = cr.expected_values; result[5] = ntasks;
    return MYQC_OK;
}

int64_t myqc_eri_plan_out_offset(const myqc_eri_plan* plan) { return plan ? plan->out_offset : -1; }
int64_t myqc_eri_plan_out_elems(const myqc_eri_plan* plan) { return plan ? plan->out_elems : -1; }

int myqc_eri_plan_execute(myqc_eri_plan* plan, double* d_out, void* stream) {
    if (!plan || (!d_out && plan->out_elems > 0)) return fail(MYQC_ERR_BAD_ARG, "null plan or output");
    CU(cudaSetDevice(plan->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nl = (int)plan->launches.size();
    if (nl == 0) return MYQC_OK;
    CU(cudaMemsetAsync(plan->d_counters, 0, sizeof(int) * (size_t)nl, st));
    CU(cudaMemsetAsync(plan->d_stats, 0, sizeof(unsigned long long) * 2 * (size_t)nl, st));
    // fork: internal streams start after whatever is already queued on the caller's stream
    CU(cudaEventRecord(plan->e_start, st));
    for (auto& sc : plan->s_comp) CU(cudaStreamWaitEvent(sc, plan->e_start, 0));
    for (int k = 0; k < nl; ++k) {
        LaunchD& L = plan->launches[k];
        StripArgs a = L.args;
        a.out = d_out;
        const int e = launch_strip(L.UT, L.TC, L.slice, a, plan->num_sms, plan->s_comp[k % myqc_eri_plan::kNumCompute]);
        if (e) return cuda_fail((cudaError_t)e, "strip kernel launch");
    }
    for (int i = 0; i < myqc_eri_plan::kNumCompute; ++i) {
        CU(cudaEventRecord(plan->e_done[i], plan->s_comp[i]));
        CU(cudaStreamWaitEvent(st, plan->e_done[i], 0));
    }
    return MYQC_OK;
}

int myqc_eri_plan_launch_count(const myqc_eri_plan* plan) { return plan ? (int)plan->launches.size() : 0; }

int myqc_eri_plan_launch_info(const myqc_eri_plan* plan, int k, int* ut, int* tc, int* slice, int64_t* ntasks) {
    if (!plan || k < 0 || k >= (int)plan->launches.size()) return fail(MYQC_ERR_BAD_ARG, "bad launch index");
    const LaunchD& L = plan->launches[k];
    if (ut) *ut = L.UT;
    if (tc) *tc = L.TC;
    if (slice) *slice = L.slice;
    if (ntasks) *ntasks = L.args.ntasks;
    return MYQC_OK;
}

int myqc_eri_plan_launch_quartets(myqc_eri_plan* plan, int64_t* nq) {
    if (!plan || !nq) return fail(MYQC_ERR_BAD_ARG, "null plan/output");
    CU(cudaSetDevice(plan->device));
    const size_t n = 2 * plan->launches.size();
    std::vector<unsigned long long> h(n);
    CU(cudaDeviceSynchronize());
    if (n) CU(cudaMemcpy(h.data(), plan->d_stats, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i) nq[i] = (int64_t)h[i];
    return MYQC_OK;
}

// Serialised pass on the caller's stream with CUDA events around every launch: the per-kernel
// durations behind bench.py's roofline (the production execute() overlaps launches on internal streams).
int myqc_eri_plan_execute_timed(myqc_eri_plan* plan, double* d_out, void* stream, float* ms) {
    if (!plan || !ms || (!d_out && plan->out_elems > 0)) return fail(MYQC_ERR_BAD_ARG, "null plan/output/ms");
    CU(cudaSetDevice(plan->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n = (int)plan->launches.size();
    if (n == 0) return MYQC_OK;
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) CU(cudaEventCreate(&e));
    CU(cudaMemsetAsync(plan->d_counters, 0, sizeof(int) * (size_t)n, st));
    CU(cudaMemsetAsync(plan->d_stats, 0, sizeof(unsigned long long) * 2 * (size_t)n, st));
    CU(cudaEventRecord(ev[0], st));
    for (int k = 0; k < n; ++k) {
        LaunchD& L = plan->launches[k];
        StripArgs a = L.args;
        a.out = d_out;
        const int e = launch_strip(L.UT, L.TC, L.slice, a, plan->num_sms, st);
        if (e) return cuda_fail((cudaError_t)e, "strip kernel launch");
        CU(cudaEventRecord(ev[k + 1], st));
    }
    CU(cudaEventSynchronize(ev[n]));
    for (int k = 0; k < n; ++k) CU(cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]));
    for (auto& x : ev) cudaEventDestroy(x);
    return MYQC_OK;
}

int myqc_fp64_peak(int device, double* tflops) {
    if (!tflops) return fail(MYQC_ERR_BAD_ARG, "null output");
    if (myqc_device_count() == 0) return fail(MYQC_ERR_NO_DEVICE, "no CUDA device");
    CU(cudaSetDevice(device));
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    double best = 0.0;
    int e = measure_dfma_peak(sms, &best);
    if (e) return cuda_fail((cudaError_t)e, "dfma peak kernel");
    *tflops = best;
    return MYQC_OK;
}

int myqc_eri_plan_stats(const myqc_eri_plan* plan_c, int64_t* nquartets, double* model_flops, int* nlaunch) {
    if (!plan_c) return fail(MYQC_ERR_BAD_ARG, "null plan");
    myqc_eri_plan* plan = const_cast<myqc_eri_plan*>(plan_c);
    if (!plan->stats_done) {
        if (plan->stats_whole)
            canonical_stats(plan->in_nnuc, plan->in_xyz.data(), plan->nset, plan->in_setl, plan->in_set.data(),
                            plan->in_setinfo.data(), plan->nquartets, &plan->model_flops);
        plan->stats_done = true;
    }
    if (nquartets) for (int c = 0; c < 6; ++c) nquartets[c] = plan->nquartets[c];
    if (model_flops) *model_flops = plan->model_flops;
    if (nlaunch) *nlaunch = (int)plan->launches.size();
    return MYQC_OK;
}

void myqc_eri_plan_destroy(myqc_eri_plan* plan) {
    if (!plan) return;
    cudaSetDevice(plan->device);
    for (auto& st : plan->s_comp) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    if (plan->e_start) cudaEventDestroy(plan->e_start);
    for (auto& e : plan->e_done) if (e) cudaEventDestroy(e);
    for (void* p : plan->dev_allocs) cudaFree(p);
    delete plan;
}

int myqc_eri_expand_dense(const double* d_packed, int norb, double* d_xx, void* stream) {
    if (!d_packed || !d_xx || norb < 1) return fail(MYQC_ERR_BAD_ARG, "bad arguments");
    int dev = 0, sms = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int e = launch_expand_dense(d_packed, norb, 0, norb, d_xx, sms, stream);
    if (e) return cuda_fail((cudaError_t)e, "expand_dense launch");
    return MYQC_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// One-shot calls with host buffers.  The plan and the device slice of the last call are kept per device
// (keyed on the bytes of every input): an SCF driver that asks for the same integrals again, or a
// benchmark loop, pays for the pair tables, the task lists and a 40 GB cudaMalloc/cudaFree once.
// myqc_eri_release_cache() drops them.
namespace myqc {

struct CacheEntry {
    uint64_t key = 0;
    myqc_eri_plan* plan = nullptr;
    double* d_out = nullptr;
    int64_t out_cap = 0;
    unsigned char* d_flags = nullptr;
    unsigned char* h_flags = nullptr;
    int64_t flags_cap = 0;
};
static std::mutex g_dev_mu[64];  // one one-shot call at a time per device: its cache entry is in use for the whole call
static CacheEntry g_cache[64];

static uint64_t fnv1a(uint64_t h, const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

static void drop_entry(CacheEntry& c) {
    if (c.plan) { myqc_eri_plan_destroy(c.plan); c.plan = nullptr; }
    if (c.d_out) { cudaFree(c.d_out); c.d_out = nullptr; c.out_cap = 0; }
    if (c.d_flags) { cudaFree(c.d_flags); c.d_flags = nullptr; }
    if (c.h_flags) { cudaFreeHost(c.h_flags); c.h_flags = nullptr; }
    c.flags_cap = 0;
    c.key = 0;
}

}  // namespace myqc

extern "C" {

void myqc_eri_release_cache(void) {
    int ndev = myqc_device_count();
    for (int d = 0; d < 64 && d < ndev; ++d) {
        std::lock_guard<std::mutex> lk(g_dev_mu[d]);
        if (!g_cache[d].plan && !g_cache[d].d_out) continue;
        cudaSetDevice(d);
        drop_entry(g_cache[d]);
    }
}

int myqc_eri_packed_shard(int nnuc, const double* xyz, int nset, int setl, const double* set,
                          const int32_t* setinfo, int ops, const double* bas, const int32_t* basinfo,
                          const double* ftab, double* packed_slice, int device, int shard, int nshards,
                          int64_t* h2d_bytes) {
    // MYQC_TRACE=1: wall-clock breakdown of the one-shot call on stderr
    const bool trace = std::getenv("MYQC_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    const auto t0 = now();
    int rc = check_args(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, ftab);
    if (rc) return rc;
    if (device < 0 || device >= 64) return fail(MYQC_ERR_BAD_ARG, "device index out of range");
    if (myqc_device_count() == 0) return fail(MYQC_ERR_NO_DEVICE, "no CUDA device: the ERI engine has no CPU fallback");
    static const bool no_cache = std::getenv("MYQC_NO_CACHE") != nullptr;
    uint64_t key = 1469598103934665603ull;
    {
        const int hdr[6] = {nnuc, nset, setl, ops, shard, nshards};
        key = fnv1a(key, hdr, sizeof(hdr));
        key = fnv1a(key, xyz, sizeof(double) * 3 * (size_t)nnuc);
        key = fnv1a(key, set, sizeof(double) * (size_t)nset);
        key = fnv1a(key, setinfo, sizeof(int32_t) * (2 + (size_t)setl * nset));
        key = fnv1a(key, bas, sizeof(double) * (size_t)ops * nset);
        key = fnv1a(key, basinfo, sizeof(int32_t) * (2 + 5 * (size_t)basinfo[1]));
        key = fnv1a(key, ftab, sizeof(double) * 2783);
        if (key == 0) key = 1;
    }
    std::lock_guard<std::mutex> dev_lock(g_dev_mu[device]);
    CacheEntry local;
    CacheEntry& ce = no_cache ? local : g_cache[device];
    bool hit = (ce.plan != nullptr && ce.key == key);
    if (!hit) {
        if (ce.plan) { cudaSetDevice(device); myqc_eri_plan_destroy(ce.plan); ce.plan = nullptr; ce.key = 0; }
        rc = myqc_eri_plan_create(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, ftab, device, shard, nshards, &ce.plan);
        if (rc) return rc;
        ce.key = key;
    }
    myqc_eri_plan* pl = ce.plan;
    CU(cudaSetDevice(device));
    const auto t1 = now();
    if (h2d_bytes) *h2d_bytes = hit ? 0 : pl->h2d_bytes;
    const int64_t n = pl->out_elems;
    if (n > 0 && !packed_slice) return fail(MYQC_ERR_BAD_ARG, "null output");
    if (n > ce.out_cap) {
        if (ce.d_out) { cudaFree(ce.d_out); ce.d_out = nullptr; ce.out_cap = 0; }
        cudaError_t e = cudaMalloc((void**)&ce.d_out, (size_t)n * sizeof(double));
        if (e != cudaSuccess) {
            if (no_cache) drop_entry(local);
            return fail(MYQC_ERR_NOMEM, std::string("device allocation of the packed slice: ") + cudaGetErrorString(e));
        }
        ce.out_cap = n;
    }
    double* d_out = ce.d_out;
    const auto t2 = now();
    rc = myqc_eri_plan_execute(pl, d_out, nullptr);
    if (trace) cudaDeviceSynchronize();
    const auto t3 = now();
    const char* xfer_kind = "cudaMemcpy";
    double xfer_frac = 1.0;
    if (!rc && n > 0) {
        // Destination in device-accessible (pinned) host memory and a slice worth the trouble: sparse
        // transfer.  ~90 % of a large molecule's integrals are exact zeros, so only the chunks that hold a
        // nonzero cross PCIe (stored by the GPU straight into the host buffer) while host threads write the
        // zeros of the other chunks locally.  Everything else: one cudaMemcpy (pinned destinations run at
        // PCIe speed; pageable ones are staged by the driver).  MYQC_SPARSE_D2H=0 forces the plain copy.
        cudaPointerAttributes pa{};
        const char* envs = std::getenv("MYQC_SPARSE_D2H");
        const bool want = !(envs && envs[0] == '0') && n >= (int64_t)(1 << 22);
        bool sparse = want && cudaPointerGetAttributes(&pa, packed_slice) == cudaSuccess &&
                      pa.type == cudaMemoryTypeHost && pa.devicePointer != nullptr;
        cudaGetLastError();  // an unregistered pointer may leave a sticky-free error code behind
        if (sparse) {
            const int64_t nchunk = (n + myqc::kXferChunk - 1) / myqc::kXferChunk;
            cudaError_t e = cudaSuccess;
            if (nchunk > ce.flags_cap) {
                if (ce.d_flags) { cudaFree(ce.d_flags); ce.d_flags = nullptr; }
                if (ce.h_flags) { cudaFreeHost(ce.h_flags); ce.h_flags = nullptr; }
                ce.flags_cap = 0;
                e = cudaMalloc((void**)&ce.d_flags, (size_t)nchunk);
                if (e == cudaSuccess) e = cudaMallocHost((void**)&ce.h_flags, (size_t)nchunk);
                if (e == cudaSuccess) ce.flags_cap = nchunk;
            }
            unsigned char* d_flags = ce.d_flags;
            unsigned char* h_flags = ce.h_flags;
            cudaEvent_t ev_flags = nullptr;
            int sms = 0;
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_flags, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
            if (e == cudaSuccess) e = (cudaError_t)myqc::launch_chunk_flags(d_out, n, d_flags, sms, nullptr);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_flags, d_flags, (size_t)nchunk, cudaMemcpyDeviceToHost, nullptr);
            if (e == cudaSuccess) e = cudaEventRecord(ev_flags, nullptr);
            if (e == cudaSuccess)
                e = (cudaError_t)myqc::launch_chunk_push(d_out, n, d_flags, static_cast<double*>(pa.devicePointer), sms, nullptr);
            if (e == cudaSuccess) e = cudaEventSynchronize(ev_flags);
            if (e == cudaSuccess) {
                // zeros of the unflagged chunks, written by host threads while the push kernel runs:
                // half the hardware threads, shared between the shards of one box, at most 8 (measured on the
                // 16-vCPU host of the B200 boxes: more writers only fight the PCIe stream for host memory)
                int nthr = std::min(8, (int)std::thread::hardware_concurrency() / 2 / std::max(1, nshards));
                if (const char* et = std::getenv("MYQC_HOST_THREADS")) nthr = std::atoi(et);
                nthr = std::max(1, std::min(nthr, 64));
                std::atomic<int64_t> sent{0};
                auto zero_range = [&](int64_t c0, int64_t c1) {
                    int64_t mine = 0, c = c0;
                    while (c < c1) {
                        if (h_flags[c]) { ++mine; ++c; continue; }
                        int64_t r = c;
                        while (r < c1 && !h_flags[r]) ++r;
                        const int64_t e0 = c * myqc::kXferChunk, e1 = std::min<int64_t>(n, r * myqc::kXferChunk);
                        std::memset(packed_slice + e0, 0, (size_t)(e1 - e0) * sizeof(double));
                        c = r;
                    }
                    sent += mine;
                };
                std::vector<std::thread> th;
                const int64_t per = (nchunk + nthr - 1) / nthr;
                for (int t = 1; t < nthr; ++t)
                    if (t * per < nchunk) th.emplace_back(zero_range, t * per, std::min<int64_t>(nchunk, (t + 1) * per));
                zero_range(0, std::min<int64_t>(nchunk, per));
                for (auto& t : th) t.join();
                xfer_frac = (double)sent.load() / (double)nchunk;
                myqc::g_last_d2h_bytes = sent.load() * myqc::kXferChunk * (int64_t)sizeof(double) + nchunk;
                xfer_kind = "sparse push";
                e = cudaStreamSynchronize(nullptr);
            }
            if (ev_flags) cudaEventDestroy(ev_flags);
            if (e != cudaSuccess) rc = cuda_fail(e, "sparse transfer of the packed slice to the host");
        } else {
            cudaError_t e = cudaMemcpy(packed_slice, d_out, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) rc = cuda_fail(e, "copy packed slice to host");
            myqc::g_last_d2h_bytes = n * (int64_t)sizeof(double);
        }
    } else if (!rc) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = cuda_fail(e, "execute");
        myqc::g_last_d2h_bytes = 0;
    }
    const auto t4 = now();
    if (no_cache) drop_entry(local);
    const auto t5 = now();
    if (trace)
        std::fprintf(stderr, "[myqc trace] shard %d/%d: plan %.1f ms%s, cudaMalloc %.1f ms, execute %.1f ms, D2H %.1f ms (%.2f GB, %.1f GB/s effective, %s, %.1f %% of the chunks sent), free %.1f ms\n",
                     shard, nshards, ms(t0, t1), hit ? " (cached)" : "", ms(t1, t2), ms(t2, t3), ms(t3, t4), 8e-9 * (double)n,
                     8e-9 * (double)n / (ms(t3, t4) * 1e-3 + 1e-12), xfer_kind, 100.0 * xfer_frac, ms(t4, t5));
    return rc;
}

static int run_packed_host(int nnuc, const double* xyz, int nset, int setl, const double* set,
                           const int32_t* setinfo, int ops, const double* bas, const int32_t* basinfo,
                           const double* ftab, double* packed, int ngpu) {
    const int ndev = myqc_device_count();
    if (ndev == 0) return fail(MYQC_ERR_NO_DEVICE, "no CUDA device: the ERI engine has no CPU fallback");
    if (ngpu <= 0 || ngpu > ndev) ngpu = ndev;
    if (!packed) return fail(MYQC_ERR_BAD_ARG, "null output");
    std::vector<int64_t> off(ngpu + 1, 0);
    int rc = myqc_eri_shard_layout(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, ngpu, off.data());
    if (rc) return rc;
    std::vector<int> rcs(ngpu, 0);
    std::vector<std::string> errs(ngpu);
    auto work = [&](int g) {
        rcs[g] = myqc_eri_packed_shard(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, ftab,
                                       packed + off[g], g, g, ngpu, nullptr);
        if (rcs[g]) errs[g] = g_last_error;
    };
    if (ngpu == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int g = 0; g < ngpu; ++g) th.emplace_back(work, g);
        for (auto& t : th) t.join();
    }
    for (int g = 0; g < ngpu; ++g)
        if (rcs[g]) return fail(rcs[g], errs[g]);
    return MYQC_OK;
}

int myqc_eri_packed(int nnuc, const double* xyz, int nset, int setl, const double* set,
                    const int32_t* setinfo, int ops, const double* bas, const int32_t* basinfo,
                    const double* ftab, double* packed, int ngpu) {
    int rc = check_args(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, ftab);
    if (rc) return rc;
    return run_packed_host(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, ftab, packed, ngpu);
}

// Dense XX on `ngpu` devices: every device computes its shard of the packed array; the slices are
// exchanged through the host so that every device holds the whole packed array, and device g expands
// the slab XX(:,:,:,h) of its range of h (the 8n^4-byte stream is produced once, in parallel).
int myqc_eri_dense(int nnuc, const double* xyz, int nset, int setl, const double* set,
                   const int32_t* setinfo, int ops, const double* bas, const int32_t* basinfo,
                   const double* ftab, double* xx, int ngpu) {
    int rc = check_args(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, ftab);
    if (rc) return rc;
    if (!xx) return fail(MYQC_ERR_BAD_ARG, "null output");
    const int ndev = myqc_device_count();
    if (ndev == 0) return fail(MYQC_ERR_NO_DEVICE, "no CUDA device: the ERI engine has no CPU fallback");
    if (ngpu <= 0 || ngpu > ndev) ngpu = ndev;
    const int64_t n = basinfo[1];
    const int64_t npair = n * (n + 1) / 2, nunique = npair * (npair + 1) / 2;
    if (ngpu > n) ngpu = (int)n;
    if (ngpu == 1) {
        myqc_eri_plan* pl = nullptr;
        rc = myqc_eri_plan_create(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, ftab, 0, 0, 1, &pl);
        if (rc) return rc;
        double *d_packed = nullptr, *d_xx = nullptr;
        cudaError_t e = cudaMalloc((void**)&d_packed, (size_t)pl->out_elems * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_xx, (size_t)(n * n * n * n) * sizeof(double));
        if (e != cudaSuccess) {
            cudaFree(d_packed); myqc_eri_plan_destroy(pl);
            return fail(MYQC_ERR_NOMEM, std::string("device allocation for dense XX: ") + cudaGetErrorString(e));
        }
        rc = myqc_eri_plan_execute(pl, d_packed, nullptr);
        if (!rc) rc = myqc_eri_expand_dense(d_packed, (int)n, d_xx, nullptr);
        if (!rc) {
            e = cudaMemcpy(xx, d_xx, (size_t)(n * n * n * n) * sizeof(double), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) rc = cuda_fail(e, "copy dense XX to host");
        }
        cudaFree(d_packed); cudaFree(d_xx);
        myqc_eri_plan_destroy(pl);
        return rc;
    }
    std::vector<double> packed;
    try { packed.resize((size_t)nunique); } catch (...) { return fail(MYQC_ERR_NOMEM, "host allocation of the packed array"); }
    rc = run_packed_host(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, ftab, packed.data(), ngpu);
    if (rc) return rc;
    std::vector<int> rcs(ngpu, 0);
    std::vector<std::string> errs(ngpu);
    auto work = [&](int g) {
        const int h0 = (int)(n * g / ngpu), h1 = (int)(n * (g + 1) / ngpu);
        const size_t slab = (size_t)(n * n * n) * (size_t)(h1 - h0);
        double *d_packed = nullptr, *d_slab = nullptr;
        int sms = 0;
        cudaError_t e = cudaSetDevice(g);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g);
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_packed, (size_t)nunique * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_slab, slab * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpy(d_packed, packed.data(), (size_t)nunique * sizeof(double), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = (cudaError_t)launch_expand_dense(d_packed, (int)n, h0, h1, d_slab, sms, nullptr);
        if (e == cudaSuccess) e = cudaMemcpy(xx + (size_t)(n * n * n) * (size_t)h0, d_slab, slab * sizeof(double), cudaMemcpyDeviceToHost);
        cudaFree(d_packed); cudaFree(d_slab);
        if (e != cudaSuccess) { rcs[g] = MYQC_ERR_CUDA; errs[g] = std::string("dense slab on device ") + std::to_string(g) + ": " + cudaGetErrorString(e); }
    };
    std::vector<std::thread> th;
    for (int g = 0; g < ngpu; ++g) th.emplace_back(work, g);
    for (auto& t : th) t.join();
    for (int g = 0; g < ngpu; ++g)
        if (rcs[g]) return fail(rcs[g], errs[g]);
    return MYQC_OK;
}

}  // extern "C"
