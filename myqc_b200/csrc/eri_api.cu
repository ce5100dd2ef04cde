// C-ABI of the ERI engine (include/myqc_eri.h): plans, one-shot host-buffer calls.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <climits>
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

struct DevList {  // device copy of a PairList
    int type = 0, n = 0, npad = 0;
    double *aos = nullptr, *soa = nullptr, *q = nullptr;
    int32_t *nprim = nullptr, *pidx = nullptr;
    PairList host;  // host copy (without the bulky record arrays) for the prefix computation
};

struct Launch {
    int UT, TT;
    // host copies of the tasks of each part until finalize_launches() merges and uploads them:
    // {task, weight of the task's row} in the part's own order (rows heaviest first, tasks of a row adjacent)
    std::vector<std::pair<int4, double>> part_tasks[kMaxParts];
    ClassArgs args;                // args.tasks/ntasks describe the whole need-sorted task array
    std::vector<int> region_task;  // [nregion+1] task ranges: region r = tasks whose writes end inside fill region r
    double weight = 0.0;           // sum over tasks of (primitives of the row) x (primitives of the lane-side range)
    bool warp = false;             // (SP SP|SP SP) only: the warp-cooperative kernel (one launch) instead of four mu-slices
};

// Below these numbers of contracted quartets per piece the warp-cooperative kernels are used (profiles/r2_notes.md
// section 9).  (SP SP|SP SP): 3 582 quartets 0.239 -> 0.057 ms, 6 481: 0.359 -> 0.104 ms, 118 708: 0.72 -> 1.06 ms.
// (S SP|SP SP) has 10x the quartets of the same molecule and the class kernel fills its lanes: the warp kernel lost on
// all three configurations (0.118 -> 0.149, 0.170 -> 0.252, 2.25 -> 3.99 ms) and is never chosen automatically.
constexpr int64_t kWarpQuartetsPP = 20000, kWarpQuartetsSP = 0;

// kernel launches one class launch issues
static inline int launch_count(const Launch& L) { return L.warp ? 1 : class_nlaunch(L.UT, L.TT); }

constexpr int kMaxCounters = 1024;

}  // namespace myqc

using namespace myqc;

struct Sub {  // one virtual sub-shard: a contiguous piece of the plan's slice with its own launches
    int64_t out_offset = 0, out_elems = 0;  // absolute packed offsets
    std::vector<Launch> launches;
    int counter_base = 0, ncounters = 0;
    // Fill regions: the slice is zero-filled in nregion consecutive pieces; the tasks of every
    // class kernel are sorted by the end of the last packed row they can write to, so the tasks
    // of region r only touch memory that pieces 0..r have already zeroed and can run while the
    // later pieces are still being filled (no extra lists, no extra arithmetic: same tasks).
    std::vector<int64_t> region_end;  // [nregion] offsets relative to the sub-shard's slice
    // screened fill (default): one launch that writes the zeros of exactly those elements no class
    // kernel writes, so it needs no ordering against them
    FillArgs fill;
    int64_t fill_zero_elems = 0;  // zeros the screened fill writes (slice elements minus screened-in ones)
    // compose mode: pair ids g of this piece = position in [mine S.S | mine S.SP | mine SP.SP | later S.S | ...]
    ComposeArgs comp;
    int listbase[6] = {0, 0, 0, 0, 0, 0};
    int ng = 0;
    std::vector<uint32_t> h_rowrel;  // [ng][6], see ComposeArgs::rowrel
    bool pp_warp = false, sp_warp = false;       // (SP SP|SP SP) / (S SP|SP SP) of this piece run on the warp-cooperative kernel
    int64_t pp_quartets = 0, sp_quartets = 0;    // their contracted quartets (reference rule + Schwarz prefix)
};

struct myqc_eri_plan {
    int device = 0, num_sms = 0;
    int norb = 0, nset = 0;
    int64_t npair = 0;
    int64_t out_offset = 0, out_elems = 0;
    std::vector<void*> dev_allocs;
    std::vector<DevList> lists;
    std::vector<Sub> subs;
    double* d_ftab = nullptr;    // [5][121][8]
    double* d_exptab = nullptr;  // [601][2] {exp(-k/10), k/10}
    int* d_counters = nullptr;   // one row counter per class-kernel launch
    int ncounters = 0;
    // internal streams: the zero fill of sub-shard k+1 overlaps the FP64 kernels of sub-shard k,
    // and class kernels of one sub-shard overlap each other's tails
    static constexpr int kNumCompute = 9;  // one per launch of an unsharded step: small launches run side by side
    cudaStream_t s_fill = nullptr, s_comp[kNumCompute] = {};
    cudaEvent_t e_start = nullptr, e_done[kNumCompute + 1] = {};
    std::vector<cudaEvent_t> e_fill;
    // zero fill by the copy engines (fill_engine 1: cudaMemsetAsync, 2: device-to-device copies from a small zero
    // buffer on fill_nstreams streams) -- no SM takes part, so the fill of piece k+1 can run under the FP64 kernels of piece k
    static constexpr int kMaxFillStreams = 4;
    int fill_engine = 1, fill_nstreams = 1;
    void* d_zero = nullptr;
    size_t zero_bytes = 0;
    cudaStream_t s_fillx[kMaxFillStreams] = {};
    cudaEvent_t e_fork = nullptr, e_join[kMaxFillStreams] = {};
    // stats (canonical primitive-quartet counts of the whole shard)
    int64_t nquartets[6] = {0, 0, 0, 0, 0, 0};
    double model_flops = 0.0;
    int nlaunch = 0;
    int64_t h2d_bytes = 0;  // bytes uploaded at plan creation (pair tables, Boys tables, prefixes)
    // inputs kept for the lazily evaluated work statistics (plan_stats)
    std::vector<double> in_xyz, in_set;
    std::vector<int32_t> in_setinfo;
    int in_nnuc = 0, in_setl = 0;
    bool stats_done = false, stats_whole = false;
    bool screened_fill = true;
    // compose mode (MYQC_OUTPUT_MODE=compose): class kernels stage quartet blocks, compose_kernel writes the slice once
    bool compose = false;
    double* d_stage = nullptr;
    int64_t stage_elems = 0;
    double schwarz_tau = 0.0;                 // 0: only the reference's own rule
    unsigned long long* d_pq = nullptr;       // primitive quartets evaluated, one counter per class launch
    int32_t *d_rk = nullptr, *d_cut = nullptr;
    std::vector<int32_t> h_rk, h_cut;
    int nrank = 0;
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

static int upload_list(myqc_eri_plan* pl, const PairList& src, DevList& d) {
    d.type = src.type; d.n = src.n; d.npad = src.npad;
    d.host.type = src.type; d.host.n = src.n; d.host.emax = src.emax; d.host.bucket = src.bucket; d.host.pidx = src.pidx; d.host.nprim = src.nprim;
    d.host.qmax = src.qmax;
    int rc;
    if ((rc = upload(pl, src.aos, &d.aos))) return rc;
    if ((rc = upload(pl, src.soa, &d.soa))) return rc;
    if ((rc = upload(pl, src.nprim, &d.nprim))) return rc;
    if ((rc = upload(pl, src.pidx, &d.pidx))) return rc;
    if ((rc = upload(pl, src.qmax, &d.q))) return rc;
    return MYQC_OK;
}

// number of lane-side pairs v with emax_u*emax_v >= 1e-14 (lane list sorted descending)
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

// ukind / tkind: 0 = the piece's own ("mine") list, 1 = its "later" list
static int add_launch(myqc_eri_plan* pl, Sub& sub, int ui, int ti, bool tri, int ukind = 0, int tkind = 0) {
    const DevList& U = pl->lists[ui];
    const DevList& T = pl->lists[ti];
    if (U.n == 0 || T.n == 0) return MYQC_OK;
    const int nregion = (int)sub.region_end.size();
    // Parts of one class are merged into one launch (one tail per class and shard instead of up to three) unless the
    // fill-region experiments (MYQC_FILL_REGIONS > 1) need a task range per region and launch.
    Launch* Lp = nullptr;
    if (nregion == 1)
        for (Launch& x : sub.launches)
            if (x.UT == U.type && x.TT == T.type && x.args.nparts < kMaxParts) Lp = &x;
    Launch Lnew;
    if (!Lp) {
        Lnew.UT = U.type; Lnew.TT = T.type;
        Lnew.warp = T.type == 2 && ((U.type == 2 && sub.pp_warp) || (U.type == 1 && sub.sp_warp));
        std::memset(&Lnew.args, 0, sizeof(Lnew.args));
        Lp = &Lnew;
    }
    Launch& L = *Lp;
    ClassArgs& a = L.args;
    const int ip = a.nparts;  // index of this part
    PartArgs pa;
    std::memset(&pa, 0, sizeof(pa));
    pa.u_aos = U.aos; pa.u_nprim = U.nprim; pa.u_pidx = U.pidx;
    const std::vector<int32_t> ntv = row_prefix(U.host, T.host, pl->schwarz_tau);
    // segments of the lane-side list: at most kTaskPairs pairs, cut at group boundaries once a
    // segment holds >= 64 pairs, so that a task holds pairs of (mostly) one kind
    // A task is one warp's unit of work and lasts as long as its longest lane.  When the whole launch has fewer
    // tasks than a few per resident warp (small and medium molecules) its duration is that of ONE task, so the
    // tasks are cut finer, down to a single chunk of 32 lane-side pairs.
    int maxpairs = class_task_pairs(U.type, T.type);
    {
        static const int occ_class[3][3] = {{7, 6, 4}, {6, 3, 2}, {4, 2, 2}};  // resident CTAs per SM (launch bounds of the class kernels)
        const double warps = 4.0 * pl->num_sms * occ_class[U.type][T.type];
        double total_pairs = 0.0;
        for (int u = 0; u < U.n; ++u) total_pairs += std::max(0, ntv[u] - (tri ? u : 0));
        while (maxpairs > 32 && total_pairs / maxpairs < 4.0 * warps) maxpairs /= 2;
        if (T.type == 2 && ((U.type == 2 && sub.pp_warp) || (U.type == 1 && sub.sp_warp))) {
            // warp-cooperative kernel: a warp works through the quartets of a task one after the other (each is one to
            // three batches of 32 primitive quartets), so tasks are a few quartets long
            maxpairs = 8;
            while (maxpairs > 1 && total_pairs / maxpairs < 4.0 * warps) maxpairs /= 2;
        }
    }
    const int mingroup = std::min(64, maxpairs);
    std::vector<int> seg;  // segment start offsets, terminated by T.n
    seg.push_back(0);
    for (int k = 1; k < T.n; ++k) {
        const int len = k - seg.back();
        if (len >= maxpairs || (len >= mingroup && T.host.bucket[k] != T.host.bucket[k - 1])) seg.push_back(k);
    }
    seg.push_back(T.n);
    // last packed element a row u can write to: end of packed row max_f P(u,f)
    std::vector<int64_t> need_u(U.n, 0);
    {
        const int nfu = pt_nf(U.type);
        for (int u = 0; u < U.n; ++u) {
            int64_t pmax = -1;
            for (int f = 0; f < nfu; ++f) pmax = std::max<int64_t>(pmax, U.host.pidx[(size_t)u * nfu + f]);
            need_u[u] = pmax < 0 ? 0 : (pmax + 1) * pl->npair - (pmax + 1) * pmax / 2 - sub.out_offset;  // row end, slice-relative
        }
    }
    struct TaskN { int4 t; int64_t need; int region; double weight; };
    std::vector<double> tprim_prefix(T.n + 1, 0.0);  // prefix sums of lane-side primitive counts
    for (int k = 0; k < T.n; ++k) tprim_prefix[k + 1] = tprim_prefix[k] + T.host.nprim[k];
    // MYQC_TASK_ORDER=0: plain heaviest-task-first order (23.5 ms on (H2O)_64 against 22.5 ms, profiles/r1_notes.md)
    static const int task_order = std::getenv("MYQC_TASK_ORDER") ? std::atoi(std::getenv("MYQC_TASK_ORDER")) : 1;
    auto emit_row = [&](int u, std::vector<TaskN>& dst) {
        const int lo = tri ? u : 0, hi = ntv[u];
        if (lo >= hi) return;
        size_t si = std::upper_bound(seg.begin(), seg.end(), lo) - seg.begin() - 1;
        for (; si + 1 < seg.size() && seg[si] < hi; ++si) {
            const int b = std::max(seg[si], lo), e = std::min(seg[si + 1], hi);
            if (b < e)
                dst.push_back({make_int4(u, b, e, ip), need_u[u], 0,
                               (double)U.host.nprim[u] * (tprim_prefix[e] - tprim_prefix[b])});
        }
    };
    std::vector<TaskN> tn;
    std::vector<int4> tasks;
    std::vector<double> task_roww;  // weight of the row each task belongs to (merge key)
    double part_weight = 0.0;
    std::vector<int> region_task(nregion + 1, 0);
    if (task_order == 1 && nregion == 1) {
        // Rows heaviest first, the tasks of one row adjacent (in lane-side order): everything a launch
        // stores into the packed rows of one uniform-side pair is stored within a short time, so
        // partial-sector stores of neighbouring lanes meet in L2.  Only the rows are sorted.
        std::vector<double> roww(U.n, 0.0);
        std::vector<int> order;
        order.reserve(U.n);
        for (int u = 0; u < U.n; ++u) {
            const int lo = tri ? u : 0, hi = ntv[u];
            if (lo >= hi) continue;
            roww[u] = (double)U.host.nprim[u] * (tprim_prefix[hi] - tprim_prefix[lo]);
            order.push_back(u);
        }
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return roww[x] > roww[y]; });
        for (int u : order) {
            tn.clear();
            emit_row(u, tn);
            for (const TaskN& t : tn) { tasks.push_back(t.t); task_roww.push_back(roww[u]); part_weight += t.weight; }
        }
        region_task[1] = (int)tasks.size();
    } else {
        for (int u = 0; u < U.n; ++u) emit_row(u, tn);
        // region of a task = first fill region that covers everything the task can write; inside a
        // region the heaviest tasks (most primitive quartets) go first so that launches end on light ones
        for (TaskN& t : tn) {
            int r = 0;
            while (r + 1 < nregion && t.need > sub.region_end[r]) ++r;
            t.region = r;
        }
        if (task_order == 1) {
            std::vector<double> roww(U.n, 0.0);
            for (const TaskN& t : tn) roww[t.t.x] += t.weight;
            std::stable_sort(tn.begin(), tn.end(), [&](const TaskN& x, const TaskN& y) {
                if (x.region != y.region) return x.region < y.region;
                if (x.t.x != y.t.x) {
                    if (roww[x.t.x] != roww[y.t.x]) return roww[x.t.x] > roww[y.t.x];
                    return x.t.x < y.t.x;
                }
                return x.t.y < y.t.y;
            });
        } else {
            std::stable_sort(tn.begin(), tn.end(), [](const TaskN& x, const TaskN& y) {
                if (x.region != y.region) return x.region < y.region;
                return x.weight > y.weight;
            });
        }
        tasks.resize(tn.size());
        task_roww.resize(tn.size());
        for (size_t k = 0; k < tn.size(); ++k) { tasks[k] = tn[k].t; task_roww[k] = tn[k].weight; part_weight += tn[k].weight; }
        region_task.assign(nregion + 1, (int)tn.size());
        region_task[0] = 0;
        for (size_t k = 0, r = 0; r < (size_t)nregion; ++r) {
            while (k < tn.size() && tn[k].region <= (int)r) ++k;
            region_task[r + 1] = (int)k;
        }
    }
    if (tasks.empty()) return MYQC_OK;
    int rc = MYQC_OK;
    pa.t_soa = T.soa; pa.t_aos = T.aos; pa.t_nprim = T.nprim; pa.t_pidx = T.pidx;
    pa.t_npad = T.npad; pa.tri = tri ? 1 : 0;
    pa.u_q = U.q; pa.t_q = T.q;
    if (ip == 0) {
        a.tau = (U.q && T.q) ? pl->schwarz_tau : 0.0;
        a.ftab_q = pl->d_ftab + (size_t)(U.type + T.type) * 121 * 8;
        a.exptab = reinterpret_cast<const double2*>(pl->d_exptab);
        a.out = nullptr;
        a.out_offset = sub.out_offset;
        a.npair = pl->npair;
        L.region_task = region_task;
    }
    L.part_tasks[ip].resize(tasks.size());
    for (size_t k = 0; k < tasks.size(); ++k) L.part_tasks[ip][k] = {tasks[k], task_roww[k]};
    L.weight += part_weight;
    if (pl->compose) {
        // quartet blocks of this launch in the staging array: row u holds the blocks of v in [lo_u, ntv[u]) one after
        // the other; stage_row[u] is the position its v = 0 block would have, so block(u,v) = stage_row[u] + v*blk
        const int64_t blk = (int64_t)pt_nf(U.type) * pt_nf(T.type);
        std::vector<int64_t> srow(U.n, INT64_MIN);
        int64_t pos = (pl->stage_elems + 15) & ~(int64_t)15;
        const int64_t lbase = pos;
        const int lidu = 2 * U.type + ukind, lidt = 2 * T.type + tkind;
        sub.comp.launch_base[lidu * 6 + lidt] = lbase;
        for (int u = 0; u < U.n; ++u) {
            const int lo = tri ? u : 0, hi = ntv[u];
            if (lo >= hi) continue;
            srow[u] = pos - (int64_t)lo * blk;
            pos += (int64_t)(hi - lo) * blk;
            sub.h_rowrel[(size_t)(sub.listbase[lidu] + u) * 6 + lidt] = (uint32_t)(srow[u] - lbase);  // mod 2^32, see ComposeArgs
        }
        if (pos - lbase >= ((int64_t)1 << 32)) return fail(MYQC_ERR_UNSUPPORTED, "one class launch stages more than 2^32 integrals");
        pl->stage_elems = pos;
        int64_t* d_srow = nullptr;
        if ((rc = upload(pl, srow, &d_srow))) return rc;
        pa.stage_row = d_srow;
    }
    a.part[ip] = pa;
    a.nparts = ip + 1;
    if (Lp == &Lnew) sub.launches.push_back(Lnew);
    return MYQC_OK;
}

// After all parts of a piece are known: merge the task lists of every launch (rows heaviest first across the parts,
// the tasks of a row still adjacent), upload them, and give the launch its task and work counters.
static int finalize_launches(myqc_eri_plan* pl, Sub& sub) {
    const int nregion = (int)sub.region_end.size();
    for (Launch& L : sub.launches) {
        ClassArgs& a = L.args;
        std::vector<int4> tasks;
        if (a.nparts == 1) {
            tasks.reserve(L.part_tasks[0].size());
            for (const auto& t : L.part_tasks[0]) tasks.push_back(t.first);
        } else {
            size_t pos[kMaxParts] = {0, 0, 0}, total = 0;
            for (int k = 0; k < a.nparts; ++k) total += L.part_tasks[k].size();
            tasks.reserve(total);
            while (tasks.size() < total) {
                int best = -1;
                for (int k = 0; k < a.nparts; ++k)
                    if (pos[k] < L.part_tasks[k].size() && (best < 0 || L.part_tasks[k][pos[k]].second > L.part_tasks[best][pos[best]].second)) best = k;
                // the whole row (its tasks are adjacent and carry the same row weight)
                const int u = L.part_tasks[best][pos[best]].first.x;
                while (pos[best] < L.part_tasks[best].size() && L.part_tasks[best][pos[best]].first.x == u) tasks.push_back(L.part_tasks[best][pos[best]++].first);
            }
            L.region_task.assign(2, 0);
            L.region_task[1] = (int)tasks.size();
        }
        for (auto& v : L.part_tasks) { v.clear(); v.shrink_to_fit(); }
        int4* d_tasks = nullptr;
        int rc = upload(pl, tasks, &d_tasks);
        if (rc) return rc;
        a.tasks = d_tasks;
        a.ntasks = (int)tasks.size();
        const int ncnt = launch_count(L) * nregion;  // one task counter per launch and region
        if (pl->ncounters + ncnt > kMaxCounters) return fail(MYQC_ERR_UNSUPPORTED, "too many launches in one plan");
        a.pq_counter = pl->d_pq + pl->ncounters;  // one per launch (slices of a launch share it)
        a.row_counter = pl->d_counters + pl->ncounters;
        pl->ncounters += ncnt;
        sub.ncounters += ncnt;
        pl->nlaunch += launch_count(L) * nregion;
    }
    return MYQC_OK;
}

// Rank / cut arrays of the screened fill (see fill_screened_kernel).  all[] are the complete pair
// lists of the molecule: every function pair P belongs to exactly one shell pair.
static int build_screen_ranks(myqc_eri_plan* pl, const PairList all[3]) {
    std::vector<double> E;
    for (int t = 0; t < 3; ++t) E.insert(E.end(), all[t].emax.begin(), all[t].emax.end());
    std::sort(E.begin(), E.end(), std::greater<double>());
    pl->nrank = (int)E.size();
    pl->h_rk.assign((size_t)pl->npair, INT32_MAX);
    pl->h_cut.assign((size_t)pl->npair, 0);
    for (int t = 0; t < 3; ++t) {
        const int nf = pt_nf(t);
        for (int q = 0; q < all[t].n; ++q) {
            const double e = all[t].emax[q];
            // first occurrence of e in the descending list
            const int32_t rank = (int32_t)(std::lower_bound(E.begin(), E.end(), e, std::greater<double>()) - E.begin());
            size_t lo = 0, hi = E.size();
            while (lo < hi) {  // first index whose product with e fails the reference's test
                const size_t mid = (lo + hi) / 2;
                if (e * E[mid] < 1.0e-14) hi = mid; else lo = mid + 1;
            }
            for (int f = 0; f < nf; ++f) {
                const int32_t P = all[t].pidx[(size_t)q * nf + f];
                if (P < 0) continue;
                pl->h_rk[P] = rank;
                pl->h_cut[P] = (int32_t)lo;
            }
        }
    }
    int rc;
    if ((rc = upload(pl, pl->h_rk, &pl->d_rk))) return rc;
    if ((rc = upload(pl, pl->h_cut, &pl->d_cut))) return rc;
    return MYQC_OK;
}

static int build_sub_fill(myqc_eri_plan* pl, Sub& sub, int64_t row_lo, int64_t row_hi) {
    FillArgs& f = sub.fill;
    std::memset(&f, 0, sizeof(f));
    f.out_offset = sub.out_offset;
    f.npair = pl->npair;
    f.row_lo = row_lo; f.row_hi = row_hi;
    f.rk = pl->d_rk; f.cut = pl->d_cut;
    f.all = std::getenv("MYQC_FILL_ALL") ? 1 : 0;
    f.sleep_ns = std::getenv("MYQC_FILL_SLEEP_NS") ? std::atoi(std::getenv("MYQC_FILL_SLEEP_NS")) : 0;
    sub.fill_zero_elems = 0;
    if (row_hi <= row_lo) return MYQC_OK;
    const int64_t ncb_all = (pl->npair + kFillCols - 1) / kFillCols;
    f.cb0 = (int)(row_lo / kFillCols);
    f.ncb = (int)(ncb_all - f.cb0);
    std::vector<int32_t> ucb(f.ncb + 1, 0);
    for (int k = 0; k < f.ncb; ++k) {
        const int64_t cend = std::min<int64_t>(((int64_t)f.cb0 + k + 1) * kFillCols, pl->npair);  // one past the last column
        const int64_t rows = std::min(row_hi, cend) - row_lo;                                    // rows r <= last column
        const int64_t nrb = rows > 0 ? (rows + kFillRows - 1) / kFillRows : 0;
        if ((int64_t)ucb[k] + nrb > INT32_MAX) return fail(MYQC_ERR_UNSUPPORTED, "too many fill units");
        ucb[k + 1] = ucb[k] + (int32_t)nrb;
    }
    f.nunits = ucb[f.ncb];
    int32_t* d_ucb = nullptr;
    int rc = upload(pl, ucb, &d_ucb);
    if (rc) return rc;
    f.ucb = d_ucb;
    if (pl->ncounters + 1 > kMaxCounters) return fail(MYQC_ERR_UNSUPPORTED, "too many launches in one plan");
    f.counter = pl->d_counters + pl->ncounters;
    pl->ncounters += 1;
    // zeros written = elements of the rows minus the screened-in ones: count pairs (P <= P') with
    // rk[P'] < cut[P] with a Fenwick tree over the ranks
    std::vector<int32_t> bit((size_t)pl->nrank + 1, 0);
    int64_t touched = 0;
    for (int64_t P = pl->npair - 1; P >= row_lo; --P) {
        const int32_t r = pl->h_rk[P];
        if (r < pl->nrank)
            for (int i = r + 1; i <= pl->nrank; i += i & (-i)) ++bit[i];
        if (P < row_hi) {
            int64_t c = 0;
            for (int i = pl->h_cut[P]; i > 0; i -= i & (-i)) c += bit[i];
            touched += c;
        }
    }
    sub.fill_zero_elems = sub.out_elems - touched;
    return MYQC_OK;
}

// canonical primitive-quartet statistics: unordered primitive pairs {a<=b},
// unordered pairs of pairs, kept iff EIJ*EGH >= 1e-14; restricted to the shard by `owner`.
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

// packed index of the first element of the first row whose leading orbital is `fn`
static int64_t packed_row_offset(int64_t fn, int64_t norb) {
    const int64_t np = norb * (norb + 1) / 2;
    if (fn >= norb) return np * (np + 1) / 2;
    const int64_t P = fn * norb - fn * (fn - 1) / 2;  // P(fn,fn)
    return P * np - P * (P - 1) / 2;
}

// Ownership rule (multi-GPU): a shell quartet belongs to the shard that owns the
// smallest first-orbital id among its four shells; all its canonical integrals then lie in packed
// rows whose leading orbital is inside that shell.  Shards are contiguous blocks of such rows.
// Pure host arithmetic: every rank computes the same answer independently.
struct CutTable {
    std::vector<int> cuts;   // candidate cut points: first orbital of each shell, ascending
    std::vector<double> w;   // estimated seconds of the block [cuts[c], cuts[c+1]): class kernels + zero fill
    std::vector<double> wclass, wfill;  // the two parts of w
};

// Fenwick tree over doubles (prefix sums with point updates)
struct Fenwick {
    std::vector<double> t;
    explicit Fenwick(int n) : t((size_t)n + 1, 0.0) {}
    void add(int i, double v) { for (++i; i < (int)t.size(); i += i & -i) t[i] += v; }
    double prefix(int n) const { double s = 0.0; for (int i = std::min(n, (int)t.size() - 1); i > 0; i -= i & -i) s += t[i]; return s; }  // sum of [0, n)
    double range(int lo, int hi) const { return hi > lo ? prefix(hi) - prefix(lo) : 0.0; }
};

// nq_ref: canonical primitive-quartet counts of the whole molecule per class (canonical_stats), or nullptr
static int build_cut_table(const std::vector<Shell>& shells, const PairList all[3], const int64_t* nq_ref, int nshards, CutTable& ct, std::string& err) {
    for (const Shell& sh : shells) {  // row blocks are closed only for contiguous orbital ranges
        int cnt = 0, mx = -1;
        for (int k = 0; k < 4; ++k) if (sh.fn[k] >= 0) { ++cnt; mx = std::max(mx, sh.fn[k]); }
        if (mx - sh.first_fn + 1 != cnt) { err = "sharding needs contiguous orbital ids per shell"; return MYQC_ERR_UNSUPPORTED; }
    }
    std::vector<int>& cuts = ct.cuts;
    cuts.clear();
    for (const Shell& sh : shells) cuts.push_back(sh.first_fn);
    std::sort(cuts.begin(), cuts.end());
    cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
    const int nc = (int)cuts.size();
    auto cut_of = [&](int fn) { return (int)(std::upper_bound(cuts.begin(), cuts.end(), fn) - cuts.begin()) - 1; };
    // Class-kernel seconds of block c = (primitive quartets of the quartets it owns) x (seconds per primitive quartet
    // of the class).  A quartet (u|v) belongs to the block min(cut(u), cut(v)); u's partners are the prefix [lo_u, L_u)
    // of the lane-side list.  The count is taken with the separable proxy nprim(u) x nprim(v) per quartet, summed
    // EXACTLY over the (row, partner) pairs with two sweeps over the blocks and a Fenwick tree each, and scaled per
    // class to the canonical primitive-quartet count of the whole molecule.  (Until session 3 of round 2 the partners
    // of a row were split over the blocks in proportion to ALL pairs of the list; the prefix holds the pairs with the
    // largest prefactors -- near pairs, which are spread evenly over the blocks, while all pairs are front-loaded --
    // so late blocks were under-weighted: the last of 8 shards of (H2O)_64 got 1.57x the primitive quartets of the first
    // where the model meant 1.38x, and was 8 % slower than the rest at every N.)
    ct.wclass.assign(nc, 0.0);
    ct.wfill.assign(nc, 0.0);
    for (int ta = 0; ta < 3; ++ta)
        for (int tb = ta; tb < 3; ++tb) {
            const PairList& A = all[ta];
            const PairList& B = all[tb];
            if (A.n == 0 || B.n == 0) continue;
            const bool tri = (ta == tb);
            std::vector<int32_t> L = row_prefix(A, B);
            for (int u = 1; u < A.n; ++u) L[u] = std::min(L[u], L[u - 1]);  // rows are in prefactor order: monotone up to rounding
            std::vector<int> cA(A.n), cB(B.n);
            std::vector<std::vector<int>> bktA(nc), bktB(nc);
            for (int u = 0; u < A.n; ++u) { cA[u] = cut_of(A.owner_fn[u]); bktA[cA[u]].push_back(u); }
            for (int v = 0; v < B.n; ++v) { cB[v] = cut_of(B.owner_fn[v]); bktB[cB[v]].push_back(v); }
            std::vector<double> acc(nc, 0.0);
            {   // quartets owned through the row: partners v in [lo_u, L_u) with cut(v) >= cut(u)
                Fenwick fb(B.n);
                for (int c = nc - 1; c >= 0; --c) {
                    for (int v : bktB[c]) fb.add(v, (double)B.nprim[v]);
                    for (int u : bktA[c]) acc[c] += (double)A.nprim[u] * fb.range(tri ? u : 0, L[u]);
                }
            }
            {   // quartets owned through the partner: rows u with lo_u <= v < L_u and cut(u) > cut(v)
                Fenwick fa(A.n);
                for (int c = nc - 1; c >= 0; --c) {
                    for (int v : bktB[c]) {
                        // rows with L_u > v are a prefix of the (monotone) list
                        int ustar = (int)(std::partition_point(L.begin(), L.end(), [&](int32_t l) { return l > v; }) - L.begin());
                        if (tri) ustar = std::min(ustar, v + 1);
                        acc[c] += (double)B.nprim[v] * fa.prefix(ustar);
                    }
                    for (int u : bktA[c]) fa.add(u, (double)A.nprim[u]);
                }
            }
            // seconds per primitive quartet: model flops / (class efficiency x DFMA peak).  The efficiencies are machine
            // constants of these kernels on a B200 (fraction of the DFMA peak in model flops when a launch fills the
            // machine: profiles/r2f_bench.json); only their ratios to each other and to the fill rate below enter the cuts.
            static const double kEff[6] = {0.46, 0.47, 0.32, 0.36, 0.30, 0.27};
            const int cid = class_id(ta, tb);
            // Launches of a 1/N share run below the efficiency of the full-size launches the constants come from (tails,
            // fewer tasks per warp): replays of the 2- / 4- / 8-way splits of (H2O)_64 on one GPU take 1.08 / 1.15 / 1.21x
            // the model's class time in every shard alike (tools/exp_shard_times.py), i.e. 1 + 0.07 log2 N.  The factor
            // only shifts weight between the class kernels and the zero fill of a shard.
            const double small_launch = 1.0 + 0.07 * std::log2((double)std::max(1, nshards));
            const double sec_per_pq = small_launch * kW[cid] / (kEff[cid] * 34.2e12);
            double proxy = 0.0;
            for (int c = 0; c < nc; ++c) proxy += acc[c];
            if (!(proxy > 0.0)) continue;
            // the proxy counts every primitive pair of both sides; the reference's rule on primitives removes a
            // class-dependent share of them, so the class total is taken from the canonical count when it is known
            const double scale = (nq_ref && nq_ref[cid] > 0) ? (double)nq_ref[cid] / proxy : 1.0;
            for (int c = 0; c < nc; ++c) ct.wclass[c] += acc[c] * scale * sec_per_pq;
        }
    // plus the zero fill of the rows the block owns (HBM bound; cudaMemsetAsync writes ~7.3 TB/s on a B200)
    {
        int norb = 0;
        for (const Shell& sh : shells)
            for (int k = 0; k < 4; ++k) norb = std::max(norb, sh.fn[k] + 1);
        for (int c = 0; c < nc; ++c) {
            const int64_t b = packed_row_offset(c == 0 ? 0 : cuts[c], norb);
            const int64_t e = packed_row_offset(c + 1 < nc ? cuts[c + 1] : norb, norb);
            ct.wfill[c] = 8.0 * (double)(e - b) / 7.3e12;
        }
    }
    ct.w.assign(nc, 0.0);
    for (int c = 0; c < nc; ++c) ct.w[c] = ct.wclass[c] + ct.wfill[c];
    return MYQC_OK;
}

// cut the candidate range [c_lo, c_hi) into m contiguous parts of ~equal weight; returns m+1 indices
static std::vector<int> split_range(const CutTable& ct, int c_lo, int c_hi, int m) {
    std::vector<int> bound(m + 1, c_hi);
    bound[0] = c_lo;
    double tot = 0;
    for (int c = c_lo; c < c_hi; ++c) tot += ct.w[c];
    double acc = 0;
    int sidx = 1;
    for (int c = c_lo; c < c_hi && sidx < m; ++c) {
        acc += ct.w[c];
        while (sidx < m && acc >= tot * sidx / m) bound[sidx++] = c + 1;
    }
    for (int k = 1; k <= m; ++k) bound[k] = std::max(bound[k], bound[k - 1]);
    bound[m] = c_hi;
    return bound;
}

static int cut_to_fn(const CutTable& ct, int c, int norb) { return c < (int)ct.cuts.size() ? (c == 0 ? 0 : ct.cuts[c]) : norb; }

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
    std::vector<Shell> shells;
    if ((rc = build_shells(nnuc, nset, setl, setinfo, ops, basinfo, shells, err))) return fail(rc, err);
    PairList all[3];
    if ((rc = build_pairs(nnuc, xyz, set, setinfo, setl, ops, bas, basinfo, shells, all, err))) return fail(rc, err);
    stage("shells + pair records");
    {
        // Default: zero the whole slice first (streaming 128-bit stores, 0.95 of HBM peak), class kernels
        // after it.  MYQC_FILL_MODE=screened selects the order-independent screened fill that runs next
        // to the class kernels; measured on (H2O)_64 it does not pay (profiles/r1_notes.md): the class
        // kernels' scattered stores and the fill compete for DRAM, the co-run takes the sum of both.
        const char* fm = std::getenv("MYQC_FILL_MODE");
        pl->screened_fill = (fm && std::strcmp(fm, "screened") == 0);
        // Default ("scatter"): zero fill of the slice, then the class kernels store their integrals into it.
        // MYQC_OUTPUT_MODE=compose: the class kernels stage dense quartet blocks and compose_kernel writes every element
        // of the slice exactly once (DRAM traffic 1.2x the algorithmic bytes instead of 1.5x, class kernels 8.2 instead
        // of 10.9 ms on (H2O)_64) -- but the element-wise gather of the compose pass costs 12 ms against 6.5 ms for
        // the plain fill, so the step is slower (20.3 against 17.1 ms; profiles/r2_notes.md section 3).
        const char* om = std::getenv("MYQC_OUTPUT_MODE");
        pl->compose = (om && std::strcmp(om, "compose") == 0) && !pl->screened_fill;

    }

    // Boys tables for the five start orders Q = 0,3,6,9,12: row t = {Ft(t,Q+k)/k!, k<7 ; t/10}
    {
        std::vector<double> h, ex;
        build_boys_tables(ftab, h, ex);
        if ((rc = upload(pl.get(), h, &pl->d_ftab))) return rc;
        if ((rc = upload(pl.get(), ex, &pl->d_exptab))) return rc;
        std::vector<int> zeros(kMaxCounters, 0);
        if ((rc = upload(pl.get(), zeros, &pl->d_counters))) return rc;
        std::vector<unsigned long long> zq(kMaxCounters, 0ull);
        if ((rc = upload(pl.get(), zq, &pl->d_pq))) return rc;
    }
    if ((pl->screened_fill || pl->compose) && (rc = build_screen_ranks(pl.get(), all))) return rc;
    stage("tables");
    // ---- Schwarz factors (north star (1); SURVEY.md 7 "Parity vs. screening").  A contracted quartet (u|v) is left out
    // when Q_u*Q_v < tau, Q = an upper bound of sqrt((ij|ij)) over the pair's function pairs: every integral of the
    // quartet is then below tau in magnitude, and an integral belongs to exactly one contracted quartet, so the
    // omission per integral is < tau (default 1e-12, two orders below the 1e-10 parity bar; MYQC_SCHWARZ_TAU=0 keeps
    // the reference's rule alone).  The diagonals come from a device pass WITHOUT the reference's screen; the primitive
    // pairs the builder dropped (E < 1e-14) are covered by the additive term below.
    {
        const char* et = std::getenv("MYQC_SCHWARZ_TAU");
        pl->schwarz_tau = et ? std::atof(et) : 1.0e-12;
        if (!(pl->schwarz_tau > 0.0)) pl->schwarz_tau = 0.0;
    }
    if (pl->schwarz_tau > 0.0) {
        double* d_diag = nullptr;
        CU(cudaMalloc((void**)&d_diag, (size_t)pl->npair * sizeof(double)));
        CU(cudaMemset(d_diag, 0, (size_t)pl->npair * sizeof(double)));
        std::vector<void*> tmp;
        int e = 0;
        for (int t = 0; t < 3 && !e; ++t) {
            if (all[t].n == 0) continue;
            double* d_aos = nullptr; int32_t *d_np = nullptr, *d_px = nullptr;
            cudaError_t ce = cudaMalloc((void**)&d_aos, all[t].aos.size() * sizeof(double));
            if (ce == cudaSuccess) { tmp.push_back(d_aos); ce = cudaMalloc((void**)&d_np, all[t].nprim.size() * sizeof(int32_t)); }
            if (ce == cudaSuccess) { tmp.push_back(d_np); ce = cudaMalloc((void**)&d_px, all[t].pidx.size() * sizeof(int32_t)); }
            if (ce == cudaSuccess) { tmp.push_back(d_px); ce = cudaMemcpy(d_aos, all[t].aos.data(), all[t].aos.size() * sizeof(double), cudaMemcpyHostToDevice); }
            if (ce == cudaSuccess) ce = cudaMemcpy(d_np, all[t].nprim.data(), all[t].nprim.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
            if (ce == cudaSuccess) ce = cudaMemcpy(d_px, all[t].pidx.data(), all[t].pidx.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
            if (ce != cudaSuccess) { e = (int)ce; break; }
            pl->h2d_bytes += (int64_t)(all[t].aos.size() * sizeof(double) + (all[t].nprim.size() + all[t].pidx.size()) * sizeof(int32_t));
            e = launch_diag(t, d_aos, d_np, d_px, all[t].n, pl->d_ftab + (size_t)(2 * t) * 121 * 8,
                            reinterpret_cast<const double2*>(pl->d_exptab), d_diag, nullptr);
        }
        std::vector<double> diag((size_t)pl->npair, 0.0);
        if (!e) e = (int)cudaMemcpy(diag.data(), d_diag, diag.size() * sizeof(double), cudaMemcpyDeviceToHost);
        for (void* p : tmp) cudaFree(p);
        cudaFree(d_diag);
        if (e) return cuda_fail((cudaError_t)e, "Schwarz diagonal pass");
        for (int t = 0; t < 3; ++t) {
            const int nf = pt_nf(t);
            all[t].qmax.assign(all[t].n, 0.0);
            for (int k = 0; k < all[t].n; ++k) {
                double q2 = 0.0;
                for (int f = 0; f < nf; ++f) {
                    const int32_t P = all[t].pidx[(size_t)k * nf + f];
                    if (P >= 0) q2 = std::max(q2, diag[(size_t)P]);
                }
                // + what the dropped primitive pairs (E < 1e-14, at most 80 quartets of O(1) x E x 1e-14 each) could add
                all[t].qmax[k] = std::sqrt(q2 + 1.0e-11 * all[t].emax[k]) * (1.0 + 1.0e-9);
            }
        }
        stage("Schwarz diagonals");
    }

    // ---- sharding: contiguous blocks of packed rows, cut where a shell's functions start -------
    // External shards (one per GPU) are cut first; this plan's shard is then cut again into
    // virtual sub-shards so that the zero fill of piece k+1 overlaps the FP64 kernels of piece k.
    CutTable ct;
    ct.cuts.push_back(0);
    ct.w.push_back(1.0);
    int nvs = 1;
    {
        const char* env = std::getenv("MYQC_VSHARDS");
        const int64_t total = pl->npair * (pl->npair + 1) / 2;
        // Measured on (H2O)_64 (profiles/r1_notes.md): cutting one GPU's shard into pieces overlaps the
        // zero fill with the FP64 kernels but costs more in extra launches, shorter task lists and
        // duplicated "later" lists than it wins (23.5 / 25.0 / 26.6 / 32.4 ms for 1 / 2 / 4 / 8 pieces),
        // so the default is one piece; MYQC_VSHARDS overrides it for experiments.
        (void)total;
        nvs = env ? std::atoi(env) : 1;
        if (nvs < 1) nvs = 1;
        if (nvs > 16) nvs = 16;
    }
    std::vector<int> sub_fn(2, 0);
    sub_fn[1] = pl->norb;
    if (nshards > 1 || nvs > 1) {
        int64_t nq_ref[6];
        double fl_ref = 0.0;
        canonical_stats(nnuc, xyz, nset, setl, set, setinfo, nq_ref, &fl_ref);
        if ((rc = build_cut_table(shells, all, nq_ref, nshards, ct, err))) return fail(rc, err);
        const int nc = (int)ct.cuts.size();
        const std::vector<int> ext = split_range(ct, 0, nc, nshards);
        const std::vector<int> sub = split_range(ct, ext[shard], ext[shard + 1], nvs);
        sub_fn.clear();
        for (int c : sub) sub_fn.push_back(cut_to_fn(ct, c, pl->norb));
        // drop empty pieces
        std::vector<int> keep;
        for (size_t k = 0; k < sub_fn.size(); ++k)
            if (k == 0 || sub_fn[k] > keep.back()) keep.push_back(sub_fn[k]);
        if (keep.size() < 2) keep.push_back(keep.back());
        sub_fn = keep;
    }
    pl->out_offset = packed_row_offset(sub_fn.front(), pl->norb);
    pl->out_elems = packed_row_offset(sub_fn.back(), pl->norb) - pl->out_offset;

    stage("shard cuts");
    const int nsub = (int)sub_fn.size() - 1;
    pl->subs.resize(nsub);
    pl->lists.reserve(6 * nsub);
    for (int k = 0; k < nsub; ++k) {
        Sub& sub = pl->subs[k];
        const int fn_lo = sub_fn[k], fn_hi = sub_fn[k + 1];
        sub.out_offset = packed_row_offset(fn_lo, pl->norb);
        sub.out_elems = packed_row_offset(fn_hi, pl->norb) - sub.out_offset;
        sub.counter_base = pl->ncounters;
        {
            const char* envr = std::getenv("MYQC_FILL_REGIONS");
            // Default 1: on (H2O)_64 the fill of region r+1 and the kernels of region r do not co-run
            // well (persistent grids starve each other): 23.2 / 24.2 / 24.8 ms for 1 / 2 / 4 regions
            // (profiles/r1_notes.md).  MYQC_FILL_REGIONS overrides it for experiments.
            int nreg = envr ? std::atoi(envr) : 1;
            if (nreg < 1 || pl->screened_fill || pl->compose) nreg = 1;
            if (nreg > 16) nreg = 16;
            sub.region_end.resize(nreg);
            for (int r = 0; r < nreg; ++r) {
                int64_t e = sub.out_elems * (r + 1) / nreg;
                e = (e + 1) & ~(int64_t)1;  // keep 16-byte aligned pieces
                sub.region_end[r] = std::min(e, sub.out_elems);
            }
            sub.region_end[nreg - 1] = sub.out_elems;
        }
        // lists: "mine" (owner key in [fn_lo,fn_hi)) and "later" (owner key >= fn_hi)
        int mine_id[3], later_id[3];
        const bool whole = (fn_lo == 0 && fn_hi >= pl->norb);
        for (int t = 0; t < 3; ++t) {
            std::vector<char> pm(all[t].n), pl8(all[t].n);
            bool any_later = false;
            for (int q = 0; q < all[t].n; ++q) {
                pm[q] = (all[t].owner_fn[q] >= fn_lo && all[t].owner_fn[q] < fn_hi);
                pl8[q] = (all[t].owner_fn[q] >= fn_hi);
                any_later = any_later || pl8[q];
            }
            pl->lists.emplace_back();
            mine_id[t] = (int)pl->lists.size() - 1;
            if ((rc = upload_list(pl.get(), whole ? all[t] : sublist(all[t], pm), pl->lists.back()))) return rc;
            later_id[t] = -1;
            if (any_later) {
                pl->lists.emplace_back();
                later_id[t] = (int)pl->lists.size() - 1;
                if ((rc = upload_list(pl.get(), sublist(all[t], pl8), pl->lists.back()))) return rc;
            }
        }
        stage("list upload");
        // pair ids of this piece (compose mode): lists in the order lid = 2*type + kind
        std::vector<int32_t> fpinfo;
        if (pl->compose) {
            std::memset(&sub.comp, 0, sizeof(sub.comp));
            for (auto& b : sub.comp.launch_base) b = INT64_MIN;
            sub.ng = 0;
            for (int lid = 0; lid < 6; ++lid) {
                sub.listbase[lid] = sub.ng;
                const int id = (lid & 1) == 0 ? mine_id[lid >> 1] : later_id[lid >> 1];
                if (id >= 0) sub.ng += pl->lists[id].n;
            }
            if (sub.ng >= (1 << 25)) return fail(MYQC_ERR_UNSUPPORTED, "too many shell pairs for the compose pass");
            sub.h_rowrel.assign((size_t)sub.ng * 6, 0u);
        }
        // (SP SP|SP SP): which kernel.  The class kernel gives a contracted quartet to one lane, so it needs tens of
        // thousands of quartets to fill the machine and lasts ~0.23 ms however few there are; the warp-cooperative
        // kernel spreads the primitive quartets of ONE quartet over a warp and wins for small pieces (kWarpQuartetsPP).
        // MYQC_PP_KERNEL=warp / slices and MYQC_SP_KERNEL=warp / class force one of them.
        {
            auto count = [&](int ui, int ti, bool tri) -> int64_t {
                if (ui < 0 || ti < 0) return 0;
                const DevList& U = pl->lists[ui];
                const DevList& T = pl->lists[ti];
                if (U.n == 0 || T.n == 0) return 0;
                const std::vector<int32_t> ntv = row_prefix(U.host, T.host, pl->schwarz_tau);
                int64_t n = 0;
                for (int u = 0; u < U.n; ++u) n += std::max(0, ntv[u] - (tri ? u : 0));
                return n;
            };
            sub.pp_quartets = count(mine_id[2], mine_id[2], true) + count(mine_id[2], later_id[2], false);
            const int mode = pp_kernel_mode();
            sub.pp_warp = mode == 1 || (mode < 0 && sub.pp_quartets < kWarpQuartetsPP);
            sub.sp_quartets = count(mine_id[1], mine_id[2], false) + count(mine_id[1], later_id[2], false) + count(later_id[1], mine_id[2], false);
            const int smode = sp_kernel_mode();
            sub.sp_warp = smode == 1 || (smode < 0 && sub.sp_quartets < kWarpQuartetsSP);
            if (trace) std::fprintf(stderr, "[myqc trace]   (SP SP|SP SP): %lld contracted quartets -> %s kernel; (S SP|SP SP): %lld -> %s kernel\n",
                                    (long long)sub.pp_quartets, sub.pp_warp ? "warp-cooperative" : "mu-slice class",
                                    (long long)sub.sp_quartets, sub.sp_warp ? "warp-cooperative" : "class");
        }
        // launches.  Class (ta,tb), ta <= tb, uniform side = ta, lane side = tb.
        for (int ta = 0; ta < 3; ++ta)
            for (int tb = ta; tb < 3; ++tb) {
                if (ta == tb) {
                    if ((rc = add_launch(pl.get(), sub, mine_id[ta], mine_id[ta], true, 0, 0))) return rc;
                    if (later_id[ta] >= 0 && (rc = add_launch(pl.get(), sub, mine_id[ta], later_id[ta], false, 0, 1))) return rc;
                } else {
                    if ((rc = add_launch(pl.get(), sub, mine_id[ta], mine_id[tb], false, 0, 0))) return rc;
                    if (later_id[tb] >= 0 && (rc = add_launch(pl.get(), sub, mine_id[ta], later_id[tb], false, 0, 1))) return rc;
                    if (later_id[ta] >= 0 && (rc = add_launch(pl.get(), sub, later_id[ta], mine_id[tb], false, 1, 0))) return rc;
                }
            }
        if ((rc = finalize_launches(pl.get(), sub))) return rc;
        if (pl->compose) {
            // tables of the compose pass: function pair -> (pair id, slot), pair id -> (list index, type, list kind),
            // Schwarz factors, stage rows; work units of kCompRows rows x kCompCols columns, row block by row block
            ComposeArgs& c = sub.comp;
            fpinfo.assign((size_t)pl->npair, -1);
            std::vector<int2> pmeta((size_t)sub.ng, make_int2(0, 0));
            std::vector<double> pq((size_t)sub.ng, 0.0);
            for (int lid = 0; lid < 6; ++lid) {
                const int t = lid >> 1;
                const int id = (lid & 1) == 0 ? mine_id[t] : later_id[t];
                if (id < 0) continue;
                const PairList& h = pl->lists[id].host;
                const int nf = pt_nf(t);
                for (int k = 0; k < h.n; ++k) {
                    const int g = sub.listbase[lid] + k;
                    pmeta[g] = make_int2(k * nf, lid);
                    if ((int)h.qmax.size() == h.n) pq[g] = h.qmax[k];
                    for (int f = 0; f < nf; ++f) {
                        const int32_t P = h.pidx[(size_t)k * nf + f];
                        if (P >= 0) fpinfo[(size_t)P] = (g << 4) | f;
                    }
                }
            }
            int32_t *d_fp = nullptr, *d_urb = nullptr;
            int2* d_pm = nullptr;
            double* d_pq = nullptr;
            uint32_t* d_rb = nullptr;
            if ((rc = upload(pl.get(), fpinfo, &d_fp))) return rc;
            if ((rc = upload(pl.get(), pmeta, &d_pm))) return rc;
            if (pl->schwarz_tau > 0.0 && (rc = upload(pl.get(), pq, &d_pq))) return rc;
            if ((rc = upload(pl.get(), sub.h_rowrel, &d_rb))) return rc;
            const int64_t n = pl->norb;
            c.row_lo = (int64_t)fn_lo * n - (int64_t)fn_lo * (fn_lo - 1) / 2;
            c.row_hi = fn_hi >= n ? pl->npair : (int64_t)fn_hi * n - (int64_t)fn_hi * (fn_hi - 1) / 2;
            const int64_t ncb_all = (pl->npair + kCompCols - 1) / kCompCols;
            c.nrb = (int)((c.row_hi - c.row_lo + kCompRows - 1) / kCompRows);
            std::vector<int32_t> urb((size_t)c.nrb + 1, 0);
            for (int rb = 0; rb < c.nrb; ++rb) {
                const int64_t r0 = c.row_lo + (int64_t)rb * kCompRows;
                const int64_t ncb = ncb_all - r0 / kCompCols;
                if ((int64_t)urb[rb] + ncb > INT32_MAX) return fail(MYQC_ERR_UNSUPPORTED, "too many compose units");
                urb[rb + 1] = urb[rb] + (int32_t)ncb;
            }
            c.nunits = urb[c.nrb];
            if ((rc = upload(pl.get(), urb, &d_urb))) return rc;
            c.out_offset = sub.out_offset; c.npair = pl->npair;
            c.rk = pl->d_rk; c.cut = pl->d_cut; c.fpinfo = d_fp; c.pmeta = d_pm; c.pq = d_pq; c.rowrel = d_rb;
            c.tau = d_pq ? pl->schwarz_tau : 0.0;
            c.urb = d_urb;
            c.all_zero = std::getenv("MYQC_COMPOSE_ZERO") ? std::atoi(std::getenv("MYQC_COMPOSE_ZERO")) : 0;
            if (pl->ncounters + 1 > kMaxCounters) return fail(MYQC_ERR_UNSUPPORTED, "too many launches in one plan");
            c.counter = pl->d_counters + pl->ncounters;
            pl->ncounters += 1;
            sub.h_rowrel.clear(); sub.h_rowrel.shrink_to_fit();
        }
        pl->nlaunch += (int)sub.region_end.size();  // zero fills of this piece (compose mode: the compose launch)
        if (pl->screened_fill) {
            // pacing table: one entry per class-kernel task counter, weighted by the launch's estimated
            // duration (upper bound of its primitive quartets x measured time per quartet of its class)
            static const double kPsPerQuartet[6] = {5.1, 9.8, 29.0, 31.0, 107.0, 640.0};  // (H2O)_64, profiles/r1_notes.md
            std::vector<int32_t> pidx, pn;
            std::vector<float> pw;
            double wsum = 0.0;
            for (const Launch& L : sub.launches) {
                const int ns = launch_count(L);
                const double w = L.weight * kPsPerQuartet[class_id(L.UT, L.TT)] / ns;
                for (int k = 0; k < ns; ++k) {
                    pidx.push_back((int32_t)(L.args.row_counter - pl->d_counters) + k);
                    pn.push_back(L.args.ntasks);
                    pw.push_back((float)w);
                    wsum += w;
                }
            }
            for (float& w : pw) w = (float)(w / (wsum > 0 ? wsum : 1.0));
            int32_t *d_pidx = nullptr, *d_pn = nullptr;
            float* d_pw = nullptr;
            if ((rc = upload(pl.get(), pidx, &d_pidx))) return rc;
            if ((rc = upload(pl.get(), pn, &d_pn))) return rc;
            if ((rc = upload(pl.get(), pw, &d_pw))) return rc;
            const int64_t n = pl->norb;
            const int64_t row_lo = (int64_t)fn_lo * n - (int64_t)fn_lo * (fn_lo - 1) / 2;
            const int64_t row_hi = (int64_t)fn_hi * n - (int64_t)fn_hi * (fn_hi - 1) / 2;
            if ((rc = build_sub_fill(pl.get(), sub, row_lo, row_hi))) return rc;
            sub.fill.counters = pl->d_counters;
            sub.fill.prog_idx = d_pidx; sub.fill.prog_n = d_pn; sub.fill.prog_w = d_pw;
            sub.fill.nprog = (int)pidx.size();
        }
    }
    pl->h_rk.clear(); pl->h_rk.shrink_to_fit();
    pl->h_cut.clear(); pl->h_cut.shrink_to_fit();
    if (pl->compose && pl->stage_elems > 0) {
        // zeroed once: blocks of quartets that are never evaluated are never written, and the compose pass may read them
        cudaError_t e = cudaMalloc((void**)&pl->d_stage, (size_t)pl->stage_elems * sizeof(double));
        if (e != cudaSuccess)
            return fail(MYQC_ERR_NOMEM, std::string("device allocation of the quartet staging array (") +
                                            std::to_string(8e-9 * (double)pl->stage_elems) + " GB): " + cudaGetErrorString(e));
        pl->dev_allocs.push_back(pl->d_stage);
        CU(cudaMemset(pl->d_stage, 0, (size_t)pl->stage_elems * sizeof(double)));
        if (trace) std::fprintf(stderr, "[myqc trace]   staging array %.3f GB for a slice of %.3f GB\n", 8e-9 * (double)pl->stage_elems, 8e-9 * (double)pl->out_elems);
    }
    stage("task lists");
    // internal streams and events
    CU(cudaStreamCreateWithFlags(&pl->s_fill, cudaStreamNonBlocking));
    for (auto& st : pl->s_comp) CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&pl->e_start, cudaEventDisableTiming));
    for (auto& e : pl->e_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    {
        size_t nfill = 0;
        for (const Sub& sub : pl->subs) nfill += sub.region_end.size();
        pl->e_fill.resize(nfill);
        for (auto& e : pl->e_fill) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }

    {
        const char* fe = std::getenv("MYQC_FILL_ENGINE");
        // Default: cudaMemsetAsync.  On a B200 the driver's fill writes 7.35 TB/s; thirty variants of an SM fill kernel
        // (128/256-bit stores, tiles, cache operators, TMA bulk stores of a zeroed shared-memory tile) all end at
        // 6.2-6.55 TB/s (profiles/r2_notes.md section 6).  MYQC_FILL_ENGINE=kernel keeps the repo's fill_zero_kernel,
        // =copy uses device-to-device copies from a zero buffer on the copy engines (4.5 TB/s, measured).
        pl->fill_engine = fe ? (std::strcmp(fe, "kernel") == 0 ? 0 : std::strcmp(fe, "copy") == 0 ? 2 : 1) : 1;
        if (pl->screened_fill || pl->compose) pl->fill_engine = 0;
        const char* fs = std::getenv("MYQC_FILL_STREAMS");
        pl->fill_nstreams = fs ? std::max(1, std::min((int)myqc_eri_plan::kMaxFillStreams, std::atoi(fs))) : 1;
        if (pl->fill_engine == 2) {
            const char* zm = std::getenv("MYQC_ZERO_MB");
            pl->zero_bytes = (size_t)(zm ? std::max(1, std::atoi(zm)) : 32) << 20;
            CU(cudaMalloc(&pl->d_zero, pl->zero_bytes));
            pl->dev_allocs.push_back(pl->d_zero);
            CU(cudaMemset(pl->d_zero, 0, pl->zero_bytes));
        }
        if (pl->fill_engine) {
            for (int i = 1; i < pl->fill_nstreams; ++i) CU(cudaStreamCreateWithFlags(&pl->s_fillx[i], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&pl->e_fork, cudaEventDisableTiming));
            for (int i = 1; i < pl->fill_nstreams; ++i) CU(cudaEventCreateWithFlags(&pl->e_join[i], cudaEventDisableTiming));
        }
    }
    stage("streams + events");
    // the canonical work statistics cost ~8 ms of host time: evaluated when plan_stats asks for them
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
    std::vector<Shell> shells;
    if ((rc = build_shells(nnuc, nset, setl, setinfo, ops, basinfo, shells, err))) return fail(rc, err);
    PairList all[3];
    if ((rc = build_pairs(nnuc, xyz, set, setinfo, setl, ops, bas, basinfo, shells, all, err))) return fail(rc, err);
    CutTable ct;
    if (nshards == 1) { offsets[0] = 0; offsets[1] = packed_row_offset(basinfo[1], basinfo[1]); return MYQC_OK; }
    int64_t nq_ref[6];
    double fl_ref = 0.0;
    canonical_stats(nnuc, xyz, nset, setl, set, setinfo, nq_ref, &fl_ref);
    if ((rc = build_cut_table(shells, all, nq_ref, nshards, ct, err))) return fail(rc, err);
    const std::vector<int> ext = split_range(ct, 0, (int)ct.cuts.size(), nshards);
    for (int k = 0; k <= nshards; ++k) offsets[k] = packed_row_offset(cut_to_fn(ct, ext[k], basinfo[1]), basinfo[1]);
    return MYQC_OK;
}

// Host only: what the cut model expects of each of `nshards` shards -- seconds in the class kernels and in the zero
// fill (class_s[nshards], fill_s[nshards]).  For tools/exp_shard_times.py, which puts the measured times next to them.
int myqc_eri_shard_model(int nnuc, const double* xyz, int nset, int setl, const double* set,
                         const int32_t* setinfo, int ops, const double* bas, const int32_t* basinfo,
                         int nshards, double* class_s, double* fill_s) {
    static const double dummy_ft[1] = {0.0};
    int rc = check_args(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, dummy_ft);
    if (rc) return rc;
    if (nshards < 1 || !class_s || !fill_s) return fail(MYQC_ERR_BAD_ARG, "bad nshards/outputs");
    std::string err;
    std::vector<Shell> shells;
    if ((rc = build_shells(nnuc, nset, setl, setinfo, ops, basinfo, shells, err))) return fail(rc, err);
    PairList all[3];
    if ((rc = build_pairs(nnuc, xyz, set, setinfo, setl, ops, bas, basinfo, shells, all, err))) return fail(rc, err);
    CutTable ct;
    int64_t nq_ref[6];
    double fl_ref = 0.0;
    canonical_stats(nnuc, xyz, nset, setl, set, setinfo, nq_ref, &fl_ref);
    if ((rc = build_cut_table(shells, all, nq_ref, nshards, ct, err))) return fail(rc, err);
    const std::vector<int> ext = split_range(ct, 0, (int)ct.cuts.size(), nshards);
    for (int k = 0; k < nshards; ++k) {
        class_s[k] = fill_s[k] = 0.0;
        for (int c = ext[k]; c < ext[k + 1]; ++c) { class_s[k] += ct.wclass[c]; fill_s[k] += ct.wfill[c]; }
    }
    return MYQC_OK;
}

int64_t myqc_eri_plan_out_offset(const myqc_eri_plan* plan) { return plan ? plan->out_offset : -1; }
int64_t myqc_eri_plan_out_elems(const myqc_eri_plan* plan) { return plan ? plan->out_elems : -1; }

// serial order of launches: for each sub-shard and fill region: its zero fill, then the class
// kernels restricted to the tasks of that region
static int plan_launch_total(const myqc_eri_plan* plan) {
    int n = 0;
    for (const Sub& sub : plan->subs)
        n += (int)sub.region_end.size() * (1 + (int)sub.launches.size());
    return n;
}

// launch the tasks of fill region r of one class launch (its own task counter per region)
static int launch_region(myqc_eri_plan* plan, Sub& sub, Launch& L, int r, int slice, double* d_sub_out, cudaStream_t st) {
    const int t0 = L.region_task[r], t1 = L.region_task[r + 1];
    if (t1 <= t0) return 0;
    ClassArgs a = L.args;
    a.out = plan->compose ? nullptr : d_sub_out;
    a.stage = plan->compose ? plan->d_stage : nullptr;
    a.tasks = L.args.tasks + t0;
    a.ntasks = t1 - t0;
    a.row_counter = L.args.row_counter + r * launch_count(L);
    return launch_class(L.UT, L.TT, L.warp ? -1 : slice, a, plan->num_sms, st);
}

static int fill_region(myqc_eri_plan* plan, Sub& sub, int r, double* d_sub_out, cudaStream_t st, bool paced = false) {
    if (plan->screened_fill) {
        FillArgs f = sub.fill;
        f.out = d_sub_out;
        static const bool no_pace = std::getenv("MYQC_FILL_NOPACE") != nullptr;
        if (!paced || no_pace) f.nprog = 0;
        return launch_fill_screened(f, plan->num_sms, st);
    }
    const int64_t b = r == 0 ? 0 : sub.region_end[r - 1], e = sub.region_end[r];
    if (plan->fill_engine) {
        if (r == 0 && sub.ncounters > 0) {
            cudaError_t ce = cudaMemsetAsync(plan->d_counters + sub.counter_base, 0, sizeof(int) * (size_t)sub.ncounters, st);
            if (ce != cudaSuccess) return (int)ce;
        }
        char* p0 = reinterpret_cast<char*>(d_sub_out + b);
        const size_t bytes = (size_t)(e - b) * sizeof(double);
        const int ns = plan->fill_nstreams;
        if (ns > 1) {
            cudaEventRecord(plan->e_fork, st);
            for (int i = 1; i < ns; ++i) cudaStreamWaitEvent(plan->s_fillx[i], plan->e_fork, 0);
        }
        // stream i takes the i-th part (cut at 1 MB); copies go in pieces of the zero buffer's size
        const size_t part = ((bytes + ns - 1) / ns + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        for (int i = 0; i < ns; ++i) {
            const size_t lo = std::min(bytes, part * i), hi = std::min(bytes, part * (i + 1));
            cudaStream_t si = i == 0 ? st : plan->s_fillx[i];
            cudaError_t ce = cudaSuccess;
            if (plan->fill_engine == 1) {
                if (hi > lo) ce = cudaMemsetAsync(p0 + lo, 0, hi - lo, si);
            } else {
                for (size_t o = lo; o < hi && ce == cudaSuccess; o += plan->zero_bytes)
                    ce = cudaMemcpyAsync(p0 + o, plan->d_zero, std::min(plan->zero_bytes, hi - o), cudaMemcpyDeviceToDevice, si);
            }
            if (ce != cudaSuccess) return (int)ce;
        }
        for (int i = 1; i < ns; ++i) {
            cudaEventRecord(plan->e_join[i], plan->s_fillx[i]);
            cudaStreamWaitEvent(st, plan->e_join[i], 0);
        }
        return (int)cudaGetLastError();
    }
    // the first fill of a sub-shard also resets all of its task counters
    return launch_fill_zero(d_sub_out + b, e - b, plan->d_counters + sub.counter_base, r == 0 ? sub.ncounters : 0,
                            plan->num_sms, st);
}

static int compose_sub(myqc_eri_plan* plan, Sub& sub, double* d_sub_out, cudaStream_t st) {
    ComposeArgs c = sub.comp;
    c.out = d_sub_out;
    c.stage = plan->d_stage;
    return launch_compose(c, plan->num_sms, st);
}

int myqc_eri_plan_execute(myqc_eri_plan* plan, double* d_out, void* stream) {
    if (!plan || (!d_out && plan->out_elems > 0)) return fail(MYQC_ERR_BAD_ARG, "null plan or output");
    CU(cudaSetDevice(plan->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // MYQC_TIMELINE=1: completion time of every launch on its internal stream, printed on stderr
    // (a debugging aid: it synchronises the device at the end of the call)
    static const bool timeline = std::getenv("MYQC_TIMELINE") != nullptr;
    std::vector<std::pair<std::string, cudaEvent_t>> tl;
    auto mark = [&](const std::string& name, cudaStream_t s) {
        if (!timeline) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        tl.emplace_back(name, e);
    };
    if (plan->screened_fill || plan->compose) CU(cudaMemsetAsync(plan->d_counters, 0, sizeof(int) * (size_t)plan->ncounters, st));
    CU(cudaMemsetAsync(plan->d_pq, 0, sizeof(unsigned long long) * (size_t)kMaxCounters, st));
    mark("start", st);
    if (plan->compose) {
        // class kernels of all pieces side by side on the internal streams (they only write the staging array),
        // then one compose launch per piece on the caller's stream
        CU(cudaEventRecord(plan->e_start, st));
        for (auto& sc : plan->s_comp) CU(cudaStreamWaitEvent(sc, plan->e_start, 0));
        int rr = 0;
        for (Sub& sub : plan->subs)
            for (Launch& L : sub.launches)
                for (int slice = 0; slice < launch_count(L); ++slice) {
                    const int si = rr++ % myqc_eri_plan::kNumCompute;
                    int e = launch_region(plan, sub, L, 0, slice, nullptr, plan->s_comp[si]);
                    if (e) return cuda_fail((cudaError_t)e, "class kernel launch");
                    mark("class{" + std::to_string(L.UT) + "," + std::to_string(L.TT) + "} on stream " + std::to_string(si), plan->s_comp[si]);
                }
        for (int i = 0; i < myqc_eri_plan::kNumCompute; ++i) {
            CU(cudaEventRecord(plan->e_done[i], plan->s_comp[i]));
            CU(cudaStreamWaitEvent(st, plan->e_done[i], 0));
        }
        for (Sub& sub : plan->subs) {
            int e = compose_sub(plan, sub, d_out + (sub.out_offset - plan->out_offset), st);
            if (e) return cuda_fail((cudaError_t)e, "compose launch");
            mark("compose", st);
        }
        if (timeline) {
            cudaDeviceSynchronize();
            for (size_t k = 1; k < tl.size(); ++k) {
                float t = 0;
                cudaEventElapsedTime(&t, tl[0].second, tl[k].second);
                std::fprintf(stderr, "[myqc timeline] %-28s done at %8.3f ms\n", tl[k].first.c_str(), t);
            }
            for (auto& x : tl) cudaEventDestroy(x.second);
        }
        return MYQC_OK;
    }
    // fork: internal streams start after whatever is already queued on the caller's stream
    CU(cudaEventRecord(plan->e_start, st));
    CU(cudaStreamWaitEvent(plan->s_fill, plan->e_start, 0));
    for (auto& sc : plan->s_comp) CU(cudaStreamWaitEvent(sc, plan->e_start, 0));
    int ef = 0;
    for (Sub& sub : plan->subs) {
        double* d_sub = d_out + (sub.out_offset - plan->out_offset);
        for (int r = 0; r < (int)sub.region_end.size(); ++r) {
            int e = fill_region(plan, sub, r, d_sub, plan->s_fill, true);
            if (e) return cuda_fail((cudaError_t)e, "fill launch");
            CU(cudaEventRecord(plan->e_fill[ef++], plan->s_fill));
            mark("fill", plan->s_fill);
        }
    }
    int rr = 0;
    ef = 0;
    for (Sub& sub : plan->subs) {
        double* d_sub = d_out + (sub.out_offset - plan->out_offset);
        for (int r = 0; r < (int)sub.region_end.size(); ++r, ++ef) {
            bool waited[myqc_eri_plan::kNumCompute] = {};
            for (Launch& L : sub.launches) {
                if (L.region_task[r + 1] <= L.region_task[r]) continue;
                // the mu-slices of (SP SP|SP SP) are independent launches: one internal stream each
                for (int slice = 0; slice < launch_count(L); ++slice) {
                    const int si = rr++ % myqc_eri_plan::kNumCompute;
                    // plain fill: the slice must be zeroed before a class kernel stores into it; the
                    // screened fill writes a disjoint set of elements and needs no ordering
                    if (!waited[si] && !plan->screened_fill) { CU(cudaStreamWaitEvent(plan->s_comp[si], plan->e_fill[ef], 0)); waited[si] = true; }
                    int e = launch_region(plan, sub, L, r, slice, d_sub, plan->s_comp[si]);
                    if (e) return cuda_fail((cudaError_t)e, "class kernel launch");
                    mark("class{" + std::to_string(L.UT) + "," + std::to_string(L.TT) + "} on stream " + std::to_string(si), plan->s_comp[si]);
                }
            }
        }
    }
    // join
    for (int i = 0; i < myqc_eri_plan::kNumCompute; ++i) {
        CU(cudaEventRecord(plan->e_done[i], plan->s_comp[i]));
        CU(cudaStreamWaitEvent(st, plan->e_done[i], 0));
    }
    CU(cudaEventRecord(plan->e_done[myqc_eri_plan::kNumCompute], plan->s_fill));
    CU(cudaStreamWaitEvent(st, plan->e_done[myqc_eri_plan::kNumCompute], 0));
    if (timeline) {
        mark("join", st);
        cudaDeviceSynchronize();
        for (size_t k = 1; k < tl.size(); ++k) {
            float t = 0;
            cudaEventElapsedTime(&t, tl[0].second, tl[k].second);
            std::fprintf(stderr, "[myqc timeline] %-28s done at %8.3f ms\n", tl[k].first.c_str(), t);
        }
        for (auto& x : tl) cudaEventDestroy(x.second);
    }
    return MYQC_OK;
}

int myqc_eri_plan_launch_count(const myqc_eri_plan* plan) {
    if (!plan) return 0;
    return plan_launch_total(plan);
}

int myqc_eri_plan_launch_info(const myqc_eri_plan* plan, int k, int* cls, int* tri, int64_t* rows) {
    if (!plan || k < 0 || k >= plan_launch_total(plan)) return fail(MYQC_ERR_BAD_ARG, "bad launch index");
    for (const Sub& sub : plan->subs) {
        for (int r = 0; r < (int)sub.region_end.size(); ++r) {
            const int n = 1 + (int)sub.launches.size();
            if (k >= n) { k -= n; continue; }
            if (plan->compose) {  // class launches first, the compose launch of the piece last (cls = -2)
                if (k == n - 1) {
                    if (cls) *cls = -2;
                    if (tri) *tri = 0;
                    if (rows) *rows = sub.out_elems;
                    return MYQC_OK;
                }
                ++k;
            }
            if (k == 0) {  // the zero fill of this region
                if (cls) *cls = -1;
                if (tri) *tri = 0;
                if (rows) *rows = plan->screened_fill ? sub.fill_zero_elems : sub.region_end[r] - (r == 0 ? 0 : sub.region_end[r - 1]);
                return MYQC_OK;
            }
            const Launch& L = sub.launches[k - 1];
            if (cls) *cls = class_id(L.UT, L.TT);
            if (tri) *tri = L.args.part[0].tri;
            if (rows) *rows = L.region_task[r + 1] - L.region_task[r];
            return MYQC_OK;
        }
    }
    return fail(MYQC_ERR_BAD_ARG, "bad launch index");
}

// Serialised pass on the caller's stream with CUDA events around every launch: the per-kernel
// durations behind bench.py's roofline (the production execute() overlaps launches on internal streams).
int myqc_eri_plan_execute_timed(myqc_eri_plan* plan, double* d_out, void* stream, float* ms) {
    if (!plan || !ms || (!d_out && plan->out_elems > 0)) return fail(MYQC_ERR_BAD_ARG, "null plan/output/ms");
    CU(cudaSetDevice(plan->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n = plan_launch_total(plan);
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) CU(cudaEventCreate(&e));
    if (plan->screened_fill || plan->compose) CU(cudaMemsetAsync(plan->d_counters, 0, sizeof(int) * (size_t)plan->ncounters, st));
    CU(cudaMemsetAsync(plan->d_pq, 0, sizeof(unsigned long long) * (size_t)kMaxCounters, st));
    CU(cudaEventRecord(ev[0], st));
    int idx = 0;
    for (Sub& sub : plan->subs) {
        double* d_sub = d_out + (sub.out_offset - plan->out_offset);
        if (plan->compose) {
            int e = 0;
            for (Launch& L : sub.launches) {
                for (int slice = 0; slice < launch_count(L) && !e; ++slice) e = launch_region(plan, sub, L, 0, slice, nullptr, st);
                if (e) return cuda_fail((cudaError_t)e, "class kernel launch");
                CU(cudaEventRecord(ev[++idx], st));
            }
            e = compose_sub(plan, sub, d_sub, st);
            if (e) return cuda_fail((cudaError_t)e, "compose launch");
            CU(cudaEventRecord(ev[++idx], st));
            continue;
        }
        for (int r = 0; r < (int)sub.region_end.size(); ++r) {
            int e = fill_region(plan, sub, r, d_sub, st);
            if (e) return cuda_fail((cudaError_t)e, "fill_zero launch");
            CU(cudaEventRecord(ev[++idx], st));
            for (Launch& L : sub.launches) {
                for (int slice = 0; slice < launch_count(L) && !e; ++slice) e = launch_region(plan, sub, L, r, slice, d_sub, st);
                if (e) return cuda_fail((cudaError_t)e, "class kernel launch");
                CU(cudaEventRecord(ev[++idx], st));
            }
        }
    }
    CU(cudaEventSynchronize(ev[n]));
    for (int k = 0; k < n; ++k) CU(cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]));
    for (auto& x : ev) cudaEventDestroy(x);
    return MYQC_OK;
}

// After an execute (synchronises the device): primitive quartets the class kernels actually evaluated, per class
// {0,0},{0,1},{0,2},{1,1},{1,2},{2,2} -- the reference's rule minus what the Schwarz skip left out (the (SP SP|SP SP)
// slices each evaluate every quartet of their class: their common counter is divided by four).
int myqc_eri_plan_executed_quartets(myqc_eri_plan* plan, int64_t* nq, double* schwarz_tau) {
    if (!plan || !nq) return fail(MYQC_ERR_BAD_ARG, "null plan/output");
    CU(cudaSetDevice(plan->device));
    CU(cudaDeviceSynchronize());
    std::vector<unsigned long long> h(kMaxCounters, 0ull);
    CU(cudaMemcpy(h.data(), plan->d_pq, sizeof(unsigned long long) * (size_t)kMaxCounters, cudaMemcpyDeviceToHost));
    for (int c = 0; c < 6; ++c) nq[c] = 0;
    for (const Sub& sub : plan->subs)
        for (const Launch& L : sub.launches) {
            const size_t idx = (size_t)(L.args.pq_counter - plan->d_pq);
            nq[class_id(L.UT, L.TT)] += (int64_t)(h[idx] / (unsigned long long)launch_count(L));
        }
    if (schwarz_tau) *schwarz_tau = plan->schwarz_tau;
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
    if (nlaunch) *nlaunch = plan->nlaunch;
    return MYQC_OK;
}

void myqc_eri_plan_destroy(myqc_eri_plan* plan) {
    if (!plan) return;
    cudaSetDevice(plan->device);
    if (plan->s_fill) { cudaStreamSynchronize(plan->s_fill); cudaStreamDestroy(plan->s_fill); }
    for (auto& st : plan->s_comp) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    if (plan->e_start) cudaEventDestroy(plan->e_start);
    for (auto& e : plan->e_done) if (e) cudaEventDestroy(e);
    for (auto& e : plan->e_fill) if (e) cudaEventDestroy(e);
    for (auto& st : plan->s_fillx) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    if (plan->e_fork) cudaEventDestroy(plan->e_fork);
    for (auto& e : plan->e_join) if (e) cudaEventDestroy(e);
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
// myqc_eri_release_cache() drops them; MYQC_NO_CACHE=1 turns the cache off.
namespace myqc {

void host_zero_stream(double* p, size_t n);  // hostmem.cpp
void host_zero_fence();

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
    const bool hit = (ce.plan != nullptr && ce.key == key);
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
    if (n > 0 && !packed_slice) { if (no_cache) drop_entry(local); return fail(MYQC_ERR_BAD_ARG, "null output"); }
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
            // chunk size of the sparse route in doubles (MYQC_XFER_CHUNK = 32 / 64 / 128 / 256)
            int chunk = myqc::kXferChunkDefault;
            if (const char* ec = std::getenv("MYQC_XFER_CHUNK")) { const int c = std::atoi(ec); if (myqc::xfer_chunk_ok(c)) chunk = c; }
            // measurement hooks (the result is then incomplete): leave out the host-side zeros or the device-side push
            const bool plain_memset = std::getenv("MYQC_HOST_MEMSET") != nullptr;  // measurement hook: glibc memset instead of streaming stores
            const bool no_host = std::getenv("MYQC_XFER_NOHOST") != nullptr, no_push = std::getenv("MYQC_XFER_NOPUSH") != nullptr;
            const int64_t nchunk = (n + chunk - 1) / chunk;
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
            if (e == cudaSuccess) e = (cudaError_t)myqc::launch_chunk_flags(d_out, n, chunk, d_flags, sms, nullptr);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_flags, d_flags, (size_t)nchunk, cudaMemcpyDeviceToHost, nullptr);
            if (e == cudaSuccess) e = cudaEventRecord(ev_flags, nullptr);
            if (e == cudaSuccess && !no_push)
                e = (cudaError_t)myqc::launch_chunk_push(d_out, n, chunk, d_flags, static_cast<double*>(pa.devicePointer), sms, nullptr);
            if (e == cudaSuccess) e = cudaEventSynchronize(ev_flags);
            const auto tf = now();
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
                        const int64_t e0 = c * chunk, e1 = std::min<int64_t>(n, r * chunk);
                        if (!no_host) {
                            if (plain_memset) std::memset(packed_slice + e0, 0, (size_t)(e1 - e0) * sizeof(double));
                            else myqc::host_zero_stream(packed_slice + e0, (size_t)(e1 - e0));
                        }
                        c = r;
                    }
                    myqc::host_zero_fence();
                    sent += mine;
                };
                std::vector<std::thread> th;
                const int64_t per = (nchunk + nthr - 1) / nthr;
                for (int t = 1; t < nthr; ++t)
                    if (t * per < nchunk) th.emplace_back(zero_range, t * per, std::min<int64_t>(nchunk, (t + 1) * per));
                zero_range(0, std::min<int64_t>(nchunk, per));
                for (auto& t : th) t.join();
                const auto th1 = now();
                xfer_frac = (double)sent.load() / (double)nchunk;
                myqc::g_last_d2h_bytes = sent.load() * chunk * (int64_t)sizeof(double) + nchunk;
                xfer_kind = "sparse push";
                e = cudaStreamSynchronize(nullptr);
                if (trace)
                    std::fprintf(stderr, "[myqc trace]   sparse route: chunk %d B, flags ready after %.1f ms, host zeros %.1f ms (%d threads), push done %.1f ms after the zeros\n",
                                 chunk * 8, ms(t3, tf), ms(tf, th1), nthr, ms(th1, now()));
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
    // MYQC_OVERSUBSCRIBE=1 (test hook): more shards than devices, mapped round-robin onto the devices there are
    const bool oversub = std::getenv("MYQC_OVERSUBSCRIBE") != nullptr;
    if (ngpu <= 0 || (ngpu > ndev && !oversub)) ngpu = ndev;
    if (!packed) return fail(MYQC_ERR_BAD_ARG, "null output");
    std::vector<int64_t> off(ngpu + 1, 0);
    int rc = myqc_eri_shard_layout(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, ngpu, off.data());
    if (rc) return rc;
    std::vector<int> rcs(ngpu, 0);
    std::vector<std::string> errs(ngpu);
    auto work = [&](int g) {
        rcs[g] = myqc_eri_packed_shard(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, ftab,
                                       packed + off[g], g % ndev, g, ngpu, nullptr);
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

// Dense XX on `ngpu` devices.  One device: packed array and XX both on it.  Several: every device computes
// its shard of the packed array into a host copy; every device then takes the whole packed array and expands
// the slab XX(:,:,:,h) of its range of h, so the 8n^4-byte stream is produced once, in parallel, and no device
// needs more than the packed array plus 1/ngpu of XX.
int myqc_eri_dense(int nnuc, const double* xyz, int nset, int setl, const double* set,
                   const int32_t* setinfo, int ops, const double* bas, const int32_t* basinfo,
                   const double* ftab, double* xx, int ngpu) {
    int rc = check_args(nnuc, xyz, nset, setl, set, setinfo, ops, bas, basinfo, ftab);
    if (rc) return rc;
    if (!xx) return fail(MYQC_ERR_BAD_ARG, "null output");
    const int ndev = myqc_device_count();
    if (ndev == 0) return fail(MYQC_ERR_NO_DEVICE, "no CUDA device: the ERI engine has no CPU fallback");
    const bool oversub = std::getenv("MYQC_OVERSUBSCRIBE") != nullptr;
    if (ngpu <= 0 || (ngpu > ndev && !oversub)) ngpu = ndev;
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
        cudaError_t e = cudaSetDevice(g % ndev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g % ndev);
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
