// Device-side interface of the quartet-class kernel family (see eri_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace myqc {

constexpr int kTaskPairs = 256;  // most lane-side pairs per task (8 warp chunks)
// The heavier the class, the shorter the tasks: a task is executed by one warp from start to finish,
// so its duration bounds the tail of the launch (a 256-pair (SP SP|SP SP) task with all 81 primitive
// quartets alive would run for more than a millisecond).
constexpr int class_task_pairs(int UT, int TT) {
    return (UT + TT <= 1) ? 256 : (UT + TT == 2) ? 128 : (UT + TT == 3) ? 64 : 32;
}

// One launch = all quartets (u, v) with u in a "uniform-side" pair list (records staged to
// shared memory by TMA bulk copy, one row of the quartet space per CTA iteration) and v in a
// "lane-side" pair list (structure-of-arrays, one pair per lane).
// The pair lists one part of a launch works on.  An unsharded plan has one part per class; a shard has up to three
// (own x own, own x later, later x own), merged into ONE launch per class so that a shard pays one launch tail per
// class, not three: task.w selects the part.
struct PartArgs {
    // uniform side (AoS records [nU][9][nfield(UT)])
    const double* u_aos;
    const int32_t* u_nprim;  // [nU]
    const int32_t* u_pidx;   // [nU][nf(UT)] packed pair index of each function pair, -1 = not stored
    const double* u_q;       // [nU] Schwarz factor of the uniform-side pairs
    // lane side (SoA [9][nfield(TT)][t_npad])
    const double* t_soa;
    const double* t_aos;     // the same records as [nT][9][nfield(TT)]: one lane reads its own pair contiguously
    const int32_t* t_nprim;  // [nT]
    const int32_t* t_pidx;   // [nT][nf(TT)]
    const double* t_q;       // [nT] Schwarz factor of the lane-side pairs
    const int64_t* stage_row;  // [nU] compose mode, see below
    int t_npad;
    int tri;  // lists are the same list: take v >= u only, and P1 <= P2 when v == u
};
constexpr int kMaxParts = 3;

// One launch = all quartets (u, v) with u in a "uniform-side" pair list (records staged to
// shared memory by TMA bulk copy, one row of the quartet space per CTA iteration) and v in a
// "lane-side" pair list (structure-of-arrays, one pair per lane).
struct ClassArgs {
    PartArgs part[kMaxParts];
    int nparts;
    // work list: task = {row u, first lane-side pair, end lane-side pair, part}; rows are cut into
    // pieces of at most kTaskPairs lane-side pairs (longest rows first)
    const int4* tasks;
    int ntasks;
    // Boys Taylor table for this class's start order Q: [121][8] = {Ft(t,Q+k)/k!, k=0..6 ; 0}
    const double* ftab_q;
    // [601] {exp(-k/10), k/10}
    const double2* exptab;
    // global row counter of this launch (zeroed by the fill kernel that precedes it)
    int* row_counter;
    // Schwarz skip (tau = 0: off): lanes leave out the quartets with u_q[u]*t_q[v] < tau
    double tau;
    unsigned long long* pq_counter;  // primitive quartets this launch evaluated (one atomic per task)
    // output
    double* out;         // this shard's slice of the packed array (scatter mode; nullptr in compose mode)
    int64_t out_offset;  // packed index of out[0]
    int64_t npair;
    // compose mode: every quartet (u, v) of a part owns a block of nf(UT)*nf(TT) doubles, [f_u][f_v], at
    // stage + stage_row[u] + v*nf(UT)*nf(TT); the lane that evaluates the quartet writes the whole block with 128-bit
    // stores and compose_kernel gathers the packed array from the blocks.  stage == nullptr: scatter into `out`.
    double* stage;
};

// Compose pass (compose_kernel): the packed slice is written ONCE, in order, by 256-byte warp stores: exact zeros
// where the reference's rule (or the Schwarz skip) leaves a quartet out, else the value gathered from the quartet
// blocks the class kernels staged.  Work unit = kCompRows packed rows x kCompCols columns.
constexpr int kCompRows = 64;
constexpr int kCompThreads = 128;
constexpr int kCompColsPerThread = 4;
constexpr int kCompCols = kCompThreads * kCompColsPerThread;
struct ComposeArgs {
    double* out;         // this slice of the packed array
    int64_t out_offset;  // packed index of out[0]
    int64_t npair;
    int64_t row_lo, row_hi;  // packed rows of the slice
    // reference rule as an integer compare (see build_screen_ranks): (P,P') is kept iff rk[P'] < cut[P]
    const int32_t* rk;       // [npair]
    const int32_t* cut;      // [npair]
    // Shell pairs of the piece are numbered g = 0..ng-1 list by list, lists in the order lid = 2*type + kind
    // (kind 0: the piece's own pairs, 1: pairs of later pieces), so that the uniform side of a quartet is simply
    // the pair with the smaller g.
    const int32_t* fpinfo;   // [npair] (g << 4) | function-pair slot f inside the shell pair; -1: no shell pair kept
    const int2* pmeta;       // [ng] {list index * nf(type), lid}
    const double* pq;        // [ng] Schwarz factors (nullptr when tau == 0)
    // block(u,v) = stage + launch_base[lid_u][lid_v] + (rowrel[g_u][lid_v] + v*nf_u*nf_v + f_u*nf_v + f_v mod 2^32)
    const uint32_t* rowrel;  // [ng][6]
    int64_t launch_base[36]; // [lid_u*6 + lid_v]; INT64_MIN: no such launch
    const double* stage;
    double tau;
    const int32_t* urb;      // [nrb+1] prefix sums of column blocks per row block
    int nrb, nunits;
    int* counter;            // unit counter (zeroed before the launch)
    int all_zero;            // experiments (MYQC_COMPOSE_ZERO=1): write zeros only, the ceiling of the store pattern
};
int launch_compose(const ComposeArgs& a, int num_sms, void* stream);

// Screened zero fill (see fill_screened_kernel): rows [row_lo,row_hi) of the packed upper triangle
constexpr int kFillRows = 64;
constexpr int kFillCols = 2048;
struct FillArgs {
    double* out;         // this slice of the packed array
    int64_t out_offset;  // packed index of out[0]
    int64_t npair;
    int64_t row_lo, row_hi;
    const int32_t* rk;   // [npair] rank of the function pair's shell-pair prefactor (INT32_MAX: no shell pair kept)
    const int32_t* cut;  // [npair] number of ranks that pass the screen against this row
    const int32_t* ucb;  // [ncb+1] prefix sums of row blocks per column block, column blocks cb0..cb0+ncb-1
    int cb0, ncb, nunits;
    int* counter;        // unit counter (zeroed before the launch)
    int all;             // experiments: ignore the screen, zero every element (only valid before the class kernels)
    // pacing: the fill is throttled to the progress of the class kernels that run next to it, so that
    // its stores are spread over their whole duration instead of saturating the memory pipeline up front.
    // progress = sum_k prog_w[k] * min(counters[prog_idx[k]], prog_n[k]) / prog_n[k]  (weights sum to 1);
    // nprog = 0: no pacing (serial timing pass, or nothing to pace against)
    const int* counters;
    const int32_t* prog_idx;
    const int32_t* prog_n;
    const float* prog_w;
    int nprog;
    int sleep_ns;  // experiments: fixed delay per four rows written (a constant-rate fill)
};
int launch_fill_screened(const FillArgs& a, int num_sms, void* stream);
// sets the kernel attributes and forces the (lazily loaded) kernels of the current device to load
int prepare_kernels();

// returns cudaError_t as int; `slice` < class_nlaunch(UT,TT) selects the mu-slice of (2,2) (0 otherwise; -1: see below);
// slice k uses the task counter a.row_counter + k
int launch_class(int UT, int TT, int slice, const ClassArgs& a, int num_sms, void* stream);
// number of kernel launches launch_class issues for (UT,TT)
int class_nlaunch(int UT, int TT);
// (SP SP|SP SP) by the warp-cooperative kernel (a warp per contracted quartet, lanes over its primitive quartets):
// launch_class(2, 2, slice = -1, ...).  pp_kernel_mode(): MYQC_PP_KERNEL = warp (1) / slices (0) / unset (-1: the plan
// decides per piece from the number of contracted quartets)
int pp_kernel_mode();
int sp_kernel_mode();  // the same for (S SP|SP SP): MYQC_SP_KERNEL = warp / class

// unscreened diagonal integrals (f|f) of the shell pairs of kind T (0..2) into diag[packed pair index]
int launch_diag(int T, const double* aos, const int32_t* nprim, const int32_t* pidx, int n, const double* ftab_q,
                const double2* exptab, double* diag, void* stream);

int measure_dfma_peak(int num_sms, double* tflops);
int launch_fill_zero(double* out, int64_t n, int* counters, int ncounters, int num_sms, void* stream);
// dense XX(:,:,:,h0:h1-1) (column-major, all 8 images) gathered from the whole packed array
int launch_expand_dense(const double* packed, int norb, int h0, int h1, double* xx_slab, int num_sms, void* stream);

// Sparse device -> host transfer of a packed slice (most of a large molecule's integrals are exact zeros):
// chunks of `chunk` doubles (32, 64, 128 or 256); flags[c] = 1 iff chunk c holds a nonzero bit pattern; the push
// kernel stores the flagged chunks straight into device-accessible (pinned) host memory.
constexpr int kXferChunkDefault = 32;  // 256 bytes: one warp store; measured best end to end (profiles/r2_notes.md section 7)
int xfer_chunk_ok(int chunk);
int launch_chunk_flags(const double* out, int64_t n, int chunk, unsigned char* flags, int num_sms, void* stream);
int launch_chunk_push(const double* out, int64_t n, int chunk, const unsigned char* flags, double* host, int num_sms, void* stream);

}  // namespace myqc
