// Device-side interface of the quartet-class kernel family (see eri_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace myqc {

constexpr int kTaskPairs = 256;  // lane-side pairs per task (8 warp chunks)

// One launch = all quartets (u, v) with u in a "uniform-side" pair list (records staged to
// shared memory by TMA bulk copy, one row of the quartet space per CTA iteration) and v in a
// "lane-side" pair list (structure-of-arrays, one pair per lane).
struct ClassArgs {
    // uniform side (AoS records [nU][9][nfield(UT)])
    const double* u_aos;
    const int32_t* u_nprim;  // [nU]
    const int32_t* u_pidx;   // [nU][nf(UT)] packed pair index of each function pair, -1 = not stored
    int nU;
    // work list: task = {row u, first lane-side pair, end lane-side pair, 0}; rows are cut into
    // pieces of at most kTaskPairs lane-side pairs (longest rows first)
    const int4* tasks;
    int ntasks;
    // lane side (SoA [9][nfield(TT)][t_npad])
    const double* t_soa;
    const int32_t* t_nprim;  // [nT]
    const int32_t* t_pidx;   // [nT][nf(TT)]
    int t_npad;
    int nT;
    int tri;  // lists are the same list: take v >= u only, and P1 <= P2 when v == u
    // Boys Taylor table for this class's start order Q: [121][8] = {Ft(t,Q+k)/k!, k=0..6 ; 0}
    const double* ftab_q;
    // [601] {exp(-k/10), k/10}
    const double2* exptab;
    // global row counter of this launch (zeroed by the fill kernel that precedes it)
    int* row_counter;
    // output
    double* out;         // this shard's slice of the packed array
    int64_t out_offset;  // packed index of out[0]
    int64_t npair;
};

// returns cudaError_t as int; `slice` selects the mu-slice for (2,2), ignored otherwise
int launch_class(int UT, int TT, const ClassArgs& a, int num_sms, void* stream);
// number of kernel launches launch_class issues for (UT,TT)
int class_nlaunch(int UT, int TT);

int measure_dfma_peak(int num_sms, double* tflops);
int launch_fill_zero(double* out, int64_t n, int* counters, int ncounters, int num_sms, void* stream);
int launch_expand_dense(const double* packed, int norb, double* xx, int num_sms, void* stream);

}  // namespace myqc
