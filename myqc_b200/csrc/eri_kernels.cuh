// Device-side interface of the quartet-class kernel family (see eri_kernels.cu).
#pragma once
#include <cstdint>

namespace myqc {

// One launch = all quartets (u, v) with u in a "uniform-side" pair list (records staged to
// shared memory by TMA bulk copy, one row of the quartet space per CTA iteration) and v in a
// "lane-side" pair list (structure-of-arrays, one pair per lane).
struct ClassArgs {
    // uniform side (AoS records [nU][9][nfield(UT)])
    const double* u_aos;
    const int32_t* u_nprim;  // [nU]
    const int32_t* u_fi;     // [nU][nf(UT)]
    const int32_t* u_fj;
    const int32_t* u_diag;   // [nU] shell A == shell B
    const int32_t* u_ntv;    // [nU] number of lane-side pairs passing emax_u*emax_v >= 1e-14
    int nU;
    // lane side (SoA [9][nfield(TT)][t_npad])
    const double* t_soa;
    const int32_t* t_nprim;  // [nT]
    const int32_t* t_fi;     // [nT][nf(TT)]
    const int32_t* t_fj;
    const int32_t* t_diag;
    int t_npad;
    int nT;
    int tri;  // lists are the same list: take v >= u only, and P1 <= P2 when v == u
    // Boys table for this class's start order Q: [121][8] = {Ft(t,Q+k)/k!, k=0..6 ; t/10}
    const double* ftab_q;
    // output
    double* out;         // this shard's slice of the packed array
    int64_t out_offset;  // packed index of out[0]
    int64_t out_elems;
    int norb;
    int64_t npair;
};

// returns cudaError_t as int; `slice` selects the mu-slice for (2,2), ignored otherwise
int launch_class(int UT, int TT, const ClassArgs& a, int num_sms, void* stream);
// number of kernel launches launch_class issues for (UT,TT)
int class_nlaunch(int UT, int TT);

int measure_dfma_peak(int num_sms, double* tflops);
int launch_fill_zero(double* out, int64_t n, int num_sms, void* stream);
int launch_expand_dense(const double* packed, int norb, double* xx, int num_sms, void* stream);

}  // namespace myqc
