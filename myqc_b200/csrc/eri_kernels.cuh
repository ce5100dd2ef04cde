// Device-side interface of the owner-row strip kernels (see eri_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace myqc {

constexpr int kBlockShells = 128;  // partner second shells D are handled in blocks of this many consecutive shells

// One launch of eri_strip_kernel<UT,TC,USL>: every task (owner shell pair u = (A,B) of type UT;
// a range of partner first shells C of type TC) computes all quartets (u | C D), D >= C, and writes
// the part of u's packed rows whose columns start in C -- values and zeros -- exactly once.
struct StripArgs {
    // shells, ordered by first orbital
    const int4* sh_fn;         // [ns] orbital ids of the slots (s,px,py,pz), -1 where absent
    const signed char* sh_type;  // [ns] 0 = S, 1 = SP
    const int32_t* sh_first;   // [ns+1] first orbital of each shell; sh_first[ns] = norb
    int ns, nblk, norb;        // nblk = ceil(ns / kBlockShells)
    int64_t npair;
    // shell-pair tables, entry (C,D), C <= D, at C*ns + D
    const int32_t* pair_rec;   // record index inside the pair's kind list, -1: no primitive pair survives
    const double* pair_emax;   // largest primitive prefactor (0 where pair_rec < 0)
    const double* blk_emax;    // [ns][nblk] max of pair_emax over D in a block, D >= C
    // partner segments: for (C, block b, shell class c) the D >= C of class c in block b that form a live pair with C,
    // by decreasing emax(C,D): entries seg_start[(C*nblk + b)*ncls + c] .. seg_start[.. + 1]
    const int32_t* seg_start;
    const unsigned short* seg_d;
    const double2* seg_eprof;  // per entry two double2: {E(1) = emax, E(3)}, {E(5), E(7)}: the pair's primitive prefactors by rank (0 past the last)
    const int32_t* cls_kind;   // [ncls] 0: shells of the class are S shells, 1: SP shells
    int ncls;
    // first shells C this launch handles: the shells of type TC in ascending order
    const int32_t* clist;
    // uniform side: records of kind UT, [n][9][nfield(UT)]
    const double* u_aos;
    const int32_t* u_nprim;
    // lane side, partner kinds TC + TD for TD = 0 (D is an S shell) and 1 (D is an SP shell)
    const double* t_aos[2];    // [n][9][nfield]
    const int32_t* t_nprim[2];
    // Boys Taylor tables [121][8] = {Ft(t,Q+k)/k!} for the two start orders, and {exp(-k/10), k/10}
    const double* ftab_q[2];
    const double2* exptab;
    // tasks {A | B << 16, first index into clist, end index, b0 | b1 << 16}: owner pair (A,B), partner first shells
    // clist[first..end), starting at partner block b0 of the first and ending before block b1 of the last;
    // pulled by warps from *counter
    const int4* tasks;
    int ntasks;
    int* counter;
    unsigned long long* stats;  // [2] primitive quartets evaluated per TD (atomics, one per chunk)
    double* out;                // this shard's slice of the packed array
    int64_t out_offset;         // packed index of out[0]
};

// Kernel variants: UT in 0..2, TC in 0..1; the SP.SP owner with SP first partner shells runs as four
// launches, one per first function of the owner (slice 0..3).
int strip_nslices(int UT, int TC);
// returns cudaError_t as int
int launch_strip(int UT, int TC, int slice, const StripArgs& a, int num_sms, void* stream);
// sets the kernel attributes and forces the (lazily loaded) kernels of the current device to load
int prepare_kernels();

int measure_dfma_peak(int num_sms, double* tflops);
// dense XX(:,:,:,h0:h1-1) (column-major, all 8 images) gathered from the whole packed array
int launch_expand_dense(const double* packed, int norb, int h0, int h1, double* xx_slab, int num_sms, void* stream);

// Sparse device -> host transfer of a packed slice (most of a large molecule's integrals are exact zeros):
// chunks of kXferChunk doubles; flags[c] = 1 iff chunk c holds a nonzero bit pattern; the push kernel
// stores the flagged chunks straight into device-accessible (pinned) host memory.
constexpr int kXferChunk = 256;
int launch_chunk_flags(const double* out, int64_t n, unsigned char* flags, int num_sms, void* stream);
int launch_chunk_push(const double* out, int64_t n, const unsigned char* flags, double* host, int num_sms, void* stream);

}  // namespace myqc
