/*
 * myqc_eri.h -- C-ABI of the B200-native two-electron-integral (ERI) engine that replaces
 * myQC's `int2e` hot path.
 *
 * The reference has no FFI/plugin interface: its boundary is (i) the executable `int2e`
 * spawned by the driver (src/myQC/myQC.f90:54) and (ii) inside it the single call
 *     CALL proc2e(bas,basinfo,atoms,options,fmem,nnuc,xyz,set,setinfo,maxL)
 * (src/integrals/int2e.f90:66, interface :78-112).  The entry points below are what an
 * iso_c_binding shim for that call binds (INTEGRATION.md shows the Fortran side), plus the
 * file-level pieces of PROGRAM int2e (int2e.f90:14-69) used by our drop-in `int2e` binary.
 *
 * Conventions
 *   - plain C types only; every array is the 0-based content of the Fortran DIMENSION(0:)
 *     array of the same name, passed by reference.
 *   - xyz is Fortran xyz(0:nnuc-1,0:2): element (i,c) at xyz[i + nnuc*c], bohr.
 *   - set[nset], setinfo[2+setl*nset] (setinfo[0]=nset, setinfo[1]=setl=7),
 *     bas[ops*nset] (ops=4), basinfo[2+5*norb] (basinfo[0]=ops, basinfo[1]=norb):
 *     exactly what buildBasis produces (src/myQC/basis.f90:101-205).
 *   - ftab[t + 121*j] = Ft(t,j), t=0..120, j=0..22: the bytes of the `Ftab` record
 *     (int2e.f90:118,161-163).  Never regenerated from a formula.
 *   - caller owns all pointers; the library copies inputs and owns all device memory (plans own their
 *     device buffers until destroyed; the one-shot calls keep the last plan and device slice per device
 *     until myqc_eri_release_cache(), see below).
 *   - return 0 on success, negative MYQC_ERR_* otherwise; the library never exits the process.
 *     The shim maps non-zero to `touch error` (int2e.f90:174-178).
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns
 *     MYQC_ERR_NO_DEVICE.
 *
 * Packed layout (8-fold-symmetry-unique integrals)
 *   pair index   P(i,j) = i*norb - i(i-1)/2 + (j-i),       0 <= i <= j < norb
 *   npair        = norb(norb+1)/2
 *   quartet index(P,P') = P*npair - P(P-1)/2 + (P'-P),     P <= P'
 *   P is monotone in the reference's key i*norb+j, so {P <= P'} is exactly the reference's
 *   canonical set (int2e.f90:686,692,695).
 */
#ifndef MYQC_ERI_H
#define MYQC_ERI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MYQC_OK 0
#define MYQC_ERR_NO_DEVICE (-1)   /* no CUDA device / driver */
#define MYQC_ERR_CUDA (-2)        /* a CUDA call failed; see myqc_last_error() */
#define MYQC_ERR_UNSUPPORTED (-3) /* l > 1, or a set layout other than S / SP / P */
#define MYQC_ERR_BAD_ARG (-4)     /* inconsistent sizes / null pointers */
#define MYQC_ERR_IO (-5)          /* file missing or malformed */
#define MYQC_ERR_NOMEM (-6)       /* host or device allocation failed */

/* Human-readable description of the last error on this thread ("" if none). */
const char *myqc_last_error(void);

/* Number of visible CUDA devices (0 if none; never fails). */
int myqc_device_count(void);

/* ---- one-shot calls with HOST buffers: the body of proc2e (int2e.f90:78-351) ---------------
 * xx:     out, caller-allocated 8*norb^4 bytes, Fortran XX(0:n-1,0:n-1,0:n-1,0:n-1) column-major,
 *         offset of (i,j,g,h) = i + n*(j + n*(g + n*h)), all 8 symmetry images filled
 *         (what fillsym leaves, int2e.f90:290-304,540-554).
 * packed: out, caller-allocated npair(npair+1)/2 doubles, layout above.
 * ngpu:   number of devices to shard over (1..myqc_device_count()); 0 = all visible.  myqc_eri_dense with
 *         ngpu > 1: every device computes its shard of the packed array, then expands the slab XX(:,:,:,h)
 *         of its range of h from the whole packed array (no device holds more than packed + XX/ngpu).  */
int myqc_eri_dense(int nnuc, const double *xyz, int nset, int setl, const double *set,
                   const int32_t *setinfo, int ops, const double *bas, const int32_t *basinfo,
                   const double *ftab, double *xx, int ngpu);

int myqc_eri_packed(int nnuc, const double *xyz, int nset, int setl, const double *set,
                    const int32_t *setinfo, int ops, const double *bas, const int32_t *basinfo,
                    const double *ftab, double *packed, int ngpu);

/* One shard of the packed array into a HOST buffer of (offsets[shard+1]-offsets[shard]) doubles
 * (myqc_eri_shard_layout), computed on `device`: what one rank of a one-process-per-GPU job calls.
 * h2d_bytes (optional) receives the bytes of pair/Boys tables uploaded for the call.          */
int myqc_eri_packed_shard(int nnuc, const double *xyz, int nset, int setl, const double *set,
                          const int32_t *setinfo, int ops, const double *bas, const int32_t *basinfo,
                          const double *ftab, double *packed_slice, int device, int shard, int nshards,
                          int64_t *h2d_bytes);

/* Bytes that crossed the device -> host link in this thread's last myqc_eri_packed_shard call.  A
 * pinned (device-accessible) destination of at least 4 Mi elements takes the sparse route: the slice
 * is cut into 256-byte chunks (MYQC_XFER_CHUNK = 32 / 64 / 128 / 256 doubles), the GPU stores the chunks
 * that hold a nonzero straight into the host buffer and host threads write the zeros of the others with
 * streaming stores (MYQC_SPARSE_D2H=0 turns it off, MYQC_HOST_THREADS sets the number of zeroing
 * threads); any other destination gets one cudaMemcpy of the slice.                                  */
int64_t myqc_eri_last_d2h_bytes(void);

/* Host only: n doubles at p (8-byte aligned) become +0.0, written with the non-temporal stores the sparse route uses for
 * the zeros of the unflagged chunks (scalar head and tail up to the first / after the last whole cache line). */
void myqc_host_zero(double *p, int64_t n);

/* The one-shot calls (myqc_eri_packed, myqc_eri_packed_shard, multi-device myqc_eri_dense) keep the plan (pair
 * tables, task lists) and the device slice of the last call per device, keyed on the bytes of every input, so a
 * caller that asks for the same integrals again pays for them once (MYQC_NO_CACHE=1 turns this off).  This
 * releases what is kept.                                                                                  */
void myqc_eri_release_cache(void);

/* ---- plan API: device-resident execution, one plan per (GPU, shard) -------------------------
 * A plan holds the shell-pair tables of one shard of the canonical quartet space on one device.
 * Shard s of nshards owns a contiguous block of rows of the packed array (rows = bra pair index
 * P), cut at shell boundaries and balanced by model flops; shards are independent (no
 * collective).  nshards=1 is the whole problem.
 * Execution knobs read at plan creation (defaults are the measured best):
 *   MYQC_FILL_ENGINE = memset (default: cudaMemsetAsync) | kernel (the repo's fill kernel) | copy (copy engines)
 *   MYQC_PP_KERNEL   = warp | slices   (SP SP|SP SP) on the warp-cooperative kernel / four mu-slices of the class
 *                      kernel; unset: per piece, warp below 20 000 contracted quartets
 *   MYQC_SP_KERNEL   = warp | class    the same choice for (S SP|SP SP); unset: class
 *   MYQC_OUTPUT_MODE = compose         staged quartet blocks + one pass that writes every element once       */
typedef struct myqc_eri_plan myqc_eri_plan;

int myqc_eri_plan_create(int nnuc, const double *xyz, int nset, int setl, const double *set,
                         const int32_t *setinfo, int ops, const double *bas,
                         const int32_t *basinfo, const double *ftab, int device, int shard,
                         int nshards, myqc_eri_plan **plan);

/* Offset (in doubles) of this shard's slice inside the full packed array, and its length. */
int64_t myqc_eri_plan_out_offset(const myqc_eri_plan *plan);
int64_t myqc_eri_plan_out_elems(const myqc_eri_plan *plan);

/* Compute the shard: d_out is a DEVICE pointer to myqc_eri_plan_out_elems() doubles on the
 * plan's device; stream is a cudaStream_t (NULL = default stream).  Asynchronous: returns
 * after enqueueing.  Every element of the slice is written (zeros where the reference's
 * screen leaves zeros).                                                                      */
int myqc_eri_plan_execute(myqc_eri_plan *plan, double *d_out, void *stream);

/* Work statistics of the shard (all optional, may be NULL):
 *   nquartets[6]  canonical primitive quartets surviving the reference screen per class
 *                 {0,0},{0,1},{0,2},{1,1},{1,2},{2,2} (class = #SP sets in bra pair, ket pair)
 *   model_flops   sum_class nquartets*W_class, W = {60,99,228,228,693,2691}
 *   nlaunch       kernels one execute() enqueues                                              */
int myqc_eri_plan_stats(const myqc_eri_plan *plan, int64_t *nquartets, double *model_flops,
                        int *nlaunch);

void myqc_eri_plan_destroy(myqc_eri_plan *plan);

/* Host-only (no device needed): offsets[0..nshards] of the shard slices in the packed array;
 * shard s owns [offsets[s], offsets[s+1]).  Every rank computes the same layout independently. */
int myqc_eri_shard_layout(int nnuc, const double *xyz, int nset, int setl, const double *set,
                          const int32_t *setinfo, int ops, const double *bas, const int32_t *basinfo,
                          int nshards, int64_t *offsets);

/* Host only: what the shard-cut cost model expects of each of `nshards` shards -- seconds in the class kernels
 * (class_s[nshards]) and in the zero fill (fill_s[nshards]); tools/exp_shard_times.py puts measured times next to them. */
int myqc_eri_shard_model(int nnuc, const double *xyz, int nset, int setl, const double *set,
                         const int32_t *setinfo, int ops, const double *bas, const int32_t *basinfo,
                         int nshards, double *class_s, double *fill_s);

/* Host-only: canonical surviving primitive-quartet counts per class and their model flops for
 * the whole molecule (same definition as myqc_eri_plan_stats).                  */
int myqc_eri_canonical_stats(int nnuc, const double *xyz, int nset, int setl, const double *set,
                              const int32_t *setinfo, int64_t *nquartets, double *model_flops);

/* Measurement hooks (bench.py): launches of one execute() = 1 zero fill + the class kernels.
 * launch_info: cls = -1 for the zero fill, else the class id 0..5 in the order of nquartets[];
 * rows = elements filled / uniform-side rows.  execute_timed brackets every launch with CUDA
 * events on `stream`, synchronises, and returns the per-launch milliseconds in ms[launch_count]. */
int myqc_eri_plan_launch_count(const myqc_eri_plan *plan);
int myqc_eri_plan_launch_info(const myqc_eri_plan *plan, int k, int *cls, int *tri, int64_t *rows);
int myqc_eri_plan_execute_timed(myqc_eri_plan *plan, double *d_out, void *stream, float *ms);

/* After an execute (synchronises the device): primitive quartets the class kernels evaluated, per class in the order of
 * nquartets[] above.  It is the canonical count of the reference's rule minus what the Schwarz skip leaves out:
 * a contracted quartet (u|v) is skipped when Q_u*Q_v < tau, Q = sqrt(max (ij|ij)) over the pair's function pairs
 * (unscreened diagonals, computed on the device at plan creation), so that every skipped integral is < tau in
 * magnitude.  tau defaults to 1e-12 (MYQC_SCHWARZ_TAU; 0 = the reference's rule alone) and is returned in *schwarz_tau. */
int myqc_eri_plan_executed_quartets(myqc_eri_plan *plan, int64_t *nq, double *schwarz_tau);

/* Register-resident DFMA microbenchmark: the FP64 (non-tensor) roofline denominator, measured
 * on the device the plan runs on (MEASURED_PEAKS.json has no FP64 figure).                    */
int myqc_fp64_peak(int device, double *tflops);

/* Expand a packed DEVICE array into the dense DEVICE array XX(n,n,n,n) (all 8 images).      */
int myqc_eri_expand_dense(const double *d_packed, int norb, double *d_xx, void *stream);

/* ---- file layer of PROGRAM int2e (int2e.f90:14-69) ------------------------------------------ */
/* getenv, src/myQC/env.f90:16-73: read envdat / nucpos / fmem from `dir`.
 * atoms[cap_nuc], xyz[3*cap_nuc] (Fortran layout with leading dim = returned nnuc),
 * options[cap_opt].                                                                          */
int myqc_read_env(const char *dir, int cap_nuc, int cap_opt, int *nnuc, int *nelcA, int *nelcB,
                  int32_t *atoms, double *xyz, double *fmem, int *nopt, int32_t *options);

/* buildBasis, src/myQC/basis.f90:23-226.  First call with null outputs to get sizes.
 * Also (re)writes `basinfo` and `setinfo` text files into out_dir unless out_dir is NULL.    */
int myqc_build_basis(const char *mybasis_path, int bkey, int nnuc, const int32_t *atoms,
                     int *nset_cap, int *norb_cap, double *set, int32_t *setinfo, double *bas,
                     int32_t *basinfo, int *maxN, int *maxL, const char *out_dir);

/* READ(1) Ft, int2e.f90:161-163.  ftab receives 2783 doubles.                               */
int myqc_read_ftab(const char *path, double *ftab);

/* WRITE(42) XX, int2e.f90:166,306-307: one Fortran unformatted sequential record, split into
 * gfortran subrecords above 2147483639 payload bytes.                                        */
int myqc_write_xx(const char *path, const double *xx, int norb);
/* same with an explicit maximum subrecord payload (bytes); used to test the subrecord framing */
int myqc_write_xx_ex(const char *path, const double *xx, int norb, int64_t max_subrecord);
int myqc_read_xx(const char *path, double *xx, int norb);

/* The whole program: what the `int2e` executable does in directory `dir`.
 * verbose mirrors the reference's stdout lines.                                              */
int myqc_int2e_main(const char *dir, int ngpu);

#ifdef __cplusplus
}
#endif
#endif /* MYQC_ERI_H */
