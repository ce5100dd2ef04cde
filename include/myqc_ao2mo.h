/* myqc_ao2mo.h -- AO -> MO four-index transformation of the packed, 8-fold-symmetry-unique ERI
 * array (myqc_eri.h) on the device: the arithmetic of the reference's
 * `ao2mo` program.
 *
 *   src/ao2mo/ao2mo.f90:1306-1439   idx1_trans .. idx4_trans: B(p,q,r,s) = sum_t A(t,q,r,s) x(t,p), ...
 *   src/ao2mo/ao2mo.f90:465-602     slow_ao2mo_MP2_RHF  (ia|jb) -> files ijab_AA, ijab_AB
 *   src/ao2mo/ao2mo.f90:614-904     slow_ao2mo_MP2_UHF  spin cases AA, BB, AB -> ijab_AA, ijab_BB, ijab_AB
 *   src/ao2mo/ao2mo.f90:919-1227    slow_ao2mo_CIS_UHF  -> ajib_AA, ajbi_AA, ajib_AB, ajib_BB, ajbi_BB
 *
 * Every one of those is the same operator with a different choice of coefficient column blocks:
 *
 *     O(p,q,r,s) = sum_{u,v,l,d} C1(u,p) C2(v,q) C3(l,r) C4(d,s) (uv|ld)
 *
 * which the reference evaluates as four explicit O(n^5) loop nests over the dense XX(n,n,n,n)
 * array read from disk.  Here the packed array stays in HBM, each quarter transformation is a
 * double-precision GEMM on the FP64 tensor pipe (DMMA m8n8k4, hand-written tiles; FP64 has no
 * tcgen05 path), and the pair symmetry halves both half-transformations:
 *     H(P; r,s)   = sum_{l,d} C3(l,r) (P|ld) C4(d,s)        for the npair unordered pairs P = (u<=v)
 *     O(p,q; r,s) = sum_{u,v} C1(u,p) H(uv; r,s) C2(v,q)
 *
 * Conventions (as myqc_eri.h / myqc_fock.h): coefficient blocks are column-major n x n_k with
 * leading dimension `norb` -- i.e. a pointer to column c0 of the reference's Cm(0:ntot-1,0:ntot-1)
 * is a valid block, exactly the array sections ao2mo.f90 passes (`Cm(0:ntot-1,noccA:ntot-1)`).
 * The result is column-major O(0:n1-1,0:n2-1,0:n3-1,0:n4-1) (p fastest), the reference's Om.
 * Returns 0 or a negative MYQC_ERR_* code (message via myqc_last_error()); no CPU fallback.
 */
#ifndef MYQC_AO2MO_H
#define MYQC_AO2MO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Device pointers; d_packed is the whole packed array (npair(npair+1)/2 doubles).  Scratch is
 * allocated and freed on `stream` (stream-ordered); the call does not synchronise.               */
int myqc_ao2mo_transform(const double *d_packed, int norb, const double *d_c1, int n1,
                         const double *d_c2, int n2, const double *d_c3, int n3,
                         const double *d_c4, int n4, double *d_out, void *stream);

/* Same with caller-provided scratch (at least myqc_ao2mo_workspace_bytes(...) bytes of device memory,
 * 16-byte aligned): nothing is allocated inside, so repeated calls (one per spin case / file) reuse the
 * buffer and the call can be captured in a CUDA graph.  The dominant part is the half-transformed
 * array H, npair * n3 * n4 doubles.                                                                 */
int64_t myqc_ao2mo_workspace_bytes(int norb, int n1, int n2, int n3, int n4);
int myqc_ao2mo_transform_ws(const double *d_packed, int norb, const double *d_c1, int n1,
                            const double *d_c2, int n2, const double *d_c3, int n3,
                            const double *d_c4, int n4, double *d_out, void *d_workspace,
                            int64_t workspace_bytes, void *stream);

/* Host buffers in and out (device 0).                                                             */
int myqc_ao2mo_transform_host(const double *packed, int norb, const double *c1, int n1,
                              const double *c2, int n2, const double *c3, int n3,
                              const double *c4, int n4, double *out);

/* Packed array from the dense XX(n,n,n,n) the reference reads (`READ(100) Km`, ao2mo.f90:494-495):
 * host helper, packed[tri(P(i,j),P(k,l))] = xx[i + n(j + n(k + n l))], i<=j, k<=l, P<=P'.          */
int myqc_pack_dense(const double *xx, int norb, double *packed);

/* PROGRAM ao2mo in directory `dir` (ao2mo.f90:22-98): reads envdat/nucpos/fmem (getenv), basinfo,
 * XX, Cui; options(1)=1 & options(3)=0 -> MP2/RHF files, options(1)=1 & options(3)=1 -> MP2/UHF
 * files, options(13)=1 & options(1)=0 & options(3)=1 -> CIS/UHF files; anything else prints the
 * reference's message and touches `error`.  Records are Fortran unformatted sequential.           */
int myqc_ao2mo_main(const char *dir);

/* FLOPs of one transform as this library executes it (for the roofline): 2 n^2 n4 npair +
 * 2 n n3 n4 npair + 2 n^2 n2 n3 n4 + 2 n n1 n2 n3 n4.                                              */
double myqc_ao2mo_flops(int norb, int n1, int n2, int n3, int n4);

/* Measured throughput of the FP64 tensor pipe of `device` in TFLOP/s (register-resident DMMA m8n8k4
 * microbenchmark): the roofline denominator of the GEMM stages, next to myqc_fp64_peak (DFMA).       */
int myqc_dmma_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* MYQC_AO2MO_H */
