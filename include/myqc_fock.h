/* myqc_fock.h -- G(D) (two-electron part of the Fock matrix) straight from the packed,
 * 8-fold-symmetry-unique ERI array that myqc_eri.h produces, on the device.
 *
 * Replaces the reference's per-iteration consumers of the dense XX file:
 *   src/I2G/RHFI2G.f90:72-95   READ(9) XX ; G(i,j) = sum_kl Da(k,l) [XX(i,j,k,l) - 1/4 XX(i,k,j,l) - 1/4 XX(i,l,j,k)]
 *   src/I2G/UHFI2G.f90:71-99   GA(i,j) = sum_kl (Da+Db)(k,l) XX(i,j,k,l) - Da(k,l) XX(i,k,j,l)   (GB with Db)
 * which scf.f90:159,323,878,1092 spawn once per SCF iteration.  The reference rereads 8 n^4 bytes
 * from disk and runs n^4 scalar iterations per call; here one pass over the packed array
 * (n^4/8 elements, exact zeros skipped) accumulates J and K, so the dense file is not needed and
 * sizes whose dense XX cannot exist (448 functions: 322 GB) still have a G build.
 *
 * Conventions
 *   - d_packed: DEVICE pointer to a slice [out_offset, out_offset+out_elems) of the packed array
 *     (layout of myqc_eri.h).  A slice must consist of whole packed rows (what
 *     myqc_eri_shard_layout returns); the result is then the slice's partial G, and the partial
 *     results of all slices add up to G (multi-GPU: one all-reduce of norb^2 doubles).
 *   - densities and results: DEVICE pointers, norb x norb doubles, column-major like the
 *     reference's Da(0:norb-1,0:norb-1) / Guv.  Densities are symmetric by construction
 *     (dens.f90:115-124,213-228 builds C C^T); the symmetric part (D + D^T)/2 is what is used,
 *     which is also exactly what the reference's RHF formula depends on.
 *   - return value 0 or a negative MYQC_ERR_* code (myqc_eri.h); message via myqc_last_error().
 *   - stream-ordered: temporaries are allocated and freed on `stream`, nothing is synchronised.
 */
#ifndef MYQC_FOCK_H
#define MYQC_FOCK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* RHFI2G.f90:80-90.  d_g receives G (overwritten). */
int myqc_fock_rhf(const double *d_packed, int64_t out_offset, int64_t out_elems, int norb,
                  const double *d_da, double *d_g, void *stream);

/* UHFI2G.f90:80-93.  d_ga / d_gb receive GuvA / GuvB (overwritten). */
int myqc_fock_uhf(const double *d_packed, int64_t out_offset, int64_t out_elems, int norb,
                  const double *d_da, const double *d_db, double *d_ga, double *d_gb, void *stream);

/* Optional sparsity mask.  The reference's screen leaves most of a large molecule's integrals exactly
 * zero; an SCF calls the G build 10-30 times on the same array.  myqc_fock_mask_build scans the slice once
 * (one streaming pass) and sets bit k of row P = (i,j) iff some (ij|kl), l >= k, is nonzero; the *_masked
 * builds then skip every all-zero (row, k) block without reading it.  Results are identical to the
 * unmasked calls (skipped blocks contribute exactly nothing).  d_mask: DEVICE pointer to
 * myqc_fock_mask_words(norb) uint32 words, indexed by the absolute packed row, so the shards of one
 * array can share one buffer; build fills only the rows of the slice it is given.                 */
int64_t myqc_fock_mask_words(int norb);
int myqc_fock_mask_build(const double *d_packed, int64_t out_offset, int64_t out_elems, int norb,
                         uint32_t *d_mask, void *stream);
int myqc_fock_rhf_masked(const double *d_packed, int64_t out_offset, int64_t out_elems, int norb,
                         const double *d_da, const uint32_t *d_mask, double *d_g, void *stream);
int myqc_fock_uhf_masked(const double *d_packed, int64_t out_offset, int64_t out_elems, int norb,
                         const double *d_da, const double *d_db, const uint32_t *d_mask,
                         double *d_ga, double *d_gb, void *stream);

/* Host-buffer convenience calls (everything copied in and out; packed is the whole array). */
int myqc_fock_rhf_host(const double *packed, int norb, const double *da, double *g);
int myqc_fock_uhf_host(const double *packed, int norb, const double *da, const double *db,
                       double *ga, double *gb);

#ifdef __cplusplus
}
#endif
#endif /* MYQC_FOCK_H */
