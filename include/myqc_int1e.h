/* myqc_int1e.h -- overlap matrix S and core Hamiltonian H = T + V on the device
 * (SURVEY.md 8f, row N2): the arithmetic of the reference's `int1e` program.
 *
 *   src/integrals/int1e.f90:14-131   PROGRAM int1e  (envdat/nucpos/fmem/mybasis/Ftab in; Suv, Huv out)
 *   src/integrals/int1e.f90:132-280  proc1e         (ordered loop over primitive sets, EIJ < 1e-14 skip)
 *   :321-386 overlap, :391-474 kinetic, :479-582 coulomb; auxilary.f90 getcoef/getDk/Boys/RNLMj
 *
 * Same conventions as myqc_eri.h: arrays are the 0-based contents of the reference's Fortran
 * arrays, xyz(0:nnuc-1,0:2) column-major in bohr, atoms(0:nnuc-1) nuclear charges, Ftab as
 * ftab[t + 121*j].  s and h receive norb x norb doubles, column-major (Suv / Huv of scf.f90:140-144).
 * Returns 0 or a negative MYQC_ERR_* code (message via myqc_last_error()); no CPU fallback.
 */
#ifndef MYQC_INT1E_H
#define MYQC_INT1E_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* CALL proc1e(S,F,bas,basinfo,atoms,options,fmem,nnuc,xyz,norb,set,setinfo), int1e.f90:103,132 */
int myqc_int1e(int nnuc, const double *xyz, const int32_t *atoms, int nset, int setl,
               const double *set, const int32_t *setinfo, int ops, const double *bas,
               const int32_t *basinfo, const double *ftab, double *s, double *h);

/* PROGRAM int1e in directory `dir`: reads envdat, nucpos, fmem, mybasis, Ftab; (re)writes basinfo and
 * setinfo; if Suv and Huv both exist it touches Sold / Hold and computes nothing (int1e.f90:98-111);
 * otherwise writes Suv and Huv as list-directed text (`WRITE(1,*) S(:,:)`, :265-271), leaves fmem net
 * unchanged, and touches `error` on failure.                                                        */
int myqc_int1e_main(const char *dir);

#ifdef __cplusplus
}
#endif
#endif /* MYQC_INT1E_H */
