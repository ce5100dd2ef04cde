/* myqc_parse.h -- the `parse` stage of myQC: ZMAT -> nucpos, envdat, fmem,
 * the three text files every later stage reads through getenv (src/myQC/env.f90:16-73).
 *
 *   src/parser/parser.f90:22-108    PROGRAM parser   (defaults :60, dispatch on the first token :74-85)
 *   :622-680  cartesian     atom lines until the END marker
 *   :683-735  read_options  KEY= VALUE lines, case-sensitive keys, unknown keys reported and ignored
 *   :147-431  getsys .. get_prop   value tables of the 16 option keys
 *   :435-547  build         close-contact check, centre-of-mass shift (masses 1,4,7,9,11,12,14,16,19,20),
 *                           Angstrom -> bohr (x 1.8897161646320724), electron counts, file writers
 *   :767-800  check_options contradictory options -> `touch error`
 *
 * Host-only (no GPU involved).  Returns 0 or a negative MYQC_ERR_* code (myqc_eri.h); message via
 * myqc_last_error().  Where the reference is undefined (its centre-of-mass accumulator `temp` is never
 * initialised) zero is used, which reproduces the geometry of examples/NO/MOLDEN.
 */
#ifndef MYQC_PARSE_H
#define MYQC_PARSE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* In-memory form.  zmat: the text of a ZMAT file.  atoms[cap_nuc], xyz[3*cap_nuc] (row i = x,y,z of
 * nucleus i, bohr, centre-of-mass frame: what `nucpos` holds), options[17] (options(0:16)).
 * Call with cap_nuc = 0 (atoms, xyz NULL) to obtain nnuc first.  *problems is a bit mask of the
 * conditions for which the reference touches `error` but still writes its files:
 *   1 close contact (< 0.2 A), 2 charge/multiplicity mismatch, 4 negative electron count,
 *   8 RHF on an open shell, 16 contradictory options (check_options).                             */
int myqc_parse_zmat(const char *zmat, int cap_nuc, int *nnuc, int32_t *atoms, double *xyz,
                    int32_t *options, int *nelcA, int *nelcB, int *problems);

/* PROGRAM parser in directory `dir`: reads ZMAT, writes nucpos / envdat / fmem (list-directed
 * readable, 17 significant digits), prints the reference's messages and option table, touches
 * `error` where the reference does.                                                              */
int myqc_parse_main(const char *dir);

#ifdef __cplusplus
}
#endif
#endif /* MYQC_PARSE_H */
