// Drop-in replacement for the reference's `int2e` executable (src/integrals/int2e.f90:14-69).
// The myQC driver runs it by name in the job directory (src/myQC/myQC.f90:54) and then tests for
// a file called `error` (myQC.f90:55-59); the exit status is not inspected by the reference but we
// return non-zero on failure anyway.
//   usage: int2e [ngpu]        (MYQC_NGPU in the environment also works; default 1)
#include <cstdio>
#include <cstdlib>

#include "../../include/myqc_eri.h"

int main(int argc, char** argv) {
    int ngpu = 1;
    if (const char* e = std::getenv("MYQC_NGPU")) ngpu = std::atoi(e);
    if (argc > 1) ngpu = std::atoi(argv[1]);
    const int rc = myqc_int2e_main(".", ngpu);
    return rc == MYQC_OK ? 0 : 1;
}
