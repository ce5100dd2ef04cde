// Drop-in `int1e` executable: what the myQC driver spawns in the job directory
// (src/myQC/myQC.f90:46; PROGRAM int1e, src/integrals/int1e.f90:14-131).
#include <cstdio>

#include "../../include/myqc_int1e.h"

int main() {
    myqc_int1e_main(".");  // failures touch `error`, which is all the driver looks at (myQC.f90:47-51)
    return 0;
}
