// Drop-in `ao2mo` executable: what the myQC driver spawns in the job directory after `scf`
// (src/myQC/myQC.f90:73; PROGRAM ao2mo, src/ao2mo/ao2mo.f90:22-98).
#include <cstdio>

#include "../../include/myqc_ao2mo.h"

int main() {
    myqc_ao2mo_main(".");  // failures touch `error`, which is all the driver looks at
    return 0;
}
