// Drop-in `parse` executable: the first stage the myQC driver spawns in the job directory
// (src/myQC/myQC.f90:29; PROGRAM parser, src/parser/parser.f90:22-108).
#include "../../include/myqc_parse.h"

int main() {
    myqc_parse_main(".");  // failures touch `error`, which is all the driver looks at (myQC.f90:30-34)
    return 0;
}
