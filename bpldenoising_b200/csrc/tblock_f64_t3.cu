// temporally blocked PDPS kernels, double, T = 3 (see tblock_kernels.h)
#define TB_REAL double
#define TB_T 3
#include "tblock_kernels.inc"
