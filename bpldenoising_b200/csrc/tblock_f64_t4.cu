// temporally blocked PDPS kernels, double, T = 4 (see tblock_kernels.h)
#define TB_REAL double
#define TB_T 4
#include "tblock_kernels.inc"
