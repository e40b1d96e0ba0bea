// temporally blocked PDPS kernels, double, T = 2 (see tblock_kernels.h)
#define TB_REAL double
#define TB_T 2
#include "tblock_kernels.inc"
