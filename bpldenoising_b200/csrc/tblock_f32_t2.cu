// temporally blocked PDPS kernels, float, T = 2 (see tblock_kernels.h)
#define TB_REAL float
#define TB_T 2
#include "tblock_kernels.inc"
