// temporally blocked PDPS kernels, float, T = 3 (see tblock_kernels.h)
#define TB_REAL float
#define TB_T 3
#include "tblock_kernels.inc"
