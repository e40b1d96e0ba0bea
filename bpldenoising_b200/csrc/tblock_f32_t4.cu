// temporally blocked PDPS kernels, float, T = 4 (see tblock_kernels.h)
#define TB_REAL float
#define TB_T 4
#include "tblock_kernels.inc"
