// gradient_nd.cu — translation unit of the nested-dissection adjoint solver (kernels: nd_solver.cuh, nd_tv.cuh;
// host driver: gradient_nd.cuh).
#include "gradient_nd.cuh"
