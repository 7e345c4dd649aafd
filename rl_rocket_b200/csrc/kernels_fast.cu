// Fast-math instantiation: FMA contraction on (nvcc default), reciprocal-sqrt quaternion normalisation.
#define RLR_NS rlr_fast
#define RLR_EXACT 0
#include "rocket_kernels.cuh"
