// Exact-math instantiation: compiled with -fmad=false; divisions and square roots as the reference performs them.
#define RLR_NS rlr_exact
#define RLR_EXACT 1
#include "rocket_kernels.cuh"
