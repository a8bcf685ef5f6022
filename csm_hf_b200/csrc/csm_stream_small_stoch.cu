// Frame kernels: engines for <= 2 sequences, stochastic top-k sampling.  See csm_stream.inl.
#define CSM_BUILD_SMALL 1
#define CSM_BUILD_STOCH 1
#include "csm_stream.inl"
