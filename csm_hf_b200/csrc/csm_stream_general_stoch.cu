// Frame kernels: engines for 3..32 sequences, stochastic top-k sampling.  See csm_stream.inl.
#define CSM_BUILD_SMALL 0
#define CSM_BUILD_STOCH 1
#include "csm_stream.inl"
