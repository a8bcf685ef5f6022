// Frame kernels: engines for <= 2 sequences, greedy.  See csm_stream.inl.
#define CSM_BUILD_SMALL 1
#define CSM_BUILD_STOCH 0
#include "csm_stream.inl"
