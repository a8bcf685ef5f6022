// Frame kernels: engines for 3..32 sequences, greedy; also the dispatching launcher.  See csm_stream.inl.
#define CSM_BUILD_SMALL 0
#define CSM_BUILD_STOCH 0
#include "csm_stream.inl"
