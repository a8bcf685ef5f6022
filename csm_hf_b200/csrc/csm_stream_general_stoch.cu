// general kernel family (engines for 3..32 sequences), stochastic top-k sampling
#define CSM_BUILD_STOCH 1
#include "csm_batch.inl"
