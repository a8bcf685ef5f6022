// general kernel family (engines for 3..32 sequences), greedy sampling
#define CSM_BUILD_STOCH 0
#include "csm_batch.inl"
