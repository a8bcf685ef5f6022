// Frame kernels of engines built for <= 2 sequences (SMALL): the leanest hot path -- fused decoder attention,
// no hang guard, no debug hooks, no poll back-off.  See csm_stream.inl.
#define CSM_BUILD_SMALL 1
#include "csm_stream.inl"
