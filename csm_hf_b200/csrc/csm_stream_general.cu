// Frame kernels of engines built for 3..32 sequences, with the hang guard, the debug hooks and the poll
// back-off, plus the launcher that dispatches between the two families.  See csm_stream.inl.
#define CSM_BUILD_SMALL 0
#include "csm_stream.inl"
