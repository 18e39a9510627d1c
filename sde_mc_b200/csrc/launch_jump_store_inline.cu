// launch_jump_store_inline.cu -- the path-storing jump-adapted kernels with the inline (dense jumps) jump source
#define SDEMC_STORE_JSRC JSRC_INLINE
#define SDEMC_STORE_ENTRY launch_jump_store_inline
#include "launch_jump_store.inc"
