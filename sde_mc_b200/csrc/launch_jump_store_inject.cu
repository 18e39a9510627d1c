// launch_jump_store_inject.cu -- the path-storing jump-adapted kernels with the injected (parity mode) jump source
#define SDEMC_STORE_JSRC JSRC_INJECT
#define SDEMC_STORE_ENTRY launch_jump_store_inject
#include "launch_jump_store.inc"
