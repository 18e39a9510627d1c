// launch_jump_store_queue.cu -- the path-storing jump-adapted kernels with the queued (sparse jumps) jump source
#define SDEMC_STORE_JSRC JSRC_QUEUE
#define SDEMC_STORE_ENTRY launch_jump_store_queue
#include "launch_jump_store.inc"
