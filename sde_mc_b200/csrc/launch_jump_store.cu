// launch_jump_store.cu -- dispatch of the path-storing jump-adapted launches to the translation unit of their jump source
#include "launch.cuh"

namespace sdemc {

int launch_jump_store_inject(const sdemc_sde& s, const LaunchArgs& a);
int launch_jump_store_queue(const sdemc_sde& s, const LaunchArgs& a);
int launch_jump_store_inline(const sdemc_sde& s, const LaunchArgs& a);

int launch_jump_store(const sdemc_sde& s, const LaunchArgs& a) {
  if (a.use_inject) return launch_jump_store_inject(s, a);
  return a.qdepth > 0 ? launch_jump_store_queue(s, a) : launch_jump_store_inline(s, a);
}

}  // namespace sdemc
