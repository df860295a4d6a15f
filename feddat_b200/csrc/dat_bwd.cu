// placeholder until the backward kernels land (next commit)
#include "feddat_b200.h"
#include "host_common.h"
extern "C" int feddat_dat_bwd_dgrad(const void*, const void*, void*, const void*, const float*,
                                    const void*, const void*, void*, void*, int, int, int64_t, int,
                                    int, float, int, int, int, void*) {
  return fd::set_error(fd::FD_ERR_UNSUPPORTED, "dat_bwd_dgrad: not built yet");
}
extern "C" int feddat_dat_bwd_wgrad(const void*, const void*, const void*, const void*, float*,
                                    float*, float*, float*, int64_t, int, int, float, int, void*) {
  return fd::set_error(fd::FD_ERR_UNSUPPORTED, "dat_bwd_wgrad: not built yet");
}
