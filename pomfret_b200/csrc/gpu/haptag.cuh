// (c) Read haplotagging kernel for untagged BAMs (-u); see engine_haptag.inc for the host side.
#ifndef POMFRET_GPU_HAPTAG_CUH
#define POMFRET_GPU_HAPTAG_CUH
#include "gpu_rt.h"
#include "types.h"
namespace pomfret_gpu {
}  // namespace pomfret_gpu
#endif
