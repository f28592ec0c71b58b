// p24_host.h — host-side helpers shared by the translation units of libp24_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include "../../include/p24.h"

#define P24_MAX_DEVICES 64
// profiling marks: 0..3 bracket the three kernels of the training chain (k_prep, k_pass, k_tail), 4..6 the two of the
// postprocess (k_post_filter, k_post_nms); p24_profile_read returns mark[i+1] - mark[i] for i = 0..7
#define P24_PROF_MARKS 9

namespace p24 {
struct DevInfo {
    int n_sm;
    unsigned attr_mask;  // which kernels' attributes were set on this device
};
DevInfo& dev_info();          // of the current device
bool dev_once(unsigned bit);  // true the first time `bit` is asked for on the current device
bool prof_on();
void prof_mark(int i, cudaStream_t st);
}  // namespace p24
