// p24_api.cu — ABI version, error strings, per-device caches and the profiling aid of libp24_b200.
#include <cuda_runtime.h>

#include <mutex>

#include "p24_host.h"

extern "C" int p24_abi_version(void) { return P24_ABI_VERSION; }

extern "C" const char* p24_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case P24_E_BADARG: return "p24: bad argument (null pointer, non-positive size or misaligned workspace)";
        case P24_E_WORKSPACE: return "p24: workspace too small (see p24_workspace_bytes)";
        case P24_E_UNSUPPORTED: return "p24: unsupported configuration";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "p24: unknown error";
}

namespace p24 {

namespace {
std::mutex g_mu;
DevInfo g_dev[P24_MAX_DEVICES];
bool g_prof_on = false, g_prof_have = false;
cudaEvent_t g_ev[P24_PROF_MARKS];
bool g_ev_set[P24_PROF_MARKS];
}  // namespace

// cudaFuncSetAttribute and the SM count are per DEVICE: a process that drives several GPUs gets one entry each
DevInfo& dev_info() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= P24_MAX_DEVICES) dev = P24_MAX_DEVICES - 1;
    std::lock_guard<std::mutex> lk(g_mu);
    DevInfo& d = g_dev[dev];
    if (!d.n_sm) {
        cudaDeviceGetAttribute(&d.n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (d.n_sm <= 0) d.n_sm = 148;
    }
    return d;
}

bool dev_once(unsigned bit) {
    DevInfo& d = dev_info();
    std::lock_guard<std::mutex> lk(g_mu);
    if (d.attr_mask & bit) return false;
    d.attr_mask |= bit;
    return true;
}

bool prof_on() { return g_prof_on; }

void prof_mark(int i, cudaStream_t st) {
    if (!g_prof_on || i < 0 || i >= P24_PROF_MARKS) return;
    cudaEventRecord(g_ev[i], st);
    g_ev_set[i] = true;
}

}  // namespace p24

extern "C" int p24_profile_enable(int on) {
    using namespace p24;
    if (on && !g_prof_have) {
        for (int i = 0; i < P24_PROF_MARKS; ++i) {
            const cudaError_t e = cudaEventCreate(&g_ev[i]);
            if (e != cudaSuccess) return (int)e;
            g_ev_set[i] = false;
        }
        g_prof_have = true;
    }
    g_prof_on = on != 0;
    return 0;
}

extern "C" int p24_profile_read(float* h_ms8) {
    using namespace p24;
    if (!g_prof_have || !h_ms8) return P24_E_BADARG;
    for (int i = 0; i < P24_PROF_MARKS - 1; ++i) {
        h_ms8[i] = 0.0f;
        if (g_ev_set[i] && g_ev_set[i + 1]) {
            cudaError_t e = cudaEventSynchronize(g_ev[i + 1]);
            if (e != cudaSuccess) return (int)e;
            e = cudaEventElapsedTime(&h_ms8[i], g_ev[i], g_ev[i + 1]);
            if (e != cudaSuccess) return (int)e;
        }
    }
    for (int i = 0; i < P24_PROF_MARKS; ++i) g_ev_set[i] = false;
    return 0;
}
