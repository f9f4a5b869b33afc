// Shared host/device helpers of the B200 hot-path library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <mutex>
#include <string>

#include "../../include/rgbdslam_b200.h"

namespace rs {

// ---- host-side error plumbing ---------------------------------------------------------------
void set_last_error(const std::string& msg);
extern std::atomic<uint64_t> g_launch_count;

#define RS_CUDA_CHECK(expr)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            rs::set_last_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +      \
                               __FILE__ + ":" + std::to_string(__LINE__));                               \
            return RS_ERR_CUDA;                                                                          \
        }                                                                                                \
    } while (0)

#define RS_LAUNCH_CHECK()                                                                                \
    do {                                                                                                 \
        rs::g_launch_count.fetch_add(1, std::memory_order_relaxed);                                      \
        RS_CUDA_CHECK(cudaGetLastError());                                                               \
    } while (0)

// Raises a kernel's dynamic shared-memory limit to what a launch needs, per device and under a lock: the library is driven
// from several host threads (two batches in flight, concurrent contexts), and a limit that only ever grows cannot be lowered
// between another thread's check and its launch. (Not simply the 227 KB maximum once: measured on B200, a kernel whose limit
// is far above what it uses runs its L1-resident traffic - register spills, local arrays - 15 % slower; the driver appears to
// size the L1 / shared memory split by the limit.)
constexpr int kMaxDynamicSmem = 232448;
struct SmemOptIn {
    std::mutex m;
    size_t configured[64] = {};
    bool carveout_set[64] = {};
    template <class Kernel>
    cudaError_t ensure(Kernel kernel, const size_t bytes, const bool prefer_shared_carveout = false)
    {
        int dev = 0;
        cudaGetDevice(&dev);
        dev &= 63;
        std::lock_guard<std::mutex> lock(m);
        cudaError_t e = cudaSuccess;
        if (bytes > configured[dev]) {
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
            if (e == cudaSuccess) configured[dev] = bytes;
        }
        if (e == cudaSuccess && prefer_shared_carveout && !carveout_set[dev]) {
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e == cudaSuccess) carveout_set[dev] = true;
        }
        return e;
    }
};

// SM count of the current device (grids are sized in multiples of it); 148 on a B200, read once per device.
inline int sm_count()
{
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

// ---- constants of the reference (src/parameters.hpp) ---------------------------------------
// depth quantisation model, parameters.hpp:16-18 / covariances.cpp:12-19
__host__ __device__ inline double depth_quantization(const double depth)
{
    const double depthSigmaError = 2.73 * ((1.0 / 1000.0) * (1.0 / 1000.0));
    const double depthSigmaMultiplier = 0.74 / 1000.0;
    const double depthSigmaMargin = -0.53;
    const double q = depthSigmaMargin + depthSigmaMultiplier * depth + depthSigmaError * (depth * depth);
    return q > 0.5 ? q : 0.5;
}

constexpr double kPi = 3.14159265358979323846;

}  // namespace rs
