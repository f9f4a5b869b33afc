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

// Raises a kernel's dynamic shared-memory limit to the sm_100 opt-in maximum (227 KB) exactly once per device, under a
// lock: the library is driven from several host threads (two batches in flight, concurrent contexts), and a limit that
// only ever holds the maximum cannot be lowered between another thread's check and its launch. The limit is a ceiling,
// not a reservation: occupancy follows the bytes each launch actually asks for.
constexpr int kMaxDynamicSmem = 232448;
struct SmemOptIn {
    std::mutex m;
    bool done[64] = {};
    template <class Kernel>
    cudaError_t ensure(Kernel kernel, const bool prefer_shared_carveout = false)
    {
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lock(m);
        if (done[dev & 63]) return cudaSuccess;
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynamicSmem);
        if (e == cudaSuccess && prefer_shared_carveout)
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e == cudaSuccess) done[dev & 63] = true;
        return e;
    }
};

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
