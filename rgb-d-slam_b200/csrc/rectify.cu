// R1 / R2 `rectify_depth` — the step in front of the CAPE path.
//
// Replaces  Depth_Map_Transformation::rectify_depth (depth_map_transformation.cpp:23-87): every valid pixel of the depth
//           camera's image is back-projected with the FLOAT tables _Xpre/_Ypre (init_matrices :147-173), moved by the
//           camera-2 -> camera-1 transformation, projected with camera 1's intrinsics and written to the pixel it falls in;
//           the serial row-major scan of the MAKE_DETERMINISTIC build makes the LAST source pixel in raster order win.
//
// The scatter has a defined winner, so it is order-free on the GPU: R1 does an atomicMax of the 64-bit key
// (source raster index + 1) << 32 | float bits of the new depth per destination pixel, R2 keeps the low word of each key
// (0 where nothing landed). Traffic per pixel: 4 B read + 8 B key RMW (+ 8 B memset) in R1, 8 B read + 4 B write in R2.
// Compiled with -fmad=false: the pixel a point falls in is a floor() of FP64 arithmetic that must round as the reference's.
#include "cape_internal.cuh"

namespace rs {

namespace {

__global__ void __launch_bounds__(256) rectify_scatter_kernel(const float4* __restrict__ depth, unsigned long long* __restrict__ keys,
                                                             const RectifyParams prm)
{
    const int W = prm.W, H = prm.H;
    const size_t perFrame4 = size_t(W) * H / 4;
    const size_t total4 = perFrame4 * prm.batch;
    for (size_t q = size_t(blockIdx.x) * blockDim.x + threadIdx.x; q < total4; q += size_t(gridDim.x) * blockDim.x) {
        const float4 v = __ldg(depth + q);
        const size_t b = q / perFrame4;
        const unsigned pix0 = unsigned(q - b * perFrame4) * 4u;   // raster index of the first of the four pixels
        const int row = int(pix0 / unsigned(W)), col0 = int(pix0 - unsigned(row) * unsigned(W));
        const float preY = static_cast<float>(prm.ky[row]);
        unsigned long long* frameKeys = keys + b * size_t(W) * H;
        const float zz[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float z = zz[t];
            if (!(z > 0.f)) continue;
            const float preX = static_cast<float>(prm.kx[col0 + t]);
            const double ox = static_cast<double>(preX * z), oy = static_cast<double>(preY * z), oz = static_cast<double>(z);
            const double px = ((prm.T[0] * ox + prm.T[1] * oy) + prm.T[2] * oz) + prm.T[3];
            const double py = ((prm.T[4] * ox + prm.T[5] * oy) + prm.T[6] * oz) + prm.T[7];
            const double pz = ((prm.T[8] * ox + prm.T[9] * oy) + prm.T[10] * oz) + prm.T[11];
            const double inv = 1.0 / pz;
            const double sx = inv * ((prm.fx * px + 0.0 * py) + prm.cx * pz);
            const double sy = inv * ((0.0 * px + prm.fy * py) + prm.cy * pz);
            if (sx != sx || sy != sy) continue;
            const double fxs = floor(sx), fys = floor(sy);
            if (!(fxs > 0.0 && fys > 0.0 && fxs < double(W) && fys < double(H))) continue;
            const unsigned dst = unsigned(int(fys)) * unsigned(W) + unsigned(int(fxs));
            const unsigned long long key =
                    (static_cast<unsigned long long>(pix0 + unsigned(t) + 1u) << 32) | __float_as_uint(static_cast<float>(pz));
            atomicMax(frameKeys + dst, key);
        }
    }
}

__global__ void __launch_bounds__(256) rectify_resolve_kernel(const ulonglong2* __restrict__ keys, float2* __restrict__ out,
                                                             const size_t n2)
{
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n2; i += size_t(gridDim.x) * blockDim.x) {
        const ulonglong2 k = keys[i];
        float2 o;
        o.x = __uint_as_float(static_cast<unsigned>(k.x));   // a key of 0 (nothing landed) gives +0.0f
        o.y = __uint_as_float(static_cast<unsigned>(k.y));
        out[i] = o;
    }
}

}  // namespace

int launch_rectify_depth(const RectifyParams& prm, const float* depth, unsigned long long* keys, float* out, cudaStream_t stream)
{
    const size_t px = size_t(prm.W) * prm.H * prm.batch;
    if ((size_t(prm.W) * prm.H) % 4 != 0) {
        set_last_error("rectify_depth: width * height must be a multiple of 4");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * px, stream));
    const int grid = 148 * 8;
    rectify_scatter_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(depth), keys, prm);
    RS_LAUNCH_CHECK();
    rectify_resolve_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const ulonglong2*>(keys), reinterpret_cast<float2*>(out), px / 2);
    RS_LAUNCH_CHECK();
    return RS_OK;
}

}  // namespace rs
