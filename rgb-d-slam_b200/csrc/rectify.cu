// R1 / R2 `rectify_depth` — the step in front of the CAPE path.
//
// Replaces  Depth_Map_Transformation::rectify_depth (depth_map_transformation.cpp:23-87): every valid pixel of the depth
//           camera's image is back-projected with the FLOAT tables _Xpre/_Ypre (init_matrices :147-173), moved by the
//           camera-2 -> camera-1 transformation, projected with camera 1's intrinsics and written to the pixel it falls in;
//           the serial row-major scan of the MAKE_DETERMINISTIC build makes the LAST source pixel in raster order win.
//
// The scatter has a defined winner, so it is order-free on the GPU. The winner of a destination pixel is found IN the output
// image itself: R1 does a 32-bit atomicMax of a key ordered like the source raster index, (row << 16 | column) + 1, per
// destination word, R2 replaces every winning key by the depth that source pixel projects to - the same device function
// evaluates that depth wherever it is needed, so the value R2 writes is bit for bit the one the scatter would have carried.
// Round 1 carried it in a 64-bit key in a scratch image twice the size of the batch (8 B per pixel of HBM; memset 8 B + key
// read-modify-write 16 B + key read 8 B + result 4 B of traffic per pixel); now: zero fill 4 B, depth 4 B, a 4-byte reduction
// in L2, key read 4 B + gathered depth 4 B + result 4 B, and no scratch.
// Compiled with -fmad=false: the pixel a point falls in is a floor() of FP64 arithmetic that must round as the reference's.
#include <algorithm>

#include "cape_internal.cuh"

namespace rs {

namespace {

// (double)f: the plain conversion. A shift-and-rebias widening on the integer pipe (as K1a uses) was tried here: it takes six
// instructions and a range branch against one F2F, and once the grid was balanced both kernels turned out to be bound by
// instruction issue, not by the conversion pipe - 0.582 ms per 256 frames with the integer widening, 0.477 with F2F
// (0.450 without the two NaN compares the rounding-down conversion makes redundant).
__device__ __forceinline__ double widen(const float f) { return static_cast<double>(f); }

// RI = the rotation block of the camera-2 -> camera-1 transformation is exactly the identity (the reference's default,
// parameters.cpp:59-74, and every pure translation): ((1 ox + 0 oy) + 0 oz) + t is ox + t bit for bit - 1 ox is exact, the
// zero products are zeros of some sign, and a zero only changes a sum that is itself zero, where it can flip the sign of a
// zero that the later fx px + cx pz absorbs (or that ends in column / row 0, which the range test drops either way). 15 of
// the 35 FP64 instructions per pixel go away. The host picks the instantiation by comparing the nine coefficients.
template <bool RI>
__device__ __forceinline__ double transform_row(const double* T, const int r, const double ox, const double oy, const double oz)
{
    if (RI) return (r == 0 ? ox : (r == 1 ? oy : oz)) + T[4 * r + 3];
    return ((T[4 * r] * ox + T[4 * r + 1] * oy) + T[4 * r + 2] * oz) + T[4 * r + 3];
}

// The depth a source pixel carries to camera 1 (depth_map_transformation.cpp:52-58,73-75): z of T * (preX z, preY z, z, 1)
template <bool RI>
__device__ __forceinline__ double rectify_depth_of(const RectifyParams& prm, const double ox, const double oy, const double oz)
{
    return transform_row<RI>(prm.T, 2, ox, oy, oz);
}

// One source pixel through depth_map_transformation.cpp:48-77: the destination pixel it falls in. false: it is dropped.
// to_screen_coordinates multiplies by the whole intrinsics matrix, (fx px + 0 py) + cx pz: a finite py leaves fx px as it is
// (a +-0 is added), a non-finite one makes sy NaN or out of range as well, so the zero terms are left out; floor + the
// unsigned cast + the range test are one conversion rounding down (it saturates, NaN gives 0: both fail 0 < i < size).
template <bool RI>
__device__ __forceinline__ bool rectify_project(const RectifyParams& prm, const float preX, const float preY, const float z,
                                                unsigned& dst)
{
    const double ox = widen(preX * z), oy = widen(preY * z), oz = widen(z);
    const double px = transform_row<RI>(prm.T, 0, ox, oy, oz);
    const double py = transform_row<RI>(prm.T, 1, ox, oy, oz);
    const double pz = rectify_depth_of<RI>(prm, ox, oy, oz);
    const double inv = 1.0 / pz;
    const double sx = inv * (prm.fx * px + prm.cx * pz);
    const double sy = inv * (prm.fy * py + prm.cy * pz);
    // no separate NaN test: a NaN converts to 0 and fails the range test (the oracle drops such a pixel; the reference itself
    // leaves the process through exit(-1) there, depth_map_transformation.cpp:79-83)
    const int ix = __double2int_rd(sx), iy = __double2int_rd(sy);
    if (!(ix > 0 && iy > 0 && ix < prm.W && iy < prm.H)) return false;
    dst = unsigned(iy) * unsigned(prm.W) + unsigned(ix);
    return true;
}

// Winner keys order source pixels by raster position and unpack without a division: (row << 16 | column) + 1.
// R1: winners[dst] = max key over the source pixels landing on dst. `winners` is the output image, zeroed. One block row of the
// grid per frame of the group (blockIdx.y), whole image rows per block (no per-pixel index arithmetic).
template <bool RI>
__global__ void __launch_bounds__(512) rectify_scatter_kernel(const float* __restrict__ depth, unsigned* __restrict__ winners,
                                                             const RectifyParams prm)
{
    const int W = prm.W, W4 = W >> 2;
    const size_t frame_px = size_t(W) * prm.H;
    const float* src = depth + blockIdx.y * frame_px;
    unsigned* frame = winners + blockIdx.y * frame_px;
    const int per_block = blockDim.x / W4 > 0 ? blockDim.x / W4 : 1;      // image rows a block covers per trip
    const int lrow = threadIdx.x / W4, c0 = threadIdx.x - lrow * W4;      // once per thread
    const int stride = blockDim.x < W4 ? blockDim.x : W4;                 // threads along a row
    for (int row = blockIdx.x * per_block + lrow; row < prm.H && lrow < per_block; row += gridDim.x * per_block) {
        const float preY = __ldg(prm.preY + row);
        const float* line = src + size_t(row) * W;
        // a warp instruction covers 32 CONSECUTIVE pixels: their destinations are (mostly) consecutive words, so the L2 sees
        // a few sectors per reduction instead of one per pixel (a float4 per lane measured 1.0 sector per pixel)
        for (int col = c0; col < W; col += 4 * stride) {
            float zz[4], pre[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int cc = col + t * stride;
                zz[t] = cc < W ? line[cc] : 0.f;
                pre[t] = cc < W ? __ldg(prm.preX + cc) : 0.f;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float z = zz[t];
                if (!(z > 0.f)) continue;
                unsigned dst;
                if (rectify_project<RI>(prm, pre[t], preY, z, dst))
                    atomicMax(frame + dst, (unsigned(row) << 16 | unsigned(col + t * stride)) + 1u);
            }
        }
    }
}

// R2, in place on the output image: winning key -> the depth that source pixel projects to; 0 (nothing landed) -> +0.0f
template <bool RI>
__global__ void __launch_bounds__(256) rectify_resolve_kernel(const float* __restrict__ depth, uint4* __restrict__ image,
                                                             const RectifyParams prm)
{
    const int W = prm.W;
    const size_t frame_px = size_t(W) * prm.H;
    const size_t perFrame4 = frame_px / 4;
    const float* frame = depth + blockIdx.y * frame_px;
    uint4* img = image + blockIdx.y * perFrame4;
    for (size_t q = size_t(blockIdx.x) * blockDim.x + threadIdx.x; q < perFrame4; q += size_t(gridDim.x) * blockDim.x) {
        const uint4 k = img[q];
        if ((k.x | k.y | k.z | k.w) == 0u) continue;   // the zero fill is already the answer
        unsigned kk[4] = {k.x, k.y, k.z, k.w};
        // the four gathers first (independent loads in flight together), then the arithmetic
        float z[4], pxf[4], pyf[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const unsigned key = kk[t] - 1u, row = key >> 16, col = key & 0xffffu;
            const bool hit = kk[t] != 0u;
            z[t] = hit ? frame[size_t(row) * W + col] : 0.f;
            pxf[t] = hit ? __ldg(prm.preX + col) : 0.f;
            pyf[t] = hit ? __ldg(prm.preY + row) : 0.f;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (kk[t] == 0u) continue;
            const double pz = rectify_depth_of<RI>(prm, widen(pxf[t] * z[t]), widen(pyf[t] * z[t]), widen(z[t]));
            kk[t] = __float_as_uint(static_cast<float>(pz));
        }
        img[q] = make_uint4(kk[0], kk[1], kk[2], kk[3]);
    }
}

}  // namespace

int launch_rectify_depth(const RectifyParams& prm, const float* depth, float* out, cudaStream_t stream)
{
    const size_t frame_px = size_t(prm.W) * prm.H;
    if (prm.W % 4 != 0 || prm.W >= 65536 || prm.H >= 65535) {
        set_last_error("rectify_depth: the width must be a multiple of 4, and width and height below 65535");
        return RS_ERR_INVALID_ARG;
    }
    if (prm.batch > 65535) {   // one grid row per frame
        set_last_error("rectify_depth: at most 65535 frames per call");
        return RS_ERR_INVALID_ARG;
    }
    // One pass over the whole batch: walking it in groups of frames that fit the L2 (so that DRAM would see the depth once and
    // the result once) measured slower at every group size - 0.79 ms per 256 frames at 48 MB groups against 0.72 - the
    // kernels are bound by the FP64 pipe (R1) and by the latency of the gather (R2), not by DRAM: profiles/README.md
    RS_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float) * frame_px * prm.batch, stream));
    // 64 blocks per SM in all (a few image rows per block): rows with many invalid pixels finish early, and with 8 blocks per
    // SM the scatter kernel ran at 52 % achieved occupancy. Measured 0.709 / 0.676 / 0.643 / 0.585 / 0.585 ms per 256 frames at
    // 8 / 16 / 32 / 64 / 256 blocks per SM.
    const dim3 grid((sm_count() * 64 + prm.batch - 1) / prm.batch, prm.batch);
    const dim3 grid2 = grid;
    // whole image rows per block: a multiple of the row's float4 count when that fits a block
    const int w4 = prm.W / 4, threads = w4 <= 512 ? w4 * std::max(1, 256 / w4) : 256;
    const double* T = prm.T;
    const bool ri = T[0] == 1.0 && T[1] == 0.0 && T[2] == 0.0 && T[4] == 0.0 && T[5] == 1.0 && T[6] == 0.0 && T[8] == 0.0 && T[9] == 0.0 &&
                    T[10] == 1.0;
    if (ri) {
        rectify_scatter_kernel<true><<<grid, threads, 0, stream>>>(depth, reinterpret_cast<unsigned*>(out), prm);
        RS_LAUNCH_CHECK();
        rectify_resolve_kernel<true><<<grid2, 256, 0, stream>>>(depth, reinterpret_cast<uint4*>(out), prm);
    }
    else {
        rectify_scatter_kernel<false><<<grid, threads, 0, stream>>>(depth, reinterpret_cast<unsigned*>(out), prm);
        RS_LAUNCH_CHECK();
        rectify_resolve_kernel<false><<<grid2, 256, 0, stream>>>(depth, reinterpret_cast<uint4*>(out), prm);
    }
    RS_LAUNCH_CHECK();
    return RS_OK;
}

}  // namespace rs
