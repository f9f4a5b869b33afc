// K1 `cape_cell_fit` — per-cell plane fit straight from the depth image (no organized cloud in HBM).
//
// Replaces, fused:  Depth_Map_Transformation::get_organized_cloud_array  (depth_map_transformation.cpp:89-142)
//                   Plane_Segment::init_plane_segment + fit_plane         (plane_segment.cpp:44-168,205-284)
//                   Primitive_Detection::init_planar_cell_fitting          (primitive_detection.cpp:187-221)
//
// Mapping: one warp owns a run of CPI consecutive cells of one cell-row ("item"). For every cell the warp's lane 0
// TMA-loads the cs x cs depth tile (a 3-D tiled tensor map over [row][cell-col][px] makes it land as a dense,
// bank-conflict-free cs*cs float tile) into a STAGES-deep per-warp ring guarded by mbarriers; lanes read float4
// columns, back-project in FP64 (cast to float, as the reference's cloud), and accumulate the nine sums of FP32
// values/products in FP64. The cross-shaped continuity test runs lane-parallel from the same tile. After the CPI
// cells are reduced, the 3x3 eigen-solves are done lane-parallel (one cell per lane) — a warp-wide solve per cell
// would spend ~5x the accumulation time on 1/32-utilised FP64 issue slots. Records (160 B) are staged in shared
// memory and written with 16-byte coalesced stores.
//
// Compiled with -fmad=false: products are FP32-rounded then accumulated in FP64 exactly as the reference does.
#include <cuda.h>

#include "cape_internal.cuh"
#include "plane_fit.cuh"
#include "tma.cuh"

namespace rs {

namespace {

constexpr int REC_STRIDE = 176;  // bytes per staged record (160 used), multiple of 16

// per-cell accumulator staged in shared memory between the accumulate and the fit phase (<= REC_STRIDE bytes)
struct CellAcc {
    double S[9];
    int count;
    int ok;          // continuity tests passed and enough positive pixels
    float p0[3];     // cloud row 0 of the cell (x,y,z) — zeros when the pixel is invalid
    float pl[3];     // cloud row P-1
};
static_assert(sizeof(CellAcc) <= REC_STRIDE, "accumulator must fit the record slot");

// Lane-parallel restatement of is_cell_{horizontal,vertical}_continuous (plane_segment.cpp:44-100).
// Elements e_i = base[i*stride], i = 0..n-1. Reference: last = max(e_0, e_1); fail if last <= 0; for i = 1..n-1:
// a positive e_i must satisfy |e_i - last| <= 4*quant(e_i) and then becomes `last`. Because any failure rejects the
// cell, `last` at step i is the nearest positive element in [1, i-1] (or the initial max when there is none).
__device__ __forceinline__ bool continuity_scan(const float* base, const int stride, const int n, const int lane)
{
    const float e0 = base[0], e1 = base[stride];
    float carry = fmaxf(e0, e1);
    if (carry <= 0.f) return false;
    bool ok = true;
    for (int first = 1; first < n; first += 32) {
        const int i = first + lane;
        const float z = (i < n) ? base[i * stride] : 0.f;
        const bool valid = z > 0.f;
        const unsigned mask = __ballot_sync(0xffffffffu, valid);
        const unsigned lower = mask & ((1u << lane) - 1u);
        const int src = lower ? (31 - __clz(lower)) : 0;
        const float pz = __shfl_sync(0xffffffffu, z, src);
        const float prev = lower ? pz : carry;
        if (valid) {
            const float diff = fabsf(z - prev);
            if (!(static_cast<double>(diff) <= 4.0 * depth_quantization(static_cast<double>(z)))) ok = false;
        }
        const int hi = mask ? (31 - __clz(mask)) : 0;
        const float cz = __shfl_sync(0xffffffffu, z, hi);
        if (mask) carry = cz;
    }
    return __all_sync(0xffffffffu, ok);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int CS, int CPI, int STAGES, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (CS <= 20 ? 2 : 1)) cape_cell_fit_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                    const CellFitParams prm, rs_cell_out* __restrict__ cells)
{
    constexpr int P = CS * CS;
    constexpr int G = CS / 4;              // float4 groups per tile row
    constexpr int RPI = 32 / G;            // tile rows covered per iteration
    constexpr int ACTIVE = RPI * G;        // lanes that own a float4 column
    constexpr int ITERS = (CS + RPI - 1) / RPI;
    constexpr int TILE_BYTES = P * 4;
    constexpr int SLOT_BYTES = (TILE_BYTES + 127) / 128 * 128;
    static_assert(CS % 4 == 0 && G <= 32, "cell side must be a multiple of 4 and <= 128");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem_raw + size_t(warp) * (size_t(STAGES) * SLOT_BYTES + size_t(CPI) * REC_STRIDE + 128);
    float* ring = reinterpret_cast<float*>(wbase);
    unsigned char* recs = wbase + size_t(STAGES) * SLOT_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(recs + size_t(CPI) * REC_STRIDE);

    const int item = blockIdx.x * WARPS + warp;
    if (item >= prm.total_items) return;
    const int ips = prm.items_per_strip;
    const int b = item / (prm.vc * ips);
    const int rem = item - b * prm.vc * ips;
    const int cr = rem / ips;
    const int c0 = (rem - cr * ips) * CPI;
    const int ncell = min(CPI, prm.hc - c0);
    const int row0 = b * prm.H + cr * CS;  // first image row of the strip in the [B*H] row dimension

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
        fence_proxy_async();
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
            if (s < ncell) {
                mbar_arrive_expect_tx(&bars[s], TILE_BYTES);
                tma_load_3d(reinterpret_cast<unsigned char*>(ring) + size_t(s) * SLOT_BYTES, &tmap, 0, c0 + s, row0, &bars[s]);
            }
    }
    __syncwarp();

    // lane -> (float4 column group g, first tile row r0)
    const bool active = lane < ACTIVE;
    const int g = lane % G, r0 = lane / G;
    double kyr[ITERS];
#pragma unroll
    for (int k = 0; k < ITERS; ++k) {
        const int r = r0 + k * RPI;
        kyr[k] = (active && r < CS) ? __ldg(prm.ky + cr * CS + r) : 0.0;
    }
    constexpr int LAST_LANE = ((CS - 1) % RPI) * G + (G - 1);
    constexpr int LAST_K = (CS - 1) / RPI;

    for (int i = 0; i < ncell; ++i) {
        const int slot = i % STAGES;
        const float* tile = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(ring) + size_t(slot) * SLOT_BYTES);
        mbar_wait(&bars[slot], (i / STAGES) & 1);

        const int col0 = (c0 + i) * CS + g * 4;
        double kxr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) kxr[j] = active ? __ldg(prm.kx + col0 + j) : 0.0;

        double S0 = 0, S1 = 0, S2 = 0, S3 = 0, S4 = 0, S5 = 0, S6 = 0, S7 = 0, S8 = 0;
        int cnt = 0;
        float fx0 = 0.f, fy0 = 0.f, fz0 = 0.f, fxl = 0.f, fyl = 0.f, fzl = 0.f;
#pragma unroll
        for (int k = 0; k < ITERS; ++k) {
            const int r = r0 + k * RPI;
            if (active && r < CS) {
                const float4 v = *reinterpret_cast<const float4*>(tile + r * CS + g * 4);
                const float zz[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float z = zz[j];
                    if (z > 0.f) {
                        ++cnt;
                        const double zd = static_cast<double>(z);
                        const float x = static_cast<float>(zd * kxr[j]);
                        const float y = static_cast<float>(zd * kyr[k]);
                        S0 += static_cast<double>(x);
                        S1 += static_cast<double>(y);
                        S2 += zd;
                        S3 += static_cast<double>(x * x);
                        S4 += static_cast<double>(y * y);
                        S5 += static_cast<double>(z * z);
                        S6 += static_cast<double>(x * y);
                        S7 += static_cast<double>(y * z);
                        S8 += static_cast<double>(x * z);
                        if (k == 0 && j == 0 && lane == 0) fx0 = x, fy0 = y, fz0 = z;
                        if (k == LAST_K && j == 3 && lane == LAST_LANE) fxl = x, fyl = y, fzl = z;
                    }
                }
            }
        }
        // cross-shaped continuity test from the tile: middle row, then middle column without its last row
        const bool hcont = continuity_scan(tile + (CS / 2) * CS, 1, CS, lane);
        const bool vcont = continuity_scan(tile + CS / 2, CS, CS - 1, lane);
        __syncwarp();
        // the tile is consumed: refill the slot with cell i + STAGES
        if (lane == 0 && i + STAGES < ncell) {
            mbar_arrive_expect_tx(&bars[slot], TILE_BYTES);
            tma_load_3d(const_cast<float*>(tile), &tmap, 0, c0 + i + STAGES, row0, &bars[slot]);
        }

        const int total = __reduce_add_sync(0xffffffffu, cnt);
        S0 = warp_sum(S0);
        S1 = warp_sum(S1);
        S2 = warp_sum(S2);
        S3 = warp_sum(S3);
        S4 = warp_sum(S4);
        S5 = warp_sum(S5);
        S6 = warp_sum(S6);
        S7 = warp_sum(S7);
        S8 = warp_sum(S8);
        CellAcc* acc = reinterpret_cast<CellAcc*>(recs + size_t(i) * REC_STRIDE);
        if (lane == 0) {
            acc->S[0] = S0, acc->S[1] = S1, acc->S[2] = S2, acc->S[3] = S3, acc->S[4] = S4;
            acc->S[5] = S5, acc->S[6] = S6, acc->S[7] = S7, acc->S[8] = S8;
            acc->count = total;
            acc->ok = (hcont && vcont && total >= P / 2) ? 1 : 0;
            acc->p0[0] = fx0, acc->p0[1] = fy0, acc->p0[2] = fz0;
        }
        if (lane == LAST_LANE) acc->pl[0] = fxl, acc->pl[1] = fyl, acc->pl[2] = fzl;
    }
    __syncwarp();

    // ---- fit phase: one cell per lane -----------------------------------------------------------
    for (int base = 0; base < ncell; base += 32) {
        const int ci = base + lane;
        rs_cell_out rec;
        if (ci < ncell) {
            const CellAcc a = *reinterpret_cast<const CellAcc*>(recs + size_t(ci) * REC_STRIDE);
            PlaneModel pm;
            plane_clear(pm);
            float tol = 0.f;
            if (a.ok) {
                pm.count = a.count;
#pragma unroll
                for (int s = 0; s < 9; ++s) pm.S[s] = a.S[s];
                if (a.count >= prm.min_zero_point_count) {
                    plane_fit(pm);
                    const double q = depth_quantization(pm.c[2]);
                    pm.planar = (pm.mse <= q * q) ? 1 : 0;
                }
                if (pm.planar) {
                    const float dx = a.pl[0] - a.p0[0], dy = a.pl[1] - a.p0[1], dz = a.pl[2] - a.p0[2];
                    const float diameter = sqrtf((dx * dx + dy * dy) + dz * dz);
                    tol = fminf(prm.merge_distance, diameter * prm.sin_merge * sqrtf(static_cast<float>(pm.count)));
                }
            }
            rec.count = pm.count;
            rec.planar = pm.planar;
#pragma unroll
            for (int s = 0; s < 9; ++s) rec.S[s] = pm.S[s];
#pragma unroll
            for (int s = 0; s < 3; ++s) rec.centroid[s] = pm.c[s], rec.normal[s] = pm.n[s];
            rec.d = pm.d;
            rec.mse = pm.mse;
            rec.score = pm.score;
            rec.tol = tol;
            rec.reserved = 0;
        }
        __syncwarp();
        if (ci < ncell) *reinterpret_cast<rs_cell_out*>(recs + size_t(ci) * REC_STRIDE) = rec;
        __syncwarp();
    }

    // ---- coalesced 16-byte stores of the ncell * 160 B record run -------------------------------------
    uint4* dst = reinterpret_cast<uint4*>(cells + (size_t(b) * prm.vc + cr) * prm.hc + c0);
    for (int t = lane; t < ncell * 10; t += 32) {
        const int r = t / 10, part = t - r * 10;
        dst[t] = *reinterpret_cast<const uint4*>(recs + size_t(r) * REC_STRIDE + part * 16);
    }
}

template <int CS, int CPI, int STAGES, int WARPS>
int launch_variant(const CUtensorMap& tmap, const CellFitParams& prm, rs_cell_out* cells, cudaStream_t stream)
{
    constexpr int SLOT_BYTES = (CS * CS * 4 + 127) / 128 * 128;
    constexpr size_t smem = size_t(WARPS) * (size_t(STAGES) * SLOT_BYTES + size_t(CPI) * REC_STRIDE + 128);
    auto kernel = cape_cell_fit_kernel<CS, CPI, STAGES, WARPS>;
    static bool configured = false;
    if (!configured) {
        RS_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = true;
    }
    CellFitParams p = prm;
    p.items_per_strip = (prm.hc + CPI - 1) / CPI;
    p.total_items = prm.batch * prm.vc * p.items_per_strip;
    const int grid = (p.total_items + WARPS - 1) / WARPS;
    kernel<<<grid, WARPS * 32, smem, stream>>>(tmap, p, cells);
    RS_LAUNCH_CHECK();
    return RS_OK;
}

}  // namespace

int launch_cape_cell_fit(const CUtensorMap& tmap, const CellFitParams& prm, rs_cell_out* cells, cudaStream_t stream)
{
    // cells per warp-item: long runs amortise the lane-parallel eigen-solves; short runs keep small batches
    // spread over all 148 SMs.
    const long cellsTotal = long(prm.batch) * prm.vc * prm.hc;
    const bool big = cellsTotal >= 148L * 16 * 32;  // >= one 32-cell item per resident warp
    if (prm.cell == 20) {
        if (big) return launch_variant<20, 32, 4, 8>(tmap, prm, cells, stream);
        return launch_variant<20, 8, 4, 8>(tmap, prm, cells, stream);
    }
    if (prm.cell == 40) {
        if (big) return launch_variant<40, 32, 3, 8>(tmap, prm, cells, stream);
        return launch_variant<40, 8, 3, 8>(tmap, prm, cells, stream);
    }
    set_last_error("cape_cell_fit: unsupported cell size (built for 20 and 40 px)");
    return RS_ERR_INVALID_ARG;
}

}  // namespace rs
