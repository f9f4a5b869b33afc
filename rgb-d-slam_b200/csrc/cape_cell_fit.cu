// K1 `cape_cell_fit` — per-cell plane fit straight from the depth image (no organized cloud in HBM).
//
// Replaces, fused:  Depth_Map_Transformation::get_organized_cloud_array  (depth_map_transformation.cpp:89-142)
//                   Plane_Segment::init_plane_segment + fit_plane         (plane_segment.cpp:44-168,205-284)
//                   Primitive_Detection::init_planar_cell_fitting          (primitive_detection.cpp:187-221)
//
// Two kernels. K1a `cape_cell_fit_kernel` streams the depth image on a persistent grid (one CTA per resident slot): the
// work unit is an "item" of 8 adjacent cells of one cell-row, items are dealt to the warps with a grid stride, FOUR LANES
// PER CELL: a 3-D tiled TMA box
// {cs px, 8 cells, R rows} (R*cs*8*4 = 2560 B) lands as [row][cell][px] in a 2-slot per-warp ring guarded by
// mbarriers; lane (c = lane/4, j = lane%4) reads the float4 j, j+4, j+8, ... of cell c's part of the box (LDS.128, at
// most 2-way bank conflicts; exactly 5 trips per box), back-projects in FP64 and accumulates the nine sums of FP32
// values / FP32 products in FP64, value for value what the reference's cloud + init_plane_segment compute. What bounds
// the loop on sm_100 is instruction issue (tools/microbench2.cu: an FP64 instruction holds the issue port ~2.2 cycles,
// F2F runs at a quarter of that rate on the XU pipe, everything else ~1 cycle), so the arithmetic is arranged for the
// fewest instructions per pixel: a float f >= 0 is widened without a conversion - its bit pattern shifted by 29 is the
// double f 2^-896 - and the 2^896 is folded, exactly, into the multiplier of the DFMA that consumes it (sums) or into
// pre-scaled back-projection factors (kxs, kys); products are FMUL on |x|, |y|, z with the sign restored on the widened
// operand (x) or on the multiplier (y: one sign per box row). Only x and y themselves cross the XU pipe (RN24 of the
// FP64 back-projection and back). 35 instructions per pixel, 11 of them FP64 (the first version of this kernel:
// 49 / 29, Veltkamp rounding inside the FP64 pipe, kept under RS_K1_FP64_PRODUCTS == 3). The pixel loop is branch-free:
// an invalid pixel (z <= 0) contributes exact zeros, as its (0,0,0) cloud row does. Per item the sums are reduced over
// the 4 lanes of a cell with two shuffle steps (all four lanes end up with the totals and each writes a quarter of the
// record); the cross-shaped continuity test runs from small per-warp copies of the middle row / middle column, split
// over the 4 lanes of the cell.
// K1b `cape_cell_finish_kernel` then fits every cell (3x3 eigen-solve, planarity, merge tolerance), one thread per cell,
// in place on the 160-byte record: inside the streaming kernel that serial chain cost 30 % of the time.
//
// Compiled with -fmad=false: products are FP32-rounded then accumulated in FP64 exactly as the reference does.
#include <cuda.h>

#include <algorithm>

#include "cape_internal.cuh"
#include "plane_fit.cuh"
#include "tma.cuh"

namespace rs {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CELLS_PER_ITEM = 8;
constexpr int WARPS = 4;
#ifndef RS_K1_MIN_CTAS
#define RS_K1_MIN_CTAS 4
#endif
#ifndef RS_K1_UNROLL
#define RS_K1_UNROLL 5          // trips of the per-box float4 loop unrolled
#endif
#ifndef RS_K1B_MIN_CTAS
#define RS_K1B_MIN_CTAS 6       // K1b: 80 registers, 24 warps / SM (measured best of 4 / 6 / 8)
#endif
#ifndef RS_K1_F2F_PRODUCTS
#define RS_K1_F2F_PRODUCTS 0    // how many of the signed products (xz, xy, yz) are widened by F2F instead of the integer route
#endif
#ifndef RS_K1_FP64_PRODUCTS
#define RS_K1_FP64_PRODUCTS 8   // 8: integer widening folded into DFMA (default); 3: Veltkamp rounding in the FP64 pipe (round-1 original)
#endif

constexpr int kUnroll = RS_K1_UNROLL;

// Continuity scan of one middle row / column, e_i = base[i], i in [0, n). Reference (plane_segment.cpp:44-100):
// last = max(e_0, e_1); fail if last <= 0; for i = 1..n-1 a positive e_i must satisfy |e_i - last| <= 4 quant(e_i) and
// then becomes `last`. Because any failure rejects the cell, `last` at step i is the nearest positive element in
// [1, i-1] (or the initial max when there is none). The four lanes of a cell take SEG consecutive elements each
// ([lo, hi), hi - lo <= SEG): a lane first finds the last positive element of its own segment, the predecessor of its
// first element then comes from the earlier lanes by shuffle, and the scan itself is branch-free but for the rare jump
// that lands within 1e-5 of the bound. Must be called by all four lanes of the cell (`lead` = its first lane).
template <int SEG>
__device__ __forceinline__ bool continuity_lane(const float* base, const int lo, const int hi, const float init, const int j,
                                                const int lead)
{
    float z[SEG];
    float last = 0.f;
#pragma unroll
    for (int t = 0; t < SEG; ++t) {
        z[t] = lo + t < hi ? base[lo + t] : 0.f;      // past the end reads as an empty pixel
        last = z[t] > 0.f ? z[t] : last;
    }
    const float l0 = __shfl_sync(FULL, last, lead), l1 = __shfl_sync(FULL, last, lead + 1), l2 = __shfl_sync(FULL, last, lead + 2);
    float prev = init;
    if (j > 0 && l0 > 0.f) prev = l0;
    if (j > 1 && l1 > 0.f) prev = l1;
    if (j > 2 && l2 > 0.f) prev = l2;
    bool ok = true;
#pragma unroll
    for (int t = 0; t < SEG; ++t) {
        const float zt = z[t];
        const bool valid = zt > 0.f;
        const float diff = fabsf(zt - prev);
        // jump <= 4 quant(z), decided in FP32 whenever the FP32 estimate of the bound is further from `diff` than its
        // own error (a few 1e-7 relative, plus the cancellation against the -0.53 offset); the reference's FP64
        // expression only runs for the rare jump that lands within 1e-5 of the bound
        const float qf = fmaxf(0.5f, fmaf(2.73e-6f * zt, zt, fmaf(0.74e-3f, zt, -0.53f)));
        const float bound = 4.0f * qf, slack = fmaf(bound, 1e-5f, 1e-5f);
        if (valid && diff > bound + slack) ok = false;
        if (valid && diff >= bound - slack && diff <= bound + slack) {
            if (!(static_cast<double>(diff) <= 4.0 * depth_quantization(static_cast<double>(zt)))) ok = false;
        }
        prev = valid ? zt : prev;
    }
    return ok;
}

// v rounded to the nearest float (ties to even), as a double: Veltkamp's split, p = v (2^29 + 1), hi = p - (p - v).
// Exact for every v whose rounded value is a normal float or zero.
__device__ __forceinline__ double round_to_float(const double v)
{
    const double p = v * 536870913.0;
    return p - (p - v);
}

// f * 2^-896 for a float f >= 0 (zero and subnormals included): the bits of f times 2^29, read as a double.
__device__ __forceinline__ double widen_scaled(const float f)
{
    // two shifts; the single-instruction form (IMAD.WIDE.U32 by 2^29) measured 5 % slower over the whole kernel (dispatch stalls)
    const unsigned b = __float_as_uint(f);
    return __hiloint2double(static_cast<int>(b >> 3), static_cast<int>(b << 29));
}
constexpr double kTwo896 = 0x1p896;
// d >= 0 with the sign bit of the 32-bit pattern b
__device__ __forceinline__ double with_sign_of(const double d, const unsigned b)
{
    return __hiloint2double(__double2hiint(d) | static_cast<int>(b & 0x80000000u), __double2loint(d));
}

template <int CS>
struct Geometry {
    static constexpr int G = CS / 4;                                 // float4 groups per cell row
    static constexpr int BOX_BYTES = 2560;                           // one TMA box: R rows of 8 cells
    static constexpr int R = BOX_BYTES / (CS * CELLS_PER_ITEM * 4);  // rows per TMA box (4 @20 px, 2 @40 px)
    static constexpr int NBOX = CS / R;                              // boxes per item
    static constexpr int F4_PER_CELL_BOX = R * G;                    // float4 of one cell in one box (20)
    static constexpr int ITERS = F4_PER_CELL_BOX / 4;                // per-lane trips per box (5, no padding trip)
    static constexpr int ROW_FLOATS = CS * CELLS_PER_ITEM;           // one box row in floats
    static constexpr int MID_BYTES = 2 * CELLS_PER_ITEM * CS * 4;    // middle row + middle column copies
    static constexpr int WARP_BYTES = (2 * BOX_BYTES + MID_BYTES + 16 + 127) / 128 * 128;  // TMA destinations stay 128-B aligned
    static_assert(CS % 4 == 0 && CS % R == 0 && R * CS * CELLS_PER_ITEM * 4 == BOX_BYTES && F4_PER_CELL_BOX % 4 == 0,
                  "unsupported cell size");
};

template <int CS>
__global__ void __launch_bounds__(WARPS * 32, RS_K1_MIN_CTAS)
        cape_cell_fit_kernel(const __grid_constant__ CUtensorMap tmap, const CellFitParams prm, rs_cell_out* __restrict__ cells)
{
    using Geo = Geometry<CS>;
    constexpr int P = CS * CS;
    constexpr int G = Geo::G, R = Geo::R, NBOX = Geo::NBOX;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem_raw + size_t(warp) * Geo::WARP_BYTES;
    float* midrow = reinterpret_cast<float*>(wbase + 2 * Geo::BOX_BYTES);   // [8][CS]
    float* midcol = midrow + CELLS_PER_ITEM * CS;                            // [8][CS] (CS-1 used)
    uint64_t* bars = reinterpret_cast<uint64_t*>(wbase + 2 * Geo::BOX_BYTES + Geo::MID_BYTES);

    // Work unit = one item (8 adjacent cells of a cell row). The grid is sized to the resident warp slots and every warp
    // walks the item list with a grid stride: ~10 items per warp at 256 frames, so the last round is 94 % full instead of
    // the 59 % that one 32-cell strip per warp left (2.59 waves). The 2-slot TMA ring runs straight through item boundaries.
    const int gwarp = blockIdx.x * WARPS + warp, nwarps = gridDim.x * WARPS;
    if (gwarp >= prm.total_items) return;
    const int nitems = (prm.total_items - gwarp + nwarps - 1) / nwarps;     // items of this warp
    const int nbox = nitems * NBOX;
    const int ipr = prm.items_per_strip;                                    // items per cell row
    // item -> (first image row in the [B*H] dimension, first cell column)
    auto item_origin = [&](const int k, int& row0, int& c0, int& cr, int& b) {
        const int item = gwarp + k * nwarps;
        b = item / (prm.vc * ipr);
        const int rem = item - b * prm.vc * ipr;
        cr = rem / ipr;
        c0 = (rem - cr * ipr) * CELLS_PER_ITEM;
        row0 = b * prm.H + cr * CS;
    };

    const uint32_t bar_a = smem_u32(bars), tile_a = smem_u32(wbase);   // shared-window addresses of the ring, converted once
    const uint64_t policy = l2_policy_evict_first();   // the depth image is read once; keep L2 for the records
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
        fence_proxy_async();
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (q < nbox) {
                int row0, c0, cr, b;
                item_origin(q / NBOX, row0, c0, cr, b);
                mbar_arrive_expect_tx_a(bar_a + q * 8, Geo::BOX_BYTES);
                tma_load_3d_hint_a(tile_a + q * Geo::BOX_BYTES, &tmap, 0, c0, row0 + (q % NBOX) * R, bar_a + q * 8, policy);
            }
    }
    __syncwarp();

    const int c = lane >> 2, j = lane & 3;             // cell inside the item, lane inside the cell
#if RS_K1_FP64_PRODUCTS == 8
    const double* const kxtab = prm.kxs;               // factors pre-scaled by 2^896 (see the pixel loop)
    const double* const kytab = prm.kys;
#else
    const double* const kxtab = prm.kx;
    const double* const kytab = prm.ky;
#endif

    double S0 = 0, S1 = 0, S2 = 0, S3 = 0, S4 = 0, S5 = 0, S6 = 0, S7 = 0, S8 = 0;
    int cnt = 0;
    float p0x = 0.f, p0y = 0.f, p0z = 0.f, plx = 0.f, ply = 0.f, plz = 0.f;

    // (item, box) of the box being consumed and of the box two ahead (the one the ring is refilled with). The item
    // coordinates cost two integer divisions: they are computed once per item, by the refill cursor, and the consume
    // cursor inherits them when it crosses into that item (NBOX >= 3: the refill cursor is then inside the same item).
    static_assert(NBOX >= 3, "the consume cursor takes its item coordinates from the refill cursor");
    int k = 0, bq = 0, row0, c0, cr, b;
    item_origin(0, row0, c0, cr, b);
    int kn = 0, bqn = 2, nrow0 = row0, nc0 = c0, ncr = cr, nb = b;
    for (int q = 0; q < nbox; ++q) {
        const int slot = q & 1;
        const double* kyrow = prm.ky + cr * CS;
        const float* tile = reinterpret_cast<const float*>(wbase + slot * Geo::BOX_BYTES);
        const float* ctile = tile + c * CS;            // this cell's columns inside a box row
        // first image column of this lane's cell. When the cell grid is not a multiple of the 8-cell item, the cells of the last
        // item that hang over the image (TMA zero-fills their pixels, their records are never written) must not index the
        // factor tables past their end: they borrow the last real cell's columns.
        const int colbase = min(c0 + c, prm.hc - 1) * CS;
        mbar_wait_a(bar_a + slot * 8, (q >> 1) & 1);

        // ---- branch-free accumulation of this lane's float4s of the box ----
        int r = 0, g = j;                              // flat index j + 4k -> (row r, group g); G >= 5 > 4
        if (g >= G) {
            g -= G;
            ++r;
        }
#pragma unroll kUnroll
        for (int it = 0; it < Geo::ITERS; ++it) {
            {
                const float4 v = *reinterpret_cast<const float4*>(ctile + r * Geo::ROW_FLOATS + g * 4);
                const double2 kxa = __ldg(reinterpret_cast<const double2*>(kxtab + colbase + g * 4));
                const double2 kxb = __ldg(reinterpret_cast<const double2*>(kxtab + colbase + g * 4 + 2));
                const double kyv = __ldg(kytab + cr * CS + bq * R + r);
                const float zz[4] = {v.x, v.y, v.z, v.w};
#if RS_K1_FP64_PRODUCTS == 8
                const double cy = __hiloint2double((__double2hiint(kyv) & 0x80000000) | 0x77f00000, 0);   // copysign(2^896, ky)
#endif
                const double kxr[4] = {kxa.x, kxa.y, kxb.x, kxb.y};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    // z = zz > 0 ? zz : 0, cnt += zz > 0: one FSETP, one FSEL and one predicated add (the compiler's own
                    // rendering of the count is an add plus a predicated move)
                    float z;
                    asm("{\n.reg .pred p;\nsetp.gt.f32 p, %2, 0f00000000;\nselp.f32 %1, %2, 0f00000000, p;\n@p add.s32 %0, %0, 1;\n}"
                        : "+r"(cnt), "=f"(z)
                        : "f"(zz[t]));
                    // The reference's values are float(z kx), float(z ky) and FP32 products widened to FP64. On sm_100 the
                    // FP32<->FP64 conversions run at a quarter of the FP64 rate (tools/microbench.cu) and eleven of them per
                    // pixel bounded this loop, so the rounding to 24 bits is done in FP64 where that is cheaper: Veltkamp's
                    // split with 2^29 + 1 returns exactly RN_24(v), ties to even included (tests/test_oracle_cape.py), and
                    // the product of two 24-bit values is exact in FP64 before it is rounded.
#if RS_K1_FP64_PRODUCTS == 8
                    // Exact widening without the XU pipe: for a float f >= 0 with bits b, the 64-bit integer b * 2^29 read as
                    // a double is f * 2^-896 (exponent field not rebiased; zero and subnormals included), one IMAD.WIDE.
                    // The 2^896 is folded, exactly, into the multiplier of the operation that consumes it: the
                    // back-projection factors are stored pre-scaled (kxs = kx 2^896) and the sums use S = fma(D, 2^896, S),
                    // bit-identical to S += (double)f. Signed products are formed from |x|, |y|: the sign of x is OR-ed
                    // into D, the sign of y (that of ky: one value per box row) rides on the multiplier cy = +-2^896.
                    // Only x and y cross the XU pipe (RN24 of the FP64 back-projection, and back).
                    const double Dz = widen_scaled(z);
                    const float x = static_cast<float>(Dz * kxr[t]);
                    const float y = static_cast<float>(Dz * kyv);
                    S0 += static_cast<double>(x);
                    S1 += static_cast<double>(y);
                    S2 = fma(Dz, kTwo896, S2);
                    const float ax = fabsf(x), ay = fabsf(y);
                    const unsigned bx = __float_as_uint(x);
                    S3 = fma(widen_scaled(__fmul_rn(ax, ax)), kTwo896, S3);
                    S4 = fma(widen_scaled(__fmul_rn(ay, ay)), kTwo896, S4);
                    S5 = fma(widen_scaled(__fmul_rn(z, z)), kTwo896, S5);
#if RS_K1_F2F_PRODUCTS >= 2
                    S6 += static_cast<double>(__fmul_rn(x, y));
#else
                    S6 = fma(with_sign_of(widen_scaled(__fmul_rn(ax, ay)), bx), cy, S6);
#endif
#if RS_K1_F2F_PRODUCTS >= 3
                    S7 += static_cast<double>(__fmul_rn(y, z));
#else
                    S7 = fma(widen_scaled(__fmul_rn(ay, z)), cy, S7);
#endif
#if RS_K1_F2F_PRODUCTS >= 1
                    S8 += static_cast<double>(__fmul_rn(x, z));
#else
                    S8 = fma(with_sign_of(widen_scaled(__fmul_rn(ax, z)), bx), kTwo896, S8);
#endif
#else
                    // round-1 original: x, y and three of the six products are rounded to 24 bits inside the FP64 pipe
                    // (Veltkamp split, bit-identical to the conversion), the other three go through FMUL + F2F
                    const double zd = static_cast<double>(z);
                    const double xd = round_to_float(zd * kxr[t]);
                    const double yd = round_to_float(zd * kyv);
                    const float x = static_cast<float>(xd);   // exact: xd already has a 24-bit significand
                    const float y = static_cast<float>(yd);
                    S0 += xd;
                    S1 += yd;
                    S2 += zd;
                    S3 += round_to_float(xd * xd);
                    S4 += round_to_float(yd * yd);
                    S5 += static_cast<double>(z * z);
                    S6 += round_to_float(xd * yd);
                    S7 += static_cast<double>(y * z);
                    S8 += static_cast<double>(x * z);
#endif
                }
            }
            g += 4;
            if (g >= G) {
                g -= G;
                ++r;
            }
        }
        // ---- first / last cloud row of the cell (merge tolerance, primitive_detection.cpp:201-220) ----
        if (j == 0 && bq == 0) {
            const float z = ctile[0];
            if (z > 0.f) {
                const double zd = static_cast<double>(z);
                p0x = static_cast<float>(zd * __ldg(prm.kx + colbase));
                p0y = static_cast<float>(zd * __ldg(kyrow));
                p0z = z;
            }
        }
        if (j == 0 && bq == NBOX - 1) {
            const float z = ctile[(R - 1) * Geo::ROW_FLOATS + CS - 1];
            if (z > 0.f) {
                const double zd = static_cast<double>(z);
                plx = static_cast<float>(zd * __ldg(prm.kx + colbase + CS - 1));
                ply = static_cast<float>(zd * __ldg(kyrow + CS - 1));
                plz = z;
            }
        }
        // ---- keep the middle row / middle column for the continuity test ----
        {
            constexpr int MIDBOX = (CS / 2) / R, MIDROW = (CS / 2) % R;
            if (bq == MIDBOX)
                for (int t = j; t < CS; t += 4) midrow[c * CS + t] = ctile[MIDROW * Geo::ROW_FLOATS + t];
            for (int t = j; t < R; t += 4) midcol[c * CS + bq * R + t] = ctile[t * Geo::ROW_FLOATS + CS / 2];
        }
        __syncwarp();
        // the box is consumed: refill the slot with box q + 2
        if (lane == 0 && q + 2 < nbox) {
            mbar_arrive_expect_tx_a(bar_a + slot * 8, Geo::BOX_BYTES);
            tma_load_3d_hint_a(tile_a + slot * Geo::BOX_BYTES, &tmap, 0, nc0, nrow0 + bqn * R, bar_a + slot * 8, policy);
        }
        // advance the refill cursor (box q + 3 next time)
        if (++bqn == NBOX) {
            bqn = 0;
            ++kn;
            if (kn < nitems) item_origin(kn, nrow0, nc0, ncr, nb);
        }
        const bool item_done = bq == NBOX - 1;
        const int cur_c0 = c0, cur_cr = cr, cur_b = b;
        // advance the consume cursor
        if (++bq == NBOX) {
            bq = 0;
            ++k;
            row0 = nrow0, c0 = nc0, cr = ncr, b = nb;   // kn == k here (stale, and unused, after the last item)
        }
        if (!item_done) continue;

        // ---- item finished: reduce the 4 lanes of each cell (all four end up with the totals), continuity test ----
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            cnt += __shfl_xor_sync(FULL, cnt, o);
            S0 += __shfl_xor_sync(FULL, S0, o);
            S1 += __shfl_xor_sync(FULL, S1, o);
            S2 += __shfl_xor_sync(FULL, S2, o);
            S3 += __shfl_xor_sync(FULL, S3, o);
            S4 += __shfl_xor_sync(FULL, S4, o);
            S5 += __shfl_xor_sync(FULL, S5, o);
            S6 += __shfl_xor_sync(FULL, S6, o);
            S7 += __shfl_xor_sync(FULL, S7, o);
            S8 += __shfl_xor_sync(FULL, S8, o);
        }
        bool cont;
        {
            // horizontal: CS elements of the middle row; vertical: CS-1 elements of the middle column
            const float* hr = midrow + c * CS;
            const float* vcq = midcol + c * CS;
            const float hinit = fmaxf(hr[0], hr[1]);
            const float vinit = fmaxf(vcq[0], vcq[1]);
            constexpr int HSEG = (CS - 1 + 3) / 4, VSEG = (CS - 2 + 3) / 4;
            const int hlo = 1 + j * HSEG, hhi = min(hlo + HSEG, CS);
            const int vlo = 1 + j * VSEG, vhi = min(vlo + VSEG, CS - 1);
            const bool hok = continuity_lane<HSEG>(hr, hlo, hhi, hinit, j, lane & ~3);
            const bool vok = continuity_lane<VSEG>(vcq, vlo, vhi, vinit, j, lane & ~3);
            cont = hinit > 0.f && vinit > 0.f && hok && vok;
        }
        const unsigned bad = __ballot_sync(FULL, !cont);
        const int okc = (((bad >> (lane & ~3)) & 0xfu) == 0u && cnt >= P / 2) ? 1 : 0;
        // the first / last cloud rows live on the cell's lane 0: hand them to its three neighbours
        const int lead = lane & ~3;
        const float a0 = __shfl_sync(FULL, p0x, lead), a1 = __shfl_sync(FULL, p0y, lead), a2 = __shfl_sync(FULL, p0z, lead);
        const float b0 = __shfl_sync(FULL, plx, lead), b1 = __shfl_sync(FULL, ply, lead), b2 = __shfl_sync(FULL, plz, lead);

        // ---- hand the cell over to the fit kernel: sums, count, continuity flag and the first / last cloud rows travel
        // in the cell's own record (the two points in the centroid / normal slots); the four lanes of a cell write two
        // 16-byte pieces each ----
        if (cur_c0 + c < prm.hc) {
            double2* dst = reinterpret_cast<double2*>(cells + (size_t(cur_b) * prm.vc + cur_cr) * prm.hc + cur_c0 + c);
            if (j == 0) {
                double2 head;
                head.x = __hiloint2double(okc, cnt);   // {int32 count, int32 planar := continuity / count test passed}
                head.y = S0;
                dst[0] = head;
                dst[1] = make_double2(S1, S2);
            }
            else if (j == 1) {
                dst[2] = make_double2(S3, S4);
                dst[3] = make_double2(S5, S6);
            }
            else if (j == 2) {
                dst[4] = make_double2(S7, S8);
                dst[5] = make_double2(static_cast<double>(a0), static_cast<double>(a1));
            }
            else {
                dst[6] = make_double2(static_cast<double>(a2), static_cast<double>(b0));
                dst[7] = make_double2(static_cast<double>(b1), static_cast<double>(b2));
            }
        }
        S0 = S1 = S2 = S3 = S4 = S5 = S6 = S7 = S8 = 0.0;
        cnt = 0;
        p0x = p0y = p0z = plx = ply = plz = 0.f;
        __syncwarp();                                   // midrow / midcol are rewritten by the next item
    }
}

// K1b: the per-cell fit (plane_segment.cpp:102-168 after the sums, primitive_detection.cpp:201-220), one THREAD per cell.
// The 3x3 eigen-solve is a serial chain of FP64 divisions and square roots with a data-dependent trip count; inside the
// streaming kernel it ran on warps that were holding TMA buffers and cost 30 % of K1's time for 4 % of its instructions.
// Here every cell of the batch is in flight at once and the records (31 MB per 256 frames) are still in L2.
__global__ void __launch_bounds__(128, RS_K1B_MIN_CTAS) cape_cell_finish_kernel(const CellFitParams prm, rs_cell_out* __restrict__ cells, const int total)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    double2* rec = reinterpret_cast<double2*>(cells + i);
    const double2 head = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3], r4 = rec[4], r5 = rec[5], r6 = rec[6], r7 = rec[7];
    const int fcount = __double2loint(head.x), fok = __double2hiint(head.x);
    PlaneModel pm;
    plane_clear(pm);
    float tol = 0.f;
    if (fok) {
        pm.count = fcount;
        pm.S[0] = head.y, pm.S[1] = r1.x, pm.S[2] = r1.y, pm.S[3] = r2.x, pm.S[4] = r2.y, pm.S[5] = r3.x, pm.S[6] = r3.y;
        pm.S[7] = r4.x, pm.S[8] = r4.y;
        if (fcount >= prm.min_zero_point_count) {
            plane_fit(pm);
            const double qz = depth_quantization(pm.c[2]);
            pm.planar = (pm.mse <= qz * qz) ? 1 : 0;
        }
        if (pm.planar) {
            const float dx = static_cast<float>(r6.y) - static_cast<float>(r5.x);
            const float dy = static_cast<float>(r7.x) - static_cast<float>(r5.y);
            const float dz = static_cast<float>(r7.y) - static_cast<float>(r6.x);
            const float diameter = sqrtf((dx * dx + dy * dy) + dz * dz);
            tol = fminf(prm.merge_distance, diameter * prm.sin_merge * sqrtf(static_cast<float>(pm.count)));
        }
    }
    double2 out0;
    out0.x = __hiloint2double(pm.planar, pm.count);    // {int32 count, int32 planar} little-endian
    out0.y = pm.S[0];
    rec[0] = out0;
    rec[1] = make_double2(pm.S[1], pm.S[2]);
    rec[2] = make_double2(pm.S[3], pm.S[4]);
    rec[3] = make_double2(pm.S[5], pm.S[6]);
    rec[4] = make_double2(pm.S[7], pm.S[8]);
    rec[5] = make_double2(pm.c[0], pm.c[1]);
    rec[6] = make_double2(pm.c[2], pm.n[0]);
    rec[7] = make_double2(pm.n[1], pm.n[2]);
    rec[8] = make_double2(pm.d, pm.mse);
    // bin of init_histogram (primitive_detection.cpp:239-265, histogram.hpp:35-62): computed here, where every cell of the
    // batch has its own thread, instead of serially per frame in the segmentation kernel
    int bin = -1;
    if (pm.planar) {
        const int cs = prm.cell;
        const double theta = acos(-pm.n[2]);
        const double phi = atan2(pm.n[0], pm.n[1]);
        const int xQ = static_cast<int>(floor((cs - 1) * (theta - 0.0) / (kPi - 0.0)));
        int yQ = 0;
        if (xQ > 0) yQ = static_cast<int>(floor((cs - 1) * (phi - (-kPi)) / (kPi - (-kPi))));
        bin = yQ * cs + xQ;
    }
    double2 tail;
    tail.x = pm.score;
    tail.y = __hiloint2double(bin, __float_as_int(tol));  // {float tol, int32 hist_bin}
    rec[9] = tail;
}

template <int CS>
int launch_variant(const CUtensorMap& tmap, const CellFitParams& prm, rs_cell_out* cells, cudaStream_t stream, cudaEvent_t streamed)
{
    using Geo = Geometry<CS>;
    constexpr size_t smem = size_t(WARPS) * Geo::WARP_BYTES;
    auto kernel = cape_cell_fit_kernel<CS>;
    static SmemOptIn optin;
    RS_CUDA_CHECK(optin.ensure(kernel, smem, true));
    CellFitParams p = prm;
    p.items_per_strip = (prm.hc + CELLS_PER_ITEM - 1) / CELLS_PER_ITEM;   // items (8 cells) per cell row
    p.total_items = prm.batch * prm.vc * p.items_per_strip;
    // persistent grid: one CTA per resident slot (SMs x RS_K1_MIN_CTAS), fewer when the batch is small
    const int slots = sm_count() * RS_K1_MIN_CTAS;
    const int grid = std::min(slots, (p.total_items + WARPS - 1) / WARPS);
    kernel<<<grid, WARPS * 32, smem, stream>>>(tmap, p, cells);
    RS_LAUNCH_CHECK();
    if (streamed) RS_CUDA_CHECK(cudaEventRecord(streamed, stream));   // the HBM-bound part is over: other streams may start
    const int total = prm.batch * prm.vc * prm.hc;
    cape_cell_finish_kernel<<<(total + 127) / 128, 128, 0, stream>>>(p, cells, total);
    RS_LAUNCH_CHECK();
    return RS_OK;
}

}  // namespace

// rows per TMA box for a cell size (the C-ABI glue encodes the tensor map with box {cell, 8, rows})
int cape_cell_fit_box_rows(int cell) { return cell == 20 ? Geometry<20>::R : (cell == 40 ? Geometry<40>::R : 0); }
int cape_cell_fit_box_cells() { return CELLS_PER_ITEM; }

int launch_cape_cell_fit(const CUtensorMap& tmap, const CellFitParams& prm, rs_cell_out* cells, cudaStream_t stream,
                         cudaEvent_t streamed)
{
    if (prm.cell == 20) return launch_variant<20>(tmap, prm, cells, stream, streamed);
    if (prm.cell == 40) return launch_variant<40>(tmap, prm, cells, stream, streamed);
    set_last_error("cape_cell_fit: unsupported cell size (built for 20 and 40 px)");
    return RS_ERR_INVALID_ARG;
}

}  // namespace rs
