// C-ABI glue for the CAPE path (include/rgbdslam_b200.h): context, device buffers, TMA descriptor, launches.
// There is no CPU fallback anywhere in this library: without an sm_100 device every entry point fails loudly.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <random>
#include <string>
#include <vector>

#include "cape_internal.cuh"

namespace rs {

static std::mutex g_err_mutex;
static std::string g_last_error;
std::atomic<uint64_t> g_launch_count{0};

void set_last_error(const std::string& msg)
{
    std::lock_guard<std::mutex> lock(g_err_mutex);
    g_last_error = msg;
}

int require_blackwell(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        set_last_error(std::string("no CUDA device visible (") + cudaGetErrorString(e) +
                       "): this library has no CPU path");
        return RS_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) {
        set_last_error("device index out of range");
        return RS_ERR_INVALID_ARG;
    }
    cudaDeviceProp prop;
    RS_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_last_error(std::string("device '") + prop.name + "' is not sm_100 (Blackwell): kernels are built for sm_100a only");
        return RS_ERR_NO_DEVICE;
    }
    RS_CUDA_CHECK(cudaSetDevice(device));
    return RS_OK;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn()
{
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
    return fn;
}

}  // namespace rs

using namespace rs;

struct rs_cape_ctx {
    int W, H, cell, hc, vc, Nc, max_batch, device, max_boundary;
    double fx, fy, cx, cy;
    double* d_kx = nullptr;
    double* d_ky = nullptr;
    float* d_pre = nullptr;   // rectify_depth: the factor tables as floats, [W] then [H]
    float* d_depth = nullptr;
    rs_cape_outputs d_out{};       // device pointers sized for max_batch
    // Cylinder-RANSAC draw tables (canonical doubles of mt19937(seed)), one slot per recently used seed. A table is copied
    // stream-ordered from the slot's own pinned staging buffer, never rewritten while a launch that reads it may be in
    // flight (a seed change takes another slot; the slot it evicts is first waited for), and every launch that reads a
    // table records an event on the slot.
    static constexpr int kUniformSlots = 4, kUniformUses = 8;
    struct UniformSlot {
        double* d = nullptr;          // device table
        double* h = nullptr;          // pinned staging
        uint32_t seed = 0;
        bool valid = false;
        uint64_t stamp = 0;           // LRU
        cudaEvent_t copied = nullptr; // the H2D copy of the table has landed
        cudaEvent_t used[kUniformUses] = {};
        int next_use = 0;
    } uniforms[kUniformSlots];
    uint64_t uniform_clock = 0;
    std::mutex uniform_mutex;
    double* d_scratch = nullptr;   // per-frame scratch of the segmentation kernel
    uint16_t* d_depth16 = nullptr; // staging for rs_cape_run_u16 (allocated on first use)
    bool rectify = false;          // rs_cape_set_rectification: rectify_depth in front of K1
    RectifyParams rect{};
    float* d_rect = nullptr;               // max_batch x H x W rectified depth
    int n_uniforms = 0;
    CellFitParams fit{};
    SegmentParams seg{};
    // cached tensor map
    const float* tmap_ptr = nullptr;
    int tmap_batch = 0;
    CUtensorMap tmap;
    cudaStream_t stream = nullptr;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;   // copy legs of the chunk pipeline of rs_cape_run
    cudaStream_t seg_streams[4] = {nullptr, nullptr, nullptr, nullptr};  // segmentation of chunk k runs on [k % 4]
    std::vector<cudaEvent_t> chunk_events;                     // 2 per chunk: depth landed, results ready
    cudaEvent_t fit_done = nullptr;   // recorded after every K1 launch (K1a + K1b): the segmentation of a chunk waits on it
    cudaEvent_t streamed = nullptr;   // recorded after K1a, the HBM-bound streaming kernel (rs_cape_stream_wait_fit)
    std::vector<cudaEvent_t> events;  // 4 per timing slot: fit start, fit end, segmentation start, segmentation end
    int timing_slots = 0;
    uint64_t run_counter = 0;
};

namespace {

// rs_cape_run moves a host batch through the GPU in chunks of this many frames: the H2D copy of chunk k+1, the kernels
// of chunk k and the D2H copy of chunk k-1 overlap (three streams), so a batch costs its PCIe time, not the sum.
constexpr int kChunkFrames = 32;

// cv::Mat::convertTo(CV_32F, alpha) of a CV_16U depth image (examples/main_TUM.cpp:242, main_CAPE.cpp:59): OpenCV's
// cvtScale 16u -> 32f works in float: dst = float(src) * float(alpha). Eight pixels per thread (16-byte load, two
// 16-byte stores); n8 = pixel count / 8.
__global__ void depth_u16_to_f32_kernel(const uint4* __restrict__ src, float4* __restrict__ dst, const size_t n8, const float alpha)
{
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += size_t(gridDim.x) * blockDim.x) {
        const uint4 v = __ldg(src + i);
        float4 a, b;
        a.x = __fmul_rn(float(v.x & 0xffffu), alpha), a.y = __fmul_rn(float(v.x >> 16), alpha);
        a.z = __fmul_rn(float(v.y & 0xffffu), alpha), a.w = __fmul_rn(float(v.y >> 16), alpha);
        b.x = __fmul_rn(float(v.z & 0xffffu), alpha), b.y = __fmul_rn(float(v.z >> 16), alpha);
        b.z = __fmul_rn(float(v.w & 0xffffu), alpha), b.w = __fmul_rn(float(v.w >> 16), alpha);
        dst[2 * i] = a;
        dst[2 * i + 1] = b;
    }
}

// Eigen's 3x3 cofactor inverse of the intrinsics applied to (u, v, 1): point_coordinates.cpp:79-83.
void backprojection_factors(const rs_cape_ctx* c, std::vector<double>& kx, std::vector<double>& ky)
{
    const double K[3][3] = {{c->fx, 0, c->cx}, {0, c->fy, c->cy}, {0, 0, 1}};
    auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return K[i1][j1] * K[i2][j2] - K[i1][j2] * K[i2][j1];
    };
    const double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    const double det = (c00 * K[0][0] + c10 * K[1][0]) + c20 * K[2][0];
    const double invdet = 1.0 / det;
    const double i00 = c00 * invdet, i01 = c10 * invdet, i02 = c20 * invdet;
    const double i10 = cof(0, 1) * invdet, i11 = cof(1, 1) * invdet, i12 = cof(2, 1) * invdet;
    kx.resize(c->W);
    ky.resize(c->H);
    for (int u = 0; u < c->W; ++u) kx[u] = (i00 * double(u) + i01 * 0.0) + i02 * 1.0;
    for (int v = 0; v < c->H; ++v) ky[v] = (i10 * 0.0 + i11 * double(v)) + i12 * 1.0;
}

int encode_tmap(rs_cape_ctx* c, const float* depth_dev, int batch)
{
    if (c->tmap_ptr == depth_dev && c->tmap_batch == batch) return RS_OK;
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
        set_last_error("cuTensorMapEncodeTiled entry point not available from the driver");
        return RS_ERR_CUDA;
    }
    // tensor view of the depth batch: dim0 = pixel column inside a cell, dim1 = cell column, dim2 = image row
    // (frames are contiguous, so batch*H rows). A {cs, 8, R} box is R rows of 8 adjacent cells, landing dense in shared
    // memory as [row][cell][px]; cells beyond the last cell column are zero-filled by the TMA unit.
    const int boxRows = cape_cell_fit_box_rows(c->cell);
    if (boxRows <= 0) {
        set_last_error("rs_cape: unsupported cell size (the plane-fit kernel is built for 20 and 40 px cells)");
        return RS_ERR_INVALID_ARG;
    }
    const cuuint64_t dims[3] = {cuuint64_t(c->cell), cuuint64_t(c->hc), cuuint64_t(batch) * cuuint64_t(c->H)};
    const cuuint64_t strides[2] = {cuuint64_t(c->cell) * 4, cuuint64_t(c->W) * 4};
    const cuuint32_t box[3] = {cuuint32_t(c->cell), cuuint32_t(cape_cell_fit_box_cells()), cuuint32_t(boxRows)};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&c->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(depth_dev), dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
        return RS_ERR_CUDA;
    }
    c->tmap_ptr = depth_dev;
    c->tmap_batch = batch;
    return RS_OK;
}

// canonical doubles of std::mt19937(seed) through std::uniform_real_distribution<double>(0,1), i.e. what
// utils::Random::get_random_double() returns on a fresh thread (random.hpp:17-31). Returns the slot holding the table of
// `seed`, ordered before anything enqueued on `stream` after this call; the caller records uniforms_used() after the
// launch that reads it.
int ensure_uniforms(rs_cape_ctx* c, uint32_t seed, cudaStream_t stream, int* slot_out)
{
    std::lock_guard<std::mutex> lock(c->uniform_mutex);
    int pick = -1;
    for (int k = 0; k < rs_cape_ctx::kUniformSlots; ++k)
        if (c->uniforms[k].valid && c->uniforms[k].seed == seed) pick = k;
    if (pick >= 0) {
        rs_cape_ctx::UniformSlot& u = c->uniforms[pick];
        u.stamp = ++c->uniform_clock;
        RS_CUDA_CHECK(cudaStreamWaitEvent(stream, u.copied, 0));   // the copy may have been enqueued on another stream
        *slot_out = pick;
        return RS_OK;
    }
    for (int k = 0; k < rs_cape_ctx::kUniformSlots; ++k)
        if (pick < 0 || c->uniforms[k].stamp < c->uniforms[pick].stamp) pick = k;
    rs_cape_ctx::UniformSlot& u = c->uniforms[pick];
    u.valid = false;
    RS_CUDA_CHECK(cudaEventSynchronize(u.copied));                 // the staging buffer is free to be rewritten
    for (cudaEvent_t e : u.used) RS_CUDA_CHECK(cudaStreamWaitEvent(stream, e, 0));   // launches still reading the old table
    std::mt19937 eng(seed);
    std::uniform_real_distribution<double> dist(0.0, 1.0);
    for (int i = 0; i < c->n_uniforms; ++i) u.h[i] = dist(eng);
    RS_CUDA_CHECK(cudaMemcpyAsync(u.d, u.h, sizeof(double) * size_t(c->n_uniforms), cudaMemcpyHostToDevice, stream));
    RS_CUDA_CHECK(cudaEventRecord(u.copied, stream));
    u.seed = seed, u.valid = true, u.stamp = ++c->uniform_clock;
    *slot_out = pick;
    return RS_OK;
}

int uniforms_used(rs_cape_ctx* c, int slot, cudaStream_t stream)
{
    std::lock_guard<std::mutex> lock(c->uniform_mutex);
    rs_cape_ctx::UniformSlot& u = c->uniforms[slot];
    RS_CUDA_CHECK(cudaEventRecord(u.used[u.next_use], stream));
    u.next_use = (u.next_use + 1) % rs_cape_ctx::kUniformUses;
    return RS_OK;
}

template <class T>
int dev_alloc(T** p, size_t n)
{
    RS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(p), sizeof(T) * n));
    RS_CUDA_CHECK(cudaMemset(*p, 0, sizeof(T) * n));
    return RS_OK;
}

int create_impl(rs_cape_ctx* c)
{
    int rc = require_blackwell(c->device);
    if (rc != RS_OK) return rc;
    std::vector<double> kx, ky;
    backprojection_factors(c, kx, ky);
    const size_t B = size_t(c->max_batch), Nc = size_t(c->Nc);
    // second half of each table: the factors times 2^896 (exact), for K1a's XU-free widening
    const size_t nkx = kx.size(), nky = ky.size();
    for (size_t i = 0; i < nkx; ++i) kx.push_back(std::ldexp(kx[i], 896));
    for (size_t i = 0; i < nky; ++i) ky.push_back(std::ldexp(ky[i], 896));
    if ((rc = dev_alloc(&c->d_kx, kx.size()))) return rc;
    if ((rc = dev_alloc(&c->d_ky, ky.size()))) return rc;
    RS_CUDA_CHECK(cudaMemcpy(c->d_kx, kx.data(), sizeof(double) * kx.size(), cudaMemcpyHostToDevice));
    RS_CUDA_CHECK(cudaMemcpy(c->d_ky, ky.data(), sizeof(double) * ky.size(), cudaMemcpyHostToDevice));
    if ((rc = dev_alloc(&c->d_depth, B * c->W * c->H))) return rc;
    if ((rc = dev_alloc(&c->d_out.cells, B * Nc))) return rc;
    if ((rc = dev_alloc(&c->d_out.plane_grid, B * Nc))) return rc;
    if ((rc = dev_alloc(&c->d_out.plane_labels, B * Nc))) return rc;
    if ((rc = dev_alloc(&c->d_out.cyl_labels, B * Nc))) return rc;
    if ((rc = dev_alloc(&c->d_out.cyl_region_seg, B * Nc))) return rc;
    if ((rc = dev_alloc(&c->d_out.planes, B * RS_MAX_PLANES))) return rc;
    if ((rc = dev_alloc(&c->d_out.cyls, B * RS_MAX_CYL_REGIONS))) return rc;
    if ((rc = dev_alloc(&c->d_out.boundary_xyz, B * size_t(c->max_boundary) * 3))) return rc;
    if ((rc = dev_alloc(&c->d_out.info, B))) return rc;
    if ((rc = dev_alloc(&c->d_scratch, B * cape_segment_scratch_doubles_per_frame(c->Nc)))) return rc;
    c->n_uniforms = 3 * RS_CYL_RANSAC_ITERS * RS_MAX_CYL_REGIONS * RS_MAX_CYL_SEGS;
    for (rs_cape_ctx::UniformSlot& u : c->uniforms) {
        if ((rc = dev_alloc(&u.d, size_t(c->n_uniforms)))) return rc;
        RS_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&u.h), sizeof(double) * size_t(c->n_uniforms), cudaHostAllocDefault));
        RS_CUDA_CHECK(cudaEventCreateWithFlags(&u.copied, cudaEventDisableTiming));
        for (cudaEvent_t& e : u.used) RS_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    RS_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    RS_CUDA_CHECK(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
    RS_CUDA_CHECK(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    for (cudaStream_t& ss : c->seg_streams) RS_CUDA_CHECK(cudaStreamCreateWithFlags(&ss, cudaStreamNonBlocking));
    c->chunk_events.assign(2 * size_t((c->max_batch + kChunkFrames - 1) / kChunkFrames), nullptr);
    for (cudaEvent_t& e : c->chunk_events) RS_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    RS_CUDA_CHECK(cudaEventCreateWithFlags(&c->fit_done, cudaEventDisableTiming));
    RS_CUDA_CHECK(cudaEventCreateWithFlags(&c->streamed, cudaEventDisableTiming));

    const unsigned P = unsigned(c->cell) * unsigned(c->cell);
    c->fit.H = c->H, c->fit.hc = c->hc, c->fit.vc = c->vc, c->fit.cell = c->cell;
    c->fit.kx = c->d_kx, c->fit.ky = c->d_ky;
    c->fit.kxs = c->d_kx + nkx, c->fit.kys = c->d_ky + nky;
    c->fit.min_zero_point_count = int(static_cast<unsigned>(std::floor(static_cast<float>(P) * 0.7f)));
    c->fit.sin_merge = sinf(static_cast<float>(18.0f * M_PI / 180.0));
    c->fit.merge_distance = 50.0f;
    c->seg.W = c->W, c->seg.H = c->H, c->seg.hc = c->hc, c->seg.vc = c->vc, c->seg.cell = c->cell;
    c->seg.kx = c->d_kx, c->seg.ky = c->d_ky;
    c->seg.cos_merge = std::cos(18.0f * M_PI / 180.0);
    c->seg.uniforms = nullptr, c->seg.n_uniforms = c->n_uniforms;   // the table of a run's seed is bound per launch
    c->seg.max_boundary = c->max_boundary;
    return RS_OK;
}

// fit = launch K1 (rectification first when enabled), segment = launch K2-K4 on the records in o->cells. A run with
// fit && segment is find_primitives; rs_cape_cell_fit_device and rs_cape_segment_device are its two halves.
int run_device_impl(rs_cape_ctx* c, const float* depth_dev, int batch, uint32_t seed, const rs_cape_outputs* o,
                    cudaStream_t stream, bool fit, bool segment, bool timing = true, cudaStream_t seg_stream = nullptr,
                    size_t scratch_frame = 0)
{
    if (!c || !depth_dev || batch <= 0 || batch > c->max_batch || !o || !o->cells) {
        set_last_error("rs_cape_run: invalid argument (null pointer or batch out of range)");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    int rc;
    if (c->rectify) {
        // rectify_depth in front of the path (rgbd_slam.cpp:85-97): K1 and the boundary step then read the rectified image
        float* rect = c->d_rect + scratch_frame * size_t(c->W) * c->H;
        if (fit) {
            RectifyParams rp = c->rect;
            rp.batch = batch;
            if ((rc = launch_rectify_depth(rp, depth_dev, rect, stream)) != RS_OK)
                return rc;
        }
        depth_dev = rect;   // a segmentation-only call reads what the preceding fit call rectified
    }
    cudaEvent_t* ev = nullptr;
    if (timing && c->timing_slots > 0) {
        // a segmentation-only call completes the slot its fit call opened
        const uint64_t run = fit ? c->run_counter : (c->run_counter ? c->run_counter - 1 : 0);
        ev = &c->events[size_t(run % uint64_t(c->timing_slots)) * 4];
    }
    if (fit) {
        rc = encode_tmap(c, depth_dev, batch);
        if (rc != RS_OK) return rc;
        CellFitParams fp = c->fit;
        fp.batch = batch;
        if (timing) ++c->run_counter;
        if (ev) RS_CUDA_CHECK(cudaEventRecord(ev[0], stream));
        if ((rc = launch_cape_cell_fit(c->tmap, fp, o->cells, stream, c->streamed)) != RS_OK) return rc;
        RS_CUDA_CHECK(cudaEventRecord(c->fit_done, stream));
        if (ev) {
            RS_CUDA_CHECK(cudaEventRecord(ev[1], stream));
            if (!segment) {   // until a segmentation-only call overwrites them
                RS_CUDA_CHECK(cudaEventRecord(ev[2], stream));
                RS_CUDA_CHECK(cudaEventRecord(ev[3], stream));
            }
        }
    }
    if (!segment) return RS_OK;
    if (!o->plane_grid || !o->plane_labels || !o->cyl_labels || !o->cyl_region_seg || !o->planes || !o->cyls ||
        !o->boundary_xyz || !o->info) {
        set_last_error("rs_cape_run_device: all device output buffers are required");
        return RS_ERR_INVALID_ARG;
    }
    cudaStream_t launch_stream = (seg_stream && seg_stream != stream) ? seg_stream : stream;
    int uslot = 0;
    if ((rc = ensure_uniforms(c, seed, launch_stream, &uslot)) != RS_OK) return rc;
    SegmentParams sp = c->seg;
    sp.batch = batch;
    sp.uniforms = c->uniforms[uslot].d;
    SegmentBuffers sb;
    sb.depth = depth_dev;
    sb.cells = o->cells;
    sb.plane_grid = o->plane_grid;
    sb.plane_labels = o->plane_labels;
    sb.cyl_labels = o->cyl_labels;
    sb.cyl_region_seg = o->cyl_region_seg;
    sb.planes = o->planes;
    sb.cyls = o->cyls;
    sb.boundary_xyz = o->boundary_xyz;
    sb.info = o->info;
    sb.scratch = c->d_scratch + scratch_frame * cape_segment_scratch_doubles_per_frame(c->Nc);
    if (seg_stream && seg_stream != stream) {
        // the latency-bound segmentation of this chunk runs beside the plane fit / segmentation of the next ones
        RS_CUDA_CHECK(cudaStreamWaitEvent(seg_stream, c->fit_done, 0));
        if ((rc = launch_cape_segment(sp, sb, seg_stream)) != RS_OK) return rc;
        return uniforms_used(c, uslot, seg_stream);
    }
    if (ev) RS_CUDA_CHECK(cudaEventRecord(ev[2], stream));
    if ((rc = launch_cape_segment(sp, sb, stream)) != RS_OK) return rc;
    if (ev) RS_CUDA_CHECK(cudaEventRecord(ev[3], stream));
    return uniforms_used(c, uslot, stream);
}

}  // namespace

extern "C" {

rs_cape_ctx* rs_cape_create(int width, int height, int cell_px, double fx, double fy, double cx, double cy, int max_batch,
                            int device)
{
    if (width <= 0 || height <= 0 || cell_px <= 0 || cell_px % 4 != 0 || width < cell_px || height < cell_px ||
        max_batch <= 0 || width % 4 != 0) {
        set_last_error("rs_cape_create: invalid geometry (cell_px and width must be multiples of 4)");
        return nullptr;
    }
    if (cape_cell_fit_box_rows(cell_px) <= 0) {
        set_last_error("rs_cape_create: unsupported cell size (the plane-fit kernel is built for 20 and 40 px cells)");
        return nullptr;
    }
    if (cape_segment_validate(width / cell_px, height / cell_px, cell_px) != RS_OK) return nullptr;   // rs_last_error() says why
    rs_cape_ctx* c = new rs_cape_ctx();
    c->W = width, c->H = height, c->cell = cell_px;
    c->hc = width / cell_px, c->vc = height / cell_px, c->Nc = c->hc * c->vc;
    c->fx = fx, c->fy = fy, c->cx = cx, c->cy = cy;
    c->max_batch = max_batch, c->device = device;
    c->max_boundary = 2 * c->Nc;
    if (create_impl(c) != RS_OK) {
        rs_cape_destroy(c);
        return nullptr;
    }
    return c;
}

void rs_cape_destroy(rs_cape_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->d_kx);
    cudaFree(c->d_ky);
    cudaFree(c->d_pre);
    cudaFree(c->d_depth);
    cudaFree(c->d_out.cells);
    cudaFree(c->d_out.plane_grid);
    cudaFree(c->d_out.plane_labels);
    cudaFree(c->d_out.cyl_labels);
    cudaFree(c->d_out.cyl_region_seg);
    cudaFree(c->d_out.planes);
    cudaFree(c->d_out.cyls);
    cudaFree(c->d_out.boundary_xyz);
    cudaFree(c->d_out.info);
    for (rs_cape_ctx::UniformSlot& u : c->uniforms) {
        cudaFree(u.d);
        if (u.h) cudaFreeHost(u.h);
        if (u.copied) cudaEventDestroy(u.copied);
        for (cudaEvent_t e : u.used)
            if (e) cudaEventDestroy(e);
    }
    cudaFree(c->d_scratch);
    cudaFree(c->d_depth16);
    cudaFree(c->d_rect);
    for (cudaEvent_t e : c->events) cudaEventDestroy(e);
    if (c->fit_done) cudaEventDestroy(c->fit_done);
    if (c->streamed) cudaEventDestroy(c->streamed);
    for (cudaEvent_t e : c->chunk_events)
        if (e) cudaEventDestroy(e);
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    for (cudaStream_t ss : c->seg_streams)
        if (ss) cudaStreamDestroy(ss);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int rs_cape_set_timing(rs_cape_ctx* c, int n_slots)
{
    if (!c || n_slots < 0 || n_slots > 4096) {
        set_last_error("rs_cape_set_timing: invalid argument");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    for (cudaEvent_t e : c->events) cudaEventDestroy(e);
    c->events.assign(size_t(n_slots) * 4, nullptr);
    for (cudaEvent_t& e : c->events) RS_CUDA_CHECK(cudaEventCreate(&e));
    c->timing_slots = n_slots;
    c->run_counter = 0;
    return RS_OK;
}

int rs_cape_kernel_ms(rs_cape_ctx* c, int slot, float ms[2])
{
    if (!c || !ms || slot < 0 || slot >= c->timing_slots || uint64_t(slot) >= c->run_counter) {
        set_last_error("rs_cape_kernel_ms: timing is off or that slot has not been recorded");
        return RS_ERR_INVALID_ARG;
    }
    cudaEvent_t* ev = &c->events[size_t(slot) * 4];
    RS_CUDA_CHECK(cudaEventSynchronize(ev[3]));
    RS_CUDA_CHECK(cudaEventElapsedTime(&ms[0], ev[0], ev[1]));
    RS_CUDA_CHECK(cudaEventElapsedTime(&ms[1], ev[2], ev[3]));
    return RS_OK;
}

int rs_cape_set_rectification(rs_cape_ctx* c, const double* cam2_to_cam1, int enable)
{
    if (!c || (enable && !cam2_to_cam1)) {
        set_last_error("rs_cape_set_rectification: invalid argument");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    if (enable) {
        const size_t px = size_t(c->max_batch) * c->W * c->H;
        if (!c->d_rect) RS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&c->d_rect), sizeof(float) * px));
        c->rect.W = c->W, c->rect.H = c->H;
        if (!c->d_pre) {
            // _Xpre / _Ypre are float images in the reference: the FP64 factors rounded once
            std::vector<double> kx, ky;
            backprojection_factors(c, kx, ky);
            std::vector<float> pre;
            for (int i = 0; i < c->W; ++i) pre.push_back(static_cast<float>(kx[i]));
            for (int i = 0; i < c->H; ++i) pre.push_back(static_cast<float>(ky[i]));
            RS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&c->d_pre), sizeof(float) * pre.size()));
            RS_CUDA_CHECK(cudaMemcpy(c->d_pre, pre.data(), sizeof(float) * pre.size(), cudaMemcpyHostToDevice));
        }
        c->rect.preX = c->d_pre, c->rect.preY = c->d_pre + c->W;
        c->rect.fx = c->fx, c->rect.fy = c->fy, c->rect.cx = c->cx, c->rect.cy = c->cy;
        for (int i = 0; i < 12; ++i) c->rect.T[i] = cam2_to_cam1[i];
    }
    c->rectify = enable != 0;
    c->tmap_ptr = nullptr;  // the plane fit reads another buffer from now on
    return RS_OK;
}

int rs_cape_rectify(rs_cape_ctx* c, const float* depth_host, int batch, float* rectified_host)
{
    if (!c || !depth_host || !rectified_host || batch <= 0 || batch > c->max_batch || !c->d_rect) {
        set_last_error("rs_cape_rectify: invalid argument, or rs_cape_set_rectification has not been called");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    const size_t px = size_t(batch) * c->W * c->H;
    RS_CUDA_CHECK(cudaMemcpyAsync(c->d_depth, depth_host, sizeof(float) * px, cudaMemcpyHostToDevice, c->stream));
    RectifyParams rp = c->rect;
    rp.batch = batch;
    const int rc = launch_rectify_depth(rp, c->d_depth, c->d_rect, c->stream);
    if (rc != RS_OK) return rc;
    RS_CUDA_CHECK(cudaMemcpyAsync(rectified_host, c->d_rect, sizeof(float) * px, cudaMemcpyDeviceToHost, c->stream));
    RS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return RS_OK;
}

int rs_cape_rectify_device(rs_cape_ctx* c, const float* depth_dev, int batch, float* rectified_dev, void* stream)
{
    if (!c || !depth_dev || !rectified_dev || batch <= 0 || batch > c->max_batch || !c->d_rect) {
        set_last_error("rs_cape_rectify_device: invalid argument, or rs_cape_set_rectification has not been called");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    RectifyParams rp = c->rect;
    rp.batch = batch;
    return launch_rectify_depth(rp, depth_dev, rectified_dev, static_cast<cudaStream_t>(stream));
}

int rs_cape_stream_wait_fit(rs_cape_ctx* c, void* stream)
{
    if (!c) {
        set_last_error("rs_cape_stream_wait_fit: null context");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), c->streamed, 0));
    return RS_OK;
}

int rs_cape_cells_per_frame(const rs_cape_ctx* c) { return c ? c->Nc : 0; }
int rs_cape_max_boundary(const rs_cape_ctx* c) { return c ? c->max_boundary : 0; }
float* rs_cape_device_depth(rs_cape_ctx* c) { return c ? c->d_depth : nullptr; }
const rs_cape_outputs* rs_cape_device_outputs(rs_cape_ctx* c) { return c ? &c->d_out : nullptr; }

int rs_cape_cell_fit_device(rs_cape_ctx* c, const float* depth_dev, int batch, rs_cell_out* cells_dev, void* stream)
{
    rs_cape_outputs o{};
    o.cells = cells_dev;
    return run_device_impl(c, depth_dev, batch, 0, &o, static_cast<cudaStream_t>(stream), true, false);
}

int rs_cape_run_device(rs_cape_ctx* c, const float* depth_dev, int batch, uint32_t seed, const rs_cape_outputs* out_dev,
                       void* stream)
{
    return run_device_impl(c, depth_dev, batch, seed, out_dev, static_cast<cudaStream_t>(stream), true, true);
}

int rs_cape_segment_device(rs_cape_ctx* c, const float* depth_dev, int batch, uint32_t seed, const rs_cape_outputs* out_dev,
                           void* stream)
{
    return run_device_impl(c, depth_dev, batch, seed, out_dev, static_cast<cudaStream_t>(stream), false, true);
}

static int run_host_impl(rs_cape_ctx* c, const float* depth_host, const uint16_t* depth16_host, float alpha, int batch,
                         uint32_t seed, const rs_cape_outputs* out)
{
    if (!c || (!depth_host && !depth16_host) || !out || batch <= 0 || batch > c->max_batch) {
        set_last_error("rs_cape_run: invalid argument (null pointer or batch out of range)");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    const size_t Nc = size_t(c->Nc), px = size_t(c->W) * c->H, mb = size_t(c->max_boundary);
    const bool cells_only = !out->plane_grid && !out->plane_labels && !out->cyl_labels && !out->cyl_region_seg &&
                            !out->planes && !out->cyls && !out->boundary_xyz && !out->info;
    const rs_cape_outputs& d = c->d_out;
    int rc;
    int k = 0;
    for (int f0 = 0; f0 < batch; f0 += kChunkFrames, ++k) {
        const size_t n = size_t(std::min(kChunkFrames, batch - f0)), o = size_t(f0);
        cudaEvent_t landed = c->chunk_events[2 * k], ready = c->chunk_events[2 * k + 1];
        if (depth16_host)
            RS_CUDA_CHECK(cudaMemcpyAsync(c->d_depth16 + o * px, depth16_host + o * px, sizeof(uint16_t) * n * px,
                                          cudaMemcpyHostToDevice, c->h2d_stream));
        else
            RS_CUDA_CHECK(cudaMemcpyAsync(c->d_depth + o * px, depth_host + o * px, sizeof(float) * n * px, cudaMemcpyHostToDevice,
                                          c->h2d_stream));
        RS_CUDA_CHECK(cudaEventRecord(landed, c->h2d_stream));
        RS_CUDA_CHECK(cudaStreamWaitEvent(c->stream, landed, 0));
        if (depth16_host) {
            depth_u16_to_f32_kernel<<<sm_count() * 8, 256, 0, c->stream>>>(reinterpret_cast<const uint4*>(c->d_depth16 + o * px),
                                                                   reinterpret_cast<float4*>(c->d_depth + o * px), n * px / 8, alpha);
            RS_LAUNCH_CHECK();
        }
        rs_cape_outputs dchunk;
        dchunk.cells = d.cells + o * Nc;
        dchunk.plane_grid = d.plane_grid + o * Nc;
        dchunk.plane_labels = d.plane_labels + o * Nc;
        dchunk.cyl_labels = d.cyl_labels + o * Nc;
        dchunk.cyl_region_seg = d.cyl_region_seg + o * Nc;
        dchunk.planes = d.planes + o * RS_MAX_PLANES;
        dchunk.cyls = d.cyls + o * RS_MAX_CYL_REGIONS;
        dchunk.boundary_xyz = d.boundary_xyz + o * mb * 3;
        dchunk.info = d.info + o;
        cudaStream_t seg = c->seg_streams[k % 4];
        if ((rc = run_device_impl(c, c->d_depth + o * px, int(n), seed, &dchunk, c->stream, true, !cells_only, false, seg, o)) != RS_OK)
            return rc;
        RS_CUDA_CHECK(cudaEventRecord(ready, cells_only ? c->stream : seg));
        RS_CUDA_CHECK(cudaStreamWaitEvent(c->d2h_stream, ready, 0));
#define RS_D2H(field, per_frame)                                                                                         \
    if (out->field)                                                                                                      \
    RS_CUDA_CHECK(cudaMemcpyAsync(out->field + o * (per_frame), d.field + o * (per_frame), sizeof(*d.field) * n * (per_frame), \
                                  cudaMemcpyDeviceToHost, c->d2h_stream))
        RS_D2H(cells, Nc);
        RS_D2H(plane_grid, Nc);
        RS_D2H(plane_labels, Nc);
        RS_D2H(cyl_labels, Nc);
        RS_D2H(cyl_region_seg, Nc);
        RS_D2H(planes, size_t(RS_MAX_PLANES));
        RS_D2H(cyls, size_t(RS_MAX_CYL_REGIONS));
        RS_D2H(boundary_xyz, mb * 3);
        RS_D2H(info, size_t(1));
#undef RS_D2H
    }
    RS_CUDA_CHECK(cudaStreamSynchronize(c->d2h_stream));
    RS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    for (cudaStream_t ss : c->seg_streams) RS_CUDA_CHECK(cudaStreamSynchronize(ss));
    return RS_OK;
}

int rs_cape_run(rs_cape_ctx* c, const float* depth_host, int batch, uint32_t seed, const rs_cape_outputs* out)
{
    return run_host_impl(c, depth_host, nullptr, 1.0f, batch, seed, out);
}

int rs_cape_run_u16(rs_cape_ctx* c, const uint16_t* depth_host, double alpha, int batch, uint32_t seed, const rs_cape_outputs* out)
{
    if (!c || !depth_host) {
        set_last_error("rs_cape_run_u16: invalid argument (null pointer)");
        return RS_ERR_INVALID_ARG;
    }
    if ((size_t(c->W) * c->H) % 8 != 0) {
        set_last_error("rs_cape_run_u16: width * height must be a multiple of 8");
        return RS_ERR_INVALID_ARG;
    }
    if (!c->d_depth16) {
        RS_CUDA_CHECK(cudaSetDevice(c->device));
        RS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&c->d_depth16), sizeof(uint16_t) * size_t(c->max_batch) * c->W * c->H));
    }
    return run_host_impl(c, nullptr, depth_host, static_cast<float>(alpha), batch, seed, out);
}

const char* rs_last_error(void)
{
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lock(g_err_mutex);
    copy = g_last_error;
    return copy.c_str();
}
const char* rs_version(void) { return "rgbdslam_b200 0.1 (sm_100a)"; }
uint64_t rs_launch_count(void) { return g_launch_count.load(); }

}  // extern "C"
