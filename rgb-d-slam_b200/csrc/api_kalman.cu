// C-ABI glue of the batched Kalman update (include/rgbdslam_b200.h). No CPU path: without an sm_100 device it fails.
#include <cuda_runtime.h>

#include <string>

#include "common.cuh"
#include "kalman_internal.cuh"

using namespace rs;

namespace {

int track_device(int dim, int n, const double* state, const double* cov, const double* meas, const double* meas_cov,
                 double process_noise, double* out_state, double* out_cov, double* out_score, uint8_t* out_moving,
                 int32_t* out_status, cudaStream_t stream)
{
    if (n < 0 || (n > 0 && (!state || !cov || !meas || !meas_cov || !out_state || !out_cov || !out_score || !out_status))) {
        set_last_error("rs_kalman_track: invalid argument (null pointer or negative count)");
        return RS_ERR_INVALID_ARG;
    }
    KalmanBatch b;
    b.n = n, b.process_noise = process_noise;
    b.state = state, b.cov = cov, b.meas = meas, b.meas_cov = meas_cov;
    b.out_state = out_state, b.out_cov = out_cov, b.out_score = out_score, b.out_moving = out_moving, b.out_status = out_status;
    return launch_kalman_track(b, dim, stream);
}

int track_host(int dim, int device, int n, const double* state, const double* cov, const double* meas, const double* meas_cov,
               double process_noise, double* out_state, double* out_cov, double* out_score, uint8_t* out_moving,
               int32_t* out_status)
{
    int rc = require_blackwell(device);
    if (rc != RS_OK) return rc;
    if (n < 0 || (n > 0 && (!state || !cov || !meas || !meas_cov || !out_state || !out_cov || !out_score || !out_status))) {
        set_last_error("rs_kalman_track: invalid argument (null pointer or negative count)");
        return RS_ERR_INVALID_ARG;
    }
    if (n == 0) return RS_OK;
    const size_t v = sizeof(double) * size_t(n) * dim, m = v * dim;
    // one allocation: state | cov | meas | meas_cov | out_state | out_cov | out_score | status | moving
    const size_t total = 2 * v + 2 * m + v + m + sizeof(double) * n + sizeof(int32_t) * n + size_t(n);
    unsigned char* d = nullptr;
    RS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d), total));
    double* d_state = reinterpret_cast<double*>(d);
    double* d_cov = d_state + size_t(n) * dim;
    double* d_meas = d_cov + size_t(n) * dim * dim;
    double* d_mcov = d_meas + size_t(n) * dim;
    double* d_ostate = d_mcov + size_t(n) * dim * dim;
    double* d_ocov = d_ostate + size_t(n) * dim;
    double* d_score = d_ocov + size_t(n) * dim * dim;
    int32_t* d_status = reinterpret_cast<int32_t*>(d_score + n);
    uint8_t* d_moving = reinterpret_cast<uint8_t*>(d_status + n);
    cudaStream_t s = nullptr;
    auto fail = [&](int code) {
        cudaFree(d);
        return code;
    };
    if (cudaMemcpyAsync(d_state, state, v, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(d_cov, cov, m, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(d_meas, meas, v, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(d_mcov, meas_cov, m, cudaMemcpyHostToDevice, s) != cudaSuccess) {
        set_last_error("rs_kalman_track: host to device copy failed");
        return fail(RS_ERR_CUDA);
    }
    rc = track_device(dim, n, d_state, d_cov, d_meas, d_mcov, process_noise, d_ostate, d_ocov, d_score, d_moving, d_status, s);
    if (rc != RS_OK) return fail(rc);
    bool ok = cudaMemcpyAsync(out_state, d_ostate, v, cudaMemcpyDeviceToHost, s) == cudaSuccess &&
              cudaMemcpyAsync(out_cov, d_ocov, m, cudaMemcpyDeviceToHost, s) == cudaSuccess &&
              cudaMemcpyAsync(out_score, d_score, sizeof(double) * n, cudaMemcpyDeviceToHost, s) == cudaSuccess &&
              cudaMemcpyAsync(out_status, d_status, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s) == cudaSuccess;
    if (ok && out_moving) ok = cudaMemcpyAsync(out_moving, d_moving, size_t(n), cudaMemcpyDeviceToHost, s) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(s) == cudaSuccess;
    if (!ok) {
        set_last_error(std::string("rs_kalman_track: device to host copy failed: ") + cudaGetErrorString(cudaGetLastError()));
        return fail(RS_ERR_CUDA);
    }
    cudaFree(d);
    return RS_OK;
}

}  // namespace

extern "C" {

int rs_kalman_track_points(int device, int n, const double* state, const double* cov, const double* meas, const double* meas_cov,
                           double process_noise, double* out_state, double* out_cov, double* out_score, uint8_t* out_moving,
                           int32_t* out_status)
{
    return track_host(3, device, n, state, cov, meas, meas_cov, process_noise, out_state, out_cov, out_score, out_moving, out_status);
}

int rs_kalman_track_planes(int device, int n, const double* state, const double* cov, const double* meas, const double* meas_cov,
                           double process_noise, double* out_state, double* out_cov, double* out_score, int32_t* out_status)
{
    return track_host(4, device, n, state, cov, meas, meas_cov, process_noise, out_state, out_cov, out_score, nullptr, out_status);
}

int rs_kalman_track_points_device(int n, const double* state, const double* cov, const double* meas, const double* meas_cov,
                                  double process_noise, double* out_state, double* out_cov, double* out_score,
                                  uint8_t* out_moving, int32_t* out_status, void* stream)
{
    return track_device(3, n, state, cov, meas, meas_cov, process_noise, out_state, out_cov, out_score, out_moving, out_status,
                        static_cast<cudaStream_t>(stream));
}

int rs_kalman_track_planes_device(int n, const double* state, const double* cov, const double* meas, const double* meas_cov,
                                  double process_noise, double* out_state, double* out_cov, double* out_score,
                                  int32_t* out_status, void* stream)
{
    return track_device(4, n, state, cov, meas, meas_cov, process_noise, out_state, out_cov, out_score, nullptr, out_status,
                        static_cast<cudaStream_t>(stream));
}

}  // extern "C"
