// Internal declarations of the batched Kalman update (not part of the public header).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rs {

struct KalmanBatch {
    int n;
    double process_noise;
    const double* state;      // n x N
    const double* cov;        // n x N x N
    const double* meas;       // n x N
    const double* meas_cov;   // n x N x N
    double* out_state;
    double* out_cov;
    double* out_score;        // |state - new state|, -1 when the update was refused
    unsigned char* out_moving;  // points only (may be null)
    int32_t* out_status;      // 0, or -1 / -2 invalid state / measurement covariance, -3 singular innovation, -4 invalid result
};

int launch_kalman_track(const KalmanBatch& b, int dim, cudaStream_t stream);
int require_blackwell(int device);

}  // namespace rs
