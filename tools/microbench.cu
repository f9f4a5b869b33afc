// Pipe-throughput probe for the K1 design decision (SURVEY.md §7 "Instruction budget"): how fast are
// F2F.F64.F32 / F2F.F32.F64 / DADD / DMUL / DFMA / FFMA per SM on this B200? Run under gpurun.
#include <cuda_runtime.h>
#include <cstdio>

template <int OP>
__global__ void probe(double* out, int iters, float seedf)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float f0 = seedf + tid, f1 = f0 + 1.f, f2 = f0 + 2.f, f3 = f0 + 3.f;
    double d0 = f0, d1 = f1, d2 = f2, d3 = f3;
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) {  // F2F.F64.F32 (+ FADD to keep a float chain alive)
            d0 += (double)f0; d1 += (double)f1; d2 += (double)f2; d3 += (double)f3;   // cvt + DADD
        } else if (OP == 1) {  // DADD only
            d0 += d1; d1 += d2; d2 += d3; d3 += d0;
        } else if (OP == 2) {  // DMUL
            d0 *= 1.0000001; d1 *= 1.0000001; d2 *= 1.0000001; d3 *= 1.0000001;
        } else if (OP == 3) {  // DFMA
            d0 = fma(d0, 1.0000001, d1); d1 = fma(d1, 1.0000001, d2); d2 = fma(d2, 1.0000001, d3); d3 = fma(d3, 1.0000001, d0);
        } else if (OP == 4) {  // F2F.F32.F64 + F2F.F64.F32 round trip
            f0 = (float)d0; d0 = (double)f0 + 1.0; f1 = (float)d1; d1 = (double)f1 + 1.0;
            f2 = (float)d2; d2 = (double)f2 + 1.0; f3 = (float)d3; d3 = (double)f3 + 1.0;
        } else if (OP == 5) {  // FFMA
            f0 = fmaf(f0, 1.0000001f, f1); f1 = fmaf(f1, 1.0000001f, f2); f2 = fmaf(f2, 1.0000001f, f3); f3 = fmaf(f3, 1.0000001f, f0);
        } else if (OP == 6) {  // cvt only: float -> double, consumed by integer xor (no FP64 pipe op besides cvt)
            long long a = __double_as_longlong((double)f0) ^ __double_as_longlong((double)f1) ^
                          __double_as_longlong((double)f2) ^ __double_as_longlong((double)f3);
            f0 += (float)(a & 1); f1 += 1.f; f2 += 1.f; f3 += 1.f;
        }
    }
    out[tid] = d0 + d1 + d2 + d3 + f0 + f1 + f2 + f3;
}

template <int OP>
void run(const char* name, double ops_per_iter)
{
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe<OP><<<blocks, threads>>>(out, iters, 1.f);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    probe<OP><<<blocks, threads>>>(out, iters, 1.f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double total = double(blocks) * threads * iters * ops_per_iter;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %8.2f Gop/s  %6.1f op/clk/SM (at %d MHz nominal)\n", name, ms, total / ms * 1e-6,
           total / (ms * 1e-3) / 148.0 / (clk * 1e3), clk / 1000);
    cudaFree(out);
}

int main()
{
    run<0>("cvt.f64.f32 + dadd (4+4)", 4);
    run<1>("dadd", 4);
    run<2>("dmul", 4);
    run<3>("dfma", 4);
    run<4>("cvt f64->f32->f64 + dadd", 4);
    run<5>("ffma", 4);
    run<6>("cvt.f64.f32 only", 4);
    return 0;
}
