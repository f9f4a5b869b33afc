// NOTE (round 1, later): probes 0 and 9-11 below feed the converted / multiplied values from registers that never change inside
// the loop, so the compiler hoists the conversion / FMUL / IMAD out of it and only the FP64 op is timed; they read too
// optimistic. tools/microbench2.cu (inline PTX, loop-carried dependencies) is the probe the design relies on.
//
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
        } else if (OP == 7) {  // IMAD.WIDE.U32 (float bits * 2^29 -> 64-bit), consumed by xor
            unsigned long long a0 = (unsigned long long)__float_as_uint(f0) * 0x20000000ull, a1 = (unsigned long long)__float_as_uint(f1) * 0x20000000ull;
            unsigned long long a2 = (unsigned long long)__float_as_uint(f2) * 0x20000000ull, a3 = (unsigned long long)__float_as_uint(f3) * 0x20000000ull;
            f0 = __uint_as_float((unsigned)(a0 >> 32) ^ (unsigned)a0); f1 = __uint_as_float((unsigned)(a1 >> 32) ^ (unsigned)a1);
            f2 = __uint_as_float((unsigned)(a2 >> 32) ^ (unsigned)a2); f3 = __uint_as_float((unsigned)(a3 >> 32) ^ (unsigned)a3);
        } else if (OP == 8) {  // the same widening with two shifts
            unsigned b0 = __float_as_uint(f0), b1 = __float_as_uint(f1), b2 = __float_as_uint(f2), b3 = __float_as_uint(f3);
            f0 = __uint_as_float((b0 >> 3) ^ (b0 << 29)); f1 = __uint_as_float((b1 >> 3) ^ (b1 << 29));
            f2 = __uint_as_float((b2 >> 3) ^ (b2 << 29)); f3 = __uint_as_float((b3 >> 3) ^ (b3 << 29));
        } else if (OP == 9) {  // DFMA fed by IMAD.WIDE (the K1a product path): 4 + 4
            d0 = fma(__longlong_as_double((unsigned long long)__float_as_uint(f0) * 0x20000000ull), 0x1p896, d0);
            d1 = fma(__longlong_as_double((unsigned long long)__float_as_uint(f1) * 0x20000000ull), 0x1p896, d1);
            d2 = fma(__longlong_as_double((unsigned long long)__float_as_uint(f2) * 0x20000000ull), 0x1p896, d2);
            d3 = fma(__longlong_as_double((unsigned long long)__float_as_uint(f3) * 0x20000000ull), 0x1p896, d3);
        } else if (OP == 10) {  // FMUL + IMAD.WIDE + DFMA: 4 + 4 + 4
            d0 = fma(__longlong_as_double((unsigned long long)__float_as_uint(f0 * f1) * 0x20000000ull), 0x1p896, d0);
            d1 = fma(__longlong_as_double((unsigned long long)__float_as_uint(f1 * f2) * 0x20000000ull), 0x1p896, d1);
            d2 = fma(__longlong_as_double((unsigned long long)__float_as_uint(f2 * f3) * 0x20000000ull), 0x1p896, d2);
            d3 = fma(__longlong_as_double((unsigned long long)__float_as_uint(f3 * f0) * 0x20000000ull), 0x1p896, d3);
        } else if (OP == 11) {  // FMUL + F2F + DADD: 4 + 4 + 4
            d0 += (double)(f0 * f1); d1 += (double)(f1 * f2); d2 += (double)(f2 * f3); d3 += (double)(f3 * f0);
        } else if (OP == 12) {  // DFMA + 3 FFMA each: 4 + 12
            d0 = fma(d0, 1.0000001, d1); d1 = fma(d1, 1.0000001, d2); d2 = fma(d2, 1.0000001, d3); d3 = fma(d3, 1.0000001, d0);
            f0 = fmaf(f0, 1.0000001f, f1); f1 = fmaf(f1, 1.0000001f, f2); f2 = fmaf(f2, 1.0000001f, f3); f3 = fmaf(f3, 1.0000001f, f0);
            f0 = fmaf(f0, 1.0000001f, f1); f1 = fmaf(f1, 1.0000001f, f2); f2 = fmaf(f2, 1.0000001f, f3); f3 = fmaf(f3, 1.0000001f, f0);
            f0 = fmaf(f0, 1.0000001f, f1); f1 = fmaf(f1, 1.0000001f, f2); f2 = fmaf(f2, 1.0000001f, f3); f3 = fmaf(f3, 1.0000001f, f0);
        } else if (OP == 13) {  // DFMA + 3 LOP3 each: 4 + 12
            d0 = fma(d0, 1.0000001, d1); d1 = fma(d1, 1.0000001, d2); d2 = fma(d2, 1.0000001, d3); d3 = fma(d3, 1.0000001, d0);
            unsigned b0 = __float_as_uint(f0), b1 = __float_as_uint(f1), b2 = __float_as_uint(f2), b3 = __float_as_uint(f3);
            b0 = (b0 & 0x7ffffff0u) ^ b1; b1 = (b1 & 0x7ffffff0u) ^ b2; b2 = (b2 & 0x7ffffff0u) ^ b3; b3 = (b3 & 0x7ffffff0u) ^ b0;
            b0 = (b0 | 0x100u) ^ b1; b1 = (b1 | 0x100u) ^ b2; b2 = (b2 | 0x100u) ^ b3; b3 = (b3 | 0x100u) ^ b0;
            b0 = (b0 & 0x7ffffff0u) ^ b1; b1 = (b1 & 0x7ffffff0u) ^ b2; b2 = (b2 & 0x7ffffff0u) ^ b3; b3 = (b3 & 0x7ffffff0u) ^ b0;
            f0 = __uint_as_float(b0); f1 = __uint_as_float(b1); f2 = __uint_as_float(b2); f3 = __uint_as_float(b3);
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
    run<7>("imad.wide.u32 (+lop3)", 4);
    run<8>("shr+shl (+lop3)", 4);
    run<9>("imad.wide + dfma (4+4)", 4);
    run<10>("fmul+imad.wide+dfma (4+4+4)", 4);
    run<11>("fmul+f2f+dadd (4+4+4)", 4);
    run<12>("dfma + 3 ffma (4+12)", 4);
    run<13>("dfma + 3 lop3 (4+12)", 4);
    return 0;
}
