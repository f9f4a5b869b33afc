// Issue-cost probe for the K1a design (round 1): per-warp-instruction cost of the instruction classes the pixel loop mixes,
// alone and next to DFMA, with real loop-carried dependencies (inline PTX, nothing hoistable). 64 warps/SM resident.
// Output: cycles per SM sub-partition per warp-instruction group. Run under gpurun: ./tools/microbench2
#include <cuda_runtime.h>
#include <cstdio>

#define REP4(X) X(0) X(1) X(2) X(3)

template <int OP>
__global__ void probe(double* out, int iters, float seedf)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float f[4] = {seedf + tid, seedf + tid + 1.f, seedf + tid + 2.f, seedf + tid + 3.f};
    unsigned b[4] = {unsigned(tid) * 2654435761u, unsigned(tid) + 12345u, unsigned(tid) ^ 0x5555u, unsigned(tid) * 31u};
    double d[4] = {f[0], f[1], f[2], f[3]};
    unsigned long long w[4] = {0, 0, 0, 0};
    for (int i = 0; i < iters; ++i) {
#define DFMA(k) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(1.0000001), "d"(0.5));
#define FMUL(k) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[k]) : "f"(1.0000001f));
#define LOP(k) asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[k]) : "r"(b[(k + 1) & 3]));
#define SHF(k) asm volatile("shr.u32 %0, %0, 3;" : "+r"(b[k]));
#define IADD(k) asm volatile("add.u32 %0, %0, %1;" : "+r"(b[k]) : "r"(b[(k + 1) & 3]));
#define WIDE(k) asm volatile("mul.wide.u32 %0, %1, 0x20000000;" : "=l"(w[k]) : "r"(b[k])); asm volatile("mov.b64 {%0, _}, %1;" : "=r"(b[k]) : "l"(w[k]));
#define WIDEONLY(k) asm volatile("mul.wide.u32 %0, %1, 0x20000000;" : "=l"(w[k]) : "r"(b[k] + i));
#define F2FW(k) asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[k]) : "f"(f[k])); asm volatile("{.reg .b32 t; mov.b64 {%0, t}, %1;}" : "=f"(f[k]) : "d"(d[k]));
#define F2FN(k) asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[k]) : "d"(d[k])); asm volatile("mov.b64 %0, {%1, %1};" : "=d"(d[k]) : "f"(f[k]));
#define FSEL(k) asm volatile("{.reg .pred p; setp.gt.f32 p, %0, 0f00000000; selp.f32 %0, %0, 0f3F800000, p;}" : "+f"(f[k]));
        if (OP == 0) { REP4(DFMA) }
        if (OP == 1) { REP4(FMUL) }
        if (OP == 2) { REP4(LOP) }
        if (OP == 3) { REP4(SHF) }
        if (OP == 4) { REP4(IADD) }
        if (OP == 5) { REP4(WIDE) }
        if (OP == 6) { REP4(F2FW) }
        if (OP == 7) { REP4(F2FN) }
        if (OP == 8) { REP4(FSEL) }
        if (OP == 10) { REP4(DFMA) REP4(FMUL) }
        if (OP == 11) { REP4(DFMA) REP4(LOP) }
        if (OP == 12) { REP4(DFMA) REP4(WIDE) }
        if (OP == 13) { REP4(DFMA) REP4(F2FW) }
        if (OP == 14) { REP4(DFMA) REP4(FMUL) REP4(FMUL) REP4(FMUL) }
        if (OP == 15) { REP4(DFMA) REP4(LOP) REP4(FMUL) REP4(F2FW) }
        if (OP == 16) { REP4(FMUL) REP4(LOP) }
        if (OP == 17) { REP4(FMUL) REP4(FMUL) REP4(LOP) REP4(LOP) }
    }
    out[tid] = d[0] + d[1] + d[2] + d[3] + f[0] + f[1] + f[2] + f[3] + b[0] + b[1] + b[2] + b[3] + double(w[0] ^ w[1] ^ w[2] ^ w[3]);
}

template <int OP>
void run(const char* name)
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
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // 16 warps per SM sub-partition, each runs `iters` groups of 4 (x pattern) instructions
    const double cyc = ms * 1e-3 * clk * 1e3 / (16.0 * iters);
    printf("%-44s %8.3f ms  %6.2f cycles per group of 4 per sub-partition\n", name, ms, cyc);
    cudaFree(out);
}

int main()
{
    run<0>("4 dfma");
    run<1>("4 fmul");
    run<2>("4 xor (lop3)");
    run<3>("4 shr");
    run<4>("4 iadd");
    run<5>("4 (mul.wide.u32 + mov)");
    run<6>("4 (cvt.f64.f32 + mov)");
    run<7>("4 (cvt.rn.f32.f64 + mov)");
    run<8>("4 (fsetp + fsel)");
    run<10>("4 dfma + 4 fmul");
    run<11>("4 dfma + 4 lop3");
    run<12>("4 dfma + 4 mul.wide");
    run<13>("4 dfma + 4 cvt.f64.f32");
    run<14>("4 dfma + 12 fmul");
    run<15>("4 dfma + 4 lop3 + 4 fmul + 4 cvt");
    run<16>("4 fmul + 4 lop3");
    run<17>("8 fmul + 8 lop3");
    return 0;
}
