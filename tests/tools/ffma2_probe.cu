// Microbenchmark: issue rate of FFMA (register x constant-bank), FFMA (3 registers) and FFMA2 (packed pairs) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o ffma2_probe ffma2_probe.cu ; prints flops per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

struct P { float c[32]; };

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const __grid_constant__ P p, float seed, int iters) {
    float a[16];
    float2 a2[16];
    for (int i = 0; i < 16; i++) { a[i] = seed + i + threadIdx.x; a2[i] = make_float2(a[i], a[i] + 1.f); }
    float x = seed * 0.5f + threadIdx.x;
    float2 x2 = make_float2(x, x + 0.25f);
    float2 cc[4];
    for (int i = 0; i < 4; i++) cc[i] = make_float2(p.c[i] + seed, p.c[i] + seed);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (MODE == 0) a[i] = __fmaf_rn(x, p.c[(r * 3 + i) & 31], a[i]);          // register x constant
                if (MODE == 1) a[i] = __fmaf_rn(x, cc[(r + i) & 3].x, a[i]);               // three registers
                if (MODE == 2) a2[i] = __ffma2_rn(x2, cc[(r + i) & 3], a2[i]);             // packed, three register pairs
                if (MODE == 3) { const float c = p.c[(r * 3 + i) & 31]; a2[i] = __ffma2_rn(x2, make_float2(c, c), a2[i]); }   // packed x constant
                if (MODE == 4) a2[i] = __ffma2_rn(a2[(i + 1) & 15], cc[(r + i) & 3], x2);   // no accumulate-in-place: 3 distinct pairs
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 16; i++) s += a[i] + a2[i].x + a2[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int per_instr) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    P p; for (int i = 0; i < 32; i++) p.c[i] = 1.0f / (i + 3);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    for (int ctas = 1; ctas <= 8; ctas *= 2) {
        k<MODE><<<148 * ctas, 256>>>(out, p, 1.0f, 10);
        cudaEventRecord(e0);
        k<MODE><<<148 * ctas, 256>>>(out, p, 1.0f, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double instr = (double)ctas * 8 /*warps*/ * iters * 8 * 16;       // warp instructions per SM
        const double clk = ms * 1e-3 * 1.965e9;
        printf("%-28s %d CTA/SM x 8 warps: %.3f warp-instr/clk/SM, %.1f fp32 fma lanes/clk/SM (%s)\n", name, ctas, instr / clk,
               instr / clk * 32 * per_instr, cudaGetErrorString(cudaGetLastError()));
    }
}

int main() {
    run<0>("FFMA reg x const", 1);
    run<1>("FFMA 3 regs", 1);
    run<2>("FFMA2 3 reg pairs", 2);
    run<3>("FFMA2 pair x const", 2);
    return 0;
}
