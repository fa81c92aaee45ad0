// Microbenchmark: shared-memory / L2 atomic increment rates on sm_100a for the Hough vote (deskew.cu).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o atoms_probe atoms_probe.cu ; prints increments per clock per SM.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int WORDS = 21860;   // four accumulator rows of a 1600 x 1131 page

// PAT 0: consecutive addresses per lane (conflict-free), 1: pseudo-random, 2: one address per warp, 3: window of 23 cells
// OP  0: atomicAdd(.., 1) (ATOMS.POPC.INC), 1: atomicAdd(.., v) (ATOMS.ADD), 2: plain load + store, 3: global atomicAdd (RED)
template <int PAT, int OP>
__global__ void __launch_bounds__(512) k(int* gacc, int iters, int one) {
    extern __shared__ int s[];
    for (int i = threadIdx.x; i < WORDS; i += 512) s[i] = 0;
    __syncthreads();
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    int* g = gacc + (size_t)(blockIdx.x % 296) * WORDS;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x = x * 1664525u + 1013904223u;
            int idx;
            if (PAT == 0) idx = (wrp * 997 + it * 8 + u) % (WORDS - 32) + lane;
            else if (PAT == 1) idx = (x >> 8) % WORDS;
            else if (PAT == 2) idx = (wrp * 997 + it * 8 + u) % WORDS;
            else idx = (wrp * 997 + (it * 8 + u) * 3) % (WORDS - 32) + (x >> 10) % 23;
            if (OP == 0) atomicAdd(&s[idx], 1);
            if (OP == 1) atomicAdd(&s[idx], one);
            if (OP == 2) s[idx] = s[idx] + 1;
            if (OP == 3) atomicAdd(&g[idx], 1);
        }
    }
    __syncthreads();
    int sum = 0;
    for (int i = threadIdx.x; i < WORDS; i += 512) sum += s[i];
    if (sum == 0x7fffffff) gacc[0] = sum;
}

template <int PAT, int OP>
void run(const char* name, int* gacc) {
    cudaFuncSetAttribute(k<PAT, OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000, ctas = 296;
    k<PAT, OP><<<ctas, 512, WORDS * 4>>>(gacc, 10, 1);
    cudaEventRecord(e0);
    k<PAT, OP><<<ctas, 512, WORDS * 4>>>(gacc, iters, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double incr = 2.0 * 512 * iters * 8;           // per SM (2 CTAs)
    printf("%-44s %.2f increments/clk/SM  (%s)\n", name, incr / (ms * 1e-3 * 1.965e9), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int* gacc; cudaMalloc(&gacc, (size_t)296 * WORDS * 4); cudaMemset(gacc, 0, (size_t)296 * WORDS * 4);
    run<0, 0>("POPC.INC consecutive", gacc);
    run<1, 0>("POPC.INC random", gacc);
    run<2, 0>("POPC.INC one address per warp", gacc);
    run<3, 0>("POPC.INC window of 23", gacc);
    run<0, 1>("ATOMS.ADD consecutive", gacc);
    run<1, 1>("ATOMS.ADD random", gacc);
    run<2, 1>("ATOMS.ADD one address per warp", gacc);
    run<3, 1>("ATOMS.ADD window of 23", gacc);
    run<0, 2>("LDS+STS consecutive", gacc);
    run<1, 2>("LDS+STS random", gacc);
    run<0, 3>("global RED consecutive", gacc);
    run<1, 3>("global RED random", gacc);
    run<3, 3>("global RED window of 23", gacc);
    return 0;
}
