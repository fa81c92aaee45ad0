// Stand-alone probe of the TMA tile load used by csrc/tcblur.cu (developer tool; run on a B200):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I smart_image_processing_b200/csrc tests/tools/tma_probe.cu -o gpurun_out/tma_probe
//   ./tma_probe <mode>   mode 0: tensor map as __grid_constant__ parameter, 1: tensor map in global memory + proxy fence
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc05.cuh"

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe_kernel(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, int use_global, int rows, int c0, int c1,
                             uint8_t* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tile = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_init_fence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (use_global == 10) {            // mbarrier only
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&bar)) : "memory");
        } else if (use_global == 11) {     // 1-D bulk copy
            tc::mbar_expect_tx(&bar, rows * 128);
            tc::bulk_load(tile, out + 65536, rows * 128, &bar);
        } else if (use_global == 1) {
            tc::tmap_acquire(gmap);
            tc::mbar_expect_tx(&bar, rows * 128);
            tc::tma_load_2d(tile, gmap, c0, c1, &bar);
        } else {
            tc::mbar_expect_tx(&bar, rows * 128);
            tc::tma_load_2d(tile, &pmap, c0, c1, &bar);
        }
    }
    if (!tc::mbar_wait_bounded(&bar, 0)) { if (threadIdx.x == 0) printf("timeout\n"); return; }
    for (int i = threadIdx.x; i < rows * 128; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int rows = argc > 2 ? atoi(argv[2]) : 160;
    const int variant = argc > 4 ? atoi(argv[4]) : 0;     // 0: 260x300 coords (-10,-10); 1: same, coords (0,0); 2: 256x256 pitch 256 coords (0,0); 3: 260x300 coords (16,16)
    const int W = variant == 2 ? 256 : 260, H = variant == 2 ? 256 : 300, pitch = variant == 2 ? 256 : 384;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no encode fn\n"); return 2; }
    std::vector<uint8_t> img((size_t)pitch * H);
    for (int y = 0; y < H; y++) for (int x = 0; x < pitch; x++) img[(size_t)y * pitch + x] = (uint8_t)(x < W ? (y * 7 + x * 3 + 1) : 0xEE);
    uint8_t *dimg, *dout; CUtensorMap* dmap;
    cudaMalloc(&dimg, img.size()); cudaMalloc(&dout, 65536 + 256 * 128); cudaMemset(dout, 7, 65536 + 256 * 128); cudaMalloc(&dmap, sizeof(CUtensorMap));
    cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice);
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch};
    const int swz = argc > 3 ? atoi(argv[3]) : 1;
    const cuuint32_t box[2] = {128, (cuuint32_t)rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = ((EncodeTiledFn)p)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dimg, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("mode %d rows %d swz %d variant %d encode=%d\n", mode, rows, swz, variant, (int)cr);
    { const uint32_t* u = (const uint32_t*)&map; printf("  map:"); for (int i = 0; i < 16; i++) printf(" %08x", u[i]); printf("\n"); }
    if (cr != CUDA_SUCCESS) return 3;
    cudaMemcpy(dmap, &map, sizeof(map), cudaMemcpyHostToDevice);
    const int c0 = variant == 0 ? -10 : (variant == 3 ? 16 : 0), c1 = c0;
    const size_t smem = 1024 + rows * 128;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<<<1, 128, smem>>>(map, dmap, mode, rows, c0, c1, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    if (mode >= 10) return 0;
    std::vector<uint8_t> out(rows * 128);
    cudaMemcpy(out.data(), dout, out.size(), cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int j = 0; j < rows; j++)
        for (int c = 0; c < 128; c++) {
            const int y = c1 + j, x = c0 + c;
            const uint8_t want = (y >= 0 && y < H && x >= 0 && x < W) ? img[(size_t)y * pitch + x] : 0;
            const uint8_t got = out[(size_t)j * 128 + (((c >> 4) ^ (swz ? (j & 7) : 0)) << 4) + (c & 15)];
            if (got != want && bad++ < 5) printf("  (%d,%d) got %d want %d\n", j, c, got, want);
        }
    printf("swizzled-layout mismatches: %ld\n", bad);
    return 0;
}
