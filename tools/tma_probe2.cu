#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#include <cute/arch/copy_sm90_tma.hpp>
#include <cutlass/arch/barrier.h>

__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, uint16_t *out, int *status)
{
    __shared__ __align__(128) uint16_t win[28][32];
    __shared__ uint64_t mbar;
    using Barrier = cutlass::arch::ClusterTransactionBarrier;
    if (threadIdx.x == 0) {
        Barrier::init(&mbar, 1);
        cutlass::arch::fence_barrier_init();
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        Barrier::arrive_and_expect_tx(&mbar, 28 * 32 * 2);
        cute::SM90_TMA_LOAD_2D::copy(&map, &mbar, 0ull /*cache hint*/, &win[0][0], x, y);
    }
    Barrier::wait(&mbar, 0);
    if (threadIdx.x == 0) status[0] = 1;
    __syncwarp();
    for (int i = threadIdx.x; i < 28 * 32; i += 32) out[i] = win[i / 32][i % 32];
}

int main()
{
    const int W = 416, H = 300;
    std::vector<uint16_t> host(W * H);
    for (int i = 0; i < W * H; i++) host[i] = (uint16_t)(i % 1000);
    uint16_t *dev, *out; int *status;
    cudaMalloc(&dev, W * H * 2); cudaMalloc(&out, 65536); cudaMalloc(&status, 8);
    cudaMemcpy(dev, host.data(), W * H * 2, cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap map;
    const cuuint64_t dims[2] = {W, H}; const cuuint64_t strides[1] = {W * 2}; const cuuint32_t box[2] = {32, 28}, es[2] = {1, 1};
    CUresult rc = ((EncodeTiled)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dev, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)rc);
    probe<<<1, 32>>>(map, 37, 11, out, status);
    cudaError_t e = cudaDeviceSynchronize();
    printf("cutlass-path kernel: %s\n", cudaGetErrorString(e));
    std::vector<uint16_t> o(28 * 32);
    cudaMemcpy(o.data(), out, 28 * 32 * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 28; r++) for (int c = 0; c < 32; c++) bad += o[r * 32 + c] != host[(11 + r) * W + 37 + c];
    printf("mismatches=%d first=%d expect=%d\n", bad, o[0], host[11 * W + 37]);
    return 0;
}
