// Third probe of the tiled TMA load: the CUDA programming guide's own example form (cuda::barrier +
// cuda::device::experimental::cp_async_bulk_tensor_2d_global_to_shared), cuTensorMapEncodeTiled fetched through the
// runtime's driver entry point.  nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe3 tma_probe3.cu
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

constexpr int BW = 32, BH = 28;

__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, uint16_t *out)
{
    __shared__ alignas(128) uint16_t win[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) {
        init(&bar, blockDim.x);
        cde::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&win, &map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(win));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BH * BW; i += blockDim.x) out[i] = win[i / BW][i % BW];
}

int main(int argc, char **argv)
{
    const int X = argc > 1 ? atoi(argv[1]) : 40, T = argc > 2 ? atoi(argv[2]) : 128, order = argc > 3 ? atoi(argv[3]) : 0;
    (void)order;
    const int W = 416, H = 300;
    std::vector<uint16_t> host(W * H);
    for (int i = 0; i < W * H; i++) host[i] = (uint16_t)(i % 1000);
    uint16_t *dev, *out;
    cudaMalloc(&dev, W * H * 2); cudaMalloc(&out, 65536);
    cudaMemcpy(dev, host.data(), W * H * 2, cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q);
    printf("entry point (by version 12000): %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, fn);
    int drv = 0, rt = 0; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("driver %d runtime %d device %s cc %d.%d\n", drv, rt, prop.name, prop.major, prop.minor);
    alignas(64) CUtensorMap map;
    const cuuint64_t dims[2] = {W, H}; const cuuint64_t strides[1] = {W * 2}; const cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
    CUresult rc = ((EncodeTiled)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dev, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)rc);
    probe<<<1, T>>>(map, X, 11, out);
    printf("x=%d threads=%d\n", X, T);
    e = cudaDeviceSynchronize();
    printf("guide-form kernel: %s\n", cudaGetErrorString(e));
    std::vector<uint16_t> o(BH * BW);
    cudaMemcpy(o.data(), out, BH * BW * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < BH; r++) for (int c = 0; c < BW; c++) bad += o[r * BW + c] != host[(11 + r) * W + X + c];
    printf("mismatches=%d first=%d expect=%d\n", bad, o[0], host[11 * W + X]);
    return 0;
}
