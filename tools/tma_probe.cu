// Stand-alone probe of the TMA window load used by k_inter (debug aid): nvcc -gencode arch=compute_100a,code=sm_100a tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

template <int MODE>
__global__ void probe(const __grid_constant__ CUtensorMap map, const CUtensorMap *gmap, int x, int y, uint16_t *out, int *status)
{
    __shared__ __align__(128) uint16_t win[28][32];
    __shared__ unsigned long long mbar;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar), dst = (uint32_t)__cvta_generic_to_shared(&win[0][0]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        if (MODE == 0) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
        } else {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(28 * 32 * 2) : "memory");
            if (MODE == 4) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(out + 4096)), "r"(28 * 32 * 2), "r"(bar) : "memory");
            }
            const uint64_t d = MODE == 1 ? reinterpret_cast<uint64_t>(&map) : reinterpret_cast<uint64_t>(gmap);
            if (MODE == 4 || MODE == 5) {
            } else if (MODE == 3)
                asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(x), "r"(y), "r"(bar) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(dst), "l"(d), "r"(x), "r"(y), "r"(bar) : "memory");
        }
    }
    uint32_t done = 0;
    unsigned spins = 0;
    while (!done && spins < (1u << 22)) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(0) : "memory");
        spins++;
    }
    if (threadIdx.x == 0) { status[0] = done; status[1] = spins; }
    __syncwarp();
    for (int i = threadIdx.x; i < 28 * 32; i += 32) out[i] = win[i / 32][i % 32];
}

#include <stdlib.h>
int main(int argc, char **argv)
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
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, fn);
    CUtensorMap map;
    const int var = argc > 2 ? atoi(argv[2]) : 0;
    const int X = argc > 3 ? atoi(argv[3]) : 37;
    cuuint64_t dims[2] = {W, H}; const cuuint64_t strides[1] = {W * 2}; cuuint32_t box[2] = {32, 28}; const cuuint32_t es[2] = {1, 1};
    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_UINT16;
    CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (var == 1) l2 = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (var == 2) { dt = CU_TENSOR_MAP_DATA_TYPE_UINT8; dims[0] = W * 2; box[0] = 64; }
    if (var == 3) { dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; }
    if (var == 4) { dt = CU_TENSOR_MAP_DATA_TYPE_UINT32; dims[0] = W / 2; box[0] = 16; }
    CUresult rc = ((EncodeTiled)fn)(&map, dt, 2, dev, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)rc);
    CUtensorMap *gmap; cudaMalloc(&gmap, sizeof(map)); cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice);
    const int mode = argc > 1 ? atoi(argv[1]) : 1;
    if (mode == 0) probe<0><<<1, 32>>>(map, gmap, X, 11, out, status);
    else if (mode == 1) probe<1><<<1, 32>>>(map, gmap, X, 11, out, status);
    else if (mode == 2) probe<2><<<1, 32>>>(map, gmap, X, 11, out, status);
    else if (mode == 3) probe<3><<<1, 32>>>(map, gmap, X, 11, out, status);
    else if (mode == 4) probe<4><<<1, 32>>>(map, gmap, X, 11, out, status);
    else probe<5><<<1, 32>>>(map, gmap, X, 11, out, status);
    e = cudaDeviceSynchronize();
    printf("mode %d kernel: %s\n", mode, cudaGetErrorString(e));
    int st[2]; std::vector<uint16_t> o(28 * 32);
    cudaMemcpy(st, status, 8, cudaMemcpyDeviceToHost); cudaMemcpy(o.data(), out, 28 * 32 * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 28; r++) for (int c = 0; c < 32; c++) bad += o[r * 32 + c] != host[(11 + r) * W + X + c];
    printf("done=%d spins=%d mismatches=%d first=%d expect=%d\n", st[0], st[1], bad, o[0], host[11 * W + X]);
    return 0;
}
