// Context lifetime, memory and timing entry points of the C ABI.
#include "common.cuh"

extern "C" {

int hb_abi_version(void) { return HB_ABI_VERSION; }

int hb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
        return 0;
    return n;
}

int hb_create(int device, hb_ctx **out)
{
    if (!out)
        return HB_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
        return HB_ERR_CUDA;   // no silent CPU fallback: without a device there is no backend
    if (device < 0 || device >= n)
        return HB_ERR_ARG;
    hb_ctx *ctx = new hb_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev_start) != cudaSuccess || cudaEventCreate(&ctx->ev_stop) != cudaSuccess) {
        delete ctx;
        return HB_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = ctx;
    return HB_OK;
}

void hb_destroy(hb_ctx *ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (void *p : ctx->allocs)
        cudaFree(p);
    for (auto &kv : ctx->scale_tabs)
        cudaFree(kv.second);
    if (ctx->bicubic_dev)
        cudaFree(ctx->bicubic_dev);
    cudaEventDestroy(ctx->ev_start);
    cudaEventDestroy(ctx->ev_stop);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *hb_last_error(const hb_ctx *ctx) { return ctx ? ctx->err : "null context"; }

int hb_sync(hb_ctx *ctx)
{
    HB_ARG(ctx, ctx != nullptr);
    HB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}

uint64_t hb_stream(const hb_ctx *ctx) { return ctx ? (uint64_t)(uintptr_t)ctx->stream : 0; }
uint64_t hb_launch_count(const hb_ctx *ctx) { return ctx ? ctx->launches : 0; }

int hb_alloc(hb_ctx *ctx, size_t bytes, hb_devptr *out)
{
    HB_ARG(ctx, ctx && out);
    void *p = nullptr;
    HB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess)
        return hb_fail(ctx, HB_ERR_NOMEM, "cudaMalloc: %s", cudaGetErrorString(e));
    ctx->allocs.push_back(p);
    *out = (hb_devptr)(uintptr_t)p;
    return HB_OK;
}

int hb_free(hb_ctx *ctx, hb_devptr p)
{
    HB_ARG(ctx, ctx != nullptr);
    for (size_t i = 0; i < ctx->allocs.size(); i++)
        if (ctx->allocs[i] == (void *)(uintptr_t)p) {
            cudaStreamSynchronize(ctx->stream);
            cudaFree(ctx->allocs[i]);
            ctx->allocs.erase(ctx->allocs.begin() + i);
            return HB_OK;
        }
    return hb_fail(ctx, HB_ERR_ARG, "hb_free: %s", "pointer not owned by this context");
}

int hb_host_alloc(size_t bytes, void **out)
{
    if (!out) return HB_ERR_ARG;
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return HB_ERR_NOMEM;
    }
    *out = p;
    return HB_OK;
}

int hb_host_free(void *p)
{
    if (p && cudaFreeHost(p) != cudaSuccess) {
        cudaGetLastError();
        return HB_ERR_CUDA;
    }
    return HB_OK;
}

int hb_upload(hb_ctx *ctx, hb_devptr dst, const void *src, size_t bytes)
{
    HB_ARG(ctx, ctx && src);
    HB_CUDA(ctx, cudaMemcpyAsync((void *)(uintptr_t)dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return HB_OK;
}

int hb_download(hb_ctx *ctx, void *dst, hb_devptr src, size_t bytes)
{
    HB_ARG(ctx, ctx && dst);
    HB_CUDA(ctx, cudaMemcpyAsync(dst, (const void *)(uintptr_t)src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return HB_OK;
}

int hb_timer_start(hb_ctx *ctx)
{
    HB_ARG(ctx, ctx != nullptr);
    HB_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    return HB_OK;
}

int hb_timer_stop(hb_ctx *ctx, float *ms)
{
    HB_ARG(ctx, ctx && ms);
    HB_CUDA(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
    HB_CUDA(ctx, cudaEventSynchronize(ctx->ev_stop));
    HB_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev_start, ctx->ev_stop));
    return HB_OK;
}

}  // extern "C"
