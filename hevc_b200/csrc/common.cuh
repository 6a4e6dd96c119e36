// Shared host-side plumbing for the hevc_b200 C ABI (see include/hevc_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/hevc_b200.h"

struct hb_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    uint64_t launches = 0;
    char err[512] = {0};
    std::vector<void *> allocs;
    // scaler tap-position tables keyed by (src << 32 | dst): device int2 {first tap index, phase}
    std::map<uint64_t, int2 *> scale_tabs;
    short *bicubic_dev = nullptr;
    std::mutex mu;
};

static inline int hb_fail(hb_ctx *ctx, int code, const char *fmt, const char *detail)
{
    if (ctx)
        snprintf(ctx->err, sizeof(ctx->err), fmt, detail);
    return code;
}

#define HB_CUDA(ctx, call)                                                                  \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return hb_fail((ctx), HB_ERR_CUDA, #call ": %s", cudaGetErrorString(e_));       \
    } while (0)

#define HB_ARG(ctx, cond)                                                                   \
    do {                                                                                    \
        if (!(cond))                                                                        \
            return hb_fail((ctx), HB_ERR_ARG, "bad argument: %s", #cond);                   \
    } while (0)

// after a <<<>>> launch: count it and surface launch-configuration errors
#define HB_LAUNCHED(ctx)                                                                    \
    do {                                                                                    \
        (ctx)->launches++;                                                                  \
        HB_CUDA((ctx), cudaGetLastError());                                                 \
    } while (0)

static inline int hb_grid_for(const hb_ctx *ctx, long long work_items, int block, int ctas_per_sm)
{
    long long need = (work_items + block - 1) / block;
    long long cap = (long long)ctx->sm_count * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}
