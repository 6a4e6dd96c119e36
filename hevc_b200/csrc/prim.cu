// Batched encoder primitives (BASELINE config 5): SAD / SATD / SA8D / DCT / DST / quant / dequant / intra,
// bit-exact against oracle/primitives.c.  Integer-ALU bound; warp-shuffle reductions, per-thread register
// butterflies (prim_dev.cuh), dense [n][H][W] block layout so every load is a full-width vector.
#include "common.cuh"
#include "prim_dev.cuh"

using namespace hb;

namespace {

__device__ __forceinline__ int group_sum(int v, int width)
{
    for (int off = width >> 1; off > 0; off >>= 1)
        v += __shfl_down_sync(0xffffffffu, v, off, width);
    return v;
}

// ------------------------------------------------------------------------------------------------ SAD
// T lanes per block (T = min(32, W*H/8)), each lane strides over 8-sample (16-byte) vectors
__global__ void __launch_bounds__(256) k_sad(const pixel *__restrict__ a, const pixel *__restrict__ b, int n, int wh, int T, int *__restrict__ out)
{
    const int lane_in = threadIdx.x % T;
    const long long first = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / T;
    const long long stride = (long long)gridDim.x * blockDim.x / T;
    const long long rounds = (n + stride - 1) / stride;   // uniform trip count keeps shuffles convergent
    const int vecs = wh >> 3;
    for (long long r = 0; r < rounds; r++) {
        const long long blk = first + r * stride;
        int s = 0;
        if (blk < n) {
            const uint4 *pa = reinterpret_cast<const uint4 *>(a + blk * wh), *pb = reinterpret_cast<const uint4 *>(b + blk * wh);
            for (int v = lane_in; v < vecs; v += T) {
                const uint4 x = __ldg(pa + v), y = __ldg(pb + v);
                const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    s += abs((int)(xs[k] & 0xffff) - (int)(ys[k] & 0xffff));
                    s += abs((int)(xs[k] >> 16) - (int)(ys[k] >> 16));
                }
            }
        }
        s = group_sum(s, T);
        if (blk < n && lane_in == 0)
            out[blk] = s;
    }
}

// x265 sad_x3 / sad_x4: one source block against R (3 or 4) reference blocks; the source vectors are loaded once
template <int R>
__global__ void __launch_bounds__(256) k_sad_multi(const pixel *__restrict__ a, const pixel *__restrict__ b0, const pixel *__restrict__ b1,
                                                   const pixel *__restrict__ b2, const pixel *__restrict__ b3, int n, int wh, int T,
                                                   int *__restrict__ out)
{
    const int lane_in = threadIdx.x % T;
    const long long first = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / T;
    const long long stride = (long long)gridDim.x * blockDim.x / T;
    const long long rounds = (n + stride - 1) / stride;
    const int vecs = wh >> 3;
    const pixel *refs[4] = {b0, b1, b2, b3};
    for (long long r = 0; r < rounds; r++) {
        const long long blk = first + r * stride;
        int s[R];
#pragma unroll
        for (int k = 0; k < R; k++) s[k] = 0;
        if (blk < n) {
            const uint4 *pa = reinterpret_cast<const uint4 *>(a + blk * wh);
            for (int v = lane_in; v < vecs; v += T) {
                const uint4 x = __ldg(pa + v);
                const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int k = 0; k < R; k++) {
                    const uint4 y = __ldg(reinterpret_cast<const uint4 *>(refs[k] + blk * wh) + v);
                    const uint32_t ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        s[k] += abs((int)(xs[j] & 0xffff) - (int)(ys[j] & 0xffff));
                        s[k] += abs((int)(xs[j] >> 16) - (int)(ys[j] >> 16));
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < R; k++) {
            const int t = group_sum(s[k], T);
            if (blk < n && lane_in == 0) out[blk * R + k] = t;
        }
    }
}

// ------------------------------------------------------------------------------------------------ SATD
// one lane per 4x4 sub-block; SATD(WxH) = sum over sub-blocks of raw/2 (raw sums are always even, so the
// x265 8x4 pairing "(raw0 + raw1) >> 1" gives the same value)
__global__ void __launch_bounds__(256) k_satd(const pixel *__restrict__ a, const pixel *__restrict__ b, int n, int w, int h, int T, int *__restrict__ out)
{
    const int lane_in = threadIdx.x % T;
    const long long first = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / T;
    const long long stride = (long long)gridDim.x * blockDim.x / T;
    const long long rounds = (n + stride - 1) / stride;
    const int sw = w >> 2, nsub = sw * (h >> 2), wh = w * h;
    for (long long r = 0; r < rounds; r++) {
        const long long blk = first + r * stride;
        int s = 0;
        if (blk < n) {
            const pixel *pa = a + blk * wh, *pb = b + blk * wh;
            for (int sb = lane_in; sb < nsub; sb += T) {
                const int ox = (sb % sw) << 2, oy = (sb / sw) << 2;
                int d[4][4];
#pragma unroll
                for (int y = 0; y < 4; y++) {
                    const uint2 x = __ldg(reinterpret_cast<const uint2 *>(pa + (oy + y) * w + ox));
                    const uint2 z = __ldg(reinterpret_cast<const uint2 *>(pb + (oy + y) * w + ox));
                    d[y][0] = (int)(x.x & 0xffff) - (int)(z.x & 0xffff);
                    d[y][1] = (int)(x.x >> 16) - (int)(z.x >> 16);
                    d[y][2] = (int)(x.y & 0xffff) - (int)(z.y & 0xffff);
                    d[y][3] = (int)(x.y >> 16) - (int)(z.y >> 16);
                }
                s += hadamard4x4_abs(d) >> 1;
            }
        }
        s = group_sum(s, T);
        if (blk < n && lane_in == 0)
            out[blk] = s;
    }
}

// ------------------------------------------------------------------------------------------------ SA8D
// one lane per 8x8 sub-block, sub-blocks numbered so that 4 consecutive lanes form one 16x16 tile
__global__ void __launch_bounds__(128) k_sa8d(const pixel *__restrict__ a, const pixel *__restrict__ b, int n, int size, int T, int *__restrict__ out)
{
    const int lane_in = threadIdx.x % T;
    const long long first = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / T;
    const long long stride = (long long)gridDim.x * blockDim.x / T;
    const long long rounds = (n + stride - 1) / stride;
    const int tiles_w = size >> 4, nsub = (size >> 3) * (size >> 3), wh = size * size;
    for (long long r = 0; r < rounds; r++) {
        const long long blk = first + r * stride;
        int total = 0;
        for (int sb0 = 0; sb0 < nsub; sb0 += T) {
            const int sb = sb0 + lane_in;
            int raw = 0;
            if (blk < n && sb < nsub) {
                int ox, oy;
                if (size == 8) { ox = oy = 0; }
                else {
                    const int tile = sb >> 2, q = sb & 3;
                    ox = ((tile % tiles_w) << 4) + ((q & 1) << 3);
                    oy = ((tile / tiles_w) << 4) + ((q >> 1) << 3);
                }
                const pixel *pa = a + blk * wh + oy * size + ox, *pb = b + blk * wh + oy * size + ox;
                int m[8][8];
#pragma unroll
                for (int y = 0; y < 8; y++) {
                    const uint4 x = __ldg(reinterpret_cast<const uint4 *>(pa + y * size));
                    const uint4 z = __ldg(reinterpret_cast<const uint4 *>(pb + y * size));
                    const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, zs[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        m[y][2 * k] = (int)(xs[k] & 0xffff) - (int)(zs[k] & 0xffff);
                        m[y][2 * k + 1] = (int)(xs[k] >> 16) - (int)(zs[k] >> 16);
                    }
                }
                raw = hadamard8x8_abs(m);
            }
            if (size == 8) {
                total += (raw + 2) >> 2;
            } else {   // (sum of the four raw 8x8 of a 16x16 tile + 2) >> 2, accumulated by the tile's first lane
                int t4 = raw + __shfl_down_sync(0xffffffffu, raw, 1, 4);
                t4 += __shfl_down_sync(0xffffffffu, t4, 2, 4);
                if ((lane_in & 3) == 0 && sb < nsub)
                    total += (t4 + 2) >> 2;
            }
        }
        total = group_sum(total, T);
        if (blk < n && lane_in == 0)
            out[blk] = total;
    }
}

// ------------------------------------------------------------------------------------------------ transforms
// lane = (block-in-warp, line).  Pass 1 reads the line straight from global memory, writes transposed into
// shared memory (padded rows); pass 2 reads a padded row, writes the coefficient row to global memory.
template <int N, bool DST, bool INVERSE>
__global__ void __launch_bounds__(128) k_transform(const int16_t *__restrict__ src, int16_t *__restrict__ dst, int n, int bit_depth)
{
    constexpr int PER_WARP = 32 / N, PAD = N + 2, L2 = ilog2c(N);
    __shared__ int16_t tile[4][PER_WARP][N][PAD];
    __shared__ int16_t otile[INVERSE ? 4 : 1][INVERSE ? PER_WARP : 1][INVERSE ? N : 1][INVERSE ? PAD : 2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / N, line = lane % N;
    const long long warps = (long long)gridDim.x * 4;
    const int shift1 = INVERSE ? 7 : L2 - 1 + (bit_depth - 8);
    const int shift2 = INVERSE ? 12 - (bit_depth - 8) : L2 + 6;
    for (long long base = (blockIdx.x * 4LL + warp) * PER_WARP; base < n; base += warps * PER_WARP) {
        const long long blk = base + sub;
        int16_t(*t)[PAD] = tile[warp][sub];
        if (blk < n) {
            if (!INVERSE)   // rows of the residual; output k of row `line` lands at t[k][line]
                fwd_line<N, DST>(src + blk * N * N + line * N, 1, &t[0][line], PAD, shift1);
            else            // first inverse stage works on columns of the coefficient block
                inv_line<N, DST>(src + blk * N * N + line, N, &t[line][0], 1, shift1);
        }
        __syncwarp();
        if constexpr (!INVERSE) {
            if (blk < n) fwd_line<N, DST>(&t[line][0], 1, dst + blk * N * N + line, N, shift2);
            __syncwarp();
        } else {
            // second inverse stage produces residual ROWS per thread: stage them in shared memory (second tile) and let the
            // whole warp write the PER_WARP blocks back as contiguous 32-bit words instead of 2-byte strided stores
            int16_t(*o)[PAD] = otile[warp][sub];
            if (blk < n) inv_line<N, DST>(&t[0][line], PAD, &o[line][0], 1, shift2);
            __syncwarp();
            const long long remaining = (long long)n - base;
            const int nblk = remaining < PER_WARP ? (int)remaining : PER_WARP;
            uint32_t *g32 = reinterpret_cast<uint32_t *>(dst + base * N * N);
            for (int w = lane; w < nblk * N * N / 2; w += 32) {
                const int e = 2 * w, b = e / (N * N), r = (e % (N * N)) / N, c = e % N;
                g32[w] = (uint32_t)(uint16_t)otile[warp][b][r][c] | ((uint32_t)(uint16_t)otile[warp][b][r][c + 1] << 16);
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------ quant / dequant
__global__ void __launch_bounds__(256) k_quant(const int16_t *__restrict__ coef, int16_t *__restrict__ level, int *__restrict__ numsig,
                                               int n, int nn, QuantParam q)
{
    // one warp per block of nn coefficients (nn >= 16): 2 coefficients per lane per step
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long blk = warp0; blk < n; blk += warps) {
        const uint32_t *c2 = reinterpret_cast<const uint32_t *>(coef + blk * nn);
        uint32_t *l2 = reinterpret_cast<uint32_t *>(level + blk * nn);
        int cnt = 0;
        for (int i = lane; i < nn / 2; i += 32) {
            const uint32_t p = __ldg(c2 + i);
            const int a = quant_one((int16_t)(p & 0xffff), q), b = quant_one((int16_t)(p >> 16), q);
            cnt += (a != 0) + (b != 0);
            l2[i] = (uint32_t)(uint16_t)a | ((uint32_t)(uint16_t)b << 16);
        }
        cnt = group_sum(cnt, 32);
        if (lane == 0)
            numsig[blk] = cnt;
    }
}

__global__ void __launch_bounds__(256) k_dequant(const int16_t *__restrict__ level, int16_t *__restrict__ coef, long long total2, QuantParam q)
{
    const uint32_t *l2 = reinterpret_cast<const uint32_t *>(level);
    uint32_t *c2 = reinterpret_cast<uint32_t *>(coef);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total2; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t p = __ldg(l2 + i);
        const int a = dequant_one((int16_t)(p & 0xffff), q), b = dequant_one((int16_t)(p >> 16), q);
        c2[i] = (uint32_t)(uint16_t)a | ((uint32_t)(uint16_t)b << 16);
    }
}

// ------------------------------------------------------------------------------------------------ intra, all 35 modes
// one warp per block: neighbours (raw + smoothed) staged in shared memory, lanes sweep mode x pixel
__global__ void __launch_bounds__(128) k_intra_all(const pixel *__restrict__ nbs, pixel *__restrict__ pred, int n, int size, int is_luma,
                                                   int strong, int bit_depth)
{
    __shared__ pixel raw[4][132], flt[4][132];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cnt = 4 * size + 1, l2 = 31 - __clz(size), nn = size * size, maxv = (1 << bit_depth) - 1;
    for (long long blk = blockIdx.x * 4LL + warp; blk < n; blk += gridDim.x * 4LL) {
        const pixel *nb = nbs + blk * cnt;
        for (int i = lane; i < cnt; i += 32)
            raw[warp][i] = nb[i];
        __syncwarp();
        bool bilinear = false;
        if (strong && size == 32) {
            const int tl = raw[warp][0], tr = raw[warp][64], bl = raw[warp][128], thr = 1 << (bit_depth - 5);
            bilinear = abs(tl + tr - 2 * raw[warp][32]) < thr && abs(tl + bl - 2 * raw[warp][96]) < thr;
        }
        for (int i = lane; i < cnt; i += 32) {
            int v;
            if (bilinear) {
                const int tl = raw[warp][0];
                if (i == 0 || i == 64 || i == 128) v = raw[warp][i];
                else if (i < 64) v = ((64 - i) * tl + i * raw[warp][64] + 32) >> 6;
                else v = ((64 - (i - 64)) * tl + (i - 64) * raw[warp][128] + 32) >> 6;
            } else {
                v = intra_filtered(raw[warp], size, i);
            }
            flt[warp][i] = (pixel)v;
        }
        __syncwarp();
        int dcs = 0;
        for (int i = lane; i < size; i += 32)
            dcs += raw[warp][1 + i] + raw[warp][1 + 2 * size + i];
        dcs = group_sum(dcs, 32);
        dcs = __shfl_sync(0xffffffffu, dcs, 0);
        const int dc = (dcs + size) >> (l2 + 1);
        pixel *out = pred + blk * 35 * nn;
        for (int mode = 0; mode < 35; mode++) {
            const bool f = is_luma && intra_use_filter(l2, mode);
            const pixel *src = f ? flt[warp] : raw[warp];
            for (int p = lane; p < nn; p += 32) {
                const int x = p & (size - 1), y = p >> l2;
                out[mode * nn + p] = (pixel)intra_sample(src, size, l2, mode, x, y, is_luma != 0, maxv, dc);
            }
        }
        __syncwarp();
    }
}

}  // namespace

extern "C" {

int hb_sad(hb_ctx *ctx, hb_devptr a, hb_devptr b, int n, int w, int h, hb_devptr out)
{
    HB_ARG(ctx, ctx && a && b && out && n >= 0 && w >= 4 && h >= 4 && (w * h) % 8 == 0 && ((a | b) & 15) == 0);
    if (!n) return HB_OK;
    int T = w * h / 8;
    T = T > 32 ? 32 : T;
    while (T & (T - 1)) T &= T - 1;   // power of two for the shuffle width
    k_sad<<<hb_grid_for(ctx, (long long)n * T, 256, 8), 256, 0, ctx->stream>>>((const pixel *)a, (const pixel *)b, n, w * h, T, (int *)out);
    HB_LAUNCHED(ctx);
    return HB_OK;
}

int hb_sad_multi(hb_ctx *ctx, hb_devptr fenc, hb_devptr ref0, hb_devptr ref1, hb_devptr ref2, hb_devptr ref3, int n_refs, int n, int w, int h,
                 hb_devptr out)
{
    HB_ARG(ctx, ctx && fenc && ref0 && ref1 && ref2 && out && (n_refs == 3 || (n_refs == 4 && ref3)) && n >= 0 && w >= 4 && h >= 4 && (w * h) % 8 == 0);
    HB_ARG(ctx, ((fenc | ref0 | ref1 | ref2 | (n_refs == 4 ? ref3 : 0)) & 15) == 0);
    if (!n) return HB_OK;
    int T = w * h / 8;
    T = T > 32 ? 32 : T;
    while (T & (T - 1)) T &= T - 1;
    const int grid = hb_grid_for(ctx, (long long)n * T, 256, 8);
    if (n_refs == 3)
        k_sad_multi<3><<<grid, 256, 0, ctx->stream>>>((const pixel *)fenc, (const pixel *)ref0, (const pixel *)ref1, (const pixel *)ref2, nullptr, n, w * h, T,
                                                      (int *)out);
    else
        k_sad_multi<4><<<grid, 256, 0, ctx->stream>>>((const pixel *)fenc, (const pixel *)ref0, (const pixel *)ref1, (const pixel *)ref2, (const pixel *)ref3, n,
                                                      w * h, T, (int *)out);
    HB_LAUNCHED(ctx);
    return HB_OK;
}

int hb_satd(hb_ctx *ctx, hb_devptr a, hb_devptr b, int n, int w, int h, hb_devptr out)
{
    HB_ARG(ctx, ctx && a && b && out && n >= 0 && w >= 4 && h >= 4 && (w % 4) == 0 && (h % 4) == 0 && ((a | b) & 7) == 0);
    if (!n) return HB_OK;
    int T = (w / 4) * (h / 4);
    T = T > 32 ? 32 : T;
    while (T & (T - 1)) T &= T - 1;
    k_satd<<<hb_grid_for(ctx, (long long)n * T, 256, 8), 256, 0, ctx->stream>>>((const pixel *)a, (const pixel *)b, n, w, h, T, (int *)out);
    HB_LAUNCHED(ctx);
    return HB_OK;
}

int hb_sa8d(hb_ctx *ctx, hb_devptr a, hb_devptr b, int n, int size, hb_devptr out)
{
    HB_ARG(ctx, ctx && a && b && out && n >= 0);
    HB_ARG(ctx, size == 4 || size == 8 || size == 16 || size == 32 || size == 64);
    if (size == 4)
        return hb_satd(ctx, a, b, n, 4, 4, out);
    HB_ARG(ctx, ((a | b) & 15) == 0);
    if (!n) return HB_OK;
    int T = (size / 8) * (size / 8);
    T = T > 32 ? 32 : T;
    k_sa8d<<<hb_grid_for(ctx, (long long)n * T, 128, 8), 128, 0, ctx->stream>>>((const pixel *)a, (const pixel *)b, n, size, T, (int *)out);
    HB_LAUNCHED(ctx);
    return HB_OK;
}

#define HB_TRANSFORM_LAUNCH(N, DST, INV)                                                                                  \
    k_transform<N, DST, INV><<<hb_grid_for(ctx, ((long long)n + (32 / N) - 1) / (32 / N) * 32, 128, 8), 128, 0, ctx->stream>>>( \
        (const int16_t *)src, (int16_t *)dst, n, bit_depth)

static int transform_dispatch(hb_ctx *ctx, hb_devptr src, int n, int size, int bit_depth, int is_dst, hb_devptr dst, bool inverse)
{
    HB_ARG(ctx, ctx && src && dst && n >= 0 && bit_depth >= 8 && bit_depth <= 12);
    HB_ARG(ctx, size == 4 || size == 8 || size == 16 || size == 32);
    HB_ARG(ctx, !is_dst || size == 4);
    if (!n) return HB_OK;
    if (inverse) {
        if (is_dst) HB_TRANSFORM_LAUNCH(4, true, true);
        else if (size == 4) HB_TRANSFORM_LAUNCH(4, false, true);
        else if (size == 8) HB_TRANSFORM_LAUNCH(8, false, true);
        else if (size == 16) HB_TRANSFORM_LAUNCH(16, false, true);
        else HB_TRANSFORM_LAUNCH(32, false, true);
    } else {
        if (is_dst) HB_TRANSFORM_LAUNCH(4, true, false);
        else if (size == 4) HB_TRANSFORM_LAUNCH(4, false, false);
        else if (size == 8) HB_TRANSFORM_LAUNCH(8, false, false);
        else if (size == 16) HB_TRANSFORM_LAUNCH(16, false, false);
        else HB_TRANSFORM_LAUNCH(32, false, false);
    }
    HB_LAUNCHED(ctx);
    return HB_OK;
}

int hb_fwd_transform(hb_ctx *ctx, hb_devptr residual, int n, int size, int bit_depth, int is_dst, hb_devptr coef)
{
    return transform_dispatch(ctx, residual, n, size, bit_depth, is_dst, coef, false);
}

int hb_inv_transform(hb_ctx *ctx, hb_devptr coef, int n, int size, int bit_depth, int is_dst, hb_devptr residual)
{
    return transform_dispatch(ctx, coef, n, size, bit_depth, is_dst, residual, true);
}

int hb_quant(hb_ctx *ctx, hb_devptr coef, int n, int size, int qp, int bit_depth, int is_intra, hb_devptr level, hb_devptr numsig)
{
    HB_ARG(ctx, ctx && coef && level && numsig && n >= 0 && qp >= 0 && qp <= 51 + 6 * (bit_depth - 8));
    HB_ARG(ctx, size == 4 || size == 8 || size == 16 || size == 32);
    if (!n) return HB_OK;
    const QuantParam q = make_quant(ilog2c(size), qp, bit_depth, is_intra);
    k_quant<<<hb_grid_for(ctx, (long long)n * 32, 256, 8), 256, 0, ctx->stream>>>((const int16_t *)coef, (int16_t *)level, (int *)numsig, n,
                                                                               size * size, q);
    HB_LAUNCHED(ctx);
    return HB_OK;
}

int hb_dequant(hb_ctx *ctx, hb_devptr level, int n, int size, int qp, int bit_depth, hb_devptr coef)
{
    HB_ARG(ctx, ctx && coef && level && n >= 0 && qp >= 0 && qp <= 51 + 6 * (bit_depth - 8));
    HB_ARG(ctx, size == 4 || size == 8 || size == 16 || size == 32);
    if (!n) return HB_OK;
    const QuantParam q = make_quant(ilog2c(size), qp, bit_depth, 0);
    const long long total2 = (long long)n * size * size / 2;
    k_dequant<<<hb_grid_for(ctx, total2, 256, 8), 256, 0, ctx->stream>>>((const int16_t *)level, (int16_t *)coef, total2, q);
    HB_LAUNCHED(ctx);
    return HB_OK;
}

int hb_intra_pred_all(hb_ctx *ctx, hb_devptr neighbours, int n, int size, int is_luma, int strong, int bit_depth, hb_devptr pred)
{
    HB_ARG(ctx, ctx && neighbours && pred && n >= 0 && bit_depth >= 8 && bit_depth <= 12);
    HB_ARG(ctx, size == 4 || size == 8 || size == 16 || size == 32);
    if (!n) return HB_OK;
    k_intra_all<<<hb_grid_for(ctx, (long long)n * 32, 128, 8), 128, 0, ctx->stream>>>((const pixel *)neighbours, (pixel *)pred, n, size,
                                                                                   is_luma, strong, bit_depth);
    HB_LAUNCHED(ctx);
    return HB_OK;
}

}  // extern "C"
