// Pre-encode pixel pipeline for sm_100a: HBM-bound, 128-bit vectorised, streaming cache hints,
// grids sized to a multiple of the SM count.  Fixed-point definitions: oracle/pixel_ref.py.
#include <math.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint4 ldg_stream(const void *p) { return __ldcs(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ void stg_stream(void *p, uint4 v) { __stcs(reinterpret_cast<uint4 *>(p), v); }

// four bytes -> two 32-bit words holding (byte << 8) in each 16-bit half
__device__ __forceinline__ void widen4(uint32_t x, uint32_t &lo, uint32_t &hi)
{
    lo = __byte_perm(x, 0, 0x1404);
    hi = __byte_perm(x, 0, 0x3424);
}

// ---------------------------------------------------------------------------------------------- pack P010
// work item = 16 luma samples of one row, or 16 U + 16 V samples of one chroma row
__global__ void __launch_bounds__(256) k_pack_p010(const uint8_t *__restrict__ y, int ys, const uint8_t *__restrict__ u, int us,
                                                   const uint8_t *__restrict__ v, int vs, int w, int h, uint8_t *__restrict__ dy,
                                                   int dys, uint8_t *__restrict__ duv, int duvs, int aligned, size_t in_fs, size_t out_fs)
{
    // blockIdx.y = frame of a batch (in_fs / out_fs: bytes between successive frames)
    y += blockIdx.y * in_fs; u += blockIdx.y * in_fs; v += blockIdx.y * in_fs;
    dy += blockIdx.y * out_fs; duv += blockIdx.y * out_fs;
    const int cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    const int vy = (w + 15) >> 4, vc = (cw + 15) >> 4;
    const long long n_y = (long long)vy * h, total = n_y + (long long)vc * ch;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        if (i < n_y) {
            const int row = (int)(i / vy), x = (int)(i % vy) << 4;
            const uint8_t *s = y + (size_t)row * ys + x;
            uint8_t *d = dy + (size_t)row * dys + 2 * x;
            if (aligned && x + 16 <= w) {
                uint4 a = ldg_stream(s), o0, o1;
                widen4(a.x, o0.x, o0.y); widen4(a.y, o0.z, o0.w);
                widen4(a.z, o1.x, o1.y); widen4(a.w, o1.z, o1.w);
                stg_stream(d, o0);
                stg_stream(d + 16, o1);
            } else {
                for (int k = 0; k < 16 && x + k < w; k++)
                    reinterpret_cast<uint16_t *>(d)[k] = (uint16_t)(s[k] << 8);
            }
        } else {
            const long long j = i - n_y;
            const int row = (int)(j / vc), x = (int)(j % vc) << 4;
            const uint8_t *su = u + (size_t)row * us + x, *sv = v + (size_t)row * vs + x;
            uint8_t *d = duv + (size_t)row * duvs + 4 * x;
            if (aligned && x + 16 <= cw) {
                uint4 a = ldg_stream(su), b = ldg_stream(sv);
                const uint32_t au[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    uint32_t t0 = __byte_perm(au[k], bv[k], 0x5140);   // u0 v0 u1 v1
                    uint32_t t1 = __byte_perm(au[k], bv[k], 0x7362);   // u2 v2 u3 v3
                    uint4 o;
                    widen4(t0, o.x, o.y);
                    widen4(t1, o.z, o.w);
                    stg_stream(d + 16 * k, o);
                }
            } else {
                for (int k = 0; k < 16 && x + k < cw; k++) {
                    reinterpret_cast<uint16_t *>(d)[2 * k] = (uint16_t)(su[k] << 8);
                    reinterpret_cast<uint16_t *>(d)[2 * k + 1] = (uint16_t)(sv[k] << 8);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- RGB -> YUV 4:2:0
struct CscCoef {
    int cy[3], ccb[3], ccr[3], yoff, coff, maxv;
};

// work item = 16 pixels x 2 rows
// PLANAR16: 16-bit planar output at DEPTH bits, samples LSB-aligned, separate U and V planes (the stream encoder's source planes)
template <int DEPTH, bool PLANAR16 = false>
__global__ void __launch_bounds__(128) k_rgb_to_yuv420(const uint8_t *__restrict__ rgb, int rs, int bgr, CscCoef c, int w, int h,
                                                       uint8_t *__restrict__ dy, int dys, uint8_t *__restrict__ du, int dus,
                                                       uint8_t *__restrict__ dv, int dvs, int aligned, size_t in_fs = 0, size_t out_fs = 0)
{
    rgb += blockIdx.y * in_fs;
    dy += blockIdx.y * out_fs; du += blockIdx.y * out_fs;
    if (dv) dv += blockIdx.y * out_fs;
    const int vpr = (w + 15) >> 4, rows2 = h >> 1;
    const long long total = (long long)vpr * rows2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / vpr) * 2, x = (int)(i % vpr) << 4;
        const int n = min(16, w - x);
        uint32_t px[2][12];
        if (aligned && n == 16) {
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const uint8_t *s = rgb + (size_t)(row + r) * rs + 3 * x;
                uint4 a = ldg_stream(s), b = ldg_stream(s + 16), d = ldg_stream(s + 32);
                px[r][0] = a.x; px[r][1] = a.y; px[r][2] = a.z; px[r][3] = a.w;
                px[r][4] = b.x; px[r][5] = b.y; px[r][6] = b.z; px[r][7] = b.w;
                px[r][8] = d.x; px[r][9] = d.y; px[r][10] = d.z; px[r][11] = d.w;
            }
        } else {
            for (int r = 0; r < 2; r++) {
                const uint8_t *s = rgb + (size_t)(row + r) * rs + 3 * x;
                for (int k = 0; k < 12; k++) {
                    uint32_t wd = 0;
                    for (int b = 0; b < 4; b++)
                        if ((4 * k + b) < 3 * n)
                            wd |= (uint32_t)s[4 * k + b] << (8 * b);
                    px[r][k] = wd;
                }
            }
        }
        int yv[2][16];
        int sum[8][3];
#pragma unroll
        for (int q = 0; q < 8; q++)
            sum[q][0] = sum[q][1] = sum[q][2] = 0;
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int k = 0; k < 16; k++) {
                int ch[3];
#pragma unroll
                for (int comp = 0; comp < 3; comp++) {
                    const int byte = 3 * k + comp;
                    ch[comp] = (px[r][byte >> 2] >> (8 * (byte & 3))) & 255;
                }
                const int R = bgr ? ch[2] : ch[0], G = ch[1], B = bgr ? ch[0] : ch[2];
                int yy = ((R * c.cy[0] + G * c.cy[1] + B * c.cy[2] + (1 << 13)) >> 14) + c.yoff;
                yv[r][k] = min(max(yy, 0), c.maxv);
                sum[k >> 1][0] += R; sum[k >> 1][1] += G; sum[k >> 1][2] += B;
            }
        int cb[8], cr[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            int b = ((sum[q][0] * c.ccb[0] + sum[q][1] * c.ccb[1] + sum[q][2] * c.ccb[2] + (1 << 15)) >> 16) + c.coff;
            int r = ((sum[q][0] * c.ccr[0] + sum[q][1] * c.ccr[1] + sum[q][2] * c.ccr[2] + (1 << 15)) >> 16) + c.coff;
            cb[q] = min(max(b, 0), c.maxv);
            cr[q] = min(max(r, 0), c.maxv);
        }
        const int crow = row >> 1, cx = x >> 1;
        if (PLANAR16) {
            if (aligned && n == 16) {
#pragma unroll
                for (int r = 0; r < 2; r++)
#pragma unroll
                    for (int hlf = 0; hlf < 2; hlf++) {
                        uint4 o;
                        uint32_t *ow = &o.x;
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            ow[k] = (uint32_t)yv[r][8 * hlf + 2 * k] | ((uint32_t)yv[r][8 * hlf + 2 * k + 1] << 16);
                        *reinterpret_cast<uint4 *>(dy + (size_t)(row + r) * dys + 2 * x + 16 * hlf) = o;
                    }
                uint4 ob, orr;
                uint32_t *pb = &ob.x, *pr = &orr.x;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    pb[k] = (uint32_t)cb[2 * k] | ((uint32_t)cb[2 * k + 1] << 16);
                    pr[k] = (uint32_t)cr[2 * k] | ((uint32_t)cr[2 * k + 1] << 16);
                }
                *reinterpret_cast<uint4 *>(du + (size_t)crow * dus + 2 * cx) = ob;
                *reinterpret_cast<uint4 *>(dv + (size_t)crow * dvs + 2 * cx) = orr;
            } else {
                for (int r = 0; r < 2; r++)
                    for (int k = 0; k < n; k++)
                        reinterpret_cast<uint16_t *>(dy + (size_t)(row + r) * dys)[x + k] = (uint16_t)yv[r][k];
                for (int q = 0; q < n / 2; q++) {
                    reinterpret_cast<uint16_t *>(du + (size_t)crow * dus)[cx + q] = (uint16_t)cb[q];
                    reinterpret_cast<uint16_t *>(dv + (size_t)crow * dvs)[cx + q] = (uint16_t)cr[q];
                }
            }
        } else if (DEPTH == 8) {
            if (aligned && n == 16) {
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    uint4 o;
                    uint32_t *ow = &o.x;
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        ow[k] = yv[r][4 * k] | (yv[r][4 * k + 1] << 8) | (yv[r][4 * k + 2] << 16) | (yv[r][4 * k + 3] << 24);
                    stg_stream(dy + (size_t)(row + r) * dys + x, o);
                }
                uint2 ob, orr;
                ob.x = cb[0] | (cb[1] << 8) | (cb[2] << 16) | (cb[3] << 24);
                ob.y = cb[4] | (cb[5] << 8) | (cb[6] << 16) | (cb[7] << 24);
                orr.x = cr[0] | (cr[1] << 8) | (cr[2] << 16) | (cr[3] << 24);
                orr.y = cr[4] | (cr[5] << 8) | (cr[6] << 16) | (cr[7] << 24);
                *reinterpret_cast<uint2 *>(du + (size_t)crow * dus + cx) = ob;
                *reinterpret_cast<uint2 *>(dv + (size_t)crow * dvs + cx) = orr;
            } else {
                for (int r = 0; r < 2; r++)
                    for (int k = 0; k < n; k++)
                        dy[(size_t)(row + r) * dys + x + k] = (uint8_t)yv[r][k];
                for (int q = 0; q < n / 2; q++) {
                    du[(size_t)crow * dus + cx + q] = (uint8_t)cb[q];
                    dv[(size_t)crow * dvs + cx + q] = (uint8_t)cr[q];
                }
            }
        } else {  // P010: value << 6, UV interleaved in `du`
            if (aligned && n == 16) {
#pragma unroll
                for (int r = 0; r < 2; r++)
#pragma unroll
                    for (int hlf = 0; hlf < 2; hlf++) {
                        uint4 o;
                        uint32_t *ow = &o.x;
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            ow[k] = ((uint32_t)yv[r][8 * hlf + 2 * k] << 6) | ((uint32_t)yv[r][8 * hlf + 2 * k + 1] << 22);
                        stg_stream(dy + (size_t)(row + r) * dys + 2 * x + 16 * hlf, o);
                    }
#pragma unroll
                for (int hlf = 0; hlf < 2; hlf++) {
                    uint4 o;
                    uint32_t *ow = &o.x;
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        ow[k] = ((uint32_t)cb[4 * hlf + k] << 6) | ((uint32_t)cr[4 * hlf + k] << 22);
                    stg_stream(du + (size_t)crow * dus + 4 * cx + 16 * hlf, o);
                }
            } else {
                for (int r = 0; r < 2; r++)
                    for (int k = 0; k < n; k++)
                        reinterpret_cast<uint16_t *>(dy + (size_t)(row + r) * dys)[x + k] = (uint16_t)(yv[r][k] << 6);
                for (int q = 0; q < n / 2; q++) {
                    reinterpret_cast<uint16_t *>(du + (size_t)crow * dus)[2 * (cx + q)] = (uint16_t)(cb[q] << 6);
                    reinterpret_cast<uint16_t *>(du + (size_t)crow * dus)[2 * (cx + q) + 1] = (uint16_t)(cr[q] << 6);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- polyphase scaler
constexpr int SC_TW = 64, SC_TH = 16, SC_FW = 144, SC_FH = 40;
__constant__ short c_bicubic[64][4];

// One CTA = one 64x16 output tile of up to two source planes (PLANES == 2 writes interleaved pairs).
// Stage the clamped source footprint in shared memory, horizontal pass to Q6 int16, vertical pass out.
template <int PLANES>
__global__ void __launch_bounds__(256) k_scale(const uint8_t *__restrict__ s0, const uint8_t *__restrict__ s1, int ss, int sw, int sh,
                                               uint8_t *__restrict__ dst, int ds, int dw, int dh, const int2 *__restrict__ xtab,
                                               const int2 *__restrict__ ytab, int out_depth, int out_shift, int step)
{
    __shared__ __align__(16) uint8_t foot[PLANES][SC_FH][SC_FW];
    __shared__ short hp[PLANES][SC_FH][SC_TW];
    __shared__ short xtap[SC_TW][4], ytap[SC_TH][4];     // per-tile copies of the phase taps (constant memory serialises on divergent phases)
    __shared__ int xfirst[SC_TW], yfirst[SC_TH];
    const int tiles_x = (dw + SC_TW - 1) / SC_TW, tiles_y = (dh + SC_TH - 1) / SC_TH;
    const int tid = threadIdx.x;
    for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
        const int ox0 = (tile % tiles_x) * SC_TW, oy0 = (tile / tiles_x) * SC_TH;
        const int tw = min(SC_TW, dw - ox0), th = min(SC_TH, dh - oy0);
        const int fx0 = xtab[ox0].x, fx1 = xtab[ox0 + tw - 1].x + 3;
        const int fy0 = ytab[oy0].x, fy1 = ytab[oy0 + th - 1].x + 3;
        const int fw = fx1 - fx0 + 1, fh = fy1 - fy0 + 1;
        const bool staged = fw <= SC_FW && fh <= SC_FH;
        const int shift = 20 - (out_depth - 8), maxv = (1 << out_depth) - 1;
        __syncthreads();
        if (tid < tw) {
            const int2 xp = xtab[ox0 + tid];
            xfirst[tid] = xp.x;
#pragma unroll
            for (int t = 0; t < 4; t++) xtap[tid][t] = c_bicubic[xp.y][t];
        } else if (tid >= 64 && tid < 64 + th) {
            const int2 yp = ytab[oy0 + tid - 64];
            yfirst[tid - 64] = yp.x;
#pragma unroll
            for (int t = 0; t < 4; t++) ytap[tid - 64][t] = c_bicubic[yp.y][t];
        }
        if (staged) {
            // footprint: interior tiles copy aligned 32-bit words (the footprint starts at fx0 rounded down to 4 bytes: `fskew`
            // bytes of slack on the left), edge tiles clamp sample by sample
            const int fxa = fx0 & ~3, fskew = fx0 - fxa, nw = (fskew + fw + 3) >> 2;
            const bool words = fxa >= 0 && fxa + 4 * nw <= sw && fy0 >= 0 && fy1 < sh && (ss & 3) == 0 && 4 * nw <= SC_FW &&
                               ((reinterpret_cast<uintptr_t>(s0) | (PLANES == 2 ? reinterpret_cast<uintptr_t>(s1) : 0)) & 3) == 0;
            const int skew = words ? fskew : 0;
            if (words) {
                for (int i = tid; i < fh * nw; i += 256) {
                    const int r = i / nw, wi = i - r * nw;
                    const size_t off = (size_t)(fy0 + r) * ss + fxa + 4 * wi;
                    reinterpret_cast<uint32_t *>(&foot[0][r][0])[wi] = *reinterpret_cast<const uint32_t *>(s0 + off);
                    if (PLANES == 2) reinterpret_cast<uint32_t *>(&foot[1][r][0])[wi] = *reinterpret_cast<const uint32_t *>(s1 + off);
                }
            } else {
                for (int r = tid >> 5; r < fh; r += 8) {
                    const int sy = min(max(fy0 + r, 0), sh - 1);
                    for (int cidx = tid & 31; cidx < fw; cidx += 32) {
                        const int sx = min(max(fx0 + cidx, 0), sw - 1);
                        foot[0][r][cidx] = s0[(size_t)sy * ss + sx];
                        if (PLANES == 2)
                            foot[1][r][cidx] = s1[(size_t)sy * ss + sx];
                    }
                }
            }
            __syncthreads();
            {   // horizontal pass: a thread keeps one output column (256 % 64 == 0), its taps and source offset live in registers
                const int ox = tid & (SC_TW - 1);
                if (ox < tw) {
                    const int base = xfirst[ox] - fx0 + skew;
                    const int t0 = xtap[ox][0], t1 = xtap[ox][1], t2 = xtap[ox][2], t3 = xtap[ox][3];
                    for (int r = tid >> 6; r < fh; r += 4) {
#pragma unroll
                        for (int p = 0; p < PLANES; p++) {
                            const uint8_t *f = &foot[p][r][base];
                            hp[p][r][ox] = (short)((128 + t0 * f[0] + t1 * f[1] + t2 * f[2] + t3 * f[3]) >> 8);
                        }
                    }
                }
            }
            __syncthreads();
        }
        for (int i = tid; i < SC_TW * SC_TH; i += 256) {
            const int oy = i >> 6, ox = i & (SC_TW - 1);
            if (ox >= tw || oy >= th) continue;
            const int2 yp = ytab[oy0 + oy];
            int outv[PLANES];
#pragma unroll
            for (int p = 0; p < PLANES; p++) {
                int acc = 1 << (shift - 1);
                if (staged) {
                    const int base = yfirst[oy] - fy0;
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        acc += ytap[oy][t] * hp[p][base + t][ox];
                } else {  // extreme down-scale ratios: read the source directly
                    const uint8_t *sp = p ? s1 : s0;
                    const int2 xp = xtab[ox0 + ox];
                    for (int t = 0; t < 4; t++) {
                        const int sy = min(max(yp.x + t, 0), sh - 1);
                        int hacc = 128;
                        for (int k = 0; k < 4; k++)
                            hacc += c_bicubic[xp.y][k] * sp[(size_t)sy * ss + min(max(xp.x + k, 0), sw - 1)];
                        acc += c_bicubic[yp.y][t] * (short)(hacc >> 8);
                    }
                }
                outv[p] = min(max(acc >> shift, 0), maxv);
            }
            uint8_t *row = dst + (size_t)(oy0 + oy) * ds;
            const int X = ox0 + ox;
            if (out_depth == 8) {
                row[X * step] = (uint8_t)outv[0];
            } else if (PLANES == 2) {
                reinterpret_cast<uint32_t *>(row)[X] = (uint32_t)(outv[0] << out_shift) | ((uint32_t)(outv[1] << out_shift) << 16);
            } else {
                reinterpret_cast<uint16_t *>(row)[X * step] = (uint16_t)(outv[0] << out_shift);
            }
        }
    }
}


// ---- vectorised scaler: one CTA = one 128x16 output tile, one thread = 8 consecutive output samples of one row
// (one 128-bit store per plane for 16-bit output).  Same arithmetic as k_scale above (oracle/pixel_ref.py scale_plane).
//   OUT 0: 8-bit planar, OUT 1: 16-bit planar (PLANES == 2: two destination planes), OUT 2: 16-bit interleaved pairs (P010 UV)
constexpr int S2_TW = 128, S2_TH = 16, S2_FW = 272, S2_FH = 40;
template <int PLANES, int OUT>
__global__ void __launch_bounds__(256) k_scale8(const uint8_t *__restrict__ s0, const uint8_t *__restrict__ s1, int ss, int sw, int sh,
                                                uint8_t *__restrict__ d0, uint8_t *__restrict__ d1, int ds, int dw, int dh,
                                                const int2 *__restrict__ xtab, const int2 *__restrict__ ytab, int out_depth, int out_shift,
                                                int aligned, size_t in_fs = 0, size_t out_fs = 0)
{
    s0 += blockIdx.y * in_fs; d0 += blockIdx.y * out_fs;
    if (PLANES == 2) { s1 += blockIdx.y * in_fs; if (d1) d1 += blockIdx.y * out_fs; }
    __shared__ __align__(16) uint8_t foot[PLANES][S2_FH][S2_FW];
    __shared__ __align__(16) short hp[PLANES][S2_FH][S2_TW];
    __shared__ short xtap[S2_TW][4], ytap[S2_TH][4];
    __shared__ int xfirst[S2_TW], yfirst[S2_TH];
    const int tiles_x = (dw + S2_TW - 1) / S2_TW;
    const int tid = threadIdx.x, tile = blockIdx.x;
    const int ox0 = (tile % tiles_x) * S2_TW, oy0 = (tile / tiles_x) * S2_TH;
    const int tw = min(S2_TW, dw - ox0), th = min(S2_TH, dh - oy0);
    const int fx0 = xtab[ox0].x, fx1 = xtab[ox0 + tw - 1].x + 3;
    const int fy0 = ytab[oy0].x, fy1 = ytab[oy0 + th - 1].x + 3;
    const int fw = fx1 - fx0 + 1, fh = fy1 - fy0 + 1;
    const int shift = 20 - (out_depth - 8), maxv = (1 << out_depth) - 1;
    const int fxa = fx0 & ~3, fskew = fx0 - fxa, nw = (fskew + fw + 3) >> 2;
    const bool staged = 4 * nw <= S2_FW && fh <= S2_FH;
    if (tid < tw) {
        const int2 xp = xtab[ox0 + tid];
        xfirst[tid] = xp.x;
#pragma unroll
        for (int t = 0; t < 4; t++) xtap[tid][t] = c_bicubic[xp.y][t];
    } else if (tid >= 128 && tid < 128 + th) {
        const int2 yp = ytab[oy0 + tid - 128];
        yfirst[tid - 128] = yp.x;
#pragma unroll
        for (int t = 0; t < 4; t++) ytap[tid - 128][t] = c_bicubic[yp.y][t];
    }
    if (!staged) {          // extreme down-scale ratios: every output straight from global memory
        for (int i = tid; i < S2_TW * S2_TH; i += 256) {
            const int oy = i >> 7, ox = i & (S2_TW - 1);
            if (ox >= tw || oy >= th) continue;
            const int2 yp = ytab[oy0 + oy], xp = xtab[ox0 + ox];
            int outv[PLANES];
#pragma unroll
            for (int p = 0; p < PLANES; p++) {
                const uint8_t *sp = p ? s1 : s0;
                int acc = 1 << (shift - 1);
                for (int t = 0; t < 4; t++) {
                    const int sy = min(max(yp.x + t, 0), sh - 1);
                    int hacc = 128;
                    for (int k = 0; k < 4; k++) hacc += c_bicubic[xp.y][k] * sp[(size_t)sy * ss + min(max(xp.x + k, 0), sw - 1)];
                    acc += c_bicubic[yp.y][t] * (short)(hacc >> 8);
                }
                outv[p] = min(max(acc >> shift, 0), maxv);
            }
            const int X = ox0 + ox;
            const size_t ro = (size_t)(oy0 + oy) * ds;
            if (OUT == 0) d0[ro + X] = (uint8_t)outv[0];
            else if (OUT == 2) reinterpret_cast<uint32_t *>(d0 + ro)[X] = (uint32_t)(outv[0] << out_shift) | ((uint32_t)(outv[PLANES - 1] << out_shift) << 16);
            else {
                reinterpret_cast<uint16_t *>(d0 + ro)[X] = (uint16_t)(outv[0] << out_shift);
                if (PLANES == 2) reinterpret_cast<uint16_t *>(d1 + ro)[X] = (uint16_t)(outv[PLANES - 1] << out_shift);
            }
        }
        return;
    }
    // footprint: interior tiles copy aligned 32-bit words, picture-edge tiles clamp sample by sample
    const bool words = fxa >= 0 && fxa + 4 * nw <= sw && fy0 >= 0 && fy1 < sh && (ss & 3) == 0 &&
                       ((reinterpret_cast<uintptr_t>(s0) | (PLANES == 2 ? reinterpret_cast<uintptr_t>(s1) : 0)) & 3) == 0;
    const int skew = words ? fskew : 0;
    if (words) {
        for (int i = tid; i < fh * nw; i += 256) {
            const int r = i / nw, wi = i - r * nw;
            const size_t off = (size_t)(fy0 + r) * ss + fxa + 4 * wi;
            reinterpret_cast<uint32_t *>(&foot[0][r][0])[wi] = __ldg(reinterpret_cast<const uint32_t *>(s0 + off));
            if (PLANES == 2) reinterpret_cast<uint32_t *>(&foot[PLANES - 1][r][0])[wi] = __ldg(reinterpret_cast<const uint32_t *>(s1 + off));
        }
    } else {
        for (int r = tid >> 5; r < fh; r += 8) {
            const int sy = min(max(fy0 + r, 0), sh - 1);
            for (int cidx = tid & 31; cidx < fw; cidx += 32) {
                const int sx = min(max(fx0 + cidx, 0), sw - 1);
                foot[0][r][cidx] = s0[(size_t)sy * ss + sx];
                if (PLANES == 2) foot[PLANES - 1][r][cidx] = s1[(size_t)sy * ss + sx];
            }
        }
    }
    __syncthreads();
    {   // horizontal pass: a thread keeps one output column, its taps and source offset live in registers
        const int ox = tid & (S2_TW - 1);
        if (ox < tw) {
            const int base = xfirst[ox] - fx0 + skew;
            const int t0 = xtap[ox][0], t1 = xtap[ox][1], t2 = xtap[ox][2], t3 = xtap[ox][3];
            for (int r = tid >> 7; r < fh; r += 2) {
#pragma unroll
                for (int p = 0; p < PLANES; p++) {
                    const uint8_t *f = &foot[p][r][base];
                    hp[p][r][ox] = (short)((128 + t0 * f[0] + t1 * f[1] + t2 * f[2] + t3 * f[3]) >> 8);
                }
            }
        }
    }
    __syncthreads();
    // vertical pass: 8 consecutive columns of one output row per thread
    const int oy = tid >> 4, xg = (tid & 15) * 8;
    if (oy >= th || xg >= tw) return;
    const int base = yfirst[oy] - fy0;
    int outv[PLANES][8];
#pragma unroll
    for (int p = 0; p < PLANES; p++) {
        int acc[8];
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = 1 << (shift - 1);
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int yt = ytap[oy][t];
            const uint4 v = *reinterpret_cast<const uint4 *>(&hp[p][base + t][xg]);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                acc[2 * k] += yt * (short)(w[k] & 0xffff);
                acc[2 * k + 1] += yt * (short)(w[k] >> 16);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) outv[p][k] = min(max(acc[k] >> shift, 0), maxv) << out_shift;
    }
    const int X = ox0 + xg, nvalid = min(8, tw - xg);
    const size_t ro = (size_t)(oy0 + oy) * ds;
    if (aligned && nvalid == 8) {
        if (OUT == 0) {
            uint2 o;
            o.x = outv[0][0] | (outv[0][1] << 8) | (outv[0][2] << 16) | (outv[0][3] << 24);
            o.y = outv[0][4] | (outv[0][5] << 8) | (outv[0][6] << 16) | (outv[0][7] << 24);
            *reinterpret_cast<uint2 *>(d0 + ro + X) = o;
        } else if (OUT == 1) {
#pragma unroll
            for (int p = 0; p < PLANES; p++) {
                uint4 o;
                o.x = outv[p][0] | (outv[p][1] << 16); o.y = outv[p][2] | (outv[p][3] << 16);
                o.z = outv[p][4] | (outv[p][5] << 16); o.w = outv[p][6] | (outv[p][7] << 16);
                *reinterpret_cast<uint4 *>((p ? d1 : d0) + ro + 2 * X) = o;
            }
        } else {
            uint4 a, b;
            a.x = outv[0][0] | (outv[PLANES - 1][0] << 16); a.y = outv[0][1] | (outv[PLANES - 1][1] << 16);
            a.z = outv[0][2] | (outv[PLANES - 1][2] << 16); a.w = outv[0][3] | (outv[PLANES - 1][3] << 16);
            b.x = outv[0][4] | (outv[PLANES - 1][4] << 16); b.y = outv[0][5] | (outv[PLANES - 1][5] << 16);
            b.z = outv[0][6] | (outv[PLANES - 1][6] << 16); b.w = outv[0][7] | (outv[PLANES - 1][7] << 16);
            *reinterpret_cast<uint4 *>(d0 + ro + 4 * X) = a;
            *reinterpret_cast<uint4 *>(d0 + ro + 4 * X + 16) = b;
        }
    } else {
        for (int k = 0; k < nvalid; k++) {
            if (OUT == 0) d0[ro + X + k] = (uint8_t)outv[0][k];
            else if (OUT == 2) reinterpret_cast<uint32_t *>(d0 + ro)[X + k] = (uint32_t)outv[0][k] | ((uint32_t)outv[PLANES - 1][k] << 16);
            else {
                reinterpret_cast<uint16_t *>(d0 + ro)[X + k] = (uint16_t)outv[0][k];
                if (PLANES == 2) reinterpret_cast<uint16_t *>(d1 + ro)[X + k] = (uint16_t)outv[PLANES - 1][k];
            }
        }
    }
}

bool is_aligned16(uint64_t p) { return (p & 15) == 0; }

void bicubic_table(short tab[64][4])
{
    const double a = -0.5;
    for (int p = 0; p < 64; p++) {
        const double t = p / 64.0;
        const double d[4] = {1 + t, t, 1 - t, 2 - t};
        int q[4], sum = 0;
        for (int k = 0; k < 4; k++) {
            const double x = d[k];
            const double wgt = x <= 1 ? (a + 2) * x * x * x - (a + 3) * x * x + 1 : a * x * x * x - 5 * a * x * x + 8 * a * x - 4 * a;
            q[k] = (int)nearbyint(wgt * 16384.0);
            sum += q[k];
        }
        q[t < 0.5 ? 1 : 2] += 16384 - sum;
        for (int k = 0; k < 4; k++)
            tab[p][k] = (short)q[k];
    }
}

// centre-aligned destination -> (first tap index, phase) table, cached per (src, dst) on the device
int scale_table(hb_ctx *ctx, int src, int dst, const int2 **out)
{
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!ctx->bicubic_dev) {
        short tab[64][4];
        bicubic_table(tab);
        HB_CUDA(ctx, cudaMemcpyToSymbolAsync(c_bicubic, tab, sizeof(tab), 0, cudaMemcpyHostToDevice, ctx->stream));
        HB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        HB_CUDA(ctx, cudaMalloc(&ctx->bicubic_dev, 16));   // marks "constant table uploaded"
    }
    const uint64_t key = ((uint64_t)(uint32_t)src << 32) | (uint32_t)dst;
    auto it = ctx->scale_tabs.find(key);
    if (it == ctx->scale_tabs.end()) {
        std::vector<int2> host(dst);
        for (int d = 0; d < dst; d++) {
            const long long num = (2LL * d + 1) * src - dst, den = 2LL * dst;
            long long ix = num >= 0 ? num / den : -((-num + den - 1) / den);
            const long long frac = num - ix * den;
            long long phase = (frac * 64 + dst) / den;
            if (phase == 64) { phase = 0; ix++; }
            host[d] = make_int2((int)ix - 1, (int)phase);
        }
        int2 *dev = nullptr;
        HB_CUDA(ctx, cudaMalloc(&dev, sizeof(int2) * dst));
        HB_CUDA(ctx, cudaMemcpyAsync(dev, host.data(), sizeof(int2) * dst, cudaMemcpyHostToDevice, ctx->stream));
        HB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        it = ctx->scale_tabs.emplace(key, dev).first;
    }
    *out = it->second;
    return HB_OK;
}

int launch_scale(hb_ctx *ctx, const uint8_t *s0, const uint8_t *s1, int ss, int sw, int sh, uint8_t *dst, int ds, int dw, int dh,
                 int out_depth, int out_shift, int step)
{
    const int2 *xt, *yt;
    int rc = scale_table(ctx, sw, dw, &xt);
    if (rc) return rc;
    rc = scale_table(ctx, sh, dh, &yt);
    if (rc) return rc;
    const int tiles = ((dw + SC_TW - 1) / SC_TW) * ((dh + SC_TH - 1) / SC_TH);
    const int grid = tiles;      // one CTA per tile: the block scheduler overlaps the staging loads of one tile with the math of another
    if (s1)
        k_scale<2><<<grid, 256, 0, ctx->stream>>>(s0, s1, ss, sw, sh, dst, ds, dw, dh, xt, yt, out_depth, out_shift, step);
    else
        k_scale<1><<<grid, 256, 0, ctx->stream>>>(s0, nullptr, ss, sw, sh, dst, ds, dw, dh, xt, yt, out_depth, out_shift, step);
    HB_LAUNCHED(ctx);
    return HB_OK;
}

bool csc_coefficients(int matrix, int depth, CscCoef &c)
{
    double kr, kb;
    switch (matrix) {
    case HB_MATRIX_BT709: kr = 0.2126; kb = 0.0722; break;
    case HB_MATRIX_BT2020: kr = 0.2627; kb = 0.0593; break;
    case HB_MATRIX_BT601: kr = 0.299; kb = 0.114; break;
    default: return false;
    }
    const double kg = 1.0 - kr - kb;
    const double sy = (219 << (depth - 8)) / 255.0 * 16384.0, sc = (224 << (depth - 8)) / 255.0 * 16384.0;
    c.cy[0] = (int)nearbyint(kr * sy); c.cy[1] = (int)nearbyint(kg * sy); c.cy[2] = (int)nearbyint(kb * sy);
    c.ccb[0] = (int)nearbyint(-kr / (2 * (1 - kb)) * sc); c.ccb[1] = (int)nearbyint(-kg / (2 * (1 - kb)) * sc); c.ccb[2] = (int)nearbyint(0.5 * sc);
    c.ccr[0] = (int)nearbyint(0.5 * sc); c.ccr[1] = (int)nearbyint(-kg / (2 * (1 - kr)) * sc); c.ccr[2] = (int)nearbyint(-kb / (2 * (1 - kr)) * sc);
    c.yoff = 16 << (depth - 8); c.coff = 128 << (depth - 8); c.maxv = (1 << depth) - 1;
    return true;
}

}  // namespace

namespace hb {

// out_mode 0: 8-bit planar, 1: 16-bit planar (s1/d1 = second plane), 2: 16-bit interleaved pairs.  Runs on ctx->stream.
int launch_scale8(hb_ctx *ctx, const uint8_t *s0, const uint8_t *s1, int ss, int sw, int sh, uint8_t *d0, uint8_t *d1, int ds, int dw, int dh,
                  int out_depth, int out_shift, int out_mode, int n_frames, size_t in_fs, size_t out_fs)
{
    const int2 *xt, *yt;
    int rc = scale_table(ctx, sw, dw, &xt);
    if (rc) return rc;
    rc = scale_table(ctx, sh, dh, &yt);
    if (rc) return rc;
    const dim3 grid(((dw + S2_TW - 1) / S2_TW) * ((dh + S2_TH - 1) / S2_TH), n_frames);
    const int aligned = is_aligned16((uint64_t)(uintptr_t)d0 | (uint64_t)(uintptr_t)d1 | (uint64_t)ds | (uint64_t)out_fs);
    cudaStream_t st = ctx->stream;
    if (s1 && out_mode == 2) k_scale8<2, 2><<<grid, 256, 0, st>>>(s0, s1, ss, sw, sh, d0, d1, ds, dw, dh, xt, yt, out_depth, out_shift, aligned, in_fs, out_fs);
    else if (s1 && out_mode == 1) k_scale8<2, 1><<<grid, 256, 0, st>>>(s0, s1, ss, sw, sh, d0, d1, ds, dw, dh, xt, yt, out_depth, out_shift, aligned, in_fs, out_fs);
    else if (!s1 && out_mode == 1) k_scale8<1, 1><<<grid, 256, 0, st>>>(s0, nullptr, ss, sw, sh, d0, nullptr, ds, dw, dh, xt, yt, out_depth, out_shift, aligned, in_fs, out_fs);
    else if (!s1 && out_mode == 0) k_scale8<1, 0><<<grid, 256, 0, st>>>(s0, nullptr, ss, sw, sh, d0, nullptr, ds, dw, dh, xt, yt, out_depth, out_shift, aligned, in_fs, out_fs);
    else return hb_fail(ctx, HB_ERR_ARG, "bad argument: %s", "scaler output mode");
    HB_LAUNCHED(ctx);
    return HB_OK;
}

// packed 8-bit RGB / BGR -> 16-bit planar 4:2:0 at `depth` bits (LSB-aligned), strides in bytes.  Runs on ctx->stream.
int launch_rgb_planar16(hb_ctx *ctx, const uint8_t *rgb, int rs, int bgr, int matrix, int depth, int w, int h, uint8_t *dy, int dys, uint8_t *du,
                        uint8_t *dv, int dcs)
{
    CscCoef c;
    if (!csc_coefficients(matrix, depth, c)) return hb_fail(ctx, HB_ERR_ARG, "bad argument: %s", "matrix");
    const int aligned = is_aligned16((uint64_t)(uintptr_t)rgb | (uint64_t)(uintptr_t)dy | (uint64_t)(uintptr_t)du | (uint64_t)(uintptr_t)dv |
                                     (uint64_t)rs | (uint64_t)dys | (uint64_t)dcs);
    const long long items = (long long)((w + 15) / 16) * (h / 2);
    const int grid = hb_grid_for(ctx, items, 128, 8);
    if (depth == 8) k_rgb_to_yuv420<8, true><<<grid, 128, 0, ctx->stream>>>(rgb, rs, bgr, c, w, h, dy, dys, du, dcs, dv, dcs, aligned);
    else k_rgb_to_yuv420<10, true><<<grid, 128, 0, ctx->stream>>>(rgb, rs, bgr, c, w, h, dy, dys, du, dcs, dv, dcs, aligned);
    HB_LAUNCHED(ctx);
    return HB_OK;
}

}  // namespace hb

extern "C" {

int hb_pack_p010(hb_ctx *ctx, hb_devptr y, int ys, hb_devptr u, int us, hb_devptr v, int vs, int w, int h, hb_devptr dy, int dys,
                 hb_devptr duv, int duvs)
{
    HB_ARG(ctx, ctx && y && u && v && dy && duv && w > 0 && h > 0);
    HB_ARG(ctx, ys >= w && us >= (w + 1) / 2 && vs >= (w + 1) / 2 && dys >= 2 * w && duvs >= 4 * ((w + 1) / 2));
    const int aligned = is_aligned16(y | u | v | dy | duv | (uint64_t)ys | (uint64_t)us | (uint64_t)vs | (uint64_t)dys | (uint64_t)duvs);
    const long long items = (long long)((w + 15) / 16) * h + (long long)(((w + 1) / 2 + 15) / 16) * ((h + 1) / 2);
    k_pack_p010<<<hb_grid_for(ctx, items, 256, 8), 256, 0, ctx->stream>>>(
        (const uint8_t *)y, ys, (const uint8_t *)u, us, (const uint8_t *)v, vs, w, h, (uint8_t *)dy, dys, (uint8_t *)duv, duvs, aligned, 0, 0);
    HB_LAUNCHED(ctx);
    return HB_OK;
}

int hb_rgb_to_yuv420(hb_ctx *ctx, hb_devptr rgb, int rs, int order, int matrix, int depth, int w, int h, hb_devptr dy, int dys,
                     hb_devptr du, int dus, hb_devptr dv, int dvs)
{
    HB_ARG(ctx, ctx && rgb && dy && du && w > 0 && h > 0 && (w % 2) == 0 && (h % 2) == 0);
    HB_ARG(ctx, depth == 8 || depth == 10);
    HB_ARG(ctx, depth == 10 || dv);
    CscCoef c;
    if (!csc_coefficients(matrix, depth, c)) return hb_fail(ctx, HB_ERR_ARG, "bad argument: %s", "matrix");
    uint64_t bits = rgb | dy | du | (uint64_t)rs | (uint64_t)dys | (uint64_t)dus;
    if (depth == 8)
        bits |= dv | (uint64_t)dvs;
    const int aligned = is_aligned16(bits);
    const long long items = (long long)((w + 15) / 16) * (h / 2);
    const int grid = hb_grid_for(ctx, items, 128, 8);
    if (depth == 8)
        k_rgb_to_yuv420<8><<<grid, 128, 0, ctx->stream>>>((const uint8_t *)rgb, rs, order == HB_BGR, c, w, h, (uint8_t *)dy, dys,
                                                          (uint8_t *)du, dus, (uint8_t *)dv, dvs, aligned);
    else
        k_rgb_to_yuv420<10><<<grid, 128, 0, ctx->stream>>>((const uint8_t *)rgb, rs, order == HB_BGR, c, w, h, (uint8_t *)dy, dys,
                                                           (uint8_t *)du, dus, (uint8_t *)dv, dvs, aligned);
    HB_LAUNCHED(ctx);
    return HB_OK;
}

int hb_scale_plane(hb_ctx *ctx, hb_devptr src, int ss, int sw, int sh, hb_devptr dst, int ds, int dw, int dh, int out_depth,
                   int out_shift, int step)
{
    HB_ARG(ctx, ctx && src && dst && sw > 0 && sh > 0 && dw > 0 && dh > 0 && step >= 1);
    HB_ARG(ctx, (out_depth == 8 && out_shift == 0) || (out_depth == 10 && out_shift >= 0 && out_shift <= 6));
    if (step == 1)
        return hb::launch_scale8(ctx, (const uint8_t *)src, nullptr, ss, sw, sh, (uint8_t *)dst, nullptr, ds, dw, dh, out_depth, out_shift,
                                 out_depth == 8 ? 0 : 1, 1, 0, 0);
    return launch_scale(ctx, (const uint8_t *)src, nullptr, ss, sw, sh, (uint8_t *)dst, ds, dw, dh, out_depth, out_shift, step);
}

int hb_scale_yuv420_to_p010(hb_ctx *ctx, hb_devptr y, int ys, hb_devptr u, int us, hb_devptr v, int vs, int sw, int sh, hb_devptr dy,
                            int dys, hb_devptr duv, int duvs, int dw, int dh)
{
    HB_ARG(ctx, ctx && y && u && v && dy && duv && sw > 0 && sh > 0 && dw > 0 && dh > 0);
    HB_ARG(ctx, us == vs && (dw % 2) == 0 && (dh % 2) == 0 && (sw % 2) == 0 && (sh % 2) == 0);
    int rc = hb::launch_scale8(ctx, (const uint8_t *)y, nullptr, ys, sw, sh, (uint8_t *)dy, nullptr, dys, dw, dh, 10, 6, 1, 1, 0, 0);
    if (rc) return rc;
    return hb::launch_scale8(ctx, (const uint8_t *)u, (const uint8_t *)v, us, sw / 2, sh / 2, (uint8_t *)duv, nullptr, duvs, dw / 2, dh / 2, 10, 6, 2, 1, 0, 0);
}


/* ---- batched forms: one launch for n_frames tightly packed frames (grid.y = frame) */
int hb_pack_p010_batch(hb_ctx *ctx, hb_devptr src, size_t src_frame_bytes, hb_devptr dst, size_t dst_frame_bytes, int w, int h, int n_frames)
{
    HB_ARG(ctx, ctx && src && dst && w > 0 && h > 0 && !(w & 1) && !(h & 1) && n_frames >= 1 && n_frames <= 65535);
    const size_t luma = (size_t)w * h, chroma = (size_t)(w / 2) * (h / 2);
    HB_ARG(ctx, src_frame_bytes >= luma + 2 * chroma && dst_frame_bytes >= 2 * (luma + 2 * chroma));
    const uint8_t *y = (const uint8_t *)src, *u = y + luma, *v = u + chroma;
    uint8_t *dy = (uint8_t *)dst, *duv = dy + 2 * luma;
    const int aligned = is_aligned16(src | dst | (uint64_t)luma | (uint64_t)chroma | (uint64_t)(w / 2) | (uint64_t)src_frame_bytes | (uint64_t)dst_frame_bytes);
    const long long items = (long long)((w + 15) / 16) * h + (long long)((w / 2 + 15) / 16) * (h / 2);
    k_pack_p010<<<dim3(hb_grid_for(ctx, items, 256, 8), n_frames), 256, 0, ctx->stream>>>(y, w, u, w / 2, v, w / 2, w, h, dy, 2 * w, duv, 2 * w, aligned,
                                                                                          src_frame_bytes, dst_frame_bytes);
    HB_LAUNCHED(ctx);
    return HB_OK;
}

int hb_rgb_to_p010_batch(hb_ctx *ctx, hb_devptr rgb, size_t rgb_frame_bytes, int order, int matrix, int w, int h, hb_devptr dst,
                         size_t dst_frame_bytes, int n_frames)
{
    HB_ARG(ctx, ctx && rgb && dst && w > 0 && h > 0 && !(w & 1) && !(h & 1) && n_frames >= 1 && n_frames <= 65535);
    HB_ARG(ctx, rgb_frame_bytes >= (size_t)3 * w * h && dst_frame_bytes >= (size_t)3 * w * h);
    CscCoef c;
    if (!csc_coefficients(matrix, 10, c)) return hb_fail(ctx, HB_ERR_ARG, "bad argument: %s", "matrix");
    uint8_t *dy = (uint8_t *)dst, *duv = dy + (size_t)2 * w * h;
    const int aligned = is_aligned16(rgb | dst | (uint64_t)(3 * w) | (uint64_t)(2 * w) | (uint64_t)rgb_frame_bytes | (uint64_t)dst_frame_bytes | (uint64_t)((size_t)2 * w * h));
    const long long items = (long long)((w + 15) / 16) * (h / 2);
    k_rgb_to_yuv420<10><<<dim3(hb_grid_for(ctx, items, 128, 8), n_frames), 128, 0, ctx->stream>>>((const uint8_t *)rgb, 3 * w, order == HB_BGR, c, w, h, dy, 2 * w, duv,
                                                                                                 2 * w, nullptr, 0, aligned, rgb_frame_bytes, dst_frame_bytes);
    HB_LAUNCHED(ctx);
    return HB_OK;
}

int hb_scale_yuv420_to_p010_batch(hb_ctx *ctx, hb_devptr src, size_t src_frame_bytes, int sw, int sh, hb_devptr dst, size_t dst_frame_bytes, int dw,
                                  int dh, int n_frames)
{
    HB_ARG(ctx, ctx && src && dst && sw > 0 && sh > 0 && dw > 0 && dh > 0 && !(sw & 1) && !(sh & 1) && !(dw & 1) && !(dh & 1) && n_frames >= 1 && n_frames <= 65535);
    const size_t sl = (size_t)sw * sh, sc = (size_t)(sw / 2) * (sh / 2), dl = (size_t)dw * dh;
    HB_ARG(ctx, src_frame_bytes >= sl + 2 * sc && dst_frame_bytes >= 3 * dl);
    const uint8_t *y = (const uint8_t *)src;
    uint8_t *dy = (uint8_t *)dst;
    int rc = hb::launch_scale8(ctx, y, nullptr, sw, sw, sh, dy, nullptr, 2 * dw, dw, dh, 10, 6, 1, n_frames, src_frame_bytes, dst_frame_bytes);
    if (rc) return rc;
    return hb::launch_scale8(ctx, y + sl, y + sl + sc, sw / 2, sw / 2, sh / 2, dy + 2 * dl, nullptr, 2 * dw, dw / 2, dh / 2, 10, 6, 2, n_frames, src_frame_bytes,
                             dst_frame_bytes);
}

}  // extern "C"
