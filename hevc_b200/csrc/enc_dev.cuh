// Encoder-wide device/host structures.  Encoder specification: DESIGN.md ("encoder specification"); the
// CPU model oracle/hevc_encode.c must produce byte-identical output.
#pragma once
#include <stdint.h>

#include "prim_dev.cuh"

namespace hb {

constexpr int kCtu = 32, kCu = 16;
constexpr int kPad = 80;            // luma border of reconstructed planes (chroma: 40)
constexpr int kCuCoefs = 384;       // 16x16 luma + 8x8 Cb + 8x8 Cr levels per CU
constexpr int kCmeRange = 12;       // quarter-resolution search range
constexpr int kMvOverhang = 64;     // predicted blocks may leave the picture by this many luma samples
constexpr int kNumCtx = 142;

struct CuInfo {
    uint8_t pred_mode;   // 0 intra, 1 inter
    uint8_t intra_mode;
    uint8_t cbf;         // bit0 Y, bit1 Cb, bit2 Cr
    uint8_t skip;
    int16_t mvx, mvy;    // quarter-sample units
};

// per-CU syntax decisions made by the mode kernel for the entropy coder
struct CuSyntax {
    int8_t merge_idx;    // -1: AMVP
    uint8_t skip;
    uint8_t mvp_idx;
    uint8_t pad;
    int16_t mvdx, mvdy;
};

struct Geom {
    int wc, hc;          // coded size (multiples of 16)
    int cuw, cuh, ctuw, ctuh;
    int bit_depth;
    int src_stride, srcc_stride;      // source planes (samples)
    int rec_stride, recc_stride;      // padded reconstruction planes (samples)
    int dsw, dsh;                     // quarter-resolution plane
};

struct Planes {
    pixel *y, *u, *v;    // pointers to sample (0,0)
};

// round(256 * sqrt(0.57 * 2^((qp - 12) / 3)))
HB_HD constexpr int lambda_q8(int qp)
{
    constexpr int t[52] = {48, 54, 61, 68, 77, 86, 97, 108, 122, 137, 153, 172, 193, 217, 244, 273, 307, 344, 387, 434, 487, 547, 614,
                           689, 773, 868, 974, 1093, 1227, 1378, 1546, 1736, 1948, 2187, 2454, 2755, 3092, 3471, 3896, 4373, 4909, 5510,
                           6185, 6942, 7792, 8747, 9818, 11020, 12370, 13884, 15585, 17493};
    return t[qp];
}

HB_HD constexpr int chroma_qp(int qp_y)
{
    constexpr int t[14] = {29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37};
    const int q = qp_y < 0 ? 0 : qp_y > 57 ? 57 : qp_y;
    return q < 30 ? q : q >= 44 ? q - 6 : t[q - 30];
}

HB_HD int mv_bits1(int v)
{
    const int a = v < 0 ? -v : v;
    int n = 0;
    while ((a + 1) >> (n + 1)) n++;
    return 2 * n + 1;
}

HB_HD int mv_cost(int lambda, int mvx, int mvy, int px, int py) { return (lambda * (mv_bits1(mvx - px) + mv_bits1(mvy - py))) >> 8; }

// decode order of CU (cx, cy): CTU raster address * 4 + z index
HB_HD int cu_order(int ctuw, int cx, int cy) { return ((cy >> 1) * ctuw + (cx >> 1)) * 4 + ((cy & 1) << 1) + (cx & 1); }
HB_HD bool cu_avail(const Geom &g, int cx, int cy, int nx, int ny)
{
    if (nx < 0 || ny < 0 || nx >= g.cuw || ny >= g.cuh) return false;
    return cu_order(g.ctuw, nx, ny) < cu_order(g.ctuw, cx, cy);
}

}  // namespace hb
