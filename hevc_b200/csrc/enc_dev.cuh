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
constexpr int kNumCtx = 144;

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

// sample adaptive offset parameters of one CTU (oracle/hevc_model.h orc_sao); padded to 32 bytes for 16-byte staging copies
struct __align__(16) SaoCtu {
    uint8_t type[2];       // [0] luma, [1] chroma: 0 off, 1 band, 2 edge
    uint8_t eo_class[2];
    uint8_t band[3];
    int8_t offset[3][4];
    uint8_t pad[13];
};
static_assert(sizeof(SaoCtu) == 32, "SaoCtu layout");

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

HB_HD int mv_bits1(int v)      // 2 * floor(log2(|v| + 1)) + 1
{
    const unsigned a = (unsigned)(v < 0 ? -v : v) + 1;
#ifdef __CUDA_ARCH__
    return 2 * (31 - __clz((int)a)) + 1;
#else
    return 2 * (31 - __builtin_clz(a)) + 1;
#endif
}

HB_HD int mv_cost(int lambda, int mvx, int mvy, int px, int py) { return (lambda * (mv_bits1(mvx - px) + mv_bits1(mvy - py))) >> 8; }

// decode order of CU (cx, cy): CTU raster address * 4 + z index
HB_HD int cu_order(int ctuw, int cx, int cy) { return ((cy >> 1) * ctuw + (cx >> 1)) * 4 + ((cy & 1) << 1) + (cx & 1); }
HB_HD bool cu_avail(const Geom &g, int cx, int cy, int nx, int ny)
{
    if (nx < 0 || ny < 0 || nx >= g.cuw || ny >= g.cuh) return false;
    return cu_order(g.ctuw, nx, ny) < cu_order(g.ctuw, cx, cy);
}

// ---------------------------------------------------------------------------------------------- rate control
// Deterministic integer controller shared (as a specification) with oracle/hevc_rc.c: the QP of a frame is chosen on the
// device from the size *estimates* of earlier frames, so the frame chain never waits for the entropy coder or the host.
struct FrameCtl {
    int qp, lambda, is_idr, redo;
    int poc, scene_cut;            // picture order count decided on the device (scene cuts move the key-frame cadence)
    QuantParam qy, qc;
    unsigned long long est16;      // size estimate accumulated by the frame kernel, 1/16 bit
    unsigned long long satd_sum;   // sum of the luma SATDs after the first merge-aware pass (gate of the P-frame intra search)
    int n_cand, pad_;              // CUs on the intra-search work list of this frame
};

struct RcState {
    long long t16, b16, fullness;
    int have[2], qp_prev[2];
    long long est_prev[2];
    int qp_i, qp_p, rate_control, bit_depth;
    // frame-type state: key frames every `keyint` frames, earlier at a detected scene cut once `min_keyint` frames have passed
    int keyint, min_keyint, scenecut, poc, started;
    int cascade;              // hb_enc_params.qp_cascade
};

// scene-cut measures of one frame, accumulated by the coarse motion search (oracle/hevc_encode.c coarse_search / scene_cut)
struct SceneStat {
    unsigned long long inter, intra;
};

HB_HD bool scene_cut(const SceneStat &s, long long n_samples) { return (long long)s.inter >= 5 * n_samples && 2 * s.inter >= 3 * s.intra; }

HB_HD long long rc_pow2_sixth(int k)      // round(65536 * 2^(k/6))
{
    constexpr unsigned t[37] = {65536, 73562, 82570, 92682, 104032, 116772, 131072, 147123, 165140, 185364, 208064, 233544, 262144,
                                294247, 330281, 370728, 416128, 467088, 524288, 588493, 660561, 741455, 832255, 934175, 1048576,
                                1176987, 1321123, 1482910, 1664511, 1868350, 2097152, 2353974, 2642246, 2965821, 3329021, 3736700, 4194304};
    return t[k];
}

HB_HD int rc_step(long long est, long long budget)
{
    if (budget < 1) budget = 1;
    if (est > budget) {
        for (int k = 1; k <= 36; k++)
            if (est * 65536 <= budget * rc_pow2_sixth(k)) return k;
        return 36;
    }
    int j = 0;
    while (j < 12 && est * rc_pow2_sixth(j + 1) <= budget * 65536) j++;
    return -j;
}

HB_HD long long rc_budget(const RcState &rc, int is_idr)
{
    long long fill = rc.fullness + rc.t16;
    if (fill > rc.b16) fill = rc.b16;
    long long budget;
    if (is_idr) {
        budget = fill / 2;
        if (budget > 8 * rc.t16) budget = 8 * rc.t16;
    } else {
        budget = rc.t16 * fill / (rc.b16 / 2);
        if (budget < rc.t16 / 2) budget = rc.t16 / 2;
        if (budget > 2 * rc.t16) budget = 2 * rc.t16;
    }
    return budget;
}

// QP cascade over P frames (oracle/hevc_rc.c k_cascade): the anchor (poc % 4 == 0) at the P-frame QP, +4 / +2 / +4 between anchors
HB_HD int rc_cascade(const RcState &rc, int is_idr, int poc) { return !is_idr && rc.cascade ? (0x4240 >> (4 * (poc & 3))) & 15 : 0; }

// `poc` = frames since the last IDR of the frame being decided; qp_prev / est_prev are kept in anchor terms (oracle orc_rc_pick_qp)
HB_HD int rc_pick_qp(const RcState &rc, int is_idr, int poc)
{
    const int off = rc_cascade(rc, is_idr, poc);
    int base = (is_idr ? rc.qp_i : rc.qp_p) + off;
    if (base > 51) base = 51;
    if (!rc.rate_control) return base;
    const long long budget = rc_budget(rc, is_idr);
    const int t = is_idr ? 1 : 0;
    int qp = base;
    if (rc.have[t]) {
        const long long est = rc.est_prev[t] * 65536 / rc_pow2_sixth(off);
        int step = rc_step(est, budget);
        if (step < 0) step = est * 5 <= budget * 4 ? -1 : 0;
        qp = rc.qp_prev[t] + step + off;
    } else if (!is_idr && rc.have[1]) {
        int step = rc_step(rc.est_prev[1] / 4 * 65536 / rc_pow2_sixth(off), budget);
        if (step < 0) step = 0;
        qp = rc.qp_prev[1] + (rc.qp_p - rc.qp_i) + step + off;
    }
    return qp < base ? base : qp > 51 ? 51 : qp;
}

// `poc` = that of the frame just coded
HB_HD void rc_update(RcState &rc, int is_idr, int qp, long long est16, int poc)
{
    const int t = is_idr ? 1 : 0;
    const int off = rc_cascade(rc, is_idr, poc);
    rc.fullness += rc.t16;
    if (rc.fullness > rc.b16) rc.fullness = rc.b16;
    rc.fullness -= est16;
    if (rc.fullness < 0) rc.fullness = 0;
    rc.have[t] = 1;
    rc.qp_prev[t] = qp - off;
    rc.est_prev[t] = est16 * rc_pow2_sixth(off) / 65536;
}

HB_HD void ctl_set_qp(FrameCtl &c, int qp, int is_idr, int bit_depth)
{
    c.qp = qp;
    c.is_idr = is_idr;
    c.lambda = lambda_q8(qp) << (bit_depth - 8);
    c.qy = make_quant(4, qp + 6 * (bit_depth - 8), bit_depth, is_idr);
    c.qc = make_quant(3, chroma_qp(qp) + 6 * (bit_depth - 8), bit_depth, is_idr);
    c.est16 = 0;
    c.satd_sum = 0;
    c.n_cand = 0; c.pad_ = 0;
}

}  // namespace hb
