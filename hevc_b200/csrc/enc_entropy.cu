// Entropy stage for sm_100a: per-CU syntax decisions (fully parallel), then WPP CABAC -- one CTA per frame,
// one warp per CTU row (rows interleaved over 32 warps), the 32 lanes stage coefficient levels and sub-block
// significance masks in shared memory while lane 0 runs the serial arithmetic coder.  Rows synchronise only
// through the context snapshot taken after the second CTU of the row above (H.265 9.3.2.2).
// Byte-identical to oracle/hevc_cabac.c.
#include "enc_kernels.cuh"

namespace hb {

// ------------------------------------------------------------------------------------------------ tables (H.265 9.3)
__constant__ uint8_t c_range_lps[64][4] = {
    {128, 176, 208, 240}, {128, 167, 197, 227}, {128, 158, 187, 216}, {123, 150, 178, 205}, {116, 142, 169, 195}, {111, 135, 160, 185},
    {105, 128, 152, 175}, {100, 122, 144, 166}, {95, 116, 137, 158}, {90, 110, 130, 150}, {85, 104, 123, 142}, {81, 99, 117, 135},
    {77, 94, 111, 128}, {73, 89, 105, 122}, {69, 85, 100, 116}, {66, 80, 95, 110}, {62, 76, 90, 104}, {59, 72, 86, 99}, {56, 69, 81, 94},
    {53, 65, 77, 89}, {51, 62, 73, 85}, {48, 59, 69, 80}, {46, 56, 66, 76}, {43, 53, 63, 72}, {41, 50, 59, 69}, {39, 48, 56, 65},
    {37, 45, 54, 62}, {35, 43, 51, 59}, {33, 41, 48, 56}, {32, 39, 46, 53}, {30, 37, 43, 50}, {29, 35, 41, 48}, {27, 33, 39, 45},
    {26, 31, 37, 43}, {24, 30, 35, 41}, {23, 28, 33, 39}, {22, 27, 32, 37}, {21, 26, 30, 35}, {20, 24, 29, 33}, {19, 23, 27, 31},
    {18, 22, 26, 30}, {17, 21, 25, 28}, {16, 20, 23, 27}, {15, 19, 22, 25}, {14, 18, 21, 24}, {14, 17, 20, 23}, {13, 16, 19, 22},
    {12, 15, 18, 21}, {12, 14, 17, 20}, {11, 14, 16, 19}, {11, 13, 15, 18}, {10, 12, 15, 17}, {10, 12, 14, 16}, {9, 11, 13, 15},
    {9, 11, 12, 14}, {8, 10, 12, 14}, {8, 9, 11, 13}, {7, 9, 11, 12}, {7, 9, 10, 12}, {7, 8, 10, 11}, {6, 8, 9, 11}, {6, 7, 9, 10},
    {6, 7, 8, 9}, {2, 2, 2, 2}};
__constant__ uint8_t c_next_lps[64] = {0, 0, 1, 2, 2, 4, 4, 5, 6, 7, 8, 9, 9, 11, 11, 12, 13, 13, 15, 15, 16, 16, 18, 18, 19, 19, 21, 21, 22, 22, 23, 24,
                                       24, 25, 26, 26, 27, 27, 28, 29, 29, 30, 30, 30, 31, 32, 32, 33, 33, 33, 34, 34, 35, 35, 35, 36, 36, 36, 37, 37, 37, 38, 38, 63};

enum {
    CX_SPLIT_CU = 0, CX_SKIP = 3, CX_PRED_MODE = 6, CX_PART_MODE = 7, CX_PREV_INTRA = 11, CX_CHROMA_PRED = 12, CX_MERGE_FLAG = 13,
    CX_MERGE_IDX = 14, CX_MVD_GR0 = 15, CX_MVD_GR1 = 16, CX_MVP_FLAG = 17, CX_ROOT_CBF = 18, CX_SPLIT_TU = 19, CX_CBF_LUMA = 22,
    CX_CBF_CHROMA = 24, CX_LAST_X = 28, CX_LAST_Y = 46, CX_CSBF = 64, CX_SIG = 68, CX_GR1 = 110, CX_GR2 = 134, CX_QP_DELTA = 140
};
static_assert(CX_QP_DELTA + 2 == kNumCtx, "context layout");

// initValue per context for initType 0 (I) and 1 (P); H.265 Tables 9-5 .. 9-37
__constant__ uint8_t c_ctx_init[2][kNumCtx] = {
    {139, 141, 157, 154, 154, 154, 154, 184, 154, 154, 154, 184, 63, 154, 154, 154, 154, 154, 154, 153, 138, 138, 111, 141, 94, 138, 182, 154,
     110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63,
     110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63,
     91, 171, 134, 141,
     111, 111, 125, 110, 110, 94, 124, 108, 124, 107, 125, 141, 179, 153, 125, 107, 125, 141, 179, 153, 125, 107, 125, 141, 179, 153, 125,
     140, 139, 182, 182, 152, 136, 152, 136, 153, 136, 139, 111, 136, 139, 111,
     140, 92, 137, 138, 140, 152, 138, 139, 153, 74, 149, 92, 139, 107, 122, 152, 140, 179, 166, 182, 140, 227, 122, 197,
     138, 153, 136, 167, 152, 152, 154, 154},
    {107, 139, 126, 197, 185, 201, 149, 154, 139, 154, 154, 154, 152, 110, 122, 140, 198, 168, 79, 124, 138, 94, 153, 111, 149, 107, 167, 154,
     125, 110, 94, 110, 95, 79, 125, 111, 110, 78, 110, 111, 111, 95, 94, 108, 123, 108,
     125, 110, 94, 110, 95, 79, 125, 111, 110, 78, 110, 111, 111, 95, 94, 108, 123, 108,
     121, 140, 61, 154,
     155, 154, 139, 153, 139, 123, 123, 63, 153, 166, 183, 140, 136, 153, 154, 166, 183, 140, 136, 153, 154, 166, 183, 140, 136, 153, 154,
     170, 153, 123, 123, 107, 121, 107, 121, 167, 151, 183, 140, 151, 183, 140,
     154, 196, 196, 167, 154, 152, 167, 182, 182, 134, 149, 136, 153, 121, 136, 137, 169, 194, 166, 167, 154, 167, 137, 182,
     107, 167, 91, 122, 107, 167, 154, 154}};

// up-right diagonal scans as raster indices (y * size + x)
__constant__ uint8_t c_diag4[16] = {0, 4, 1, 8, 5, 2, 12, 9, 6, 3, 13, 10, 7, 14, 11, 15};
__constant__ uint8_t c_diag2[4] = {0, 2, 1, 3};
__constant__ uint8_t c_group_idx[32] = {0, 1, 2, 3, 4, 4, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9, 9, 9};
__constant__ uint8_t c_min_in_group[10] = {0, 1, 2, 3, 4, 6, 8, 12, 16, 24};

// ------------------------------------------------------------------------------------------------ syntax decisions
struct MvPair { int x, y; };

__device__ __forceinline__ bool inter_at(const ModeParams &p, int cx, int cy, int nx, int ny, MvPair &mv)
{
    if (!cu_avail(p.g, cx, cy, nx, ny)) return false;
    const CuInfo c = p.cus[ny * p.g.cuw + nx];
    if (c.pred_mode != 1) return false;
    mv.x = c.mvx; mv.y = c.mvy;
    return true;
}

__device__ __forceinline__ int mvd_bits(int d)
{
    const int a = abs(d);
    if (a == 0) return 1;
    if (a == 1) return 3;
    int v = a - 2, k = 1, bits = 3;
    while (v >= (1 << k)) { v -= 1 << k; k++; bits++; }
    return bits + 1 + k;
}

// one thread per CU: merge index / skip / AMVP predictor + difference (P), or MPM index / remaining mode (I)
__global__ void __launch_bounds__(256) k_modes(ModeParams p)
{
    const Geom &g = p.g;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.cuw * g.cuh) return;
    const int cx = idx % g.cuw, cy = idx / g.cuw;
    const CuInfo me = p.cus[idx];
    CuSyntax out;
    out.merge_idx = -1; out.skip = 0; out.mvp_idx = 0; out.pad = 0; out.mvdx = 0; out.mvdy = 0;
    if (p.is_intra) {
        int a = 1, b = 1;
        if (cu_avail(g, cx, cy, cx - 1, cy) && p.cus[idx - 1].pred_mode == 0) a = p.cus[idx - 1].intra_mode;
        if ((cy & 1) && cu_avail(g, cx, cy, cx, cy - 1) && p.cus[idx - g.cuw].pred_mode == 0) b = p.cus[idx - g.cuw].intra_mode;
        int m0, m1, m2;
        if (a == b) {
            if (a < 2) { m0 = 0; m1 = 1; m2 = 26; }
            else { m0 = a; m1 = 2 + ((a + 29) & 31); m2 = 2 + ((a - 1) & 31); }
        } else {
            m0 = a; m1 = b;
            m2 = (a != 0 && b != 0) ? 0 : (a != 1 && b != 1) ? 1 : 26;
        }
        const int mode = me.intra_mode;
        if (mode == m0) out.merge_idx = 0;
        else if (mode == m1) out.merge_idx = 1;
        else if (mode == m2) out.merge_idx = 2;
        else {      // remaining mode: subtract the number of smaller candidates
            out.mvdx = (int16_t)(mode - (m0 < mode) - (m1 < mode) - (m2 < mode));
        }
        p.syn[idx] = out;
        return;
    }
    MvPair A1, B1, B0, A0, B2;
    const bool aA1 = inter_at(p, cx, cy, cx - 1, cy, A1), aB1 = inter_at(p, cx, cy, cx, cy - 1, B1);
    const bool aB0 = inter_at(p, cx, cy, cx + 1, cy - 1, B0), aA0 = inter_at(p, cx, cy, cx - 1, cy + 1, A0);
    const bool aB2 = inter_at(p, cx, cy, cx - 1, cy - 1, B2);
    auto same = [](const MvPair &u, const MvPair &v) { return u.x == v.x && u.y == v.y; };
    // merge list: A1, B1, B0, A0, B2 with the normative pruning, zero candidates after
    const bool fA1 = aA1, fB1 = aB1 && !(aA1 && same(A1, B1)), fB0 = aB0 && !(aB1 && same(B1, B0)), fA0 = aA0 && !(aA1 && same(A1, A0));
    const bool fB2 = aB2 && !(aA1 && same(A1, B2)) && !(aB1 && same(B1, B2)) && ((int)fA0 + fA1 + fB0 + fB1 != 4);
    const MvPair mine{me.mvx, me.mvy};
    int n = 0, found = -1;
    if (fA1) { if (found < 0 && same(A1, mine)) found = n; n++; }
    if (fB1) { if (found < 0 && same(B1, mine)) found = n; n++; }
    if (fB0) { if (found < 0 && same(B0, mine)) found = n; n++; }
    if (fA0) { if (found < 0 && same(A0, mine)) found = n; n++; }
    if (fB2) { if (found < 0 && same(B2, mine)) found = n; n++; }
    if (found < 0 && n < 5 && mine.x == 0 && mine.y == 0) found = n;
    if (found >= 0) {
        out.merge_idx = (int8_t)found;
        out.skip = me.cbf == 0;
    } else {
        // AMVP: A = first of (A0, A1), B = first of (B0, B1, B2); without A, A takes B's vector
        bool haveA = aA0 || aA1, haveB = aB0 || aB1 || aB2;
        MvPair a = aA0 ? A0 : A1, b = aB0 ? B0 : aB1 ? B1 : B2;
        if (!haveA && haveB) { a = b; haveA = true; }
        MvPair c[2] = {{0, 0}, {0, 0}};
        int k = 0;
        if (haveA) c[k++] = a;
        if (haveB && !(haveA && same(a, b))) c[k++] = b;
        const int b0 = mvd_bits(mine.x - c[0].x) + mvd_bits(mine.y - c[0].y);
        const int b1 = mvd_bits(mine.x - c[1].x) + mvd_bits(mine.y - c[1].y);
        const int sel = b1 < b0;
        out.mvp_idx = (uint8_t)sel;
        out.mvdx = (int16_t)(mine.x - c[sel].x);
        out.mvdy = (int16_t)(mine.y - c[sel].y);
    }
    p.syn[idx] = out;
}

// ------------------------------------------------------------------------------------------------ CABAC engine (lane 0)
struct Cabac {
    uint32_t low, range;
    int bits_left, buffered;
    uint32_t held;
    uint8_t *out;
    uint32_t pos, cap;
    uint8_t *ctx;          // shared memory, kNumCtx entries: (state << 1) | mps
};

__device__ __forceinline__ void cb_byte(Cabac &c, uint32_t v)
{
    if (c.pos < c.cap) c.out[c.pos] = (uint8_t)v;
    c.pos++;
}

__device__ __noinline__ void cb_write_out(Cabac &c)
{
    const uint32_t lead = c.low >> (24 - c.bits_left);
    c.bits_left += 8;
    c.low &= 0xffffffffu >> c.bits_left;
    if (lead == 0xff) {
        c.buffered++;
    } else if (c.buffered > 0) {
        const uint32_t carry = lead >> 8;
        cb_byte(c, c.held + carry);
        c.held = lead & 0xff;
        const uint32_t fill = (0xff + carry) & 0xff;
        while (c.buffered > 1) { cb_byte(c, fill); c.buffered--; }
    } else {
        c.buffered = 1;
        c.held = lead;
    }
}

// Not inlined on purpose: the syntax code calls these from ~100 sites; inlining them made a 45 k-instruction kernel that
// spent 57 % of its stall samples waiting for instruction fetch (profiles/round1_summary.md).
__device__ __noinline__ void cb_bin(Cabac &c, int ctx, int bin)
{
    const uint32_t s = c.ctx[ctx];
    const uint32_t st = s >> 1, mps = s & 1;
    const uint32_t lps = c_range_lps[st][(c.range >> 6) & 3];
    c.range -= lps;
    if ((uint32_t)bin != mps) {
        const int nb = __clz(lps) - 23;             // renormalisation shift: lps in [6, 240] -> 1..6
        c.low = (c.low + c.range) << nb;
        c.range = lps << nb;
        c.ctx[ctx] = (uint8_t)((c_next_lps[st] << 1) | (st == 0 ? 1 - mps : mps));
        c.bits_left -= nb;
    } else {
        c.ctx[ctx] = (uint8_t)(((st < 62 ? st + 1 : st) << 1) | mps);
        if (c.range >= 256) return;
        c.low <<= 1;
        c.range <<= 1;
        c.bits_left--;
    }
    if (c.bits_left < 12) cb_write_out(c);
}

__device__ __forceinline__ void cb_bypass_inl(Cabac &c, int bin)
{
    c.low <<= 1;
    if (bin) c.low += c.range;
    c.bits_left--;
    if (c.bits_left < 12) cb_write_out(c);
}

__device__ __noinline__ void cb_bypass(Cabac &c, int bin) { cb_bypass_inl(c, bin); }

__device__ __noinline__ void cb_bypass_bits(Cabac &c, uint32_t v, int n)
{
    for (int i = n - 1; i >= 0; i--) cb_bypass_inl(c, (v >> i) & 1);
}

__device__ __noinline__ void cb_terminate(Cabac &c, int bin)
{
    c.range -= 2;
    if (bin) {
        c.low += c.range;
        c.low <<= 7;
        c.range = 2 << 7;
        c.bits_left -= 7;
    } else if (c.range >= 256) {
        return;
    } else {
        c.low <<= 1;
        c.range <<= 1;
        c.bits_left--;
    }
    if (c.bits_left < 12) cb_write_out(c);
}

__device__ __noinline__ uint32_t cb_finish(Cabac &c)
{
    if (c.low >> (32 - c.bits_left)) {
        cb_byte(c, c.held + 1);
        while (c.buffered > 1) { cb_byte(c, 0x00); c.buffered--; }
        c.low -= 1u << (32 - c.bits_left);
    } else {
        if (c.buffered > 0) cb_byte(c, c.held);
        while (c.buffered > 1) { cb_byte(c, 0xff); c.buffered--; }
    }
    int n = 24 - c.bits_left + 1;
    unsigned long long v = ((unsigned long long)(c.low >> 8) << 1) | 1;   // remaining bits + rbsp stop bit
    const int pad = (8 - (n & 7)) & 7;
    v <<= pad;
    n += pad;
    for (int i = n - 8; i >= 0; i -= 8) cb_byte(c, (uint32_t)(v >> i) & 0xff);
    return c.pos;
}

// ------------------------------------------------------------------------------------------------ residual_coding (lane 0)
__device__ __noinline__ void write_remaining(Cabac &c, int value, int rice)
{
    if (value < (3 << rice)) {
        const int len = value >> rice;
        cb_bypass_bits(c, (1u << (len + 1)) - 2, len + 1);
        cb_bypass_bits(c, value & ((1 << rice) - 1), rice);
    } else {
        int len = rice;
        value -= 3 << rice;
        while (value >= (1 << len)) { value -= 1 << len; len++; }
        const int pre = 3 + len + 1 - rice;
        cb_bypass_bits(c, (1u << pre) - 2, pre);
        cb_bypass_bits(c, value, len);
    }
}

// lv: raster levels of the transform block in shared memory; masks: per sub-block (raster sub-block index)
// 16-bit significance mask in diagonal scan order.  log2n is 4 (luma) or 3 (chroma); diagonal scan only.
__device__ __noinline__ void residual_coding(Cabac &c, const int16_t *lv, const uint16_t *masks, int log2n, int c_idx)
{
    const int n = 1 << log2n, sbw = n >> 2, nsb = sbw * sbw;
    const uint8_t *sbscan = sbw == 4 ? c_diag4 : c_diag2;
    int last_sb = 0;
    for (int i = nsb - 1; i >= 0; i--)
        if (masks[sbscan[i]]) { last_sb = i; break; }
    const int last_mask = masks[sbscan[last_sb]];
    const int last_pos = 31 - __clz(last_mask);
    {
        const int sr = sbscan[last_sb], pr = c_diag4[last_pos];
        const int px = ((sr % sbw) << 2) + (pr & 3), py = ((sr / sbw) << 2) + (pr >> 2);
        const int gx = c_group_idx[px], gy = c_group_idx[py], cmax = c_group_idx[n - 1];
        int off, shift;
        if (c_idx == 0) { off = 3 * (log2n - 2) + ((log2n - 1) >> 2); shift = (log2n + 1) >> 2; }
        else { off = 15; shift = log2n - 2; }
        for (int i = 0; i < gx; i++) cb_bin(c, CX_LAST_X + off + (i >> shift), 1);
        if (gx < cmax) cb_bin(c, CX_LAST_X + off + (gx >> shift), 0);
        for (int i = 0; i < gy; i++) cb_bin(c, CX_LAST_Y + off + (i >> shift), 1);
        if (gy < cmax) cb_bin(c, CX_LAST_Y + off + (gy >> shift), 0);
        if (gx > 3) cb_bypass_bits(c, px - c_min_in_group[gx], (gx - 2) >> 1);
        if (gy > 3) cb_bypass_bits(c, py - c_min_in_group[gy], (gy - 2) >> 1);
    }
    int greater1_ctx = 1;
    for (int i = last_sb; i >= 0; i--) {
        const int sr = sbscan[i], xs = sr % sbw, ys = sr / sbw;
        const int mask = masks[sr];
        // coded_sub_block_flag of the right / below neighbours: coded (or inferred) sub-blocks are exactly those
        // with a non-zero mask, plus the DC sub-block and the last one which are inferred 1
        auto csbf_of = [&](int x, int y) -> int {
            if (x >= sbw || y >= sbw) return 0;
            const int r = y * sbw + x;
            return masks[r] != 0 || r == 0 || r == sbscan[last_sb];
        };
        const int right = csbf_of(xs + 1, ys), below = csbf_of(xs, ys + 1);
        int infer_dc = 0;
        if (i < last_sb && i > 0) {
            cb_bin(c, CX_CSBF + ((right | below) ? 1 : 0) + (c_idx ? 2 : 0), mask != 0);
            infer_dc = 1;
            if (!mask) continue;
        }
        const int start = i == last_sb ? last_pos : 15;
        const int prev_csbf = right | (below << 1);
        int abs_lv[16];
        uint32_t signs = 0;
        int cnt = 0;
        for (int k = start; k >= 0; k--) {
            const int pr = c_diag4[k], xp = pr & 3, yp = pr >> 2;
            const int x = (xs << 2) + xp, y = (ys << 2) + yp;
            const int sigf = (mask >> k) & 1;
            const bool is_last = i == last_sb && k == last_pos;
            if (!is_last && (k > 0 || !infer_dc)) {
                int sig;
                if (x == 0 && y == 0) {
                    sig = 0;
                } else {
                    if (prev_csbf == 0) sig = (xp + yp == 0) ? 2 : (xp + yp < 3) ? 1 : 0;
                    else if (prev_csbf == 1) sig = yp == 0 ? 2 : yp == 1 ? 1 : 0;
                    else if (prev_csbf == 2) sig = xp == 0 ? 2 : xp == 1 ? 1 : 0;
                    else sig = 2;
                    if (c_idx == 0) {
                        if (xs > 0 || ys > 0) sig += 3;
                        sig += log2n == 3 ? 9 : 21;
                    } else {
                        sig += log2n == 3 ? 9 : 12;
                    }
                }
                cb_bin(c, CX_SIG + (c_idx == 0 ? sig : 27 + sig), sigf);
                if (sigf) infer_dc = 0;
            }
            if (sigf) {
                const int v = lv[y * n + x];
                abs_lv[cnt++] = abs(v);
                signs = (signs << 1) | (v < 0 ? 1u : 0u);
            }
        }
        if (!cnt) continue;
        int ctx_set = (i > 0 && c_idx == 0) ? 2 : 0;
        if (i != last_sb && greater1_ctx == 0) ctx_set++;
        greater1_ctx = 1;
        int first_g1 = -1;
        const int n_g1 = cnt < 8 ? cnt : 8;
        for (int k = 0; k < n_g1; k++) {
            const int g1 = abs_lv[k] > 1;
            cb_bin(c, CX_GR1 + (ctx_set << 2) + greater1_ctx + (c_idx ? 16 : 0), g1);
            if (g1) {
                greater1_ctx = 0;
                if (first_g1 < 0) first_g1 = k;
            } else if (greater1_ctx > 0 && greater1_ctx < 3) {
                greater1_ctx++;
            }
        }
        if (first_g1 >= 0) cb_bin(c, CX_GR2 + ctx_set + (c_idx ? 4 : 0), abs_lv[first_g1] > 2);
        cb_bypass_bits(c, signs, cnt);
        int rice = 0;
        for (int k = 0; k < cnt; k++) {
            const int base = k < 8 ? (k == first_g1 ? 3 : 2) : 1;
            if (abs_lv[k] >= base) {
                write_remaining(c, abs_lv[k] - base, rice);
                if (abs_lv[k] > 3 * (1 << rice)) rice = rice < 4 ? rice + 1 : 4;
            }
        }
    }
}

__device__ __noinline__ void write_mvd(Cabac &c, int dx, int dy)
{
    const int ax = abs(dx), ay = abs(dy);
    cb_bin(c, CX_MVD_GR0, ax > 0);
    cb_bin(c, CX_MVD_GR0, ay > 0);
    if (ax > 0) cb_bin(c, CX_MVD_GR1, ax > 1);
    if (ay > 0) cb_bin(c, CX_MVD_GR1, ay > 1);
    for (int comp = 0; comp < 2; comp++) {
        const int a = comp ? ay : ax, neg = (comp ? dy : dx) < 0;
        if (a == 0) continue;
        if (a > 1) {
            int v = a - 2, k = 1;
            while (v >= (1 << k)) { cb_bypass(c, 1); v -= 1 << k; k++; }
            cb_bypass(c, 0);
            cb_bypass_bits(c, v, k);
        }
        cb_bypass(c, neg);
    }
}

// ------------------------------------------------------------------------------------------------ WPP CABAC kernel
// grid = (ceil(ctuh / kEntropyWarps), frames); CTA = kEntropyWarps warps, one CTU row per warp.  The CTAs are deliberately small
// (4 warps, ~8 K registers, ~4 KB shared memory) so that they fit into the hole a retiring k_inter CTA leaves and really
// overlap the frame chain.  Rows hand the context snapshot down through global memory (sync area of the frame); CTAs of a
// grid are dispatched in index order, so the CTA owning row r-1 is always resident (or done) when row r waits for it.
__global__ void __launch_bounds__(kEntropyWarps * 32) k_entropy(EntropyParams p)
{
    __shared__ EntropyWarpScratch ws[kEntropyWarps];
    const Geom &g = p.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const EntropyFrame fr = p.frames[blockIdx.y];
    uint8_t *ctx_save = fr.ctx_save;                              // [ctuh][kNumCtx]
    volatile int *row_ready = fr.row_ready;                       // [ctuh], zeroed before the launch
    EntropyWarpScratch &s = ws[warp];
    const int init_type = fr.is_intra ? 0 : 1;

    const int row = blockIdx.x * kEntropyWarps + warp;
    if (row < g.ctuh) {
        // ---- context initialisation: fresh for row 0 (or 1-CTU-wide pictures), else the snapshot of the row above
        if (row == 0 || g.ctuw < 2) {
            const int q = min(max(fr.ctl->qp, 0), 51);
            for (int i = lane; i < kNumCtx; i += 32) {
                const int v = c_ctx_init[init_type][i];
                const int m = (v >> 4) * 5 - 45, n = ((v & 15) << 3) - 16;
                const int pre = min(max(((m * q) >> 4) + n, 1), 126);
                const int mps = pre > 63;
                s.ctx[i] = (uint8_t)(((mps ? pre - 64 : 63 - pre) << 1) | mps);
            }
        } else {
            if (lane == 0) {
                unsigned ns = 200, spins = 0;     // back off: polling warps share issue slots with the coding warps
                while (!row_ready[row - 1]) {
                    __nanosleep(ns);
                    ns = ns < 4000 ? ns * 2 : 4000;
                    if (++spins > (1u << 24)) __trap();       // ~1 minute: a lost hand-off must fail loudly, not hang the device
                }
            }
            __syncwarp();
            __threadfence();
            for (int i = lane; i < kNumCtx; i += 32) s.ctx[i] = __ldcg(ctx_save + (row - 1) * kNumCtx + i);
        }
        __syncwarp();
        Cabac c;
        c.low = 0; c.range = 510; c.bits_left = 23; c.buffered = 0; c.held = 0xff;
        c.out = fr.out + (size_t)row * p.row_cap; c.pos = 0; c.cap = p.row_cap; c.ctx = s.ctx;

        for (int x = 0; x < g.ctuw; x++) {
            const int cx0 = 2 * x, cy0 = 2 * row;
            for (int k = 0; k < 4; k++) {
                const int cx = cx0 + (k & 1), cy = cy0 + (k >> 1);
                if (cx >= g.cuw || cy >= g.cuh) continue;
                const int idx = cy * g.cuw + cx;
                const CuInfo cu = fr.cus[idx];
                const CuSyntax sy = fr.syn[idx];
                // ---- all lanes: stage this CU's levels and build the sub-block significance masks
                if (cu.cbf) {
                    const uint2 *src = reinterpret_cast<const uint2 *>(fr.coefs + (size_t)idx * kCuCoefs);
                    uint2 *dst = reinterpret_cast<uint2 *>(s.lv);
                    for (int i = lane; i < kCuCoefs / 4; i += 32) dst[i] = src[i];
                    __syncwarp();
                    if (lane < 24) {
                        const int16_t *blk;
                        int stride, sx, sy2;
                        if (lane < 16) { blk = s.lv; stride = 16; sx = (lane & 3) * 4; sy2 = (lane >> 2) * 4; }
                        else { const int q = lane - 16; blk = s.lv + 256 + (q >> 2) * 64; stride = 8; sx = (q & 1) * 4; sy2 = ((q >> 1) & 1) * 4; }
                        uint32_t m = 0;
#pragma unroll
                        for (int t = 0; t < 16; t++) {
                            const int pr = c_diag4[t];
                            m |= (blk[(sy2 + (pr >> 2)) * stride + sx + (pr & 3)] != 0 ? 1u : 0u) << t;
                        }
                        s.masks[lane] = (uint16_t)m;
                    }
                    __syncwarp();
                }
                // ---- lane 0: syntax elements
                if (lane == 0) {
                    if (k == 0 || (cx == cx0 && cy == cy0)) {
                        if (32 * x + 32 <= g.wc && 32 * row + 32 <= g.hc) {
                            const int inc = (cu_avail(g, cx0, cy0, cx0 - 1, cy0) ? 1 : 0) + (cu_avail(g, cx0, cy0, cx0, cy0 - 1) ? 1 : 0);
                            cb_bin(c, CX_SPLIT_CU + inc, 1);
                        }
                    }
                    cb_bin(c, CX_SPLIT_CU, 0);
                    const int cb_y = cu.cbf & 1, cb_u = (cu.cbf >> 1) & 1, cb_v = (cu.cbf >> 2) & 1;
                    bool coded_residual = true;
                    if (!fr.is_intra) {
                        const int availL = cu_avail(g, cx, cy, cx - 1, cy), availA = cu_avail(g, cx, cy, cx, cy - 1);
                        const int ctx = (availL && fr.syn[idx - 1].skip ? 1 : 0) + (availA && fr.syn[idx - g.cuw].skip ? 1 : 0);
                        cb_bin(c, CX_SKIP + ctx, sy.skip);
                        if (sy.merge_idx >= 0) {
                            if (!sy.skip) {
                                cb_bin(c, CX_PRED_MODE, 0);
                                cb_bin(c, CX_PART_MODE, 1);
                                cb_bin(c, CX_MERGE_FLAG, 1);
                            }
                            cb_bin(c, CX_MERGE_IDX, sy.merge_idx > 0);
                            if (sy.merge_idx > 0)
                                for (int t = 1; t < 4; t++) {
                                    cb_bypass(c, sy.merge_idx > t);
                                    if (sy.merge_idx <= t) break;
                                }
                            if (sy.skip) coded_residual = false;
                        } else {
                            cb_bin(c, CX_PRED_MODE, 0);
                            cb_bin(c, CX_PART_MODE, 1);
                            cb_bin(c, CX_MERGE_FLAG, 0);
                            write_mvd(c, sy.mvdx, sy.mvdy);
                            cb_bin(c, CX_MVP_FLAG, sy.mvp_idx);
                            cb_bin(c, CX_ROOT_CBF, cu.cbf != 0);
                            if (!cu.cbf) coded_residual = false;
                        }
                    } else {
                        cb_bin(c, CX_PREV_INTRA, sy.merge_idx >= 0);
                        if (sy.merge_idx >= 0) {
                            cb_bypass(c, sy.merge_idx > 0);
                            if (sy.merge_idx > 0) cb_bypass(c, sy.merge_idx > 1);
                        } else {
                            cb_bypass_bits(c, (uint32_t)sy.mvdx, 5);
                        }
                        cb_bin(c, CX_CHROMA_PRED, 0);
                    }
                    if (coded_residual) {
                        cb_bin(c, CX_CBF_CHROMA, cb_u);
                        cb_bin(c, CX_CBF_CHROMA, cb_v);
                        if (fr.is_intra || cb_u || cb_v) cb_bin(c, CX_CBF_LUMA + 1, cb_y);
                        if (cb_y) residual_coding(c, s.lv, s.masks, 4, 0);
                        if (cb_u) residual_coding(c, s.lv + 256, s.masks + 16, 3, 1);
                        if (cb_v) residual_coding(c, s.lv + 320, s.masks + 20, 3, 2);
                    }
                }
                __syncwarp();
            }
            // ---- end of CTU
            if (x == 1) {       // snapshot for the row below (taken before the terminating bin, contexts only)
                __syncwarp();
                for (int i = lane; i < kNumCtx; i += 32) ctx_save[row * kNumCtx + i] = s.ctx[i];
                __threadfence();
                __syncwarp();
                if (lane == 0) row_ready[row] = 1;
            }
            if (lane == 0) {
                const bool last_in_pic = row == g.ctuh - 1 && x == g.ctuw - 1;
                cb_terminate(c, last_in_pic);
                if (x == g.ctuw - 1 && !last_in_pic) cb_terminate(c, 1);
            }
        }
        if (lane == 0) {
            const uint32_t n = cb_finish(c);
            fr.row_len[row] = n;
            if (n > p.row_cap) atomicExch(p.overflow, 1);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ compaction
// single CTA: exclusive prefix sum over all (frame, row) lengths
__global__ void __launch_bounds__(1024) k_pack_scan(PackParams p)
{
    __shared__ uint32_t part[1024];
    const int total = p.n_frames * p.rows, tid = threadIdx.x;
    const int per = (total + 1023) / 1024;
    uint32_t sum = 0;
    for (int i = tid * per; i < min(total, (tid + 1) * per); i++) sum += p.frames[i / p.rows].row_len[i % p.rows];
    part[tid] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const uint32_t v = tid >= off ? part[tid - off] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    uint32_t run = tid ? part[tid - 1] : 0;
    for (int i = tid * per; i < min(total, (tid + 1) * per); i++) {
        p.offsets[i] = run;
        run += p.frames[i / p.rows].row_len[i % p.rows];
    }
    if (tid == 1023) p.offsets[total] = part[1023];
}

// one CTA per (frame, row): contiguous copy of the sub-stream
__global__ void __launch_bounds__(128) k_pack_copy(PackParams p)
{
    const int i = blockIdx.x, f = i / p.rows, r = i % p.rows;
    const uint8_t *src = p.frames[f].out + (size_t)r * p.row_cap;
    const uint32_t n = p.frames[f].row_len[r];
    uint8_t *dst = p.packed + p.offsets[i];
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) dst[k] = src[k];
}

}  // namespace hb
