/* TEST INFRASTRUCTURE -- CPU oracle: CABAC engine (H.265 9.3) and slice-data syntax (7.3.8) for the subset of
 * tools this encoder uses: CTU 32, CU 16x16 2Nx2N, one 16x16 luma + two 8x8 chroma transform blocks per CU,
 * I and P slices with a single reference, merge / AMVP, WPP sub-streams.  Context initialisation values are
 * H.265 Tables 9-5..9-37; they were cross-checked against the tables inside the bundled libavcodec. */
#include <stdlib.h>
#include <string.h>

#include "hevc_model.h"

/* ------------------------------------------------------------------ tables */

static const uint8_t k_range_lps[64][4] = {
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
static const uint8_t k_next_lps[64] = {0, 0, 1, 2, 2, 4, 4, 5, 6, 7, 8, 9, 9, 11, 11, 12, 13, 13, 15, 15, 16, 16, 18, 18, 19, 19, 21, 21, 22, 22, 23, 24,
                                       24, 25, 26, 26, 27, 27, 28, 29, 29, 30, 30, 30, 31, 32, 32, 33, 33, 33, 34, 34, 35, 35, 35, 36, 36, 36, 37, 37, 37, 38, 38, 63};
static const uint8_t k_renorm[32] = {6, 5, 4, 4, 3, 3, 3, 3, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};

/* context layout */
enum {
    CX_SPLIT_CU = 0,        /* 3 */
    CX_SKIP = 3,            /* 3 */
    CX_PRED_MODE = 6,       /* 1 */
    CX_PART_MODE = 7,       /* 4 */
    CX_PREV_INTRA = 11,     /* 1 */
    CX_CHROMA_PRED = 12,    /* 1 */
    CX_MERGE_FLAG = 13,     /* 1 */
    CX_MERGE_IDX = 14,      /* 1 */
    CX_MVD_GR0 = 15,        /* 1 */
    CX_MVD_GR1 = 16,        /* 1 */
    CX_MVP_FLAG = 17,       /* 1 */
    CX_ROOT_CBF = 18,       /* 1 */
    CX_SPLIT_TU = 19,       /* 3 */
    CX_CBF_LUMA = 22,       /* 2 */
    CX_CBF_CHROMA = 24,     /* 4 */
    CX_LAST_X = 28,         /* 18 */
    CX_LAST_Y = 46,         /* 18 */
    CX_CSBF = 64,           /* 4 */
    CX_SIG = 68,            /* 42 */
    CX_GR1 = 110,           /* 24 */
    CX_GR2 = 134,           /* 6 */
    CX_QP_DELTA = 140,      /* 2 */
    CX_SAO_MERGE = 142,     /* 1 */
    CX_SAO_TYPE = 143,      /* 1 */
    CX_COUNT = 144
};

#define CNU 154
/* [initType][ctx]: initType 0 = I slice, 1 = P slice (cabac_init_flag = 0) */
static const uint8_t k_init[2][CX_COUNT] = {
    {   139, 141, 157,                                   /* split_cu_flag */
        CNU, CNU, CNU,                                   /* cu_skip_flag */
        CNU,                                             /* pred_mode_flag */
        184, CNU, CNU, CNU,                              /* part_mode */
        184,                                             /* prev_intra_luma_pred_flag */
        63,                                              /* intra_chroma_pred_mode */
        CNU, CNU,                                        /* merge_flag, merge_idx */
        CNU, CNU,                                        /* abs_mvd_greater0/1 */
        CNU,                                             /* mvp_lx_flag */
        CNU,                                             /* rqt_root_cbf */
        153, 138, 138,                                   /* split_transform_flag */
        111, 141,                                        /* cbf_luma */
        94, 138, 182, 154,                               /* cbf_cb / cbf_cr */
        110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63,   /* last x prefix */
        110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63,   /* last y prefix */
        91, 171, 134, 141,                               /* coded_sub_block_flag */
        111, 111, 125, 110, 110, 94, 124, 108, 124, 107, 125, 141, 179, 153, 125, 107, 125, 141, 179, 153, 125, 107, 125, 141, 179, 153, 125,
        140, 139, 182, 182, 152, 136, 152, 136, 153, 136, 139, 111, 136, 139, 111,                  /* sig_coeff_flag */
        140, 92, 137, 138, 140, 152, 138, 139, 153, 74, 149, 92, 139, 107, 122, 152, 140, 179, 166, 182, 140, 227, 122, 197,   /* greater1 */
        138, 153, 136, 167, 152, 152,                    /* greater2 */
        154, 154,                                        /* cu_qp_delta_abs */
        153, 200 },                                      /* sao_merge_left/up_flag, sao_type_idx */
    {   107, 139, 126,
        197, 185, 201,
        149,
        154, 139, 154, 154,
        154,
        152,
        110, 122,
        140, 198,
        168,
        79,
        124, 138, 94,
        153, 111,
        149, 107, 167, 154,
        125, 110, 94, 110, 95, 79, 125, 111, 110, 78, 110, 111, 111, 95, 94, 108, 123, 108,
        125, 110, 94, 110, 95, 79, 125, 111, 110, 78, 110, 111, 111, 95, 94, 108, 123, 108,
        121, 140, 61, 154,
        155, 154, 139, 153, 139, 123, 123, 63, 153, 166, 183, 140, 136, 153, 154, 166, 183, 140, 136, 153, 154, 166, 183, 140, 136, 153, 154,
        170, 153, 123, 123, 107, 121, 107, 121, 167, 151, 183, 140, 151, 183, 140,
        154, 196, 196, 167, 154, 152, 167, 182, 182, 134, 149, 136, 153, 121, 136, 137, 169, 194, 166, 167, 154, 167, 137, 182,
        107, 167, 91, 122, 107, 167,
        154, 154,
        153, 185 }};

/* ------------------------------------------------------------------ arithmetic encoder (HM TEncBinCABAC form of 9.3.4) */

typedef struct {
    uint32_t low, range;
    int bits_left, buffered;
    uint32_t held;                 /* buffered byte */
    uint8_t *out;
    size_t pos, cap;
    int overflow;
    uint8_t ctx[CX_COUNT];         /* (pStateIdx << 1) | valMps */
} cabac;

static void cb_byte(cabac *c, uint32_t v)
{
    if (c->pos < c->cap) c->out[c->pos] = (uint8_t)v;
    else c->overflow = 1;
    c->pos++;
}

static void cb_start(cabac *c, uint8_t *out, size_t cap)
{
    c->low = 0; c->range = 510; c->bits_left = 23; c->buffered = 0; c->held = 0xff;
    c->out = out; c->pos = 0; c->cap = cap; c->overflow = 0;
}

static void cb_init_contexts(cabac *c, int init_type, int qp)
{
    const int q = qp < 0 ? 0 : qp > 51 ? 51 : qp;
    for (int i = 0; i < CX_COUNT; i++) {
        const int v = k_init[init_type][i];
        const int m = (v >> 4) * 5 - 45, n = ((v & 15) << 3) - 16;
        int pre = ((m * q) >> 4) + n;
        pre = pre < 1 ? 1 : pre > 126 ? 126 : pre;
        const int mps = pre > 63;
        const int st = mps ? pre - 64 : 63 - pre;
        c->ctx[i] = (uint8_t)((st << 1) | mps);
    }
}

static void cb_write_out(cabac *c)
{
    const uint32_t lead = c->low >> (24 - c->bits_left);
    c->bits_left += 8;
    c->low &= 0xffffffffu >> c->bits_left;
    if (lead == 0xff) {
        c->buffered++;
    } else if (c->buffered > 0) {
        const uint32_t carry = lead >> 8;
        cb_byte(c, c->held + carry);
        c->held = lead & 0xff;
        const uint32_t fill = (0xff + carry) & 0xff;
        while (c->buffered > 1) { cb_byte(c, fill); c->buffered--; }
    } else {
        c->buffered = 1;
        c->held = lead;
    }
}

static void cb_bin(cabac *c, int ctx, int bin)
{
    uint8_t *s = &c->ctx[ctx];
    const int st = *s >> 1, mps = *s & 1;
    const uint32_t lps = k_range_lps[st][(c->range >> 6) & 3];
    c->range -= lps;
    if (bin != mps) {
        const int nb = k_renorm[lps >> 3];
        c->low = (c->low + c->range) << nb;
        c->range = lps << nb;
        *s = (uint8_t)((k_next_lps[st] << 1) | (st == 0 ? 1 - mps : mps));
        c->bits_left -= nb;
    } else {
        *s = (uint8_t)(((st < 62 ? st + 1 : st) << 1) | mps);
        if (c->range >= 256)
            return;
        c->low <<= 1;
        c->range <<= 1;
        c->bits_left--;
    }
    if (c->bits_left < 12)
        cb_write_out(c);
}

static void cb_bypass(cabac *c, int bin)
{
    c->low <<= 1;
    if (bin) c->low += c->range;
    c->bits_left--;
    if (c->bits_left < 12)
        cb_write_out(c);
}

static void cb_bypass_bits(cabac *c, uint32_t v, int n)
{
    for (int i = n - 1; i >= 0; i--)
        cb_bypass(c, (v >> i) & 1);
}

static void cb_terminate(cabac *c, int bin)
{
    c->range -= 2;
    if (bin) {
        c->low += c->range;
        c->low <<= 7;
        c->range = 2 << 7;
        c->bits_left -= 7;
    } else if (c->range >= 256) {
        return;
    } else {
        c->low <<= 1;
        c->range <<= 1;
        c->bits_left--;
    }
    if (c->bits_left < 12)
        cb_write_out(c);
}

/* flush after a terminating bin equal to 1, then rbsp stop bit + alignment; returns bytes written */
static size_t cb_finish(cabac *c)
{
    if (c->low >> (32 - c->bits_left)) {
        cb_byte(c, c->held + 1);
        while (c->buffered > 1) { cb_byte(c, 0x00); c->buffered--; }
        c->low -= 1u << (32 - c->bits_left);
    } else {
        if (c->buffered > 0) cb_byte(c, c->held);
        while (c->buffered > 1) { cb_byte(c, 0xff); c->buffered--; }
    }
    /* remaining (24 - bits_left) bits of low >> 8, then the '1' stop bit and zero padding */
    int n = 24 - c->bits_left;
    uint64_t v = ((uint64_t)(c->low >> 8) << 1) | 1;
    n += 1;
    const int pad = (8 - (n & 7)) & 7;
    v <<= pad;
    n += pad;
    for (int i = n - 8; i >= 0; i -= 8)
        cb_byte(c, (uint32_t)(v >> i) & 0xff);
    return c->pos;
}

/* ------------------------------------------------------------------ helpers on the CU grid */

static inline const orc_cu *cu_at(const orc_frame_syntax *f, int cx, int cy) { return &f->cu[cy * f->cuw + cx]; }

/* decode order of CU (cx, cy): CTU raster address * 4 + z index inside the CTU */
static inline int cu_order(const orc_frame_syntax *f, int cx, int cy)
{
    return ((cy >> 1) * f->ctuw + (cx >> 1)) * 4 + ((cy & 1) << 1) + (cx & 1);
}

/* 6.4.1 z-scan availability of neighbour CU (nx, ny) seen from (cx, cy) */
static inline int cu_avail(const orc_frame_syntax *f, int cx, int cy, int nx, int ny)
{
    if (nx < 0 || ny < 0 || nx >= f->cuw || ny >= f->cuh)
        return 0;
    return cu_order(f, nx, ny) < cu_order(f, cx, cy);
}

static inline int inter_avail(const orc_frame_syntax *f, int cx, int cy, int nx, int ny)
{
    return cu_avail(f, cx, cy, nx, ny) && cu_at(f, nx, ny)->pred_mode == 1;
}

static inline int same_mv(const orc_cu *a, const orc_cu *b) { return a->mvx == b->mvx && a->mvy == b->mvy; }

/* 8.5.3.2.2-3: merge candidates of a 16x16 2Nx2N PU in a P slice with one reference, no temporal candidate */
int orc_merge_candidates(const orc_frame_syntax *f, int cx, int cy, int16_t cand[5][2])
{
    int n = 0;
    const int aA1 = inter_avail(f, cx, cy, cx - 1, cy), aB1 = inter_avail(f, cx, cy, cx, cy - 1);
    const int aB0 = inter_avail(f, cx, cy, cx + 1, cy - 1), aA0 = inter_avail(f, cx, cy, cx - 1, cy + 1);
    const int aB2 = inter_avail(f, cx, cy, cx - 1, cy - 1);
    const orc_cu *A1 = aA1 ? cu_at(f, cx - 1, cy) : NULL, *B1 = aB1 ? cu_at(f, cx, cy - 1) : NULL;
    const orc_cu *B0 = aB0 ? cu_at(f, cx + 1, cy - 1) : NULL, *A0 = aA0 ? cu_at(f, cx - 1, cy + 1) : NULL;
    const orc_cu *B2 = aB2 ? cu_at(f, cx - 1, cy - 1) : NULL;
    int fA1 = aA1, fB1 = aB1 && !(aA1 && same_mv(A1, B1));
    int fB0 = aB0 && !(aB1 && same_mv(B1, B0));
    int fA0 = aA0 && !(aA1 && same_mv(A1, A0));
    int fB2 = aB2 && !(aA1 && same_mv(A1, B2)) && !(aB1 && same_mv(B1, B2)) && (fA0 + fA1 + fB0 + fB1 != 4);
    if (fA1) { cand[n][0] = A1->mvx; cand[n][1] = A1->mvy; n++; }
    if (fB1) { cand[n][0] = B1->mvx; cand[n][1] = B1->mvy; n++; }
    if (fB0) { cand[n][0] = B0->mvx; cand[n][1] = B0->mvy; n++; }
    if (fA0) { cand[n][0] = A0->mvx; cand[n][1] = A0->mvy; n++; }
    if (fB2) { cand[n][0] = B2->mvx; cand[n][1] = B2->mvy; n++; }
    while (n < 5) { cand[n][0] = 0; cand[n][1] = 0; n++; }   /* zero candidates, refIdx 0 */
    return 5;
}

/* 8.5.3.2.6-7: AMVP candidates (same single reference everywhere, so no scaling) */
int orc_amvp_candidates(const orc_frame_syntax *f, int cx, int cy, int16_t cand[2][2])
{
    const int aA0 = inter_avail(f, cx, cy, cx - 1, cy + 1), aA1 = inter_avail(f, cx, cy, cx - 1, cy);
    const int aB0 = inter_avail(f, cx, cy, cx + 1, cy - 1), aB1 = inter_avail(f, cx, cy, cx, cy - 1);
    const int aB2 = inter_avail(f, cx, cy, cx - 1, cy - 1);
    int haveA = 0, haveB = 0;
    int16_t a[2] = {0, 0}, b[2] = {0, 0};
    if (aA0) { a[0] = cu_at(f, cx - 1, cy + 1)->mvx; a[1] = cu_at(f, cx - 1, cy + 1)->mvy; haveA = 1; }
    else if (aA1) { a[0] = cu_at(f, cx - 1, cy)->mvx; a[1] = cu_at(f, cx - 1, cy)->mvy; haveA = 1; }
    const orc_cu *bc = aB0 ? cu_at(f, cx + 1, cy - 1) : aB1 ? cu_at(f, cx, cy - 1) : aB2 ? cu_at(f, cx - 1, cy - 1) : NULL;
    if (bc) { b[0] = bc->mvx; b[1] = bc->mvy; haveB = 1; }
    if (!(aA0 || aA1) && haveB) {           /* isScaledFlag == 0: A takes B's vector, B is re-derived (identical here) */
        a[0] = b[0]; a[1] = b[1]; haveA = 1;
    }
    int n = 0;
    if (haveA) { cand[n][0] = a[0]; cand[n][1] = a[1]; n++; }
    if (haveB && !(haveA && a[0] == b[0] && a[1] == b[1])) { cand[n][0] = b[0]; cand[n][1] = b[1]; n++; }
    while (n < 2) { cand[n][0] = 0; cand[n][1] = 0; n++; }
    return 2;
}

/* ------------------------------------------------------------------ residual_coding (7.3.8.11) */

static const uint8_t k_group_idx[32] = {0, 1, 2, 3, 4, 4, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9, 9, 9};
static const uint8_t k_min_in_group[10] = {0, 1, 2, 3, 4, 6, 8, 12, 16, 24};
static const uint8_t k_sig_ctx_4x4[16] = {0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8};

/* up-right diagonal scan of a blk x blk array (6.5.3): out[i] = (y << 4) | x */
static void diag_scan(int blk, uint8_t *out)
{
    int i = 0, x = 0, y = 0, stop = 0;
    while (!stop) {
        while (y >= 0) {
            if (x < blk && y < blk)
                out[i++] = (uint8_t)((y << 4) | x);
            y--; x++;
        }
        y = x; x = 0;
        if (i >= blk * blk) stop = 1;
    }
}

static void make_scan(int blk, int scan_idx, uint8_t *out)
{
    if (scan_idx == 0) { diag_scan(blk, out); return; }
    for (int i = 0; i < blk * blk; i++) {
        const int a = i / blk, b = i % blk;
        out[i] = scan_idx == 1 ? (uint8_t)((a << 4) | b) : (uint8_t)((b << 4) | a);   /* horizontal : vertical */
    }
}

static void write_remaining(cabac *c, int value, int rice)
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

/* lv: raster [n][n] levels; at least one non-zero */
static void residual_coding(cabac *c, const int16_t *lv, int log2n, int c_idx, int scan_idx)
{
    const int n = 1 << log2n, sbw = n >> 2;
    uint8_t sb_scan[64], pos_scan[16], csbf[8][8];
    make_scan(sbw, scan_idx, sb_scan);
    make_scan(4, scan_idx, pos_scan);
    memset(csbf, 0, sizeof csbf);

    int last_sb = -1, last_pos = -1;
    for (int i = 0; i < sbw * sbw; i++) {
        const int xs = sb_scan[i] & 15, ys = sb_scan[i] >> 4;
        for (int k = 0; k < 16; k++) {
            const int x = (xs << 2) + (pos_scan[k] & 15), y = (ys << 2) + (pos_scan[k] >> 4);
            if (lv[y * n + x]) { last_sb = i; last_pos = k; csbf[ys][xs] = 1; }
        }
    }
    int last_x = ((sb_scan[last_sb] & 15) << 2) + (pos_scan[last_pos] & 15);
    int last_y = ((sb_scan[last_sb] >> 4) << 2) + (pos_scan[last_pos] >> 4);
    {
        int px = last_x, py = last_y;
        if (scan_idx == 2) { int t = px; px = py; py = t; }
        const int gx = k_group_idx[px], gy = k_group_idx[py], cmax = k_group_idx[n - 1];
        int off, shift;
        if (c_idx == 0) { off = 3 * (log2n - 2) + ((log2n - 1) >> 2); shift = (log2n + 1) >> 2; }
        else { off = 15; shift = log2n - 2; }
        for (int i = 0; i < gx; i++) cb_bin(c, CX_LAST_X + off + (i >> shift), 1);
        if (gx < cmax) cb_bin(c, CX_LAST_X + off + (gx >> shift), 0);
        for (int i = 0; i < gy; i++) cb_bin(c, CX_LAST_Y + off + (i >> shift), 1);
        if (gy < cmax) cb_bin(c, CX_LAST_Y + off + (gy >> shift), 0);
        if (gx > 3) cb_bypass_bits(c, px - k_min_in_group[gx], (gx - 2) >> 1);
        if (gy > 3) cb_bypass_bits(c, py - k_min_in_group[gy], (gy - 2) >> 1);
    }

    int greater1_ctx = 1;
    for (int i = last_sb; i >= 0; i--) {
        const int xs = sb_scan[i] & 15, ys = sb_scan[i] >> 4;
        const int right = xs + 1 < sbw ? csbf[ys][xs + 1] : 0, below = ys + 1 < sbw ? csbf[ys + 1][xs] : 0;
        int infer_dc = 0;
        if (i < last_sb && i > 0) {
            cb_bin(c, CX_CSBF + ((right | below) ? 1 : 0) + (c_idx ? 2 : 0), csbf[ys][xs]);
            infer_dc = 1;
        } else {
            csbf[ys][xs] = 1;       /* inferred for the last and the DC sub-block */
        }
        if (!csbf[ys][xs])
            continue;
        /* significance */
        int abs_lv[16], sign[16], cnt = 0;
        const int start = i == last_sb ? last_pos : 15;
        const int prev_csbf = right | (below << 1);
        for (int k = start; k >= 0; k--) {
            const int xp = pos_scan[k] & 15, yp = pos_scan[k] >> 4;
            const int x = (xs << 2) + xp, y = (ys << 2) + yp;
            const int v = lv[y * n + x];
            const int is_last = i == last_sb && k == last_pos;
            if (!is_last && (k > 0 || !infer_dc)) {
                int sig;
                if (log2n == 2) {
                    sig = k_sig_ctx_4x4[(yp << 2) + xp];
                } else if (x == 0 && y == 0) {
                    sig = 0;
                } else {
                    if (prev_csbf == 0) sig = (xp + yp == 0) ? 2 : (xp + yp < 3) ? 1 : 0;
                    else if (prev_csbf == 1) sig = yp == 0 ? 2 : yp == 1 ? 1 : 0;
                    else if (prev_csbf == 2) sig = xp == 0 ? 2 : xp == 1 ? 1 : 0;
                    else sig = 2;
                    if (c_idx == 0) {
                        if (xs > 0 || ys > 0) sig += 3;
                        sig += log2n == 3 ? (scan_idx == 0 ? 9 : 15) : 21;
                    } else {
                        sig += log2n == 3 ? 9 : 12;
                    }
                }
                cb_bin(c, CX_SIG + (c_idx == 0 ? sig : 27 + sig), v != 0);
                if (v) infer_dc = 0;
            }
            if (v) { abs_lv[cnt] = abs(v); sign[cnt] = v < 0; cnt++; }
        }
        if (!cnt)
            continue;
        /* greater1 / greater2 */
        int ctx_set = (i > 0 && c_idx == 0) ? 2 : 0;
        if (i != last_sb && greater1_ctx == 0)
            ctx_set++;
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
        if (first_g1 >= 0)
            cb_bin(c, CX_GR2 + ctx_set + (c_idx ? 4 : 0), abs_lv[first_g1] > 2);
        uint32_t signs = 0;
        for (int k = 0; k < cnt; k++)
            signs = (signs << 1) | (uint32_t)sign[k];
        cb_bypass_bits(c, signs, cnt);
        int rice = 0;
        for (int k = 0; k < cnt; k++) {
            const int base = k < 8 ? (k == first_g1 ? 3 : 2) : 1;
            if (abs_lv[k] >= base) {
                write_remaining(c, abs_lv[k] - base, rice);
                if (abs_lv[k] > 3 * (1 << rice))
                    rice = rice < 4 ? rice + 1 : 4;
            }
        }
    }
}

/* ------------------------------------------------------------------ coding unit */

static int intra_mpm(const orc_frame_syntax *f, int cx, int cy, int mpm[3])
{
    int a = 1, b = 1;   /* DC when unavailable / not intra */
    if (cu_avail(f, cx, cy, cx - 1, cy) && cu_at(f, cx - 1, cy)->pred_mode == 0)
        a = cu_at(f, cx - 1, cy)->intra_mode;
    if ((cy & 1) && cu_avail(f, cx, cy, cx, cy - 1) && cu_at(f, cx, cy - 1)->pred_mode == 0)
        b = cu_at(f, cx, cy - 1)->intra_mode;     /* above must lie in the same CTU row */
    if (a == b) {
        if (a < 2) { mpm[0] = 0; mpm[1] = 1; mpm[2] = 26; }
        else { mpm[0] = a; mpm[1] = 2 + ((a + 29) % 32); mpm[2] = 2 + ((a - 2 + 1) % 32); }
    } else {
        mpm[0] = a; mpm[1] = b;
        mpm[2] = (a != 0 && b != 0) ? 0 : (a != 1 && b != 1) ? 1 : 26;
    }
    return 0;
}

/* exposed to the encoder for mode-bit estimation */
int orc_intra_mpm(const orc_frame_syntax *f, int cx, int cy, int mpm[3]) { return intra_mpm(f, cx, cy, mpm); }

static int mvd_bits(int d)
{
    const int a = abs(d);
    if (a == 0) return 1;
    if (a == 1) return 3;
    int v = a - 2, k = 1, bits = 3;      /* gr0 + gr1 + sign */
    while (v >= (1 << k)) { v -= 1 << k; k++; bits++; }
    return bits + 1 + k;
}

static void write_mvd(cabac *c, int dx, int dy)
{
    const int ax = abs(dx), ay = abs(dy);
    cb_bin(c, CX_MVD_GR0, ax > 0);
    cb_bin(c, CX_MVD_GR0, ay > 0);
    if (ax > 0) cb_bin(c, CX_MVD_GR1, ax > 1);
    if (ay > 0) cb_bin(c, CX_MVD_GR1, ay > 1);
    for (int comp = 0; comp < 2; comp++) {
        const int a = comp ? ay : ax, neg = (comp ? dy : dx) < 0;
        if (a == 0) continue;
        if (a > 1) {                      /* abs_mvd_minus2: EG1 */
            int v = a - 2, k = 1;
            while (v >= (1 << k)) { cb_bypass(c, 1); v -= 1 << k; k++; }
            cb_bypass(c, 0);
            cb_bypass_bits(c, v, k);
        }
        cb_bypass(c, neg);
    }
}

static void write_transform_unit(cabac *c, const orc_cu *cu, const int16_t *coef, int intra)
{
    const int cb_y = cu->cbf & 1, cb_u = (cu->cbf >> 1) & 1, cb_v = (cu->cbf >> 2) & 1;
    cb_bin(c, CX_CBF_CHROMA + 0, cb_u);       /* trafoDepth 0 */
    cb_bin(c, CX_CBF_CHROMA + 0, cb_v);
    if (intra || cb_u || cb_v)
        cb_bin(c, CX_CBF_LUMA + 1, cb_y);     /* ctxInc = 1 at trafoDepth 0 */
    if (cb_y) residual_coding(c, coef, 4, 0, 0);
    if (cb_u) residual_coding(c, coef + 256, 3, 1, 0);
    if (cb_v) residual_coding(c, coef + 320, 3, 2, 0);
}

static void write_cu(cabac *c, orc_frame_syntax *f, int cx, int cy)
{
    orc_cu *cu = &f->cu[cy * f->cuw + cx];
    const int16_t *coef = f->coef + (size_t)(cy * f->cuw + cx) * ORC_CU_COEFS;
    const int availL = cu_avail(f, cx, cy, cx - 1, cy), availA = cu_avail(f, cx, cy, cx, cy - 1);
    cu->skip = 0;
    if (!f->is_intra && cu->pred_mode == 0) {
        /* intra CU in a P slice: cu_skip_flag = 0, pred_mode_flag = 1; part_mode is not sent (the CU is larger than the minimum size) */
        const int ctx = (availL && cu_at(f, cx - 1, cy)->skip) + (availA && cu_at(f, cx, cy - 1)->skip);
        cb_bin(c, CX_SKIP + ctx, 0);
        cb_bin(c, CX_PRED_MODE, 1);
    } else if (!f->is_intra) {
        int16_t mc[5][2];
        int merge_idx = -1;
        orc_merge_candidates(f, cx, cy, mc);
        for (int k = 0; k < 5; k++)
            if (mc[k][0] == cu->mvx && mc[k][1] == cu->mvy) { merge_idx = k; break; }
        const int skip = merge_idx >= 0 && cu->cbf == 0;
        const int ctx = (availL && cu_at(f, cx - 1, cy)->skip) + (availA && cu_at(f, cx, cy - 1)->skip);
        cb_bin(c, CX_SKIP + ctx, skip);
        cu->skip = (uint8_t)skip;
        if (skip || merge_idx >= 0) {
            if (!skip) {
                cb_bin(c, CX_PRED_MODE, 0);
                cb_bin(c, CX_PART_MODE, 1);       /* PART_2Nx2N */
                cb_bin(c, CX_MERGE_FLAG, 1);
            }
            cb_bin(c, CX_MERGE_IDX, merge_idx > 0);
            if (merge_idx > 0)
                for (int k = 1; k < 4; k++) {
                    cb_bypass(c, merge_idx > k);
                    if (merge_idx <= k) break;
                }
            if (skip)
                return;
            /* merge + 2Nx2N: rqt_root_cbf inferred 1 */
        } else {
            int16_t ac[2][2];
            cb_bin(c, CX_PRED_MODE, 0);
            cb_bin(c, CX_PART_MODE, 1);
            cb_bin(c, CX_MERGE_FLAG, 0);
            orc_amvp_candidates(f, cx, cy, ac);
            const int b0 = mvd_bits(cu->mvx - ac[0][0]) + mvd_bits(cu->mvy - ac[0][1]);
            const int b1 = mvd_bits(cu->mvx - ac[1][0]) + mvd_bits(cu->mvy - ac[1][1]);
            const int idx = b1 < b0;
            write_mvd(c, cu->mvx - ac[idx][0], cu->mvy - ac[idx][1]);
            cb_bin(c, CX_MVP_FLAG, idx);
            cb_bin(c, CX_ROOT_CBF, cu->cbf != 0);
            if (!cu->cbf)
                return;
        }
        write_transform_unit(c, cu, coef, 0);
        return;
    }
    /* intra 2Nx2N: in an I slice no skip / pred_mode flags; no part_mode (CU larger than the minimum size) */
    int mpm[3], idx = -1;
    intra_mpm(f, cx, cy, mpm);
    for (int k = 0; k < 3; k++)
        if (mpm[k] == cu->intra_mode) idx = k;
    cb_bin(c, CX_PREV_INTRA, idx >= 0);
    if (idx >= 0) {
        cb_bypass(c, idx > 0);
        if (idx > 0) cb_bypass(c, idx > 1);
    } else {
        int m = cu->intra_mode;
        if (mpm[0] > mpm[1]) { int t = mpm[0]; mpm[0] = mpm[1]; mpm[1] = t; }
        if (mpm[0] > mpm[2]) { int t = mpm[0]; mpm[0] = mpm[2]; mpm[2] = t; }
        if (mpm[1] > mpm[2]) { int t = mpm[1]; mpm[1] = mpm[2]; mpm[2] = t; }
        for (int k = 2; k >= 0; k--)
            if (m > mpm[k]) m--;
        cb_bypass_bits(c, m, 5);
    }
    cb_bin(c, CX_CHROMA_PRED, 0);             /* intra_chroma_pred_mode = 4 (derived from luma) */
    write_transform_unit(c, cu, coef, 1);
}

static int sao_equal(const orc_sao *a, const orc_sao *b) { return memcmp(a, b, sizeof *a) == 0; }

/* sao() of one CTU (7.3.8.3).  Merging is by identity: the flag is set when the left (else the upper) CTU carries exactly the
 * same parameters, so it never changes the reconstruction, only the signalling. */
static void write_sao(cabac *c, const orc_frame_syntax *f, int rx, int ry)
{
    const orc_sao *s = &f->sao[ry * f->ctuw + rx];
    const int cmax = (1 << ((f->bit_depth < 10 ? f->bit_depth : 10) - 5)) - 1;
    if (rx > 0) {
        const int m = sao_equal(s, s - 1);
        cb_bin(c, CX_SAO_MERGE, m);
        if (m) return;
    }
    if (ry > 0) {
        const int m = sao_equal(s, s - f->ctuw);
        cb_bin(c, CX_SAO_MERGE, m);
        if (m) return;
    }
    for (int ci = 0; ci < 3; ci++) {
        const int g = ci ? 1 : 0, type = s->type[g];
        if (ci < 2) {                              /* sao_type_idx_luma / _chroma: TR cMax 2, first bin context coded */
            cb_bin(c, CX_SAO_TYPE, type != 0);
            if (type) cb_bypass(c, type == 2);
        }
        if (!type) continue;
        for (int i = 0; i < 4; i++) {              /* sao_offset_abs: TR, bypass */
            const int a = abs(s->offset[ci][i]);
            for (int k = 0; k < a; k++) cb_bypass(c, 1);
            if (a < cmax) cb_bypass(c, 0);
        }
        if (type == 1) {
            for (int i = 0; i < 4; i++)
                if (s->offset[ci][i]) cb_bypass(c, s->offset[ci][i] < 0);
            cb_bypass_bits(c, s->band[ci], 5);
        } else if (ci < 2) {
            cb_bypass_bits(c, s->eo_class[g], 2);
        }
    }
}

/* coding_quadtree at CTU level: split to four 16x16 CUs (those inside the picture) */
static void write_ctu(cabac *c, orc_frame_syntax *f, int ctx_x, int ctx_y)
{
    const int x0 = ctx_x * 32, y0 = ctx_y * 32;
    const int cx0 = ctx_x * 2, cy0 = ctx_y * 2;
    if (f->sao) write_sao(c, f, ctx_x, ctx_y);
    if (x0 + 32 <= f->wc && y0 + 32 <= f->hc) {
        /* depth-0 split flag: neighbours, when available, are at depth 1 > 0 */
        const int inc = cu_avail(f, cx0, cy0, cx0 - 1, cy0) + cu_avail(f, cx0, cy0, cx0, cy0 - 1);
        cb_bin(c, CX_SPLIT_CU + inc, 1);
    }
    for (int k = 0; k < 4; k++) {
        const int cx = cx0 + (k & 1), cy = cy0 + (k >> 1);
        if (cx >= f->cuw || cy >= f->cuh)
            continue;
        cb_bin(c, CX_SPLIT_CU + 0, 0);        /* depth 1: no neighbour is deeper */
        write_cu(c, f, cx, cy);
    }
}

int orc_cabac_encode_frame(orc_frame_syntax *f, uint8_t *out, size_t cap, uint32_t *row_off, uint32_t *row_len)
{
    cabac c;
    uint8_t saved[CX_COUNT];
    size_t pos = 0;
    const int init_type = f->is_intra ? 0 : 1;
    memset(saved, 0, sizeof saved);
    for (int r = 0; r < f->ctuh; r++) {
        cb_start(&c, out + pos, cap - pos);
        if (r == 0 || f->ctuw < 2)
            cb_init_contexts(&c, init_type, f->qp);
        else
            memcpy(c.ctx, saved, sizeof saved);          /* WPP: state after the 2nd CTU of the row above */
        for (int x = 0; x < f->ctuw; x++) {
            write_ctu(&c, f, x, r);
            if (x == 1 || (f->ctuw == 1 && x == 0))
                memcpy(saved, c.ctx, sizeof saved);
            const int last_in_pic = r == f->ctuh - 1 && x == f->ctuw - 1;
            cb_terminate(&c, last_in_pic);                /* end_of_slice_segment_flag */
            if (x == f->ctuw - 1 && !last_in_pic)
                cb_terminate(&c, 1);                      /* end_of_subset_one_bit */
        }
        const size_t n = cb_finish(&c);
        if (c.overflow)
            return -1;
        row_off[r] = (uint32_t)pos;
        row_len[r] = (uint32_t)n;
        pos += n;
    }
    return 0;
}
