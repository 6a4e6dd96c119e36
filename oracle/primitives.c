/* TEST INFRASTRUCTURE -- CPU oracle, not part of the product path.
 *
 * Scalar restatement of the encoder-side primitives the reference's CPU path runs inside libx265
 * (reached through `ffmpeg -c:v libx265`, reference core/transcoder.py:398-412,506).  libx265 is an
 * un-vendored, un-pinned dependency of the reference and is absent from this image, so these follow the
 * published definitions (x265 source/common/{pixel,dct,intrapred}.cpp C primitives; H.265 sections
 * 8.4.4.2 and 8.6) as listed in SURVEY.md section 8(c).
 *
 * PARITY STATUS: the inverse-side functions (idct/idst, dequant, intra prediction, interpolation) are
 * normative and are pinned through the FFmpeg hevc decoder (tests/test_oracle_encoder.py: decoder output
 * == model reconstruction).  The forward-side functions (SAD, SATD, SA8D, forward DCT/DST, quant) are
 * pinned only by these definitions and by algebraic properties: "parity unpinned" against x265 itself.
 *
 * pixel = uint16_t for every bit depth (x265 HIGH_BIT_DEPTH layout); coefficients int16_t.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint16_t pixel;

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : v > hi ? hi : v; }

/* ------------------------------------------------------------------ SAD / SATD / SA8D */

/* x265 pixel.cpp sad<lx,ly>: sum of absolute differences */
int orc_sad(const pixel *a, int sa, const pixel *b, int sb, int w, int h)
{
    int s = 0;
    for (int y = 0; y < h; y++, a += sa, b += sb)
        for (int x = 0; x < w; x++)
            s += abs((int)a[x] - (int)b[x]);
    return s;
}

/* 4x4 Hadamard of the difference block, sum of |.|, >> 1  (x265 satd_4x4) */
static int satd4x4_raw(const pixel *a, int sa, const pixel *b, int sb)
{
    int d[4][4], t[4][4], s = 0;
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++)
            d[y][x] = (int)a[y * sa + x] - (int)b[y * sb + x];
    for (int y = 0; y < 4; y++) {
        int s01 = d[y][0] + d[y][1], d01 = d[y][0] - d[y][1];
        int s23 = d[y][2] + d[y][3], d23 = d[y][2] - d[y][3];
        t[y][0] = s01 + s23; t[y][1] = d01 + d23; t[y][2] = s01 - s23; t[y][3] = d01 - d23;
    }
    for (int x = 0; x < 4; x++) {
        int s01 = t[0][x] + t[1][x], d01 = t[0][x] - t[1][x];
        int s23 = t[2][x] + t[3][x], d23 = t[2][x] - t[3][x];
        s += abs(s01 + s23) + abs(d01 + d23) + abs(s01 - s23) + abs(d01 - d23);
    }
    return s;
}

/* satd_8x4 in x265 is two horizontally adjacent 4x4 Hadamards summed before the >> 1 */
static int satd8x4(const pixel *a, int sa, const pixel *b, int sb)
{
    return (satd4x4_raw(a, sa, b, sb) + satd4x4_raw(a + 4, sa, b + 4, sb)) >> 1;
}

/* SATD of a WxH block: sum of 8x4 tiles, or 4x4 tiles when a dimension is 4 or 12 */
int orc_satd(const pixel *a, int sa, const pixel *b, int sb, int w, int h)
{
    int s = 0;
    if ((w % 8) == 0 && (h % 4) == 0 && w != 4 && w != 12) {
        for (int y = 0; y < h; y += 4)
            for (int x = 0; x < w; x += 8)
                s += satd8x4(a + y * sa + x, sa, b + y * sb + x, sb);
    } else {
        for (int y = 0; y < h; y += 4)
            for (int x = 0; x < w; x += 4)
                s += satd4x4_raw(a + y * sa + x, sa, b + y * sb + x, sb) >> 1;
    }
    return s;
}

/* raw 8x8 Hadamard sum (no normalisation) */
static int sa8d8x8_raw(const pixel *a, int sa, const pixel *b, int sb)
{
    int m[8][8], s = 0;
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            m[y][x] = (int)a[y * sa + x] - (int)b[y * sb + x];
    for (int y = 0; y < 8; y++)          /* rows */
        for (int len = 1; len < 8; len <<= 1)
            for (int i = 0; i < 8; i += len << 1)
                for (int j = i; j < i + len; j++) {
                    int u = m[y][j], v = m[y][j + len];
                    m[y][j] = u + v; m[y][j + len] = u - v;
                }
    for (int x = 0; x < 8; x++)          /* columns */
        for (int len = 1; len < 8; len <<= 1)
            for (int i = 0; i < 8; i += len << 1)
                for (int j = i; j < i + len; j++) {
                    int u = m[j][x], v = m[j + len][x];
                    m[j][x] = u + v; m[j + len][x] = u - v;
                }
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            s += abs(m[y][x]);
    return s;
}

/* sa8d: 8x8 = (raw+2)>>2; 16x16 = (sum of four raw 8x8 + 2)>>2; 32/64 = sum of 16x16 tiles; 4x4 = satd_4x4 */
int orc_sa8d(const pixel *a, int sa, const pixel *b, int sb, int n)
{
    if (n == 4)
        return satd4x4_raw(a, sa, b, sb) >> 1;
    if (n == 8)
        return (sa8d8x8_raw(a, sa, b, sb) + 2) >> 2;
    int s = 0;
    for (int y = 0; y < n; y += 16)
        for (int x = 0; x < n; x += 16) {
            int r = 0;
            for (int j = 0; j < 16; j += 8)
                for (int i = 0; i < 16; i += 8)
                    r += sa8d8x8_raw(a + (y + j) * sa + x + i, sa, b + (y + j) * sb + x + i, sb);
            s += (r + 2) >> 2;
        }
    return s;
}

/* ------------------------------------------------------------------ transforms (HM matrices) */

static const int8_t g_c4[3] = {64, 83, 36};
static const int8_t g_o8[4] = {89, 75, 50, 18};
static const int8_t g_o16[8] = {90, 87, 80, 70, 57, 43, 25, 9};
static const int8_t g_o32[16] = {90, 90, 88, 85, 82, 78, 73, 67, 61, 54, 46, 38, 31, 22, 13, 4};
static const int8_t g_dst4[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};

/* Row k, column n of the N-point HEVC core transform: c(k) * cos((2n+1) k pi / 2N) realised by the
 * nested even/odd structure of the HM tables. */
static int tmat(int N, int k, int n)
{
    if (k == 0)
        return 64;
    if (N == 2)                         /* k == 1 */
        return n == 0 ? 64 : -64;
    if ((k & 1) == 0) {                 /* even rows come from the N/2-point transform, mirrored */
        int m = n < N / 2 ? n : N - 1 - n;
        return tmat(N / 2, k / 2, m);
    }
    /* odd rows: value = +-odd[idx], from cos((2n+1)k pi/2N) folded into the first quadrant */
    const int8_t *odd = N == 4 ? (const int8_t[]){83, 36} : N == 8 ? g_o8 : N == 16 ? g_o16 : g_o32;
    int p = ((2 * n + 1) * k) % (4 * N);       /* angle in units of pi/2N, period 4N */
    int sign = 1;
    if (p > 2 * N) p = 4 * N - p;               /* cos(2pi - x) = cos x */
    if (p > N) { p = 2 * N - p; sign = -1; }    /* cos(pi - x) = -cos x */
    /* p is odd in [1, N-1]; odd[j] corresponds to angle (2j+1) */
    return sign * odd[(p - 1) / 2];
}

void orc_transform_matrix(int N, int is_dst, int16_t *out /* N*N */)
{
    (void)g_c4;
    for (int k = 0; k < N; k++)
        for (int n = 0; n < N; n++)
            out[k * N + n] = (int16_t)(is_dst ? g_dst4[k][n] : tmat(N, k, n));
}

static int ilog2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }

/* one separable pass: dst[k][line] = (sum_n M[k][n] * src[line][n] + add) >> shift   (output transposed) */
static void fwd_pass(const int16_t *M, const int32_t *src, int32_t *dst, int N, int shift)
{
    int add = shift > 0 ? 1 << (shift - 1) : 0;
    for (int line = 0; line < N; line++)
        for (int k = 0; k < N; k++) {
            int64_t acc = 0;
            for (int n = 0; n < N; n++)
                acc += (int64_t)M[k * N + n] * src[line * N + n];
            dst[k * N + line] = (int32_t)((acc + add) >> shift);
        }
}

/* forward DCT (or DST-VII for 4x4 intra luma): x265 dctN_c / dst4_c.  src: residual with stride. */
void orc_fwd_transform(const int16_t *src, int stride, int16_t *dst, int N, int bit_depth, int is_dst)
{
    int16_t M[32 * 32];
    int32_t a[32 * 32], b[32 * 32];
    int l2 = ilog2(N);
    orc_transform_matrix(N, is_dst, M);
    for (int y = 0; y < N; y++)
        for (int x = 0; x < N; x++)
            a[y * N + x] = src[y * stride + x];
    fwd_pass(M, a, b, N, l2 - 1 + (bit_depth - 8));
    fwd_pass(M, b, a, N, l2 + 6);
    for (int i = 0; i < N * N; i++)
        dst[i] = (int16_t)a[i];
}

/* inverse pass: dst[line][n] = clip16((sum_k M[k][n] * src[k][line] + add) >> shift)  (output transposed) */
static void inv_pass(const int16_t *M, const int32_t *src, int32_t *dst, int N, int shift)
{
    int add = 1 << (shift - 1);
    for (int line = 0; line < N; line++)
        for (int n = 0; n < N; n++) {
            int64_t acc = 0;
            for (int k = 0; k < N; k++)
                acc += (int64_t)M[k * N + n] * src[k * N + line];
            dst[line * N + n] = clip3(-32768, 32767, (int32_t)((acc + add) >> shift));
        }
}

/* inverse DCT/DST: H.265 8.6.4.2 (x265 idctN_c / idst4_c): shifts 7 and 12 - (bitDepth - 8) */
void orc_inv_transform(const int16_t *src, int16_t *dst, int stride, int N, int bit_depth, int is_dst)
{
    int16_t M[32 * 32];
    int32_t a[32 * 32], b[32 * 32];
    orc_transform_matrix(N, is_dst, M);
    for (int i = 0; i < N * N; i++)
        a[i] = src[i];
    inv_pass(M, a, b, N, 7);
    inv_pass(M, b, a, N, 12 - (bit_depth - 8));
    for (int y = 0; y < N; y++)
        for (int x = 0; x < N; x++)
            dst[y * stride + x] = (int16_t)a[y * N + x];
}

/* ------------------------------------------------------------------ quant / dequant (flat lists, RDOQ off) */

static const int g_quant_scale[6] = {26214, 23302, 20560, 18396, 16384, 14564};
static const int g_inv_quant_scale[6] = {40, 45, 51, 57, 64, 72};

/* x265 nquant_c with the flat scaling list; qp is the luma/chroma QP *including* QpBdOffset.
 * returns the number of non-zero levels */
int orc_quant(const int16_t *coef, int16_t *level, int N, int qp, int bit_depth, int is_intra)
{
    int l2 = ilog2(N), tshift = 15 - bit_depth - l2;
    int qbits = 14 + qp / 6 + tshift;
    int scale = g_quant_scale[qp % 6];
    int64_t add = (int64_t)(is_intra ? 171 : 85) << (qbits - 9);
    int nsig = 0;
    for (int i = 0; i < N * N; i++) {
        int c = coef[i], sign = c < 0 ? -1 : 1;
        int64_t t = (int64_t)abs(c) * scale;
        int lvl = (int)((t + add) >> qbits);
        lvl = clip3(-32768, 32767, sign * lvl);
        level[i] = (int16_t)lvl;
        nsig += lvl != 0;
    }
    return nsig;
}

/* x265 dequant_normal_c == H.265 8.6.3 with m = 16 */
void orc_dequant(const int16_t *level, int16_t *coef, int N, int qp, int bit_depth)
{
    int l2 = ilog2(N), tshift = 15 - bit_depth - l2;
    int shift = 20 - 14 - tshift;                 /* QUANT_IQUANT_SHIFT - QUANT_SHIFT - transformShift */
    int scale = g_inv_quant_scale[qp % 6] << (qp / 6);
    for (int i = 0; i < N * N; i++) {
        int64_t v = (int64_t)level[i] * scale;
        int r;
        if (shift > 0)
            r = (int)((v + (1 << (shift - 1))) >> shift);
        else
            r = (int)(v << -shift);
        coef[i] = (int16_t)clip3(-32768, 32767, r);
    }
}

/* ------------------------------------------------------------------ intra prediction (H.265 8.4.4.2) */

static const int8_t g_intra_angle[35] = {0, 0, 32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26,
                                         -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32};
static const int16_t g_inv_angle[35] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -4096, -1638, -910, -630, -482, -390, -315,
                                        -256, -315, -390, -482, -630, -910, -1638, -4096, 0, 0, 0, 0, 0, 0, 0, 0, 0};

/* neighbour buffer layout (x265): nb[0] = top-left, nb[1..2N] = top (left to right),
 * nb[2N+1..4N] = left (top to bottom).  All already substituted for availability. */

/* [1 2 1]/4 reference smoothing (8.4.4.2.3), optional bilinear "strong" variant for 32x32 */
void orc_intra_filter(const pixel *nb, pixel *out, int N, int strong, int bit_depth)
{
    int n2 = 2 * N;
    const pixel *top = nb + 1, *left = nb + 1 + n2;
    pixel *ftop = out + 1, *fleft = out + 1 + n2;
    if (strong && N == 32) {
        int thr = 1 << (bit_depth - 5);
        int tl = nb[0], tr = top[63], bl = left[63];
        if (abs(tl + tr - 2 * top[31]) < thr && abs(tl + bl - 2 * left[31]) < thr) {
            out[0] = nb[0];
            for (int i = 0; i < 63; i++) {
                fleft[i] = (pixel)(((63 - i) * tl + (i + 1) * bl + 32) >> 6);
                ftop[i] = (pixel)(((63 - i) * tl + (i + 1) * tr + 32) >> 6);
            }
            fleft[63] = left[63];
            ftop[63] = top[63];
            return;
        }
    }
    out[0] = (pixel)((left[0] + 2 * nb[0] + top[0] + 2) >> 2);
    for (int i = 0; i < n2 - 1; i++) {
        int tp = i == 0 ? nb[0] : top[i - 1];
        int lp = i == 0 ? nb[0] : left[i - 1];
        ftop[i] = (pixel)((tp + 2 * top[i] + top[i + 1] + 2) >> 2);
        fleft[i] = (pixel)((lp + 2 * left[i] + left[i + 1] + 2) >> 2);
    }
    ftop[n2 - 1] = top[n2 - 1];
    fleft[n2 - 1] = left[n2 - 1];
}

/* does mode use the smoothed reference?  luma only (8.4.4.2.3) */
int orc_intra_use_filter(int N, int mode)
{
    static const int thr[6] = {0, 0, 10 /*4x4: never*/, 7, 1, 0};
    if (mode == 1 || N == 4)
        return 0;
    if (mode == 0)
        return 1;
    int d1 = abs(mode - 26), d2 = abs(mode - 10);
    int md = d1 < d2 ? d1 : d2;
    return md > thr[ilog2(N)];
}

/* predict one block from an (already filtered or not) neighbour buffer.
 * edge: apply the DC / pure horizontal / pure vertical boundary smoothing (luma and N < 32) */
void orc_intra_pred(const pixel *nb, pixel *dst, int stride, int N, int mode, int edge, int bit_depth)
{
    int n2 = 2 * N, maxv = (1 << bit_depth) - 1;
    const pixel *top = nb + 1, *left = nb + 1 + n2;
    if (mode == 0) {                                  /* planar 8.4.4.2.4 */
        int l2 = ilog2(N);
        for (int y = 0; y < N; y++)
            for (int x = 0; x < N; x++)
                dst[y * stride + x] = (pixel)(((N - 1 - x) * left[y] + (x + 1) * top[N] + (N - 1 - y) * top[x] +
                                               (y + 1) * left[N] + N) >> (l2 + 1));
        return;
    }
    if (mode == 1) {                                  /* DC 8.4.4.2.5 */
        int s = N, l2 = ilog2(N);
        for (int i = 0; i < N; i++)
            s += top[i] + left[i];
        int dc = s >> (l2 + 1);
        for (int y = 0; y < N; y++)
            for (int x = 0; x < N; x++)
                dst[y * stride + x] = (pixel)dc;
        if (edge && N < 32) {
            dst[0] = (pixel)((left[0] + 2 * dc + top[0] + 2) >> 2);
            for (int x = 1; x < N; x++)
                dst[x] = (pixel)((top[x] + 3 * dc + 2) >> 2);
            for (int y = 1; y < N; y++)
                dst[y * stride] = (pixel)((left[y] + 3 * dc + 2) >> 2);
        }
        return;
    }
    /* angular 8.4.4.2.6 */
    int ang = g_intra_angle[mode], vertical = mode >= 18;
    int refbuf[3 * 32 + 1];
    int *ref = refbuf + 32;                           /* ref[-N .. 2N] */
    const pixel *main_ = vertical ? top : left, *side = vertical ? left : top;
    ref[0] = nb[0];
    for (int i = 1; i <= n2; i++)
        ref[i] = main_[i - 1];
    if (ang < 0) {
        int last = (N * ang) >> 5;
        if (last < -1) {
            int inv = g_inv_angle[mode];
            for (int i = last; i <= -1; i++) {
                int idx = -1 + ((i * inv + 128) >> 8);   /* side index: -1 = corner */
                ref[i] = idx < 0 ? nb[0] : side[idx];
            }
        }
    }
    for (int j = 0; j < N; j++) {                     /* j: along the prediction direction */
        int pos = (j + 1) * ang, idx = pos >> 5, fr = pos & 31;
        for (int i = 0; i < N; i++) {
            int v = fr ? ((32 - fr) * ref[i + idx + 1] + fr * ref[i + idx + 2] + 16) >> 5 : ref[i + idx + 1];
            if (vertical)
                dst[j * stride + i] = (pixel)v;
            else
                dst[i * stride + j] = (pixel)v;
        }
    }
    if (edge && N < 32 && ang == 0) {
        if (vertical)
            for (int y = 0; y < N; y++)
                dst[y * stride] = (pixel)clip3(0, maxv, top[0] + (((int)left[y] - (int)nb[0]) >> 1));
        else
            for (int x = 0; x < N; x++)
                dst[x] = (pixel)clip3(0, maxv, left[0] + (((int)top[x] - (int)nb[0]) >> 1));
    }
}

/* all 35 luma or chroma predictions of one block, applying the normative smoothing decision per mode.
 * out: [35][N*N].  is_luma selects smoothing + edge filters (chroma 4:2:0 has neither). */
void orc_intra_pred_all(const pixel *nb, pixel *out, int N, int is_luma, int strong, int bit_depth)
{
    pixel filt[4 * 32 + 1];
    orc_intra_filter(nb, filt, N, strong, bit_depth);
    for (int mode = 0; mode < 35; mode++) {
        int f = is_luma && orc_intra_use_filter(N, mode);
        orc_intra_pred(f ? filt : nb, out + mode * N * N, N, N, mode, is_luma, bit_depth);
    }
}

/* ------------------------------------------------------------------ sub-pel interpolation (H.265 8.5.3.3.3) */

static const int8_t g_luma_taps[4][8] = {{0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0},
                                         {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1}};
static const int8_t g_chroma_taps[8][4] = {{0, 64, 0, 0}, {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4},
                                           {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2}};

/* uni-directional luma prediction sample block: ref points at the integer sample (x + mvx>>2, y + mvy>>2);
 * the caller guarantees 3 samples left/top and 4 right/bottom are readable */
void orc_interp_luma(const pixel *ref, int rs, pixel *dst, int ds, int w, int h, int fx, int fy, int bit_depth)
{
    int shift1 = bit_depth - 8, maxv = (1 << bit_depth) - 1;
    int s14 = 14 - bit_depth, off14 = 1 << (s14 - 1);
    if (!fx && !fy) {
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                dst[y * ds + x] = ref[y * rs + x];  /* (v << s14 + off) >> s14 == v */
        return;
    }
    if (!fy) {
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                int a = 0;
                for (int t = 0; t < 8; t++)
                    a += g_luma_taps[fx][t] * ref[y * rs + x + t - 3];
                dst[y * ds + x] = (pixel)clip3(0, maxv, ((a >> shift1) + off14) >> s14);
            }
        return;
    }
    if (!fx) {
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                int a = 0;
                for (int t = 0; t < 8; t++)
                    a += g_luma_taps[fy][t] * ref[(y + t - 3) * rs + x];
                dst[y * ds + x] = (pixel)clip3(0, maxv, ((a >> shift1) + off14) >> s14);
            }
        return;
    }
    int16_t *tmp = (int16_t *)malloc(sizeof(int16_t) * (size_t)w * (h + 7));
    for (int y = -3; y < h + 4; y++)
        for (int x = 0; x < w; x++) {
            int a = 0;
            for (int t = 0; t < 8; t++)
                a += g_luma_taps[fx][t] * ref[y * rs + x + t - 3];
            tmp[(y + 3) * w + x] = (int16_t)(a >> shift1);
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int a = 0;
            for (int t = 0; t < 8; t++)
                a += g_luma_taps[fy][t] * tmp[(y + t) * w + x];
            dst[y * ds + x] = (pixel)clip3(0, maxv, ((a >> 6) + off14) >> s14);
        }
    free(tmp);
}

/* chroma, fractions in 1/8 sample; 1 sample left/top and 2 right/bottom readable */
void orc_interp_chroma(const pixel *ref, int rs, pixel *dst, int ds, int w, int h, int fx, int fy, int bit_depth)
{
    int shift1 = bit_depth - 8, maxv = (1 << bit_depth) - 1;
    int s14 = 14 - bit_depth, off14 = 1 << (s14 - 1);
    if (!fx && !fy) {
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                dst[y * ds + x] = ref[y * rs + x];
        return;
    }
    if (!fy || !fx) {
        int step = fy ? rs : 1, f = fy ? fy : fx;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                int a = 0;
                for (int t = 0; t < 4; t++)
                    a += g_chroma_taps[f][t] * ref[y * rs + x + (t - 1) * step];
                dst[y * ds + x] = (pixel)clip3(0, maxv, ((a >> shift1) + off14) >> s14);
            }
        return;
    }
    int16_t *tmp = (int16_t *)malloc(sizeof(int16_t) * (size_t)w * (h + 3));
    for (int y = -1; y < h + 2; y++)
        for (int x = 0; x < w; x++) {
            int a = 0;
            for (int t = 0; t < 4; t++)
                a += g_chroma_taps[fx][t] * ref[y * rs + x + t - 1];
            tmp[(y + 1) * w + x] = (int16_t)(a >> shift1);
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int a = 0;
            for (int t = 0; t < 4; t++)
                a += g_chroma_taps[fy][t] * tmp[(y + t) * w + x];
            dst[y * ds + x] = (pixel)clip3(0, maxv, ((a >> 6) + off14) >> s14);
        }
    free(tmp);
}
