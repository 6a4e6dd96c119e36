// Device-side integer primitives shared by the batched primitive kernels (prim.cu) and the encoder kernels.
// One thread runs a whole 1-D N-point partial butterfly in registers; the HM matrix entries are folded into
// IMAD immediates at compile time (constexpr coefficient function + full unrolling), so the transform makes
// no table loads at all.  Semantics: H.265 8.6.4.2 / x265 dct.cpp C primitives.
#pragma once
#include <stdint.h>

namespace hb {

typedef uint16_t pixel;

#define HB_DEV __device__ __forceinline__
#define HB_HD __host__ __device__ __forceinline__

HB_HD constexpr int clip3i(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
HB_HD constexpr int ilog2c(int n) { return n <= 1 ? 0 : 1 + ilog2c(n >> 1); }

// odd-row magnitude j of the N-point transform (angle (2j+1) * pi / 2N)
HB_HD constexpr int odd_coef(int N, int j)
{
    constexpr int o4[2] = {83, 36};
    constexpr int o8[4] = {89, 75, 50, 18};
    constexpr int o16[8] = {90, 87, 80, 70, 57, 43, 25, 9};
    constexpr int o32[16] = {90, 90, 88, 85, 82, 78, 73, 67, 61, 54, 46, 38, 31, 22, 13, 4};
    return N == 4 ? o4[j] : N == 8 ? o8[j] : N == 16 ? o16[j] : o32[j];
}

// entry (odd row k, column n < N/2) of the N-point matrix: sign * odd_coef, by folding the cosine angle
HB_HD constexpr int odd_entry(int N, int k, int n)
{
    int p = ((2 * n + 1) * k) % (4 * N);
    int sign = 1;
    if (p > 2 * N) p = 4 * N - p;
    if (p > N) { p = 2 * N - p; sign = -1; }
    return sign * odd_coef(N, (p - 1) / 2);
}

constexpr int kDst4[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};
HB_HD constexpr int dst_entry(int k, int n)
{
    constexpr int t[16] = {29, 55, 74, 84, 74, 74, 0, -74, 84, -29, -74, 55, 55, -84, 74, -29};
    return t[k * 4 + n];
}

// ---- forward: out[k] = sum_n M[k][n] in[n]   (raw sums, caller applies rounding shift)
template <int N>
struct Fwd1D {
    static HB_DEV void run(const int (&in)[N], int (&out)[N])
    {
        int e[N / 2], o[N / 2], ee[N / 2];
#pragma unroll
        for (int k = 0; k < N / 2; k++) {
            e[k] = in[k] + in[N - 1 - k];
            o[k] = in[k] - in[N - 1 - k];
        }
        Fwd1D<N / 2>::run(e, ee);
#pragma unroll
        for (int j = 0; j < N / 2; j++) {
            out[2 * j] = ee[j];
            int acc = 0;
#pragma unroll
            for (int k = 0; k < N / 2; k++)
                acc += odd_entry(N, 2 * j + 1, k) * o[k];
            out[2 * j + 1] = acc;
        }
    }
};
template <>
struct Fwd1D<2> {
    static HB_DEV void run(const int (&in)[2], int (&out)[2])
    {
        out[0] = 64 * (in[0] + in[1]);
        out[1] = 64 * (in[0] - in[1]);
    }
};

// ---- inverse: out[n] = sum_k M[k][n] in[k]
template <int N>
struct Inv1D {
    static HB_DEV void run(const int (&in)[N], int (&out)[N])
    {
        int ev[N / 2], e[N / 2];
#pragma unroll
        for (int k = 0; k < N / 2; k++)
            ev[k] = in[2 * k];
        Inv1D<N / 2>::run(ev, e);
#pragma unroll
        for (int n = 0; n < N / 2; n++) {
            int o = 0;
#pragma unroll
            for (int j = 0; j < N / 2; j++)
                o += odd_entry(N, 2 * j + 1, n) * in[2 * j + 1];
            out[n] = e[n] + o;
            out[N - 1 - n] = e[n] - o;
        }
    }
};
template <>
struct Inv1D<2> {
    static HB_DEV void run(const int (&in)[2], int (&out)[2])
    {
        out[0] = 64 * (in[0] + in[1]);
        out[1] = 64 * (in[0] - in[1]);
    }
};

HB_DEV void dst4_fwd(const int (&in)[4], int (&out)[4])
{
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int acc = 0;
#pragma unroll
        for (int n = 0; n < 4; n++)
            acc += dst_entry(k, n) * in[n];
        out[k] = acc;
    }
}
HB_DEV void dst4_inv(const int (&in)[4], int (&out)[4])
{
#pragma unroll
    for (int n = 0; n < 4; n++) {
        int acc = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
            acc += dst_entry(k, n) * in[k];
        out[n] = acc;
    }
}

// One separable pass for one line held by this thread.
//  forward: reads N values at src[i*sstep], writes (sum + add) >> shift to dst[k*dstep]
//  inverse: same with int16 clipping
template <int N, bool DST>
HB_DEV void fwd_line(const int16_t *src, int sstep, int16_t *dst, int dstep, int shift)
{
    int in[N], out[N];
#pragma unroll
    for (int i = 0; i < N; i++)
        in[i] = src[i * sstep];
    if constexpr (DST)
        dst4_fwd(in, out);
    else
        Fwd1D<N>::run(in, out);
    const int add = shift > 0 ? 1 << (shift - 1) : 0;
#pragma unroll
    for (int k = 0; k < N; k++)
        dst[k * dstep] = (int16_t)((out[k] + add) >> shift);
}

template <int N, bool DST>
HB_DEV void inv_line(const int16_t *src, int sstep, int16_t *dst, int dstep, int shift)
{
    int in[N], out[N];
#pragma unroll
    for (int i = 0; i < N; i++)
        in[i] = src[i * sstep];
    if constexpr (DST)
        dst4_inv(in, out);
    else
        Inv1D<N>::run(in, out);
    const int add = 1 << (shift - 1);
#pragma unroll
    for (int n = 0; n < N; n++)
        dst[n * dstep] = (int16_t)clip3i(-32768, 32767, (out[n] + add) >> shift);
}

// ---- quantisation (flat scaling list, no RDOQ) -- x265 nquant / dequant_normal
HB_HD constexpr int quant_scale(int rem)
{
    constexpr int t[6] = {26214, 23302, 20560, 18396, 16384, 14564};
    return t[rem];
}
HB_HD constexpr int dequant_scale(int rem)
{
    constexpr int t[6] = {40, 45, 51, 57, 64, 72};
    return t[rem];
}

struct QuantParam {
    int scale, qbits;
    long long add;
    int dq_scale, dq_shift;   // dq_shift may be <= 0 (left shift)
};

HB_HD QuantParam make_quant(int log2n, int qp, int bit_depth, int is_intra)
{
    QuantParam q;
    const int tshift = 15 - bit_depth - log2n;
    q.qbits = 14 + qp / 6 + tshift;
    q.scale = quant_scale(qp % 6);
    q.add = (long long)(is_intra ? 171 : 85) << (q.qbits - 9);
    q.dq_scale = dequant_scale(qp % 6) << (qp / 6);
    q.dq_shift = 20 - 14 - tshift;
    return q;
}

HB_DEV int quant_one(int c, const QuantParam &q)
{
    const int a = c < 0 ? -c : c;
    int lvl = (int)(((long long)a * q.scale + q.add) >> q.qbits);
    lvl = c < 0 ? -lvl : lvl;
    return clip3i(-32768, 32767, lvl);
}

HB_DEV int dequant_one(int lvl, const QuantParam &q)
{
    const long long v = (long long)lvl * q.dq_scale;
    int r;
    if (q.dq_shift > 0)
        r = (int)((v + (1 << (q.dq_shift - 1))) >> q.dq_shift);
    else
        r = (int)(v << -q.dq_shift);
    return clip3i(-32768, 32767, r);
}

// ---- Hadamard cost kernels
// |H4 D H4| summed, un-normalised, of a 4x4 difference block held as 4 rows of 4 ints
HB_DEV int hadamard4x4_abs(int d[4][4])
{
    int s = 0;
#pragma unroll
    for (int y = 0; y < 4; y++) {
        const int s01 = d[y][0] + d[y][1], d01 = d[y][0] - d[y][1];
        const int s23 = d[y][2] + d[y][3], d23 = d[y][2] - d[y][3];
        d[y][0] = s01 + s23; d[y][1] = d01 + d23; d[y][2] = s01 - s23; d[y][3] = d01 - d23;
    }
#pragma unroll
    for (int x = 0; x < 4; x++) {
        const int s01 = d[0][x] + d[1][x], d01 = d[0][x] - d[1][x];
        const int s23 = d[2][x] + d[3][x], d23 = d[2][x] - d[3][x];
        s += abs(s01 + s23) + abs(d01 + d23) + abs(s01 - s23) + abs(d01 - d23);
    }
    return s;
}

HB_DEV int hadamard8x8_abs(int m[8][8])
{
#pragma unroll
    for (int y = 0; y < 8; y++)
#pragma unroll
        for (int len = 1; len < 8; len <<= 1)
#pragma unroll
            for (int i = 0; i < 8; i += len << 1)
#pragma unroll
                for (int j = i; j < i + len; j++) {
                    const int u = m[y][j], v = m[y][j + len];
                    m[y][j] = u + v; m[y][j + len] = u - v;
                }
    int s = 0;
#pragma unroll
    for (int x = 0; x < 8; x++) {
#pragma unroll
        for (int len = 1; len < 8; len <<= 1)
#pragma unroll
            for (int i = 0; i < 8; i += len << 1)
#pragma unroll
                for (int j = i; j < i + len; j++) {
                    const int u = m[j][x], v = m[j + len][x];
                    m[j][x] = u + v; m[j + len][x] = u - v;
                }
#pragma unroll
        for (int y = 0; y < 8; y++)
            s += abs(m[y][x]);
    }
    return s;
}

// ---- intra prediction (H.265 8.4.4.2).  Neighbour buffer: nb[0] corner, nb[1..2N] top, nb[2N+1..4N] left.
// mode -> intraPredAngle / invAngle (H.265 Tables 8-4, 8-5); constant memory: indexed by a run-time mode
static __constant__ int8_t c_intra_angle[35] = {0, 0, 32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26,
                                                -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32};
static __constant__ int16_t c_intra_inv_angle[15] = {-4096, -1638, -910, -630, -482, -390, -315, -256, -315, -390, -482, -630, -910, -1638, -4096};
HB_DEV int intra_angle(int mode) { return c_intra_angle[mode]; }
HB_DEV int intra_inv_angle(int mode) { return (mode >= 11 && mode <= 25) ? c_intra_inv_angle[mode - 11] : 0; }

HB_HD bool intra_use_filter(int log2n, int mode)
{
    if (mode == 1 || log2n == 2)
        return false;
    if (mode == 0)
        return true;
    const int d1 = mode > 26 ? mode - 26 : 26 - mode, d2 = mode > 10 ? mode - 10 : 10 - mode;
    const int md = d1 < d2 ? d1 : d2;
    const int thr = log2n == 3 ? 7 : log2n == 4 ? 1 : 0;
    return md > thr;
}

// smoothed neighbour sample i of the [1 2 1]/4 filter (strong bilinear variant decided by the caller)
template <typename T>
HB_DEV int intra_filtered(const T *nb, int n, int i)
{
    const int n2 = 2 * n;
    if (i == 0)
        return (nb[1 + n2] + 2 * nb[0] + nb[1] + 2) >> 2;
    if (i == n2 || i == 2 * n2)
        return nb[i];
    if (i <= n2) {                        // top[i-1]
        const int prev = i == 1 ? nb[0] : nb[i - 1];
        return (prev + 2 * nb[i] + nb[i + 1] + 2) >> 2;
    }
    const int prev = i == n2 + 1 ? nb[0] : nb[i - 1];
    return (prev + 2 * nb[i] + nb[i + 1] + 2) >> 2;
}

// one predicted sample at (x, y) of an NxN block for `mode`, from neighbour buffer nb (already the
// filtered or unfiltered version as the mode requires).  edge: DC / pure H / pure V boundary smoothing.
template <typename T>
HB_DEV int intra_sample(const T *nb, int n, int log2n, int mode, int x, int y, bool edge, int maxv, int dc)
{
    const int n2 = 2 * n;
    const T *top = nb + 1, *left = nb + 1 + n2;
    if (mode == 0)
        return ((n - 1 - x) * left[y] + (x + 1) * top[n] + (n - 1 - y) * top[x] + (y + 1) * left[n] + n) >> (log2n + 1);
    if (mode == 1) {
        if (edge && n < 32) {
            if (x == 0 && y == 0) return (left[0] + 2 * dc + top[0] + 2) >> 2;
            if (y == 0) return (top[x] + 3 * dc + 2) >> 2;
            if (x == 0) return (left[y] + 3 * dc + 2) >> 2;
        }
        return dc;
    }
    const int ang = intra_angle(mode);
    const bool vertical = mode >= 18;
    const int i = vertical ? x : y, j = vertical ? y : x;      // i across the reference, j along the direction
    const T *mainr = vertical ? top : left, *side = vertical ? left : top;
    if (ang == 0) {
        int v = mainr[i];
        if (edge && n < 32 && i == 0)
            v = clip3i(0, maxv, v + (((int)side[j] - (int)nb[0]) >> 1));
        return v;
    }
    const int pos = (j + 1) * ang, idx = pos >> 5, fr = pos & 31;
    const int inv = intra_inv_angle(mode);
    auto ref = [&](int r) -> int {     // r in [-n, 2n]
        if (r > 0) return mainr[r - 1];
        if (r == 0) return nb[0];
        const int s = -1 + ((r * inv + 128) >> 8);
        return s < 0 ? nb[0] : side[s];
    };
    const int a = ref(i + idx + 1);
    if (fr == 0)
        return a;
    const int b = ref(i + idx + 2);
    return ((32 - fr) * a + fr * b + 16) >> 5;
}

// ---- sub-pel interpolation taps (H.265 8.5.3.3.3)
HB_HD constexpr int luma_tap(int frac, int t)
{
    constexpr int k[32] = {0, 0, 0, 64, 0, 0, 0, 0, -1, 4, -10, 58, 17, -5, 1, 0, -1, 4, -11, 40, 40, -11, 4, -1, 0, 1, -5, 17, 58, -10, 4, -1};
    return k[frac * 8 + t];
}
HB_HD constexpr int chroma_tap(int frac, int t)
{
    constexpr int k[32] = {0, 64, 0, 0, -2, 58, 10, -2, -4, 54, 16, -2, -6, 46, 28, -4, -4, 36, 36, -4, -4, 28, 46, -6, -2, 16, 54, -4, -2, 10, 58, -2};
    return k[frac * 4 + t];
}

}  // namespace hb
