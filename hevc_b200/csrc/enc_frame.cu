// Frame-level encoder kernels for sm_100a: ingest (+ quarter-resolution plane), hierarchical motion search,
// fused inter prediction / transform / quantisation / reconstruction (one warp per 16x16 CU), wavefront intra
// frames (one CTA per CTU row), border extension.  Bit-exact against oracle/hevc_encode.c.
#include <cuda.h>

#include "enc_kernels.cuh"

namespace hb {

__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }

// interpolation taps in constant memory: the fraction is warp-uniform, so these are broadcast LDCs
__constant__ int8_t c_luma_taps[4][8] = {{0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1}};
// the same luma taps packed for dp2a: {E0123, E4567, O(0,t0,t1,t2), O(t3,t4,t5,t6), O(t7,0,0,0)} (bytes, little end first)
__constant__ int c_luma_pack[4][5];
__constant__ int8_t c_chroma_taps[8][4] = {{0, 64, 0, 0}, {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4},
                                           {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2}};
__device__ __forceinline__ int clampd(int v, int lo, int hi) { return min(max(v, lo), hi); }
// clamp to [0, hi] in one instruction (VIMNMX.RELU)
__device__ __forceinline__ int clamp0(int v, int hi)
{
    int r;
    asm("min.relu.s32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(hi));
    return r;
}
// sum of absolute differences of four packed bytes, accumulated: one VABSDIFF4.U8.ACC
__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t acc)
{
    uint32_t r;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(acc));
    return r;
}
// 8 most significant bits of four 16-bit samples held as two pair words -> four packed bytes
__device__ __forceinline__ uint32_t pack4_msb8(uint32_t lo, uint32_t hi, int sh)
{
    const uint32_t a = (lo >> sh) & 0x00ff00ffu, b = (hi >> sh) & 0x00ff00ffu;     // bytes 0,2 of each
    return __byte_perm(a, b, 0x6420);
}

// ================================================================================================ ingest
// one thread per 4x4 luma block: copies (with edge replication into the coded-size padding), converts the
// input sample format and produces the quarter-resolution sample (sum + 8) >> 4
__global__ void __launch_bounds__(256) k_ingest(IngestParams p)
{
    const Geom &g = p.g;
    const int total = g.dsw * g.dsh;
    const int cw = p.w >> 1, ch = p.h >> 1;
    for (int blk = blockIdx.x * blockDim.x + threadIdx.x; blk < total; blk += gridDim.x * blockDim.x) {
        const int bx = blk % g.dsw, by = blk / g.dsw;
        int acc = 8;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int yy = min(4 * by + j, p.h - 1);
            const uint8_t *row = p.in_y + (size_t)yy * p.in_ys;
            uint32_t px[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int xx = min(4 * bx + i, p.w - 1);
                uint32_t v;
                if (p.fmt == HB_FMT_YUV420P8) v = (uint32_t)row[xx] << p.up_shift;
                else v = min((((uint32_t)reinterpret_cast<const uint16_t *>(row)[xx] + p.round_add) >> p.down_shift) << p.up_shift, (uint32_t)p.maxv);
                px[i] = v;
                acc += v;
            }
            uint2 o = make_uint2(px[0] | (px[1] << 16), px[2] | (px[3] << 16));
            *reinterpret_cast<uint2 *>(p.src.y + (size_t)(4 * by + j) * g.src_stride + 4 * bx) = o;
        }
        p.ds[(size_t)by * g.dsw + bx] = (uint8_t)((acc >> 4) >> (g.bit_depth - 8));     // coarse search works on the 8 MSBs
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int yy = min(2 * by + j, ch - 1);
            uint32_t uu[2], vv[2];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const int xx = min(2 * bx + i, cw - 1);
                if (p.fmt == HB_FMT_YUV420P8) {
                    uu[i] = (uint32_t)p.in_u[(size_t)yy * p.in_us + xx] << p.up_shift;
                    vv[i] = (uint32_t)p.in_v[(size_t)yy * p.in_vs + xx] << p.up_shift;
                } else if (p.fmt == HB_FMT_P010) {      // interleaved UV plane in in_u
                    const uint16_t *r = reinterpret_cast<const uint16_t *>(p.in_u + (size_t)yy * p.in_us);
                    uu[i] = (uint32_t)r[2 * xx] >> p.down_shift;
                    vv[i] = (uint32_t)r[2 * xx + 1] >> p.down_shift;
                } else {
                    uu[i] = min((((uint32_t)reinterpret_cast<const uint16_t *>(p.in_u + (size_t)yy * p.in_us)[xx] + p.round_add) >> p.down_shift) << p.up_shift, (uint32_t)p.maxv);
                    vv[i] = min((((uint32_t)reinterpret_cast<const uint16_t *>(p.in_v + (size_t)yy * p.in_vs)[xx] + p.round_add) >> p.down_shift) << p.up_shift, (uint32_t)p.maxv);
                }
            }
            *reinterpret_cast<uint32_t *>(p.src.u + (size_t)(2 * by + j) * g.srcc_stride + 2 * bx) = uu[0] | (uu[1] << 16);
            *reinterpret_cast<uint32_t *>(p.src.v + (size_t)(2 * by + j) * g.srcc_stride + 2 * bx) = vv[0] | (vv[1] << 16);
        }
    }
}

// After a fused scale / colour-conversion ingest (pixel.cu) wrote the display area of the source planes: replicate the last
// row / column into the coded-size padding and build the quarter-resolution plane.  One thread per 4x4 luma block; reads
// touch only the display area, writes only the padding, so there is no ordering between threads.
__global__ void __launch_bounds__(256) k_pad_ds(Geom g, Planes src, int w, int h, uint8_t *ds)
{
    const int total = g.dsw * g.dsh, cw = w >> 1, ch = h >> 1;
    for (int blk = blockIdx.x * blockDim.x + threadIdx.x; blk < total; blk += gridDim.x * blockDim.x) {
        const int bx = blk % g.dsw, by = blk / g.dsw;
        int acc = 8;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int y = 4 * by + j, yy = min(y, h - 1);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int x = 4 * bx + i, xx = min(x, w - 1);
                const pixel v = src.y[(size_t)yy * g.src_stride + xx];
                acc += v;
                if (y != yy || x != xx) src.y[(size_t)y * g.src_stride + x] = v;
            }
        }
        ds[(size_t)by * g.dsw + bx] = (uint8_t)((acc >> 4) >> (g.bit_depth - 8));
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int y = 2 * by + j, yy = min(y, ch - 1);
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const int x = 2 * bx + i, xx = min(x, cw - 1);
                if (y != yy || x != xx) {
                    src.u[(size_t)y * g.srcc_stride + x] = src.u[(size_t)yy * g.srcc_stride + xx];
                    src.v[(size_t)y * g.srcc_stride + x] = src.v[(size_t)yy * g.srcc_stride + xx];
                }
            }
        }
    }
}

// ================================================================================================ border extension
// blockIdx.y = plane; threads enumerate only the border samples of the padded plane
__global__ void __launch_bounds__(256) k_border(Planes rec, Geom g)
{
    const int pl = blockIdx.y;
    pixel *base = pl == 0 ? rec.y : pl == 1 ? rec.u : rec.v;
    const int w = pl ? g.wc >> 1 : g.wc, h = pl ? g.hc >> 1 : g.hc, pad = pl ? kPad >> 1 : kPad;
    const int stride = pl ? g.recc_stride : g.rec_stride;
    const int fullw = w + 2 * pad;
    const int n_tb = 2 * pad * fullw, n_side = h * 2 * pad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_tb + n_side; i += gridDim.x * blockDim.x) {
        int x, y;
        if (i < n_tb) {
            const int r = i / fullw;
            x = i % fullw - pad;
            y = r < pad ? r - pad : h + (r - pad);
        } else {
            const int j = i - n_tb, r = j / (2 * pad), c = j % (2 * pad);
            y = r;
            x = c < pad ? c - pad : w + (c - pad);
        }
        base[(ptrdiff_t)y * stride + x] = base[(ptrdiff_t)clampd(y, 0, h - 1) * stride + clampd(x, 0, w - 1)];
    }
}

// ================================================================================================ coarse search
// one CTA per CTU (8x8 block of the 8-bit quarter-resolution plane), blockIdx.y = frame of the batch.
// 625 candidates (+-12), byte SAD (VABSDIFF4: four samples + accumulate per instruction) + |dx| + |dy|; winner = min over
// (cost << 10 | raster index).  The 32x32 search window is kept in four byte-shifted copies so that every candidate reads
// aligned words; the current block sits in 16 registers.
constexpr int kCoarseCopyWords = 32 * 9 + 8;       // row stride 36 bytes; +8 words so that the copies start 8 banks apart
__global__ void __launch_bounds__(128) k_coarse(CoarseParams p)
{
    __shared__ uint32_t curw[16];
    __shared__ uint32_t win[4][kCoarseCopyWords];
    __shared__ unsigned long long best[4];
    __shared__ unsigned grad[2][2];
    const Geom &g = p.g;
    const int tx = blockIdx.x % g.ctuw, ty = blockIdx.x / g.ctuw, f = blockIdx.y;
    const uint8_t *dcur = p.ds + (size_t)(f + 1) * p.ds_frame_stride, *dprev = p.ds + (size_t)f * p.ds_frame_stride;
    const int tid = threadIdx.x;
    if (tid < 64) {
        const int i = tid & 7, j = tid >> 3;
        reinterpret_cast<uint8_t *>(curw)[tid] = dcur[(size_t)clampd(ty * 8 + j, 0, g.dsh - 1) * g.dsw + clampd(tx * 8 + i, 0, g.dsw - 1)];
    }
    // window: interior CTUs read aligned words and derive the three byte-shifted copies with funnel shifts (each thread owns
    // two row-words plus their right neighbours); picture-edge CTUs clamp sample by sample
    const int wx = tx * 8 - kCmeRange, wy = ty * 8 - kCmeRange;
    if (wx >= 0 && wy >= 0 && wx + 32 <= g.dsw && wy + 32 <= g.dsh && (g.dsw & 3) == 0) {
        for (int k = tid; k < 32 * 8; k += 128) {
            const int j = k >> 3, wi = k & 7;
            const uint32_t *row = reinterpret_cast<const uint32_t *>(dprev + (size_t)(wy + j) * g.dsw + wx);
            const uint32_t w0 = row[wi], w1 = wi < 7 ? row[wi + 1] : 0;
            win[0][j * 9 + wi] = w0;
#pragma unroll
            for (int sft = 1; sft < 4; sft++) win[sft][j * 9 + wi] = __funnelshift_r(w0, w1, 8 * sft);
        }
    } else {
        for (int k = tid; k < 32 * 32; k += 128) {
            const int i = k & 31, j = k >> 5;
            const uint8_t v = dprev[(size_t)clampd(wy + j, 0, g.dsh - 1) * g.dsw + clampd(wx + i, 0, g.dsw - 1)];
#pragma unroll
            for (int sft = 0; sft < 4; sft++)
                if (i >= sft) reinterpret_cast<uint8_t *>(win[sft])[j * 36 + i - sft] = v;
        }
    }
    __syncthreads();
    if (tid < 64) {         // intra measure of the scene-cut detector: horizontal / vertical neighbour differences inside the block
        const uint8_t *cb = reinterpret_cast<const uint8_t *>(curw);
        const int i = tid & 7, j = tid >> 3, v = cb[tid];
        const unsigned hs = __reduce_add_sync(0xffffffffu, i ? (unsigned)abs(v - (int)cb[tid - 1]) : 0u);
        const unsigned vs = __reduce_add_sync(0xffffffffu, j ? (unsigned)abs(v - (int)cb[tid - 8]) : 0u);
        if ((tid & 31) == 0) { grad[tid >> 5][0] = hs; grad[tid >> 5][1] = vs; }
    }
    uint32_t c[16];
#pragma unroll
    for (int t = 0; t < 16; t++) c[t] = curw[t];
    unsigned long long mine = ~0ull;
    for (int cand = tid; cand < 625; cand += 128) {
        const int dy = cand / 25, dx = cand % 25;     // offsets already shifted by +12
        const uint32_t *w = win[dx & 3] + dy * 9 + (dx >> 2);
        uint32_t sad = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            sad = sad4(c[2 * j], w[j * 9], sad);
            sad = sad4(c[2 * j + 1], w[j * 9 + 1], sad);
        }
        const unsigned long long key = ((unsigned long long)(sad + abs(dx - kCmeRange) + abs(dy - kCmeRange)) << 10) | (unsigned)cand;
        mine = key < mine ? key : mine;
    }
    for (int off = 16; off; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, mine, off);
        mine = o < mine ? o : mine;
    }
    if ((tid & 31) == 0) best[tid >> 5] = mine;
    __syncthreads();
    if (tid == 0) {
        unsigned long long b = best[0];
        for (int k = 1; k < 4; k++) b = best[k] < b ? best[k] : b;
        const int cbest = (int)(b & 1023);
        int16_t *out = p.cmv + ((size_t)f * g.ctuw * g.ctuh + blockIdx.x) * 2;
        out[0] = (int16_t)(cbest % 25 - kCmeRange);
        out[1] = (int16_t)(cbest / 25 - kCmeRange);
        const unsigned hs = grad[0][0] + grad[1][0], vs = grad[0][1] + grad[1][1];
        atomicAdd(&p.scene[f].inter, b >> 10);
        atomicAdd(&p.scene[f].intra, (unsigned long long)min(hs, vs));
    }
}

// ================================================================================================ inter frame
struct __align__(128) WarpScratch {
    pixel win[28][40];          // reference window: 28 rows of 80 bytes, each written by one TMA bulk copy (cp.async.bulk)
    pixel src[16][16];
    int16_t tmpT[16][24];       // horizontal-pass output, TRANSPOSED: [column][row], rows 0..22 used
    pixel pred[16][16];
    int16_t a[16][18], b[16][18];
    unsigned long long mbar;    // mbarrier the TMA load completes on
};

// ---- TMA (cp.async.bulk.tensor) + mbarrier helpers, PTX as in the CUDA 12.9 ISA
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (SASS: UBLKCP): 16-byte aligned source / destination, size a multiple of 16
__device__ __forceinline__ void tma_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
                 : "memory");
}
// TMA tiled copy global -> shared (SASS: UTMALDG.2D): one instruction moves the whole box.  The innermost coordinate times the
// element size must be a multiple of 16 bytes (profiles/tma_probe_r2.log: x = 37 faults with 'illegal instruction', x = 40 works).
__device__ __forceinline__ void tma_tile_load_2d(uint32_t dst, const CUtensorMap *map, int x, int y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (unsigned spins = 0; !done; spins++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (spins > (1u << 26)) __trap();       // a broken descriptor must fail loudly, not hang the device
    }
}

struct MV { int x, y; };

__device__ __forceinline__ MV clamp_mv(const Geom &g, int x0, int y0, MV m)
{
    m.x = clampd(m.x, -(x0 + kMvOverhang) * 4, (g.wc - x0 - 16 + kMvOverhang) * 4);
    m.y = clampd(m.y, -(y0 + kMvOverhang) * 4, (g.hc - y0 - 16 + kMvOverhang) * 4);
    return m;
}

__device__ __forceinline__ MV ctu_mv(const Geom &g, const int16_t *cmv, int tx, int ty)
{
    tx = clampd(tx, 0, g.ctuw - 1); ty = clampd(ty, 0, g.ctuh - 1);
    const int16_t *c = cmv + (ty * g.ctuw + tx) * 2;
    return MV{c[0] * 16, c[1] * 16};
}

// 8-tap filter on sample pairs held as 32-bit words: result for the sample at position `rel` (in samples) of w[]
//   even rel: taps (t0,t1)(t2,t3)(t4,t5)(t6,t7) on words rel/2 ..;  odd rel: (0,t0)(t1,t2)(t3,t4)(t5,t6)(t7,0) on words (rel-1)/2 ..
template <int REL>
__device__ __forceinline__ int tap8(const int (&w)[8], const int (&t)[5], int init = 0)
{
    if constexpr ((REL & 1) == 0) {
        constexpr int b = REL / 2;
        int acc = __dp2a_lo(w[b], t[0], init);
        acc = __dp2a_hi(w[b + 1], t[0], acc);
        acc = __dp2a_lo(w[b + 2], t[1], acc);
        return __dp2a_hi(w[b + 3], t[1], acc);
    } else {
        constexpr int b = (REL - 1) / 2;
        int acc = __dp2a_lo(w[b], t[2], init);
        acc = __dp2a_hi(w[b + 1], t[2], acc);
        acc = __dp2a_lo(w[b + 2], t[3], acc);
        acc = __dp2a_hi(w[b + 3], t[3], acc);
        return __dp2a_lo(w[b + 4], t[4], acc);
    }
}

template <int PAR>
__device__ __forceinline__ void hpass8(const int (&w)[8], const int (&t)[5], int shift1, int16_t *dst /* tmpT[c0][r], column stride 24 */)
{
    dst[0 * 24] = (int16_t)(tap8<0 + PAR>(w, t) >> shift1);
    dst[1 * 24] = (int16_t)(tap8<1 + PAR>(w, t) >> shift1);
    dst[2 * 24] = (int16_t)(tap8<2 + PAR>(w, t) >> shift1);
    dst[3 * 24] = (int16_t)(tap8<3 + PAR>(w, t) >> shift1);
    dst[4 * 24] = (int16_t)(tap8<4 + PAR>(w, t) >> shift1);
    dst[5 * 24] = (int16_t)(tap8<5 + PAR>(w, t) >> shift1);
    dst[6 * 24] = (int16_t)(tap8<6 + PAR>(w, t) >> shift1);
    dst[7 * 24] = (int16_t)(tap8<7 + PAR>(w, t) >> shift1);
}

// Luma prediction of the 16x16 block whose integer sample (0,0) sits at window position (ix, iy), in the unified two-pass
// form (bit-exact with the normative one-pass cases).  Both passes run on packed sample pairs with dp2a.
// The horizontal pass filters `rows` (23 or 24) window rows starting at iy0 - 3 into tmpT[column][row]; the vertical pass
// reads 16 of them starting at row `ro` (0 or 1), so up to three candidates that differ only vertically share one
// horizontal pass.  Lane mapping of the result: lane -> column c = lane >> 1, rows r0 = 8 * (lane & 1) .. r0 + 7.
__device__ __forceinline__ void interp_hpass(WarpScratch &s, int ix, int iy0, int fx, int rows, int bd, int lane)
{
    const int shift1 = bd - 8;
    int th[5];
#pragma unroll
    for (int k = 0; k < 5; k++) th[k] = c_luma_pack[fx][k];
    if (fx == 0) {           // integer column: taps (0,0,0,64,0,0,0,0) -- a shift, no filtering
        const int sh = 6 - shift1;
        for (int task = lane; task < 2 * rows; task += 32) {
            const int r = task >> 1, c0 = (task & 1) * 8;
            const pixel *wp = &s.win[iy0 - 3 + r][ix + c0];
#pragma unroll
            for (int k = 0; k < 8; k++) s.tmpT[c0 + k][r] = (int16_t)((int)wp[k] << sh);
        }
        __syncwarp();
        return;
    }
    for (int task = lane; task < 2 * rows; task += 32) {
        const int r = task >> 1, c0 = (task & 1) * 8;
        const int s0 = ix - 3 + c0, w0 = s0 >> 1;
        const int *wp = reinterpret_cast<const int *>(&s.win[iy0 - 3 + r][0]) + w0;
        int w[8];
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = wp[k];
        int16_t *dst = &s.tmpT[c0][r];
        if (s0 & 1) hpass8<1>(w, th, shift1, dst);
        else hpass8<0>(w, th, shift1, dst);
    }
    __syncwarp();
}

__device__ __forceinline__ void interp_vpass(WarpScratch &s, int ro, int fy, int bd, int lane, int (&pv)[8])
{
    // ((acc >> 6) + off14) >> s14 == (acc + (off14 << 6)) >> (6 + s14) (nested floor divisions), so the rounding offset is the
    // accumulator's start value and one shift remains
    const int s14 = 14 - bd, rnd = (1 << (s14 - 1)) << 6, sh = 6 + s14, maxv = (1 << bd) - 1;
    int tv[5];
#pragma unroll
    for (int k = 0; k < 5; k++) tv[k] = c_luma_pack[fy][k];
    // one column segment of 8 outputs per lane, two 128-bit loads
    const int c = lane >> 1, r0 = (lane & 1) * 8;
    const int4 lo = *reinterpret_cast<const int4 *>(&s.tmpT[c][r0]), hi = *reinterpret_cast<const int4 *>(&s.tmpT[c][r0 + 8]);
    const int w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    int acc[8];
    if (ro) {
        acc[0] = tap8<1>(w, tv, rnd); acc[1] = tap8<2>(w, tv, rnd); acc[2] = tap8<3>(w, tv, rnd); acc[3] = tap8<4>(w, tv, rnd);
        acc[4] = tap8<5>(w, tv, rnd); acc[5] = tap8<6>(w, tv, rnd); acc[6] = tap8<7>(w, tv, rnd); acc[7] = tap8<8>(w, tv, rnd);
    } else {
        acc[0] = tap8<0>(w, tv, rnd); acc[1] = tap8<1>(w, tv, rnd); acc[2] = tap8<2>(w, tv, rnd); acc[3] = tap8<3>(w, tv, rnd);
        acc[4] = tap8<4>(w, tv, rnd); acc[5] = tap8<5>(w, tv, rnd); acc[6] = tap8<6>(w, tv, rnd); acc[7] = tap8<7>(w, tv, rnd);
    }
#pragma unroll
    for (int k = 0; k < 8; k++)
        pv[k] = clamp0(acc[k] >> sh, maxv);
}

__device__ __forceinline__ void interp_cols(WarpScratch &s, int ix, int iy, int fx, int fy, int bd, int lane, int (&pv)[8])
{
    if ((fx | fy) == 0) {        // integer position (e.g. the zero vector): the two-pass filter reduces to the identity
        const pixel *w = &s.win[iy + (lane & 1) * 8][ix + (lane >> 1)];
#pragma unroll
        for (int k = 0; k < 8; k++) pv[k] = w[k * 40];
        return;
    }
    interp_hpass(s, ix, iy, fx, 23, bd, lane);
    interp_vpass(s, 0, fy, bd, lane, pv);
    __syncwarp();
}

// SATD of the 16x16 block from column segments held in registers: vertical 4-point Hadamard in the lane, horizontal one
// across the four lanes that hold the four columns of a 4x4 block (lane xor 2, xor 4).  Sum of |H D H| over all 16 blocks
// is even per block, so the per-block ">> 1" of satd_4x4 can be applied to the total.
__device__ __forceinline__ int satd_cols(const int (&st)[8], const int (&pv)[8], int lane)
{
    int v[8];
#pragma unroll
    for (int g = 0; g < 8; g += 4) {
        const int d0 = st[g] - pv[g], d1 = st[g + 1] - pv[g + 1], d2 = st[g + 2] - pv[g + 2], d3 = st[g + 3] - pv[g + 3];
        const int a0 = d0 + d1, a1 = d0 - d1, a2 = d2 + d3, a3 = d2 - d3;
        v[g] = a0 + a2; v[g + 1] = a1 + a3; v[g + 2] = a0 - a2; v[g + 3] = a1 - a3;
    }
#pragma unroll
    for (int m = 2; m <= 4; m <<= 1) {
        int sgn = (lane & m) ? -1 : 1;                // butterfly as one multiply-add per value: partner + sign * own
        asm("" : "+r"(sgn));                          // opaque, or the compiler turns the product back into select + negate + add
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = __shfl_xor_sync(0xffffffffu, v[k], m) + sgn * v[k];
    }
    int sum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) sum += abs(v[k]);
    return warp_sum(sum) >> 1;
}

// Stage the 28-row x 40-sample luma window around integer position `centre` (quarter-sample vector) with ONE tiled TMA load
// (cp.async.bulk.tensor.2d, box 40 x 28) completing on the warp's mbarrier (first use of the barrier: phase 0).  The column is
// rounded down to 16 bytes (8 samples) -- the alignment the tiled form needs -- and the remainder is returned.  Covers +-2
// integer positions plus the 8-tap support; the padded plane keeps the box inside the tensor, so nothing is ever zero-filled.
__device__ __forceinline__ int stage_window(const InterParams &p, uint32_t bar, uint32_t win_smem, int x0, int y0, MV centre, int lane)
{
    const int wx = x0 + (centre.x >> 2) - 6, wy = y0 + (centre.y >> 2) - 6;
    const int ax = wx & ~7, woff = wx - ax;
    if (lane == 0) {
        mbar_expect_tx(bar, 28 * 40 * sizeof(pixel));
        tma_tile_load_2d(win_smem, &p.ref_map, ax + kPad, wy + kPad, bar);      // tensor coordinates count from the padded allocation
    }
    __syncwarp();
    mbar_wait(bar, 0);
    return woff;
}

__device__ __forceinline__ uint32_t pack_mv(MV m) { return (uint32_t)(uint16_t)m.x | ((uint32_t)(uint16_t)m.y << 16); }
__device__ __forceinline__ MV unpack_mv(uint32_t v) { return MV{(int)(int16_t)(v & 0xffff), (int)(int16_t)(v >> 16)}; }

// One merge-aware pass for one CU (warp-uniform): compare the CU's own vector with the vectors its five merge-candidate
// neighbours (A1, B1, B0, A0, B2) held after the previous pass and with the zero vector (oracle/hevc_encode.c, "merge-aware
// passes").  The window in `s.win` is centred on the own vector; candidates further than +-2 integer samples are not tried.
__device__ __forceinline__ void merge_decide(const InterParams &p, WarpScratch &s, int cx, int cy, int lambda, int wx0, int wy0,
                                             const int (&st)[8], int lane, MV &best, int &bsatd, int &bcost_out)
{
    const Geom &g = p.g;
    const int x0 = cx * 16, y0 = cy * 16, idx = cy * g.cuw + cx;
    const MV own = unpack_mv(p.mv_in[idx]);
    // lane k < 6 owns candidate k (A1, B1, B0, A0, B2, zero): its vector, validity, duplicate and range checks are computed
    // once, in parallel; the evaluation loop then walks the surviving candidates in order
    bool okk = lane == 5;
    uint32_t mine = 0;                                  // packed vector of this lane's candidate (zero vector for lane 5)
    if (lane < 5) {
        const int nx = cx + (lane == 2 ? 1 : lane == 1 ? 0 : -1), ny = cy + (lane == 0 ? 0 : lane == 3 ? 1 : -1);
        okk = nx >= 0 && ny >= 0 && nx < g.cuw && ny < g.cuh;
        if (okk) mine = p.mv_in[ny * g.cuw + nx];
    }
    const MV m = unpack_mv(mine);
    const bool same_as_own = m.x == own.x && m.y == own.y;
    const int bts = !okk ? 0x7fffffff : same_as_own ? 0 : mv_bits1(own.x - m.x) + mv_bits1(own.y - m.y);
    const int own_bits = __reduce_min_sync(0xffffffffu, bts);
    const unsigned okmask = __ballot_sync(0xffffffffu, okk);
    bool dup = same_as_own;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const uint32_t mj = __shfl_sync(0xffffffffu, mine, j);
        dup |= j < lane && ((okmask >> j) & 1) && mj == mine;
    }
    const MV cm = clamp_mv(g, x0, y0, m);
    const bool eval = okk && !dup && cm.x == m.x && cm.y == m.y && abs((m.x >> 2) - (own.x >> 2)) <= 2 && abs((m.y >> 2) - (own.y >> 2)) <= 2;
    unsigned todo = __ballot_sync(0xffffffffu, eval) & 0x3fu;
    best = own;
    bsatd = p.satd_in[idx];
    int bcost = bsatd + ((lambda * (own_bits + 2)) >> 8) + 1;         // +1: ties go to a merge candidate
    // (loop kept rolled: six inlined copies of interpolation + SATD made the kernel miss the instruction cache)
#pragma unroll 1
    while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        const MV c = unpack_mv(__shfl_sync(0xffffffffu, mine, k));
        int pv[8];
        interp_cols(s, (c.x >> 2) - wx0, (c.y >> 2) - wy0, c.x & 3, c.y & 3, g.bit_depth, lane, pv);
        const int sd = satd_cols(st, pv, lane);
        const int cost = sd + ((lambda * 2) >> 8);
        if (cost < bcost) { bcost = cost; best = c; bsatd = sd; }
    }
    bcost_out = bcost;
}

// pass 1 of the motion search: per CU, integer candidates -> 5x5 -> half / quarter sample; writes the vector field and its SATD
__global__ void __launch_bounds__(128, 8) k_me(const __grid_constant__ InterParams p)
{
    __shared__ WarpScratch scratch[4];
    const Geom &g = p.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx = blockIdx.x % g.ctuw, ty = blockIdx.x / g.ctuw;
    const int cx = 2 * tx + (warp & 1), cy = 2 * ty + (warp >> 1);
    if (cx >= g.cuw || cy >= g.cuh || p.ctl->is_idr)        // (frame types are decided on the device: key frames skip the inter kernels)
        return;
    WarpScratch &s = scratch[warp];
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s.mbar), win_smem = (uint32_t)__cvta_generic_to_shared(&s.win[0][0]);
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    const int x0 = cx * 16, y0 = cy * 16, bd = g.bit_depth;
    const int lambda = p.ctl->lambda;
    const int row = lane >> 1, c0 = (lane & 1) * 8;

    // source block: shared memory (for SATD) + its 8 most significant bits as packed bytes in registers (for the integer SADs)
    const int sh8 = bd - 8, lambda8 = lambda >> sh8;          // the integer stages compare 8-bit samples, lambda at 8-bit scale
    uint32_t s8lo, s8hi;
    {
        const uint4 v = *reinterpret_cast<const uint4 *>(p.src.y + (size_t)(y0 + row) * g.src_stride + x0 + c0);
        *reinterpret_cast<uint4 *>(&s.src[row][c0]) = v;
        s8lo = pack4_msb8(v.x, v.y, sh8);
        s8hi = pack4_msb8(v.z, v.w, sh8);
    }
    const MV pred = clamp_mv(g, x0, y0, ctu_mv(g, p.cmv, tx, ty));
    MV best{0, 0};
    int bcost = 0x7fffffff;
    // ---- integer stage 1: six candidates straight from global memory (lane -> row lane >> 1, 8 samples).  Lane k < 6 fetches
    //      and clamps candidate k and prices its vector once; the loop broadcasts them.
    uint32_t my_c = 0;
    int my_vc = 0;
    if (lane < 6) {
        MV c = lane == 0 ? MV{0, 0} : lane == 1 ? pred : lane == 2 ? ctu_mv(g, p.cmv, tx - 1, ty) : lane == 3 ? ctu_mv(g, p.cmv, tx, ty - 1)
               : lane == 4 ? ctu_mv(g, p.cmv, tx + 1, ty) : ctu_mv(g, p.cmv, tx, ty + 1);
        c = clamp_mv(g, x0, y0, c);
        my_c = pack_mv(c);
        my_vc = mv_cost(lambda8, c.x, c.y, pred.x, pred.y);
    }
#pragma unroll 1
    for (int k = 0; k < 6; k++) {
        const MV c = unpack_mv(__shfl_sync(0xffffffffu, my_c, k));
        const int vc = __shfl_sync(0xffffffffu, my_vc, k);
        const pixel *r = p.ref.y + (ptrdiff_t)(y0 + (c.y >> 2) + row) * g.rec_stride + x0 + (c.x >> 2) + c0;
        uint32_t rw[4];
#pragma unroll
        for (int i = 0; i < 4; i++) rw[i] = (uint32_t)r[2 * i] | ((uint32_t)r[2 * i + 1] << 16);
        const uint32_t sad = sad4(s8hi, pack4_msb8(rw[2], rw[3], sh8), sad4(s8lo, pack4_msb8(rw[0], rw[1], sh8), 0));
        const int cost = warp_sum((int)sad) + vc;
        if (cost < bcost) { bcost = cost; best = c; }
    }
    const MV centre = best;
    const int woff = stage_window(p, bar, win_smem, x0, y0, centre, lane);
    // ---- integer stage 2: 5x5 square by byte SAD.  The window's 8 MSBs are packed once into a byte copy (aliased onto the
    //      interpolation scratch, unused so far); a candidate row is three aligned words funnel-shifted to the block position.
    {
        uint32_t *win8 = reinterpret_cast<uint32_t *>(&s.tmpT[0][0]);           // [28][10] words = 28 rows x 40 bytes
        for (int t = lane, r = (lane * 13) >> 7, c = lane - 10 * ((lane * 13) >> 7); t < 28 * 10; t += 32) {       // r = t / 10, c = t % 10
            const uint2 v = *reinterpret_cast<const uint2 *>(&s.win[r][c * 4]);
            win8[t] = pack4_msb8(v.x, v.y, sh8);
            c += 2; r += 3;                                                   // t += 32
            if (c >= 10) { c -= 10; r++; }
        }
        __syncwarp();
        // lane t < 25 owns position t (raster over dy, dx): its validity and vector cost are computed once, in parallel
        int my_cost = -1;
        if (lane < 25 && lane != 12) {
            const MV m{centre.x + 4 * (lane % 5 - 2), centre.y + 4 * (lane / 5 - 2)};
            const MV cm = clamp_mv(g, x0, y0, m);
            if (cm.x == m.x && cm.y == m.y) my_cost = mv_cost(lambda8, m.x, m.y, pred.x, pred.y);
        }
        unsigned todo = __ballot_sync(0xffffffffu, my_cost >= 0);             // valid positions, bit t = raster index
        const uint32_t *wrow = win8 + (4 + row) * 10;                         // window row of dy = -2
        const int colb0 = 4 + woff + c0;                                      // byte column of this lane's first sample at dx = -2
#pragma unroll 1
        while (todo) {
            const int t = __ffs(todo) - 1;
            todo &= todo - 1;
            const int vc = __shfl_sync(0xffffffffu, my_cost, t);
            const int dyi = (t * 13) >> 6, dxi = t - 5 * dyi;                 // t / 5, t % 5 for t < 25
            const int colb = colb0 + dxi;
            const uint32_t *w = wrow + dyi * 10 + (colb >> 2);
            const int fs = (colb & 3) * 8;
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
            const uint32_t sad = sad4(s8hi, __funnelshift_r(w1, w2, fs), sad4(s8lo, __funnelshift_r(w0, w1, fs), 0));
            const int cost = warp_sum((int)sad) + vc;
            if (cost < bcost) { bcost = cost; best = MV{centre.x + 4 * (dxi - 2), centre.y + 4 * (dyi - 2)}; }
        }
        __syncwarp();
    }
    // ---- sub-sample stages: SATD on the normative interpolation
    const int wx0 = (centre.x >> 2) - 6 - woff, wy0 = (centre.y >> 2) - 6;     // window origin (sample [0][0]) relative to the block position
    int st[8], pv[8];        // source / prediction column segments: column lane >> 1, rows 8 * (lane & 1) + k
#pragma unroll
    for (int k = 0; k < 8; k++) st[k] = s.src[(lane & 1) * 8 + k][lane >> 1];
    {   // `best` is an integer position here: its prediction is the window itself (the two-pass filter reduces to the identity)
        const pixel *w = &s.win[(best.y >> 2) - wy0 + (lane & 1) * 8][(best.x >> 2) - wx0 + (lane >> 1)];
#pragma unroll
        for (int k = 0; k < 8; k++) pv[k] = w[k * 40];
    }
    const int satd_int = satd_cols(st, pv, lane);
    bcost = satd_int + mv_cost(lambda, best.x, best.y, pred.x, pred.y);
    // a block the integer vector already predicts to within a quarter grey level per sample (static, clean content) has nothing
    // to gain from sub-sample refinement: the search stops here (warp-uniform)
#pragma unroll 1
    for (int step = satd_int > (64 << (bd - 8)) ? 2 : 0; step >= 1; step--) {
        const MV c2 = best;
        const int iy_min = (c2.y - step) >> 2;                    // the three rows of candidates start at integer row iy_min or iy_min + 1
        // lane t < 9 owns candidate (dx, dy) = (t / 3 - 1, t % 3 - 1): validity and vector cost once, in parallel
        int my_cost = -1;
        if (lane < 9 && lane != 4) {
            const MV m{c2.x + step * (lane / 3 - 1), c2.y + step * (lane % 3 - 1)};
            const MV cm = clamp_mv(g, x0, y0, m);
            if (cm.x == m.x && cm.y == m.y) my_cost = mv_cost(lambda, m.x, m.y, pred.x, pred.y);
        }
        const unsigned okmask = __ballot_sync(0xffffffffu, my_cost >= 0);
#pragma unroll 1
        for (int dx = -1; dx <= 1; dx++) {
            const int mx = c2.x + step * dx;
            if (!((okmask >> (3 * (dx + 1))) & 7u)) continue;
            interp_hpass(s, (mx >> 2) - wx0, iy_min - wy0, mx & 3, 24, bd, lane);       // shared by the column's candidates
#pragma unroll
            for (int dy = -1; dy <= 1; dy++) {
                const int t = 3 * (dx + 1) + dy + 1;
                if (!((okmask >> t) & 1)) continue;
                const int vc = __shfl_sync(0xffffffffu, my_cost, t);
                const MV m{mx, c2.y + step * dy};
                interp_vpass(s, (m.y >> 2) - iy_min, m.y & 3, bd, lane, pv);
                const int cost = satd_cols(st, pv, lane) + vc;
                if (cost < bcost) { bcost = cost; best = m; }
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        p.mv_out[cy * g.cuw + cx] = pack_mv(best);
        p.satd_out[cy * g.cuw + cx] = bcost - mv_cost(lambda, best.x, best.y, pred.x, pred.y);
    }
}

// one merge-aware (Jacobi) pass over the motion field: mv_in / satd_in -> mv_out / satd_out
__global__ void __launch_bounds__(128, 7) k_merge(const __grid_constant__ InterParams p)
{
    __shared__ WarpScratch scratch[4];
    const Geom &g = p.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx = blockIdx.x % g.ctuw, ty = blockIdx.x / g.ctuw;
    const int cx = 2 * tx + (warp & 1), cy = 2 * ty + (warp >> 1);
    if (cx >= g.cuw || cy >= g.cuh || p.ctl->is_idr)        // (frame types are decided on the device: key frames skip the inter kernels)
        return;
    WarpScratch &s = scratch[warp];
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s.mbar), win_smem = (uint32_t)__cvta_generic_to_shared(&s.win[0][0]);
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    const int x0 = cx * 16, y0 = cy * 16, bd = g.bit_depth;
    const int lambda = p.ctl->lambda;
    const int row = lane >> 1, c0 = (lane & 1) * 8;
    *reinterpret_cast<uint4 *>(&s.src[row][c0]) = *reinterpret_cast<const uint4 *>(p.src.y + (size_t)(y0 + row) * g.src_stride + x0 + c0);
    const MV own = unpack_mv(p.mv_in[cy * g.cuw + cx]);
    const int woff = stage_window(p, bar, win_smem, x0, y0, own, lane);
    const int wx0 = (own.x >> 2) - 6 - woff, wy0 = (own.y >> 2) - 6;
    int st[8];
#pragma unroll
    for (int k = 0; k < 8; k++) st[k] = s.src[(lane & 1) * 8 + k][lane >> 1];
    MV best;
    int bsatd, bcost;
    merge_decide(p, s, cx, cy, lambda, wx0, wy0, st, lane, best, bsatd, bcost);
    if (lane == 0) {
        p.mv_out[cy * g.cuw + cx] = pack_mv(best);
        p.satd_out[cy * g.cuw + cx] = bsatd;
        atomicAdd(&p.ctl->satd_sum, (unsigned long long)bsatd);
    }
}

// last merge-aware pass + reconstruction: prediction, transform, quantisation and reconstruction of every CU of a P frame
__global__ void __launch_bounds__(128, 7) k_inter(const __grid_constant__ InterParams p)
{
    __shared__ WarpScratch scratch[4];
    const Geom &g = p.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx = blockIdx.x % g.ctuw, ty = blockIdx.x / g.ctuw;
    const int cx = 2 * tx + (warp & 1), cy = 2 * ty + (warp >> 1);
    if (cx >= g.cuw || cy >= g.cuh || p.ctl->is_idr)        // (frame types are decided on the device: key frames skip the inter kernels)
        return;
    WarpScratch &s = scratch[warp];
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s.mbar), win_smem = (uint32_t)__cvta_generic_to_shared(&s.win[0][0]);
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    const int x0 = cx * 16, y0 = cy * 16, bd = g.bit_depth;
    const int maxv = (1 << bd) - 1;
    const FrameCtl ctl = *p.ctl;
    const int lambda = ctl.lambda;
    const int row = lane >> 1, c0 = (lane & 1) * 8;

    // source block: registers (for SAD) + shared memory (for SATD)
    int sp[8];
    {
        const uint4 v = *reinterpret_cast<const uint4 *>(p.src.y + (size_t)(y0 + row) * g.src_stride + x0 + c0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) { sp[2 * k] = w[k] & 0xffff; sp[2 * k + 1] = w[k] >> 16; }
        *reinterpret_cast<uint4 *>(&s.src[row][c0]) = v;
    }
    const MV own = unpack_mv(p.mv_in[cy * g.cuw + cx]);
    const int woff = stage_window(p, bar, win_smem, x0, y0, own, lane);
    const int wx0 = (own.x >> 2) - 6 - woff, wy0 = (own.y >> 2) - 6;
    int st[8], pv[8];        // source / prediction column segments: column lane >> 1, rows 8 * (lane & 1) + k
#pragma unroll
    for (int k = 0; k < 8; k++) st[k] = s.src[(lane & 1) * 8 + k][lane >> 1];
    MV best;
    {
        int bsatd, bcost;
        merge_decide(p, s, cx, cy, lambda, wx0, wy0, st, lane, best, bsatd, bcost);
        // Intra CU in a P frame (oracle/hevc_encode.c): the best intra prediction searched on source neighbours, plus its
        // signalling, against 9/8 of the final inter choice (the SATD of the source-neighbour search understates what the
        // reconstructed-neighbour prediction and its flatter residual save); the wavefront kernel then reconstructs the CU
        if (p.intra_best) {
            const int ibest = p.intra_best[cy * g.cuw + cx];
            const long long icost = (long long)ibest + ((lambda * 12) >> 8);
            if (ibest != 0x7fffffff && icost * 8 < (long long)bcost * 9) {
                if (lane == 0) {
                    CuInfo ci;
                    ci.pred_mode = 0; ci.intra_mode = 0; ci.cbf = 0; ci.skip = 0; ci.mvx = 0; ci.mvy = 0;
                    p.cus[cy * g.cuw + cx] = ci;
                }
                return;
            }
        }
    }
    // ---- luma: predict, transform, quantise, reconstruct
    interp_cols(s, (best.x >> 2) - wx0, (best.y >> 2) - wy0, best.x & 3, best.y & 3, bd, lane, pv);
#pragma unroll
    for (int k = 0; k < 8; k++) s.pred[(lane & 1) * 8 + k][lane >> 1] = (pixel)pv[k];    // back to row-major for the residual path
    __syncwarp();
    int16_t *coef = p.coefs + (size_t)(cy * g.cuw + cx) * kCuCoefs;
    int pr[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        pr[i] = s.pred[row][c0 + i];
        s.a[row][c0 + i] = (int16_t)(sp[i] - pr[i]);
    }
    __syncwarp();
    if (lane < 16) fwd_line<16, false>(&s.a[lane][0], 1, &s.b[0][lane], 18, 3 + (bd - 8));
    __syncwarp();
    if (lane < 16) fwd_line<16, false>(&s.b[lane][0], 1, &s.a[0][lane], 18, 10);
    __syncwarp();
    const QuantParam qy = ctl.qy;
    int lv[8];
    bool nz = false;
    int e_nnz = 0, e_slog = 0;         // size estimate (rate control): non-zero count, sum floor(log2 |level|)
    unsigned nz_lo = 0, nz_hi = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        lv[i] = quant_one(s.a[row][c0 + i], qy);
        if (lv[i]) {
            e_nnz++;
            e_slog += 31 - __clz(abs(lv[i]));
            if (i < 4) nz_lo = 1; else nz_hi = 1;
        }
        nz |= lv[i] != 0;
    }
    {   // a block that carries nothing but one or two isolated +-1 levels is dropped (oracle/hevc_encode.c, code_block)
        const int nnz_y = warp_sum(e_nnz), slog_y = warp_sum(e_slog);
        if (nnz_y > 0 && nnz_y <= 2 && slog_y == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) lv[i] = 0;
            nz = false; e_nnz = 0; e_slog = 0; nz_lo = 0; nz_hi = 0;
        }
    }
    int e_nsb;
    {   // coded 4x4 sub-blocks of the 16x16 block: a sub-block is 4 rows (lanes l, l+2, l+4, l+6) x one half-row group
        const unsigned mlo = __ballot_sync(0xffffffffu, nz_lo), mhi = __ballot_sync(0xffffffffu, nz_hi);
        const int sy = (lane >> 2) & 3, sxx = lane & 3, h = sxx >> 1;
        const unsigned m = (sxx & 1) ? mhi : mlo;
        const bool f = lane < 16 && ((m >> (8 * sy + h)) & 0x55u) != 0;
        e_nsb = __popc(__ballot_sync(0xffffffffu, f));
    }
    {
        uint4 o;
        o.x = (uint32_t)(uint16_t)lv[0] | ((uint32_t)(uint16_t)lv[1] << 16);
        o.y = (uint32_t)(uint16_t)lv[2] | ((uint32_t)(uint16_t)lv[3] << 16);
        o.z = (uint32_t)(uint16_t)lv[4] | ((uint32_t)(uint16_t)lv[5] << 16);
        o.w = (uint32_t)(uint16_t)lv[6] | ((uint32_t)(uint16_t)lv[7] << 16);
        *reinterpret_cast<uint4 *>(coef + row * 16 + c0) = o;
    }
    const bool cbf_y = __any_sync(0xffffffffu, nz);
    int rec[8];
    if (cbf_y) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            s.a[row][c0 + i] = (int16_t)dequant_one(lv[i], qy);
        __syncwarp();
        if (lane < 16) inv_line<16, false>(&s.a[0][lane], 18, &s.b[lane][0], 1, 7);
        __syncwarp();
        if (lane < 16) inv_line<16, false>(&s.b[0][lane], 18, &s.a[lane][0], 1, 12 - (bd - 8));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; i++)
            rec[i] = clamp0(pr[i] + s.a[row][c0 + i], maxv);
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++)
            rec[i] = pr[i];
    }
    {
        uint4 o;
        o.x = rec[0] | (rec[1] << 16); o.y = rec[2] | (rec[3] << 16); o.z = rec[4] | (rec[5] << 16); o.w = rec[6] | (rec[7] << 16);
        *reinterpret_cast<uint4 *>(p.rec.y + (size_t)(y0 + row) * g.rec_stride + x0 + c0) = o;
    }
    __syncwarp();
    // ---- chroma: both planes at once (lanes 0-15 -> Cb, 16-31 -> Cr for the per-sample stages)
    const int fxc = best.x & 7, fyc = best.y & 7;
    const int shift1 = bd - 8, s14 = 14 - bd, off14 = 1 << (s14 - 1);
    pixel *cwin = &s.win[0][0];                 // [2][11][12]
    int16_t *ctmp = &s.tmpT[0][0];              // [2][11][8]
    {
        const ptrdiff_t off = (ptrdiff_t)(cy * 8 + (best.y >> 3) - 1) * g.recc_stride + cx * 8 + (best.x >> 3) - 1;
        // lane -> (plane, column); the 11 rows are 11 independent loads in flight (no per-sample index arithmetic)
        const int pl = lane >> 4, c = lane & 15;
        if (c < 11) {
            const pixel *src = (pl ? p.ref.v : p.ref.u) + off + c;
            pixel v[11];
#pragma unroll
            for (int r = 0; r < 11; r++) v[r] = src[(ptrdiff_t)r * g.recc_stride];
#pragma unroll
            for (int r = 0; r < 11; r++) cwin[pl * 132 + r * 12 + c] = v[r];
        }
    }
    __syncwarp();
    for (int o = lane; o < 2 * 11 * 8; o += 32) {
        const int pl = o / 88, r = (o % 88) >> 3, c = o & 7;
        const pixel *w = cwin + pl * 132 + r * 12 + c;
        int acc = 0;
#pragma unroll
        for (int t = 0; t < 4; t++)
            acc += c_chroma_taps[fxc][t] * w[t];
        ctmp[pl * 88 + r * 8 + c] = (int16_t)(acc >> shift1);
    }
    __syncwarp();
    // each lane: 4 consecutive samples of one plane: plane = lane >> 4, row = (lane & 15) >> 1, col0 = (lane & 1) * 4
    const int cpl = lane >> 4, crow = (lane & 15) >> 1, cc0 = (lane & 1) * 4;
    int cpr[4], csrc[4];
    {
        const pixel *sp2 = (cpl ? p.src.v : p.src.u) + (size_t)(cy * 8 + crow) * g.srcc_stride + cx * 8 + cc0;
        const uint2 v = *reinterpret_cast<const uint2 *>(sp2);
        csrc[0] = v.x & 0xffff; csrc[1] = v.x >> 16; csrc[2] = v.y & 0xffff; csrc[3] = v.y >> 16;
    }
    int16_t(*ca)[8][10] = reinterpret_cast<int16_t(*)[8][10]>(&s.a[0][0]);     // [2][8][10]
    int16_t(*cb)[8][10] = reinterpret_cast<int16_t(*)[8][10]>(&s.b[0][0]);
#pragma unroll
    for (int c = 0; c < 4; c++) {
        int acc = 0;
#pragma unroll
        for (int t = 0; t < 4; t++)
            acc += c_chroma_taps[fyc][t] * ctmp[cpl * 88 + (crow + t) * 8 + cc0 + c];
        cpr[c] = clamp0(((acc >> 6) + off14) >> s14, maxv);
        ca[cpl][crow][cc0 + c] = (int16_t)(csrc[c] - cpr[c]);
    }
    __syncwarp();
    if (lane < 16) fwd_line<8, false>(&ca[lane >> 3][lane & 7][0], 1, &cb[lane >> 3][0][lane & 7], 10, 2 + (bd - 8));
    __syncwarp();
    if (lane < 16) fwd_line<8, false>(&cb[lane >> 3][lane & 7][0], 1, &ca[lane >> 3][0][lane & 7], 10, 9);
    __syncwarp();
    const QuantParam qc = ctl.qc;
    int clv[4];
    bool cnz = false;
    int c_nnz = 0, c_slog = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        clv[c] = quant_one(ca[cpl][crow][cc0 + c], qc);
        if (clv[c]) {
            c_nnz++;
            c_slog += 31 - __clz(abs(clv[c]));
        }
        cnz |= clv[c] != 0;
    }
    e_nnz += c_nnz;
    e_slog += c_slog;
    *reinterpret_cast<uint2 *>(coef + 256 + cpl * 64 + crow * 8 + cc0) =
        make_uint2((uint32_t)(uint16_t)clv[0] | ((uint32_t)(uint16_t)clv[1] << 16), (uint32_t)(uint16_t)clv[2] | ((uint32_t)(uint16_t)clv[3] << 16));
    const unsigned cmask = __ballot_sync(0xffffffffu, cnz);
    const int cbf = (cbf_y ? 1 : 0) | ((cmask & 0xffffu) ? 2 : 0) | ((cmask >> 16) ? 4 : 0);
    {   // chroma sub-blocks: plane (lane >> 4), 4-row group ((lane >> 3) & 1), column half (lane & 1)
        const int pl = (lane >> 2) & 1, sy = (lane >> 1) & 1, h = lane & 1;
        const bool f = lane < 8 && ((cmask >> (16 * pl + 8 * sy + h)) & 0x55u) != 0;
        e_nsb += __popc(__ballot_sync(0xffffffffu, f));
        const int nnz = warp_sum(e_nnz), slog = warp_sum(e_slog);
        if (lane == 0)
            atomicAdd(&p.ctl->est16, cbf ? (unsigned long long)(47 * nnz + 22 * slog + 104 * e_nsb + 160) : 80ull);
    }
    if (cmask) {
#pragma unroll
        for (int c = 0; c < 4; c++)
            ca[cpl][crow][cc0 + c] = (int16_t)dequant_one(clv[c], qc);
        __syncwarp();
        if (lane < 16) inv_line<8, false>(&ca[lane >> 3][0][lane & 7], 10, &cb[lane >> 3][lane & 7][0], 1, 7);
        __syncwarp();
        if (lane < 16) inv_line<8, false>(&cb[lane >> 3][0][lane & 7], 10, &ca[lane >> 3][lane & 7][0], 1, 12 - (bd - 8));
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; c++)
            cpr[c] = clamp0(cpr[c] + ca[cpl][crow][cc0 + c], maxv);
    }
    {
        pixel *rp = (cpl ? p.rec.v : p.rec.u) + (size_t)(cy * 8 + crow) * g.recc_stride + cx * 8 + cc0;
        *reinterpret_cast<uint2 *>(rp) = make_uint2(cpr[0] | (cpr[1] << 16), cpr[2] | (cpr[3] << 16));
    }
    if (lane == 0) {
        CuInfo ci;
        ci.pred_mode = 1; ci.intra_mode = 1; ci.cbf = (uint8_t)cbf; ci.skip = 0;
        ci.mvx = (int16_t)best.x; ci.mvy = (int16_t)best.y;
        p.cus[cy * g.cuw + cx] = ci;
    }
}


// ================================================================================================ intra frame
struct IntraScratch {
    pixel nb[65], flt[65];
    pixel cnb[2][33];
    pixel src[16][16];
    int16_t a[16][18], b[16][18];         // luma residual / coefficient ping-pong
    int16_t ca[2][8][10], cb[2][8][10];   // chroma (both planes)
    int cost[35];
    int dc, best_mode;
    int e_nnz, e_slog, sbflag[24], cbf[3];
    int left_mode[2], top_mode[2];        // luma modes of the CUs left of / above the next CUs (most-probable-mode derivation)
    int lambda;
    QuantParam qy, qc;
    int16_t m16[16 * 16], m8[8 * 8];      // core transform matrices (H.265 8.6.4.2) for the matrix-form, all-threads transforms
};

// neighbour sample `i` of the (4N + 1)-sample reference array of an NxN block of CU (cx, cy), with the
// substitution rule of 8.4.4.2.2 evaluated in closed form per 5 availability segments
__device__ __forceinline__ pixel gather_one(const pixel *plane, int stride, const Geom &g, int cx, int cy, int N, int i, int bd)
{
    const int n2 = 2 * N, x0 = cx * N, y0 = cy * N;
    // availability of the five neighbour segments as a bit mask (bit s = segment s), kept in a register
    const unsigned av = (cu_avail(g, cx, cy, cx - 1, cy + 1) ? 1u : 0u) | (cu_avail(g, cx, cy, cx - 1, cy) ? 2u : 0u) |
                        (cu_avail(g, cx, cy, cx - 1, cy - 1) ? 4u : 0u) | (cu_avail(g, cx, cy, cx, cy - 1) ? 8u : 0u) |
                        (cu_avail(g, cx, cy, cx + 1, cy - 1) ? 16u : 0u);
    // walk order: bottom-left (bottom to top), left, corner, top, top-right
    int seg, px, py;
    if (i == 0) { seg = 2; px = -1; py = -1; }
    else if (i <= n2) { seg = i <= N ? 3 : 4; px = i - 1; py = -1; }
    else { const int k = i - 1 - n2; seg = k < N ? 1 : 0; px = -1; py = k; }
    if (!((av >> seg) & 1)) {
        int s2 = seg - 1;
        while (s2 >= 0 && !((av >> s2) & 1)) s2--;
        if (s2 >= 0) {          // last sample (in walk order) of the nearest earlier available segment
            if (s2 == 0) { px = -1; py = N; }
            else if (s2 == 1) { px = -1; py = 0; }
            else if (s2 == 2) { px = -1; py = -1; }
            else { px = N - 1; py = -1; }
        } else {                // first sample of the first later available segment
            s2 = seg + 1;
            while (s2 < 5 && !((av >> s2) & 1)) s2++;
            if (s2 >= 5) return (pixel)(1 << (bd - 1));
            if (s2 == 1) { px = -1; py = N - 1; }
            else if (s2 == 2) { px = -1; py = -1; }
            else if (s2 == 3) { px = 0; py = -1; }
            else { px = N; py = -1; }
        }
    }
    return __ldcg(plane + (ptrdiff_t)(y0 + py) * stride + x0 + px);
}

// entry (k, n) of the N-point core transform at run time (same folding as the compile-time odd_entry)
__device__ int tmat_rt(int N, int k, int n)
{
    while (true) {
        if (k == 0) return 64;
        if (N == 2) return n == 0 ? 64 : -64;
        if (k & 1) return odd_entry(N, k, n);
        n = n < N / 2 ? n : N - 1 - n;
        N >>= 1;
        k >>= 1;
    }
}

// one output sample of a separable transform stage in matrix form (exact integer sums, so identical to the butterflies):
//   forward: out[i][j] = (sum_t M[i][t] * in[j][t] + add) >> shift          inverse: out[i][j] = clip16((sum_t M[t][j] * in[t][i] + add) >> shift)
template <int N, int LD>
__device__ __forceinline__ void fwd_stage(const int16_t *M, const int16_t *in, int16_t *out, int i, int j, int shift)
{
    int acc = shift > 0 ? 1 << (shift - 1) : 0;
#pragma unroll
    for (int t = 0; t < N; t++) acc += M[i * N + t] * in[j * LD + t];
    out[i * LD + j] = (int16_t)(acc >> shift);
}
template <int N, int LD>
__device__ __forceinline__ void inv_stage(const int16_t *M, const int16_t *in, int16_t *out, int i, int j, int shift)
{
    int acc = 1 << (shift - 1);
#pragma unroll
    for (int t = 0; t < N; t++) acc += M[t * N + j] * in[t * LD + i];
    out[i * LD + j] = (int16_t)clampd(acc >> shift, -32768, 32767);
}

// One intra CU on a 384-thread CTA: threads 0-255 own one luma sample each, threads 256-383 one chroma sample each (both
// planes), so the luma and chroma pipelines run through the same nine barriers.  The mode is the argmin of the precomputed
// source-neighbour distortions (k_intra_search) plus the signalling cost against the real most-probable modes.
__device__ void intra_cu(IntraParams &p, IntraScratch &s, int cx, int cy)
{
    const Geom &g = p.g;
    const int tid = threadIdx.x, bd = g.bit_depth, maxv = (1 << bd) - 1;
    const int x0 = cx * 16, y0 = cy * 16, k = (cx & 1) | ((cy & 1) << 1);
    const bool luma = tid < 256, chroma = tid >= 256;
    const int py = (tid >> 4) & 15, px = tid & 15;
    const int ct = tid - 256, cpl = (ct >> 6) & 1, cyy = (ct >> 3) & 7, cxx = ct & 7;
    // ---- stage 0: neighbours (luma + both chroma planes), source samples, search result
    if (tid < 65) s.nb[tid] = gather_one(p.rec.y, g.rec_stride, g, cx, cy, 16, tid, bd);
    else if (tid >= 96 && tid < 96 + 66) {
        const int i = tid - 96, pl = i / 33;
        s.cnb[pl][i % 33] = gather_one(pl ? p.rec.v : p.rec.u, g.recc_stride, g, cx, cy, 8, i % 33, bd);
    } else if (tid >= 192 && tid < 192 + 35) {
        s.cost[tid - 192] = __ldcg(p.mode_cost + (size_t)(cy * g.cuw + cx) * 35 + tid - 192);
    } else if (tid >= 228 && tid < 252) {
        s.sbflag[tid - 228] = 0;
    } else if (tid >= 252 && tid < 256) {
        if (tid == 252) { s.e_nnz = 0; s.e_slog = 0; }
        else s.cbf[tid - 253] = 0;
    }
    int srcv;
    if (luma) srcv = p.src.y[(size_t)(y0 + py) * g.src_stride + x0 + px];
    else srcv = (cpl ? p.src.v : p.src.u)[(size_t)(cy * 8 + cyy) * g.srcc_stride + cx * 8 + cxx];
    __syncthreads();
    // ---- stage 1: smoothing filter, DC, mode decision (warp 3)
    if (tid < 65) s.flt[tid] = (pixel)intra_filtered(s.nb, 16, tid);
    if (tid >= 64 && tid < 96) {
        const int l = tid - 64;
        int v = l < 16 ? s.nb[1 + l] + s.nb[33 + l] : 0;
        v = warp_sum(v);
        if (l == 0) s.dc = (v + 16) >> 5;
    } else if (tid >= 96 && tid < 128) {
        const int l = tid - 96;
        int a = 1, b = 1, mpm0, mpm1, mpm2;
        if (cx > 0) a = s.left_mode[k >> 1];
        if (k >> 1) b = s.top_mode[k & 1];
        if (a == b) {
            if (a < 2) { mpm0 = 0; mpm1 = 1; mpm2 = 26; }
            else { mpm0 = a; mpm1 = 2 + ((a + 29) & 31); mpm2 = 2 + ((a - 1) & 31); }
        } else {
            mpm0 = a; mpm1 = b;
            mpm2 = (a != 0 && b != 0) ? 0 : (a != 1 && b != 1) ? 1 : 26;
        }
        unsigned key = 0xffffffffu;
#pragma unroll
        for (int m = l; m < 35; m += 32) {
            const int bits = m == mpm0 ? 2 : (m == mpm1 || m == mpm2) ? 3 : 6;
            key = min(key, ((unsigned)(s.cost[m] + ((s.lambda * bits) >> 8)) << 6) | (unsigned)m);    // ties: lowest mode
        }
        key = __reduce_min_sync(0xffffffffu, key);
        if (l == 0) s.best_mode = (int)(key & 63);
    }
    __syncthreads();
    const int mode = s.best_mode;
    int16_t *coef = p.coefs + (size_t)(cy * g.cuw + cx) * kCuCoefs;
    // ---- stage 2: prediction and residual
    int pv;
    if (luma) {
        pv = intra_sample(intra_use_filter(4, mode) ? s.flt : s.nb, 16, 4, mode, px, py, true, maxv, s.dc);
        s.a[py][px] = (int16_t)(srcv - pv);
    } else {
        const pixel *nbuf = s.cnb[cpl];
        int dcs = 8;
#pragma unroll
        for (int i = 0; i < 8; i++) dcs += nbuf[1 + i] + nbuf[17 + i];
        pv = intra_sample(nbuf, 8, 3, mode, cxx, cyy, false, maxv, dcs >> 4);
        s.ca[cpl][cyy][cxx] = (int16_t)(srcv - pv);
    }
    __syncthreads();
    // ---- stages 3-4: forward transform
    if (luma) fwd_stage<16, 18>(s.m16, &s.a[0][0], &s.b[0][0], py, px, 3 + (bd - 8));
    else fwd_stage<8, 10>(s.m8, &s.ca[cpl][0][0], &s.cb[cpl][0][0], cyy, cxx, 2 + (bd - 8));
    __syncthreads();
    if (luma) fwd_stage<16, 18>(s.m16, &s.b[0][0], &s.a[0][0], py, px, 10);
    else fwd_stage<8, 10>(s.m8, &s.cb[cpl][0][0], &s.ca[cpl][0][0], cyy, cxx, 9);
    __syncthreads();
    // ---- stage 5: quantisation, levels out, size-estimate statistics, dequantisation in place
    int lv;
    if (luma) {
        lv = quant_one(s.a[py][px], s.qy);
        coef[tid] = (int16_t)lv;
        s.a[py][px] = (int16_t)dequant_one(lv, s.qy);
        if (lv) { s.sbflag[(py >> 2) * 4 + (px >> 2)] = 1; s.cbf[0] = 1; }
    } else {
        lv = quant_one(s.ca[cpl][cyy][cxx], s.qc);
        coef[tid] = (int16_t)lv;                                  // tid - 256 + 256
        s.ca[cpl][cyy][cxx] = (int16_t)dequant_one(lv, s.qc);
        if (lv) { s.sbflag[16 + cpl * 4 + (cyy >> 2) * 2 + (cxx >> 2)] = 1; s.cbf[1 + cpl] = 1; }
    }
    {
        const int nnz = warp_sum(lv != 0), slog = warp_sum(lv ? 31 - __clz(abs(lv)) : 0);
        if ((tid & 31) == 0 && nnz) { atomicAdd(&s.e_nnz, nnz); atomicAdd(&s.e_slog, slog); }
    }
    __syncthreads();
    // ---- stages 6-7: inverse transform
    if (luma) inv_stage<16, 18>(s.m16, &s.a[0][0], &s.b[0][0], py, px, 7);
    else inv_stage<8, 10>(s.m8, &s.ca[cpl][0][0], &s.cb[cpl][0][0], cyy, cxx, 7);
    __syncthreads();
    if (luma) inv_stage<16, 18>(s.m16, &s.b[0][0], &s.a[0][0], py, px, 12 - (bd - 8));
    else inv_stage<8, 10>(s.m8, &s.cb[cpl][0][0], &s.ca[cpl][0][0], cyy, cxx, 12 - (bd - 8));
    __syncthreads();
    // ---- stage 8: reconstruction, CU record, size estimate
    if (luma) p.rec.y[(size_t)(y0 + py) * g.rec_stride + x0 + px] = (pixel)clamp0(pv + s.a[py][px], maxv);
    else (cpl ? p.rec.v : p.rec.u)[(size_t)(cy * 8 + cyy) * g.recc_stride + cx * 8 + cxx] = (pixel)clamp0(pv + s.ca[cpl][cyy][cxx], maxv);
    if (tid < 32) {
        const int nsb = __popc(__ballot_sync(0xffffffffu, tid < 24 && s.sbflag[tid < 24 ? tid : 0] != 0));
        if (tid == 0) {
            CuInfo ci;
            ci.pred_mode = 0; ci.intra_mode = (uint8_t)mode; ci.cbf = (uint8_t)((s.cbf[0] ? 1 : 0) | (s.cbf[1] ? 2 : 0) | (s.cbf[2] ? 4 : 0));
            ci.skip = 0; ci.mvx = 0; ci.mvy = 0;
            p.cus[cy * g.cuw + cx] = ci;
            atomicAdd(&p.ctl->est16, ci.cbf ? (unsigned long long)(47 * s.e_nnz + 22 * s.e_slog + 104 * nsb + 160) : 80ull);
            s.left_mode[k >> 1] = mode;
            if (!(k >> 1)) s.top_mode[k & 1] = mode;
        }
    }
    __syncthreads();
}

// Intra mode search for every CU of a key frame at once: SATD of the 35 luma predictions built from the SOURCE picture's
// neighbour samples (same availability / substitution rules as the real prediction), so that nothing here depends on the
// reconstruction.  Writes the 35 distortions per CU; the wavefront kernel adds the signalling cost (which needs the real
// most-probable modes), picks the mode and reconstructs.  One CTA walks CUs in a grid-stride loop.
struct __align__(16) IntraSearchScratch {
    pixel nb[68], flt[68];
    pixel src[16][16];
    int dc;
};
// Work list of the intra search: every CU of a key frame; in a P frame only the CUs that inter prediction serves badly
// (oracle/hevc_encode.c intra_search_all) -- they cluster, so handing them out through a list (rather than by CU index) is
// what keeps all warps busy.  One thread per CU; appends are warp-aggregated.
__global__ void __launch_bounds__(256) k_intra_list(IntraParams p)
{
    const Geom &g = p.g;
    const int cu = blockIdx.x * blockDim.x + threadIdx.x, ncu = g.cuw * g.cuh, lane = threadIdx.x & 31;
    const bool key = p.ctl->is_idr != 0;
    if (!key && !p.intra_in_p) return;
    bool cand = false;
    if (cu < ncu) {
        if (key) cand = true;
        else {
            const long long satd_sum = (long long)p.ctl->satd_sum;
            const int gate_thr = (256 * p.ctl->lambda) >> 9, s1 = p.satd1[cu];
            cand = s1 > gate_thr && ((long long)s1 * ncu > 2 * satd_sum || s1 > 8 * gate_thr || s1 > (2048 << (g.bit_depth - 8)));
        }
        p.intra_best[cu] = 0x7fffffff;
    }
    const unsigned m = __ballot_sync(0xffffffffu, cand);
    if (!m) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(&p.ctl->n_cand, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (cand) p.cand_list[base + __popc(m & ((1u << lane) - 1))] = cu;
}

// one warp per work item = (listed CU, third of the 35 modes), grid-stride, no block barriers: lanes 0-15 / 16-31 take the 16
// sub-blocks of two modes per round and reduce their SATD with shuffles; the per-CU minimum is combined with atomicMin
__global__ void __launch_bounds__(kIntraSearchThreads) k_intra_search(IntraParams p)
{
    __shared__ IntraSearchScratch scratch[kIntraSearchThreads / 32];
    const Geom &g = p.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, bd = g.bit_depth, maxv = (1 << bd) - 1;
    IntraSearchScratch &s = scratch[warp];
    if (!p.ctl->is_idr && !p.intra_in_p) return;
    const int nwarps = gridDim.x * (kIntraSearchThreads / 32);
    const int n_items = 3 * __ldcg(&p.ctl->n_cand);
    for (int item = blockIdx.x * (kIntraSearchThreads / 32) + warp; item < n_items; item += nwarps) {
        const int cu = __ldcg(p.cand_list + item / 3), part = item % 3;
        const int cx = cu % g.cuw, cy = cu / g.cuw, x0 = cx * 16, y0 = cy * 16;
        for (int i = lane; i < 65; i += 32) s.nb[i] = gather_one(p.src.y, g.src_stride, g, cx, cy, 16, i, bd);
        {
            const int row = lane >> 1, c0 = (lane & 1) * 8;
            *reinterpret_cast<uint4 *>(&s.src[row][c0]) = *reinterpret_cast<const uint4 *>(p.src.y + (size_t)(y0 + row) * g.src_stride + x0 + c0);
        }
        __syncwarp();
        int fl[3];
#pragma unroll
        for (int r = 0; r < 3; r++) fl[r] = lane + 32 * r < 65 ? intra_filtered(s.nb, 16, lane + 32 * r) : 0;
        {
            int v = lane < 16 ? s.nb[1 + lane] + s.nb[33 + lane] : 0;
            v = warp_sum(v);
            if (lane == 0) s.dc = (v + 16) >> 5;
        }
#pragma unroll
        for (int r = 0; r < 3; r++) if (lane + 32 * r < 65) s.flt[lane + 32 * r] = (pixel)fl[r];
        __syncwarp();
        const int sb = lane & 15, sx = (sb & 3) * 4, sy = (sb >> 2) * 4;
        int sv[4][4];
#pragma unroll
        for (int y = 0; y < 4; y++)
#pragma unroll
            for (int x = 0; x < 4; x++) sv[y][x] = s.src[sy + y][sx + x];
        int cbest = 0x7fffffff;
#pragma unroll 1
        for (int m0 = 12 * part; m0 < 12 * part + 12; m0 += 2) {
            const int mode = min(m0 + (lane >> 4), 34);          // the upper half idles on mode 34 twice in the last round
            const pixel *nbuf = intra_use_filter(4, mode) ? s.flt : s.nb;
            int d[4][4];
#pragma unroll
            for (int y = 0; y < 4; y++)
#pragma unroll
                for (int x = 0; x < 4; x++)
                    d[y][x] = sv[y][x] - intra_sample(nbuf, 16, 4, mode, sx + x, sy + y, true, maxv, s.dc);
            int c = hadamard4x4_abs(d) >> 1;
#pragma unroll
            for (int o = 8; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (sb == 0 && m0 + (lane >> 4) < 35) p.mode_cost[(size_t)cu * 35 + mode] = c;
            cbest = min(cbest, c);
        }
        cbest = min(cbest, __shfl_xor_sync(0xffffffffu, cbest, 16));
        if (lane == 0) atomicMin(&p.intra_best[cu], cbest);
        __syncwarp();
    }
}

// one CTA per CTU row; row r may process CTU x once row r-1 has finished CTU x+1 (top-right dependency).
// I slices: every CU.  P slices: only the CUs k_inter marked intra -- the row's CTUs that hold such a CU are listed up front and
// progress[] jumps from one listed CTU to the next (everything between is final since k_inter), so rows only wait on real
// dependencies and a frame without intra CUs costs one scan.
__global__ void __launch_bounds__(kIntraReconThreads) k_intra(IntraParams p)
{
    __shared__ IntraScratch s;
    __shared__ uint8_t has_intra[256];
    __shared__ short ctu_list[256];
    __shared__ int n_list;
    const Geom &g = p.g;
    const int r = blockIdx.x, tid = threadIdx.x;
    const bool slice_intra = p.ctl->is_idr != 0;
    if (p.second_pass && !p.ctl->redo)
        return;
    if (!slice_intra && !p.intra_in_p)
        return;
    for (int x = tid; x < g.ctuw; x += kIntraReconThreads) {
        bool any = slice_intra;
        for (int k = 0; k < 4 && !any; k++) {
            const int cx = 2 * x + (k & 1), cy = 2 * r + (k >> 1);
            any = cx < g.cuw && cy < g.cuh && p.cus[cy * g.cuw + cx].pred_mode == 0;
        }
        has_intra[x] = any;
    }
    for (int i = tid; i < 256; i += kIntraReconThreads) s.m16[i] = (int16_t)tmat_rt(16, i >> 4, i & 15);
    for (int i = tid; i < 64; i += kIntraReconThreads) s.m8[i] = (int16_t)tmat_rt(8, i >> 3, i & 7);
    if (tid == 0) { s.lambda = p.ctl->lambda; s.qy = p.ctl->qy; s.qc = p.ctl->qc; }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int x = 0; x < g.ctuw; x++)
            if (has_intra[x]) ctu_list[n++] = (short)x;
        n_list = n;
        atomicExch(p.progress + r, n ? (int)ctu_list[0] : g.ctuw);
    }
    __syncthreads();
    const int n = n_list;
    for (int li = 0; li < n; li++) {
        const int x = ctu_list[li];
        // which of the CTU's CUs are intra: final since k_inter, so the four records are fetched together and before the wait
        // (one global round trip instead of up to four dependent ones on the wavefront's critical path)
        unsigned intra_mask = 0xfu;
        if (!slice_intra) {
            uint8_t pm[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int cx = min(2 * x + (k & 1), g.cuw - 1), cy = min(2 * r + (k >> 1), g.cuh - 1);
                pm[k] = p.cus[cy * g.cuw + cx].pred_mode;
            }
            intra_mask = (pm[0] == 0 ? 1u : 0u) | (pm[1] == 0 ? 2u : 0u) | (pm[2] == 0 ? 4u : 0u) | (pm[3] == 0 ? 8u : 0u);
        }
        if (r > 0) {
            if (tid == 0) {
                const int need = min(x + 2, g.ctuw);
                while (*reinterpret_cast<volatile int *>(p.progress + r - 1) < need)
                    __nanosleep(64);
            }
            __syncthreads();
            __threadfence();
        }
        if (tid == 0 && (li == 0 || ctu_list[li - 1] != x - 1)) { s.left_mode[0] = 1; s.left_mode[1] = 1; }   // inter (or no) CUs to the left: DC
        for (int k = 0; k < 4; k++) {
            const int cx = 2 * x + (k & 1), cy = 2 * r + (k >> 1);
            if (cx >= g.cuw || cy >= g.cuh) continue;
            if ((intra_mask >> k) & 1) {
                intra_cu(p, s, cx, cy);
            } else if (tid == 0) {         // an inter CU counts as DC in the most-probable-mode derivation of its neighbours
                s.left_mode[k >> 1] = 1;
                if (!(k >> 1)) s.top_mode[k & 1] = 1;
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0)
            atomicExch(p.progress + r, li + 1 < n ? (int)ctu_list[li + 1] : g.ctuw);
    }
}

// ================================================================================================ deblocking filter
// H.265 8.7.2 restricted to this encoder's edges: CU boundaries (every 16 luma samples).  One thread per 4-line edge
// segment (plus the two chroma lines it covers when the boundary strength is 2).  dir 0: vertical edges, dir 1: horizontal
// edges; the horizontal pass is a second launch because it consumes the output of the vertical one.
__constant__ uint8_t c_tc_table[54] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4,
                                       5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24};
__constant__ uint8_t c_beta_table[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 20, 22, 24,
                                         26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64};

__device__ __forceinline__ void deblock_luma_segment(pixel *s, ptrdiff_t step, ptrdiff_t line, int beta, int tc, int maxv)
{
    int p[4][4], q[4][4];        // [line][distance from the edge]
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            p[k][i] = s[k * line - (i + 1) * step];
            q[k][i] = s[k * line + i * step];
        }
    const int dp0 = abs(p[0][2] - 2 * p[0][1] + p[0][0]), dp3 = abs(p[3][2] - 2 * p[3][1] + p[3][0]);
    const int dq0 = abs(q[0][2] - 2 * q[0][1] + q[0][0]), dq3 = abs(q[3][2] - 2 * q[3][1] + q[3][0]);
    const int dpq0 = dp0 + dq0, dpq3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3;
    if (dpq0 + dpq3 >= beta) return;
    const bool strong0 = 2 * dpq0 < (beta >> 2) && abs(p[0][3] - p[0][0]) + abs(q[0][0] - q[0][3]) < (beta >> 3) &&
                         abs(p[0][0] - q[0][0]) < ((5 * tc + 1) >> 1);
    const bool strong3 = 2 * dpq3 < (beta >> 2) && abs(p[3][3] - p[3][0]) + abs(q[3][0] - q[3][3]) < (beta >> 3) &&
                         abs(p[3][0] - q[3][0]) < ((5 * tc + 1) >> 1);
    const bool strong = strong0 && strong3;
    const bool dep = dp < ((beta + (beta >> 1)) >> 3), deq = dq < ((beta + (beta >> 1)) >> 3);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int p0 = p[k][0], p1 = p[k][1], p2 = p[k][2], p3 = p[k][3], q0 = q[k][0], q1 = q[k][1], q2 = q[k][2], q3 = q[k][3];
        pixel *l = s + k * line;
        if (strong) {
            l[-1 * step] = (pixel)clampd((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3, p0 - 2 * tc, p0 + 2 * tc);
            l[-2 * step] = (pixel)clampd((p2 + p1 + p0 + q0 + 2) >> 2, p1 - 2 * tc, p1 + 2 * tc);
            l[-3 * step] = (pixel)clampd((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3, p2 - 2 * tc, p2 + 2 * tc);
            l[0] = (pixel)clampd((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3, q0 - 2 * tc, q0 + 2 * tc);
            l[step] = (pixel)clampd((p0 + q0 + q1 + q2 + 2) >> 2, q1 - 2 * tc, q1 + 2 * tc);
            l[2 * step] = (pixel)clampd((p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3, q2 - 2 * tc, q2 + 2 * tc);
        } else {
            int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
            if (abs(delta) < 10 * tc) {
                delta = clampd(delta, -tc, tc);
                l[-1 * step] = (pixel)clamp0(p0 + delta, maxv);
                l[0] = (pixel)clamp0(q0 - delta, maxv);
                if (dep) l[-2 * step] = (pixel)clamp0(p1 + clampd((((p2 + p0 + 1) >> 1) - p1 + delta) >> 1, -(tc >> 1), tc >> 1), maxv);
                if (deq) l[step] = (pixel)clamp0(q1 + clampd((((q2 + q0 + 1) >> 1) - q1 - delta) >> 1, -(tc >> 1), tc >> 1), maxv);
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_deblock(DeblockParams p)
{
    const Geom &g = p.g;
    const int total = g.cuw * g.cuh * 4;
    const int qp = p.ctl->qp, bd = g.bit_depth, maxv = (1 << bd) - 1;
    const int beta = c_beta_table[min(max(qp, 0), 51)] << (bd - 8);
    for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < total; item += gridDim.x * blockDim.x) {
        const int cu = item >> 2, seg = item & 3, cx = cu % g.cuw, cy = cu / g.cuw;
        if ((p.dir == 0 && cx == 0) || (p.dir == 1 && cy == 0)) continue;
        const CuInfo q = p.cus[cu], pp = p.cus[p.dir == 0 ? cu - 1 : cu - g.cuw];
        int bs = 0;
        if (pp.pred_mode == 0 || q.pred_mode == 0) bs = 2;
        else if (((pp.cbf | q.cbf) & 1) || abs(pp.mvx - q.mvx) >= 4 || abs(pp.mvy - q.mvy) >= 4) bs = 1;
        if (!bs) continue;
        const int tc = c_tc_table[min(max(qp + 2 * (bs - 1), 0), 53)] << (bd - 8);
        pixel *y = p.rec.y + (size_t)cy * 16 * g.rec_stride + cx * 16;
        if (p.dir == 0) deblock_luma_segment(y + (size_t)seg * 4 * g.rec_stride, 1, g.rec_stride, beta, tc, maxv);
        else deblock_luma_segment(y + seg * 4, g.rec_stride, 1, beta, tc, maxv);
        if (bs == 2) {
            const int tcc = c_tc_table[min(max(chroma_qp(qp) + 2, 0), 53)] << (bd - 8);
#pragma unroll
            for (int c = 0; c < 2; c++) {
                pixel *u = (c ? p.rec.v : p.rec.u) + (size_t)cy * 8 * g.recc_stride + cx * 8;
                const ptrdiff_t step = p.dir == 0 ? 1 : g.recc_stride, line = p.dir == 0 ? g.recc_stride : 1;
                for (int k = 2 * seg; k < 2 * seg + 2; k++) {
                    pixel *l = u + k * line;
                    const int p0 = l[-step], p1 = l[-2 * step], q0 = l[0], q1 = l[step];
                    const int delta = clampd((((q0 - p0) << 2) + p1 - q1 + 4) >> 3, -tcc, tcc);
                    l[-step] = (pixel)clamp0(p0 + delta, maxv);
                    l[0] = (pixel)clamp0(q0 - delta, maxv);
                }
            }
        }
    }
}

cudaError_t upload_inter_constants(cudaStream_t st)
{
    static const int8_t taps[4][8] = {{0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1}};
    int pack[4][5];
    auto b = [](int v) { return (unsigned)(uint8_t)(int8_t)v; };
    for (int f = 0; f < 4; f++) {
        const int8_t *t = taps[f];
        pack[f][0] = (int)(b(t[0]) | (b(t[1]) << 8) | (b(t[2]) << 16) | (b(t[3]) << 24));
        pack[f][1] = (int)(b(t[4]) | (b(t[5]) << 8) | (b(t[6]) << 16) | (b(t[7]) << 24));
        pack[f][2] = (int)((b(t[0]) << 8) | (b(t[1]) << 16) | (b(t[2]) << 24));
        pack[f][3] = (int)(b(t[3]) | (b(t[4]) << 8) | (b(t[5]) << 16) | (b(t[6]) << 24));
        pack[f][4] = (int)b(t[7]);
    }
    return cudaMemcpyToSymbolAsync(c_luma_pack, pack, sizeof(pack), 0, cudaMemcpyHostToDevice, st);
}

// ================================================================================================ rate control steps
// single-thread kernels: close the books of the frame that just finished and choose the QP of the next one
__global__ void k_rc_step(RcState *rc, FrameCtl *done, FrameCtl *next, int force_idr, const SceneStat *scene, long long ds_samples)
{
    if (threadIdx.x || blockIdx.x) return;
    RcState s = *rc;
    if (done) rc_update(s, done->is_idr, done->qp, (long long)done->est16, done->poc);
    if (next) {
        // frame type (oracle/hevc_encode.c orc_enc_frame): forced / first frame / keyint reached / scene cut past min-keyint
        const int cut = s.started && s.scenecut && scene && scene_cut(*scene, ds_samples) && s.poc + 1 >= s.min_keyint;
        const int idr = force_idr || !s.started || s.poc + 1 >= s.keyint || cut;
        s.poc = idr ? 0 : s.poc + 1;
        s.started = 1;
        FrameCtl c;
        ctl_set_qp(c, rc_pick_qp(s, idr, s.poc), idr, s.bit_depth);
        c.redo = 0; c.poc = s.poc; c.scene_cut = cut;
        *next = c;
    }
    *rc = s;
}

// first key frame of a stream under rate control: if the first try overshoots its budget, ask for a second pass
__global__ void k_rc_redo(RcState *rc, FrameCtl *ctl)
{
    if (threadIdx.x || blockIdx.x) return;
    const RcState s = *rc;
    FrameCtl c = *ctl;
    c.redo = 0;
    if (s.rate_control && !s.have[1]) {
        const long long budget = rc_budget(s, 1), est = (long long)c.est16;
        int qp2 = c.qp + rc_step(est, budget);
        qp2 = qp2 < c.qp ? c.qp : qp2 > 51 ? 51 : qp2;
        if (est > budget && qp2 != c.qp) {
            ctl_set_qp(c, qp2, 1, s.bit_depth);
            c.redo = 1;
        }
    }
    *ctl = c;
}

}  // namespace hb
