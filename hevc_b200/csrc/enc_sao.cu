// Sample adaptive offset (H.265 8.7.3) for sm_100a: per-CTU statistics + parameter decision, and application.
// One CTA per CTU in both kernels; every CTU decides on its own statistics, so a frame is one launch each.
// Bit-exact against oracle/hevc_encode.c (sao_collect / sao_offset / sao_decide_group / sao_apply_ctb).
#include "enc_kernels.cuh"

namespace hb {

namespace {

constexpr int kSaoThreads = 256;

__device__ __forceinline__ int sgn(int v) { return (v > 0) - (v < 0); }
// edge category of 8.7.3.2: 1 valley, 2 / 3 edges, 4 peak, 0 none
__device__ __forceinline__ int sao_category(int r, int a, int b)
{
    const int e = 2 + sgn(r - a) + sgn(r - b);
    return (0x43021 >> (4 * e)) & 7;          // {1, 2, 0, 3, 4}
}

struct SaoStats {
    int cnt[4][4], sum[4][4];                 // edge classes x categories 1..4
    int bcnt[32], bsum[32];                   // bands
};

struct SaoScratch {
    pixel tile[34][36];                       // component tile with a one-sample halo (luma 34x34, chroma 18x18 in the corner)
    SaoStats st[3];
    long long ocost[3][48];                   // best cost per (component, edge class x category | band)
    int8_t ooff[3][48];
    int copied;
};

// best offset in [lo, hi] (oracle sao_offset): minimises (cnt o^2 - 2 o sum) * 65536 + lam * bits, from the rounded mean towards zero
__device__ int sao_offset(long long cnt, long long sum, int lo, int hi, long long lam, int cmax, int sign_bit, long long &cost_out)
{
    long long best = lam;
    int bo = 0;
    if (cnt) {
        int o = (int)(sum >= 0 ? (sum + cnt / 2) / cnt : -((-sum + cnt / 2) / cnt));
        o = min(max(o, lo), hi);
        for (int t = o; t != 0; t += t > 0 ? -1 : 1) {
            const int a = abs(t);
            const long long cost = (cnt * t * t - 2 * t * sum) * 65536 + lam * ((a < cmax ? a + 1 : cmax) + sign_bit);
            if (cost < best) { best = cost; bo = t; }
        }
    }
    cost_out = best;
    return bo;
}

// stage the N x N block at (x0, y0) of `plane` with a one-sample halo; samples outside the picture are never read back
template <int N>
__device__ __forceinline__ void stage_tile(SaoScratch &s, const pixel *plane, int stride, int x0, int y0, int w, int h, int tid)
{
    for (int i = tid; i < (N + 2) * (N + 2); i += kSaoThreads) {
        const int ty = i / (N + 2), tx = i - ty * (N + 2);
        const int x = min(max(x0 + tx - 1, 0), w - 1), y = min(max(y0 + ty - 1, 0), h - 1);
        s.tile[ty][tx] = plane[(ptrdiff_t)y * stride + x];
    }
}

// statistics of one component.  Every thread takes samples i = tid, tid + 256, ... of the N x N block and classifies each for
// all four edge classes from the halo tile (the sample and its eight neighbours are read once).  Count and error sum of a
// category share one accumulator, (count << 22) + sum: a thread sees at most 4 samples, a warp 128 (|sum| < 2^18), so the signed
// sum never reaches bit 21 and the pair comes apart again after the warp reduction -- 16 accumulators instead of 32, which keeps
// the kernel at 6 resident CTAs per SM.  Band statistics go through warp-aggregated shared-memory atomics.
// EDGE = false: the block and its halo lie inside the picture, every picture-border test folds away.
template <int N, bool EDGE>
__device__ __forceinline__ void collect(SaoScratch &s, SaoStats &st, const pixel *src, int src_stride, int x0, int y0, int w, int h,
                                        int bshift, int tid)
{
    constexpr int logn = N == 32 ? 5 : 4;
    int acc[4][4];
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[k][c] = 0;
#pragma unroll 1
    for (int i = tid; i < N * N; i += kSaoThreads) {
        const int ly = i >> logn, lx = i & (N - 1), x = x0 + lx, y = y0 + ly;
        const bool inside = !EDGE || (x < w && y < h);
        const int r = s.tile[ly + 1][lx + 1];
        const int d = inside ? (int)__ldg(src + (size_t)y * src_stride + x) - r : 0;
        {   // band statistics: lanes with the same band pool their contribution
            const int band = inside ? r >> bshift : 32 + (tid & 31);      // lanes outside the picture form singleton groups and do nothing
            const unsigned grp = __match_any_sync(0xffffffffu, band);
            const int gs = __reduce_add_sync(grp, d), gc = __popc(grp);
            if (inside && (int)(__ffs(grp) - 1) == (tid & 31)) { atomicAdd(&st.bcnt[band], gc); atomicAdd(&st.bsum[band], gs); }
        }
        const bool l = !EDGE || x > 0, rr = !EDGE || x + 1 < w, u = !EDGE || y > 0, dn = !EDGE || y + 1 < h;
        const bool okh = inside && l && rr, okv = inside && u && dn, okd = okh && okv;
        const int cat[4] = {okh ? sao_category(r, s.tile[ly + 1][lx], s.tile[ly + 1][lx + 2]) : 0,
                            okv ? sao_category(r, s.tile[ly][lx + 1], s.tile[ly + 2][lx + 1]) : 0,
                            okd ? sao_category(r, s.tile[ly][lx], s.tile[ly + 2][lx + 2]) : 0,
                            okd ? sao_category(r, s.tile[ly][lx + 2], s.tile[ly + 2][lx]) : 0};
        const int contrib = (1 << 22) + d;
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[k][c] += cat[k] == c + 1 ? contrib : 0;
    }
    // warp totals: lane k * 4 + c keeps accumulator (k, c), so the unpacking and the shared-memory atomics run once, 16 lanes wide
    int mine = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int t = __reduce_add_sync(0xffffffffu, acc[k][c]);
            if ((tid & 31) == k * 4 + c) mine = t;
        }
    if ((tid & 31) < 16) {
        const int ts = (int)((unsigned)mine << 10) >> 10, tc = (mine - ts) >> 22;
        if (tc) { atomicAdd(&st.cnt[0][tid & 15], tc); atomicAdd(&st.sum[0][tid & 15], ts); }
    }
}

// final choice of one component group from the per-entry best costs (oracle sao_decide_group); comps = first component, count
__device__ void decide_group(const SaoScratch &s, int c0, int ncomp, long long lam, SaoCtu &out, int g)
{
    long long best = lam;
    int type = 0, eo_class = 0, band[2] = {0, 0};
    for (int k = 0; k < 4; k++) {
        long long cost = lam * 4;
        for (int c = 0; c < ncomp; c++)
            for (int cat = 0; cat < 4; cat++) cost += s.ocost[c0 + c][k * 4 + cat];
        if (cost < best) { best = cost; type = 2; eo_class = k; }
    }
    {
        long long cost = lam * 2;
        int bp[2] = {0, 0};
        for (int c = 0; c < ncomp; c++) {
            const long long *bc = &s.ocost[c0 + c][16];
            long long wbest = 0;
            for (int s0 = 0; s0 <= 28; s0++) {
                const long long wv = bc[s0] + bc[s0 + 1] + bc[s0 + 2] + bc[s0 + 3];
                if (s0 == 0 || wv < wbest) { wbest = wv; bp[c] = s0; }
            }
            cost += wbest + lam * 5;
        }
        if (cost < best) { best = cost; type = 1; band[0] = bp[0]; band[1] = bp[1]; }
    }
    out.type[g] = (uint8_t)type;
    out.eo_class[g] = (uint8_t)(type == 2 ? eo_class : 0);
    for (int c = 0; c < ncomp; c++) {
        out.band[g + c] = (uint8_t)(type == 1 ? band[c] : 0);
        for (int i = 0; i < 4; i++)
            out.offset[g + c][i] = type == 2 ? s.ooff[c0 + c][eo_class * 4 + i] : type == 1 ? s.ooff[c0 + c][16 + band[c] + i] : 0;
    }
}

}  // namespace

__global__ void __launch_bounds__(kSaoThreads, 6) k_sao_decide(SaoParams p)
{
    __shared__ SaoScratch s;
    const Geom &g = p.g;
    const int tid = threadIdx.x, rx = blockIdx.x % g.ctuw, ry = blockIdx.x / g.ctuw;
    const int bd = g.bit_depth, cmax = (1 << (min(bd, 10) - 5)) - 1, bshift = bd - 5;
    if (tid == 0) {
        // a CTU whose CUs are all inter without residual is a copy of reference samples that already went through SAO: left off
        int copied = 1;
        for (int k = 0; k < 4; k++) {
            const int cx = 2 * rx + (k & 1), cy = 2 * ry + (k >> 1);
            if (cx < g.cuw && cy < g.cuh) {
                const CuInfo c = p.cus[cy * g.cuw + cx];
                if (c.pred_mode == 0 || c.cbf) copied = 0;
            }
        }
        s.copied = copied;
    }
    for (int i = tid; i < (int)(sizeof(s.st) / sizeof(int)); i += kSaoThreads) reinterpret_cast<int *>(s.st)[i] = 0;
    __syncthreads();
    if (s.copied) {
        if (tid < 8) reinterpret_cast<uint32_t *>(&p.sao[blockIdx.x])[tid] = 0;
        return;
    }
    for (int comp = 0; comp < 3; comp++) {
        const int N = comp ? 16 : 32, w = comp ? g.wc >> 1 : g.wc, h = comp ? g.hc >> 1 : g.hc;
        const pixel *pre = comp == 0 ? p.pre.y : comp == 1 ? p.pre.u : p.pre.v, *src = comp == 0 ? p.src.y : comp == 1 ? p.src.u : p.src.v;
        if (comp == 0) stage_tile<32>(s, pre, g.rec_stride, rx * N, ry * N, w, h, tid);
        else stage_tile<16>(s, pre, g.recc_stride, rx * N, ry * N, w, h, tid);
        __syncthreads();
        const bool edge = rx == 0 || ry == 0 || (rx + 1) * N >= w || (ry + 1) * N >= h;
        if (comp == 0) {
            if (edge) collect<32, true>(s, s.st[comp], src, g.src_stride, rx * N, ry * N, w, h, bshift, tid);
            else collect<32, false>(s, s.st[comp], src, g.src_stride, rx * N, ry * N, w, h, bshift, tid);
        } else {
            if (edge) collect<16, true>(s, s.st[comp], src, g.srcc_stride, rx * N, ry * N, w, h, bshift, tid);
            else collect<16, false>(s, s.st[comp], src, g.srcc_stride, rx * N, ry * N, w, h, bshift, tid);
        }
        __syncthreads();
    }
    const int qp = p.ctl->qp;
    const long long ly = (long long)(lambda_q8(qp) << (bd - 8)), lc = (long long)(lambda_q8(chroma_qp(qp)) << (bd - 8));
    const long long lam_y = 3 * ly * ly, lam_c = 3 * lc * lc;
    if (tid < 144) {          // one (component, entry) per thread: entries 0..15 edge class x category, 16..47 bands
        const int comp = tid / 48, e = tid % 48;
        const long long lam = comp ? lam_c : lam_y;
        const SaoStats &st = s.st[comp];
        long long cost;
        int o;
        if (e < 16) {
            const int cat = e & 3;
            o = sao_offset(st.cnt[e >> 2][cat], st.sum[e >> 2][cat], cat < 2 ? 0 : -cmax, cat < 2 ? cmax : 0, lam, cmax, 0, cost);
        } else {
            o = sao_offset(st.bcnt[e - 16], st.bsum[e - 16], -cmax, cmax, lam, cmax, 1, cost);
        }
        s.ocost[comp][e] = cost;
        s.ooff[comp][e] = (int8_t)o;
    }
    __syncthreads();
    __shared__ SaoCtu result;
    if (tid < 8) reinterpret_cast<uint32_t *>(&result)[tid] = 0;
    __syncthreads();
    if (tid == 0) decide_group(s, 0, 1, lam_y, result, 0);
    if (tid == 32) decide_group(s, 1, 2, lam_c, result, 1);
    __syncthreads();
    if (tid < 8) reinterpret_cast<uint32_t *>(&p.sao[blockIdx.x])[tid] = reinterpret_cast<const uint32_t *>(&result)[tid];
}

// pre (deblocked) -> out: every CTU is written (copied when its type is 0), so `out` is complete after this launch
__global__ void __launch_bounds__(kSaoThreads) k_sao_apply(SaoParams p)
{
    __shared__ pixel tile[34][36];
    __shared__ SaoCtu sc;
    const Geom &g = p.g;
    const int tid = threadIdx.x, rx = blockIdx.x % g.ctuw, ry = blockIdx.x / g.ctuw;
    const int bd = g.bit_depth, maxv = (1 << bd) - 1, bshift = bd - 5;
    if (tid < 8) reinterpret_cast<uint32_t *>(&sc)[tid] = reinterpret_cast<const uint32_t *>(&p.sao[blockIdx.x])[tid];
    __syncthreads();
    for (int comp = 0; comp < 3; comp++) {
        const int N = comp ? 16 : 32, w = comp ? g.wc >> 1 : g.wc, h = comp ? g.hc >> 1 : g.hc, stride = comp ? g.recc_stride : g.rec_stride;
        const pixel *pre = comp == 0 ? p.pre.y : comp == 1 ? p.pre.u : p.pre.v;
        pixel *out = comp == 0 ? p.out.y : comp == 1 ? p.out.u : p.out.v;
        const int x0 = rx * N, y0 = ry * N, lq = comp ? 2 : 3;      // log2(N / 4)
        const int gi = comp ? 1 : 0, type = sc.type[gi], cls = sc.eo_class[gi], band = sc.band[comp];
        if (type != 2) {
            // off or band offset: no neighbours involved -- straight from global memory, four samples per work item
            for (int i = tid; i < N * N / 4; i += kSaoThreads) {
                const int ly = i >> lq, lx = (i & (N / 4 - 1)) * 4, y = y0 + ly;
                if (y >= h || x0 + lx >= w) continue;
                const ptrdiff_t o = (ptrdiff_t)y * stride + x0 + lx;
                uint2 v = *reinterpret_cast<const uint2 *>(pre + o);
                if (type == 1) {
                    uint32_t q[4] = {v.x & 0xffff, v.x >> 16, v.y & 0xffff, v.y >> 16};
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int b = (((int)q[k] >> bshift) - band) & 31;
                        if (b < 4) q[k] = (uint32_t)min(max((int)q[k] + sc.offset[comp][b], 0), maxv);
                    }
                    v = make_uint2(q[0] | (q[1] << 16), q[2] | (q[3] << 16));
                }
                *reinterpret_cast<uint2 *>(out + o) = v;
            }
            continue;
        }
        __syncthreads();
        for (int i = tid; i < (N + 2) * (N + 2); i += kSaoThreads) {
            const int ty = i / (N + 2), tx = i - ty * (N + 2);
            const int x = min(max(x0 + tx - 1, 0), w - 1), y = min(max(y0 + ty - 1, 0), h - 1);
            tile[ty][tx] = pre[(ptrdiff_t)y * stride + x];
        }
        __syncthreads();
        int dxa = 0, dya = 0;
        if (cls == 0) dxa = -1; else if (cls == 1) dya = -1; else if (cls == 2) { dxa = -1; dya = -1; } else { dxa = 1; dya = -1; }
        // four consecutive samples per work item: 64-bit stores
        for (int i = tid; i < N * N / 4; i += kSaoThreads) {
            const int ly = i >> lq, lx = (i & (N / 4 - 1)) * 4, y = y0 + ly;
            if (y >= h || x0 + lx >= w) continue;
            uint32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int x = x0 + lx + k, r = tile[ly + 1][lx + k + 1];
                int o = r;
                {
                    const int xa = x + dxa, ya = y + dya, xb = x - dxa, yb = y - dya;
                    if (xa >= 0 && xb >= 0 && ya >= 0 && yb >= 0 && xa < w && xb < w && ya < h && yb < h) {
                        const int cat = sao_category(r, tile[ly + 1 + dya][lx + k + 1 + dxa], tile[ly + 1 - dya][lx + k + 1 - dxa]);
                        if (cat) o = min(max(r + sc.offset[comp][cat - 1], 0), maxv);
                    }
                }
                v[k] = (uint32_t)o;
            }
            *reinterpret_cast<uint2 *>(out + (ptrdiff_t)y * stride + x0 + lx) = make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
        }
    }
}

}  // namespace hb
