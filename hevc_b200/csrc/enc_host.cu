// Host driver of the stream encoder: batch scheduling of the kernels, parameter sets / SEI / slice headers
// (Rec. ITU-T H.265 7.3, Annex D, Annex E), emulation prevention and access-unit assembly.
#include <atomic>
#include <memory>

#include "common.cuh"
#include "enc_kernels.cuh"

using namespace hb;

namespace {

// ------------------------------------------------------------------------------------------------ bit writer
class BitWriter {
public:
    void put(uint32_t v, int n)
    {
        for (int i = n - 1; i >= 0; i--) {
            acc_ = (uint8_t)((acc_ << 1) | ((v >> i) & 1));
            if (++fill_ == 8) { bytes_.push_back(acc_); acc_ = 0; fill_ = 0; }
        }
    }
    void flag(bool b) { put(b ? 1 : 0, 1); }
    void ue(uint32_t v)
    {
        const uint64_t x = (uint64_t)v + 1;
        int len = 0;
        while ((x >> (len + 1)) != 0) len++;
        put(0, len);
        for (int i = len; i >= 0; i--) put((uint32_t)((x >> i) & 1), 1);
    }
    void se(int v) { ue(v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * (long long)v)); }
    void trailing()          // rbsp_trailing_bits() / byte_alignment()
    {
        put(1, 1);
        while (fill_) put(0, 1);
    }
    bool aligned() const { return fill_ == 0; }
    const std::vector<uint8_t> &bytes() const { return bytes_; }

private:
    std::vector<uint8_t> bytes_;
    uint8_t acc_ = 0;
    int fill_ = 0;
};

// emulation prevention (H.265 7.4.2): 0x03 before any byte <= 3 that follows two zero bytes.  Zero-free stretches (almost all of
// a CABAC payload) are found with memchr and copied in bulk.
void append_escaped(std::vector<uint8_t> &out, const uint8_t *in, size_t n)
{
    size_t start = 0, i = 0;
    int zeros = 0;
    while (i < n) {
        if (zeros == 0) {
            const uint8_t *z = static_cast<const uint8_t *>(memchr(in + i, 0, n - i));
            if (!z) break;
            i = (size_t)(z - in);
        }
        if (zeros >= 2 && in[i] <= 3) {
            out.insert(out.end(), in + start, in + i);
            out.push_back(3);
            start = i;
            zeros = 0;
        }
        zeros = in[i] == 0 ? zeros + 1 : 0;
        i++;
    }
    out.insert(out.end(), in + start, in + n);
}

void append_nal(std::vector<uint8_t> &out, int type, const std::vector<uint8_t> &rbsp, bool long_start)
{
    if (long_start) out.push_back(0);
    out.push_back(0); out.push_back(0); out.push_back(1);
    out.push_back((uint8_t)(type << 1));
    out.push_back(1);
    append_escaped(out, rbsp.data(), rbsp.size());
}

enum { NAL_TRAIL_R = 1, NAL_IDR_W_RADL = 19, NAL_VPS = 32, NAL_SPS = 33, NAL_PPS = 34, NAL_AUD = 35, NAL_SEI_PREFIX = 39, NAL_SEI_SUFFIX = 40 };

void write_ptl(BitWriter &b, const hb_enc_params &p)
{
    b.put(0, 2); b.put(p.tier, 1); b.put(p.profile_idc, 5);
    for (int j = 0; j < 32; j++) b.flag(j == p.profile_idc || (p.profile_idc == 1 && j == 2));
    b.flag(true); b.flag(false); b.flag(false); b.flag(true);      // progressive, !interlaced, !non-packed, frame-only
    b.put(0, 32); b.put(0, 11); b.put(0, 1);
    b.put(p.level_idc, 8);
}

std::vector<uint8_t> make_vps(const hb_enc_params &p)
{
    BitWriter b;
    b.put(0, 4); b.put(3, 2); b.put(0, 6); b.put(0, 3); b.flag(true); b.put(0xffff, 16);
    write_ptl(b, p);
    b.flag(true); b.ue(1); b.ue(0); b.ue(0);
    b.put(0, 6); b.ue(0); b.flag(false); b.flag(false);
    b.trailing();
    return b.bytes();
}

std::vector<uint8_t> make_sps(const hb_enc_params &p, int wc, int hc)
{
    BitWriter b;
    b.put(0, 4); b.put(0, 3); b.flag(true);
    write_ptl(b, p);
    b.ue(0); b.ue(1); b.ue(wc); b.ue(hc);
    if (wc != p.width || hc != p.height) {
        b.flag(true); b.ue(0); b.ue((wc - p.width) / 2); b.ue(0); b.ue((hc - p.height) / 2);
    } else {
        b.flag(false);
    }
    b.ue(p.bit_depth - 8); b.ue(p.bit_depth - 8);
    b.ue(4);                                         // POC lsb: 8 bits
    b.flag(true); b.ue(1); b.ue(0); b.ue(0);         // DPB 2 pictures, no reordering
    b.ue(0); b.ue(2);                                // CB 8..32
    b.ue(0); b.ue(3);                                // TB 4..32
    b.ue(0); b.ue(0);                                // transform hierarchy depth inter / intra
    b.flag(false); b.flag(false); b.flag(p.sao != 0); b.flag(false);   // scaling lists, AMP, SAO, PCM
    b.ue(1); b.ue(1); b.ue(0); b.ue(0); b.flag(true);             // one RPS: previous picture
    b.flag(false); b.flag(false); b.flag(false);                  // long-term, temporal MVP, strong smoothing
    b.flag(true);                                                 // VUI
    b.flag(true); b.put(1, 8);                                    //   square samples
    b.flag(false);
    b.flag(true); b.put(5, 3); b.flag(p.full_range != 0); b.flag(p.vui_colour != 0);
    if (p.vui_colour) { b.put(p.colour_primaries, 8); b.put(p.transfer_characteristics, 8); b.put(p.matrix_coeffs, 8); }
    if (p.chroma_loc >= 0) { b.flag(true); b.ue(p.chroma_loc); b.ue(p.chroma_loc); }
    else b.flag(false);
    b.flag(false); b.flag(false); b.flag(false); b.flag(false);
    b.flag(true); b.put((uint32_t)p.fps_den, 32); b.put((uint32_t)p.fps_num, 32); b.flag(false);
    b.flag(p.hrd != 0);
    if (p.hrd) {
        b.flag(true); b.flag(false); b.flag(false);
        b.put(0, 4); b.put(0, 4);                                 // 64 bit/s and 16 bit units
        b.put(23, 5); b.put(23, 5); b.put(23, 5);
        b.flag(false); b.flag(false); b.flag(false); b.ue(0);
        b.ue((uint32_t)((long long)p.vbv_maxrate_kbps * 1000 / 64 - 1));
        b.ue((uint32_t)((long long)p.vbv_bufsize_kbit * 1000 / 16 - 1));
        b.flag(false);
    }
    b.flag(false);                                                // bitstream restriction
    b.flag(false);                                                // extension
    b.trailing();
    return b.bytes();
}

std::vector<uint8_t> make_pps(const hb_enc_params &p)
{
    BitWriter b;
    b.ue(0); b.ue(0);
    b.flag(false); b.flag(false); b.put(0, 3); b.flag(false); b.flag(false);
    b.ue(0); b.ue(0); b.se(0);
    b.flag(false); b.flag(false); b.flag(false);                  // constrained intra, transform skip, cu_qp_delta
    b.se(0); b.se(0); b.flag(false);
    b.flag(false); b.flag(false); b.flag(false); b.flag(false);   // weighted x2, transquant bypass, tiles
    b.flag(true);                                                 // entropy_coding_sync_enabled
    b.flag(false);                                                // loop filter across slices
    b.flag(true); b.flag(false); b.flag(!p.deblock);              // deblocking control: no override, enabled per params
    if (p.deblock) { b.se(0); b.se(0); }                          //   beta / tc offsets
    b.flag(false); b.flag(false); b.ue(0); b.flag(false); b.flag(false);
    b.trailing();
    return b.bytes();
}

std::vector<uint8_t> make_sei(int type, const std::vector<uint8_t> &payload)
{
    BitWriter b;
    int t = type;
    size_t s = payload.size();
    while (t >= 255) { b.put(255, 8); t -= 255; }
    b.put(t, 8);
    while (s >= 255) { b.put(255, 8); s -= 255; }
    b.put((uint32_t)s, 8);
    for (uint8_t v : payload) b.put(v, 8);
    b.trailing();
    return b.bytes();
}

// RFC 1321
void md5(const uint8_t *data, size_t n, uint8_t out[16])
{
    static const uint32_t K[64] = {
        0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1,
        0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453,
        0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942,
        0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05,
        0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d,
        0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
    static const int S[4][4] = {{7, 12, 17, 22}, {5, 9, 14, 20}, {4, 11, 16, 23}, {6, 10, 15, 21}};
    uint32_t h[4] = {0x67452301, 0xefcdab89, 0x98badcfe, 0x10325476};
    std::vector<uint8_t> msg(data, data + n);
    msg.push_back(0x80);
    while (msg.size() % 64 != 56) msg.push_back(0);
    for (int i = 0; i < 8; i++) msg.push_back((uint8_t)(((uint64_t)n * 8) >> (8 * i)));
    for (size_t off = 0; off < msg.size(); off += 64) {
        uint32_t M[16], a = h[0], b = h[1], c = h[2], d = h[3];
        for (int i = 0; i < 16; i++) memcpy(&M[i], &msg[off + 4 * i], 4);
        for (int i = 0; i < 64; i++) {
            uint32_t f;
            int g;
            switch (i >> 4) {
            case 0: f = (b & c) | (~b & d); g = i; break;
            case 1: f = (d & b) | (~d & c); g = (5 * i + 1) & 15; break;
            case 2: f = b ^ c ^ d; g = (3 * i + 5) & 15; break;
            default: f = c ^ (b | ~d); g = (7 * i) & 15; break;
            }
            const uint32_t x = a + f + K[i] + M[g];
            const int r = S[i >> 4][i & 3];
            a = d; d = c; c = b;
            b = b + ((x << r) | (x >> (32 - r)));
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d;
    }
    memcpy(out, h, 16);
}

constexpr int kChunk = 8;            // frames per upload / entropy chunk
constexpr int kMaxChunks = 1024 / kChunk + 1;
constexpr int kEntropyStreams = 8;
constexpr int kGroupFrames = 32;     // frames per drain group (a multiple of kChunk)
constexpr int kMaxGroups = 1024 / kGroupFrames + 1;

struct FrameSlot {
    Planes src{};
    CuInfo *cus = nullptr;
    CuSyntax *syn = nullptr;
    int16_t *coefs = nullptr;
    uint8_t *rows = nullptr;
    uint32_t *row_len = nullptr;
    uint8_t *ctx_save = nullptr;      // WPP hand-off area of the entropy stage
    int *row_ready = nullptr;
    SaoCtu *sao = nullptr;            // [ctus] SAO parameters (SAO on)
    Planes keep{};            // verification copy of the reconstruction (unpadded strides = rec strides)
};

}  // namespace

// Everything one batch owns between being enqueued and being drained.  Two sets alternate, so that the frame chain of the
// next batch (and its uploads) run while the CABAC tail of the previous one finishes and the host assembles its access units.
struct BatchSet {
    std::vector<FrameSlot> slot;      // src planes are shared between the sets (the frame chain is serial)
    uint8_t *staging = nullptr;       // raw input frames
    int *overflow = nullptr, *row_ready_all = nullptr;
    FrameCtl *ctl_dev = nullptr, *ctl_host = nullptr;
    EntropyFrame *eframes_dev = nullptr;
    unsigned long long *trace_dev = nullptr;   // debug: per-row CABAC timestamps, only with HB_ENTROPY_TRACE=<file>
    uint32_t *offsets_dev = nullptr;
    uint8_t *packed_dev = nullptr;
    // pinned host
    uint32_t *offsets_host = nullptr;
    uint8_t *packed_host = nullptr;
    int *overflow_host = nullptr;
    cudaEvent_t ev[4] = {};
    cudaEvent_t ev_misc[2] = {};
    cudaEvent_t ev_grp[kMaxGroups][kEntropyStreams] = {}, ev_off[kMaxGroups] = {};
    std::vector<cudaEvent_t> ev_chunk;   // [0, kMaxChunks): upload done, [kMaxChunks, 2 kMaxChunks): chain done
    std::vector<cudaEvent_t> kev;     // 2 per frame of a batch + 6 per batch: per-kernel-class timing
    std::vector<cudaEvent_t> kev_me;  // after the motion-search kernel of each P frame
    // host bookkeeping of the batch in flight
    bool pending = false;
    int n = 0;
    bool forced_first = false;
    long long first_frame_no = 0;
    std::vector<int> is_idr, qps, pocs;
};

struct hb_encoder {
    hb_ctx *ctx = nullptr;
    hb_enc_params prm{};
    Geom g{};
    int max_batch = 0;
    uint32_t row_cap = 0;
    // device memory
    std::vector<void *> dev;          // everything to free
    BatchSet set[2];
    int next_set = 0;                 // set the next batch is enqueued into
    BatchSet *last_drained = nullptr; // what read_recon / read_decisions look at
    pixel *rec_base[3][3] = {};
    Planes rec[2], pre{};
    CUtensorMap ref_map[2];           // tiled TMA descriptors of the two padded luma reconstruction planes (search-window box)             // pre: reconstruction before SAO (SAO on); rec[cur] then receives the SAO output
    uint8_t *ds = nullptr;            // [max_batch + 1][dsh * dsw] quarter-resolution planes (8 MSBs)
    int16_t *cmv = nullptr;           // [max_batch][ctus][2]
    int *mode_cost = nullptr;         // [cus][35] intra mode search result of the frame in flight
    int *intra_best = nullptr;        // [cus] best of those per CU
    int *cand_list = nullptr;         // [cus] work list of the intra search
    SceneStat *scene = nullptr;       // [max_batch] scene-cut measures of the batch in flight (frame chain is serial: one array)
    uint32_t *mvf[2] = {nullptr, nullptr};   // [cus] motion field, ping-pong between the merge-aware passes of one frame
    int *satdf[2] = {nullptr, nullptr};
    size_t staging_bytes = 0;
    uint8_t *csc_tmp = nullptr;       // 8-bit 4:2:0 scratch frame of the scaled-RGB ingest path
    size_t csc_tmp_bytes = 0;
    int *progress = nullptr;
    RcState *rc_dev = nullptr;
    const char *trace_path = nullptr;
    size_t packed_cap = 0;
    // stream state
    int cur = 0;                      // reconstruction buffer being written
    long long frame_no = 0;
    int since_bp = 0;
    std::atomic<int> stop{0}, done{0};
    cudaStream_t st_copy = nullptr;
    cudaStream_t st_entropy[kEntropyStreams] = {};   // CABAC launches round-robin over these, concurrent with the frame chain
    int next_entropy_stream = 0;
    // drain: every kGroupFrames frames the finished CABAC payload is compacted and downloaded (own streams) and the host
    // assembles those access units while the GPU is still encoding the rest of the batch
    cudaStream_t st_drain = nullptr, st_dl = nullptr;
    size_t frame_cap = 0;             // bytes of the packed buffer reserved per frame
    int profiling = 0;
    float prof_ms[8] = {};            // inter (3 kernels), intra, coarse, entropy(+pack), ingest, chain, k_me alone, spare
    int prof_launches[8] = {};
    float last_total_ms = 0, last_kernel_ms = 0;
    cudaEvent_t ev_mark = nullptr, ev_last_done = nullptr;   // region timing across pipelined calls (hb_enc_mark / hb_enc_elapsed)
    bool have_last_done = false, have_mark = false;
    std::vector<uint8_t> vps, sps, pps;
};

namespace {

template <typename T>
int dev_alloc(hb_encoder *e, T **out, size_t count)
{
    void *p = nullptr;
    cudaError_t err = cudaMalloc(&p, count * sizeof(T) + 256);
    if (err != cudaSuccess)
        return hb_fail(e->ctx, HB_ERR_NOMEM, "cudaMalloc: %s", cudaGetErrorString(err));
    e->dev.push_back(p);
    *out = reinterpret_cast<T *>(p);
    return HB_OK;
}

#define HB_TRY(x)                  \
    do {                           \
        int rc_ = (x);             \
        if (rc_ != HB_OK) return rc_; \
    } while (0)

int alloc_planes(hb_encoder *e, Planes *pl, int wc, int hc)
{
    HB_TRY(dev_alloc(e, &pl->y, (size_t)wc * hc));
    HB_TRY(dev_alloc(e, &pl->u, (size_t)(wc / 2) * (hc / 2)));
    HB_TRY(dev_alloc(e, &pl->v, (size_t)(wc / 2) * (hc / 2)));
    return HB_OK;
}

RcState initial_rc(const hb_enc_params &p)
{
    RcState rc{};
    rc.t16 = (long long)p.vbv_maxrate_kbps * 1000 * 16 * p.fps_den / p.fps_num;
    rc.b16 = (long long)p.vbv_bufsize_kbit * 1000 * 16;
    rc.fullness = rc.b16 * 9 / 10;
    rc.qp_i = p.qp_i; rc.qp_p = p.qp_p; rc.rate_control = p.rate_control; rc.bit_depth = p.bit_depth;
    rc.keyint = p.keyint; rc.min_keyint = p.min_keyint; rc.scenecut = p.scenecut; rc.poc = 0; rc.started = 0;
    rc.cascade = p.qp_cascade;
    return rc;
}

int upload_initial_rc(hb_encoder *e)
{
    const RcState rc = initial_rc(e->prm);
    HB_CUDA(e->ctx, cudaMemcpyAsync(e->rc_dev, &rc, sizeof(rc), cudaMemcpyHostToDevice, e->ctx->stream));
    HB_CUDA(e->ctx, cudaStreamSynchronize(e->ctx->stream));
    return HB_OK;
}

// launch the WPP CABAC kernel for frames [first, first + count) of the current batch on the next side stream, ordered after
// everything issued so far on the main stream
int launch_entropy(hb_encoder *e, BatchSet &B, int first, int count, cudaEvent_t ev)
{
    hb_ctx *ctx = e->ctx;
    cudaStream_t q = e->st_entropy[e->next_entropy_stream];
    e->next_entropy_stream = (e->next_entropy_stream + 1) % kEntropyStreams;
    HB_CUDA(ctx, cudaEventRecord(ev, ctx->stream));
    HB_CUDA(ctx, cudaStreamWaitEvent(q, ev, 0));
    EntropyParams ep;
    ep.g = e->g; ep.frames = B.eframes_dev + first; ep.row_cap = e->row_cap; ep.overflow = B.overflow;
    k_entropy<<<dim3((e->g.ctuh + kEntropyWarps - 1) / kEntropyWarps, count), kEntropyWarps * 32, 0, q>>>(ep);
    HB_LAUNCHED(ctx);
    return HB_OK;
}

size_t input_frame_bytes(const hb_enc_params &p, int fmt, int sw = 0, int sh = 0)
{
    const int w = sw > 0 ? sw : p.width, h = sh > 0 ? sh : p.height;
    const size_t luma = (size_t)w * h, chroma = (size_t)(w / 2) * (h / 2);
    if (fmt == HB_PIX_BGR24 || fmt == HB_PIX_RGB24) return 3 * luma;
    return fmt == HB_PIX_YUV420P8 ? luma + 2 * chroma : 2 * (luma + 2 * chroma);
}


// Enqueue everything one batch needs on the GPU -- uploads, frame chain, CABAC, per-group compaction -- without waiting for
// any of it.  `fr` frames [base, base + n) go into batch set B, whose previous batch has been drained.
int enqueue_batch(hb_encoder *e, BatchSet &B, const hb_frames *fr, int base, int n, bool force_first)
{
    hb_ctx *ctx = e->ctx;
    const hb_enc_params &p = e->prm;
    const Geom &g = e->g;
    const size_t fbytes = input_frame_bytes(p, fr->format, fr->src_width, fr->src_height);
    cudaStream_t st = ctx->stream;
    const int ncu = g.cuw * g.cuh, nctu = g.ctuw * g.ctuh;
    const size_t ds_stride = (size_t)g.dsw * g.dsh;
    {
        if (e->stop.load()) return hb_fail(ctx, HB_ERR_STOPPED, "%s", "stopped");
        HB_CUDA(ctx, cudaEventRecord(B.ev[0], st));
        // ---- frame types are decided on the device (k_rc_step: key-frame cadence + scene cuts) and read back with the frame
        //      controls; the host only knows which frames it FORCES to be key frames (first of the stream / of a segment)
        B.n = n; B.forced_first = force_first; B.first_frame_no = e->frame_no;
        B.is_idr.assign(n, 0); B.qps.assign(n, 0); B.pocs.assign(n, 0);
        std::vector<int> &is_idr = B.is_idr;
        for (int i = 0; i < n; i++) is_idr[i] = (force_first && i == 0) || e->frame_no + i == 0;      // forced key frames
        std::vector<EntropyFrame> ef(n);
        for (int i = 0; i < n; i++) {
            FrameSlot &s = B.slot[i];
            ef[i].cus = s.cus; ef[i].syn = s.syn; ef[i].coefs = s.coefs; ef[i].out = s.rows; ef[i].row_len = s.row_len;
            ef[i].ctl = B.ctl_dev + i; ef[i].sao = s.sao; ef[i].ctx_save = s.ctx_save; ef[i].row_ready = s.row_ready;
            ef[i].trace = B.trace_dev ? B.trace_dev + (size_t)i * g.ctuh * kTraceWords : nullptr;
        }
        HB_CUDA(ctx, cudaMemcpyAsync(B.eframes_dev, ef.data(), sizeof(EntropyFrame) * n, cudaMemcpyHostToDevice, st));
        HB_CUDA(ctx, cudaMemsetAsync(B.overflow, 0, sizeof(int), st));
        HB_CUDA(ctx, cudaMemsetAsync(e->scene, 0, sizeof(SceneStat) * n, st));
        HB_CUDA(ctx, cudaMemsetAsync(B.row_ready_all, 0, sizeof(int) * (size_t)n * g.ctuh, st));
        // (the set's staging buffer and sync areas are free: its previous batch has been drained)
        HB_CUDA(ctx, cudaEventRecord(B.ev_misc[0], st));
        for (auto &q : e->st_entropy) HB_CUDA(ctx, cudaStreamWaitEvent(q, B.ev_misc[0], 0));
        const uint8_t *in = static_cast<const uint8_t *>(fr->data) + (size_t)base * fr->frame_bytes;
        const size_t kb = (size_t)2 * e->max_batch;      // batch-level profiling events start here
        float ingest_ms = 0, coarse_ms = 0;
        (void)ingest_ms; (void)coarse_ms;
        // CABAC launches for the P frames [pend, end) not handed over yet (runs between key frames)
        int pend = 0, cur_chunk = 0;
        auto launch_p_runs = [&](int end) -> int {
            for (int i0 = pend; i0 < end;) {
                if (is_idr[i0]) { i0++; continue; }
                int i1 = i0;
                while (i1 < end && !is_idr[i1]) i1++;
                HB_TRY(launch_entropy(e, B, i0, i1 - i0, B.ev_chunk[kMaxChunks + cur_chunk]));
                i0 = i1;
            }
            pend = end;
            return HB_OK;
        };
        // ---- chunks: upload (copy stream) | ingest + coarse search + frame chain (main stream) | CABAC (entropy stream)
        for (int c0 = 0, chunk = 0; c0 < n; c0 += kChunk, chunk++) {
            const int cn = std::min(kChunk, n - c0);
            cur_chunk = chunk;
            const uint8_t *dev_in = in + (size_t)c0 * fr->frame_bytes;
            size_t dev_fb = fr->frame_bytes;
            if (!fr->on_device) {
                if (fr->frame_bytes == fbytes) {
                    HB_CUDA(ctx, cudaMemcpyAsync(B.staging + (size_t)c0 * fbytes, dev_in, fbytes * cn, cudaMemcpyHostToDevice, e->st_copy));
                } else {
                    for (int i = 0; i < cn; i++)
                        HB_CUDA(ctx, cudaMemcpyAsync(B.staging + (size_t)(c0 + i) * fbytes, dev_in + (size_t)i * fr->frame_bytes, fbytes,
                                                     cudaMemcpyHostToDevice, e->st_copy));
                }
                HB_CUDA(ctx, cudaEventRecord(B.ev_chunk[chunk], e->st_copy));
                HB_CUDA(ctx, cudaStreamWaitEvent(st, B.ev_chunk[chunk], 0));
                dev_in = B.staging + (size_t)c0 * fbytes;
                dev_fb = fbytes;
            }
            if (c0 == 0) {
                HB_CUDA(ctx, cudaEventRecord(B.ev[1], st));
                if (e->profiling) HB_CUDA(ctx, cudaEventRecord(B.kev[kb + 0], st));
            }
            const bool scaled = fr->src_width > 0 && (fr->src_width != p.width || fr->src_height != p.height);
            const bool packed_rgb = fr->format == HB_PIX_BGR24 || fr->format == HB_PIX_RGB24;
            for (int i = 0; i < cn && (scaled || packed_rgb); i++) {
                // fused pre-encode pixel pipeline: the scaler / colour conversion writes the encoder's source planes directly
                const uint8_t *f = dev_in + (size_t)i * dev_fb;
                const Planes &sp = B.slot[c0 + i].src;
                int m = fr->matrix;
                // default: the matrix the stream will signal; unspecified (2) -> BT.601, which is what libswscale (and so OpenCV's
                // decoder front end) assumes for untagged YUV sources when it produced these RGB samples
                if (!m) m = p.matrix_coeffs == 9 ? HB_MATRIX_BT2020 : p.matrix_coeffs == 1 ? HB_MATRIX_BT709 : HB_MATRIX_BT601;
                if (packed_rgb && scaled) {
                    // colour conversion at the source size into an 8-bit 4:2:0 scratch frame, then the scaler (frames are serial on
                    // this stream, so one scratch frame is enough)
                    const int sw = fr->src_width, sh = fr->src_height;
                    const size_t need = (size_t)sw * sh * 3 / 2;
                    if (e->csc_tmp_bytes < need) {
                        uint8_t *t = nullptr;
                        HB_TRY(dev_alloc(e, &t, need));
                        e->csc_tmp = t; e->csc_tmp_bytes = need;
                    }
                    uint8_t *ty = e->csc_tmp, *tu = ty + (size_t)sw * sh, *tv = tu + (size_t)(sw / 2) * (sh / 2);
                    HB_TRY(hb_rgb_to_yuv420(ctx, (hb_devptr)(uintptr_t)f, 3 * sw, fr->format == HB_PIX_BGR24 ? HB_BGR : HB_RGB, m, 8, sw, sh,
                                            (hb_devptr)(uintptr_t)ty, sw, (hb_devptr)(uintptr_t)tu, sw / 2, (hb_devptr)(uintptr_t)tv, sw / 2));
                    HB_TRY(launch_scale8(ctx, ty, nullptr, sw, sw, sh, reinterpret_cast<uint8_t *>(sp.y), nullptr, g.src_stride * 2, p.width, p.height,
                                         p.bit_depth, 0, 1));
                    HB_TRY(launch_scale8(ctx, tu, tv, sw / 2, sw / 2, sh / 2, reinterpret_cast<uint8_t *>(sp.u), reinterpret_cast<uint8_t *>(sp.v),
                                         g.srcc_stride * 2, p.width / 2, p.height / 2, p.bit_depth, 0, 1));
                } else if (packed_rgb) {
                    HB_TRY(launch_rgb_planar16(ctx, f, 3 * p.width, fr->format == HB_PIX_BGR24, m, p.bit_depth, p.width, p.height,
                                               reinterpret_cast<uint8_t *>(sp.y), g.src_stride * 2, reinterpret_cast<uint8_t *>(sp.u),
                                               reinterpret_cast<uint8_t *>(sp.v), g.srcc_stride * 2));
                } else {
                    const int sw = fr->src_width, sh = fr->src_height;
                    const uint8_t *fu = f + (size_t)sw * sh, *fv = fu + (size_t)(sw / 2) * (sh / 2);
                    HB_TRY(launch_scale8(ctx, f, nullptr, sw, sw, sh, reinterpret_cast<uint8_t *>(sp.y), nullptr, g.src_stride * 2, p.width, p.height,
                                         p.bit_depth, 0, 1));
                    HB_TRY(launch_scale8(ctx, fu, fv, sw / 2, sw / 2, sh / 2, reinterpret_cast<uint8_t *>(sp.u), reinterpret_cast<uint8_t *>(sp.v),
                                         g.srcc_stride * 2, p.width / 2, p.height / 2, p.bit_depth, 0, 1));
                }
                k_pad_ds<<<hb_grid_for(ctx, (long long)g.dsw * g.dsh, 256, 8), 256, 0, st>>>(g, sp, p.width, p.height,
                                                                                           e->ds + (size_t)(c0 + i + 1) * ds_stride);
                HB_LAUNCHED(ctx);
            }
            for (int i = 0; i < cn && !(scaled || packed_rgb); i++) {
                IngestParams ip;
                ip.g = g; ip.round_add = 0; ip.maxv = (1 << p.bit_depth) - 1;
                const uint8_t *f = dev_in + (size_t)i * dev_fb;
                const size_t luma = (size_t)p.width * p.height, chroma = (size_t)(p.width / 2) * (p.height / 2);
                if (fr->format == HB_PIX_YUV420P8) {
                    ip.in_y = f; ip.in_u = f + luma; ip.in_v = f + luma + chroma;
                    ip.in_ys = p.width; ip.in_us = ip.in_vs = p.width / 2;
                    ip.up_shift = p.bit_depth - 8; ip.down_shift = 0;
                } else if (fr->format == HB_PIX_P010) {
                    ip.in_y = f; ip.in_u = f + 2 * luma; ip.in_v = nullptr;
                    ip.in_ys = 2 * p.width; ip.in_us = ip.in_vs = 2 * p.width;
                    ip.up_shift = 0; ip.down_shift = 16 - p.bit_depth;
                } else {
                    ip.in_y = f; ip.in_u = f + 2 * luma; ip.in_v = f + 2 * luma + 2 * chroma;
                    ip.in_ys = 2 * p.width; ip.in_us = ip.in_vs = p.width;
                    const int sd = fr->src_bit_depth ? fr->src_bit_depth : p.bit_depth;
                    ip.up_shift = sd < p.bit_depth ? p.bit_depth - sd : 0;
                    ip.down_shift = sd > p.bit_depth ? sd - p.bit_depth : 0;
                    ip.round_add = ip.down_shift ? 1 << (ip.down_shift - 1) : 0;
                }
                ip.fmt = fr->format; ip.w = p.width; ip.h = p.height;
                ip.src = B.slot[c0 + i].src;
                ip.ds = e->ds + (size_t)(c0 + i + 1) * ds_stride;
                k_ingest<<<hb_grid_for(ctx, (long long)g.dsw * g.dsh, 256, 8), 256, 0, st>>>(ip);
                HB_LAUNCHED(ctx);
            }
            // coarse motion search for the whole chunk in one launch (source-based, independent of the reconstruction chain)
            {
                CoarseParams cp;
                cp.g = g; cp.ds = e->ds + (size_t)c0 * ds_stride; cp.ds_frame_stride = ds_stride; cp.cmv = e->cmv + (size_t)c0 * nctu * 2; cp.scene = e->scene + c0;
                k_coarse<<<dim3(nctu, cn), 128, 0, st>>>(cp);
                HB_LAUNCHED(ctx);
            }
            // frame chain
            for (int i = c0; i < c0 + cn; i++) {
                const bool forced = is_idr[i] != 0;
                const bool first_of_stream = e->frame_no == 0;
                FrameSlot &s = B.slot[i];
                const Planes &fin = e->rec[e->cur], &ref = e->rec[1 - e->cur];
                const Planes &rec = p.sao ? e->pre : fin;      // the frame is reconstructed and deblocked here; SAO writes `fin`
                // rate control + frame type on the device: account for the previous frame, choose this frame's type and QP
                k_rc_step<<<1, 32, 0, st>>>(e->rc_dev, i > 0 ? B.ctl_dev + i - 1 : nullptr, B.ctl_dev + i, forced ? 1 : 0, e->scene + i,
                                            (long long)nctu * 64);
                HB_LAUNCHED(ctx);
                HB_CUDA(ctx, cudaMemsetAsync(e->progress, 0, sizeof(int) * g.ctuh, st));
                if (e->profiling) HB_CUDA(ctx, cudaEventRecord(B.kev[2 * i], st));
                // Every frame gets both kernel families; each kernel looks at the frame type in its control block and returns
                // at once when it does not apply (key frame <-> inter kernels), so nothing waits for the host.
                IntraParams np;
                np.g = g; np.src = s.src; np.rec = rec; np.cus = s.cus; np.coefs = s.coefs; np.progress = e->progress;
                np.ctl = B.ctl_dev + i; np.second_pass = 0; np.mode_cost = e->mode_cost; np.intra_best = e->intra_best;
                np.intra_in_p = p.intra_in_p; np.satd1 = e->satdf[1]; np.cand_list = e->cand_list;
                if (!forced) {
                    InterParams ip;
                    ip.ref_map = e->ref_map[1 - e->cur];
                    ip.g = g; ip.src = s.src; ip.ref = ref; ip.rec = rec; ip.cmv = e->cmv + (size_t)i * nctu * 2;
                    ip.cus = s.cus; ip.coefs = s.coefs; ip.ctl = B.ctl_dev + i;
                    ip.intra_best = p.intra_in_p ? e->intra_best : nullptr;
                    // motion search, then two merge-aware passes over the field (the second one inside k_inter)
                    ip.mv_in = nullptr; ip.satd_in = nullptr; ip.mv_out = e->mvf[0]; ip.satd_out = e->satdf[0];
                    static const int xs = getenv("HB_DEBUG_EXTRA_SMEM") ? atoi(getenv("HB_DEBUG_EXTRA_SMEM")) : 0;      // occupancy experiments
                    k_me<<<nctu, 128, xs, st>>>(ip);
                    HB_LAUNCHED(ctx);
                    if (e->profiling) HB_CUDA(ctx, cudaEventRecord(B.kev_me[i], st));
                    ip.mv_in = e->mvf[0]; ip.satd_in = e->satdf[0]; ip.mv_out = e->mvf[1]; ip.satd_out = e->satdf[1];
                    k_merge<<<nctu, 128, xs, st>>>(ip);
                    HB_LAUNCHED(ctx);
                    k_intra_list<<<(ncu + 255) / 256, 256, 0, st>>>(np);
                    HB_LAUNCHED(ctx);
                    k_intra_search<<<std::min((3 * ncu + 3) / 4, 16 * ctx->sm_count), kIntraSearchThreads, 0, st>>>(np);
                    HB_LAUNCHED(ctx);
                    ip.mv_in = e->mvf[1]; ip.satd_in = e->satdf[1]; ip.mv_out = nullptr; ip.satd_out = nullptr;
                    k_inter<<<nctu, 128, xs, st>>>(ip);
                    HB_LAUNCHED(ctx);
                } else {
                    if (e->profiling) HB_CUDA(ctx, cudaEventRecord(B.kev_me[i], st));
                    k_intra_list<<<(ncu + 255) / 256, 256, 0, st>>>(np);
                    HB_LAUNCHED(ctx);
                    k_intra_search<<<std::min((3 * ncu + 3) / 4, 16 * ctx->sm_count), kIntraSearchThreads, 0, st>>>(np);
                    HB_LAUNCHED(ctx);
                }
                k_intra<<<g.ctuh, kIntraReconThreads, 0, st>>>(np);
                HB_LAUNCHED(ctx);
                if (p.rate_control && first_of_stream) {     // first key frame of the stream: second try if it overshot
                    k_rc_redo<<<1, 32, 0, st>>>(e->rc_dev, B.ctl_dev + i);
                    HB_LAUNCHED(ctx);
                    HB_CUDA(ctx, cudaMemsetAsync(e->progress, 0, sizeof(int) * g.ctuh, st));
                    np.second_pass = 1;
                    k_intra<<<g.ctuh, kIntraReconThreads, 0, st>>>(np);
                    HB_LAUNCHED(ctx);
                }
                if (e->profiling) HB_CUDA(ctx, cudaEventRecord(B.kev[2 * i + 1], st));
                if (p.deblock) {
                    DeblockParams dp;
                    dp.g = g; dp.rec = rec; dp.cus = s.cus; dp.ctl = B.ctl_dev + i;
                    for (dp.dir = 0; dp.dir < 2; dp.dir++) {
                        k_deblock<<<hb_grid_for(ctx, (long long)ncu * 4, 256, 8), 256, 0, st>>>(dp);
                        HB_LAUNCHED(ctx);
                    }
                }
                if (p.sao) {
                    SaoParams sp;
                    sp.g = g; sp.src = s.src; sp.pre = rec; sp.out = fin; sp.cus = s.cus; sp.ctl = B.ctl_dev + i; sp.sao = s.sao;
                    k_sao_decide<<<nctu, 256, 0, st>>>(sp);
                    HB_LAUNCHED(ctx);
                    k_sao_apply<<<nctu, 256, 0, st>>>(sp);
                    HB_LAUNCHED(ctx);
                }
                {
                    const int border = 2 * kPad * (g.wc + 2 * kPad) + g.hc * 2 * kPad;
                    k_border<<<dim3((border + 255) / 256, 3), 256, 0, st>>>(fin, g);
                    HB_LAUNCHED(ctx);
                }
                if (p.keep_recon || p.hash_sei) {
                    const pixel *srcp[3] = {fin.y, fin.u, fin.v};
                    pixel *dstp[3] = {s.keep.y, s.keep.u, s.keep.v};
                    for (int c = 0; c < 3; c++) {
                        const int w = c ? g.wc / 2 : g.wc, h = c ? g.hc / 2 : g.hc, stride = c ? g.recc_stride : g.rec_stride;
                        HB_CUDA(ctx, cudaMemcpy2DAsync(dstp[c], (size_t)w * sizeof(pixel), srcp[c], (size_t)stride * sizeof(pixel),
                                                       (size_t)w * sizeof(pixel), h, cudaMemcpyDeviceToDevice, st));
                    }
                }
                ModeParams mp;
                mp.g = g; mp.cus = s.cus; mp.syn = s.syn; mp.ctl = B.ctl_dev + i;
                k_modes<<<(ncu + 255) / 256, 256, 0, st>>>(mp);
                HB_LAUNCHED(ctx);
                if (forced) HB_TRY(launch_entropy(e, B, i, 1, B.ev_chunk[kMaxChunks + chunk]));
                e->cur = 1 - e->cur;
                e->frame_no++;
                // last chunk of the batch: hand its frames to CABAC in shrinking groups (.., 2, 1, 1), so that when the chain
                // ends only the last frame still has its whole entropy latency ahead of it
                if (c0 + cn == n) {
                    const int left = c0 + cn - (i + 1);
                    if (left == 4 || left == 2 || left == 1) HB_TRY(launch_p_runs(i + 1));
                }
            }
            // CABAC of this chunk on a side stream, overlapping the frame chain of the following chunks (key frames were
            // already launched on their own right after their mode kernel: they carry ~8x the bins of a P frame)
            HB_TRY(launch_p_runs(c0 + cn));
            // ---- end of a drain group: compaction + download of its sub-stream sizes behind its CABAC kernels
            if ((c0 + cn) % kGroupFrames == 0 || c0 + cn == n) {
                const int grp = c0 / kGroupFrames, f0 = grp * kGroupFrames, cnt = c0 + cn - f0;
                for (int k = 0; k < kEntropyStreams; k++) {
                    HB_CUDA(ctx, cudaEventRecord(B.ev_grp[grp][k], e->st_entropy[k]));
                    HB_CUDA(ctx, cudaStreamWaitEvent(e->st_drain, B.ev_grp[grp][k], 0));
                }
                PackParams pp;
                pp.frames = B.eframes_dev + f0; pp.n_frames = cnt; pp.rows = g.ctuh; pp.row_cap = e->row_cap;
                pp.packed = B.packed_dev + (size_t)f0 * e->frame_cap; pp.cap = (unsigned long long)cnt * e->frame_cap; pp.offsets = B.offsets_dev + (size_t)f0 * g.ctuh + grp;
                k_pack_scan<<<1, 1024, 0, e->st_drain>>>(pp);
                HB_LAUNCHED(ctx);
                k_pack_copy<<<cnt * g.ctuh, 128, 0, e->st_drain>>>(pp);
                HB_LAUNCHED(ctx);
                HB_CUDA(ctx, cudaMemcpyAsync(B.offsets_host + (size_t)f0 * g.ctuh + grp, pp.offsets, sizeof(uint32_t) * ((size_t)cnt * g.ctuh + 1),
                                             cudaMemcpyDeviceToHost, e->st_drain));
                HB_CUDA(ctx, cudaMemcpyAsync(B.ctl_host + f0, B.ctl_dev + f0, sizeof(FrameCtl) * cnt, cudaMemcpyDeviceToHost, e->st_drain));
                if (c0 + cn == n) {
                    HB_CUDA(ctx, cudaMemcpyAsync(B.overflow_host, B.overflow, sizeof(int), cudaMemcpyDeviceToHost, e->st_drain));
                    HB_CUDA(ctx, cudaEventRecord(B.ev[2], e->st_drain));
                    if (e->profiling) HB_CUDA(ctx, cudaEventRecord(B.kev[kb + 4], e->st_drain));
                }
                HB_CUDA(ctx, cudaEventRecord(B.ev_off[grp], e->st_drain));
            }
        }
        k_rc_step<<<1, 32, 0, st>>>(e->rc_dev, B.ctl_dev + n - 1, nullptr, 0, nullptr, 0);
        HB_LAUNCHED(ctx);
        // the last frame's quarter-resolution plane becomes slot 0 (the predecessor) of the next batch / call
        HB_CUDA(ctx, cudaMemcpyAsync(e->ds, e->ds + (size_t)n * ds_stride, ds_stride * sizeof(uint8_t), cudaMemcpyDeviceToDevice, st));
        if (e->profiling) HB_CUDA(ctx, cudaEventRecord(B.kev[kb + 3], st));
        HB_CUDA(ctx, cudaEventRecord(B.ev_misc[1], st));        // end of this batch's frame chain
        B.pending = true;
    }
    return HB_OK;
}


// Wait for a batch group by group, download its payload and assemble its access units into out[written...]; stats (may be
// null) is indexed from the batch's first frame.
int drain_batch(hb_encoder *e, BatchSet &B, uint8_t *out, size_t cap, size_t &written, hb_frame_stat *stats, float &total_ms, float &kernel_ms)
{
    hb_ctx *ctx = e->ctx;
    const hb_enc_params &p = e->prm;
    const Geom &g = e->g;
    const int n = B.n;
    const size_t kb = (size_t)2 * e->max_batch;
    std::vector<int> &is_idr = B.is_idr, &qps = B.qps, &pocs = B.pocs;
    {
        // ---- drain: per group wait for its sizes, download its payload, assemble its access units (the GPU keeps encoding)
        std::vector<uint8_t> au, slice;
        std::vector<uint16_t> hostrec;
        const int n_groups = (n + kGroupFrames - 1) / kGroupFrames;
        for (int grp = 0; grp < n_groups; grp++) {
        const int f0 = grp * kGroupFrames, f1 = std::min(n, f0 + kGroupFrames);
        const uint32_t *goff = B.offsets_host + (size_t)f0 * g.ctuh + grp;
        const uint8_t *gpay = B.packed_host + (size_t)f0 * e->frame_cap;
        HB_CUDA(ctx, cudaEventSynchronize(B.ev_off[grp]));
        for (int i = f0; i < f1; i++) { qps[i] = B.ctl_host[i].qp; is_idr[i] = B.ctl_host[i].is_idr; pocs[i] = B.ctl_host[i].poc; }
        {
            const uint32_t total = goff[(size_t)(f1 - f0) * g.ctuh];
            if (total > (size_t)(f1 - f0) * e->frame_cap) return hb_fail(ctx, HB_ERR_SPACE, "%s", "packed bitstream exceeds the download buffer");
            HB_CUDA(ctx, cudaMemcpyAsync(B.packed_host + (size_t)f0 * e->frame_cap, B.packed_dev + (size_t)f0 * e->frame_cap, total,
                                         cudaMemcpyDeviceToHost, e->st_dl));
            if (grp == n_groups - 1) HB_CUDA(ctx, cudaEventRecord(B.ev[3], e->st_dl));
            HB_CUDA(ctx, cudaStreamSynchronize(e->st_dl));
        }
        if (grp == n_groups - 1) {
            HB_CUDA(ctx, cudaEventRecord(e->ev_last_done, e->st_dl));
            e->have_last_done = true;
            HB_CUDA(ctx, cudaEventSynchronize(B.ev_misc[1]));        // the batch's frame chain (the next batch may already be running)
            if (*B.overflow_host) return hb_fail(ctx, HB_ERR_SPACE, "%s", "CABAC sub-stream exceeded its row buffer");
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, B.ev[0], B.ev[3]);
            cudaEventElapsedTime(&b, B.ev[1], B.ev[2]);
            total_ms += a; kernel_ms += b;
            if (e->profiling) {
                float ms = 0;
                for (int i = 0; i < n; i++) {
                    cudaEventElapsedTime(&ms, B.kev[2 * i], B.kev[2 * i + 1]);
                    const int k = is_idr[i] ? 1 : 0;
                    e->prof_ms[k] += ms; e->prof_launches[k]++;
                    if (!is_idr[i]) { cudaEventElapsedTime(&ms, B.kev[2 * i], B.kev_me[i]); e->prof_ms[6] += ms; e->prof_launches[6]++; }
                }
                cudaEventElapsedTime(&ms, B.kev[kb + 3], B.kev[kb + 4]); e->prof_ms[3] += ms; e->prof_launches[3] += 2;     // entropy tail + compaction
                cudaEventElapsedTime(&ms, B.kev[kb + 0], B.kev[kb + 3]); e->prof_ms[5] += ms; e->prof_launches[5] += 4 * n;   // ingest + coarse + frame chain
            }
            if (B.trace_dev) {      // debug dump: n, ctuh, then per frame / row {start ns, end ns, bytes, binarisation cycles, coding cycles, list entries}
                std::vector<unsigned long long> tr((size_t)n * g.ctuh * kTraceWords);
                cudaMemcpy(tr.data(), B.trace_dev, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
                if (FILE *f = fopen(e->trace_path, "wb")) {
                    const unsigned long long hdr[2] = {(unsigned long long)n, (unsigned long long)g.ctuh};
                    fwrite(hdr, sizeof(hdr), 1, f);
                    for (size_t k = 0; k < (size_t)n * g.ctuh; k++) {
                        const size_t fi = k / g.ctuh, gi = fi / kGroupFrames;
                        const uint32_t *o = B.offsets_host + k + gi;
                        const unsigned long long *t = &tr[kTraceWords * k];
                        const unsigned long long rec[6] = {t[0], t[1], (unsigned long long)(o[1] - o[0]), t[2], t[3], t[4]};
                        fwrite(rec, sizeof(rec), 1, f);
                    }
                    fclose(f);
                }
            }
        }
        for (int i = f0; i < f1; i++) {

            au.clear();
            const bool idr = is_idr[i] != 0;
            const bool first_of_stream = B.first_frame_no + i == 0;
            if (p.aud) { BitWriter b; b.put(idr ? 0 : 1, 3); b.trailing(); append_nal(au, NAL_AUD, b.bytes(), true); }
            if (idr && (first_of_stream || p.repeat_headers || (B.forced_first && i == 0))) {
                append_nal(au, NAL_VPS, e->vps, true);
                append_nal(au, NAL_SPS, e->sps, true);
                append_nal(au, NAL_PPS, e->pps, true);
            }
            if (p.hrd) {
                // au_cpb_removal_delay counts from the most recent buffering period in a PRECEDING access unit (D.3.2), so an
                // access unit that carries a buffering period itself still reports its distance to the previous one
                const int delay = e->since_bp;
                if (idr) {
                    BitWriter m;
                    const uint32_t delay = (uint32_t)((long long)90000 * 9 * p.vbv_bufsize_kbit / ((long long)10 * p.vbv_maxrate_kbps));
                    m.ue(0); m.flag(false); m.flag(false); m.put(0, 24); m.put(delay, 24); m.put(0, 24);
                    if (!m.aligned()) m.trailing();
                    append_nal(au, NAL_SEI_PREFIX, make_sei(0, m.bytes()), false);
                    e->since_bp = 0;
                }
                BitWriter m;
                m.put((uint32_t)(delay > 0 ? delay - 1 : 0), 24); m.put(0, 24);
                append_nal(au, NAL_SEI_PREFIX, make_sei(1, m.bytes()), false);
                e->since_bp++;
            }
            if (idr && p.hdr10) {
                BitWriter m;
                for (int k = 0; k < 8; k++) m.put(p.master_display[k], 16);
                m.put(p.master_display[8], 32); m.put(p.master_display[9], 32);
                append_nal(au, NAL_SEI_PREFIX, make_sei(137, m.bytes()), false);
                BitWriter c;
                c.put(p.max_cll, 16); c.put(p.max_fall, 16);
                append_nal(au, NAL_SEI_PREFIX, make_sei(144, c.bytes()), false);
            }
            // slice: sub-streams are escaped independently (none can end in 0x00), entry points count escaped bytes
            slice.clear();
            std::vector<uint32_t> entry(g.ctuh);
            for (int r = 0; r < g.ctuh; r++) {
                const uint32_t o0 = goff[(size_t)(i - f0) * g.ctuh + r], o1 = goff[(size_t)(i - f0) * g.ctuh + r + 1];
                const size_t before = slice.size();
                append_escaped(slice, gpay + o0, o1 - o0);
                entry[r] = (uint32_t)(slice.size() - before);
            }
            BitWriter h;
            const int nal_type = idr ? NAL_IDR_W_RADL : NAL_TRAIL_R;
            h.flag(true);
            if (idr) h.flag(false);
            h.ue(0);
            h.ue(idr ? 2 : 1);
            if (!idr) { h.put(pocs[i] & 255, 8); h.flag(true); }
            if (p.sao) { h.flag(true); h.flag(true); }              // slice_sao_luma_flag, slice_sao_chroma_flag
            if (!idr) { h.flag(false); h.ue(0); }
            h.se(qps[i] - 26);
            h.ue(g.ctuh - 1);
            if (g.ctuh > 1) {
                uint32_t mx = 0;
                for (int r = 0; r + 1 < g.ctuh; r++) mx = std::max(mx, entry[r] - 1);
                int len = 1;
                while (len < 32 && (mx >> len) != 0) len++;
                h.ue(len - 1);
                for (int r = 0; r + 1 < g.ctuh; r++) h.put(entry[r] - 1, len);
            }
            h.trailing();
            const bool long_start = au.empty();
            if (long_start) au.push_back(0);
            au.push_back(0); au.push_back(0); au.push_back(1);
            au.push_back((uint8_t)(nal_type << 1)); au.push_back(1);
            append_escaped(au, h.bytes().data(), h.bytes().size());
            au.insert(au.end(), slice.begin(), slice.end());
            if (p.hash_sei) {
                std::vector<uint8_t> pl(49, 0);
                const pixel *kp[3] = {B.slot[i].keep.y, B.slot[i].keep.u, B.slot[i].keep.v};
                for (int c = 0; c < 3; c++) {
                    const size_t cnt = (size_t)(c ? g.wc / 2 : g.wc) * (c ? g.hc / 2 : g.hc);
                    hostrec.resize(cnt);
                    HB_CUDA(ctx, cudaMemcpy(hostrec.data(), kp[c], cnt * sizeof(pixel), cudaMemcpyDeviceToHost));
                    if (p.bit_depth > 8) {
                        md5(reinterpret_cast<const uint8_t *>(hostrec.data()), cnt * 2, &pl[1 + 16 * c]);
                    } else {
                        std::vector<uint8_t> b8(cnt);
                        for (size_t k = 0; k < cnt; k++) b8[k] = (uint8_t)hostrec[k];
                        md5(b8.data(), cnt, &pl[1 + 16 * c]);
                    }
                }
                append_nal(au, NAL_SEI_SUFFIX, make_sei(132, pl), false);
            }
            if (written + au.size() > cap) return hb_fail(ctx, HB_ERR_SPACE, "%s", "output buffer too small");
            memcpy(out + written, au.data(), au.size());
            written += au.size();
            if (stats) {
                hb_frame_stat &s = stats[i];
                s.is_idr = idr; s.poc = pocs[i]; s.qp = qps[i]; s.bytes = (uint32_t)au.size(); s.n_skip = 0; s.n_merge = 0;
            }
            e->done.fetch_add(1);
        }
        }      // drain groups
    }
    B.pending = false;
    e->last_drained = &B;
    return HB_OK;
}

}  // namespace

extern "C" {

int hb_enc_create(hb_ctx *ctx, const hb_enc_params *params, int max_batch, hb_encoder **out)
{
    HB_ARG(ctx, ctx && params && out && max_batch >= 1 && max_batch <= 1024);
    const hb_enc_params &p = *params;
    HB_ARG(ctx, p.width >= 16 && p.height >= 16 && (p.width % 2) == 0 && (p.height % 2) == 0 && p.width <= 8192 && p.height <= 8192);
    HB_ARG(ctx, p.bit_depth == 8 || p.bit_depth == 10);
    HB_ARG(ctx, !p.rate_control || (p.vbv_maxrate_kbps > 0 && p.vbv_bufsize_kbit > 0));
    HB_ARG(ctx, p.qp_i >= 0 && p.qp_i <= 51 && p.qp_p >= 0 && p.qp_p <= 51 && p.keyint >= 1 && p.fps_num > 0 && p.fps_den > 0);
    HB_CUDA(ctx, cudaSetDevice(ctx->device));
    std::unique_ptr<hb_encoder> e(new hb_encoder());
    e->ctx = ctx;
    e->prm = p;
    e->max_batch = max_batch;
    Geom &g = e->g;
    g.wc = (p.width + 15) & ~15; g.hc = (p.height + 15) & ~15;
    g.cuw = g.wc / 16; g.cuh = g.hc / 16; g.ctuw = (g.wc + 31) / 32; g.ctuh = (g.hc + 31) / 32;
    g.bit_depth = p.bit_depth;
    g.src_stride = g.wc; g.srcc_stride = g.wc / 2;
    g.rec_stride = g.wc + 2 * kPad; g.recc_stride = g.wc / 2 + kPad;
    g.dsw = g.wc / 4; g.dsh = g.hc / 4;
    // worst-case CABAC payload of one CTU row: half the raw samples, at least 16 KiB
    e->row_cap = (uint32_t)std::max<size_t>(16384, (size_t)32 * g.wc * 3 / 2 * (p.bit_depth > 8 ? 2 : 1) / 2);
    e->row_cap = (e->row_cap + 255) & ~255u;
    const int ncu = g.cuw * g.cuh, nctu = g.ctuw * g.ctuh;
    hb_encoder *E = e.get();
    // source planes: one set of slots, shared by both batch sets
    std::vector<Planes> src(max_batch);
    for (int i = 0; i < max_batch; i++) HB_TRY(alloc_planes(E, &src[i], g.wc, g.hc));
    E->staging_bytes = (size_t)max_batch * input_frame_bytes(p, HB_PIX_P010);
    E->frame_cap = ((size_t)g.ctuh * E->row_cap / 4 + 65536 + 255) & ~(size_t)255;
    E->packed_cap = (size_t)max_batch * E->frame_cap;
    E->trace_path = getenv("HB_ENTROPY_TRACE");
    if (!(E->trace_path && *E->trace_path)) E->trace_path = nullptr;
    for (BatchSet &B : E->set) {
        B.slot.resize(max_batch);
        for (int i = 0; i < max_batch; i++) {
            FrameSlot &s = B.slot[i];
            s.src = src[i];
            HB_TRY(dev_alloc(E, &s.cus, (size_t)ncu));
            HB_TRY(dev_alloc(E, &s.syn, (size_t)ncu));
            HB_TRY(dev_alloc(E, &s.coefs, (size_t)ncu * kCuCoefs));
            HB_TRY(dev_alloc(E, &s.rows, (size_t)g.ctuh * E->row_cap));
            HB_TRY(dev_alloc(E, &s.row_len, (size_t)g.ctuh));
            HB_TRY(dev_alloc(E, &s.ctx_save, (size_t)g.ctuh * kNumCtx));
            if (p.keep_recon || p.hash_sei) HB_TRY(alloc_planes(E, &s.keep, g.wc, g.hc));
            if (p.sao) HB_TRY(dev_alloc(E, &s.sao, (size_t)nctu));
        }
        HB_TRY(dev_alloc(E, &B.row_ready_all, (size_t)max_batch * g.ctuh));
        for (int i = 0; i < max_batch; i++) B.slot[i].row_ready = B.row_ready_all + (size_t)i * g.ctuh;
        HB_TRY(dev_alloc(E, &B.staging, E->staging_bytes));
        HB_TRY(dev_alloc(E, &B.overflow, 1));
        HB_TRY(dev_alloc(E, &B.ctl_dev, (size_t)max_batch));
        HB_CUDA(ctx, cudaMallocHost(&B.ctl_host, sizeof(FrameCtl) * max_batch));
        HB_TRY(dev_alloc(E, &B.eframes_dev, (size_t)max_batch));
        if (E->trace_path) HB_TRY(dev_alloc(E, &B.trace_dev, (size_t)max_batch * g.ctuh * kTraceWords));
        HB_TRY(dev_alloc(E, &B.offsets_dev, (size_t)max_batch * g.ctuh + kMaxGroups + 1));
        HB_TRY(dev_alloc(E, &B.packed_dev, E->packed_cap));
        HB_CUDA(ctx, cudaMallocHost(&B.offsets_host, ((size_t)max_batch * g.ctuh + kMaxGroups + 1) * sizeof(uint32_t)));
        HB_CUDA(ctx, cudaMallocHost(&B.packed_host, E->packed_cap));
        HB_CUDA(ctx, cudaMallocHost(&B.overflow_host, sizeof(int)));
        for (auto &ev : B.ev) HB_CUDA(ctx, cudaEventCreate(&ev));
        for (auto &ev : B.ev_misc) HB_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        B.ev_chunk.resize((size_t)2 * kMaxChunks);
        for (auto &ev : B.ev_chunk) HB_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto &row : B.ev_grp) for (auto &ev : row) HB_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto &ev : B.ev_off) HB_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        B.kev.resize((size_t)2 * max_batch + 8);
        for (auto &ev : B.kev) HB_CUDA(ctx, cudaEventCreate(&ev));
        B.kev_me.resize((size_t)max_batch);
        for (auto &ev : B.kev_me) HB_CUDA(ctx, cudaEventCreate(&ev));
    }
    for (int k = 0; k < (p.sao ? 3 : 2); k++)
        for (int c = 0; c < 3; c++) {
            const int h = c ? g.hc / 2 : g.hc, pad = c ? kPad / 2 : kPad;
            const int stride = c ? g.recc_stride : g.rec_stride;
            HB_TRY(dev_alloc(E, &E->rec_base[k][c], (size_t)stride * (h + 2 * pad)));
            HB_CUDA(ctx, cudaMemsetAsync(E->rec_base[k][c], 0, (size_t)stride * (h + 2 * pad) * sizeof(pixel), ctx->stream));
            pixel *origin = E->rec_base[k][c] + (size_t)pad * stride + pad;
            Planes &dst = k == 2 ? E->pre : E->rec[k];
            if (c == 0) dst.y = origin; else if (c == 1) dst.u = origin; else dst.v = origin;
        }
    {   // tiled TMA descriptors (cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time libcuda dependency)
        typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                        const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        HB_CUDA(ctx, cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) return hb_fail(ctx, HB_ERR_CUDA, "%s", "cuTensorMapEncodeTiled is not available");
        for (int k = 0; k < 2; k++) {
            const cuuint64_t dims[2] = {(cuuint64_t)g.rec_stride, (cuuint64_t)(g.hc + 2 * kPad)};
            const cuuint64_t strides[1] = {(cuuint64_t)g.rec_stride * sizeof(pixel)};
            const cuuint32_t box[2] = {40, 28}, es[2] = {1, 1};
            const CUresult rc = ((EncodeTiled)fn)(&E->ref_map[k], CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, E->rec_base[k][0], dims, strides, box, es,
                                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rc != CUDA_SUCCESS) return hb_fail(ctx, HB_ERR_CUDA, "%s", "cuTensorMapEncodeTiled failed");
        }
    }
    HB_TRY(dev_alloc(E, &E->ds, (size_t)(max_batch + 1) * g.dsw * g.dsh));
    HB_TRY(dev_alloc(E, &E->cmv, (size_t)max_batch * nctu * 2));
    HB_TRY(dev_alloc(E, &E->mode_cost, (size_t)g.cuw * g.cuh * 35));
    HB_TRY(dev_alloc(E, &E->intra_best, (size_t)g.cuw * g.cuh));
    HB_TRY(dev_alloc(E, &E->cand_list, (size_t)g.cuw * g.cuh));
    HB_TRY(dev_alloc(E, &E->scene, (size_t)max_batch));
    for (int k = 0; k < 2; k++) {
        HB_TRY(dev_alloc(E, &E->mvf[k], (size_t)g.cuw * g.cuh));
        HB_TRY(dev_alloc(E, &E->satdf[k], (size_t)g.cuw * g.cuh));
    }
    HB_TRY(dev_alloc(E, &E->progress, (size_t)g.ctuh));
    HB_TRY(dev_alloc(E, &E->rc_dev, 1));
    HB_TRY(upload_initial_rc(E));
    HB_CUDA(ctx, cudaEventCreate(&E->ev_mark));
    HB_CUDA(ctx, cudaEventCreate(&E->ev_last_done));
    HB_CUDA(ctx, cudaStreamCreateWithFlags(&E->st_copy, cudaStreamNonBlocking));
    HB_CUDA(ctx, cudaStreamCreateWithFlags(&E->st_drain, cudaStreamNonBlocking));
    HB_CUDA(ctx, cudaStreamCreateWithFlags(&E->st_dl, cudaStreamNonBlocking));
    {   // CABAC launches run at the highest stream priority: their CTAs would otherwise starve behind the thousands of pending
        // CTAs of the frame chain
        int lo = 0, hi = 0;
        HB_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        for (auto &q : E->st_entropy) HB_CUDA(ctx, cudaStreamCreateWithPriority(&q, cudaStreamNonBlocking, hi));
    }
    HB_CUDA(ctx, upload_inter_constants(ctx->stream));
    E->vps = make_vps(p);
    E->sps = make_sps(p, g.wc, g.hc);
    E->pps = make_pps(p);
    HB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = e.release();
    return HB_OK;
}

void hb_enc_destroy(hb_encoder *e)
{
    if (!e) return;
    cudaSetDevice(e->ctx->device);
    cudaStreamSynchronize(e->ctx->stream);
    if (e->st_copy) { cudaStreamSynchronize(e->st_copy); cudaStreamDestroy(e->st_copy); }
    for (auto &q : e->st_entropy) if (q) { cudaStreamSynchronize(q); cudaStreamDestroy(q); }
    if (e->st_drain) { cudaStreamSynchronize(e->st_drain); cudaStreamDestroy(e->st_drain); }
    if (e->st_dl) { cudaStreamSynchronize(e->st_dl); cudaStreamDestroy(e->st_dl); }
    for (void *p : e->dev) cudaFree(p);
    for (BatchSet &B : e->set) {
        if (B.offsets_host) cudaFreeHost(B.offsets_host);
        if (B.packed_host) cudaFreeHost(B.packed_host);
        if (B.overflow_host) cudaFreeHost(B.overflow_host);
        if (B.ctl_host) cudaFreeHost(B.ctl_host);
        for (auto &ev : B.ev) if (ev) cudaEventDestroy(ev);
        for (auto &ev : B.ev_misc) if (ev) cudaEventDestroy(ev);
        for (auto &row : B.ev_grp) for (auto &ev : row) if (ev) cudaEventDestroy(ev);
        for (auto &ev : B.ev_off) if (ev) cudaEventDestroy(ev);
        for (auto &ev : B.ev_chunk) if (ev) cudaEventDestroy(ev);
        for (auto &ev : B.kev) if (ev) cudaEventDestroy(ev);
        for (auto &ev : B.kev_me) if (ev) cudaEventDestroy(ev);
    }
    if (e->ev_mark) cudaEventDestroy(e->ev_mark);
    if (e->ev_last_done) cudaEventDestroy(e->ev_last_done);
    delete e;
}

int hb_enc_reset(hb_encoder *e)
{
    if (!e) return HB_ERR_ARG;
    hb_ctx *ctx = e->ctx;
    if (e->set[0].pending || e->set[1].pending)
        return hb_fail(ctx, HB_ERR_ARG, "%s", "hb_enc_reset with frames in flight: flush hb_enc_encode_delayed first");
    HB_CUDA(ctx, cudaSetDevice(ctx->device));
    HB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    e->frame_no = 0; e->since_bp = 0; e->cur = 0;
    e->stop.store(0); e->done.store(0);
    return upload_initial_rc(e);
}

int hb_enc_headers(hb_encoder *e, uint8_t *out, size_t cap, size_t *len)
{
    if (!e) return HB_ERR_ARG;
    HB_ARG(e->ctx, out && len);
    std::vector<uint8_t> buf;
    append_nal(buf, NAL_VPS, e->vps, true);
    append_nal(buf, NAL_SPS, e->sps, true);
    append_nal(buf, NAL_PPS, e->pps, true);
    if (buf.size() > cap) return hb_fail(e->ctx, HB_ERR_SPACE, "%s", "header buffer too small");
    memcpy(out, buf.data(), buf.size());
    *len = buf.size();
    return HB_OK;
}

int hb_param_sets(const hb_enc_params *params, uint8_t *out, size_t cap, size_t *len)
{
    if (!params || !out || !len) return HB_ERR_ARG;
    const hb_enc_params &p = *params;
    if (p.width < 16 || p.height < 16 || (p.width % 2) || (p.height % 2) || p.width > 8192 || p.height > 8192) return HB_ERR_ARG;
    if (p.bit_depth != 8 && p.bit_depth != 10) return HB_ERR_ARG;
    const int wc = (p.width + 15) & ~15, hc = (p.height + 15) & ~15;
    std::vector<uint8_t> buf;
    append_nal(buf, NAL_VPS, make_vps(p), true);
    append_nal(buf, NAL_SPS, make_sps(p, wc, hc), true);
    append_nal(buf, NAL_PPS, make_pps(p), true);
    if (buf.size() > cap) return HB_ERR_SPACE;
    memcpy(out, buf.data(), buf.size());
    *len = buf.size();
    return HB_OK;
}

int hb_rc_simulate(const hb_enc_params *params, const long long *est16, const int *is_idr, int n, int *qps)
{
    if (!params || !est16 || !is_idr || !qps || n < 0) return HB_ERR_ARG;
    const hb_enc_params &p = *params;
    if (p.fps_num <= 0 || p.fps_den <= 0) return HB_ERR_ARG;
    RcState rc = initial_rc(p);       // exactly the state hb_enc_create uploads, stepped by the functions the device kernels run
    int poc = 0;
    for (int i = 0; i < n; i++) {
        poc = is_idr[i] ? 0 : poc + 1;
        qps[i] = rc_pick_qp(rc, is_idr[i] != 0, poc);
        rc_update(rc, is_idr[i] != 0, qps[i], est16[i], poc);
    }
    return HB_OK;
}

int hb_enc_coded_size(const hb_encoder *e, int *wc, int *hc)
{
    if (!e || !wc || !hc) return HB_ERR_ARG;
    *wc = e->g.wc; *hc = e->g.hc;
    return HB_OK;
}

int hb_enc_request_stop(hb_encoder *e)
{
    if (!e) return HB_ERR_ARG;
    e->stop.store(1);
    return HB_OK;
}

int hb_enc_poll_progress(const hb_encoder *e, int *frames_done)
{
    if (!e || !frames_done) return HB_ERR_ARG;
    *frames_done = e->done.load();
    return HB_OK;
}

int hb_enc_last_timing(const hb_encoder *e, float *total_ms, float *kernel_ms)
{
    if (!e) return HB_ERR_ARG;
    if (total_ms) *total_ms = e->last_total_ms;
    if (kernel_ms) *kernel_ms = e->last_kernel_ms;
    return HB_OK;
}

namespace {
int encode_common(hb_encoder *e, const hb_frames *fr, int force_idr, uint8_t *out, size_t cap, size_t *out_len, hb_frame_stat *stats,
                  int *frames_out, bool delayed)
{
    hb_ctx *ctx = e->ctx;
    HB_ARG(ctx, out && out_len);
    HB_ARG(ctx, delayed || fr);
    if (fr) {
        HB_ARG(ctx, fr->data && fr->n_frames >= 0);
        HB_ARG(ctx, fr->format >= HB_PIX_YUV420P8 && fr->format <= HB_PIX_RGB24);
        HB_ARG(ctx, (fr->src_width == 0 && fr->src_height == 0) ||
                        (fr->src_width == e->prm.width && fr->src_height == e->prm.height) ||
                        ((fr->format == HB_PIX_YUV420P8 || fr->format == HB_PIX_BGR24 || fr->format == HB_PIX_RGB24) && fr->src_width >= 16 && fr->src_height >= 16 && !(fr->src_width & 1) && !(fr->src_height & 1)));
        HB_ARG(ctx, fr->frame_bytes >= input_frame_bytes(e->prm, fr->format, fr->src_width, fr->src_height));
        HB_ARG(ctx, input_frame_bytes(e->prm, fr->format, fr->src_width, fr->src_height) <= e->staging_bytes / e->max_batch);
        HB_ARG(ctx, fr->src_bit_depth == 0 || (fr->format == HB_PIX_YUV420P16 && fr->src_bit_depth >= 8 && fr->src_bit_depth <= 16));
    }
    // the synchronous call sizes its output for its own frames: frames still in flight from the pipelined entry point must be
    // flushed first
    if (!delayed && (e->set[0].pending || e->set[1].pending))
        return hb_fail(ctx, HB_ERR_ARG, "%s", "hb_enc_encode with frames in flight: flush hb_enc_encode_delayed first");
    HB_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t written = 0;
    float total_ms = 0, kernel_ms = 0;
    int emitted = 0;
    e->done.store(0);
    const int n_frames = fr ? fr->n_frames : 0;
    for (int base = 0; base < n_frames; base += e->max_batch) {
        const int n = std::min(e->max_batch, n_frames - base);
        BatchSet &B = e->set[e->next_set];
        BatchSet &prev = e->set[e->next_set ^ 1];
        HB_TRY(enqueue_batch(e, B, fr, base, n, force_idr && base == 0));
        e->next_set ^= 1;
        if (prev.pending) {      // drain the batch before this one while the GPU works on this one
            const int pn = prev.n;
            HB_TRY(drain_batch(e, prev, out, cap, written, stats ? stats + emitted : nullptr, total_ms, kernel_ms));
            emitted += pn;
        }
    }
    if (!delayed || !fr) {       // synchronous call, or flush: nothing stays in flight
        for (int k = 0; k < 2; k++) {
            BatchSet &B = e->set[(e->next_set + k) & 1];      // older batch first
            if (!B.pending) continue;
            const int pn = B.n;
            HB_TRY(drain_batch(e, B, out, cap, written, stats ? stats + emitted : nullptr, total_ms, kernel_ms));
            emitted += pn;
        }
    }
    e->last_total_ms = total_ms;
    e->last_kernel_ms = kernel_ms;
    *out_len = written;
    if (frames_out) *frames_out = emitted;
    return HB_OK;
}
}  // namespace

int hb_enc_encode(hb_encoder *e, const hb_frames *fr, int force_idr, uint8_t *out, size_t cap, size_t *out_len, hb_frame_stat *stats)
{
    if (!e) return HB_ERR_ARG;
    return encode_common(e, fr, force_idr, out, cap, out_len, stats, nullptr, false);
}

int hb_enc_encode_delayed(hb_encoder *e, const hb_frames *fr, int force_idr, uint8_t *out, size_t cap, size_t *out_len, hb_frame_stat *stats,
                          int *frames_out)
{
    if (!e) return HB_ERR_ARG;
    return encode_common(e, fr, force_idr, out, cap, out_len, stats, frames_out, true);
}

size_t hb_escape_rbsp(const uint8_t *in, size_t n, uint8_t *out, size_t cap)
{
    if ((!in && n) || !out) return 0;
    std::vector<uint8_t> buf;
    buf.reserve(n + n / 64 + 8);
    append_escaped(buf, in, n);
    if (buf.size() > cap) return 0;
    if (!buf.empty()) memcpy(out, buf.data(), buf.size());
    return buf.size();
}

int hb_enc_mark(hb_encoder *e)
{
    if (!e) return HB_ERR_ARG;
    HB_CUDA(e->ctx, cudaSetDevice(e->ctx->device));
    HB_CUDA(e->ctx, cudaEventRecord(e->ev_mark, e->ctx->stream));
    e->have_last_done = false;
    e->have_mark = true;
    return HB_OK;
}

int hb_enc_elapsed(hb_encoder *e, float *ms)
{
    if (!e || !ms) return HB_ERR_ARG;
    HB_ARG(e->ctx, e->have_last_done && e->have_mark);      // (an unrecorded event would leave a sticky CUDA error behind)
    HB_CUDA(e->ctx, cudaEventSynchronize(e->ev_last_done));
    HB_CUDA(e->ctx, cudaEventElapsedTime(ms, e->ev_mark, e->ev_last_done));
    return HB_OK;
}

int hb_enc_profile(hb_encoder *e, int enable, float ms[8], int launches[8])
{
    if (!e) return HB_ERR_ARG;
    if (ms) memcpy(ms, e->prof_ms, sizeof(e->prof_ms));
    if (launches) memcpy(launches, e->prof_launches, sizeof(e->prof_launches));
    if (enable >= 0) {
        e->profiling = enable;
        memset(e->prof_ms, 0, sizeof(e->prof_ms));
        memset(e->prof_launches, 0, sizeof(e->prof_launches));
    }
    return HB_OK;
}

int hb_enc_read_recon(hb_encoder *e, int i, uint16_t *y, uint16_t *u, uint16_t *v)
{
    if (!e) return HB_ERR_ARG;
    hb_ctx *ctx = e->ctx;
    HB_ARG(ctx, (e->prm.keep_recon || e->prm.hash_sei) && e->last_drained && i >= 0 && i < e->max_batch && y && u && v);
    const Geom &g = e->g;
    HB_CUDA(ctx, cudaSetDevice(ctx->device));
    HB_CUDA(ctx, cudaMemcpy(y, e->last_drained->slot[i].keep.y, (size_t)g.wc * g.hc * 2, cudaMemcpyDeviceToHost));
    HB_CUDA(ctx, cudaMemcpy(u, e->last_drained->slot[i].keep.u, (size_t)g.wc * g.hc / 2, cudaMemcpyDeviceToHost));
    HB_CUDA(ctx, cudaMemcpy(v, e->last_drained->slot[i].keep.v, (size_t)g.wc * g.hc / 2, cudaMemcpyDeviceToHost));
    return HB_OK;
}

int hb_enc_read_decisions(hb_encoder *e, int i, void *cus, int16_t *coefs)
{
    if (!e) return HB_ERR_ARG;
    hb_ctx *ctx = e->ctx;
    HB_ARG(ctx, e->last_drained && i >= 0 && i < e->max_batch);
    const size_t ncu = (size_t)e->g.cuw * e->g.cuh;
    HB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (cus) HB_CUDA(ctx, cudaMemcpy(cus, e->last_drained->slot[i].cus, ncu * sizeof(CuInfo), cudaMemcpyDeviceToHost));
    if (coefs) HB_CUDA(ctx, cudaMemcpy(coefs, e->last_drained->slot[i].coefs, ncu * kCuCoefs * sizeof(int16_t), cudaMemcpyDeviceToHost));
    return HB_OK;
}

}  // extern "C"
