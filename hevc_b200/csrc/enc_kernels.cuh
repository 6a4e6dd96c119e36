// Kernel parameter blocks shared between the encoder kernels (enc_frame.cu, enc_entropy.cu) and the host
// driver (enc_host.cu).
#pragma once
#include <cstddef>
#include <cuda.h>

#include "enc_dev.cuh"

struct hb_ctx;

enum { HB_FMT_YUV420P8 = 0, HB_FMT_P010 = 1, HB_FMT_YUV420P16 = 2, HB_FMT_BGR24 = 3, HB_FMT_RGB24 = 4 };

namespace hb {

struct IngestParams {
    Geom g;
    const uint8_t *in_y, *in_u, *in_v;   // P010: in_u is the interleaved UV plane
    int in_ys, in_us, in_vs;             // bytes
    int fmt, w, h;                       // display size
    int up_shift, down_shift;
    int round_add, maxv;                 // added before the down shift (16-bit planar sources deeper than the encoder), sample clamp
    Planes src;
    uint8_t *ds;                         // quarter-resolution plane, 8 most significant bits
};

struct CoarseParams {
    Geom g;
    const uint8_t *ds;                   // [slots][dsh * dsw]; frame f of the batch = slot f + 1, slot f = its predecessor
    size_t ds_frame_stride;
    int16_t *cmv;                        // [frames][ctuh * ctuw][2]
    SceneStat *scene;                    // [frames] scene-cut measures, zeroed before the launch
};

struct InterParams {
    CUtensorMap ref_map;                 // tiled TMA descriptor of the padded luma reference plane: box = 40 x 28 samples (the search window)
    Geom g;
    Planes src, ref, rec;
    const int16_t *cmv;                  // this frame's coarse vectors
    CuInfo *cus;
    int16_t *coefs;
    FrameCtl *ctl;                       // QP / lambda / quantisers chosen on the device; receives the size estimate
    const uint32_t *mv_in;               // [cuh][cuw] motion field of the previous pass (x | y << 16, quarter samples) ...
    const int *satd_in;                  // ... and the luma SATD of each vector
    uint32_t *mv_out;
    int *satd_out;
    const int *intra_best;               // [cus] best intra SATD of each CU (k_intra_search), or null: no intra CUs in P frames
};

struct IntraParams {
    Geom g;
    Planes src, rec;
    CuInfo *cus;
    int16_t *coefs;
    int *progress;                       // [ctuh], zeroed before launch
    FrameCtl *ctl;
    int *mode_cost;                      // [cus][35] luma SATD per intra mode from k_intra_search
    int *intra_best;                     // [cus] min over the 35 modes (INT_MAX: not searched)
    const int *satd1;                    // [cus] luma SATD after the first merge-aware pass: gate of the P-frame search
    int *cand_list;                      // [cus] work list of the intra search (k_intra_list -> k_intra_search)
    int second_pass;                     // 1: run only when ctl->redo is set (first key frame of a stream under rate control)
    int intra_in_p;                      // P frames: search every CU and reconstruct the CUs the inter kernel marked intra
};

struct DeblockParams {
    Geom g;
    Planes rec;
    const CuInfo *cus;
    const FrameCtl *ctl;
    int dir;                             // 0 vertical edges, 1 horizontal edges
};

struct SaoParams {
    Geom g;
    Planes src, pre, out;                // source, deblocked reconstruction (padded strides), final reconstruction (padded strides)
    const CuInfo *cus;
    const FrameCtl *ctl;
    SaoCtu *sao;                         // [ctuh * ctuw]
};

struct ModeParams {
    Geom g;
    const CuInfo *cus;
    CuSyntax *syn;
    const FrameCtl *ctl;                 // slice type decided on the device
};

struct EntropyFrame {
    const CuInfo *cus;
    const CuSyntax *syn;
    const int16_t *coefs;
    uint8_t *out;                        // [ctuh][row_cap]
    uint32_t *row_len;                   // [ctuh]
    const FrameCtl *ctl;                 // slice QP (context initialisation)
    const SaoCtu *sao;                   // [ctuh * ctuw] SAO parameters, or null: SAO disabled
    uint8_t *ctx_save;                   // [ctuh][kNumCtx] WPP context snapshots (after the 2nd CTU of each row)
    int *row_ready;                      // [ctuh] snapshot-published flags, zeroed before the launch
    unsigned long long *trace;           // debug (HB_ENTROPY_TRACE): [ctuh][2] globaltimer at row start (after the hand-off) / end, or null
};

struct EntropyParams {
    Geom g;
    const EntropyFrame *frames;          // device array, one per CTA
    uint32_t row_cap;
    int *overflow;
};

struct PackParams {
    const EntropyFrame *frames;
    int n_frames, rows;
    uint32_t row_cap;
    uint8_t *packed;                     // contiguous: frame 0 rows, frame 1 rows, ...
    uint32_t *offsets;                   // [n_frames * rows + 1] exclusive prefix of row lengths
    unsigned long long cap;              // bytes reserved at `packed` for this group
};

constexpr int kEntropyWarps = 4;
constexpr int kTraceWords = 5;             // debug trace per CTU row: start ns, end ns, binarisation cycles, coding cycles, list entries
constexpr int kIntraSearchThreads = 128;   // k_intra_search CTA: four warps, one CU each
constexpr int kIntraReconThreads = 384;   // k_intra CTA: one thread per luma sample + one per chroma sample (both planes) of the CU
constexpr int kBinStride = 61;            // 32-bit words per sub-block bin list (<= 60 entries, odd stride: no bank conflicts)
constexpr int kHdrBins = 40, kTuHdrBins = 24;
struct CuStage {
    CuInfo info;
    CuSyntax syn;
};
// Per-warp (= per CTU row) working set of the CABAC kernel.  Bin lists are written by the binarisation lanes and consumed by
// lane 0; an entry is kind (bits 0-7: context index, or one of the kBypass / kUnary / kTerminate codes), argument (bits 8-15:
// bin value or bin count) and bypass bits (16-31).
struct __align__(16) EntropyWarpScratch {
    int16_t lv[2][kCuCoefs];             // levels of the CU being coded / of the next coded CU (cp.async, one coded CU ahead)
    uint2 ctx[kNumCtx + 2];              // x: the four rangeTabLps bytes of the context's state; y: state << 1 | mps, next-LPS state << 8
    CuStage cu[3][4];                    // ring of staged CTUs (cp.async, two CTUs ahead); 8-byte aligned for the copies
    CuSyntax above[3][2];                // syntax of the two CUs above each staged CTU (skip-flag context)
    SaoCtu sao_cur[3], sao_up[3];        // SAO parameters of the staged CTUs and of the CTUs above them (cp.async)
    SaoCtu sao_left;                     // ... of the CTU coded last in this row
    uint32_t saob[52];                   // sao() bin list of the current CTU
    uint32_t nsao, pad3[3];
    // byte-output side of the arithmetic coder (touched only when a byte leaves it)
    uint8_t *out;
    uint32_t pos, cap;
    int buffered;
    uint32_t held;
    uint32_t bins[24][kBinStride];       // per sub-block bin lists
    uint32_t hdr[4][kHdrBins];           // per CU: syntax up to the coded block flags
    uint32_t tuh[3][kTuHdrBins];         // per transform block: last significant coefficient position
    uint16_t masks[24];                  // per sub-block significance masks in diagonal scan order, raster sub-block index
    uint8_t nbins[24];
    uint8_t nhdr[4], ntuh[3], pad0;
    int8_t last_sb[3], pad1;
    uint32_t pad2;
};
static_assert(offsetof(EntropyWarpScratch, cu) % 8 == 0 && offsetof(EntropyWarpScratch, above) % 8 == 0, "cp.async destinations");
static_assert(offsetof(EntropyWarpScratch, sao_cur) % 16 == 0 && offsetof(EntropyWarpScratch, sao_up) % 16 == 0, "cp.async destinations");
static_assert(sizeof(EntropyWarpScratch) % 16 == 0, "per-warp scratch must keep 16-byte alignment");
// uploads the dp2a-packed interpolation taps into constant memory (call once per process/device before k_inter)
cudaError_t upload_inter_constants(cudaStream_t st);

__global__ void k_ingest(IngestParams p);
__global__ void k_pad_ds(Geom g, Planes src, int w, int h, uint8_t *ds);
__global__ void k_border(Planes rec, Geom g);
__global__ void k_coarse(CoarseParams p);
__global__ void k_me(const __grid_constant__ InterParams p);
__global__ void k_merge(const __grid_constant__ InterParams p);
__global__ void k_inter(const __grid_constant__ InterParams p);
__global__ void k_intra_list(IntraParams p);
__global__ void k_intra_search(IntraParams p);
__global__ void k_intra(IntraParams p);
__global__ void k_deblock(DeblockParams p);
__global__ void k_modes(ModeParams p);
__global__ void k_sao_decide(SaoParams p);
__global__ void k_sao_apply(SaoParams p);
__global__ void k_rc_step(RcState *rc, FrameCtl *done, FrameCtl *next, int force_idr, const SceneStat *scene, long long ds_samples);
__global__ void k_rc_redo(RcState *rc, FrameCtl *ctl);
__global__ void k_entropy(EntropyParams p);
__global__ void k_pack_scan(PackParams p);
__global__ void k_pack_copy(PackParams p);


// pixel.cu: pre-encode pixel pipeline launchers the ingest stage uses (run on ctx->stream)
int launch_scale8(hb_ctx *ctx, const uint8_t *s0, const uint8_t *s1, int ss, int sw, int sh, uint8_t *d0, uint8_t *d1, int ds, int dw, int dh,
                  int out_depth, int out_shift, int out_mode, int n_frames = 1, size_t in_fs = 0, size_t out_fs = 0);
int launch_rgb_planar16(hb_ctx *ctx, const uint8_t *rgb, int rs, int bgr, int matrix, int depth, int w, int h, uint8_t *dy, int dys, uint8_t *du,
                        uint8_t *dv, int dcs);

}  // namespace hb
