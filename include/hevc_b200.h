/* hevc_b200 -- C ABI of the B200-native HEVC encode backend.
 *
 * This is the drop-in boundary for the encode step of the reference pipeline.  The reference has no FFI:
 * its "operator API" to the codec is the ffmpeg argv built at core/transcoder.py:452-495 and executed at
 * core/transcoder.py:497-535 (run_ffmpeg -> subprocess.Popen).  Each entry point below cites the part of
 * that child process it replaces.  Plain C types only; device pointers are CUDA device addresses of the
 * calling process (any allocator: cudaMalloc, torch, ...).  Every function returns 0 on success and a
 * negative hb_status on failure; hb_last_error() gives the text.  One hb_ctx owns one device + one
 * CUDA stream; use one per worker thread (reference threading model: gui/mainwindow.py:289-301).
 */
#ifndef HEVC_B200_H
#define HEVC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_ABI_VERSION 1

typedef enum hb_status {
    HB_OK = 0,
    HB_ERR_ARG = -1,       /* bad argument */
    HB_ERR_CUDA = -2,      /* CUDA runtime / driver failure */
    HB_ERR_NOMEM = -3,
    HB_ERR_SPACE = -4,     /* caller buffer too small */
    HB_ERR_STOPPED = -5,   /* hb_enc_request_stop() was honoured */
    HB_ERR_STATE = -6
} hb_status;

typedef struct hb_ctx hb_ctx;
typedef uint64_t hb_devptr;

int hb_abi_version(void);
int hb_device_count(void);
/* create a context on `device` with its own non-blocking stream */
int hb_create(int device, hb_ctx **out);
void hb_destroy(hb_ctx *ctx);
const char *hb_last_error(const hb_ctx *ctx);
int hb_sync(hb_ctx *ctx);
/* the context's cudaStream_t, for callers that record their own CUDA events around launches */
uint64_t hb_stream(const hb_ctx *ctx);
/* number of kernels this context has launched since creation */
uint64_t hb_launch_count(const hb_ctx *ctx);
/* device scratch allocation owned by the context (freed by hb_free or hb_destroy) */
int hb_alloc(hb_ctx *ctx, size_t bytes, hb_devptr *out);
int hb_free(hb_ctx *ctx, hb_devptr p);
/* page-locked host memory for frame buffers (cudaMallocHost): uploads from it run at full PCIe rate and truly asynchronously,
 * which is what lets the reader fill batch k+1 while batch k is still being copied */
int hb_host_alloc(size_t bytes, void **out);
int hb_host_free(void *p);
int hb_upload(hb_ctx *ctx, hb_devptr dst, const void *src, size_t bytes);     /* async on the ctx stream */
int hb_download(hb_ctx *ctx, void *dst, hb_devptr src, size_t bytes);         /* async on the ctx stream */
/* event timing on the context's stream: start/stop bracket launches, elapsed is in milliseconds */
int hb_timer_start(hb_ctx *ctx);
int hb_timer_stop(hb_ctx *ctx, float *ms);

/* ---------------------------------------------------------------------------------------------------------
 * Pre-encode pixel pipeline.  Replaces the libswscale stage ffmpeg auto-inserts for `-pix_fmt`
 * (reference core/transcoder.py:464; p010le requested at :364) and the geometry of the upscale path
 * (upscale_gui_final.py:81-87).  Strides are in BYTES.  All pointers are device pointers.
 * ------------------------------------------------------------------------------------------------------- */

/* 8-bit planar 4:2:0 -> P010 (16-bit containers, value << 8, UV interleaved).  4.5 bytes / pixel. */
int hb_pack_p010(hb_ctx *ctx, hb_devptr y, int y_stride, hb_devptr u, int u_stride, hb_devptr v, int v_stride,
                 int width, int height, hb_devptr dst_y, int dst_y_stride, hb_devptr dst_uv, int dst_uv_stride);

typedef enum hb_matrix { HB_MATRIX_BT709 = 1, HB_MATRIX_BT601 = 6, HB_MATRIX_BT2020 = 9 } hb_matrix;
typedef enum hb_rgb_order { HB_RGB = 0, HB_BGR = 1 } hb_rgb_order;

/* packed 8-bit full-range RGB/BGR -> limited-range 4:2:0, Q14 matrix, 2x2 box chroma.
 * depth 8: dst_y/dst_u/dst_v are 8-bit planes.  depth 10: dst_y is a P010 luma plane, dst_u a P010
 * interleaved UV plane (dst_v ignored).  6 bytes / pixel at depth 10. */
int hb_rgb_to_yuv420(hb_ctx *ctx, hb_devptr rgb, int rgb_stride, int order, int matrix, int depth, int width, int height,
                     hb_devptr dst_y, int dst_y_stride, hb_devptr dst_u, int dst_u_stride, hb_devptr dst_v, int dst_v_stride);

/* 4-tap Catmull-Rom polyphase scaler (64 phases, Q14 taps, Q6 int16 intermediate) of one 8-bit plane.
 * out_depth 8 -> 8-bit plane; out_depth 10 -> 16-bit samples, shifted left by `out_shift` (6 = P010).
 * dst_pixel_step = distance between successive output samples in SAMPLES (2 to write one half of an
 * interleaved UV plane). */
int hb_scale_plane(hb_ctx *ctx, hb_devptr src, int src_stride, int src_w, int src_h,
                   hb_devptr dst, int dst_stride, int dst_w, int dst_h, int out_depth, int out_shift, int dst_pixel_step);

/* fused upscale path: 8-bit planar 4:2:0 (src_w x src_h) -> P010 (dst_w x dst_h), three hb_scale_plane passes */
int hb_scale_yuv420_to_p010(hb_ctx *ctx, hb_devptr y, int y_stride, hb_devptr u, int u_stride, hb_devptr v, int v_stride,
                            int src_w, int src_h, hb_devptr dst_y, int dst_y_stride, hb_devptr dst_uv, int dst_uv_stride,
                            int dst_w, int dst_h);

/* Batched forms: ONE launch converts n_frames consecutive, tightly packed frames (source side: the hb_frames layout -- Y plane,
 * U plane, V plane, or packed RGB; destination side: P010 frames -- Y plane, then the interleaved UV plane).  The byte distance
 * between successive frames is given for both sides.  This is how a batch worker feeds whole read batches to the device. */
int hb_pack_p010_batch(hb_ctx *ctx, hb_devptr src, size_t src_frame_bytes, hb_devptr dst, size_t dst_frame_bytes, int width, int height,
                       int n_frames);
int hb_rgb_to_p010_batch(hb_ctx *ctx, hb_devptr rgb, size_t rgb_frame_bytes, int order, int matrix, int width, int height, hb_devptr dst,
                         size_t dst_frame_bytes, int n_frames);
int hb_scale_yuv420_to_p010_batch(hb_ctx *ctx, hb_devptr src, size_t src_frame_bytes, int src_w, int src_h, hb_devptr dst,
                                  size_t dst_frame_bytes, int dst_w, int dst_h, int n_frames);

/* ---------------------------------------------------------------------------------------------------------
 * Encoder primitives, batched over blocks (BASELINE config 5).  These are the functions libx265 spends
 * its time in behind `-c:v libx265` (reference core/transcoder.py:398-412): x265 primitives sad / satd /
 * sa8d / dct / idct / dst / nquant / dequant_normal / intra_pred.  Samples are uint16_t for every bit
 * depth, coefficients int16_t.  Blocks are dense: block i of a WxH batch starts at i*W*H samples.
 * ------------------------------------------------------------------------------------------------------- */

/* out[i] = SAD / SATD of block i of `a` against block i of `b`; out is int32[n] */
int hb_sad(hb_ctx *ctx, hb_devptr a, hb_devptr b, int n, int w, int h, hb_devptr out);
/* x265 sad_x3 / sad_x4: one source block against n_refs (3 or 4) reference blocks, the source read once; out int32[n][n_refs] */
int hb_sad_multi(hb_ctx *ctx, hb_devptr fenc, hb_devptr ref0, hb_devptr ref1, hb_devptr ref2, hb_devptr ref3, int n_refs, int n,
                 int w, int h, hb_devptr out);
int hb_satd(hb_ctx *ctx, hb_devptr a, hb_devptr b, int n, int w, int h, hb_devptr out);
/* square blocks, size 4 (== satd 4x4), 8, 16, 32, 64 */
int hb_sa8d(hb_ctx *ctx, hb_devptr a, hb_devptr b, int n, int size, hb_devptr out);
/* forward / inverse core transform, size 4/8/16/32; is_dst selects DST-VII (size 4 only) */
int hb_fwd_transform(hb_ctx *ctx, hb_devptr residual, int n, int size, int bit_depth, int is_dst, hb_devptr coef);
int hb_inv_transform(hb_ctx *ctx, hb_devptr coef, int n, int size, int bit_depth, int is_dst, hb_devptr residual);
/* flat-matrix quantisation without RDOQ (x265 nquant): qp already includes QpBdOffset; numsig int32[n] */
int hb_quant(hb_ctx *ctx, hb_devptr coef, int n, int size, int qp, int bit_depth, int is_intra, hb_devptr level, hb_devptr numsig);
int hb_dequant(hb_ctx *ctx, hb_devptr level, int n, int size, int qp, int bit_depth, hb_devptr coef);
/* all 35 intra predictions per block.  neighbours: uint16[n][4*size+1] = {top-left, top[2N], left[2N]};
 * pred: uint16[n][35][size*size].  is_luma enables reference smoothing + DC/H/V edge filters. */
int hb_intra_pred_all(hb_ctx *ctx, hb_devptr neighbours, int n, int size, int is_luma, int strong_smoothing,
                      int bit_depth, hb_devptr pred);


/* ---------------------------------------------------------------------------------------------------------
 * Stream encoder.  Replaces `ffmpeg ... -c:v libx265 -x265-params ...` (argv built at reference
 * core/transcoder.py:398-412,452-495, executed by run_ffmpeg :497-535).  hb_enc_params carries what that
 * argv carries: the outputs of calculate_apple_hevc_level (:174), calculate_dynamic_values (:263) and
 * build_hdr_metadata (core/utils.py:29).  Output is an Annex-B elementary stream (the mov muxing the
 * reference leaves to ffmpeg is done by the host package, hevc_b200/mp4.py).
 * ------------------------------------------------------------------------------------------------------- */
typedef struct hb_enc_params {
    int width, height;                 /* display size, even */
    int fps_num, fps_den;
    int bit_depth;                     /* 8 = Main, 10 = Main10 */
    int profile_idc, level_idc, tier;  /* level_idc = 30 * level */
    int qp_i, qp_p;                    /* base QPs derived from crf= by the host (hevc_b200/encoder.py) */
    int keyint, min_keyint;
    int vbv_maxrate_kbps, vbv_bufsize_kbit;
    int colour_primaries, transfer_characteristics, matrix_coeffs;
    int vui_colour;                    /* signal the colour description */
    int chroma_loc;                    /* chromaloc=; -1 = not signalled */
    int full_range;                    /* -color_range tv -> 0 */
    int aud, repeat_headers, hrd, hdr10;
    uint32_t master_display[10];       /* Gx,Gy,Bx,By,Rx,Ry,WPx,WPy (0.00002), Lmax,Lmin (0.0001 cd/m2) */
    int max_cll, max_fall;
    int hash_sei;                      /* verification: emit MD5 decoded-picture-hash SEI (reads back every reconstruction) */
    int keep_recon;                    /* verification: keep every reconstruction / decision of the last batch readable */
    int rate_control;                  /* 0 = constant QP, 1 = VBV-constrained */
    int deblock;                       /* in-loop deblocking filter (x265 default: on) */
    int scenecut;                      /* key frame at a detected scene cut once min_keyint frames have passed (x265 default: on;
                                        * min-keyint= from core/transcoder.py:405) */
    int intra_in_p;                    /* intra CUs in P frames */
    int sao;                           /* sample adaptive offset (x265 default: on) */
    int qp_cascade;                    /* P-frame QP cascade: qp_p applies to every fourth P frame (poc % 4 == 0), +4 / +2 / +4 between.
                                          In place of the P / B QP ratio of x265 (pbratio): no B frames here */
    int reserved[3];
} hb_enc_params;

typedef enum hb_pix_fmt {
    HB_PIX_YUV420P8 = 0, HB_PIX_P010 = 1, HB_PIX_YUV420P16 = 2,
    HB_PIX_BGR24 = 3, HB_PIX_RGB24 = 4     /* packed 8-bit full-range RGB: converted by the ingest stage (what swscale does for -pix_fmt) */
} hb_pix_fmt;

/* a run of consecutive frames, tightly packed per frame: Y plane, then U and V planes (P010: Y then UV) */
typedef struct hb_frames {
    const void *data;                  /* host pointer, or device pointer when on_device != 0 */
    int on_device;
    int format;                        /* hb_pix_fmt */
    int n_frames;
    int src_bit_depth;                 /* HB_PIX_YUV420P16 only: significant bits per sample (8..16, LSB-aligned); 0 = the
                                        * encoder's bit depth.  Deeper sources are rounded down to the encoder's depth, shallower
                                        * ones shifted up (what swscale does for -pix_fmt, core/transcoder.py:464) */
    size_t frame_bytes;                /* distance between successive frames */
    int src_width, src_height;         /* 0, 0 = the encoder's display size.  Other sizes (HB_PIX_YUV420P8 and the RGB formats): every frame is
                                        * resampled on the device by the polyphase scaler straight into the encoder's source
                                        * planes -- the upscale path (reference upscale_gui_final.py:81-87) without a P010 round trip */
    int matrix;                        /* HB_PIX_BGR24 / RGB24: HB_MATRIX_*; 0 = follow params.matrix_coeffs (1 -> BT.709, 9 -> BT.2020, anything else -> BT.601) */
    int reserved0;
} hb_frames;

typedef struct hb_frame_stat {
    int is_idr, poc, qp;
    uint32_t bytes;                    /* size of the access unit */
    uint32_t n_skip, n_merge;
} hb_frame_stat;

typedef struct hb_encoder hb_encoder;

int hb_enc_create(hb_ctx *ctx, const hb_enc_params *params, int max_batch, hb_encoder **out);
void hb_enc_destroy(hb_encoder *enc);
/* Start a new stream on an existing encoder (same parameters): frame counter, POC, rate-control state and SEI timing return to
 * their initial values, so the next frame is the first key frame of an independent stream.  Nothing may be in flight.  Lets a
 * batch worker reuse one encoder (and its device memory) for successive files of the same geometry instead of paying the
 * allocation per file (reference: one ffmpeg child per file, gui/worker.py:30-41). */
int hb_enc_reset(hb_encoder *enc);
/* VPS + SPS + PPS as Annex-B (for the hvcC box) */
int hb_enc_headers(hb_encoder *enc, uint8_t *out, size_t cap, size_t *len);
int hb_enc_coded_size(const hb_encoder *enc, int *coded_w, int *coded_h);
/* The same parameter sets without an encoder or a device (pure host code): what a muxer needs before the first frame exists. */
int hb_param_sets(const hb_enc_params *params, uint8_t *out, size_t cap, size_t *len);
/* Rate controller on the host (pure host code, the same functions the device kernels k_rc_step run): QP of each of n frames
 * given their types and size estimates in 1/16 bit.  Test hook: lets the controller be checked against the oracle without a GPU. */
int hb_rc_simulate(const hb_enc_params *params, const long long *est16, const int *is_idr, int n, int *qps);
/* Encode frames->n_frames consecutive frames, continuing the stream (closed GOP of params.keyint frames;
 * force_idr restarts the GOP at the first frame: closed-GOP segment sharding).  Appends one access unit per
 * frame to out; stats (optional) receives one entry per frame.  The call returns when the bytes are in out. */
int hb_enc_encode(hb_encoder *enc, const hb_frames *frames, int force_idr, uint8_t *out, size_t cap, size_t *out_len,
                  hb_frame_stat *stats);
/* Pipelined form (the way libx265's x265_encoder_encode delivers its output with a delay): the frames are enqueued and the
 * call returns the access units of the frames submitted by EARLIER calls that were still in flight, so that the frame
 * chain of this batch overlaps the CABAC tail, download and access-unit assembly of the previous one.  frames == NULL
 * flushes: everything still in flight is returned.  *frames_out = number of access units written (stats likewise).
 * Host input buffers must stay valid until the frames have been returned. */
int hb_enc_encode_delayed(hb_encoder *enc, const hb_frames *frames, int force_idr, uint8_t *out, size_t cap, size_t *out_len,
                          hb_frame_stat *stats, int *frames_out);
/* region timing across pipelined calls: hb_enc_mark records a CUDA event on the encoder's stream; hb_enc_elapsed returns the
 * device time from that mark to the end of the last completed bitstream download */
int hb_enc_mark(hb_encoder *enc);
int hb_enc_elapsed(hb_encoder *enc, float *ms);
/* Host-only helper (no device needed): emulation prevention of one RBSP / sub-stream exactly as the access-unit assembly applies
 * it (H.265 7.4.2: 0x03 before any byte <= 3 that follows two zero bytes).  Returns the escaped size, or 0 if cap is too small. */
size_t hb_escape_rbsp(const uint8_t *in, size_t n, uint8_t *out, size_t cap);
/* device time of the last hb_enc_encode call, measured with CUDA events on the encoder's stream:
 * total (first upload to last download) and kernels only */
int hb_enc_last_timing(const hb_encoder *enc, float *total_ms, float *kernel_ms);
/* per-kernel-class device time (CUDA events around each launch on the encoder's stream), accumulated since the
 * last call: index 0 inter-frame kernel, 1 intra-frame kernel, 2 coarse search, 3 entropy + compaction, 4 ingest,
 * 5 whole frame chain (inter/intra + border + mode kernels), 6 the motion-search kernel k_me alone, 7 unused.  Index 0 spans
 * the three inter kernels (k_me, k_merge, k_inter) of a P frame.  enable: 1/0 switches the event recording and resets the
 * accumulators, -1 only reads. */
int hb_enc_profile(hb_encoder *enc, int enable, float ms[8], int launches[8]);
/* cooperative cancel (reference stop_event, core/transcoder.py:511-516): polled between frames */
int hb_enc_request_stop(hb_encoder *enc);
int hb_enc_poll_progress(const hb_encoder *enc, int *frames_done);
/* verification access (needs keep_recon): frame i of the last batch, coded size, 16-bit samples */
int hb_enc_read_recon(hb_encoder *enc, int i, uint16_t *y, uint16_t *u, uint16_t *v);
/* per-CU decisions {pred_mode,intra_mode,cbf,skip,mvx,mvy} (8 bytes each) and levels (384 int16 each) of frame i */
int hb_enc_read_decisions(hb_encoder *enc, int i, void *cus, int16_t *coefs);

#ifdef __cplusplus
}
#endif
#endif /* HEVC_B200_H */
