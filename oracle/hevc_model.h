/* TEST INFRASTRUCTURE -- CPU oracle of the B200 HEVC encoder (not on the product path).
 *
 * A plain-C restatement of the encode step the reference delegates to `ffmpeg -c:v libx265`
 * (core/transcoder.py:398-412,506).  libx265 itself is absent and un-pinned, so this model defines the
 * encoder algorithm the CUDA path implements (DESIGN.md "encoder specification") and is itself pinned by the
 * normative side: every stream it writes must decode under the FFmpeg hevc decoder to exactly its own
 * reconstruction (tests/test_oracle_encoder.py).  The CUDA encoder is then required to be byte-identical to
 * this model.
 */
#ifndef ORC_HEVC_MODEL_H
#define ORC_HEVC_MODEL_H
#include <stddef.h>
#include <stdint.h>

typedef uint16_t pixel;

typedef struct orc_enc_params {
    int width, height;            /* display size (even) */
    int fps_num, fps_den;
    int bit_depth;                /* 8 or 10 */
    int profile_idc, level_idc, tier;
    int qp_i, qp_p;               /* constant QPs per slice type */
    int keyint;
    int colour_primaries, transfer_characteristics, matrix_coeffs;
    int vui_colour;               /* write colour description */
    int chroma_loc;               /* -1: not signalled */
    int full_range;
    int aud, repeat_headers, hrd, hdr10;
    int vbv_maxrate_kbps, vbv_bufsize_kbit;
    uint32_t master_display[10];  /* Gx,Gy,Bx,By,Rx,Ry,WPx,WPy,Lmax,Lmin */
    int max_cll, max_fall;
    int hash_sei;                 /* emit decoded-picture-hash (MD5) suffix SEI */
    int deblock;                  /* in-loop deblocking filter enabled */
    int rate_control;             /* 0 constant QP, 1 VBV-constrained (vbv_maxrate / vbv_bufsize) */
    int min_keyint;               /* min-keyint= (reference core/transcoder.py:405): a scene cut closer than this to the last IDR stays a P frame */
    int scenecut;                 /* scene-cut detection: key frame at a detected cut (x265 default: on) */
    int intra_in_p;               /* intra CUs in P frames */
    int sao;                      /* sample adaptive offset */
    int qp_cascade;               /* P-frame QP cascade (hevc_rc.c): qp_p applies to every fourth P frame, +4 / +2 / +4 between */
} orc_enc_params;

/* per-CU side information, also the interface between the decide/reconstruct stage and the entropy stage */
typedef struct orc_cu {
    uint8_t pred_mode;            /* 0 intra, 1 inter */
    uint8_t intra_mode;           /* luma mode 0..34 */
    uint8_t cbf;                  /* bit0 Y, bit1 Cb, bit2 Cr */
    uint8_t skip;                 /* filled by the entropy stage */
    int16_t mvx, mvy;             /* quarter-sample units */
} orc_cu;

/* sample adaptive offset parameters of one CTU (7.3.8.3): per component group type 0 off / 1 band / 2 edge */
typedef struct orc_sao {
    uint8_t type[2];              /* [0] luma, [1] chroma (Cb and Cr share type and edge class) */
    uint8_t eo_class[2];
    uint8_t band[3];              /* band position per component */
    int8_t offset[3][4];          /* signed offsets per component (edge: categories 1..4; band: the four bands) */
    uint8_t pad;
} orc_sao;

#define ORC_CU_COEFS 384          /* 16x16 luma + 8x8 Cb + 8x8 Cr, raster inside each block */
#define ORC_PAD 80                /* luma border of reconstructed planes */

/* ---- bit writer / NAL layer (hevc_bits.c) */
typedef struct orc_bits {
    uint8_t *buf;
    size_t cap, pos;              /* pos in bytes */
    uint32_t cur;
    int nbits;                    /* bits held in cur */
    int overflow;
} orc_bits;

void orc_bits_init(orc_bits *b, uint8_t *buf, size_t cap);
void orc_put(orc_bits *b, uint32_t v, int n);
void orc_put_ue(orc_bits *b, uint32_t v);
void orc_put_se(orc_bits *b, int v);
void orc_trailing(orc_bits *b);                 /* rbsp_trailing_bits / byte_alignment */
size_t orc_bits_flush(orc_bits *b);
/* start code + NAL header + emulation prevention over `rbsp` */
size_t orc_write_nal(uint8_t *out, size_t cap, int nal_type, const uint8_t *rbsp, size_t n, int long_start);
size_t orc_escape(uint8_t *out, size_t cap, const uint8_t *in, size_t n);   /* emulation prevention only */
void orc_md5(const uint8_t *data, size_t n, uint8_t out[16]);

size_t orc_write_vps(const orc_enc_params *p, uint8_t *out, size_t cap);
size_t orc_write_sps(const orc_enc_params *p, uint8_t *out, size_t cap);
size_t orc_write_pps(const orc_enc_params *p, uint8_t *out, size_t cap);
size_t orc_write_aud(int pic_type, uint8_t *out, size_t cap);
size_t orc_write_sei_hdr10(const orc_enc_params *p, uint8_t *out, size_t cap);
size_t orc_write_sei_buffering_period(const orc_enc_params *p, uint8_t *out, size_t cap);
size_t orc_write_sei_pic_timing(const orc_enc_params *p, int cpb_removal_delay, uint8_t *out, size_t cap);
size_t orc_write_sei_hash(const uint8_t md5[3][16], uint8_t *out, size_t cap);
/* slice segment header up to and including byte_alignment; entry = sizes of all but the last sub-stream */
size_t orc_write_slice_header(const orc_enc_params *p, int nal_type, int slice_type, int poc, int qp,
                              const uint32_t *entry, int n_entry, uint8_t *out, size_t cap);

/* ---- CABAC / syntax layer (hevc_cabac.c) */
typedef struct orc_frame_syntax {
    int wc, hc;                   /* coded size (multiples of 16) */
    int cuw, cuh;                 /* CU grid (16x16 units) */
    int ctuw, ctuh;               /* CTU grid (32x32) */
    int is_intra;                 /* I slice */
    int qp;
    orc_cu *cu;                   /* [cuh][cuw] */
    const int16_t *coef;          /* [cuh*cuw][ORC_CU_COEFS] */
    const orc_sao *sao;           /* [ctuh][ctuw], or NULL: SAO disabled */
    int bit_depth;
} orc_frame_syntax;

/* encode all CTU rows as WPP sub-streams; row r -> out + row_off[r], size row_len[r].  returns 0 / -1 */
int orc_cabac_encode_frame(orc_frame_syntax *f, uint8_t *out, size_t cap, uint32_t *row_off, uint32_t *row_len);
/* merge candidate list / AMVP helpers exposed for tests */
int orc_merge_candidates(const orc_frame_syntax *f, int cx, int cy, int16_t cand[5][2]);
int orc_amvp_candidates(const orc_frame_syntax *f, int cx, int cy, int16_t cand[2][2]);

/* ---- encoder (hevc_encode.c) */
typedef struct orc_encoder orc_encoder;
typedef struct orc_frame_info {
    int is_idr, poc, qp, bytes;
    int n_skip, n_merge, n_intra;
    double psnr_y;
    long long est_bits16;         /* rate-control size estimate of this frame, 1/16 bit units */
} orc_frame_info;

orc_encoder *orc_enc_create(const orc_enc_params *p);
void orc_enc_destroy(orc_encoder *e);
size_t orc_enc_headers(orc_encoder *e, uint8_t *out, size_t cap);      /* VPS + SPS + PPS (+ HDR10 SEI) */
/* encode one frame; y/u/v are display-size planes of 16-bit samples (strides in samples) */
long orc_enc_frame(orc_encoder *e, const pixel *y, int ys, const pixel *u, const pixel *v, int cs, int force_idr,
                   uint8_t *out, size_t cap, orc_frame_info *info);
/* reconstruction of the last encoded frame at CODED size (what the decoder outputs before cropping) */
void orc_enc_get_recon(const orc_encoder *e, pixel *y, pixel *u, pixel *v);
void orc_enc_coded_size(const orc_encoder *e, int *wc, int *hc);
const orc_cu *orc_enc_last_cus(const orc_encoder *e);
const int16_t *orc_enc_last_coefs(const orc_encoder *e);
const int16_t *orc_enc_last_coarse_mv(const orc_encoder *e);

/* ---- rate control (hevc_rc.c): deterministic, integer-only, driven by a size estimate computed from the levels */
typedef struct orc_rc {
    long long t16, b16;           /* bits*16 per frame interval, buffer size */
    long long fullness;           /* bits*16 currently in the decoder buffer model */
    int have[2], qp_prev[2];      /* [0] P frames, [1] IDR frames */
    long long est_prev[2];
    int poc;                      /* frames since the last IDR, of the frame last passed to orc_rc_update */
    int cascade;                  /* orc_enc_params.qp_cascade */
} orc_rc;
void orc_rc_init(orc_rc *rc, const orc_enc_params *p);
int orc_rc_pick_qp(const orc_rc *rc, const orc_enc_params *p, int is_idr);
void orc_rc_update(orc_rc *rc, int is_idr, int qp, long long est16);
long long orc_rc_budget(const orc_rc *rc, int is_idr);
int orc_rc_step(long long est, long long budget);
/* size estimate of one CU from its levels, 1/16 bit: 47 nnz + 22 sum floor(log2|l|) + 104 coded 4x4 sub-blocks + 160 (coded) | 80 */
long long orc_rc_cu_estimate(const int16_t *coef, int cbf);

/* from primitives.c */
int orc_sad(const pixel *a, int sa, const pixel *b, int sb, int w, int h);
int orc_satd(const pixel *a, int sa, const pixel *b, int sb, int w, int h);
void orc_fwd_transform(const int16_t *src, int stride, int16_t *dst, int N, int bit_depth, int is_dst);
void orc_inv_transform(const int16_t *src, int16_t *dst, int stride, int N, int bit_depth, int is_dst);
int orc_quant(const int16_t *coef, int16_t *level, int N, int qp, int bit_depth, int is_intra);
void orc_dequant(const int16_t *level, int16_t *coef, int N, int qp, int bit_depth);
void orc_intra_filter(const pixel *nb, pixel *out, int N, int strong, int bit_depth);
int orc_intra_use_filter(int N, int mode);
void orc_intra_pred(const pixel *nb, pixel *dst, int stride, int N, int mode, int edge, int bit_depth);
void orc_interp_luma(const pixel *ref, int rs, pixel *dst, int ds, int w, int h, int fx, int fy, int bit_depth);
void orc_interp_chroma(const pixel *ref, int rs, pixel *dst, int ds, int w, int h, int fx, int fy, int bit_depth);

#endif
