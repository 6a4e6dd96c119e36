/* TEST INFRASTRUCTURE -- CPU oracle: frame-level encoder model (decisions, reconstruction, access units).
 *
 * This is the executable specification of the B200 encoder (DESIGN.md "encoder specification"); the CUDA
 * path must produce byte-identical access units.  It replaces what the reference gets from
 * `ffmpeg -c:v libx265` (core/transcoder.py:398-412,506) for the tool subset described there:
 *   IDR + P frames (one reference, closed GOP), CTU 32, CU 16x16, 35-mode intra search by SATD,
 *   hierarchical motion search (quarter-resolution full search, integer refinement by SAD, half- and
 *   quarter-sample refinement by SATD), 16x16/8x8 DCT with flat dead-zone quantisation, WPP CABAC.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hevc_model.h"

int orc_intra_mpm(const orc_frame_syntax *f, int cx, int cy, int mpm[3]);

/* round(256 * sqrt(0.57 * 2^((qp - 12) / 3))) */
static const uint16_t k_lambda_q8[52] = {48, 54, 61, 68, 77, 86, 97, 108, 122, 137, 153, 172, 193, 217, 244, 273, 307, 344, 387, 434, 487, 547, 614,
                                         689, 773, 868, 974, 1093, 1227, 1378, 1546, 1736, 1948, 2187, 2454, 2755, 3092, 3471, 3896, 4373, 4909, 5510,
                                         6185, 6942, 7792, 8747, 9818, 11020, 12370, 13884, 15585, 17493};
static const uint8_t k_chroma_qp[14] = {29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37};   /* qPi 30..43 */

#define CME_RANGE 12      /* quarter-resolution search range (+-48 luma samples) */
#define MV_OVERHANG 64    /* a predicted block may leave the picture by this many luma samples */
#define MERGE_PASSES 2    /* merge-aware refinement passes over the motion field */

typedef struct plane {
    pixel *base, *p;      /* p points at sample (0,0) inside the padded allocation */
    int stride, w, h, pad;
} plane;

struct orc_encoder {
    orc_enc_params prm;
    int wc, hc, cuw, cuh, ctuw, ctuh;
    plane src[3], rec[2][3];
    plane pre[3];                    /* SAO on: reconstruction before SAO (deblocked); rec[cur] then receives the SAO output */
    orc_sao *sao;                    /* [ctuh][ctuw] SAO parameters of the current frame */
    int cur;                         /* index of the reconstruction being written */
    plane ds[2];                     /* quarter-resolution source: [cur_ds], [1 - cur_ds] = previous frame */
    int cur_ds;
    orc_cu *cus;
    int16_t *coefs;
    int16_t *cmv;                    /* [ctuh][ctuw][2] coarse vectors, quarter-resolution samples */
    int16_t *mvf[2];                 /* [cuh][cuw][2] motion field, ping-pong between the merge-aware passes */
    int32_t *satdf[2];               /* [cuh][cuw] SATD of that vector */
    int32_t *costf;                  /* [cuh][cuw] cost of the final inter choice (SATD + lambda * vector bits) */
    int32_t *mode_cost;              /* [cuh * cuw][35] luma SATD of every intra mode, predicted from SOURCE neighbours */
    long long sc_inter, sc_intra;    /* scene-cut measures of the current frame (coarse_search) */
    int frame_no, poc, since_bp;
    uint8_t *payload;                /* CABAC sub-streams */
    size_t payload_cap;
    uint32_t *row_off, *row_len;
    orc_rc rc;
};

static int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

static void plane_alloc(plane *pl, int w, int h, int pad)
{
    pl->w = w; pl->h = h; pl->pad = pad;
    pl->stride = w + 2 * pad;
    pl->base = (pixel *)calloc((size_t)pl->stride * (h + 2 * pad), sizeof(pixel));
    pl->p = pl->base + (size_t)pad * pl->stride + pad;
}

static void plane_extend(plane *pl)
{
    for (int y = 0; y < pl->h; y++) {
        pixel *row = pl->p + (size_t)y * pl->stride;
        for (int x = 1; x <= pl->pad; x++) { row[-x] = row[0]; row[pl->w - 1 + x] = row[pl->w - 1]; }
    }
    for (int y = 1; y <= pl->pad; y++) {
        memcpy(pl->p + (size_t)(-y) * pl->stride - pl->pad, pl->p - pl->pad, sizeof(pixel) * pl->stride);
        memcpy(pl->p + (size_t)(pl->h - 1 + y) * pl->stride - pl->pad, pl->p + (size_t)(pl->h - 1) * pl->stride - pl->pad,
               sizeof(pixel) * pl->stride);
    }
}

orc_encoder *orc_enc_create(const orc_enc_params *p)
{
    orc_encoder *e = (orc_encoder *)calloc(1, sizeof *e);
    e->prm = *p;
    e->wc = (p->width + 15) & ~15; e->hc = (p->height + 15) & ~15;
    e->cuw = e->wc / 16; e->cuh = e->hc / 16;
    e->ctuw = (e->wc + 31) / 32; e->ctuh = (e->hc + 31) / 32;
    for (int c = 0; c < 3; c++) {
        const int sh = c ? 1 : 0;
        plane_alloc(&e->src[c], e->wc >> sh, e->hc >> sh, 0);
        plane_alloc(&e->rec[0][c], e->wc >> sh, e->hc >> sh, ORC_PAD >> sh);
        plane_alloc(&e->rec[1][c], e->wc >> sh, e->hc >> sh, ORC_PAD >> sh);
        plane_alloc(&e->pre[c], e->wc >> sh, e->hc >> sh, ORC_PAD >> sh);
    }
    e->sao = (orc_sao *)calloc((size_t)e->ctuw * e->ctuh, sizeof(orc_sao));
    plane_alloc(&e->ds[0], e->wc / 4, e->hc / 4, 0);
    plane_alloc(&e->ds[1], e->wc / 4, e->hc / 4, 0);
    e->cus = (orc_cu *)calloc((size_t)e->cuw * e->cuh, sizeof(orc_cu));
    e->coefs = (int16_t *)calloc((size_t)e->cuw * e->cuh * ORC_CU_COEFS, sizeof(int16_t));
    e->cmv = (int16_t *)calloc((size_t)e->ctuw * e->ctuh * 2, sizeof(int16_t));
    for (int k = 0; k < 2; k++) {
        e->mvf[k] = (int16_t *)calloc((size_t)e->cuw * e->cuh * 2, sizeof(int16_t));
        e->satdf[k] = (int32_t *)calloc((size_t)e->cuw * e->cuh, sizeof(int32_t));
    }
    e->costf = (int32_t *)calloc((size_t)e->cuw * e->cuh, sizeof(int32_t));
    e->mode_cost = (int32_t *)calloc((size_t)e->cuw * e->cuh * 35, sizeof(int32_t));
    e->payload_cap = (size_t)e->wc * e->hc * 3 + 65536;
    e->payload = (uint8_t *)malloc(e->payload_cap);
    e->row_off = (uint32_t *)calloc(e->ctuh, sizeof(uint32_t));
    e->row_len = (uint32_t *)calloc(e->ctuh, sizeof(uint32_t));
    orc_rc_init(&e->rc, p);
    return e;
}

void orc_enc_destroy(orc_encoder *e)
{
    if (!e) return;
    for (int c = 0; c < 3; c++) { free(e->src[c].base); free(e->rec[0][c].base); free(e->rec[1][c].base); free(e->pre[c].base); }
    free(e->sao);
    free(e->ds[0].base); free(e->ds[1].base);
    free(e->cus); free(e->coefs); free(e->cmv); free(e->mvf[0]); free(e->mvf[1]); free(e->satdf[0]); free(e->satdf[1]); free(e->costf); free(e->mode_cost); free(e->payload); free(e->row_off); free(e->row_len);
    free(e);
}

void orc_enc_coded_size(const orc_encoder *e, int *wc, int *hc) { *wc = e->wc; *hc = e->hc; }
const orc_cu *orc_enc_last_cus(const orc_encoder *e) { return e->cus; }
const int16_t *orc_enc_last_coefs(const orc_encoder *e) { return e->coefs; }
const int16_t *orc_enc_last_coarse_mv(const orc_encoder *e) { return e->cmv; }

void orc_enc_get_recon(const orc_encoder *e, pixel *y, pixel *u, pixel *v)
{
    pixel *dst[3] = {y, u, v};
    const plane *r = e->rec[1 - e->cur];      /* cur was flipped after the frame was finished */
    for (int c = 0; c < 3; c++)
        for (int row = 0; row < r[c].h; row++)
            memcpy(dst[c] + (size_t)row * r[c].w, r[c].p + (size_t)row * r[c].stride, sizeof(pixel) * r[c].w);
}

size_t orc_enc_headers(orc_encoder *e, uint8_t *out, size_t cap)
{
    size_t o = 0;
    o += orc_write_vps(&e->prm, out + o, cap - o);
    o += orc_write_sps(&e->prm, out + o, cap - o);
    o += orc_write_pps(&e->prm, out + o, cap - o);
    return o;
}

/* planes the frame is reconstructed and deblocked in: with SAO on that is the `pre` set, and SAO writes the final picture */
static plane *work_planes(orc_encoder *e) { return e->prm.sao ? e->pre : e->rec[e->cur]; }

/* ------------------------------------------------------------------ ingest */

static void load_source(orc_encoder *e, const pixel *y, int ys, const pixel *u, const pixel *v, int cs)
{
    const pixel *in[3] = {y, u, v};
    for (int c = 0; c < 3; c++) {
        const int sh = c ? 1 : 0, w = e->prm.width >> sh, h = e->prm.height >> sh, st = c ? cs : ys;
        plane *d = &e->src[c];
        for (int row = 0; row < d->h; row++) {
            const pixel *s = in[c] + (size_t)(row < h ? row : h - 1) * st;
            pixel *o = d->p + (size_t)row * d->stride;
            memcpy(o, s, sizeof(pixel) * w);
            for (int x = w; x < d->w; x++) o[x] = s[w - 1];      /* replicate into the coded-size padding */
        }
    }
    plane *ds = &e->ds[e->cur_ds];
    const plane *s = &e->src[0];
#pragma omp parallel for
    for (int yy = 0; yy < ds->h; yy++)
        for (int xx = 0; xx < ds->w; xx++) {
            int acc = 8;
            for (int j = 0; j < 4; j++)
                for (int i = 0; i < 4; i++)
                    acc += s->p[(size_t)(4 * yy + j) * s->stride + 4 * xx + i];
            ds->p[(size_t)yy * ds->stride + xx] = (pixel)(acc >> 4);
        }
}

/* ------------------------------------------------------------------ shared residual path */

static int chroma_qp(int qp_y)
{
    const int qpi = clampi(qp_y, 0, 57);
    return qpi < 30 ? qpi : qpi >= 44 ? qpi - 6 : k_chroma_qp[qpi - 30];
}

/* transform + quantise + reconstruct one NxN block; returns cbf */
static int code_block(const orc_encoder *e, const pixel *src, int ss, const pixel *pred, int ps, pixel *rec, int rs, int N, int qp,
                      int intra_slice, int inter_cu, int16_t *level)
{
    int16_t res[32 * 32], coef[32 * 32], deq[32 * 32];
    const int bd = e->prm.bit_depth, maxv = (1 << bd) - 1;
    for (int y = 0; y < N; y++)
        for (int x = 0; x < N; x++)
            res[y * N + x] = (int16_t)((int)src[y * ss + x] - (int)pred[y * ps + x]);
    orc_fwd_transform(res, N, coef, N, bd, 0);
    int nsig = orc_quant(coef, level, N, qp + 6 * (bd - 8), bd, intra_slice);      /* dead zone: 171/512 in I slices, 85/512 in P slices */
    /* An inter luma block that carries nothing but one or two isolated +-1 levels is dropped: such levels cost far more bits
     * (last position, flags, and often the difference between a coded and a skipped CU) than the distortion they remove:
     * -7..-8 % bits for -0.1..-0.2 dB luma on the calibration clip.  Chroma blocks are kept (dropping them bought 1 % for up to
     * 0.6 dB of chroma PSNR). */
    if (nsig && inter_cu && N == 16) {
        int sum = 0, mx = 0;
        for (int i = 0; i < N * N; i++) { const int a = abs(level[i]); sum += a; if (a > mx) mx = a; }
        if (mx <= 1 && sum <= 2) { memset(level, 0, sizeof(int16_t) * N * N); nsig = 0; }
    }
    if (nsig) {
        orc_dequant(level, deq, N, qp + 6 * (bd - 8), bd);
        orc_inv_transform(deq, res, N, N, bd, 0);
    }
    for (int y = 0; y < N; y++)
        for (int x = 0; x < N; x++)
            rec[y * rs + x] = (pixel)(nsig ? clampi(pred[y * ps + x] + res[y * N + x], 0, maxv) : pred[y * ps + x]);
    return nsig != 0;
}

/* ------------------------------------------------------------------ intra frame */

static int cu_order(const orc_encoder *e, int cx, int cy) { return ((cy >> 1) * e->ctuw + (cx >> 1)) * 4 + ((cy & 1) << 1) + (cx & 1); }
static int cu_avail(const orc_encoder *e, int cx, int cy, int nx, int ny)
{
    if (nx < 0 || ny < 0 || nx >= e->cuw || ny >= e->cuh) return 0;
    return cu_order(e, nx, ny) < cu_order(e, cx, cy);
}

/* 8.4.4.2.2 reference sample gathering + substitution for the NxN block of CU (cx, cy) in plane `r` */
static void gather_neighbours(const orc_encoder *e, const plane *r, int cx, int cy, int N, pixel *nb)
{
    const int x0 = cx * N, y0 = cy * N, n2 = 2 * N, bd = e->prm.bit_depth;
    /* availability per N-sample segment: [bottom-left, left, corner, top, top-right] */
    const int a_bl = cu_avail(e, cx, cy, cx - 1, cy + 1), a_l = cu_avail(e, cx, cy, cx - 1, cy), a_c = cu_avail(e, cx, cy, cx - 1, cy - 1);
    const int a_t = cu_avail(e, cx, cy, cx, cy - 1), a_tr = cu_avail(e, cx, cy, cx + 1, cy - 1);
    pixel *top = nb + 1, *left = nb + 1 + n2;
    if (!(a_bl | a_l | a_c | a_t | a_tr)) {
        for (int i = 0; i < 4 * N + 1; i++) nb[i] = (pixel)(1 << (bd - 1));
        return;
    }
    /* walk from the bottom-most left sample up to the corner, then along the top to the right */
    int have = 0;
    pixel last = 0;
    /* first available sample in walk order */
    if (a_bl) last = r->p[(size_t)(y0 + n2 - 1) * r->stride + x0 - 1];
    else if (a_l) last = r->p[(size_t)(y0 + N - 1) * r->stride + x0 - 1];
    else if (a_c) last = r->p[(size_t)(y0 - 1) * r->stride + x0 - 1];
    else if (a_t) last = r->p[(size_t)(y0 - 1) * r->stride + x0];
    else last = r->p[(size_t)(y0 - 1) * r->stride + x0 + N];
    (void)have;
    for (int i = n2 - 1; i >= 0; i--) {
        const int av = i >= N ? a_bl : a_l;
        if (av) last = r->p[(size_t)(y0 + i) * r->stride + x0 - 1];
        left[i] = last;
    }
    if (a_c) last = r->p[(size_t)(y0 - 1) * r->stride + x0 - 1];
    nb[0] = last;
    for (int i = 0; i < n2; i++) {
        const int av = i >= N ? a_tr : a_t;
        if (av) last = r->p[(size_t)(y0 - 1) * r->stride + x0 + i];
        top[i] = last;
    }
}

/* Intra mode search of every CU at once: luma SATD of the 35 predictions built from the SOURCE picture's neighbour samples
 * (same availability / substitution rules as the real prediction), so that nothing here depends on the reconstruction and the
 * GPU runs it as one launch.  The wavefront stage adds the signalling cost against the real most-probable modes. */
static void intra_search_all(orc_encoder *e, const int32_t *satd1, int lambda)
{
    const int bd = e->prm.bit_depth, ncu = e->cuw * e->cuh;
    /* P frames (satd1 = luma SATD of every CU's vector after the first merge-aware pass): only CUs that inter prediction
     * serves badly are searched -- a residual worth coding at this quantiser (SATD above lambda / 2 per sample) that is also
     * more than twice the frame's mean SATD, or above 4 lambda / 8 grey levels per sample outright */
    const long thr = (256L * lambda) >> 9;
    long long sum = 0;
    if (satd1)
        for (int i = 0; i < ncu; i++) sum += satd1[i];
#pragma omp parallel for schedule(dynamic, 8)
    for (int idx = 0; idx < ncu; idx++) {
        const int cx = idx % e->cuw, cy = idx / e->cuw;
        if (satd1 && !(satd1[idx] > thr && ((long long)satd1[idx] * ncu > 2 * sum || satd1[idx] > 8 * thr || satd1[idx] > (2048 << (bd - 8))))) {
            e->mode_cost[idx * 35] = -1;          /* not a candidate */
            continue;
        }
        const pixel *src = e->src[0].p + (size_t)cy * 16 * e->src[0].stride + cx * 16;
        pixel nbs[65], flts[65], pred[256];
        gather_neighbours(e, &e->src[0], cx, cy, 16, nbs);
        orc_intra_filter(nbs, flts, 16, 0, bd);
        for (int mode = 0; mode < 35; mode++) {
            orc_intra_pred(orc_intra_use_filter(16, mode) ? flts : nbs, pred, 16, 16, mode, 1, bd);
            e->mode_cost[idx * 35 + mode] = orc_satd(src, e->src[0].stride, pred, 16, 16, 16);
        }
    }
}

/* one intra CU: mode = argmin(searched SATD + lambda * bits against the real MPM list), prediction from the reconstructed
 * neighbours (inter or intra: constrained_intra_pred is off), transform / quantisation / reconstruction */
static void intra_cu(orc_encoder *e, const orc_frame_syntax *fs, plane *rec, int cx, int cy, int qp, int lambda, int intra_slice)
{
    const int bd = e->prm.bit_depth, idx = cy * e->cuw + cx;
    orc_cu *cu = &e->cus[idx];
    int16_t *coef = e->coefs + (size_t)idx * ORC_CU_COEFS;
    pixel nb[65], flt[65], pred[256];
    int mpm[3];
    const pixel *src = e->src[0].p + (size_t)cy * 16 * e->src[0].stride + cx * 16;
    gather_neighbours(e, &rec[0], cx, cy, 16, nb);
    orc_intra_filter(nb, flt, 16, 0, bd);
    cu->pred_mode = 0; cu->mvx = cu->mvy = 0; cu->skip = 0;
    orc_intra_mpm(fs, cx, cy, mpm);
    long best = -1;
    int best_mode = 0;
    for (int mode = 0; mode < 35; mode++) {
        const int bits = mode == mpm[0] ? 2 : (mode == mpm[1] || mode == mpm[2]) ? 3 : 6;
        const long cost = e->mode_cost[idx * 35 + mode] + ((lambda * bits) >> 8);
        if (best < 0 || cost < best) { best = cost; best_mode = mode; }
    }
    cu->intra_mode = (uint8_t)best_mode;
    orc_intra_pred(orc_intra_use_filter(16, best_mode) ? flt : nb, pred, 16, 16, best_mode, 1, bd);
    pixel *ry = rec[0].p + (size_t)cy * 16 * rec[0].stride + cx * 16;
    int cbf = code_block(e, src, e->src[0].stride, pred, 16, ry, rec[0].stride, 16, qp, intra_slice, 0, coef);
    for (int c = 1; c < 3; c++) {
        pixel cnb[33], cpred[64];
        const pixel *cs = e->src[c].p + (size_t)cy * 8 * e->src[c].stride + cx * 8;
        pixel *rc = rec[c].p + (size_t)cy * 8 * rec[c].stride + cx * 8;
        gather_neighbours(e, &rec[c], cx, cy, 8, cnb);
        orc_intra_pred(cnb, cpred, 8, 8, best_mode, 0, bd);
        cbf |= code_block(e, cs, e->src[c].stride, cpred, 8, rc, rec[c].stride, 8, chroma_qp(qp), intra_slice, 0, coef + (c == 1 ? 256 : 320)) << c;
    }
    cu->cbf = (uint8_t)cbf;
}

/* the wavefront stage in decoding order; `only_flagged`: P frame, reconstruct only the CUs the decision stage marked intra */
static void intra_pass(orc_encoder *e, int qp, int intra_slice, int only_flagged)
{
    const int lambda = k_lambda_q8[qp] << (e->prm.bit_depth - 8);
    plane *rec = work_planes(e);
    orc_frame_syntax fs = {e->wc, e->hc, e->cuw, e->cuh, e->ctuw, e->ctuh, intra_slice, qp, e->cus, e->coefs, NULL, e->prm.bit_depth};
    for (int ctu = 0; ctu < e->ctuw * e->ctuh; ctu++)
        for (int k = 0; k < 4; k++) {
            const int cx = (ctu % e->ctuw) * 2 + (k & 1), cy = (ctu / e->ctuw) * 2 + (k >> 1);
            if (cx >= e->cuw || cy >= e->cuh) continue;
            if (only_flagged && e->cus[cy * e->cuw + cx].pred_mode != 0) continue;
            intra_cu(e, &fs, rec, cx, cy, qp, lambda, intra_slice);
        }
}

static void encode_intra_frame(orc_encoder *e, int qp)
{
    for (int i = 0; i < e->cuw * e->cuh; i++) e->cus[i].pred_mode = 0;
    intra_pass(e, qp, 1, 0);
}

/* ------------------------------------------------------------------ inter frame */

static int mv_bits1(int v)
{
    int a = abs(v), n = 0;
    while ((a + 1) >> (n + 1)) n++;
    return 2 * n + 1;
}
static int mv_cost(int lambda, int mvx, int mvy, int px, int py) { return (lambda * (mv_bits1(mvx - px) + mv_bits1(mvy - py))) >> 8; }

static pixel ds_at(const plane *d, int x, int y) { return d->p[(size_t)clampi(y, 0, d->h - 1) * d->stride + clampi(x, 0, d->w - 1)]; }

static void coarse_search(orc_encoder *e)
{
    const plane *cur = &e->ds[e->cur_ds], *prev = &e->ds[1 - e->cur_ds];
    /* the search runs on the 8 most significant bits of the quarter-resolution samples (byte SAD on the GPU) */
    const int sh = e->prm.bit_depth - 8, bias = 1;
    long long sc_inter = 0, sc_intra = 0;
#pragma omp parallel for schedule(dynamic) reduction(+ : sc_inter, sc_intra)
    for (int ty = 0; ty < e->ctuh; ty++)
        for (int tx = 0; tx < e->ctuw; tx++) {
            long best = -1;
            int bdx = 0, bdy = 0;
            for (int dy = -CME_RANGE; dy <= CME_RANGE; dy++)
                for (int dx = -CME_RANGE; dx <= CME_RANGE; dx++) {
                    long sad = 0;
                    for (int j = 0; j < 8; j++)
                        for (int i = 0; i < 8; i++)
                            sad += abs((int)(ds_at(cur, tx * 8 + i, ty * 8 + j) >> sh) - (int)(ds_at(prev, tx * 8 + i + dx, ty * 8 + j + dy) >> sh));
                    const long cost = sad + bias * (abs(dx) + abs(dy));
                    if (best < 0 || cost < best) { best = cost; bdx = dx; bdy = dy; }
                }
            e->cmv[(ty * e->ctuw + tx) * 2] = (int16_t)bdx;
            e->cmv[(ty * e->ctuw + tx) * 2 + 1] = (int16_t)bdy;
            /* scene-cut measures: the winning inter cost against an intra measure of the same 8x8 block = the smaller of its
             * horizontal / vertical neighbour-difference sums (56 sample pairs each) */
            long hs = 0, vs = 0;
            for (int j = 0; j < 8; j++)
                for (int i = 0; i < 8; i++) {
                    const int c = ds_at(cur, tx * 8 + i, ty * 8 + j) >> sh;
                    if (i) hs += abs(c - (int)(ds_at(cur, tx * 8 + i - 1, ty * 8 + j) >> sh));
                    if (j) vs += abs(c - (int)(ds_at(cur, tx * 8 + i, ty * 8 + j - 1) >> sh));
                }
            sc_inter += best;
            sc_intra += hs < vs ? hs : vs;
        }
    e->sc_inter = sc_inter; e->sc_intra = sc_intra;
}

/* a scene cut: the motion-compensated quarter-resolution picture is no better a predictor than the picture's own neighbour
 * samples (inter >= 1.5 x intra) and the mismatch is large in absolute terms (mean >= 5 per sample at 8 bits) */
static int scene_cut(const orc_encoder *e)
{
    const long long n = (long long)e->ctuw * e->ctuh * 64;
    return e->sc_inter >= 5 * n && 2 * e->sc_inter >= 3 * e->sc_intra;
}

typedef struct { int x, y; } mv_t;

/* SAD of a 16x16 block on the 8 most significant bits of the samples (integer search stages; byte SAD on the GPU) */
static long sad8_16x16(const pixel *a, int sa, const pixel *b, int sb, int sh)
{
    long s = 0;
    for (int y = 0; y < 16; y++)
        for (int x = 0; x < 16; x++)
            s += abs((int)(a[y * sa + x] >> sh) - (int)(b[y * sb + x] >> sh));
    return s;
}

static mv_t clamp_mv(const orc_encoder *e, int x0, int y0, mv_t m)
{
    m.x = clampi(m.x, -(x0 + MV_OVERHANG) * 4, (e->wc - x0 - 16 + MV_OVERHANG) * 4);
    m.y = clampi(m.y, -(y0 + MV_OVERHANG) * 4, (e->hc - y0 - 16 + MV_OVERHANG) * 4);
    return m;
}

static mv_t ctu_mv(const orc_encoder *e, int tx, int ty)
{
    tx = clampi(tx, 0, e->ctuw - 1); ty = clampi(ty, 0, e->ctuh - 1);
    mv_t m = {e->cmv[(ty * e->ctuw + tx) * 2] * 16, e->cmv[(ty * e->ctuw + tx) * 2 + 1] * 16};
    return m;
}

static void predict_luma(const orc_encoder *e, const plane *ref, int x0, int y0, mv_t m, pixel *dst)
{
    const pixel *r = ref->p + (size_t)(y0 + (m.y >> 2)) * ref->stride + x0 + (m.x >> 2);
    orc_interp_luma(r, ref->stride, dst, 16, 16, 16, m.x & 3, m.y & 3, e->prm.bit_depth);
}

static void encode_inter_frame(orc_encoder *e, int qp)
{
    const int bd = e->prm.bit_depth;
    const int lambda = k_lambda_q8[qp] << (bd - 8);
    plane *rec = work_planes(e);
    const plane *ref = e->rec[1 - e->cur];
    /* every loop over CUs below is order-independent (that is what lets the GPU run them as one launch each), so the model
     * spreads them over the host cores */
#pragma omp parallel for collapse(2) schedule(dynamic, 8)
    for (int cy = 0; cy < e->cuh; cy++)
        for (int cx = 0; cx < e->cuw; cx++) {
            const int x0 = cx * 16, y0 = cy * 16, tx = cx >> 1, ty = cy >> 1;
            const pixel *src = e->src[0].p + (size_t)y0 * e->src[0].stride + x0;
            const int ss = e->src[0].stride;
            const mv_t pred = clamp_mv(e, x0, y0, ctu_mv(e, tx, ty));
            mv_t cand[6] = {{0, 0}, pred, ctu_mv(e, tx - 1, ty), ctu_mv(e, tx, ty - 1), ctu_mv(e, tx + 1, ty), ctu_mv(e, tx, ty + 1)};
            mv_t best = {0, 0};
            long bcost = -1;
            /* integer stages: SAD on 8-bit samples, lambda at 8-bit scale */
            const int lambda8 = k_lambda_q8[qp], sh8 = bd - 8;
            for (int k = 0; k < 6; k++) {
                const mv_t m = clamp_mv(e, x0, y0, cand[k]);
                const pixel *r = ref[0].p + (size_t)(y0 + (m.y >> 2)) * ref[0].stride + x0 + (m.x >> 2);
                const long cost = sad8_16x16(src, ss, r, ref[0].stride, sh8) + mv_cost(lambda8, m.x, m.y, pred.x, pred.y);
                if (bcost < 0 || cost < bcost) { bcost = cost; best = m; }
            }
            {
                const mv_t centre = best;
                for (int dy = -2; dy <= 2; dy++)
                    for (int dx = -2; dx <= 2; dx++) {
                        if (!dx && !dy) continue;
                        mv_t m = {centre.x + 4 * dx, centre.y + 4 * dy};
                        const mv_t cm = clamp_mv(e, x0, y0, m);
                        if (cm.x != m.x || cm.y != m.y) continue;
                        const pixel *r = ref[0].p + (size_t)(y0 + (m.y >> 2)) * ref[0].stride + x0 + (m.x >> 2);
                        const long cost = sad8_16x16(src, ss, r, ref[0].stride, sh8) + mv_cost(lambda8, m.x, m.y, pred.x, pred.y);
                        if (cost < bcost) { bcost = cost; best = m; }
                    }
            }
            /* sub-sample stages: SATD on the normative interpolation */
            pixel blk[256];
            predict_luma(e, &ref[0], x0, y0, best, blk);
            const long satd_int = orc_satd(src, ss, blk, 16, 16, 16);
            bcost = satd_int + mv_cost(lambda, best.x, best.y, pred.x, pred.y);
            /* a block the integer vector already predicts to within a quarter grey level per sample (static, clean content) has
             * nothing to gain from sub-sample refinement: the search stops here */
            for (int step = satd_int > (64 << (bd - 8)) ? 2 : 0; step >= 1; step--) {
                const mv_t centre = best;
                /* column by column: the three candidates of a column share their horizontal filter pass on the GPU */
                for (int dx = -1; dx <= 1; dx++)
                    for (int dy = -1; dy <= 1; dy++) {
                        if (!dx && !dy) continue;
                        mv_t m = {centre.x + step * dx, centre.y + step * dy};
                        const mv_t cm = clamp_mv(e, x0, y0, m);
                        if (cm.x != m.x || cm.y != m.y) continue;
                        predict_luma(e, &ref[0], x0, y0, m, blk);
                        const long cost = orc_satd(src, ss, blk, 16, 16, 16) + mv_cost(lambda, m.x, m.y, pred.x, pred.y);
                        if (cost < bcost) { bcost = cost; best = m; }
                    }
            }
            e->mvf[0][(cy * e->cuw + cx) * 2] = (int16_t)best.x;
            e->mvf[0][(cy * e->cuw + cx) * 2 + 1] = (int16_t)best.y;
            e->satdf[0][cy * e->cuw + cx] = (int32_t)(bcost - mv_cost(lambda, best.x, best.y, pred.x, pred.y));
        }
    /* Merge-aware passes (MERGE_PASSES Jacobi iterations).  Each CU compares its vector with the vectors its five merge-candidate
     * neighbours (A1, B1, B0, A0, B2) held after the previous pass, and with the zero vector: taking a neighbour's vector costs
     * a merge index (~2 bits) instead of a vector difference.  Reading the previous pass keeps every pass order-independent
     * (data-parallel); the normative merge / AMVP derivation still happens afterwards on the final field. */
    for (int pass = 0; pass < MERGE_PASSES; pass++) {
        const int16_t *mvi = e->mvf[pass & 1];
        const int32_t *sdi = e->satdf[pass & 1];
        int16_t *mvo = e->mvf[(pass + 1) & 1];
        int32_t *sdo = e->satdf[(pass + 1) & 1];
#pragma omp parallel for collapse(2) schedule(dynamic, 8)
        for (int cy = 0; cy < e->cuh; cy++)
            for (int cx = 0; cx < e->cuw; cx++) {
                const int x0 = cx * 16, y0 = cy * 16, idx = cy * e->cuw + cx;
                const pixel *src = e->src[0].p + (size_t)y0 * e->src[0].stride + x0;
                const int ss = e->src[0].stride;
                const mv_t own = {mvi[idx * 2], mvi[idx * 2 + 1]};
                static const int nbx[5] = {-1, 0, 1, -1, -1}, nby[5] = {0, -1, -1, 1, -1};
                mv_t cand[6];
                int ncand = 0;
                for (int k = 0; k < 5; k++) {
                    const int nx = cx + nbx[k], ny = cy + nby[k];
                    if (nx < 0 || ny < 0 || nx >= e->cuw || ny >= e->cuh) continue;
                    cand[ncand].x = mvi[(ny * e->cuw + nx) * 2];
                    cand[ncand].y = mvi[(ny * e->cuw + nx) * 2 + 1];
                    ncand++;
                }
                cand[ncand].x = 0; cand[ncand].y = 0; ncand++;
                int own_bits = -1;
                for (int k = 0; k < ncand; k++) {
                    const int bts = (cand[k].x == own.x && cand[k].y == own.y) ? 0 : mv_bits1(own.x - cand[k].x) + mv_bits1(own.y - cand[k].y);
                    if (own_bits < 0 || bts < own_bits) own_bits = bts;
                }
                mv_t best = own;
                long bsatd = sdi[idx];
                long bcost = bsatd + ((lambda * (own_bits + 2)) >> 8) + 1;      /* +1: ties go to a merge candidate */
                pixel blk[256];
                for (int k = 0; k < ncand; k++) {
                    const mv_t m = cand[k];
                    int dup = (m.x == own.x && m.y == own.y);
                    for (int j = 0; j < k && !dup; j++) dup = cand[j].x == m.x && cand[j].y == m.y;
                    if (dup) continue;
                    const mv_t cm = clamp_mv(e, x0, y0, m);
                    if (cm.x != m.x || cm.y != m.y) continue;
                    if (abs((m.x >> 2) - (own.x >> 2)) > 2 || abs((m.y >> 2) - (own.y >> 2)) > 2) continue;   /* inside the staged window */
                    predict_luma(e, &ref[0], x0, y0, m, blk);
                    const long sd = orc_satd(src, ss, blk, 16, 16, 16);
                    const long cost = sd + ((lambda * 2) >> 8);
                    if (cost < bcost) { bcost = cost; best = m; bsatd = sd; }
                }
                mvo[idx * 2] = (int16_t)best.x; mvo[idx * 2 + 1] = (int16_t)best.y;
                sdo[idx] = (int32_t)bsatd;
                e->costf[idx] = (int32_t)bcost;
            }
    }
    /* Intra CUs in P frames: a CU goes intra when the best intra prediction (searched on source neighbours) plus its
     * signalling costs less than 9/8 of the final inter choice.  Only CUs the gate in intra_search_all() passes are candidates
     * -- inter residuals worth coding -- and for those the measured optimum sits past parity (BD-rate against the margin, 360p
     * calibration clips: 3/4 -3.9 %, 1 -6.2 %, 9/8 -7.0 %, 5/4 -7.3 % with static scenes starting to lose): an intra residual
     * after a good angular prediction is flatter and quantises cheaper than an inter residual of equal SATD. */
    if (e->prm.intra_in_p) intra_search_all(e, e->satdf[1], lambda);      /* [1] = output of the first merge-aware pass */
    const int16_t *mvfinal = e->mvf[MERGE_PASSES & 1];
#pragma omp parallel for collapse(2) schedule(dynamic, 8)
    for (int cy = 0; cy < e->cuh; cy++)
        for (int cx = 0; cx < e->cuw; cx++) {
            orc_cu *cu = &e->cus[cy * e->cuw + cx];
            int16_t *coef = e->coefs + (size_t)(cy * e->cuw + cx) * ORC_CU_COEFS;
            const int x0 = cx * 16, y0 = cy * 16;
            const pixel *src = e->src[0].p + (size_t)y0 * e->src[0].stride + x0;
            const int ss = e->src[0].stride;
            const mv_t best = {mvfinal[(cy * e->cuw + cx) * 2], mvfinal[(cy * e->cuw + cx) * 2 + 1]};
            pixel blk[256];
            cu->pred_mode = 1; cu->intra_mode = 1; cu->skip = 0;
            cu->mvx = (int16_t)best.x; cu->mvy = (int16_t)best.y;
            if (e->prm.intra_in_p) {
                const int idx = cy * e->cuw + cx;
                long ibest = e->mode_cost[idx * 35];
                for (int m = 1; m < 35 && ibest >= 0; m++) if (e->mode_cost[idx * 35 + m] < ibest) ibest = e->mode_cost[idx * 35 + m];
                const long icost = ibest + ((lambda * 12) >> 8);
                if (ibest >= 0 && icost * 8 < (long)e->costf[idx] * 9) {
                    cu->pred_mode = 0; cu->mvx = cu->mvy = 0; cu->cbf = 0;     /* reconstructed by the wavefront stage below */
                    continue;
                }
            }
            /* reconstruct */
            predict_luma(e, &ref[0], x0, y0, best, blk);
            pixel *ry = rec[0].p + (size_t)y0 * rec[0].stride + x0;
            int cbf = code_block(e, src, ss, blk, 16, ry, rec[0].stride, 16, qp, 0, 1, coef);
            for (int c = 1; c < 3; c++) {
                pixel cpred[64];
                const pixel *r = ref[c].p + (size_t)(cy * 8 + (best.y >> 3)) * ref[c].stride + cx * 8 + (best.x >> 3);
                orc_interp_chroma(r, ref[c].stride, cpred, 8, 8, 8, best.x & 7, best.y & 7, bd);
                const pixel *cs = e->src[c].p + (size_t)cy * 8 * e->src[c].stride + cx * 8;
                pixel *rc = rec[c].p + (size_t)cy * 8 * rec[c].stride + cx * 8;
                cbf |= code_block(e, cs, e->src[c].stride, cpred, 8, rc, rec[c].stride, 8, chroma_qp(qp), 0, 1, coef + (c == 1 ? 256 : 320)) << c;
            }
            cu->cbf = (uint8_t)cbf;
        }
    if (e->prm.intra_in_p) intra_pass(e, qp, 0, 1);
}

/* ------------------------------------------------------------------ in-loop deblocking filter (H.265 8.7.2) */

static const uint8_t k_tc_table[54] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4,
                                       5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24};
static const uint8_t k_beta_table[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 20, 22, 24,
                                         26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64};

/* boundary strength between two neighbouring CUs (8.7.2.4): 2 intra, 1 coefficients or vectors apart by a sample, else 0 */
static int boundary_strength(const orc_cu *p, const orc_cu *q)
{
    if (p->pred_mode == 0 || q->pred_mode == 0) return 2;
    if ((p->cbf | q->cbf) & 1) return 1;
    if (abs(p->mvx - q->mvx) >= 4 || abs(p->mvy - q->mvy) >= 4) return 1;
    return 0;
}

/* one luma edge segment of 4 lines; s points at q0 of line 0; `step` = distance across the edge, `line` = distance between lines */
static void deblock_luma_segment(pixel *s, ptrdiff_t step, ptrdiff_t line, int beta, int tc, int maxv)
{
#define P(i, k) ((int)s[(k) * line - ((i) + 1) * step])
#define Q(i, k) ((int)s[(k) * line + (i) * step])
    const int dp0 = abs(P(2, 0) - 2 * P(1, 0) + P(0, 0)), dp3 = abs(P(2, 3) - 2 * P(1, 3) + P(0, 3));
    const int dq0 = abs(Q(2, 0) - 2 * Q(1, 0) + Q(0, 0)), dq3 = abs(Q(2, 3) - 2 * Q(1, 3) + Q(0, 3));
    const int dpq0 = dp0 + dq0, dpq3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3, d = dpq0 + dpq3;
    if (d >= beta) return;
    const int strong0 = 2 * dpq0 < (beta >> 2) && abs(P(3, 0) - P(0, 0)) + abs(Q(0, 0) - Q(3, 0)) < (beta >> 3) &&
                        abs(P(0, 0) - Q(0, 0)) < ((5 * tc + 1) >> 1);
    const int strong3 = 2 * dpq3 < (beta >> 2) && abs(P(3, 3) - P(0, 3)) + abs(Q(0, 3) - Q(3, 3)) < (beta >> 3) &&
                        abs(P(0, 3) - Q(0, 3)) < ((5 * tc + 1) >> 1);
    const int strong = strong0 && strong3;
    const int dep = dp < ((beta + (beta >> 1)) >> 3), deq = dq < ((beta + (beta >> 1)) >> 3);
    for (int k = 0; k < 4; k++) {
        const int p0 = P(0, k), p1 = P(1, k), p2 = P(2, k), p3 = P(3, k), q0 = Q(0, k), q1 = Q(1, k), q2 = Q(2, k), q3 = Q(3, k);
        if (strong) {
            s[k * line - 1 * step] = (pixel)clampi((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3, p0 - 2 * tc, p0 + 2 * tc);
            s[k * line - 2 * step] = (pixel)clampi((p2 + p1 + p0 + q0 + 2) >> 2, p1 - 2 * tc, p1 + 2 * tc);
            s[k * line - 3 * step] = (pixel)clampi((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3, p2 - 2 * tc, p2 + 2 * tc);
            s[k * line + 0 * step] = (pixel)clampi((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3, q0 - 2 * tc, q0 + 2 * tc);
            s[k * line + 1 * step] = (pixel)clampi((p0 + q0 + q1 + q2 + 2) >> 2, q1 - 2 * tc, q1 + 2 * tc);
            s[k * line + 2 * step] = (pixel)clampi((p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3, q2 - 2 * tc, q2 + 2 * tc);
        } else {
            int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
            if (abs(delta) < 10 * tc) {
                delta = clampi(delta, -tc, tc);
                s[k * line - 1 * step] = (pixel)clampi(p0 + delta, 0, maxv);
                s[k * line + 0 * step] = (pixel)clampi(q0 - delta, 0, maxv);
                if (dep) {
                    const int dlt = clampi((((p2 + p0 + 1) >> 1) - p1 + delta) >> 1, -(tc >> 1), tc >> 1);
                    s[k * line - 2 * step] = (pixel)clampi(p1 + dlt, 0, maxv);
                }
                if (deq) {
                    const int dlt = clampi((((q2 + q0 + 1) >> 1) - q1 - delta) >> 1, -(tc >> 1), tc >> 1);
                    s[k * line + 1 * step] = (pixel)clampi(q1 + dlt, 0, maxv);
                }
            }
        }
    }
#undef P
#undef Q
}

static void deblock_chroma_segment(pixel *s, ptrdiff_t step, ptrdiff_t line, int tc, int maxv, int lines)
{
    for (int k = 0; k < lines; k++) {
        const int p0 = s[k * line - step], p1 = s[k * line - 2 * step], q0 = s[k * line], q1 = s[k * line + step];
        const int delta = clampi((((q0 - p0) << 2) + p1 - q1 + 4) >> 3, -tc, tc);
        s[k * line - step] = (pixel)clampi(p0 + delta, 0, maxv);
        s[k * line] = (pixel)clampi(q0 - delta, 0, maxv);
    }
}

/* all vertical CU edges of the picture first, then all horizontal ones (8.7.2); CU edges are the only transform /
 * prediction edges in this encoder and all lie on the 8x8 luma (8-sample chroma) grid */
static void deblock_frame(orc_encoder *e, plane *rec, int qp)
{
    const int bd = e->prm.bit_depth, maxv = (1 << bd) - 1;
    const int beta = k_beta_table[clampi(qp, 0, 51)] << (bd - 8);
    for (int dir = 0; dir < 2; dir++)
        for (int cy = 0; cy < e->cuh; cy++)
            for (int cx = 0; cx < e->cuw; cx++) {
                if ((dir == 0 && cx == 0) || (dir == 1 && cy == 0)) continue;
                const orc_cu *q = &e->cus[cy * e->cuw + cx], *p = dir == 0 ? q - 1 : q - e->cuw;
                const int bs = boundary_strength(p, q);
                if (!bs) continue;
                const int tc = k_tc_table[clampi(qp + 2 * (bs - 1), 0, 53)] << (bd - 8);
                pixel *y = rec[0].p + (size_t)cy * 16 * rec[0].stride + cx * 16;
                for (int seg = 0; seg < 4; seg++) {
                    if (dir == 0) deblock_luma_segment(y + (size_t)seg * 4 * rec[0].stride, 1, rec[0].stride, beta, tc, maxv);
                    else deblock_luma_segment(y + seg * 4, rec[0].stride, 1, beta, tc, maxv);
                }
                if (bs == 2) {
                    const int tcc = k_tc_table[clampi(chroma_qp(qp) + 2, 0, 53)] << (bd - 8);
                    for (int c = 1; c < 3; c++) {
                        pixel *u = rec[c].p + (size_t)cy * 8 * rec[c].stride + cx * 8;
                        if (dir == 0) deblock_chroma_segment(u, 1, rec[c].stride, tcc, maxv, 8);
                        else deblock_chroma_segment(u, rec[c].stride, 1, tcc, maxv, 8);
                    }
                }
            }
}

/* ------------------------------------------------------------------ sample adaptive offset (H.265 8.7.3) */

typedef struct {
    long long cnt[4][5], sum[4][5];       /* edge classes x categories 1..4: samples, sum of (source - reconstruction) */
    long long bcnt[32], bsum[32];         /* bands */
} sao_stats;

static const int8_t k_sao_dx[4][2] = {{-1, 1}, {0, 0}, {-1, 1}, {1, -1}}, k_sao_dy[4][2] = {{0, 0}, {-1, 1}, {-1, 1}, {-1, 1}};

static int sgn(int v) { return (v > 0) - (v < 0); }

/* edge category of 8.7.3.2: 1 valley, 2 / 3 edges, 4 peak, 0 none */
static int sao_category(int r, int a, int b)
{
    static const uint8_t map[5] = {1, 2, 0, 3, 4};
    return map[2 + sgn(r - a) + sgn(r - b)];
}

static void sao_collect(const orc_encoder *e, const plane *rec, const plane *src, int N, int rx, int ry, sao_stats *st)
{
    const int x0 = rx * N, y0 = ry * N, bshift = e->prm.bit_depth - 5;
    memset(st, 0, sizeof *st);
    for (int y = y0; y < y0 + N && y < rec->h; y++)
        for (int x = x0; x < x0 + N && x < rec->w; x++) {
            const int r = rec->p[(size_t)y * rec->stride + x], d = (int)src->p[(size_t)y * src->stride + x] - r;
            st->bcnt[r >> bshift]++;
            st->bsum[r >> bshift] += d;
            for (int k = 0; k < 4; k++) {
                const int xa = x + k_sao_dx[k][0], ya = y + k_sao_dy[k][0], xb = x + k_sao_dx[k][1], yb = y + k_sao_dy[k][1];
                if (xa < 0 || xb < 0 || ya < 0 || yb < 0 || xa >= rec->w || xb >= rec->w || ya >= rec->h || yb >= rec->h) continue;
                const int cat = sao_category(r, rec->p[(size_t)ya * rec->stride + xa], rec->p[(size_t)yb * rec->stride + xb]);
                if (cat) { st->cnt[k][cat]++; st->sum[k][cat] += d; }
            }
        }
}

/* best offset in [lo, hi] for `cnt` samples whose errors sum to `sum`: minimises (cnt o^2 - 2 o sum) * 65536 + lam * bits,
 * walking from the rounded mean towards zero (first minimum wins).  sign_bit: band offsets carry a sign bit when non-zero. */
static int sao_offset(long long cnt, long long sum, int lo, int hi, long long lam, int cmax, int sign_bit, long long *cost_out)
{
    long long best = lam * 1;                 /* offset 0: one bin */
    int bo = 0;
    if (cnt) {
        int o = (int)(sum >= 0 ? (sum + cnt / 2) / cnt : -((-sum + cnt / 2) / cnt));
        o = clampi(o, lo, hi);
        for (int t = o; t != 0; t += t > 0 ? -1 : 1) {
            const int a = abs(t);
            const long long cost = (cnt * t * t - 2 * t * sum) * 65536 + lam * ((a < cmax ? a + 1 : cmax) + sign_bit);
            if (cost < best) { best = cost; bo = t; }
        }
    }
    *cost_out = best;
    return bo;
}

/* parameters of one component group (luma, or Cb + Cr) from the statistics of its `ncomp` components */
static void sao_decide_group(const sao_stats *st, int ncomp, long long lam, int cmax, orc_sao *out, int g)
{
    long long best = lam * 1;                 /* off: sao_type_idx = 0 */
    int type = 0, eo_class = 0, band[2] = {0, 0};
    int8_t off[2][4];
    memset(off, 0, sizeof off);
    for (int k = 0; k < 4; k++) {             /* edge classes */
        long long cost = lam * 4;             /* type (2 bins) + class (2 bins) */
        int8_t o[2][4];
        for (int c = 0; c < ncomp; c++)
            for (int cat = 1; cat <= 4; cat++) {
                long long cc;
                o[c][cat - 1] = (int8_t)sao_offset(st[c].cnt[k][cat], st[c].sum[k][cat], cat <= 2 ? 0 : -cmax, cat <= 2 ? cmax : 0, lam, cmax, 0, &cc);
                cost += cc;
            }
        if (cost < best) { best = cost; type = 2; eo_class = k; memcpy(off, o, sizeof off); }
    }
    {                                         /* band offset: per component the best window of four consecutive bands */
        long long cost = lam * 2;
        int8_t o[2][4];
        int bp[2] = {0, 0};
        for (int c = 0; c < ncomp; c++) {
            long long bc[32], wbest = 0;
            int8_t bo_[32];
            for (int b = 0; b < 32; b++) bo_[b] = (int8_t)sao_offset(st[c].bcnt[b], st[c].bsum[b], -cmax, cmax, lam, cmax, 1, &bc[b]);
            for (int s0 = 0; s0 <= 28; s0++) {
                const long long w = bc[s0] + bc[s0 + 1] + bc[s0 + 2] + bc[s0 + 3];
                if (s0 == 0 || w < wbest) { wbest = w; bp[c] = s0; }
            }
            for (int i = 0; i < 4; i++) o[c][i] = bo_[bp[c] + i];
            cost += wbest + lam * 5;
        }
        if (cost < best) { best = cost; type = 1; band[0] = bp[0]; band[1] = bp[1]; memcpy(off, o, sizeof off); }
    }
    out->type[g] = (uint8_t)type;
    out->eo_class[g] = (uint8_t)(type == 2 ? eo_class : 0);
    for (int c = 0; c < ncomp; c++) {
        out->band[g + c] = (uint8_t)(type == 1 ? band[c] : 0);
        for (int i = 0; i < 4; i++) out->offset[g + c][i] = type ? off[c][i] : 0;
    }
}

static void sao_apply_ctb(const orc_encoder *e, const plane *in, plane *out, int N, int rx, int ry, int type, int eo_class, int band,
                          const int8_t *off)
{
    const int x0 = rx * N, y0 = ry * N, bd = e->prm.bit_depth, maxv = (1 << bd) - 1, bshift = bd - 5;
    for (int y = y0; y < y0 + N && y < in->h; y++)
        for (int x = x0; x < x0 + N && x < in->w; x++) {
            const int r = in->p[(size_t)y * in->stride + x];
            int v = r;
            if (type == 1) {
                const int k = ((r >> bshift) - band) & 31;
                if (k < 4) v = clampi(r + off[k], 0, maxv);
            } else if (type == 2) {
                const int xa = x + k_sao_dx[eo_class][0], ya = y + k_sao_dy[eo_class][0], xb = x + k_sao_dx[eo_class][1], yb = y + k_sao_dy[eo_class][1];
                if (!(xa < 0 || xb < 0 || ya < 0 || yb < 0 || xa >= in->w || xb >= in->w || ya >= in->h || yb >= in->h)) {
                    const int cat = sao_category(r, in->p[(size_t)ya * in->stride + xa], in->p[(size_t)yb * in->stride + xb]);
                    if (cat) v = clampi(r + off[cat - 1], 0, maxv);
                }
            }
            out->p[(size_t)y * out->stride + x] = (pixel)v;
        }
}

/* decide (per CTU, independently: one launch on the GPU) and apply: `pre` (deblocked) -> rec[cur].  The Lagrangian is three
 * times the squared SATD-domain lambda: the rest of this encoder decides without RDO, and at lambda^2 SAO bought its PSNR at a
 * worse rate than a QP change does (BD-rate +1..+3 % on the calibration clips; 2 lambda^2 and 3 lambda^2 are within 0.4 % of
 * each other on moving content, 3 lambda^2 halves the cost on static scenes; tools/model_rd.py). */
static void sao_frame(orc_encoder *e, int qp)
{
    const int bd = e->prm.bit_depth, cmax = (1 << ((bd < 10 ? bd : 10) - 5)) - 1;
    const long long ly = (long long)(k_lambda_q8[qp] << (bd - 8)), lc = (long long)(k_lambda_q8[chroma_qp(qp)] << (bd - 8));
    plane *out = e->rec[e->cur];
#pragma omp parallel for schedule(dynamic, 4)
    for (int ctu = 0; ctu < e->ctuw * e->ctuh; ctu++) {
        const int rx = ctu % e->ctuw, ry = ctu / e->ctuw;
        orc_sao *s = &e->sao[ctu];
        sao_stats st[2];
        memset(s, 0, sizeof *s);
        {   /* a CTU whose CUs are all inter without residual is a copy of reference samples that already went through SAO:
             * offsetting it again fits noise and costs bits in every frame -- left off */
            int copied = 1;
            for (int k = 0; k < 4; k++) {
                const int cx = 2 * rx + (k & 1), cy = 2 * ry + (k >> 1);
                if (cx < e->cuw && cy < e->cuh && (e->cus[cy * e->cuw + cx].pred_mode == 0 || e->cus[cy * e->cuw + cx].cbf)) copied = 0;
            }
            if (copied) continue;
        }
        sao_collect(e, &e->pre[0], &e->src[0], 32, rx, ry, &st[0]);
        sao_decide_group(st, 1, 3 * ly * ly, cmax, s, 0);
        sao_collect(e, &e->pre[1], &e->src[1], 16, rx, ry, &st[0]);
        sao_collect(e, &e->pre[2], &e->src[2], 16, rx, ry, &st[1]);
        sao_decide_group(st, 2, 3 * lc * lc, cmax, s, 1);
    }
#pragma omp parallel for schedule(dynamic, 4)
    for (int ctu = 0; ctu < e->ctuw * e->ctuh; ctu++) {
        const int rx = ctu % e->ctuw, ry = ctu / e->ctuw;
        const orc_sao *s = &e->sao[ctu];
        sao_apply_ctb(e, &e->pre[0], &out[0], 32, rx, ry, s->type[0], s->eo_class[0], s->band[0], s->offset[0]);
        sao_apply_ctb(e, &e->pre[1], &out[1], 16, rx, ry, s->type[1], s->eo_class[1], s->band[1], s->offset[1]);
        sao_apply_ctb(e, &e->pre[2], &out[2], 16, rx, ry, s->type[1], s->eo_class[1], s->band[2], s->offset[2]);
    }
}

/* ------------------------------------------------------------------ access unit */

static void picture_md5(const orc_encoder *e, const plane *rec, uint8_t md5[3][16])
{
    const int bps = e->prm.bit_depth > 8 ? 2 : 1;
    for (int c = 0; c < 3; c++) {
        const size_t n = (size_t)rec[c].w * rec[c].h * bps;
        uint8_t *buf = (uint8_t *)malloc(n), *o = buf;
        for (int y = 0; y < rec[c].h; y++)
            for (int x = 0; x < rec[c].w; x++) {
                const pixel v = rec[c].p[(size_t)y * rec[c].stride + x];
                *o++ = (uint8_t)v;
                if (bps == 2) *o++ = (uint8_t)(v >> 8);
            }
        orc_md5(buf, n, md5[c]);
        free(buf);
    }
}

long orc_enc_frame(orc_encoder *e, const pixel *y, int ys, const pixel *u, const pixel *v, int cs, int force_idr, uint8_t *out,
                   size_t cap, orc_frame_info *info)
{
    const orc_enc_params *p = &e->prm;
    load_source(e, y, ys, u, v, cs);
    int cut = 0;
    if (e->frame_no > 0) {          /* coarse motion search against the previous source picture; also measures scene cuts */
        coarse_search(e);
        cut = p->scenecut && scene_cut(e) && e->poc + 1 >= p->min_keyint;
    }
    const int idr = force_idr || e->frame_no == 0 || (p->keyint > 0 && e->poc + 1 >= p->keyint) || cut;
    if (idr) e->poc = 0; else e->poc++;
    int qp = clampi(orc_rc_pick_qp(&e->rc, p, idr), 0, 51);
    if (idr) { intra_search_all(e, NULL, 0); encode_intra_frame(e, qp); } else encode_inter_frame(e, qp);
    long long est16 = 0;
    for (int i = 0; i < e->cuw * e->cuh; i++)
        est16 += orc_rc_cu_estimate(e->coefs + (size_t)i * ORC_CU_COEFS, e->cus[i].cbf);
    if (p->rate_control && idr && !e->rc.have[1]) {
        /* first key frame of a stream: no history to predict from, so it is coded twice when the first try overshoots */
        const long long budget = orc_rc_budget(&e->rc, 1);
        const int qp2 = clampi(qp + orc_rc_step(est16, budget), qp, 51);
        if (est16 > budget && qp2 != qp) {
            qp = qp2;
            encode_intra_frame(e, qp);
            est16 = 0;
            for (int i = 0; i < e->cuw * e->cuh; i++)
                est16 += orc_rc_cu_estimate(e->coefs + (size_t)i * ORC_CU_COEFS, e->cus[i].cbf);
        }
    }
    if (p->deblock) deblock_frame(e, work_planes(e), qp);
    if (p->sao) sao_frame(e, qp);
    plane *rec = e->rec[e->cur];
    for (int c = 0; c < 3; c++) plane_extend(&rec[c]);
    orc_rc_update(&e->rc, idr, qp, est16);

    orc_frame_syntax fs = {e->wc, e->hc, e->cuw, e->cuh, e->ctuw, e->ctuh, idr, qp, e->cus, e->coefs, p->sao ? e->sao : NULL, p->bit_depth};
    if (orc_cabac_encode_frame(&fs, e->payload, e->payload_cap, e->row_off, e->row_len) != 0)
        return -1;

    size_t o = 0;
    if (p->aud) o += orc_write_aud(idr ? 0 : 1, out + o, cap - o);
    if (idr && (e->frame_no == 0 || p->repeat_headers || force_idr)) o += orc_enc_headers(e, out + o, cap - o);
    if (p->hrd) {
        /* au_cpb_removal_delay counts from the most recent buffering period in a PRECEDING access unit (D.3.2) */
        const int delay = e->since_bp;
        if (idr) { o += orc_write_sei_buffering_period(p, out + o, cap - o); e->since_bp = 0; }
        o += orc_write_sei_pic_timing(p, delay > 0 ? delay : 1, out + o, cap - o);
        e->since_bp++;
    }
    if (idr && p->hdr10) o += orc_write_sei_hdr10(p, out + o, cap - o);

    /* slice NAL: header + escaped sub-streams (entry points count escaped bytes) */
    {
        size_t total = 0;
        uint32_t *entry = (uint32_t *)malloc(sizeof(uint32_t) * e->ctuh);
        uint8_t *esc = (uint8_t *)malloc(e->payload_cap + e->payload_cap / 2 + 64);
        for (int r = 0; r < e->ctuh; r++) {
            const size_t n = orc_escape(esc + total, e->payload_cap * 3 / 2 + 64 - total, e->payload + e->row_off[r], e->row_len[r]);
            entry[r] = (uint32_t)n;
            total += n;
        }
        uint8_t hdr[1024], hdr_esc[1600];
        const int nal_type = idr ? 19 : 1;      /* IDR_W_RADL : TRAIL_R */
        const size_t hn = orc_write_slice_header(p, nal_type, idr ? 2 : 1, e->poc, qp, entry, e->ctuh - 1, hdr, sizeof hdr);
        const size_t he = orc_escape(hdr_esc, sizeof hdr_esc, hdr, hn);
        if (o + 6 + he + total > cap) { free(entry); free(esc); return -1; }
        const int long_start = o == 0;
        if (long_start) out[o++] = 0;
        out[o++] = 0; out[o++] = 0; out[o++] = 1;
        out[o++] = (uint8_t)(nal_type << 1); out[o++] = 1;
        memcpy(out + o, hdr_esc, he); o += he;
        memcpy(out + o, esc, total); o += total;
        free(entry); free(esc);
    }
    if (p->hash_sei) {
        uint8_t md5[3][16];
        picture_md5(e, rec, md5);
        o += orc_write_sei_hash(md5, out + o, cap - o);
    }
    if (info) {
        info->is_idr = idr; info->poc = e->poc; info->qp = qp; info->bytes = (int)o; info->est_bits16 = est16;
        info->n_skip = info->n_merge = info->n_intra = 0;
        double sse = 0;
        for (int i = 0; i < e->cuw * e->cuh; i++) { info->n_skip += e->cus[i].skip; info->n_intra += e->cus[i].pred_mode == 0; }
        for (int yy = 0; yy < p->height; yy++)
            for (int xx = 0; xx < p->width; xx++) {
                const double d = (double)e->src[0].p[(size_t)yy * e->src[0].stride + xx] - rec[0].p[(size_t)yy * rec[0].stride + xx];
                sse += d * d;
            }
        const double peak = (double)((1 << p->bit_depth) - 1);
        info->psnr_y = sse > 0 ? 10.0 * log10(peak * peak * p->width * p->height / sse) : 99.0;
    }
    e->cur = 1 - e->cur;
    e->cur_ds = 1 - e->cur_ds;
    e->frame_no++;
    return (long)o;
}
