/* TEST INFRASTRUCTURE -- CPU oracle: VBV-constrained rate control of the encoder model.
 *
 * The reference asks libx265 for crf= plus vbv-maxrate= / vbv-bufsize= (core/transcoder.py:398-403, values from
 * calculate_dynamic_values :263-354).  x265's own controller is not available, so this defines a deterministic,
 * integer-only one both implementations share: the CRF-derived QP is a quality ceiling (never go
 * below it) and the QP is raised so that a leaky-bucket model of the decoder buffer does not underflow.  Frame sizes are
 * *estimated* from the quantised levels (so the GPU never has to wait for the entropy coder): within a few percent of
 * the real size (tools/calibrate_rc.py -> profiles/rc_estimate_calibration_r2.jsonl). */
#include <stdlib.h>

#include "hevc_model.h"

/* round(65536 * 2^(k/6)) */
static const unsigned k_pow2_sixth[37] = {65536, 73562, 82570, 92682, 104032, 116772, 131072, 147123, 165140, 185364, 208064, 233544, 262144,
                                          294247, 330281, 370728, 416128, 467088, 524288, 588493, 660561, 741455, 832255, 934175, 1048576,
                                          1176987, 1321123, 1482910, 1664511, 1868350, 2097152, 2353974, 2642246, 2965821, 3329021, 3736700,
                                          4194304};

/* QP cascade over P frames: every fourth P frame (poc % 4 == 0) is the anchor and is coded at the P-frame QP, the three between
 * anchors at +4 / +2 / +4.  One reference picture and no B frames leave this as the way to spend bits where they are inherited:
 * the anchors refresh the quality that the skipped and merged CUs of the following frames copy (the role the P / B QP ratio
 * plays in x265, pbratio).  Offsets are >= 0 and the anchor's is 0, so the controller's state stays in anchor terms. */
static const int8_t k_cascade[4] = {0, 4, 2, 4};
int orc_rc_cascade(const orc_rc *rc, int is_idr, int poc) { return !is_idr && rc->cascade ? k_cascade[poc & 3] : 0; }

void orc_rc_init(orc_rc *rc, const orc_enc_params *p)
{
    rc->poc = 0;
    rc->cascade = p->qp_cascade;
    rc->t16 = (long long)p->vbv_maxrate_kbps * 1000 * 16 * p->fps_den / p->fps_num;
    rc->b16 = (long long)p->vbv_bufsize_kbit * 1000 * 16;
    rc->fullness = rc->b16 * 9 / 10;
    rc->have[0] = rc->have[1] = 0;
    rc->qp_prev[0] = rc->qp_prev[1] = 0;
    rc->est_prev[0] = rc->est_prev[1] = 0;
}

/* QP change that maps a frame of size `est` onto `budget`, assuming the size halves every 6 QP steps */
int orc_rc_step(long long est, long long budget)
{
    if (budget < 1) budget = 1;
    if (est > budget) {
        for (int k = 1; k <= 36; k++)
            if (est * 65536 <= budget * (long long)k_pow2_sixth[k]) return k;
        return 36;
    }
    int j = 0;
    while (j < 12 && est * (long long)k_pow2_sixth[j + 1] <= budget * 65536) j++;
    return -j;
}

/* size budget of the next frame: an IDR may take half of what the buffer holds (at most 8 frame intervals); a P frame
 * gets the per-frame drain scaled by the buffer fullness relative to half full, within [T/2, 2T] */
long long orc_rc_budget(const orc_rc *rc, int is_idr)
{
    long long fill = rc->fullness + rc->t16;
    if (fill > rc->b16) fill = rc->b16;
    long long budget;
    if (is_idr) {
        budget = fill / 2;
        if (budget > 8 * rc->t16) budget = 8 * rc->t16;
    } else {
        budget = rc->t16 * fill / (rc->b16 / 2);
        if (budget < rc->t16 / 2) budget = rc->t16 / 2;
        if (budget > 2 * rc->t16) budget = 2 * rc->t16;
    }
    return budget;
}

/* QP of the next frame.  qp_prev / est_prev are kept in anchor terms (cascade offset removed, size scaled up by 2^(offset / 6)),
 * so the size expected of this frame is the remembered one scaled down by its own offset. */
int orc_rc_pick_qp(const orc_rc *rc, const orc_enc_params *p, int is_idr)
{
    const int off = orc_rc_cascade(rc, is_idr, is_idr ? 0 : rc->poc + 1);
    int base = (is_idr ? p->qp_i : p->qp_p) + off;
    if (base > 51) base = 51;
    if (!p->rate_control) return base;
    const long long budget = orc_rc_budget(rc, is_idr);
    const int t = is_idr ? 1 : 0;
    int qp = base;
    if (rc->have[t]) {
        const long long est = rc->est_prev[t] * 65536 / k_pow2_sixth[off];
        int step = orc_rc_step(est, budget);
        if (step < 0)                                  /* come down one step at a time and only with 25 % headroom */
            step = est * 5 <= budget * 4 ? -1 : 0;
        qp = rc->qp_prev[t] + step + off;
    } else if (!is_idr && rc->have[1]) {
        /* first P frame: start from the IDR's operating point (P frames are ~1/4 of an IDR at the same QP) */
        int step = orc_rc_step(rc->est_prev[1] / 4 * 65536 / k_pow2_sixth[off], budget);
        if (step < 0) step = 0;
        qp = rc->qp_prev[1] + (p->qp_p - p->qp_i) + step + off;
    }
    if (qp < base) qp = base;
    if (qp > 51) qp = 51;
    return qp;
}

void orc_rc_update(orc_rc *rc, int is_idr, int qp, long long est16)
{
    const int t = is_idr ? 1 : 0;
    rc->poc = is_idr ? 0 : rc->poc + 1;
    const int off = orc_rc_cascade(rc, is_idr, rc->poc);
    rc->fullness += rc->t16;
    if (rc->fullness > rc->b16) rc->fullness = rc->b16;
    rc->fullness -= est16;
    if (rc->fullness < 0) rc->fullness = 0;
    rc->have[t] = 1;
    rc->qp_prev[t] = qp - off;
    rc->est_prev[t] = est16 * k_pow2_sixth[off] / 65536;
}

long long orc_rc_cu_estimate(const int16_t *coef, int cbf)
{
    if (!cbf) return 80;
    long long nnz = 0, slog = 0, nsb = 0;
    for (int blk = 0; blk < 3; blk++) {
        const int n = blk == 0 ? 16 : 8;
        const int16_t *c = coef + (blk == 0 ? 0 : blk == 1 ? 256 : 320);
        for (int sy = 0; sy < n; sy += 4)
            for (int sx = 0; sx < n; sx += 4) {
                int any = 0;
                for (int y = 0; y < 4; y++)
                    for (int x = 0; x < 4; x++) {
                        int a = abs(c[(sy + y) * n + sx + x]);
                        if (a) {
                            any = 1;
                            nnz++;
                            while (a > 1) { slog++; a >>= 1; }
                        }
                    }
                nsb += any;
            }
    }
    return 47 * nnz + 22 * slog + 104 * nsb + 160;
}
