/* TEST INFRASTRUCTURE -- CPU oracle: bit writer, NAL layer, parameter sets, SEI, slice header, MD5.
 * Syntax follows Rec. ITU-T H.265 sections 7.3.1-7.3.6 and Annex D/E; the values signalled are the ones the
 * reference requests from libx265 (core/transcoder.py:398-412, core/utils.py:58-69). */
#include <string.h>

#include "hevc_model.h"

void orc_bits_init(orc_bits *b, uint8_t *buf, size_t cap)
{
    b->buf = buf; b->cap = cap; b->pos = 0; b->cur = 0; b->nbits = 0; b->overflow = 0;
}

void orc_put(orc_bits *b, uint32_t v, int n)
{
    while (n > 0) {
        int take = n > 8 ? 8 : n;                       /* feed at most a byte at a time */
        uint32_t part = (v >> (n - take)) & ((1u << take) - 1);
        b->cur = (b->cur << take) | part;
        b->nbits += take;
        n -= take;
        while (b->nbits >= 8) {
            if (b->pos < b->cap)
                b->buf[b->pos] = (uint8_t)(b->cur >> (b->nbits - 8));
            else
                b->overflow = 1;
            b->pos++;
            b->nbits -= 8;
        }
    }
}

void orc_put_ue(orc_bits *b, uint32_t v)
{
    uint64_t x = (uint64_t)v + 1;
    int len = 0;
    while ((x >> (len + 1)) != 0) len++;
    orc_put(b, 0, len);
    orc_put(b, (uint32_t)(x >> 16 >> 16), len + 1 > 32 ? len + 1 - 32 : 0);
    orc_put(b, (uint32_t)x, len + 1 > 32 ? 32 : len + 1);
}

void orc_put_se(orc_bits *b, int v) { orc_put_ue(b, v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v)); }

void orc_trailing(orc_bits *b)
{
    orc_put(b, 1, 1);
    if (b->nbits)
        orc_put(b, 0, 8 - b->nbits);
}

size_t orc_bits_flush(orc_bits *b)
{
    if (b->nbits)
        orc_put(b, 0, 8 - b->nbits);
    return b->overflow ? 0 : b->pos;
}

size_t orc_escape(uint8_t *out, size_t cap, const uint8_t *in, size_t n)
{
    size_t o = 0;
    int zeros = 0;
    for (size_t i = 0; i < n; i++) {
        if (zeros >= 2 && in[i] <= 3) {
            if (o < cap) out[o] = 3;
            o++;
            zeros = 0;
        }
        if (o < cap) out[o] = in[i];
        o++;
        zeros = in[i] == 0 ? zeros + 1 : 0;
    }
    return o;
}

size_t orc_write_nal(uint8_t *out, size_t cap, int nal_type, const uint8_t *rbsp, size_t n, int long_start)
{
    size_t o = 0;
    if (cap < n + n / 2 + 8)
        return 0;
    if (long_start) out[o++] = 0;
    out[o++] = 0; out[o++] = 0; out[o++] = 1;
    out[o++] = (uint8_t)(nal_type << 1);          /* forbidden_zero, type, layer id msb */
    out[o++] = 1;                                 /* layer id lsbs = 0, temporal_id_plus1 = 1 */
    o += orc_escape(out + o, cap - o, rbsp, n);
    return o;
}

/* ------------------------------------------------------------------ parameter sets */

static void profile_tier_level(orc_bits *b, const orc_enc_params *p)
{
    orc_put(b, 0, 2);                             /* general_profile_space */
    orc_put(b, p->tier, 1);
    orc_put(b, p->profile_idc, 5);
    for (int j = 0; j < 32; j++)                  /* Main streams are also Main10-decodable */
        orc_put(b, j == p->profile_idc || (p->profile_idc == 1 && j == 2), 1);
    orc_put(b, 1, 1);                             /* progressive_source */
    orc_put(b, 0, 1);                             /* interlaced_source */
    orc_put(b, 0, 1);                             /* non_packed_constraint */
    orc_put(b, 1, 1);                             /* frame_only_constraint */
    orc_put(b, 0, 32); orc_put(b, 0, 11);         /* 43 reserved zero bits */
    orc_put(b, 0, 1);                             /* inbld / reserved */
    orc_put(b, p->level_idc, 8);
}

size_t orc_write_vps(const orc_enc_params *p, uint8_t *out, size_t cap)
{
    uint8_t tmp[128];
    orc_bits b;
    orc_bits_init(&b, tmp, sizeof tmp);
    orc_put(&b, 0, 4);                            /* vps id */
    orc_put(&b, 3, 2);                            /* base_layer_internal, base_layer_available */
    orc_put(&b, 0, 6);                            /* max_layers_minus1 */
    orc_put(&b, 0, 3);                            /* max_sub_layers_minus1 */
    orc_put(&b, 1, 1);                            /* temporal_id_nesting */
    orc_put(&b, 0xffff, 16);
    profile_tier_level(&b, p);
    orc_put(&b, 1, 1);                            /* sub_layer_ordering_info_present */
    orc_put_ue(&b, 1);                            /* max_dec_pic_buffering_minus1 */
    orc_put_ue(&b, 0);                            /* max_num_reorder_pics */
    orc_put_ue(&b, 0);                            /* max_latency_increase_plus1 */
    orc_put(&b, 0, 6);                            /* max_layer_id */
    orc_put_ue(&b, 0);                            /* num_layer_sets_minus1 */
    orc_put(&b, 0, 1);                            /* timing_info_present */
    orc_put(&b, 0, 1);                            /* extension */
    orc_trailing(&b);
    return orc_write_nal(out, cap, 32, tmp, orc_bits_flush(&b), 1);
}

static void hrd_parameters(orc_bits *b, const orc_enc_params *p)
{
    orc_put(b, 1, 1);                             /* nal_hrd_parameters_present */
    orc_put(b, 0, 1);                             /* vcl_hrd_parameters_present */
    orc_put(b, 0, 1);                             /* sub_pic_hrd_params_present */
    orc_put(b, 0, 4);                             /* bit_rate_scale: units of 64 bit/s */
    orc_put(b, 0, 4);                             /* cpb_size_scale: units of 16 bits */
    orc_put(b, 23, 5);                            /* initial_cpb_removal_delay_length_minus1 */
    orc_put(b, 23, 5);                            /* au_cpb_removal_delay_length_minus1 */
    orc_put(b, 23, 5);                            /* dpb_output_delay_length_minus1 */
    orc_put(b, 0, 1);                             /* fixed_pic_rate_general */
    orc_put(b, 0, 1);                             /* fixed_pic_rate_within_cvs */
    orc_put(b, 0, 1);                             /* low_delay_hrd */
    orc_put_ue(b, 0);                             /* cpb_cnt_minus1 */
    orc_put_ue(b, (uint32_t)((int64_t)p->vbv_maxrate_kbps * 1000 / 64 - 1));
    orc_put_ue(b, (uint32_t)((int64_t)p->vbv_bufsize_kbit * 1000 / 16 - 1));
    orc_put(b, 0, 1);                             /* cbr_flag */
}

size_t orc_write_sps(const orc_enc_params *p, uint8_t *out, size_t cap)
{
    uint8_t tmp[256];
    orc_bits b;
    const int wc = (p->width + 15) & ~15, hc = (p->height + 15) & ~15;
    orc_bits_init(&b, tmp, sizeof tmp);
    orc_put(&b, 0, 4);                            /* vps id */
    orc_put(&b, 0, 3);                            /* max_sub_layers_minus1 */
    orc_put(&b, 1, 1);                            /* temporal_id_nesting */
    profile_tier_level(&b, p);
    orc_put_ue(&b, 0);                            /* sps id */
    orc_put_ue(&b, 1);                            /* chroma_format_idc 4:2:0 */
    orc_put_ue(&b, wc);
    orc_put_ue(&b, hc);
    if (wc != p->width || hc != p->height) {
        orc_put(&b, 1, 1);                        /* conformance window, chroma units */
        orc_put_ue(&b, 0); orc_put_ue(&b, (wc - p->width) / 2);
        orc_put_ue(&b, 0); orc_put_ue(&b, (hc - p->height) / 2);
    } else {
        orc_put(&b, 0, 1);
    }
    orc_put_ue(&b, p->bit_depth - 8);
    orc_put_ue(&b, p->bit_depth - 8);
    orc_put_ue(&b, 4);                            /* log2_max_pic_order_cnt_lsb_minus4 -> 8 bits */
    orc_put(&b, 1, 1);                            /* sub_layer_ordering_info_present */
    orc_put_ue(&b, 1); orc_put_ue(&b, 0); orc_put_ue(&b, 0);
    orc_put_ue(&b, 0);                            /* log2_min_luma_coding_block_size_minus3: 8 */
    orc_put_ue(&b, 2);                            /* log2_diff_max_min: CTU 32 */
    orc_put_ue(&b, 0);                            /* log2_min_luma_transform_block_size_minus2: 4 */
    orc_put_ue(&b, 3);                            /* log2_diff_max_min transform: 32 */
    orc_put_ue(&b, 0);                            /* max_transform_hierarchy_depth_inter */
    orc_put_ue(&b, 0);                            /* max_transform_hierarchy_depth_intra */
    orc_put(&b, 0, 1);                            /* scaling_list_enabled */
    orc_put(&b, 0, 1);                            /* amp_enabled */
    orc_put(&b, p->sao ? 1 : 0, 1);               /* sample_adaptive_offset_enabled */
    orc_put(&b, 0, 1);                            /* pcm_enabled */
    orc_put_ue(&b, 1);                            /* num_short_term_ref_pic_sets */
    orc_put_ue(&b, 1);                            /*   num_negative_pics */
    orc_put_ue(&b, 0);                            /*   num_positive_pics */
    orc_put_ue(&b, 0);                            /*   delta_poc_s0_minus1 */
    orc_put(&b, 1, 1);                            /*   used_by_curr_pic_s0 */
    orc_put(&b, 0, 1);                            /* long_term_ref_pics_present */
    orc_put(&b, 0, 1);                            /* sps_temporal_mvp_enabled */
    orc_put(&b, 0, 1);                            /* strong_intra_smoothing_enabled */
    orc_put(&b, 1, 1);                            /* vui_parameters_present */
    /* vui_parameters() */
    orc_put(&b, 1, 1); orc_put(&b, 1, 8);         /* aspect_ratio_info: square samples */
    orc_put(&b, 0, 1);                            /* overscan_info_present */
    orc_put(&b, 1, 1);                            /* video_signal_type_present */
    orc_put(&b, 5, 3);                            /*   video_format unspecified */
    orc_put(&b, p->full_range, 1);
    orc_put(&b, p->vui_colour ? 1 : 0, 1);
    if (p->vui_colour) {
        orc_put(&b, p->colour_primaries, 8);
        orc_put(&b, p->transfer_characteristics, 8);
        orc_put(&b, p->matrix_coeffs, 8);
    }
    if (p->chroma_loc >= 0) {
        orc_put(&b, 1, 1);
        orc_put_ue(&b, p->chroma_loc); orc_put_ue(&b, p->chroma_loc);
    } else {
        orc_put(&b, 0, 1);
    }
    orc_put(&b, 0, 1);                            /* neutral_chroma_indication */
    orc_put(&b, 0, 1);                            /* field_seq */
    orc_put(&b, 0, 1);                            /* frame_field_info_present */
    orc_put(&b, 0, 1);                            /* default_display_window */
    orc_put(&b, 1, 1);                            /* vui_timing_info_present */
    orc_put(&b, (uint32_t)p->fps_den, 32);
    orc_put(&b, (uint32_t)p->fps_num, 32);
    orc_put(&b, 0, 1);                            /* poc_proportional_to_timing */
    orc_put(&b, p->hrd ? 1 : 0, 1);               /* vui_hrd_parameters_present */
    if (p->hrd)
        hrd_parameters(&b, p);
    orc_put(&b, 0, 1);                            /* bitstream_restriction */
    orc_put(&b, 0, 1);                            /* sps_extension_present */
    orc_trailing(&b);
    return orc_write_nal(out, cap, 33, tmp, orc_bits_flush(&b), 1);
}

size_t orc_write_pps(const orc_enc_params *p, uint8_t *out, size_t cap)
{
    uint8_t tmp[64];
    orc_bits b;
    orc_bits_init(&b, tmp, sizeof tmp);
    orc_put_ue(&b, 0); orc_put_ue(&b, 0);         /* pps id, sps id */
    orc_put(&b, 0, 1);                            /* dependent_slice_segments_enabled */
    orc_put(&b, 0, 1);                            /* output_flag_present */
    orc_put(&b, 0, 3);                            /* num_extra_slice_header_bits */
    orc_put(&b, 0, 1);                            /* sign_data_hiding_enabled */
    orc_put(&b, 0, 1);                            /* cabac_init_present */
    orc_put_ue(&b, 0); orc_put_ue(&b, 0);         /* num_ref_idx_l0/l1_default_active_minus1 */
    orc_put_se(&b, 0);                            /* init_qp_minus26 */
    orc_put(&b, 0, 1);                            /* constrained_intra_pred */
    orc_put(&b, 0, 1);                            /* transform_skip_enabled */
    orc_put(&b, 0, 1);                            /* cu_qp_delta_enabled */
    orc_put_se(&b, 0); orc_put_se(&b, 0);         /* cb / cr qp offsets */
    orc_put(&b, 0, 1);                            /* slice_chroma_qp_offsets_present */
    orc_put(&b, 0, 1); orc_put(&b, 0, 1);         /* weighted_pred, weighted_bipred */
    orc_put(&b, 0, 1);                            /* transquant_bypass_enabled */
    orc_put(&b, 0, 1);                            /* tiles_enabled */
    orc_put(&b, 1, 1);                            /* entropy_coding_sync_enabled (WPP) */
    orc_put(&b, 0, 1);                            /* loop_filter_across_slices_enabled */
    orc_put(&b, 1, 1);                            /* deblocking_filter_control_present */
    orc_put(&b, 0, 1);                            /*   deblocking_filter_override_enabled */
    orc_put(&b, p->deblock ? 0 : 1, 1);           /*   pps_deblocking_filter_disabled */
    if (p->deblock) {
        orc_put_se(&b, 0); orc_put_se(&b, 0);     /*   beta_offset_div2, tc_offset_div2 */
    }
    orc_put(&b, 0, 1);                            /* scaling_list_data_present */
    orc_put(&b, 0, 1);                            /* lists_modification_present */
    orc_put_ue(&b, 0);                            /* log2_parallel_merge_level_minus2 */
    orc_put(&b, 0, 1);                            /* slice_segment_header_extension_present */
    orc_put(&b, 0, 1);                            /* pps_extension_present */
    orc_trailing(&b);
    return orc_write_nal(out, cap, 34, tmp, orc_bits_flush(&b), 1);
}

size_t orc_write_aud(int pic_type, uint8_t *out, size_t cap)
{
    uint8_t tmp[4];
    orc_bits b;
    orc_bits_init(&b, tmp, sizeof tmp);
    orc_put(&b, pic_type, 3);
    orc_trailing(&b);
    return orc_write_nal(out, cap, 35, tmp, orc_bits_flush(&b), 1);
}

static void sei_message(orc_bits *b, int type, const uint8_t *payload, int n)
{
    int t = type, s = n;
    while (t >= 255) { orc_put(b, 255, 8); t -= 255; }
    orc_put(b, t, 8);
    while (s >= 255) { orc_put(b, 255, 8); s -= 255; }
    orc_put(b, s, 8);
    for (int i = 0; i < n; i++)
        orc_put(b, payload[i], 8);
}

size_t orc_write_sei_hdr10(const orc_enc_params *p, uint8_t *out, size_t cap)
{
    uint8_t tmp[96], pl[32];
    orc_bits b, m;
    size_t o = 0;
    /* mastering display colour volume (137): G, B, R primaries, white point, max / min luminance */
    orc_bits_init(&m, pl, sizeof pl);
    for (int i = 0; i < 8; i++)
        orc_put(&m, p->master_display[i], 16);
    orc_put(&m, p->master_display[8], 32);
    orc_put(&m, p->master_display[9], 32);
    orc_bits_init(&b, tmp, sizeof tmp);
    sei_message(&b, 137, pl, (int)orc_bits_flush(&m));
    orc_trailing(&b);
    o += orc_write_nal(out + o, cap - o, 39, tmp, orc_bits_flush(&b), 0);
    /* content light level (144) */
    orc_bits_init(&m, pl, sizeof pl);
    orc_put(&m, p->max_cll, 16);
    orc_put(&m, p->max_fall, 16);
    orc_bits_init(&b, tmp, sizeof tmp);
    sei_message(&b, 144, pl, (int)orc_bits_flush(&m));
    orc_trailing(&b);
    o += orc_write_nal(out + o, cap - o, 39, tmp, orc_bits_flush(&b), 0);
    return o;
}

size_t orc_write_sei_buffering_period(const orc_enc_params *p, uint8_t *out, size_t cap)
{
    uint8_t tmp[64], pl[32];
    orc_bits b, m;
    /* initial removal delay at 90 % buffer fullness, 90 kHz units */
    const uint32_t delay = (uint32_t)((int64_t)90000 * 9 * p->vbv_bufsize_kbit / ((int64_t)10 * p->vbv_maxrate_kbps));
    orc_bits_init(&m, pl, sizeof pl);
    orc_put_ue(&m, 0);                            /* bp_seq_parameter_set_id */
    orc_put(&m, 0, 1);                            /* irap_cpb_params_present */
    orc_put(&m, 0, 1);                            /* concatenation_flag */
    orc_put(&m, 0, 24);                           /* au_cpb_removal_delay_delta_minus1 */
    orc_put(&m, delay, 24);                       /* nal_initial_cpb_removal_delay */
    orc_put(&m, 0, 24);                           /* nal_initial_cpb_removal_offset */
    if (m.nbits) orc_trailing(&m);                /* payload bit-alignment */
    orc_bits_init(&b, tmp, sizeof tmp);
    sei_message(&b, 0, pl, (int)orc_bits_flush(&m));
    orc_trailing(&b);
    return orc_write_nal(out, cap, 39, tmp, orc_bits_flush(&b), 0);
}

size_t orc_write_sei_pic_timing(const orc_enc_params *p, int cpb_removal_delay, uint8_t *out, size_t cap)
{
    uint8_t tmp[32], pl[16];
    orc_bits b, m;
    (void)p;
    orc_bits_init(&m, pl, sizeof pl);
    orc_put(&m, (uint32_t)(cpb_removal_delay > 0 ? cpb_removal_delay - 1 : 0), 24);   /* au_cpb_removal_delay_minus1 */
    orc_put(&m, 0, 24);                                                               /* pic_dpb_output_delay */
    orc_bits_init(&b, tmp, sizeof tmp);
    sei_message(&b, 1, pl, (int)orc_bits_flush(&m));
    orc_trailing(&b);
    return orc_write_nal(out, cap, 39, tmp, orc_bits_flush(&b), 0);
}

size_t orc_write_sei_hash(const uint8_t md5[3][16], uint8_t *out, size_t cap)
{
    uint8_t tmp[80], pl[49];
    orc_bits b;
    pl[0] = 0;                                    /* hash_type MD5 */
    memcpy(pl + 1, md5, 48);
    orc_bits_init(&b, tmp, sizeof tmp);
    sei_message(&b, 132, pl, 49);
    orc_trailing(&b);
    return orc_write_nal(out, cap, 40, tmp, orc_bits_flush(&b), 0);   /* suffix SEI */
}

size_t orc_write_slice_header(const orc_enc_params *p, int nal_type, int slice_type, int poc, int qp,
                              const uint32_t *entry, int n_entry, uint8_t *out, size_t cap)
{
    orc_bits b;
    orc_bits_init(&b, out, cap);
    orc_put(&b, 1, 1);                            /* first_slice_segment_in_pic */
    if (nal_type >= 16 && nal_type <= 23)
        orc_put(&b, 0, 1);                        /* no_output_of_prior_pics */
    orc_put_ue(&b, 0);                            /* pps id */
    orc_put_ue(&b, slice_type);                   /* 0 B, 1 P, 2 I */
    if (nal_type != 19 && nal_type != 20) {
        orc_put(&b, poc & 255, 8);                /* slice_pic_order_cnt_lsb */
        orc_put(&b, 1, 1);                        /* short_term_ref_pic_set_sps_flag (single set: no idx bits) */
    }
    if (p->sao) {
        orc_put(&b, 1, 1);                        /* slice_sao_luma_flag */
        orc_put(&b, 1, 1);                        /* slice_sao_chroma_flag */
    }
    if (slice_type == 1) {
        orc_put(&b, 0, 1);                        /* num_ref_idx_active_override */
        orc_put_ue(&b, 0);                        /* five_minus_max_num_merge_cand -> 5 candidates */
    }
    orc_put_se(&b, qp - 26);                      /* slice_qp_delta */
    orc_put_ue(&b, n_entry);                      /* num_entry_point_offsets */
    if (n_entry > 0) {
        uint32_t mx = 0;
        int len = 1;
        for (int i = 0; i < n_entry; i++)
            if (entry[i] - 1 > mx) mx = entry[i] - 1;
        while (len < 32 && (mx >> len) != 0) len++;
        orc_put_ue(&b, len - 1);                  /* offset_len_minus1 */
        for (int i = 0; i < n_entry; i++)
            orc_put(&b, entry[i] - 1, len);
    }
    orc_trailing(&b);                             /* byte_alignment() */
    return orc_bits_flush(&b);
}

/* ------------------------------------------------------------------ MD5 (RFC 1321) */

static uint32_t rol(uint32_t x, int c) { return (x << c) | (x >> (32 - c)); }

void orc_md5(const uint8_t *data, size_t n, uint8_t out[16])
{
    static const uint8_t S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
                                  4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
    static uint32_t K[64];
    static int init = 0;
    if (!init) {
        /* floor(2^32 * |sin(i + 1)|) without libm dependency on exactness: table values are standard */
        static const uint32_t k[64] = {
            0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1,
            0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453,
            0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942,
            0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05,
            0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d,
            0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
        memcpy(K, k, sizeof K);
        init = 1;
    }
    uint32_t h0 = 0x67452301, h1 = 0xefcdab89, h2 = 0x98badcfe, h3 = 0x10325476;
    const size_t total = ((n + 8) / 64 + 1) * 64;
    for (size_t off = 0; off < total; off += 64) {
        uint8_t blk[64];
        for (int i = 0; i < 64; i++) {
            const size_t idx = off + i;
            if (idx < n) blk[i] = data[idx];
            else if (idx == n) blk[i] = 0x80;
            else if (idx >= total - 8) blk[i] = (uint8_t)(((uint64_t)n * 8) >> (8 * (idx - (total - 8))));
            else blk[i] = 0;
        }
        uint32_t M[16], a = h0, b = h1, c = h2, d = h3;
        for (int i = 0; i < 16; i++)
            M[i] = (uint32_t)blk[4 * i] | ((uint32_t)blk[4 * i + 1] << 8) | ((uint32_t)blk[4 * i + 2] << 16) | ((uint32_t)blk[4 * i + 3] << 24);
        for (int i = 0; i < 64; i++) {
            uint32_t f;
            int g;
            if (i < 16) { f = (b & c) | (~b & d); g = i; }
            else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
            else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
            else { f = c ^ (b | ~d); g = (7 * i) & 15; }
            const uint32_t t = d;
            d = c; c = b;
            b = b + rol(a + f + K[i] + M[g], S[i]);
            a = t;
        }
        h0 += a; h1 += b; h2 += c; h3 += d;
    }
    const uint32_t h[4] = {h0, h1, h2, h3};
    for (int i = 0; i < 16; i++)
        out[i] = (uint8_t)(h[i / 4] >> (8 * (i % 4)));
}
