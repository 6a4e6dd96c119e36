// Entropy stage for sm_100a: per-CU syntax decisions (fully parallel), then WPP CABAC -- one CTA per frame,
// one warp per CTU row (rows interleaved over 32 warps), the 32 lanes stage coefficient levels and sub-block
// significance masks in shared memory while lane 0 runs the serial arithmetic coder.  Rows synchronise only
// through the context snapshot taken after the second CTU of the row above (H.265 9.3.2.2).
// Byte-identical to oracle/hevc_cabac.c.
#include "enc_kernels.cuh"

namespace hb {

// ------------------------------------------------------------------------------------------------ tables (H.265 9.3)
__constant__ uint8_t c_range_lps[64][4] = {
    {128, 176, 208, 240}, {128, 167, 197, 227}, {128, 158, 187, 216}, {123, 150, 178, 205}, {116, 142, 169, 195}, {111, 135, 160, 185},
    {105, 128, 152, 175}, {100, 122, 144, 166}, {95, 116, 137, 158}, {90, 110, 130, 150}, {85, 104, 123, 142}, {81, 99, 117, 135},
    {77, 94, 111, 128}, {73, 89, 105, 122}, {69, 85, 100, 116}, {66, 80, 95, 110}, {62, 76, 90, 104}, {59, 72, 86, 99}, {56, 69, 81, 94},
    {53, 65, 77, 89}, {51, 62, 73, 85}, {48, 59, 69, 80}, {46, 56, 66, 76}, {43, 53, 63, 72}, {41, 50, 59, 69}, {39, 48, 56, 65},
    {37, 45, 54, 62}, {35, 43, 51, 59}, {33, 41, 48, 56}, {32, 39, 46, 53}, {30, 37, 43, 50}, {29, 35, 41, 48}, {27, 33, 39, 45},
    {26, 31, 37, 43}, {24, 30, 35, 41}, {23, 28, 33, 39}, {22, 27, 32, 37}, {21, 26, 30, 35}, {20, 24, 29, 33}, {19, 23, 27, 31},
    {18, 22, 26, 30}, {17, 21, 25, 28}, {16, 20, 23, 27}, {15, 19, 22, 25}, {14, 18, 21, 24}, {14, 17, 20, 23}, {13, 16, 19, 22},
    {12, 15, 18, 21}, {12, 14, 17, 20}, {11, 14, 16, 19}, {11, 13, 15, 18}, {10, 12, 15, 17}, {10, 12, 14, 16}, {9, 11, 13, 15},
    {9, 11, 12, 14}, {8, 10, 12, 14}, {8, 9, 11, 13}, {7, 9, 11, 12}, {7, 9, 10, 12}, {7, 8, 10, 11}, {6, 8, 9, 11}, {6, 7, 9, 10},
    {6, 7, 8, 9}, {2, 2, 2, 2}};
__constant__ uint8_t c_next_lps[64] = {0, 0, 1, 2, 2, 4, 4, 5, 6, 7, 8, 9, 9, 11, 11, 12, 13, 13, 15, 15, 16, 16, 18, 18, 19, 19, 21, 21, 22, 22, 23, 24,
                                       24, 25, 26, 26, 27, 27, 28, 29, 29, 30, 30, 30, 31, 32, 32, 33, 33, 33, 34, 34, 35, 35, 35, 36, 36, 36, 37, 37, 37, 38, 38, 63};

enum {
    CX_SPLIT_CU = 0, CX_SKIP = 3, CX_PRED_MODE = 6, CX_PART_MODE = 7, CX_PREV_INTRA = 11, CX_CHROMA_PRED = 12, CX_MERGE_FLAG = 13,
    CX_MERGE_IDX = 14, CX_MVD_GR0 = 15, CX_MVD_GR1 = 16, CX_MVP_FLAG = 17, CX_ROOT_CBF = 18, CX_SPLIT_TU = 19, CX_CBF_LUMA = 22,
    CX_CBF_CHROMA = 24, CX_LAST_X = 28, CX_LAST_Y = 46, CX_CSBF = 64, CX_SIG = 68, CX_GR1 = 110, CX_GR2 = 134, CX_QP_DELTA = 140,
    CX_SAO_MERGE = 142, CX_SAO_TYPE = 143
};
static_assert(CX_SAO_TYPE + 1 == kNumCtx, "context layout");

// initValue per context for initType 0 (I) and 1 (P); H.265 Tables 9-5 .. 9-37
__constant__ uint8_t c_ctx_init[2][kNumCtx] = {
    {139, 141, 157, 154, 154, 154, 154, 184, 154, 154, 154, 184, 63, 154, 154, 154, 154, 154, 154, 153, 138, 138, 111, 141, 94, 138, 182, 154,
     110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63,
     110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63,
     91, 171, 134, 141,
     111, 111, 125, 110, 110, 94, 124, 108, 124, 107, 125, 141, 179, 153, 125, 107, 125, 141, 179, 153, 125, 107, 125, 141, 179, 153, 125,
     140, 139, 182, 182, 152, 136, 152, 136, 153, 136, 139, 111, 136, 139, 111,
     140, 92, 137, 138, 140, 152, 138, 139, 153, 74, 149, 92, 139, 107, 122, 152, 140, 179, 166, 182, 140, 227, 122, 197,
     138, 153, 136, 167, 152, 152, 154, 154, 153, 200},
    {107, 139, 126, 197, 185, 201, 149, 154, 139, 154, 154, 154, 152, 110, 122, 140, 198, 168, 79, 124, 138, 94, 153, 111, 149, 107, 167, 154,
     125, 110, 94, 110, 95, 79, 125, 111, 110, 78, 110, 111, 111, 95, 94, 108, 123, 108,
     125, 110, 94, 110, 95, 79, 125, 111, 110, 78, 110, 111, 111, 95, 94, 108, 123, 108,
     121, 140, 61, 154,
     155, 154, 139, 153, 139, 123, 123, 63, 153, 166, 183, 140, 136, 153, 154, 166, 183, 140, 136, 153, 154, 166, 183, 140, 136, 153, 154,
     170, 153, 123, 123, 107, 121, 107, 121, 167, 151, 183, 140, 151, 183, 140,
     154, 196, 196, 167, 154, 152, 167, 182, 182, 134, 149, 136, 153, 121, 136, 137, 169, 194, 166, 167, 154, 167, 137, 182,
     107, 167, 91, 122, 107, 167, 154, 154, 153, 185}};

// up-right diagonal scans as raster indices (y * size + x)
__constant__ uint8_t c_diag4[16] = {0, 4, 1, 8, 5, 2, 12, 9, 6, 3, 13, 10, 7, 14, 11, 15};
__constant__ uint8_t c_diag2[4] = {0, 2, 1, 3};
__constant__ uint8_t c_group_idx[32] = {0, 1, 2, 3, 4, 4, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9, 9, 9};
__constant__ uint8_t c_min_in_group[10] = {0, 1, 2, 3, 4, 6, 8, 12, 16, 24};

// ------------------------------------------------------------------------------------------------ syntax decisions
struct MvPair { int x, y; };

__device__ __forceinline__ bool inter_at(const ModeParams &p, int cx, int cy, int nx, int ny, MvPair &mv)
{
    if (!cu_avail(p.g, cx, cy, nx, ny)) return false;
    const CuInfo c = p.cus[ny * p.g.cuw + nx];
    if (c.pred_mode != 1) return false;
    mv.x = c.mvx; mv.y = c.mvy;
    return true;
}

__device__ __forceinline__ int mvd_bits(int d)
{
    const int a = abs(d);
    if (a == 0) return 1;
    if (a == 1) return 3;
    int v = a - 2, k = 1, bits = 3;
    while (v >= (1 << k)) { v -= 1 << k; k++; bits++; }
    return bits + 1 + k;
}

// one thread per CU: merge index / skip / AMVP predictor + difference (P), or MPM index / remaining mode (I)
__global__ void __launch_bounds__(256) k_modes(ModeParams p)
{
    const Geom &g = p.g;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.cuw * g.cuh) return;
    const int cx = idx % g.cuw, cy = idx / g.cuw;
    const CuInfo me = p.cus[idx];
    CuSyntax out;
    out.merge_idx = -1; out.skip = 0; out.mvp_idx = 0; out.pad = 0; out.mvdx = 0; out.mvdy = 0;
    if (p.ctl->is_idr || me.pred_mode == 0) {        // intra CU (I slice, or intra CU of a P slice): most-probable-mode signalling
        int a = 1, b = 1;
        if (cu_avail(g, cx, cy, cx - 1, cy) && p.cus[idx - 1].pred_mode == 0) a = p.cus[idx - 1].intra_mode;
        if ((cy & 1) && cu_avail(g, cx, cy, cx, cy - 1) && p.cus[idx - g.cuw].pred_mode == 0) b = p.cus[idx - g.cuw].intra_mode;
        int m0, m1, m2;
        if (a == b) {
            if (a < 2) { m0 = 0; m1 = 1; m2 = 26; }
            else { m0 = a; m1 = 2 + ((a + 29) & 31); m2 = 2 + ((a - 1) & 31); }
        } else {
            m0 = a; m1 = b;
            m2 = (a != 0 && b != 0) ? 0 : (a != 1 && b != 1) ? 1 : 26;
        }
        const int mode = me.intra_mode;
        if (mode == m0) out.merge_idx = 0;
        else if (mode == m1) out.merge_idx = 1;
        else if (mode == m2) out.merge_idx = 2;
        else {      // remaining mode: subtract the number of smaller candidates
            out.mvdx = (int16_t)(mode - (m0 < mode) - (m1 < mode) - (m2 < mode));
        }
        p.syn[idx] = out;
        return;
    }
    MvPair A1, B1, B0, A0, B2;
    const bool aA1 = inter_at(p, cx, cy, cx - 1, cy, A1), aB1 = inter_at(p, cx, cy, cx, cy - 1, B1);
    const bool aB0 = inter_at(p, cx, cy, cx + 1, cy - 1, B0), aA0 = inter_at(p, cx, cy, cx - 1, cy + 1, A0);
    const bool aB2 = inter_at(p, cx, cy, cx - 1, cy - 1, B2);
    auto same = [](const MvPair &u, const MvPair &v) { return u.x == v.x && u.y == v.y; };
    // merge list: A1, B1, B0, A0, B2 with the normative pruning, zero candidates after
    const bool fA1 = aA1, fB1 = aB1 && !(aA1 && same(A1, B1)), fB0 = aB0 && !(aB1 && same(B1, B0)), fA0 = aA0 && !(aA1 && same(A1, A0));
    const bool fB2 = aB2 && !(aA1 && same(A1, B2)) && !(aB1 && same(B1, B2)) && ((int)fA0 + fA1 + fB0 + fB1 != 4);
    const MvPair mine{me.mvx, me.mvy};
    int n = 0, found = -1;
    if (fA1) { if (found < 0 && same(A1, mine)) found = n; n++; }
    if (fB1) { if (found < 0 && same(B1, mine)) found = n; n++; }
    if (fB0) { if (found < 0 && same(B0, mine)) found = n; n++; }
    if (fA0) { if (found < 0 && same(A0, mine)) found = n; n++; }
    if (fB2) { if (found < 0 && same(B2, mine)) found = n; n++; }
    if (found < 0 && n < 5 && mine.x == 0 && mine.y == 0) found = n;
    if (found >= 0) {
        out.merge_idx = (int8_t)found;
        out.skip = me.cbf == 0;
    } else {
        // AMVP: A = first of (A0, A1), B = first of (B0, B1, B2); without A, A takes B's vector
        bool haveA = aA0 || aA1, haveB = aB0 || aB1 || aB2;
        MvPair a = aA0 ? A0 : A1, b = aB0 ? B0 : aB1 ? B1 : B2;
        if (!haveA && haveB) { a = b; haveA = true; }
        MvPair c[2] = {{0, 0}, {0, 0}};
        int k = 0;
        if (haveA) c[k++] = a;
        if (haveB && !(haveA && same(a, b))) c[k++] = b;
        const int b0 = mvd_bits(mine.x - c[0].x) + mvd_bits(mine.y - c[0].y);
        const int b1 = mvd_bits(mine.x - c[1].x) + mvd_bits(mine.y - c[1].y);
        const int sel = b1 < b0;
        out.mvp_idx = (uint8_t)sel;
        out.mvdx = (int16_t)(mine.x - c[sel].x);
        out.mvdy = (int16_t)(mine.y - c[sel].y);
    }
    p.syn[idx] = out;
}

// ------------------------------------------------------------------------------------------------ CABAC engine (lane 0)
// Per-warp scratch at file scope so that the (deliberately not inlined) coding functions address it as shared memory.
// per state: x = the four rangeTabLps bytes, y = the state after an LPS | the renormalisation shifts of the four LPS ranges << 8
// (four bits each; built once per CTA from the constant tables).  One struct, so every shared-space address derives from one base.
struct EntropyShared {
    EntropyWarpScratch ws[kEntropyWarps];
    uint2 state_tab[64];
};
__shared__ EntropyShared g_es;
#define g_ews g_es.ws
#define g_state_tab g_es.state_tab

// Arithmetic-coder state, passed and returned BY VALUE so that it lives in registers (a struct passed by reference to
// non-inlined functions sat in local memory and put local loads / stores on the critical path of every bin).
// x: low; y: range | bits_left << 16.
typedef uint2 CabacState;
__device__ __forceinline__ CabacState cs_make(uint32_t low, uint32_t range, int bits_left) { return make_uint2(low, range | ((uint32_t)bits_left << 16)); }

__device__ __forceinline__ void cb_byte(EntropyWarpScratch &o, uint32_t v)
{
    if (o.pos < o.cap) o.out[o.pos] = (uint8_t)v;
    o.pos++;
}

// one byte leaves the coder (bits_left < 12); the byte-output state stays in shared memory, it is touched once per 8 bits
__device__ __noinline__ CabacState cb_write_out(uint32_t low, uint32_t range, int bits_left, int w)
{
    EntropyWarpScratch &o = g_ews[w];
    const uint32_t lead = low >> (24 - bits_left);
    bits_left += 8;
    low &= 0xffffffffu >> bits_left;
    if (lead == 0xff) {
        o.buffered++;
    } else if (o.buffered > 0) {
        const uint32_t carry = lead >> 8;
        cb_byte(o, o.held + carry);
        o.held = lead & 0xff;
        const uint32_t fill = (0xff + carry) & 0xff;
        while (o.buffered > 1) { cb_byte(o, fill); o.buffered--; }
    } else {
        o.buffered = 1;
        o.held = lead;
    }
    return cs_make(low, range, bits_left);
}

__device__ __noinline__ uint32_t cb_finish(CabacState st, int w)
{
    EntropyWarpScratch &o = g_ews[w];
    uint32_t low = st.x;
    const int bits_left = (int)(st.y >> 16);
    if (low >> (32 - bits_left)) {
        cb_byte(o, o.held + 1);
        while (o.buffered > 1) { cb_byte(o, 0x00); o.buffered--; }
        low -= 1u << (32 - bits_left);
    } else {
        if (o.buffered > 0) cb_byte(o, o.held);
        while (o.buffered > 1) { cb_byte(o, 0xff); o.buffered--; }
    }
    int n = 24 - bits_left + 1;
    unsigned long long v = ((unsigned long long)(low >> 8) << 1) | 1;   // remaining bits + rbsp stop bit
    const int pad = (8 - (n & 7)) & 7;
    v <<= pad;
    n += pad;
    for (int i = n - 8; i >= 0; i -= 8) cb_byte(o, (uint32_t)(v >> i) & 0xff);
    return o.pos;
}

constexpr uint32_t kBypass = 254, kUnary = 253, kTerminate = 252;
__device__ __forceinline__ uint32_t bin_ctx(int ctx, int bin) { return (uint32_t)ctx | ((uint32_t)bin << 8); }
__device__ __forceinline__ uint32_t bin_bypass(uint32_t v, int n) { return kBypass | ((uint32_t)n << 8) | (v << 16); }
__device__ __forceinline__ uint32_t bin_unary(int ones) { return kUnary | ((uint32_t)ones << 8); }      // `ones` 1-bins, then a 0-bin
__device__ __forceinline__ uint2 ctx_entry(int state, int mps)
{
    const uint2 t = g_state_tab[state];
    return make_uint2(t.x, (t.y << 8) | ((uint32_t)state << 1) | (uint32_t)mps);
}

// Shared-memory accesses of the coding loop by 32-bit shared-space address: the generic-address forms recomputed the shared
// window base (S2UR SR_CgaCtaId + three uniform-datapath instructions) in every iteration, and the explicit order of the
// accesses is what the software pipeline below relies on.
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint2 lds64(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }

// Lane 0: the arithmetic coder proper, one tight loop over a bin list in shared memory (word offset `off` inside the warp's
// scratch), software-pipelined by hand: the list entry is fetched two bins ahead and the context entry of the NEXT bin one bin
// ahead (patched when it is the context this bin updates), and the table entry of the state this bin moves to is requested as
// soon as the bin value is compared with the MPS -- before the range arithmetic, which then covers its latency.  The MPS / LPS
// paths are one branch-free update (the LPS renormalisation shift comes out of the context entry, four bits per range quarter),
// so what stays on the critical path from bin to bin is the range / low recurrence alone: quarter select, LPS byte select,
// subtract, renormalise.  One step codes `cur`, refills it with the entry two places on and returns the context entry of `nxt`;
// the loop alternates the two registers so nothing is moved between iterations.
// (Entries are read up to two words past the end of a list and context entries are fetched for non-context kinds: both stay
// inside the warp's scratch and are never used.)
struct CoderRegs {
    uint32_t low, range;
    int bits_left;
};
__device__ __forceinline__ uint2 code_step(CoderRegs &r, uint32_t &cur, const uint32_t nxt, const uint2 c, uint32_t &ep, const uint32_t cbase,
                                           const uint32_t tbase, const int w)
{
    const uint32_t kind = cur & 0xff, arg = __byte_perm(cur, 0, 0x4441), bypass = cur >> 16, kn = nxt & 0xff;
    cur = lds32(ep);
    ep += 4;
    uint2 cn = lds64(cbase + (kn << 3));
    if (kind < kTerminate) {
        const uint32_t mps = c.y & 1, state = (c.y >> 1) & 63;
        const bool is_lps = arg != mps;
        const uint32_t next = is_lps ? __byte_perm(c.y, 0, 0x4441) : (state < 62 ? state + 1 : state);
        const uint32_t nmps = (is_lps && state == 0) ? mps ^ 1 : mps;
        const uint2 t = lds64(tbase + (next << 3));
        const uint32_t q = (r.range >> 6) & 3;
        const uint32_t lps = __byte_perm(c.x, 0, 0x4440u | q);
        const uint32_t nbl = (c.y >> (16 + 4 * q)) & 15;             // renormalisation shift of this LPS range: clz(lps) - 23 = 1..6
        const uint32_t rmps = r.range - lps;
        const uint32_t nb = is_lps ? nbl : (rmps < 256 ? 1u : 0u);
        r.low = (r.low + (is_lps ? rmps : 0u)) << nb;
        r.range = (is_lps ? lps : rmps) << nb;
        r.bits_left -= (int)nb;
        // (r.range >> 9 is always 0 -- a renormalised range is below 512 -- and makes the first use of the table entry depend on
        // the finished range update: without it the assembler schedules that use right behind the load, and the in-order issue
        // of the single active lane waits there for the whole shared-memory latency before the range arithmetic starts)
        const uint2 ne = make_uint2(t.x, (t.y * 256u + (r.range >> 9)) | (next << 1) | nmps);
        sts64(cbase + (kind << 3), ne);
        if (kn == kind) cn = ne;
        if (r.bits_left < 12) { const CabacState o = cb_write_out(r.low, r.range, r.bits_left, w); r.low = o.x; r.bits_left = (int)(o.y >> 16); }
    } else if (kind == kTerminate) {
        r.range -= 2;
        if (arg) {
            r.low = (r.low + r.range) << 7;
            r.range = 2 << 7;
            r.bits_left -= 7;
        } else if (r.range < 256) {
            r.low <<= 1; r.range <<= 1; r.bits_left--;
        }
        if (r.bits_left < 12) { const CabacState o = cb_write_out(r.low, r.range, r.bits_left, w); r.low = o.x; r.bits_left = (int)(o.y >> 16); }
    } else {
        // equiprobable bins, up to 8 at a time (H.265 9.3.4.3.4 applied c times: low = low * 2^c + range * bits)
        const uint32_t bits = kind == kBypass ? bypass : (2u << arg) - 2;
        int nbit = kind == kBypass ? (int)arg : (int)arg + 1;
        while (nbit > 0) {
            const int cnt = nbit > 8 ? 8 : nbit;
            nbit -= cnt;
            r.low = (r.low << cnt) + r.range * ((bits >> nbit) & ((1u << cnt) - 1));
            r.bits_left -= cnt;
            if (r.bits_left < 12) { const CabacState o = cb_write_out(r.low, r.range, r.bits_left, w); r.low = o.x; r.bits_left = (int)(o.y >> 16); }
        }
    }
    return cn;
}

// `esb`: shared-space address of the CTA's entropy scratch (computed once per kernel: taking the address of a shared variable
// costs an S2UR and three uniform-datapath instructions wherever the compiler rematerialises it)
__device__ __noinline__ CabacState code_list(CabacState st, uint32_t esb, int w, int off, int n)
{
    if (n <= 0) return st;
    const uint32_t sbase = esb + (uint32_t)(offsetof(EntropyShared, ws) + w * sizeof(EntropyWarpScratch));
    const uint32_t tbase = esb + (uint32_t)offsetof(EntropyShared, state_tab);
    const uint32_t cbase = sbase + (uint32_t)offsetof(EntropyWarpScratch, ctx);
    uint32_t ep = sbase + 4u * (uint32_t)off;
    CoderRegs r{st.x, st.y & 0xffffu, (int)(st.y >> 16)};
    uint32_t v0 = lds32(ep), v1 = lds32(ep + 4);
    uint2 c0 = lds64(cbase + ((v0 & 0xff) << 3)), c1;
    ep += 8;
    for (;;) {
        c1 = code_step(r, v0, v1, c0, ep, cbase, tbase, w);
        if (--n == 0) break;
        c0 = code_step(r, v1, v0, c1, ep, cbase, tbase, w);
        if (--n == 0) break;
    }
    return cs_make(r.low, r.range, r.bits_left);
}

#define EWS_OFF(member) ((int)(offsetof(EntropyWarpScratch, member) / 4))

// lane 0: everything CU `k` of the current CTU contributes: header list, then per coded transform block its last-position list
// and the sub-block lists in coding order
__device__ __noinline__ CabacState code_cu(CabacState st, uint32_t esb, int w, int k, int cbf)
{
    EntropyWarpScratch &s = g_ews[w];
    st = code_list(st, esb, w, EWS_OFF(hdr) + k * kHdrBins, s.nhdr[k]);
    for (int tu = 0; tu < 3; tu++) {
        if (!((cbf >> tu) & 1)) continue;
        const int base = tu == 0 ? 0 : 12 + 4 * tu;
        st = code_list(st, esb, w, EWS_OFF(tuh) + tu * kTuHdrBins, s.ntuh[tu]);
        for (int i = s.last_sb[tu]; i >= 0; i--) st = code_list(st, esb, w, EWS_OFF(bins) + (base + i) * kBinStride, s.nbins[base + i]);
    }
    return st;
}

// lane 0: one stand-alone list entry (end_of_slice_segment_flag / end_of_subset_one_bit)
__device__ __noinline__ CabacState code_terminate(CabacState st, uint32_t esb, int w, int bin)
{
    EntropyWarpScratch &s = g_ews[w];
    s.hdr[0][0] = kTerminate | ((uint32_t)bin << 8);
    return code_list(st, esb, w, EWS_OFF(hdr), 1);
}

// ------------------------------------------------------------------------------------------------ residual coding
// Two stages per coded CU.  Binarisation (lanes 0-23, one 4x4 sub-block each, lane = TU base + sub-block scan index): every
// lane turns its sub-block into a list of bins in shared memory -- context index + value for context-coded bins, (count, bits)
// for bypass runs.  All context selection (9.3.4.2.4 - 9.3.4.2.7) happens here, in parallel.  Coding (lane 0): walks the lists
// in coding order through one tight arithmetic-coder loop.  Bin entry: bits 0-7 kind (context index, kBypass or kUnary),
// bits 8-15 bin value / bin count, bits 16-31 bypass bits.
// Binarise the three transform blocks of the CU whose levels sit in s.lv[buf] and whose masks are in s.masks.  Called by the
// whole warp; returns (per lane, lane-uniform within a TU) the mask of coded sub-blocks in scan order for lanes of that TU.
__device__ __forceinline__ void binarise_cu(EntropyWarpScratch &s, int buf, int lane)
{
    const int tu = lane < 16 ? 0 : lane < 20 ? 1 : 2, base = tu == 0 ? 0 : tu == 1 ? 16 : 20;
    const int log2n = tu == 0 ? 4 : 3, n = 1 << log2n, sbw = n >> 2, nsb = sbw * sbw, c_idx = tu;
    const int i = lane - base;                                   // scan index of this lane's sub-block
    const bool lane_ok = lane < 24;
    const int sr = lane_ok ? (sbw == 4 ? c_diag4[i] : c_diag2[i]) : 0;
    const uint16_t *masks = s.masks + base;
    const int16_t *lv = s.lv[buf] + (tu == 0 ? 0 : 256 + (tu - 1) * 64);
    const int mask = lane_ok ? masks[sr] : 0;
    const uint32_t coded_all = __ballot_sync(0xffffffffu, mask != 0);
    const uint32_t tmask = (coded_all >> base) & ((1u << nsb) - 1);
    const int last_sb = tmask ? 31 - __clz(tmask) : -1;
    const bool active = lane_ok && i <= last_sb;
    uint32_t *out = s.bins[lane_ok ? lane : 0];
    int nb = 0;
    const int xs = sr % sbw, ys = sr / sbw;
    int cnt = 0, first_g1 = -1, a_first_g1 = 0, start = 15;
    uint32_t signs = 0;
    bool any_g1 = false;
    if (active) {
        const int last_sr = sbw == 4 ? c_diag4[last_sb] : c_diag2[last_sb];
        // coded_sub_block_flag of the right / below neighbours: coded (or inferred) sub-blocks are exactly those with a
        // non-zero mask, plus the DC sub-block and the last one which are inferred 1
        auto csbf_of = [&](int x, int y) -> int {
            if (x >= sbw || y >= sbw) return 0;
            const int r = y * sbw + x;
            return masks[r] != 0 || r == 0 || r == last_sr;
        };
        const int right = csbf_of(xs + 1, ys), below = csbf_of(xs, ys + 1);
        int infer_dc = 0;
        if (i < last_sb && i > 0) {
            out[nb++] = bin_ctx(CX_CSBF + ((right | below) ? 1 : 0) + (c_idx ? 2 : 0), mask != 0);
            infer_dc = 1;
        }
        if (mask || i == 0) {            // the DC sub-block is inferred coded: all-zero it still carries its 16 flags
            const int last_pos = 31 - __clz(mask);
            if (i == last_sb) start = last_pos;
            const int prev_csbf = right | (below << 1);
            for (int k = start; k >= 0; k--) {
                const int pr = c_diag4[k], xp = pr & 3, yp = pr >> 2;
                const int x = (xs << 2) + xp, y = (ys << 2) + yp;
                const int sigf = (mask >> k) & 1;
                const bool is_last = i == last_sb && k == last_pos;
                if (!is_last && (k > 0 || !infer_dc)) {
                    int sig;
                    if (x == 0 && y == 0) {
                        sig = 0;
                    } else {
                        if (prev_csbf == 0) sig = (xp + yp == 0) ? 2 : (xp + yp < 3) ? 1 : 0;
                        else if (prev_csbf == 1) sig = yp == 0 ? 2 : yp == 1 ? 1 : 0;
                        else if (prev_csbf == 2) sig = xp == 0 ? 2 : xp == 1 ? 1 : 0;
                        else sig = 2;
                        if (c_idx == 0) {
                            if (xs > 0 || ys > 0) sig += 3;
                            sig += log2n == 3 ? 9 : 21;
                        } else {
                            sig += log2n == 3 ? 9 : 12;
                        }
                    }
                    out[nb++] = bin_ctx(CX_SIG + (c_idx == 0 ? sig : 27 + sig), sigf);
                    if (sigf) infer_dc = 0;
                }
                if (sigf) {
                    const int v = lv[y * n + x];
                    if (cnt < 8 && abs(v) > 1) {
                        any_g1 = true;
                        if (first_g1 < 0) { first_g1 = cnt; a_first_g1 = abs(v); }
                    }
                    cnt++;
                    signs = (signs << 1) | (v < 0 ? 1u : 0u);
                }
            }
        }
    }
    // greater1 context set: +1 when the previously coded sub-block (next higher scan index with levels) ended with
    // greater1Ctx == 0, i.e. had a level above 1 among its first eight
    const uint32_t g1_all = __ballot_sync(0xffffffffu, any_g1);
    if (active && cnt) {
        const uint32_t tu_lanes = ((1u << nsb) - 1) << base;
        const uint32_t higher = coded_all & tu_lanes & ~((2u << lane) - 1);
        const bool prev_zero = higher && ((g1_all >> (__ffs(higher) - 1)) & 1);
        const int ctx_set = ((i > 0 && c_idx == 0) ? 2 : 0) + (prev_zero ? 1 : 0);
        int greater1_ctx = 1, seen = 0;
        for (int k = start; k >= 0 && seen < 8; k--) {
            if (!((mask >> k) & 1)) continue;
            const int pr = c_diag4[k];
            const int a = abs((int)lv[((ys << 2) + (pr >> 2)) * n + (xs << 2) + (pr & 3)]);
            const int g1 = a > 1;
            out[nb++] = bin_ctx(CX_GR1 + (ctx_set << 2) + greater1_ctx + (c_idx ? 16 : 0), g1);
            if (g1) greater1_ctx = 0;
            else if (greater1_ctx > 0 && greater1_ctx < 3) greater1_ctx++;
            seen++;
        }
        // coeff_abs_level_greater2 (first level above 1 only), signs, then the remaining levels
        if (first_g1 >= 0) out[nb++] = bin_ctx(CX_GR2 + ctx_set + (c_idx ? 4 : 0), a_first_g1 > 2);
        out[nb++] = bin_bypass(signs, cnt);
        int rice = 0;
        seen = 0;
        for (int k = start; k >= 0; k--) {
            if (!((mask >> k) & 1)) continue;
            const int pr = c_diag4[k];
            const int a = abs((int)lv[((ys << 2) + (pr >> 2)) * n + (xs << 2) + (pr & 3)]);
            const int basev = seen < 8 ? (seen == first_g1 ? 3 : 2) : 1;
            if (a >= basev) {
                int value = a - basev;
                if (value < (3 << rice)) {
                    out[nb++] = kUnary | ((uint32_t)(value >> rice) << 8);
                    if (rice) out[nb++] = bin_bypass(value & ((1 << rice) - 1), rice);
                } else {
                    int len = rice;
                    value -= 3 << rice;
                    while (value >= (1 << len)) { value -= 1 << len; len++; }
                    out[nb++] = kUnary | ((uint32_t)(3 + len - rice) << 8);
                    out[nb++] = bin_bypass(value, len);
                }
                if (a > 3 * (1 << rice)) rice = rice < 4 ? rice + 1 : 4;
            }
            seen++;
        }
    }
    if (lane_ok) s.nbins[lane] = (uint8_t)nb;
}


// Lanes 28-30: last significant coefficient position of transform block `tu` (9.3.3.x prefixes context-coded, suffixes bypass)
__device__ __forceinline__ void binarise_last_pos(EntropyWarpScratch &s, int tu)
{
    const int base = tu == 0 ? 0 : 12 + 4 * tu, log2n = tu == 0 ? 4 : 3, c_idx = tu;
    const int n = 1 << log2n, sbw = n >> 2, nsb = sbw * sbw;
    const uint8_t *sbscan = sbw == 4 ? c_diag4 : c_diag2;
    const uint16_t *masks = s.masks + base;
    int last_sb = -1;
    for (int i = nsb - 1; i >= 0; i--)
        if (masks[sbscan[i]]) { last_sb = i; break; }
    s.last_sb[tu] = (int8_t)last_sb;
    uint32_t *out = s.tuh[tu];
    int nb = 0;
    if (last_sb >= 0) {
        const int last_pos = 31 - __clz((int)masks[sbscan[last_sb]]);
        const int sr = sbscan[last_sb], pr = c_diag4[last_pos];
        const int px = ((sr % sbw) << 2) + (pr & 3), py = ((sr / sbw) << 2) + (pr >> 2);
        const int gx = c_group_idx[px], gy = c_group_idx[py], cmax = c_group_idx[n - 1];
        int off, shift;
        if (c_idx == 0) { off = 3 * (log2n - 2) + ((log2n - 1) >> 2); shift = (log2n + 1) >> 2; }
        else { off = 15; shift = log2n - 2; }
        for (int i = 0; i < gx; i++) out[nb++] = bin_ctx(CX_LAST_X + off + (i >> shift), 1);
        if (gx < cmax) out[nb++] = bin_ctx(CX_LAST_X + off + (gx >> shift), 0);
        for (int i = 0; i < gy; i++) out[nb++] = bin_ctx(CX_LAST_Y + off + (i >> shift), 1);
        if (gy < cmax) out[nb++] = bin_ctx(CX_LAST_Y + off + (gy >> shift), 0);
        if (gx > 3) out[nb++] = bin_bypass(px - c_min_in_group[gx], (gx - 2) >> 1);
        if (gy > 3) out[nb++] = bin_bypass(py - c_min_in_group[gy], (gy - 2) >> 1);
    }
    s.ntuh[tu] = (uint8_t)nb;
}

// Lanes 24-27: syntax of CU `k` of the CTU up to (and including) its coded block flags
struct CuHeaderIn {
    CuInfo cu;
    CuSyntax sy;
    int is_intra, split_flag_coded, first_in_ctu, split_inc, skip_l, skip_a;
};
__device__ __forceinline__ void binarise_header(EntropyWarpScratch &s, int k, const CuHeaderIn &h)
{
    uint32_t *out = s.hdr[k];
    int nb = 0;
    const CuInfo &cu = h.cu;
    const CuSyntax &sy = h.sy;
    if (h.first_in_ctu && h.split_flag_coded) out[nb++] = bin_ctx(CX_SPLIT_CU + h.split_inc, 1);
    out[nb++] = bin_ctx(CX_SPLIT_CU, 0);
    const int cb_y = cu.cbf & 1, cb_u = (cu.cbf >> 1) & 1, cb_v = (cu.cbf >> 2) & 1;
    bool coded_residual = true;
    const bool intra_cu = h.is_intra || cu.pred_mode == 0;
    if (!h.is_intra && intra_cu) {
        // intra CU in a P slice: cu_skip_flag = 0, pred_mode_flag = 1 (no part_mode: the CU is larger than the minimum size)
        out[nb++] = bin_ctx(CX_SKIP + (h.skip_l ? 1 : 0) + (h.skip_a ? 1 : 0), 0);
        out[nb++] = bin_ctx(CX_PRED_MODE, 1);
    }
    if (!intra_cu) {
        out[nb++] = bin_ctx(CX_SKIP + (h.skip_l ? 1 : 0) + (h.skip_a ? 1 : 0), sy.skip);
        if (sy.merge_idx >= 0) {
            if (!sy.skip) {
                out[nb++] = bin_ctx(CX_PRED_MODE, 0);
                out[nb++] = bin_ctx(CX_PART_MODE, 1);
                out[nb++] = bin_ctx(CX_MERGE_FLAG, 1);
            }
            out[nb++] = bin_ctx(CX_MERGE_IDX, sy.merge_idx > 0);
            if (sy.merge_idx > 0) {                      // truncated unary, cMax 4: bins 1.. are bypass
                const int ones = sy.merge_idx - 1;
                if (sy.merge_idx < 4) out[nb++] = bin_unary(ones);
                else out[nb++] = bin_bypass(7, 3);
            }
            if (sy.skip) coded_residual = false;
        } else {
            out[nb++] = bin_ctx(CX_PRED_MODE, 0);
            out[nb++] = bin_ctx(CX_PART_MODE, 1);
            out[nb++] = bin_ctx(CX_MERGE_FLAG, 0);
            const int dx = sy.mvdx, dy = sy.mvdy, ax = abs(dx), ay = abs(dy);
            out[nb++] = bin_ctx(CX_MVD_GR0, ax > 0);
            out[nb++] = bin_ctx(CX_MVD_GR0, ay > 0);
            if (ax > 0) out[nb++] = bin_ctx(CX_MVD_GR1, ax > 1);
            if (ay > 0) out[nb++] = bin_ctx(CX_MVD_GR1, ay > 1);
            for (int comp = 0; comp < 2; comp++) {
                const int a = comp ? ay : ax, neg = (comp ? dy : dx) < 0;
                if (a == 0) continue;
                if (a > 1) {                            // abs_mvd_minus2: Exp-Golomb order 1
                    int v = a - 2, kk = 1;
                    while (v >= (1 << kk)) { v -= 1 << kk; kk++; }
                    out[nb++] = bin_unary(kk - 1);
                    out[nb++] = bin_bypass(v, kk);
                }
                out[nb++] = bin_bypass(neg, 1);
            }
            out[nb++] = bin_ctx(CX_MVP_FLAG, sy.mvp_idx);
            out[nb++] = bin_ctx(CX_ROOT_CBF, cu.cbf != 0);
            if (!cu.cbf) coded_residual = false;
        }
    } else {
        out[nb++] = bin_ctx(CX_PREV_INTRA, sy.merge_idx >= 0);
        if (sy.merge_idx >= 0) {
            if (sy.merge_idx == 0) out[nb++] = bin_bypass(0, 1);
            else out[nb++] = bin_bypass(2 | (sy.merge_idx > 1 ? 1 : 0), 2);
        } else {
            out[nb++] = bin_bypass((uint32_t)sy.mvdx & 31, 5);
        }
        out[nb++] = bin_ctx(CX_CHROMA_PRED, 0);
    }
    if (coded_residual) {
        out[nb++] = bin_ctx(CX_CBF_CHROMA, cb_u);
        out[nb++] = bin_ctx(CX_CBF_CHROMA, cb_v);
        if (intra_cu || cb_u || cb_v) out[nb++] = bin_ctx(CX_CBF_LUMA + 1, cb_y);
    }
    s.nhdr[k] = (uint8_t)nb;
}

// Lane 31: sao() of the CTU (7.3.8.3).  Merging is by identity with the left (else the upper) CTU's parameters.
__device__ __forceinline__ bool sao_same(const SaoCtu &a, const SaoCtu &b)
{
    const uint32_t *x = reinterpret_cast<const uint32_t *>(&a), *y = reinterpret_cast<const uint32_t *>(&b);
    return x[0] == y[0] && x[1] == y[1] && x[2] == y[2] && x[3] == y[3] && x[4] == y[4];
}
__device__ __forceinline__ void binarise_sao(EntropyWarpScratch &s, const SaoCtu &c, bool have_left, bool have_up, const SaoCtu &up, int bit_depth)
{
    uint32_t *out = s.saob;
    int nb = 0;
    const int cmax = (1 << (min(bit_depth, 10) - 5)) - 1;
    bool merged = false;
    if (have_left) {
        merged = sao_same(c, s.sao_left);
        out[nb++] = bin_ctx(CX_SAO_MERGE, merged);
    }
    if (!merged && have_up) {
        merged = sao_same(c, up);
        out[nb++] = bin_ctx(CX_SAO_MERGE, merged);
    }
    if (!merged) {
        for (int ci = 0; ci < 3; ci++) {
            const int g = ci ? 1 : 0, type = c.type[g];
            if (ci < 2) {
                out[nb++] = bin_ctx(CX_SAO_TYPE, type != 0);
                if (type) out[nb++] = bin_bypass(type == 2, 1);
            }
            if (!type) continue;
            for (int i = 0; i < 4; i++) {              // sao_offset_abs: truncated unary, bypass
                const int a = abs((int)c.offset[ci][i]);
                if (a < cmax) out[nb++] = bin_unary(a);
                else {
                    if (a > 16) out[nb++] = bin_bypass(0xffff, 16);
                    const int rest = a > 16 ? a - 16 : a;
                    out[nb++] = bin_bypass((1u << rest) - 1, rest);
                }
            }
            if (type == 1) {
                for (int i = 0; i < 4; i++)
                    if (c.offset[ci][i]) out[nb++] = bin_bypass(c.offset[ci][i] < 0, 1);
                out[nb++] = bin_bypass(c.band[ci], 5);
            } else if (ci < 2) {
                out[nb++] = bin_bypass(c.eo_class[g], 2);
            }
        }
    }
    s.nsao = (uint32_t)nb;
}

// ------------------------------------------------------------------------------------------------ WPP CABAC kernel
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// stage the CU records of CTU x of this row (and the syntax of the two CUs above it) into ring slot x % 3
__device__ __forceinline__ void stage_ctu(EntropyWarpScratch &s, const EntropyFrame &fr, const Geom &g, int row, int x, int lane)
{
    if (x >= g.ctuw) return;
    const int slot = x % 3;
    if (lane < 4) {
        const int cx = 2 * x + (lane & 1), cy = 2 * row + (lane >> 1);
        if (cx < g.cuw && cy < g.cuh) {
            cp_async8(&s.cu[slot][lane].info, fr.cus + cy * g.cuw + cx);
            cp_async8(&s.cu[slot][lane].syn, fr.syn + cy * g.cuw + cx);
        }
    } else if (lane < 6) {
        const int cx = 2 * x + (lane - 4), cy = 2 * row - 1;
        if (cx < g.cuw && cy >= 0) cp_async8(&s.above[slot][lane - 4], fr.syn + cy * g.cuw + cx);
    } else if (lane < 10 && fr.sao) {       // SAO parameters of this CTU (lanes 6, 7) and of the one above (lanes 8, 9): 2 x 16 bytes each
        const int half = lane & 1, up = lane >= 8;
        if (!up || row > 0) {
            const SaoCtu *src = fr.sao + (row - up) * g.ctuw + x;
            cp_async16(reinterpret_cast<uint4 *>(up ? &s.sao_up[slot] : &s.sao_cur[slot]) + half, reinterpret_cast<const uint4 *>(src) + half);
        }
    }
}

// first coded CU after (x, k) within CTU x and the (already staged) CTU x + 1; -1 if none
__device__ __forceinline__ int next_coded_cu(const EntropyWarpScratch &s, const Geom &g, int row, int x, int k)
{
    for (int t = k + 1; t < 8; t++) {
        const int xx = x + (t >> 2), kk = t & 3;
        if (xx >= g.ctuw) break;
        const int cx = 2 * xx + (kk & 1), cy = 2 * row + (kk >> 1);
        if (cx >= g.cuw || cy >= g.cuh) continue;
        if (s.cu[xx % 3][kk].info.cbf) return cy * g.cuw + cx;
    }
    return -1;
}

// grid = (ceil(ctuh / kEntropyWarps), frames); CTA = kEntropyWarps warps, one CTU row per warp.  The CTAs are deliberately small
// so that they fit beside the resident motion-search CTAs and really overlap the frame chain.  Rows hand the context snapshot
// down through global memory (sync area of the frame); CTAs of a grid are dispatched in index order, so the CTA owning row
// r-1 is always resident (or done) when row r waits for it.  Nothing on the per-CU path waits for global memory: CU records
// are staged two CTUs ahead and levels one coded CU ahead with cp.async.
__global__ void __launch_bounds__(kEntropyWarps * 32, 8) k_entropy(EntropyParams p)
{
    const Geom &g = p.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 64) {
        const int t = threadIdx.x;
        uint32_t lps4 = 0, nb4 = 0;
        for (int q = 0; q < 4; q++) {
            const uint32_t lps = c_range_lps[t][q];
            lps4 |= lps << (8 * q);
            nb4 |= (uint32_t)(__clz(lps) - 23) << (4 * q);
        }
        g_state_tab[t] = make_uint2(lps4, (uint32_t)c_next_lps[t] | (nb4 << 8));
    }
    uint32_t esb = (uint32_t)__cvta_generic_to_shared(&g_es);
    asm volatile("" : "+r"(esb));                     // opaque: keep it in a register instead of rematerialising the address
    __syncthreads();
    const EntropyFrame fr = p.frames[blockIdx.y];
    uint8_t *ctx_save = fr.ctx_save;                              // [ctuh][kNumCtx]
    volatile int *row_ready = fr.row_ready;                       // [ctuh], zeroed before the launch
    EntropyWarpScratch &s = g_ews[warp];
    const int w = warp;
    const int slice_intra = fr.ctl->is_idr;                       // slice type is decided on the device (scene cuts)
    const int init_type = slice_intra ? 0 : 1;

    const int row = blockIdx.x * kEntropyWarps + warp;
    if (row >= g.ctuh) return;
    stage_ctu(s, fr, g, row, 0, lane);
    stage_ctu(s, fr, g, row, 1, lane);
    cp_async_commit();
    // ---- context initialisation: fresh for row 0 (or 1-CTU-wide pictures), else the snapshot of the row above
    if (row == 0 || g.ctuw < 2) {
        const int q = min(max(fr.ctl->qp, 0), 51);
        for (int i = lane; i < kNumCtx; i += 32) {
            const int v = c_ctx_init[init_type][i];
            const int m = (v >> 4) * 5 - 45, n = ((v & 15) << 3) - 16;
            const int pre = min(max(((m * q) >> 4) + n, 1), 126);
            const int mps = pre > 63;
            s.ctx[i] = ctx_entry(mps ? pre - 64 : 63 - pre, mps);
        }
    } else {
        if (lane == 0) {
            unsigned ns = 100, spins = 0;     // back off: polling warps share issue slots with the coding warps
            while (!row_ready[row - 1]) {
                __nanosleep(ns);
                ns = ns < 1600 ? ns * 2 : 1600;
                if (++spins > (1u << 25)) __trap();       // ~1 minute: a lost hand-off must fail loudly, not hang the device
            }
        }
        __syncwarp();
        __threadfence();
        for (int i = lane; i < kNumCtx; i += 32) {
            const int b = __ldcg(ctx_save + (row - 1) * kNumCtx + i);
            s.ctx[i] = ctx_entry(b >> 1, b & 1);
        }
    }
    unsigned long long t_start = 0;
    const bool tracing = fr.trace != nullptr;                     // debug (HB_ENTROPY_TRACE): phase cycle counters of lane 0
    if (tracing && lane == 0) s.pad3[0] = s.pad3[1] = s.pad3[2] = 0;   // all-lane phase cycles, coding cycles, list entries (shared: no registers)
    if (fr.trace && lane == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    CabacState st = cs_make(0, 510, 23);
    if (lane == 0) { s.buffered = 0; s.held = 0xff; s.out = fr.out + (size_t)row * p.row_cap; s.pos = 0; s.cap = p.row_cap; }
    cp_async_wait<0>();
    __syncwarp();

    int pf_idx = -1, pf_buf = 0;          // CU whose levels are in flight / ready in s.lv[pf_buf]
    bool pf_before_stage = false;         // that copy was committed before the newest CTU staging group
    int left_skip0 = 0, left_skip1 = 0;   // skip flags of the two CUs left of the CTU
    for (int x = 0; x < g.ctuw; x++) {
        const int slot = x % 3;
        // CTU x and x + 1 are staged (the wait at the end of the previous iteration); start on x + 2
        stage_ctu(s, fr, g, row, x + 2, lane);
        cp_async_commit();
        pf_before_stage = true;
        const int cx0 = 2 * x, cy0 = 2 * row;
        // ---- lanes 24-27: syntax of the four CUs up to the coded block flags
        if (lane >= 24 && lane < 28) {
            const int k = lane - 24, cx = cx0 + (k & 1), cy = cy0 + (k >> 1);
            if (cx < g.cuw && cy < g.cuh) {
                CuHeaderIn h;
                h.cu = s.cu[slot][k].info;
                h.sy = s.cu[slot][k].syn;
                h.is_intra = slice_intra;
                h.first_in_ctu = k == 0;
                h.split_flag_coded = 32 * x + 32 <= g.wc && 32 * row + 32 <= g.hc;
                h.split_inc = (x > 0 ? 1 : 0) + (row > 0 ? 1 : 0);
                // skip-flag context: left / above CUs always precede this one in decoding order when they exist
                h.skip_l = (k & 1) ? s.cu[slot][k - 1].syn.skip : (x > 0 ? ((k >> 1) ? left_skip1 : left_skip0) : 0);
                h.skip_a = (k >> 1) ? s.cu[slot][k - 2].syn.skip : (row > 0 ? s.above[slot][k].skip : 0);
                binarise_header(s, k, h);
            }
        }
        if (lane == 31 && fr.sao) binarise_sao(s, s.sao_cur[slot], x > 0, row > 0, s.sao_up[slot], g.bit_depth);
        left_skip0 = s.cu[slot][1].syn.skip;
        left_skip1 = s.cu[slot][3].syn.skip;
        __syncwarp();
        if (fr.sao) {
            if (lane == 0) st = code_list(st, esb, w, EWS_OFF(saob), (int)s.nsao);
            if (lane < 8) reinterpret_cast<uint32_t *>(&s.sao_left)[lane] = reinterpret_cast<const uint32_t *>(&s.sao_cur[slot])[lane];
            __syncwarp();
        }
        for (int k = 0; k < 4; k++) {
            const int cx = cx0 + (k & 1), cy = cy0 + (k >> 1);
            if (cx >= g.cuw || cy >= g.cuh) continue;
            const int idx = cy * g.cuw + cx;
            const int cbf = s.cu[slot][k].info.cbf;
            // ---- all lanes: levels of this CU (normally already in flight), sub-block masks, next prefetch, binarisation
            if (tracing && lane == 0) s.pad2 = (uint32_t)clock();
            if (cbf) {
                int buf;
                if (pf_idx == idx) {
                    buf = pf_buf;
                    if (pf_before_stage) cp_async_wait<1>(); else cp_async_wait<0>();
                } else {
                    buf = pf_buf ^ 1;
                    const uint4 *src = reinterpret_cast<const uint4 *>(fr.coefs + (size_t)idx * kCuCoefs);
                    for (int i = lane; i < kCuCoefs / 8; i += 32) cp_async16(reinterpret_cast<uint4 *>(s.lv[buf]) + i, src + i);
                    cp_async_commit();
                    cp_async_wait<0>();
                }
                __syncwarp();
                if (lane < 24) {
                    const int16_t *blk;
                    int stride, sx, sy2;
                    if (lane < 16) { blk = s.lv[buf]; stride = 16; sx = (lane & 3) * 4; sy2 = (lane >> 2) * 4; }
                    else { const int q = lane - 16; blk = s.lv[buf] + 256 + (q >> 2) * 64; stride = 8; sx = (q & 1) * 4; sy2 = ((q >> 1) & 1) * 4; }
                    uint32_t m = 0;
#pragma unroll
                    for (int t = 0; t < 16; t++) {
                        const int pr = c_diag4[t];
                        m |= (blk[(sy2 + (pr >> 2)) * stride + sx + (pr & 3)] != 0 ? 1u : 0u) << t;
                    }
                    s.masks[lane] = (uint16_t)m;
                }
                __syncwarp();
                pf_idx = next_coded_cu(s, g, row, x, k);
                pf_buf = buf ^ 1;
                pf_before_stage = false;
                if (pf_idx >= 0) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(fr.coefs + (size_t)pf_idx * kCuCoefs);
                    for (int i = lane; i < kCuCoefs / 8; i += 32) cp_async16(reinterpret_cast<uint4 *>(s.lv[pf_buf]) + i, src + i);
                }
                cp_async_commit();
                binarise_cu(s, buf, lane);
                if (lane >= 28 && lane < 31 && ((cbf >> (lane - 28)) & 1)) binarise_last_pos(s, lane - 28);
                __syncwarp();
            }
            // ---- lane 0: the arithmetic coder over this CU's lists
            if (tracing && lane == 0) { const uint32_t t = (uint32_t)clock(); s.pad3[0] += t - s.pad2; s.pad2 = t; }
            if (lane == 0) st = code_cu(st, esb, w, k, cbf);
            __syncwarp();
            if (tracing && lane == 0) {
                s.pad3[1] += (uint32_t)clock() - s.pad2;
                uint32_t ne = s.nhdr[k];
                for (int tu = 0; tu < 3; tu++)
                    if ((cbf >> tu) & 1) {
                        ne += s.ntuh[tu];
                        for (int i = s.last_sb[tu]; i >= 0; i--) ne += s.nbins[(tu == 0 ? 0 : 12 + 4 * tu) + i];
                    }
                s.pad3[2] += ne;
            }
        }
        // ---- end of CTU
        if (x == 1) {       // snapshot for the row below (taken before the terminating bin, contexts only)
            for (int i = lane; i < kNumCtx; i += 32) ctx_save[row * kNumCtx + i] = (uint8_t)(s.ctx[i].y & 0xff);
            __threadfence();
            __syncwarp();
            if (lane == 0) row_ready[row] = 1;
        }
        if (lane == 0) {
            const bool last_in_pic = row == g.ctuh - 1 && x == g.ctuw - 1;
            st = code_terminate(st, esb, w, last_in_pic);
            if (x == g.ctuw - 1 && !last_in_pic) st = code_terminate(st, esb, w, 1);
        }
        // the staging of CTU x + 2 has had a whole CTU of time; the level prefetch (if any) is the only younger group
        if (pf_before_stage) cp_async_wait<0>(); else cp_async_wait<1>();
        __syncwarp();
    }
    if (lane == 0) {
        const uint32_t n = cb_finish(st, w);
        fr.row_len[row] = n;
        if (fr.trace) {
            unsigned long long t_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            unsigned long long *t = fr.trace + (size_t)kTraceWords * row;
            t[0] = t_start; t[1] = t_end; t[2] = s.pad3[0]; t[3] = s.pad3[1]; t[4] = s.pad3[2];
        }
        if (n > p.row_cap) atomicExch(p.overflow, 1);
    }
}

// ------------------------------------------------------------------------------------------------ compaction
// single CTA: exclusive prefix sum over all (frame, row) lengths
__global__ void __launch_bounds__(1024) k_pack_scan(PackParams p)
{
    __shared__ uint32_t part[1024];
    const int total = p.n_frames * p.rows, tid = threadIdx.x;
    const int per = (total + 1023) / 1024;
    uint32_t sum = 0;
    for (int i = tid * per; i < min(total, (tid + 1) * per); i++) sum += p.frames[i / p.rows].row_len[i % p.rows];
    part[tid] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const uint32_t v = tid >= off ? part[tid - off] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    uint32_t run = tid ? part[tid - 1] : 0;
    for (int i = tid * per; i < min(total, (tid + 1) * per); i++) {
        p.offsets[i] = run;
        run += p.frames[i / p.rows].row_len[i % p.rows];
    }
    if (tid == 1023) p.offsets[total] = part[1023];
}

// one CTA per (frame, row): contiguous copy of the sub-stream
__global__ void __launch_bounds__(128) k_pack_copy(PackParams p)
{
    const int i = blockIdx.x, f = i / p.rows, r = i % p.rows;
    const uint8_t *src = p.frames[f].out + (size_t)r * p.row_cap;
    const uint32_t n = p.frames[f].row_len[r];
    // a group whose sub-streams exceed the space reserved for it (or a row that overran its own buffer) is not copied: the
    // host sees offsets[total] > cap / the overflow flag and fails the call, and nothing is written outside the group's region
    if (n > p.row_cap || (unsigned long long)p.offsets[i] + n > p.cap) return;
    uint8_t *dst = p.packed + p.offsets[i];
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) dst[k] = src[k];
}

}  // namespace hb
