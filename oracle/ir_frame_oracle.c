/*
 * ir_frame_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's frame classifiers: access code, IBC header BCH(7,3),
 * de-interleaving, BCH(31,21) + parity with Chase decoding on the LLRs, IRA / IBC field extraction
 * (frame_decode.c:51-598), and further down the IDA burst decoder: link control word, payload
 * de-interleave, BCH(31,20)/Chase, header fields, CRC (ida_decode.c:33-396, 543-662).  Nothing in the
 * product imports it; tests/test_frame_oracle.py pins it to the reference's own frame_decode() and
 * ida_decode() compiled unmodified (oracle/_ref/libref_frame.so) on generated IRA / IBC / IDA frames with
 * and without bit errors.  It is the checker of k_classify_frames (tests/test_zz_gpu_classify.py) and of that
 * kernel's arithmetic compiled for the host (tests/test_frame_classify_host.py).
 *
 * Representation differs from the reference on purpose (it is what a bit-parallel device kernel would
 * use): a de-interleaved 32-bit block is one word, first bit in bit 31, so the 31-bit codeword is
 * w >> 1 and the overall parity bit is w & 1; corrections are XOR masks on that word.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct {
    int32_t ret;             /* frame_decode()'s return value: 1 decoded, 0 not */
    int32_t type;            /* frame_type_t: 0 unknown, 1 IRA, 2 IBC (frame_decode.h:20-24) */
    int32_t sat_id, beam_id;
    double lat, lon;
    int32_t alt;
    int32_t pos_xyz[3];
    int32_t n_pages;
    uint32_t tmsi[12];
    int32_t msc_id[12];
    int32_t timeslot, sv_blocking, bc_type;
    uint32_t iri_time;
} orc_frame_t;

enum { POLY_RA = 1207, POLY_HDR = 29, N_FLIP = 5 };   /* frame_decode.c:35-48 */

static const uint8_t k_access_dl[24] = {0,0,1,1,0,0,0,0,0,0,1,1,0,0,0,0,1,1,1,1,0,0,1,1};   /* :51-53 */
static const uint8_t k_access_ul[24] = {1,1,0,0,1,1,0,0,0,0,1,1,1,1,0,0,1,1,1,1,1,1,0,0};   /* :54-56 */

/* GF(2) remainder of v modulo poly (frame_decode.c:82-92) */
static uint32_t poly_rem(uint32_t poly, uint32_t v) {
    const int deg = 31 - __builtin_clz(poly);
    while (v >> deg) {
        const int top = 31 - __builtin_clz(v);
        v ^= poly << (top - deg);
    }
    return v;
}

/* syndrome -> (number of errors, XOR mask) for every pattern of up to `t` errors in n bits
 * (frame_decode.c:95-135: single errors first, a double only where no entry exists yet) */
typedef struct { int8_t errs; uint32_t mask; } syn_t;
static syn_t g_syn_ra[1024], g_syn_hdr[16];
static int g_ready;

static void fill_table(syn_t *tab, int size, uint32_t poly, int n, int t) {
    for (int i = 0; i < size; i++) { tab[i].errs = -1; tab[i].mask = 0; }
    for (int a = 0; a < n; a++) {
        const uint32_t m = 1u << a, r = poly_rem(poly, m);
        if (r < (uint32_t)size) { tab[r].errs = 1; tab[r].mask = m; }
    }
    if (t < 2) return;
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++) {
            const uint32_t m = (1u << a) | (1u << b), r = poly_rem(poly, m);
            if (r < (uint32_t)size && tab[r].errs < 0) { tab[r].errs = 2; tab[r].mask = m; }
        }
}

static void init_tables(void) {
    if (g_ready) return;
    fill_table(g_syn_ra, 1024, POLY_RA, 31, 2);      /* :133 */
    fill_table(g_syn_hdr, 16, POLY_HDR, 7, 1);       /* :134 */
    g_ready = 1;
}

static uint32_t take_bits(const uint8_t *b, int n) {            /* MSB first (:66-72) */
    uint32_t v = 0;
    for (int i = 0; i < n; i++) v = (v << 1) | (b[i] & 1u);
    return v;
}

/* Symbol s of a run of dibits is (in[2s], in[2s+1]).  de_interleave (:156-176) sends symbols
 * 31,29,..,1 to the first block and 30,28,..,0 to the second; de_interleave3 (:178-199) sends
 * 47,44,..,2 / 46,43,..,1 / 45,42,..,0.  `first`, `step`: first symbol and stride of one block. */
static uint32_t gather_block(const uint8_t *in, int first, int step) {
    uint32_t w = 0;
    for (int s = first; s >= 0; s -= step) w = (w << 2) | ((uint32_t)(in[2 * s] & 1u) << 1) | (in[2 * s + 1] & 1u);
    return w;                                                   /* 16 symbols: first bit ends up in bit 31 */
}
static void gather_llr(const float *in, int first, int step, float *out) {
    int p = 0;
    for (int s = first; s >= 0; s -= step) { out[p++] = in[2 * s]; out[p++] = in[2 * s + 1]; }
}

/* One 32-bit block: BCH(31,21) with up to 2 corrections, else Chase over the 5 least reliable of the
 * 31 code bits (frame_decode.c:224-295), then the overall parity (:399-408).  Returns the number of
 * errors the BCH step corrected (0..2) and the corrected codeword, or -1. */
static int decode_block(uint32_t w, const float *llr, uint32_t *code_out) {
    const uint32_t code = w >> 1;
    uint32_t s = poly_rem(POLY_RA, code);
    if (s == 0) { *code_out = code; return 0; }
    if (g_syn_ra[s].errs >= 0) { *code_out = code ^ g_syn_ra[s].mask; return g_syn_ra[s].errs; }
    if (!llr) return -1;
    /* the reference's partial selection sort, ties and all: position i takes the least reliable of what
     * is left, scanning the current arrangement front to back with a strict comparison, by swapping */
    int order[31];
    for (int i = 0; i < 31; i++) order[i] = i;
    for (int i = 0; i < N_FLIP; i++) {
        int m = i;
        for (int j = i + 1; j < 31; j++)
            if (llr[order[j]] < llr[order[m]]) m = j;
        const int t = order[i]; order[i] = order[m]; order[m] = t;
    }
    for (int sel = 1; sel < (1 << N_FLIP); sel++) {              /* subsets in counting order; first hit wins */
        uint32_t c = code;
        for (int b = 0; b < N_FLIP; b++)
            if (sel & (1 << b)) c ^= 1u << (30 - order[b]);     /* code bit k sits at bit 30-k */
        s = poly_rem(POLY_RA, c);
        if (s == 0) { *code_out = c; return 0; }
        if (g_syn_ra[s].errs >= 0) { *code_out = c ^ g_syn_ra[s].mask; return g_syn_ra[s].errs; }
    }
    return -1;
}
static int parity_ok(uint32_t w, uint32_t code) { return ((__builtin_popcount(code) + (int)(w & 1u)) & 1) == 0; }

static void put_data(uint8_t *stream, int *len, uint32_t code) {   /* the 21 data bits, MSB first */
    for (int i = 20; i >= 0; i--) stream[(*len)++] = (uint8_t)((code >> (10 + i)) & 1u);
}

static int sgn12(const uint8_t *b) {                              /* :299-307 */
    const int mag = (int)take_bits(b + 1, 11);
    return b[0] ? mag - 2048 : mag;
}

static void fields_ira(const uint8_t *d, int n, orc_frame_t *o) {  /* :317-366 */
    if (n < 63) return;
    o->sat_id = (int)take_bits(d, 7);
    o->beam_id = (int)take_bits(d + 7, 6);
    const int x = sgn12(d + 13), y = sgn12(d + 25), z = sgn12(d + 37);
    o->pos_xyz[0] = x; o->pos_xyz[1] = y; o->pos_xyz[2] = z;
    const double xy = sqrt((double)x * x + (double)y * y);
    o->lat = atan2((double)z, xy) * 180.0 / M_PI;
    o->lon = atan2((double)y, (double)x) * 180.0 / M_PI;
    o->alt = (int)(sqrt((double)x * x + (double)y * y + (double)z * z) * 4.0) - 6378 + 23;
    for (int off = 63; off + 42 <= n && o->n_pages < 12; off += 42) {
        const uint8_t *pg = d + off;
        int ones = 0;
        for (int i = 0; i < 42; i++) ones += pg[i] != 0;
        if (ones == 42) break;                                     /* all-ones terminator */
        o->tmsi[o->n_pages] = take_bits(pg, 32);
        o->msc_id[o->n_pages] = (int)take_bits(pg + 34, 5);
        o->n_pages++;
    }
}

static void fields_ibc(const uint8_t *d, int n, int hdr_type, orc_frame_t *o) {   /* :368-393 */
    o->bc_type = hdr_type;
    if (n < 42) return;
    o->sat_id = (int)take_bits(d, 7);
    o->beam_id = (int)take_bits(d + 7, 6);
    o->timeslot = d[14];
    o->sv_blocking = d[15];
    if (n >= 84 && take_bits(d + 42, 6) == 1) o->iri_time = take_bits(d + 52, 32);
}

/* pairs of blocks from `off` on while they decode and their parity holds (:495-516, :570-590) */
static void more_pairs(const uint8_t *data, const float *llr, int off, int limit, uint8_t *stream, int *len, int cap) {
    float l1[32], l2[32];
    while (off + 64 <= limit && *len + 42 <= cap) {
        const uint32_t w1 = gather_block(data + off, 31, 2), w2 = gather_block(data + off, 30, 2);
        if (llr) { gather_llr(llr + off, 31, 2, l1); gather_llr(llr + off, 30, 2, l2); }
        uint32_t c1, c2;
        if (decode_block(w1, llr ? l1 : 0, &c1) < 0 || decode_block(w2, llr ? l2 : 0, &c2) < 0) break;
        if (!parity_ok(w1, c1)) break;
        if (!parity_ok(w2, c2)) break;
        put_data(stream, len, c1);
        put_data(stream, len, c2);
        off += 64;
    }
}

/* frame_decode() (frame_decode.c:414-598): bits = one byte per bit, llr may be NULL */
int orc_frame_decode(const uint8_t *bits, const float *llr, int n_bits, orc_frame_t *o) {
    init_tables();
    memset(o, 0, sizeof(*o));
    if (n_bits < 24) return 0;
    if (memcmp(bits, k_access_dl, 24) != 0 && memcmp(bits, k_access_ul, 24) != 0) return 0;
    const uint8_t *data = bits + 24;
    const float *dl = llr ? llr + 24 : 0;
    const int n = n_bits - 24;

    /* ---- IBC: 6-bit header under BCH(7,3), then pairs of blocks (:440-523) */
    if (n >= 6 + 64) {
        uint32_t hv = take_bits(data, 6);
        const uint32_t hs = poly_rem(POLY_HDR, hv);
        int hdr_ok = hs == 0;
        if (!hdr_ok && hs < 16 && g_syn_hdr[hs].errs >= 0) { hv ^= g_syn_hdr[hs].mask; hdr_ok = 1; }
        if (hdr_ok) {
            float l1[32], l2[32];
            const uint32_t w1 = gather_block(data + 6, 31, 2), w2 = gather_block(data + 6, 30, 2);
            if (dl) { gather_llr(dl + 6, 31, 2, l1); gather_llr(dl + 6, 30, 2, l2); }
            uint32_t c1, c2;
            const int e1 = decode_block(w1, dl ? l1 : 0, &c1), e2 = decode_block(w2, dl ? l2 : 0, &c2);
            if (e1 >= 0 && e2 >= 0 && parity_ok(w1, c1) && parity_ok(w2, c2)) {
                uint8_t stream[256];
                int len = 0;
                put_data(stream, &len, c1);
                put_data(stream, &len, c2);
                more_pairs(data, dl, 6 + 64, n < 262 ? n : 262, stream, &len, (int)sizeof(stream));
                o->type = 2;
                fields_ibc(stream, len, (int)((hv >> 4) & 7u), o);
                return o->ret = 1;
            }
        }
    }
    /* ---- IRA: three header blocks from the first 96 bits, then pairs (:531-595) */
    if (n >= 96) {
        float l[3][32];
        uint32_t w[3], cw[3];
        int ok = 1;
        for (int k = 0; k < 3; k++) {
            w[k] = gather_block(data, 47 - k, 3);
            if (dl) gather_llr(dl, 47 - k, 3, l[k]);
        }
        for (int k = 0; k < 3; k++) ok = (decode_block(w[k], dl ? l[k] : 0, &cw[k]) >= 0) && ok;
        for (int k = 0; k < 3 && ok; k++) ok = parity_ok(w[k], cw[k]);
        if (ok) {
            uint8_t stream[512];
            int len = 0;
            for (int k = 0; k < 3; k++) put_data(stream, &len, cw[k]);
            more_pairs(data, dl, 96, n, stream, &len, (int)sizeof(stream));
            o->type = 1;
            fields_ira(stream, len, o);
            return o->ret = 1;
        }
    }
    return 0;
}

/* =========================================================================================
 * IDA bursts: link control word, payload de-interleave + BCH(31,20)/Chase, header fields, CRC
 * (ida_decode.c:33-396, 543-662).  Not restated: the LCW pretty-printer (ida_decode.c:398-541) and
 * the multi-burst reassembly (:669-748) -- host-side text and bookkeeping that would stay on the host.
 * ========================================================================================= */
typedef struct {
    int32_t ret;
    int32_t ft, lcw_ok, lcw_ft, lcw_code, ec_lcw;
    uint32_t lcw3_val;
    int32_t da_ctr, da_len, cont, payload_len, crc_ok, fixederrs, bch_len;
    uint16_t stored_crc, computed_crc;
    uint8_t payload[32];
    uint8_t bch_stream[256];
} orc_ida_t;

enum { POLY_DA = 3545, POLY_L1 = 29, POLY_L2 = 465, POLY_L3 = 41 };   /* ida_decode.c:34-43 */
static syn_t g_syn_da[2048], g_syn_l1[16], g_syn_l2[256], g_syn_l3[32];
static int g_ida_ready;

/* where each LCW bit comes from, 1-based, after the dibit swap (ida_decode.c:54-60) */
static const uint8_t k_lcw_from[46] = {40, 39, 36, 35, 32, 31, 28, 27, 24, 23, 20, 19, 16, 15, 12, 11, 8, 7, 4, 3,
                                       41, 38, 37, 34, 33, 30, 29, 26, 25, 22, 21, 18, 17, 14, 13, 10, 9, 6, 5, 2,
                                       1, 46, 45, 44, 43, 42};

static void init_ida_tables(void) {
    if (g_ida_ready) return;
    fill_table(g_syn_da, 2048, POLY_DA, 31, 2);      /* :98-101 */
    fill_table(g_syn_l1, 16, POLY_L1, 7, 1);
    fill_table(g_syn_l2, 256, POLY_L2, 14, 1);
    fill_table(g_syn_l3, 32, POLY_L3, 26, 2);
    g_ida_ready = 1;
}

/* one LCW component: zero syndrome, or a table correction, or failure (:216-243) */
static int lcw_fix(uint32_t *v, uint32_t poly, const syn_t *tab, int size, int *corrected) {
    const uint32_t s = poly_rem(poly, *v);
    *corrected = s != 0;
    if (s == 0) return 1;
    if (s >= (uint32_t)size || tab[s].errs < 0) return 0;
    *v ^= tab[s].mask;
    return 1;
}

/* BCH(31,20) block with Chase fall-back (:107-172): code = the 31 bits, first bit in bit 30.
 * *fixed = 1 whenever anything but a clean first look decoded it. */
static int decode_da(uint32_t code, const float *llr, uint32_t *out, int *fixed) {
    uint32_t s = poly_rem(POLY_DA, code);
    *fixed = 0;
    if (s == 0) { *out = code; return 0; }
    *fixed = 1;
    if (g_syn_da[s].errs >= 0) { *out = code ^ g_syn_da[s].mask; return g_syn_da[s].errs; }
    if (!llr) return -1;
    int order[31];
    for (int i = 0; i < 31; i++) order[i] = i;
    for (int i = 0; i < N_FLIP; i++) {
        int m = i;
        for (int j = i + 1; j < 31; j++)
            if (llr[order[j]] < llr[order[m]]) m = j;
        const int t = order[i]; order[i] = order[m]; order[m] = t;
    }
    for (int sel = 1; sel < (1 << N_FLIP); sel++) {
        uint32_t c = code;
        for (int b = 0; b < N_FLIP; b++)
            if (sel & (1 << b)) c ^= 1u << (30 - order[b]);
        s = poly_rem(POLY_DA, c);
        if (s == 0) { *out = c; return 0; }
        if (g_syn_da[s].errs >= 0) { *out = c ^ g_syn_da[s].mask; return g_syn_da[s].errs; }
    }
    return -1;
}

static void put_data20(uint8_t *stream, int *len, uint32_t code) {
    for (int i = 19; i >= 0; i--) stream[(*len)++] = (uint8_t)((code >> (11 + i)) & 1u);
}

/* two-way de-interleave of n_sym dibits (:259-272): odd symbols from the top down, then even ones */
static void split_halves(const uint8_t *in, const float *lin, int n_sym, uint8_t *h1, uint8_t *h2, float *l1, float *l2) {
    int p = 0;
    for (int s = n_sym - 1; s >= 1; s -= 2, p += 2) {
        h1[p] = in[2 * s]; h1[p + 1] = in[2 * s + 1];
        if (lin) { l1[p] = lin[2 * s]; l1[p + 1] = lin[2 * s + 1]; }
    }
    p = 0;
    for (int s = n_sym - 2; s >= 0; s -= 2, p += 2) {
        h2[p] = in[2 * s]; h2[p + 1] = in[2 * s + 1];
        if (lin) { l2[p] = lin[2 * s]; l2[p + 1] = lin[2 * s + 1]; }
    }
}

/* payload: 124-bit blocks of four code words taken in the order 4th, 2nd, 3rd, 1st, then a partial
 * block whose halves lose their first bit and swap places (:276-377) */
static int payload_stream(const uint8_t *data, const float *llr, int n, uint8_t *stream, int cap, int *fixederrs) {
    int len = 0;
    *fixederrs = 0;
    const int n_full = n / 124, rest = n % 124;
    static const int k_take[4] = {3, 1, 2, 0};
    for (int blk = 0; blk < n_full; blk++) {
        uint8_t bits[124];
        float rel[124];
        const float *bl = llr ? llr + blk * 124 : 0;
        split_halves(data + blk * 124, bl, 62, bits, bits + 62, rel, rel + 62);
        for (int c = 0; c < 4; c++) {
            if (len + 20 > cap) break;
            const int off = k_take[c] * 31;
            uint32_t out;
            int fixed;
            if (decode_da(take_bits(bits + off, 31), bl ? rel + off : 0, &out, &fixed) < 0) return len;
            *fixederrs += fixed;
            put_data20(stream, &len, out);
        }
    }
    if (rest >= 4 && len + 2 * (rest / 2 - 1) <= cap) {
        const int ns = rest / 2;
        /* (for an odd number of symbols the reference reads one element past what it wrote into its
         * first half -- uninitialised stack; zero here, and the tests keep to even counts) */
        uint8_t h1[64] = {0}, h2[64] = {0}, bits[128];
        float l1[64] = {0}, l2[64] = {0}, rel[128];
        const float *ll = llr ? llr + n_full * 124 : 0;
        split_halves(data + n_full * 124, ll, ns, h1, h2, l1, l2);
        if (ns > 1 && len + 20 <= cap) {
            int m = 0;
            for (int i = 1; i < ns && m < 128; i++, m++) { bits[m] = h2[i]; if (ll) rel[m] = l2[i]; }
            for (int i = 1; i < ns && m < 128; i++, m++) { bits[m] = h1[i]; if (ll) rel[m] = l1[i]; }
            for (int pos = 0; pos + 31 <= m && len + 20 <= cap; pos += 31) {
                uint32_t out;
                int fixed;
                if (decode_da(take_bits(bits + pos, 31), ll ? rel + pos : 0, &out, &fixed) < 0) break;
                *fixederrs += fixed;
                put_data20(stream, &len, out);
            }
        }
    }
    return len;
}

static uint16_t crc16_ccitt_false(const uint8_t *p, int n) {     /* :381-394 */
    uint16_t crc = 0xFFFF;
    for (int i = 0; i < n; i++) {
        crc ^= (uint16_t)(p[i] << 8);
        for (int k = 0; k < 8; k++) crc = (uint16_t)((crc & 0x8000) ? (crc << 1) ^ 0x1021 : crc << 1);
    }
    return crc;
}

/* ida_decode() (ida_decode.c:543-662) minus the fields copied from the frame and the LCW text;
 * direction: 1 downlink, 2 uplink (burst_downmix.h:32-36) */
int orc_ida_decode(const uint8_t *bits, const float *llr, int n_bits, int direction, orc_ida_t *o) {
    init_tables();
    init_ida_tables();
    memset(o, 0, sizeof(*o));
    if (n_bits < 24 + 46 + 124) return 0;
    if (direction != 1 && direction != 2) return 0;
    const uint8_t *data = bits + 24;
    const float *dl = llr ? llr + 24 : 0;
    const int n = n_bits - 24;

    /* ---- link control word (:193-255): dibit swap, permutation, three short BCH codes */
    uint8_t lb[46];
    for (int i = 0; i < 46; i++) lb[i] = data[(k_lcw_from[i] - 1) ^ 1];
    uint32_t v1 = take_bits(lb, 7), v2 = take_bits(lb + 7, 13) << 1, v3 = take_bits(lb + 20, 26);
    int c1, c2, c3;
    if (!lcw_fix(&v1, POLY_L1, g_syn_l1, 16, &c1)) return 0;
    if (!lcw_fix(&v2, POLY_L2, g_syn_l2, 256, &c2)) return 0;
    if (!lcw_fix(&v3, POLY_L3, g_syn_l3, 32, &c3)) return 0;
    const int ft = (int)(v1 >> 4) & 7, l2 = (int)(v2 >> 8) & 0x3f;
    if (ft != 2) return 0;

    if (n - 46 < 124) return 0;
    uint8_t stream[512];
    int fixederrs = 0;
    const int len = payload_stream(data + 46, dl ? dl + 46 : 0, n - 46, stream, (int)sizeof(stream), &fixederrs);
    if (len < 196) return 0;                                      /* 20 header + 160 payload + 16 CRC */
    const int da_len = (int)take_bits(stream + 11, 5);
    if (take_bits(stream + 17, 3) != 0 || da_len > 20) return 0;

    o->ft = ft; o->lcw_ok = 1; o->lcw_ft = (l2 >> 4) & 3; o->lcw_code = l2 & 15; o->lcw3_val = v3 >> 5;
    o->ec_lcw = c1 + c2 + c3;
    o->cont = stream[3];
    o->da_ctr = (int)take_bits(stream + 5, 3);
    o->da_len = da_len;
    o->fixederrs = fixederrs;
    o->payload_len = da_len > 0 ? da_len : 20;
    for (int i = 0; i < o->payload_len; i++) o->payload[i] = (uint8_t)take_bits(stream + 20 + 8 * i, 8);
    if (da_len > 0) {
        o->stored_crc = (uint16_t)take_bits(stream + 180, 16);
        /* CRC over: bits 0-19, twelve zero bits, bits 20 .. len-5, packed MSB first */
        const int crc_bits = 20 + 12 + (len - 24);
        uint8_t buf[64];
        if ((crc_bits + 7) / 8 <= (int)sizeof(buf)) {
            memset(buf, 0, sizeof(buf));
            int pos = 0;
            for (int i = 0; i < 20; i++, pos++) buf[pos >> 3] |= (uint8_t)(stream[i] << (7 - (pos & 7)));
            pos += 12;
            for (int i = 20; i < len - 4; i++, pos++) buf[pos >> 3] |= (uint8_t)(stream[i] << (7 - (pos & 7)));
            o->computed_crc = crc16_ccitt_false(buf, (pos + 7) / 8);
            o->crc_ok = o->computed_crc == 0;
        }
    }
    o->bch_len = len;
    memcpy(o->bch_stream, stream, len < 256 ? len : 256);
    return o->ret = 1;
}
