/*
 * ir_frame_oracle.c -- TEST INFRASTRUCTURE ONLY (groundwork for SURVEY.md section 8f rank 3).
 *
 * CPU restatement of the reference's frame classifier: access code, IBC header BCH(7,3),
 * de-interleaving, BCH(31,21) + parity with Chase decoding on the LLRs, IRA / IBC field extraction
 * (frame_decode.c:51-598).  Nothing in the product imports it; tests/test_frame_oracle.py pins it to the
 * reference's own frame_decode() compiled unmodified (oracle/_ref/libref_frame.so) on generated IRA / IBC
 * frames with and without bit errors.  No CUDA kernel consumes it yet: parity for this row is "oracle
 * pinned, device path not built".
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
