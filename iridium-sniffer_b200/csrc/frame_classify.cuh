// frame_classify.cuh -- frame classification + BCH/Chase decoding of demodulated frames, per frame:
// what frame_decode() (frame_decode.c:414-598: access code, IBC header, IRA / IBC blocks, field
// extraction) and ida_decode() (ida_decode.c:543-662: link control word, payload blocks, header
// fields, CRC) compute from demod_frame_t.bits / .llr.  The reference calls both on every frame
// (main.c:320-350), so one pass fills both halves of ir_frame_class_t.
//
// Written as __host__ __device__ code on purpose: k_classify.cu runs it one warp per frame on the
// GPU (frames are few and small next to the IQ stream; the work is integer/bit arithmetic on <= 16
// 32-bit blocks), and tests compile the very same header for the host to compare it with the CPU
// oracle and with the reference's own functions without a GPU.  The library itself never classifies a
// frame on the host (only the public bit helpers of frame_decode.h, host functions by contract, reuse
// fc_rem / fc_take / fc_fill in refapi_frames.cu).
//
// A de-interleaved 32-bit block is one word with its first bit in bit 31: the 31-bit code word is
// w >> 1, the overall parity bit w & 1, corrections are XOR masks.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/iridium_b200.h"

#ifdef __CUDACC__
#define IR_HD __host__ __device__ __forceinline__
#else
#define IR_HD inline
#endif

namespace ir {

// syndrome -> (errors, XOR mask) tables of the five BCH codes involved, built once on the host
struct FcSyn { int8_t errs; uint32_t mask; };
struct FcTables {
    FcSyn ra[1024];      // BCH(31,21), generator 1207, t = 2   (frame_decode.c:35,133)
    FcSyn hdr[16];       // BCH(7,3),   generator 29,   t = 1   (:36,134; also LCW part 1, ida_decode.c:40)
    FcSyn da[2048];      // BCH(31,20), generator 3545, t = 2   (ida_decode.c:34,98)
    FcSyn l2[256];       // 14-bit,     generator 465,  t = 1   (:41,100)
    FcSyn l3[32];        // 26-bit,     generator 41,   t = 2   (:42,101)
};

IR_HD int fc_top_bit(uint32_t v) {           // index of the highest set bit, v != 0
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}
IR_HD int fc_popc(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

// GF(2) remainder of v modulo the generator g of degree deg (frame_decode.c:82-92)
IR_HD uint32_t fc_rem(uint32_t g, int deg, uint32_t v) {
    while (v >> deg) v ^= g << (fc_top_bit(v) - deg);
    return v;
}

inline void fc_fill(FcSyn *tab, int size, uint32_t g, int deg, int n, int t) {      // frame_decode.c:95-129
    for (int i = 0; i < size; i++) { tab[i].errs = -1; tab[i].mask = 0; }
    for (int a = 0; a < n; a++) {
        const uint32_t m = 1u << a, r = fc_rem(g, deg, m);
        if (r < (uint32_t)size) { tab[r].errs = 1; tab[r].mask = m; }
    }
    if (t < 2) return;
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++) {
            const uint32_t m = (1u << a) | (1u << b), r = fc_rem(g, deg, m);
            if (r < (uint32_t)size && tab[r].errs < 0) { tab[r].errs = 2; tab[r].mask = m; }
        }
}
inline void fc_build_tables(FcTables &T) {
    fc_fill(T.ra, 1024, 1207, 10, 31, 2);
    fc_fill(T.hdr, 16, 29, 4, 7, 1);
    fc_fill(T.da, 2048, 3545, 11, 31, 2);
    fc_fill(T.l2, 256, 465, 8, 14, 1);
    fc_fill(T.l3, 32, 41, 5, 26, 2);
}

IR_HD uint32_t fc_take(const uint8_t *b, int n) {               // n bits, first bit most significant
    uint32_t v = 0;
    for (int i = 0; i < n; i++) v = (v << 1) | (b[i] & 1u);
    return v;
}

// 16 dibits (in[2s], in[2s+1]) for s = first, first-step, ... as one word, and their reliabilities
IR_HD uint32_t fc_gather(const uint8_t *in, int first, int step) {
    uint32_t w = 0;
    for (int s = first; s >= 0; s -= step) w = (w << 2) | ((uint32_t)(in[2 * s] & 1u) << 1) | (in[2 * s + 1] & 1u);
    return w;
}
IR_HD void fc_gather_llr(const float *in, int first, int step, float *out) {
    int p = 0;
    for (int s = first; s >= 0; s -= step) { out[p++] = in[2 * s]; out[p++] = in[2 * s + 1]; }
}

// Candidate `sel` of the Chase search over the five least reliable positions `order` (sel = 0: the word as
// received; bit b of sel flips position order[b]; code bit k sits at bit 30-k): does it decode, to what,
// with how many table corrections.
IR_HD bool fc_candidate(uint32_t code, const int *order, int sel, uint32_t g, int deg, const FcSyn *tab, uint32_t *out,
                        int *errs, bool *clean) {
    uint32_t c = code;
    for (int b = 0; b < 5; b++)
        if (sel & (1 << b)) c ^= 1u << (30 - order[b]);
    const uint32_t s = fc_rem(g, deg, c);
    *clean = s == 0;
    if (s == 0) { *out = c; *errs = 0; return true; }
    if (tab[s].errs >= 0) { *out = c ^ tab[s].mask; *errs = tab[s].errs; return true; }
    return false;
}

// One 31-bit code word: clean / table correction / Chase over the 5 least reliable bits
// (frame_decode.c:224-295, ida_decode.c:107-172).  Returns the table's error count (0 for a clean
// candidate) or -1; *touched = 1 unless the first look was clean.  The selection of the five positions
// and the order of the candidates are the reference's, so that ties and multiple decodable candidates
// resolve the same way: the reference tries the 31 non-empty subsets in counting order and keeps the first
// that decodes.
//
// On the device the whole warp runs the classifier in lock step on the same frame (same data, same
// branches), and here -- the only expensive step -- the lanes part ways: lane L tries subset L, a ballot
// collects who decoded, and "first in counting order" is the lowest set bit.  The host build runs the same
// selection over an array of 32 results, so that the CPU parity test covers the ballot logic too.
IR_HD int fc_decode31(uint32_t code, const float *llr, uint32_t g, int deg, const FcSyn *tab, uint32_t *out, int *touched) {
    int order[31];
    for (int i = 0; i < 5; i++) order[i] = 0;
    int errs;
    bool clean;
    if (fc_candidate(code, order, 0, g, deg, tab, out, &errs, &clean)) { *touched = !clean; return errs; }
    *touched = 1;
    if (!llr) return -1;
    for (int i = 0; i < 31; i++) order[i] = i;
    for (int i = 0; i < 5; i++) {              // the reference's partial selection sort, ties and all
        int m = i;
        for (int j = i + 1; j < 31; j++)
            if (llr[order[j]] < llr[order[m]]) m = j;
        const int t = order[i]; order[i] = order[m]; order[m] = t;
    }
#ifdef __CUDA_ARCH__
    const int lane = (int)(threadIdx.x & 31u);
    uint32_t mine = 0;
    int my_errs = 0;
    const bool hit = lane > 0 && fc_candidate(code, order, lane, g, deg, tab, &mine, &my_errs, &clean);
    const uint32_t who = __ballot_sync(0xffffffffu, hit);
    if (who == 0) return -1;
    const int first = __ffs((int)who) - 1;
    *out = __shfl_sync(0xffffffffu, mine, first);
    return __shfl_sync(0xffffffffu, my_errs, first);
#else
    uint32_t who = 0, cand[32];
    int cand_errs[32];
    for (int lane = 1; lane < 32; lane++)
        if (fc_candidate(code, order, lane, g, deg, tab, &cand[lane], &cand_errs[lane], &clean)) who |= 1u << lane;
    if (who == 0) return -1;
    const int first = __builtin_ffs((int)who) - 1;
    *out = cand[first];
    return cand_errs[first];
#endif
}

IR_HD bool fc_block_ra(const FcTables &T, uint32_t w, const float *llr, uint32_t *code) {   // BCH(31,21) + parity (:399-408)
    int touched;
    if (fc_decode31(w >> 1, llr, 1207, 10, T.ra, code, &touched) < 0) return false;
    return ((fc_popc(*code) + (int)(w & 1u)) & 1) == 0;
}
// (the reference decodes BOTH blocks of a pair before looking at either parity; no side effects, same outcome)

IR_HD void fc_put(uint8_t *stream, int *len, uint32_t code, int n_data, int n_check) {
    for (int i = n_data - 1; i >= 0; i--) stream[(*len)++] = (uint8_t)((code >> (n_check + i)) & 1u);
}

IR_HD int fc_s12(const uint8_t *b) { const int mag = (int)fc_take(b + 1, 11); return b[0] ? mag - 2048 : mag; }   // :299-307

// pairs of BCH(31,21) blocks from bit `off` on while both decode and both parities hold (:495-516, :570-590)
IR_HD void fc_pairs(const FcTables &T, const uint8_t *data, const float *llr, int off, int limit, uint8_t *stream, int *len, int cap) {
    float l1[32], l2[32];
    while (off + 64 <= limit && *len + 42 <= cap) {
        const uint32_t w1 = fc_gather(data + off, 31, 2), w2 = fc_gather(data + off, 30, 2);
        if (llr) { fc_gather_llr(llr + off, 31, 2, l1); fc_gather_llr(llr + off, 30, 2, l2); }
        uint32_t c1, c2;
        int t1, t2;
        const int e1 = fc_decode31(w1 >> 1, llr ? l1 : nullptr, 1207, 10, T.ra, &c1, &t1);
        const int e2 = fc_decode31(w2 >> 1, llr ? l2 : nullptr, 1207, 10, T.ra, &c2, &t2);
        if (e1 < 0 || e2 < 0) break;
        if (((fc_popc(c1) + (int)(w1 & 1u)) & 1) != 0) break;
        if (((fc_popc(c2) + (int)(w2 & 1u)) & 1) != 0) break;
        fc_put(stream, len, c1, 21, 10);
        fc_put(stream, len, c2, 21, 10);
        off += 64;
    }
}

// ---- frame_decode() (frame_decode.c:414-598): fills the IRA / IBC half of o, returns 1 if decoded
IR_HD int fc_frame(const FcTables &T, const uint8_t *bits, const float *llr, int n_bits, ir_frame_class_t *o) {
    const uint8_t dl_code[24] = {0,0,1,1,0,0,0,0,0,0,1,1,0,0,0,0,1,1,1,1,0,0,1,1};   // :51-53
    const uint8_t ul_code[24] = {1,1,0,0,1,1,0,0,0,0,1,1,1,1,0,0,1,1,1,1,1,1,0,0};   // :54-56
    if (n_bits < 24) return 0;
    bool is_dl = true, is_ul = true;
    for (int i = 0; i < 24; i++) { is_dl = is_dl && bits[i] == dl_code[i]; is_ul = is_ul && bits[i] == ul_code[i]; }
    if (!is_dl && !is_ul) return 0;
    const uint8_t *data = bits + 24;
    const float *dl = llr ? llr + 24 : nullptr;
    const int n = n_bits - 24;
    // IBC: 6-bit header under BCH(7,3), then pairs of blocks (:440-523)
    if (n >= 6 + 64) {
        uint32_t hv = fc_take(data, 6);
        const uint32_t hs = fc_rem(29, 4, hv);
        bool hdr_ok = hs == 0;
        if (!hdr_ok && hs < 16 && T.hdr[hs].errs >= 0) { hv ^= T.hdr[hs].mask; hdr_ok = true; }
        if (hdr_ok) {
            float l1[32], l2[32];
            const uint32_t w1 = fc_gather(data + 6, 31, 2), w2 = fc_gather(data + 6, 30, 2);
            if (dl) { fc_gather_llr(dl + 6, 31, 2, l1); fc_gather_llr(dl + 6, 30, 2, l2); }
            uint32_t c1, c2;
            const bool ok1 = fc_block_ra(T, w1, dl ? l1 : nullptr, &c1), ok2 = fc_block_ra(T, w2, dl ? l2 : nullptr, &c2);
            if (ok1 && ok2) {
                uint8_t stream[256];
                int len = 0;
                fc_put(stream, &len, c1, 21, 10);
                fc_put(stream, &len, c2, 21, 10);
                fc_pairs(T, data, dl, 6 + 64, n < 262 ? n : 262, stream, &len, 256);
                o->frame_type = IR_FRAME_IBC;
                o->bc_type = (int32_t)((hv >> 4) & 7u);                              // :368-393
                if (len >= 42) {
                    o->sat_id = (int32_t)fc_take(stream, 7);
                    o->beam_id = (int32_t)fc_take(stream + 7, 6);
                    o->timeslot = stream[14];
                    o->sv_blocking = stream[15];
                    if (len >= 84 && fc_take(stream + 42, 6) == 1) o->iri_time = fc_take(stream + 52, 32);
                }
                return 1;
            }
        }
    }
    // IRA: three header blocks out of the first 96 bits, then pairs (:531-595)
    if (n >= 96) {
        float l[3][32];
        uint32_t w[3], cw[3];
        bool ok = true;
        for (int k = 0; k < 3; k++) {
            w[k] = fc_gather(data, 47 - k, 3);
            if (dl) fc_gather_llr(dl, 47 - k, 3, l[k]);
        }
        for (int k = 0; k < 3; k++) { int t; ok = (fc_decode31(w[k] >> 1, dl ? l[k] : nullptr, 1207, 10, T.ra, &cw[k], &t) >= 0) && ok; }
        for (int k = 0; k < 3 && ok; k++) ok = ((fc_popc(cw[k]) + (int)(w[k] & 1u)) & 1) == 0;
        if (ok) {
            uint8_t stream[512];
            int len = 0;
            for (int k = 0; k < 3; k++) fc_put(stream, &len, cw[k], 21, 10);
            fc_pairs(T, data, dl, 96, n, stream, &len, 512);
            o->frame_type = IR_FRAME_IRA;                                            // :317-366
            o->sat_id = (int32_t)fc_take(stream, 7);
            o->beam_id = (int32_t)fc_take(stream + 7, 6);
            const int x = fc_s12(stream + 13), y = fc_s12(stream + 25), z = fc_s12(stream + 37);
            o->pos_xyz[0] = x; o->pos_xyz[1] = y; o->pos_xyz[2] = z;
            for (int off = 63; off + 42 <= len && o->n_pages < 12; off += 42) {
                const uint8_t *pg = stream + off;
                int ones = 0;
                for (int i = 0; i < 42; i++) ones += pg[i] != 0;
                if (ones == 42) break;
                o->tmsi[o->n_pages] = fc_take(pg, 32);
                o->msc_id[o->n_pages] = (int32_t)fc_take(pg + 34, 5);
                o->n_pages++;
            }
            return 1;
        }
    }
    return 0;
}

// ---- ida_decode() (ida_decode.c:543-662) minus the LCW text and the fields copied from the frame
IR_HD bool fc_lcw_part(uint32_t *v, uint32_t g, int deg, const FcSyn *tab, int size, int *corrected) {   // :216-243
    const uint32_t s = fc_rem(g, deg, *v);
    *corrected = s != 0;
    if (s == 0) return true;
    if (s >= (uint32_t)size || tab[s].errs < 0) return false;
    *v ^= tab[s].mask;
    return true;
}

IR_HD void fc_halves(const uint8_t *in, const float *lin, int n_sym, uint8_t *h1, uint8_t *h2, float *l1, float *l2) {   // :259-272
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

IR_HD int fc_ida(const FcTables &T, const uint8_t *bits, const float *llr, int n_bits, int direction, ir_frame_class_t *o) {
    const uint8_t from[46] = {40, 39, 36, 35, 32, 31, 28, 27, 24, 23, 20, 19, 16, 15, 12, 11, 8, 7, 4, 3,
                              41, 38, 37, 34, 33, 30, 29, 26, 25, 22, 21, 18, 17, 14, 13, 10, 9, 6, 5, 2,
                              1, 46, 45, 44, 43, 42};                                  // ida_decode.c:54-60
    if (n_bits < 24 + 46 + 124) return 0;
    if (direction != IR_DIR_DOWNLINK && direction != IR_DIR_UPLINK) return 0;
    const uint8_t *data = bits + 24;
    const float *dl = llr ? llr + 24 : nullptr;
    const int n = n_bits - 24;
    uint8_t lb[46];
    for (int i = 0; i < 46; i++) lb[i] = data[(from[i] - 1) ^ 1];                      // dibit swap + permutation (:199-212)
    uint32_t v1 = fc_take(lb, 7), v2 = fc_take(lb + 7, 13) << 1, v3 = fc_take(lb + 20, 26);
    int c1, c2, c3;
    if (!fc_lcw_part(&v1, 29, 4, T.hdr, 16, &c1)) return 0;
    if (!fc_lcw_part(&v2, 465, 8, T.l2, 256, &c2)) return 0;
    if (!fc_lcw_part(&v3, 41, 5, T.l3, 32, &c3)) return 0;
    if (((v1 >> 4) & 7u) != 2) return 0;                                               // frame type 2 = IDA
    if (n - 46 < 124) return 0;
    // payload: 124-bit blocks of four code words read 4th, 2nd, 3rd, 1st; then a partial block whose halves
    // lose their first bit and swap places (:276-377)
    const uint8_t *pd = data + 46;
    const float *pl = dl ? dl + 46 : nullptr;
    const int pn = n - 46, n_full = pn / 124, rest = pn % 124;
    uint8_t stream[512];
    int len = 0, fixederrs = 0;
    bool failed = false;
    const int take[4] = {3, 1, 2, 0};
    for (int blk = 0; blk < n_full && !failed; blk++) {
        uint8_t cb[124];
        float rel[124];
        const float *bl = pl ? pl + blk * 124 : nullptr;
        fc_halves(pd + blk * 124, bl, 62, cb, cb + 62, rel, rel + 62);
        for (int c = 0; c < 4; c++) {
            if (len + 20 > 512) break;
            const int off = take[c] * 31;
            uint32_t out;
            int touched;
            if (fc_decode31(fc_take(cb + off, 31), bl ? rel + off : nullptr, 3545, 11, T.da, &out, &touched) < 0) { failed = true; break; }
            fixederrs += touched;
            fc_put(stream, &len, out, 20, 11);
        }
    }
    if (!failed && rest >= 4 && len + 2 * (rest / 2 - 1) <= 512) {
        const int ns = rest / 2;
        uint8_t h1[64], h2[64], cb[128];
        float l1[64], l2[64], rel[128];
        for (int i = 0; i < 64; i++) { h1[i] = 0; h2[i] = 0; l1[i] = 0.0f; l2[i] = 0.0f; }   // (odd ns: the reference reads past what it wrote)
        const float *ll = pl ? pl + n_full * 124 : nullptr;
        fc_halves(pd + n_full * 124, ll, ns, h1, h2, l1, l2);
        if (ns > 1 && len + 20 <= 512) {
            int m = 0;
            for (int i = 1; i < ns && m < 128; i++, m++) { cb[m] = h2[i]; if (ll) rel[m] = l2[i]; }
            for (int i = 1; i < ns && m < 128; i++, m++) { cb[m] = h1[i]; if (ll) rel[m] = l1[i]; }
            for (int pos = 0; pos + 31 <= m && len + 20 <= 512; pos += 31) {
                uint32_t out;
                int touched;
                if (fc_decode31(fc_take(cb + pos, 31), ll ? rel + pos : nullptr, 3545, 11, T.da, &out, &touched) < 0) break;
                fixederrs += touched;
                fc_put(stream, &len, out, 20, 11);
            }
        }
    }
    if (len < 196) return 0;                                                           // 20 header + 160 payload + 16 CRC
    const int da_len = (int)fc_take(stream + 11, 5);
    if (fc_take(stream + 17, 3) != 0 || da_len > 20) return 0;
    const int l2d = (int)(v2 >> 8) & 0x3f;
    o->ida_ok = 1;
    o->lcw_ft = (l2d >> 4) & 3; o->lcw_code = l2d & 15; o->lcw3_val = v3 >> 5; o->ec_lcw = c1 + c2 + c3;
    o->cont = stream[3];
    o->da_ctr = (int32_t)fc_take(stream + 5, 3);
    o->da_len = da_len;
    o->fixederrs = fixederrs;
    o->payload_len = da_len > 0 ? da_len : 20;
    for (int i = 0; i < o->payload_len; i++) o->payload[i] = (uint8_t)fc_take(stream + 20 + 8 * i, 8);
    o->bch_len = len;
    for (int i = 0; i < len && i < 256; i++) o->bch_stream[i] = stream[i];
    if (da_len > 0) {                                                                  // CRC-CCITT-FALSE (:381-394, 612-638)
        o->stored_crc = (uint16_t)fc_take(stream + 180, 16);
        if ((20 + 12 + (len - 24) + 7) / 8 <= 64) {
            uint16_t crc = 0xFFFF;
            uint32_t acc = 0;
            int nb = 0;
            // bits 0-19, twelve zero bits, bits 20 .. len-5, a byte at a time, zero-padded at the end
            const int total = 32 + (len - 24);
            for (int k = 0; k < ((total + 7) / 8) * 8; k++) {
                uint32_t bit = 0;
                if (k < 20) bit = stream[k];
                else if (k >= 32 && k < total) bit = stream[k - 12];
                acc = (acc << 1) | bit;
                if (++nb == 8) {
                    crc ^= (uint16_t)(acc << 8);
                    for (int j = 0; j < 8; j++) crc = (uint16_t)((crc & 0x8000) ? (crc << 1) ^ 0x1021 : crc << 1);
                    acc = 0; nb = 0;
                }
            }
            o->computed_crc = crc;
            o->crc_ok = crc == 0;
        }
    }
    return 1;
}

// lat / lon / alt of an IRA frame from its integer position (frame_decode.c:338-347), in double like the
// reference.  alt is exact (products of 12-bit integers and an IEEE square root); lat / lon go through atan2,
// where the device's double-precision atan2 may differ from the host C library's in the last bits (as two C
// libraries may from each other): the GPU tests allow 1e-11 degrees on these two fields and nothing elsewhere.
IR_HD void fc_geo(ir_frame_class_t *o) {
    if (o->frame_type != IR_FRAME_IRA) return;
    const int x = o->pos_xyz[0], y = o->pos_xyz[1], z = o->pos_xyz[2];
    const double xy = sqrt((double)x * x + (double)y * y);
    o->lat = atan2((double)z, xy) * 180.0 / 3.14159265358979323846;
    o->lon = atan2((double)y, (double)x) * 180.0 / 3.14159265358979323846;
    o->alt = (int32_t)(sqrt((double)x * x + (double)y * y + (double)z * z) * 4.0) - 6378 + 23;
}

// both classifiers on one frame, like main.c:320-350
IR_HD void fc_classify(const FcTables &T, const uint8_t *bits, const float *llr, int n_bits, int direction, ir_frame_class_t *o) {
    uint8_t *z = reinterpret_cast<uint8_t *>(o);
    for (unsigned i = 0; i < sizeof(*o); i++) z[i] = 0;
    fc_frame(T, bits, llr, n_bits, o);
    fc_geo(o);
    fc_ida(T, bits, llr, n_bits, direction, o);
}

}  // namespace ir
