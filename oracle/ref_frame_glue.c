/*
 * ref_frame_glue.c -- TEST INFRASTRUCTURE ONLY.  Calls the reference's own frame_decode()
 * (frame_decode.c, compiled unmodified from /root/reference into oracle/_ref/libref_frame.so) and
 * flattens decoded_frame_t (frame_decode.h:26-60) into the struct tests compare with
 * oracle/ir_frame_oracle.c's.
 */
#include <string.h>

#include "frame_decode.h"

typedef struct {
    int32_t ret;
    int32_t type;
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

int ref_frame_decode(const uint8_t *bits, const float *llr, int n_bits, orc_frame_t *o) {
    static int ready;
    if (!ready) { frame_decode_init(); ready = 1; }
    demod_frame_t f;
    decoded_frame_t d;
    memset(&f, 0, sizeof(f));
    memset(o, 0, sizeof(*o));
    f.bits = (uint8_t *)bits;
    f.llr = (float *)llr;
    f.n_bits = n_bits;
    o->ret = frame_decode(&f, &d);
    o->type = (int32_t)d.type;
    if (d.type == FRAME_IRA) {
        o->sat_id = d.ira.sat_id; o->beam_id = d.ira.beam_id;
        o->lat = d.ira.lat; o->lon = d.ira.lon; o->alt = d.ira.alt;
        for (int i = 0; i < 3; i++) o->pos_xyz[i] = d.ira.pos_xyz[i];
        o->n_pages = d.ira.n_pages;
        for (int i = 0; i < 12; i++) { o->tmsi[i] = d.ira.pages[i].tmsi; o->msc_id[i] = d.ira.pages[i].msc_id; }
    } else if (d.type == FRAME_IBC) {
        o->sat_id = d.ibc.sat_id; o->beam_id = d.ibc.beam_id;
        o->timeslot = d.ibc.timeslot; o->sv_blocking = d.ibc.sv_blocking;
        o->bc_type = d.ibc.bc_type; o->iri_time = d.ibc.iri_time;
    }
    return o->ret;
}

/* ---- ida_decode() (ida_decode.c compiled unmodified), flattened the same way ---- */
#include "ida_decode.h"

typedef struct {
    int32_t ret;
    int32_t ft, lcw_ok, lcw_ft, lcw_code, ec_lcw;
    uint32_t lcw3_val;
    int32_t da_ctr, da_len, cont, payload_len, crc_ok, fixederrs, bch_len;
    uint16_t stored_crc, computed_crc;
    uint8_t payload[32];
    uint8_t bch_stream[256];
} orc_ida_t;

int ref_ida_decode(const uint8_t *bits, const float *llr, int n_bits, int direction, orc_ida_t *o) {
    static int ready;
    if (!ready) { frame_decode_init(); ida_decode_init(); ready = 1; }
    demod_frame_t f;
    static ida_burst_t b;
    memset(&f, 0, sizeof(f));
    memset(o, 0, sizeof(*o));
    f.bits = (uint8_t *)bits;
    f.llr = (float *)llr;
    f.n_bits = n_bits;
    f.direction = (ir_direction_t)direction;
    o->ret = ida_decode(&f, &b);
    if (!o->ret) return 0;
    o->ft = b.lcw.ft; o->lcw_ok = b.lcw.lcw_ok; o->lcw_ft = b.lcw.lcw_ft; o->lcw_code = b.lcw.lcw_code;
    o->ec_lcw = b.lcw.ec_lcw; o->lcw3_val = b.lcw.lcw3_val;
    o->da_ctr = b.da_ctr; o->da_len = b.da_len; o->cont = b.cont; o->payload_len = b.payload_len;
    o->crc_ok = b.crc_ok; o->fixederrs = b.fixederrs; o->bch_len = b.bch_len;
    o->stored_crc = b.stored_crc; o->computed_crc = b.computed_crc;
    memcpy(o->payload, b.payload, sizeof(o->payload));
    memcpy(o->bch_stream, b.bch_stream, sizeof(o->bch_stream));
    return 1;
}

/* ---- the reference's --parsed sink: ida_decode() + frame_output_print_ida() (frame_output.c compiled
 * unmodified), with stdout captured into the caller's buffer.  frame_output.c keeps its time origin in a
 * static that the first printed line sets (frame_output.c:144-158): ref_print_prime() prints one throw-away
 * line so that the origin is a known whole second for everything printed afterwards in this process. ---- */
#include <stdio.h>
#include <unistd.h>

#include "frame_output.h"

int diagnostic_mode, parsed_mode, acars_enabled;      /* main.c's globals that frame_output.c refers to */

static int capture_begin(FILE **tmp) {
    fflush(stdout);
    const int saved = dup(1);
    *tmp = tmpfile();
    if (saved < 0 || !*tmp) return -1;
    dup2(fileno(*tmp), 1);
    return saved;
}
static int capture_end(int saved, FILE *tmp, char *out, int cap) {
    fflush(stdout);
    dup2(saved, 1);
    close(saved);
    rewind(tmp);
    const int n = (int)fread(out, 1, cap > 0 ? cap - 1 : 0, tmp);
    fclose(tmp);
    if (cap > 0) out[n] = 0;
    return n;
}

void ref_print_prime(uint64_t timestamp_ns) {
    static ida_burst_t b;
    char sink[512];
    FILE *tmp;
    memset(&b, 0, sizeof(b));
    b.timestamp = timestamp_ns;
    const int saved = capture_begin(&tmp);
    if (saved < 0) return;
    frame_output_print_ida(&b);
    capture_end(saved, tmp, sink, sizeof(sink));
}

/* returns the length of the IDA line (0 if ida_decode() refuses the frame); lcw_header gets ida_burst_t.lcw_header */
int ref_print_ida(const uint8_t *bits, const float *llr, int n_bits, int direction, uint64_t timestamp,
                  double center_frequency, float magnitude, float noise, float level, int confidence,
                  int n_payload_symbols, char *out, int cap, char *lcw_header) {
    static int ready;
    if (!ready) { frame_decode_init(); ida_decode_init(); ready = 1; }
    demod_frame_t f;
    static ida_burst_t b;
    memset(&f, 0, sizeof(f));
    f.bits = (uint8_t *)bits; f.llr = (float *)llr; f.n_bits = n_bits;
    f.direction = (ir_direction_t)direction;
    f.timestamp = timestamp; f.center_frequency = center_frequency;
    f.magnitude = magnitude; f.noise = noise; f.level = level; f.confidence = confidence;
    f.n_payload_symbols = n_payload_symbols; f.n_symbols = n_payload_symbols + 12;
    if (cap > 0) out[0] = 0;
    if (!ida_decode(&f, &b)) return 0;
    if (lcw_header) memcpy(lcw_header, b.lcw_header, sizeof(b.lcw_header));
    FILE *tmp;
    const int saved = capture_begin(&tmp);
    if (saved < 0) return -1;
    frame_output_print_ida(&b);
    return capture_end(saved, tmp, out, cap);
}

/* ---- the reference's structs as they come out of frame_decode() / ida_decode(), byte for byte ---- */
static void fill_frame(demod_frame_t *f, const uint8_t *bits, const float *llr, int n_bits, int direction,
                       uint64_t timestamp, double center_frequency, float magnitude, float noise, float level,
                       int confidence, int n_payload_symbols) {
    memset(f, 0, sizeof(*f));
    f->bits = (uint8_t *)bits; f->llr = (float *)llr; f->n_bits = n_bits;
    f->direction = (ir_direction_t)direction;
    f->timestamp = timestamp; f->center_frequency = center_frequency;
    f->magnitude = magnitude; f->noise = noise; f->level = level; f->confidence = confidence;
    f->n_payload_symbols = n_payload_symbols; f->n_symbols = n_payload_symbols + 12;
}
int ref_sizeof_decoded_frame(void) { return (int)sizeof(decoded_frame_t); }
int ref_sizeof_ida_burst(void) { return (int)sizeof(ida_burst_t); }
int ref_sizeof_ida_context(void) { return (int)sizeof(ida_context_t); }
int ref_frame_decode_raw(const uint8_t *bits, const float *llr, int n_bits, int direction, uint64_t timestamp,
                         double center_frequency, float magnitude, float noise, float level, int confidence,
                         int n_payload_symbols, void *out) {
    static int ready;
    if (!ready) { frame_decode_init(); ida_decode_init(); ready = 1; }
    demod_frame_t f;
    fill_frame(&f, bits, llr, n_bits, direction, timestamp, center_frequency, magnitude, noise, level, confidence, n_payload_symbols);
    return frame_decode(&f, (decoded_frame_t *)out);
}
int ref_ida_decode_raw(const uint8_t *bits, const float *llr, int n_bits, int direction, uint64_t timestamp,
                       double center_frequency, float magnitude, float noise, float level, int confidence,
                       int n_payload_symbols, void *out) {
    static int ready;
    if (!ready) { frame_decode_init(); ida_decode_init(); ready = 1; }
    demod_frame_t f;
    fill_frame(&f, bits, llr, n_bits, direction, timestamp, center_frequency, magnitude, noise, level, confidence, n_payload_symbols);
    return ida_decode(&f, (ida_burst_t *)out);
}
