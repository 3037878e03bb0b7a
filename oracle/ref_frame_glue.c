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
