// parsed_output.cu -- the reference's --parsed sink (host text, no device code): the "IDA:" line of
// frame_output_print_ida() (frame_output.c:203-357) and the LCW(...) header ida_decode() prepares for it
// (ida_decode.c:398-541), from an ir_frame_t + the ir_frame_class_t the classification kernel produced.
// Like the RAW formatter this is the step right after the device path; it is byte-identical to the
// reference's output (tests/test_parsed_output.py: against the reference's own functions and golden lines).
//
// The LCW header is table-driven here: every (type, code) pair has a pattern for its "C:" part and one for
// the remainder, with placeholders that pull bit fields out of the 21 data bits of the third LCW word:
//   {d:a:n} decimal   {+:a:n} decimal of value+1   {x:a:n} %01x   {X:a:n} %02x   {b:a:n} the bits as text
//   {c:a}   'P' if the bit is 0 else 'S'            {k}     the 4-bit code itself, decimal
// (a = first bit counted from the most significant of the 21, n = width).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "../../include/iridium_b200.h"

namespace {

struct LcwPattern { int ft, code; const char *type, *c, *rest; };   // code -1 = any other code of this type

const LcwPattern kPatterns[] = {
    {0, 0, "maint", "sync[status:{d:1:1},dtoa:{d:3:10},dfoa:{d:13:8}]", "{b:0:1}|{b:2:1}"},
    {0, 1, "maint", "switch[dtoa:{d:3:10},dfoa:{d:13:8}]", "{b:0:3}"},
    {0, 3, "maint", "maint[2][lqi:{d:1:2},power:{d:3:3},f_dtoa:{d:6:7},f_dfoa:{d:13:7}]", "{b:0:1}|{b:20:1}"},
    {0, 6, "maint", "geoloc", "{b:0:21}"},
    {0, 12, "maint", "maint[1][lqi:{d:19:2},power:{d:16:3}]", "{b:0:16}"},
    {0, 15, "maint", "<silent>", "{b:0:21}"},
    {0, -1, "maint", "rsrvd({k})", "{b:0:21}"},
    {1, 1, "acchl", "acchl[msg_type:{x:1:3},bloc_num:{x:4:1},sapi_code:{x:5:3},segm_list:{b:8:8}]", "{b:0:1},{X:16:5}"},
    {1, -1, "acchl", "rsrvd({k})", "{b:0:21}"},
    {2, 3, "hndof",
     "handoff_resp[cand:{c:2},denied:{d:3:1},ref:{d:4:1},slot:{+:6:2},sband_up:{d:8:5},sband_dn:{d:13:5},access:{+:18:3}]",
     "{b:0:2},{b:5:1}"},
    {2, 12, "hndof", "handoff_cand", "{b:0:11},{b:11:10}"},
    {2, 15, "hndof", "<silent>", "{b:0:21}"},
    {2, -1, "hndof", "rsrvd({k})", "{b:0:21}"},
    {3, -1, "rsrvd", "<{k}>", "{b:0:21}"},
};

uint32_t field(uint32_t v21, int at, int n) { return (v21 >> (21 - at - n)) & ((1u << n) - 1u); }

void expand(const char *pat, uint32_t v21, int code, std::string &out) {
    char num[16];
    for (const char *p = pat; *p;) {
        if (*p != '{') { out += *p++; continue; }
        const char kind = p[1];
        int at = 0, n = 1;
        const char *q = p + 2;
        if (*q == ':') { at = (int)strtol(q + 1, (char **)&q, 10); }
        if (*q == ':') { n = (int)strtol(q + 1, (char **)&q, 10); }
        const uint32_t v = field(v21, at, n);
        switch (kind) {
        case 'd': snprintf(num, sizeof num, "%u", v); out += num; break;
        case '+': snprintf(num, sizeof num, "%u", v + 1); out += num; break;
        case 'x': snprintf(num, sizeof num, "%01x", v); out += num; break;
        case 'X': snprintf(num, sizeof num, "%02x", v); out += num; break;
        case 'b': for (int i = n - 1; i >= 0; i--) out += (char)('0' + ((v >> i) & 1u)); break;
        case 'c': out += v ? 'S' : 'P'; break;
        case 'k': snprintf(num, sizeof num, "%d", code); out += num; break;
        }
        p = q + 1;      // past '}'
    }
}

std::string lcw_header(const ir_frame_class_t *c) {
    int ft = c->lcw_ft;
    if (ft < 0 || ft > 3) ft = 3;
    const LcwPattern *use = nullptr;
    for (const LcwPattern &p : kPatterns)
        if (p.ft == ft && (p.code == c->lcw_code || p.code < 0)) { use = &p; break; }
    std::string s = "LCW(2,T:";      // ida_decode() only succeeds on frame type 2 (ida_decode.c:566-567)
    s += use->type;
    s += ",C:";
    expand(use->c, c->lcw3_val & 0x1fffffu, c->lcw_code, s);
    s += ",";
    expand(use->rest, c->lcw3_val & 0x1fffffu, c->lcw_code, s);
    s += ")";
    if (s.size() > 127) s.resize(127);            // the reference's 128-byte scratch
    if (s.size() < 110) s.append(110 - s.size(), ' ');
    s += ' ';
    if (s.size() > 127) s.resize(127);            // lcw_header[128]
    return s;
}

void appendf(std::string &s, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
void appendf(std::string &s, const char *fmt, ...) {
    char tmp[256];
    va_list ap;
    va_start(ap, fmt);
    const int n = vsnprintf(tmp, sizeof tmp, fmt, ap);
    va_end(ap);
    if (n > 0) s.append(tmp, (size_t)(n < (int)sizeof tmp ? n : (int)sizeof tmp - 1));
}

int emit(const std::string &s, char *dst, size_t cap) {
    if (!dst || cap == 0) return -1;
    if (s.size() + 1 > cap) return -1;
    memcpy(dst, s.data(), s.size());
    dst[s.size()] = 0;
    return (int)s.size();
}

}  // namespace

extern "C" int ir_format_lcw(char *dst, size_t cap, const ir_frame_class_t *c) {
    if (!c || !c->ida_ok) return -1;
    return emit(lcw_header(c), dst, cap);
}

extern "C" int ir_format_ida(char *dst, size_t cap, uint64_t t0, const ir_frame_t *f, const ir_frame_class_t *c) {
    return ir_format_ida_hdr(dst, cap, t0, f, c, nullptr);
}

extern "C" int ir_format_ida_hdr(char *dst, size_t cap, uint64_t t0, const ir_frame_t *f, const ir_frame_class_t *c,
                                 const char *lcw_text) {
    if (!f || !c || !c->ida_ok) return -1;
    std::string s;
    const double ts_ms = (double)(f->timestamp - t0) / 1000000.0;
    const int fhz = (int)(f->center_frequency + 0.5);
    const double leveldb = f->level > 0 ? 20.0 * log10((double)f->level) : -99.99;
    int syms = f->n_payload_symbols;
    if (syms < 0) syms = 0;
    appendf(s, "IDA: p-%llu %014.4f %010d %3d%% %06.2f|%07.2f|%05.2f %3d %s ", (unsigned long long)(t0 / 1000000000ULL), ts_ms,
            fhz, f->confidence, leveldb, (double)f->noise, (double)f->magnitude, syms,
            f->direction == IR_DIR_UPLINK ? "UL" : "DL");
    if (lcw_text) s += lcw_text; else s += lcw_header(c);
    const uint8_t *bs = c->bch_stream;
    const int bch_len = c->bch_len < 256 ? c->bch_len : 256;      // what bch_stream holds
    auto bits = [&](int a, int n) { for (int i = a; i < a + n; i++) s += (char)('0' + bs[i]); };
    if (bch_len >= 20) {
        bits(0, 3);
        s += " cont="; bits(3, 1);
        s += ' '; bits(4, 1);
        s += " ctr="; bits(5, 3);
        s += ' '; bits(8, 3);
        appendf(s, " len=%02d", c->da_len);
        s += " 0:"; bits(16, 4);
        // payload as hex: da_len bytes when nothing but zeros follows, else all 20 with '!' at the boundary
        int shown = 20;
        if (c->da_len > 0) {
            bool rest_zero = true;
            for (int i = c->da_len + 1; i < 20; i++) rest_zero = rest_zero && c->payload[i] == 0;
            if (rest_zero) shown = c->da_len < 32 ? c->da_len : 32;      // (da_len <= 20 from the classifier; never read past payload[])
        }
        s += " [";
        for (int i = 0; i < shown; i++) {
            if (i > 0) s += (shown == 20 && c->da_len > 0 && c->da_len < 20 && i == c->da_len) ? '!' : '.';
            appendf(s, "%02x", c->payload[i]);
        }
        s += ']';
        for (int i = shown * 3; i < 60; i++) s += ' ';              // hex + bracket padded to 60 columns
        if (c->da_len > 0) appendf(s, " %04x/%04x %s", c->stored_crc, c->computed_crc, c->crc_ok ? "CRC:OK" : "CRC:no");
        else s += "  ---   ";
        if (bch_len > 196) { s += ' '; bits(196, bch_len - 196); }
        else s += " 0000";
        if (c->da_len > 0 && bch_len >= 180) {
            s += " SBD: ";
            for (int i = 0; i < 20; i++) {
                int byte = 0;
                for (int b = 0; b < 8; b++) byte = (byte << 1) | bs[20 + 8 * i + b];
                s += (byte >= 32 && byte < 127) ? (char)byte : '.';
            }
        }
    }
    s += '\n';
    return emit(s, dst, cap);
}
