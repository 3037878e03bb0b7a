// host_tables.cpp -- constants of the path, computed on the HOST with the same float
// expressions and the same libm the reference uses, then uploaded.  Nothing here is
// recomputed with device intrinsics (SURVEY.md hard part 2): taps, windows, sync
// templates, twiddles, detector parameters.
//
// Build note: compiled with -ffp-contract=off so the host compiler fuses nothing.
#include <math.h>
#include <string.h>

#include <complex>

#include "ir_device.cuh"
#include "ir_internal.h"

namespace ir {

static const float kPi = (float)M_PI;

// ---- detector parameters: burst_detector_create (burst_detect.c:180-226, :292-296)
void derive_det_config(DetConfig &c, int sample_rate, int fft_size, int burst_width_hz,
                       float threshold_db) {
    memset(&c, 0, sizeof(c));
    c.sample_rate = sample_rate;
    if (fft_size > 0) {
        c.N = fft_size;
    } else {
        int e = (int)round(log2(sample_rate / 1000.0));
        c.N = 1 << e;
    }
    c.L = 0;
    while ((1 << c.L) < c.N) c.L++;
    c.pre_len = 2 * c.N;
    c.post_len = (int)(sample_rate * 16e-3);
    if (burst_width_hz <= 0) burst_width_hz = 40000;
    c.burst_width = burst_width_hz / (sample_rate / c.N);
    c.half_bw = c.burst_width / 2;
    c.max_bursts = (int)((sample_rate / (float)burst_width_hz) * 0.8f);
    c.max_burst_len = (int)(sample_rate * 0.09);
    c.hist_size = 512;
    if (!(threshold_db > 0)) threshold_db = 16.0f;
    const float enbw = 1.72f;
    c.thr = powf(10.0f, threshold_db / 10.0f) / c.hist_size / enbw;
    uint64_t ring = (uint64_t)c.max_burst_len + c.pre_len + c.post_len + (uint64_t)c.N * 4;
    if (ring < (uint64_t)(2 * (int64_t)sample_rate)) ring = 2 * (uint64_t)sample_rate;
    c.ringbuf_size = ring;
}

// ---- window_func.c:19-24
static void blackman_f(float *w, int n) {
    for (int i = 0; i < n; i++)
        w[i] = 0.42f - 0.5f * cosf(2.0f * kPi * i / (n - 1)) + 0.08f * cosf(4.0f * kPi * i / (n - 1));
}

// ---- fir_filter.c:143-182 (windowed sinc, Blackman-Harris, unity DC gain)
static std::vector<float> lowpass(float gain, float fs, float cutoff, float transition) {
    int nt = (int)(4.0f / (transition / fs));
    nt |= 1;
    std::vector<float> h(nt);
    const int mid = nt / 2;
    const float wc = 2.0f * kPi * cutoff / fs;
    float acc = 0;
    for (int i = 0; i < nt; i++) {
        float k = i - mid;
        float sinc = fabsf(k) < 1e-10f ? wc / kPi : sinf(wc * k) / (kPi * k);
        float win = 0.35875f - 0.48829f * cosf(2.0f * kPi * i / (nt - 1))
                             + 0.14128f * cosf(4.0f * kPi * i / (nt - 1))
                             - 0.01168f * cosf(6.0f * kPi * i / (nt - 1));
        h[i] = sinc * win;
        acc += h[i];
    }
    if (fabsf(acc) > 0) {
        float sc = gain / acc;
        for (auto &v : h) v *= sc;
    }
    return h;
}

// ---- fir_filter.c:74-111 (unit-energy RRC)
static std::vector<float> root_raised_cosine(float gain, float fs, float sym, float alpha, int nt) {
    nt |= 1;
    std::vector<float> h(nt);
    const float sps = fs / sym;
    const int mid = nt / 2;
    float energy = 0;
    for (int i = 0; i < nt; i++) {
        float t = (i - mid) / sps;
        if (fabsf(t) < 1e-10f) {
            h[i] = (1.0f - alpha + 4.0f * alpha / kPi);
        } else if (fabsf(fabsf(t) - 1.0f / (4.0f * alpha)) < 1e-6f) {
            h[i] = alpha / sqrtf(2.0f) *
                   ((1.0f + 2.0f / kPi) * sinf(kPi / (4.0f * alpha)) +
                    (1.0f - 2.0f / kPi) * cosf(kPi / (4.0f * alpha)));
        } else {
            float num = sinf(kPi * t * (1.0f - alpha)) + 4.0f * alpha * t * cosf(kPi * t * (1.0f + alpha));
            float den = kPi * t * (1.0f - (4.0f * alpha * t) * (4.0f * alpha * t));
            h[i] = num / den;
        }
        energy += h[i] * h[i];
    }
    float sc = gain / sqrtf(energy);
    for (auto &v : h) v *= sc;
    return h;
}

static float sinc_pi(float x) {      // fir_filter.c:67-70
    if (fabsf(x) < 1e-10f) return 1.0f;
    return sinf(kPi * x) / (kPi * x);
}

// ---- fir_filter.c:115-139
static std::vector<float> raised_cosine(float fs, float sym, float alpha, int nt) {
    nt |= 1;
    std::vector<float> h(nt);
    const float sps = fs / sym;
    const int mid = nt / 2;
    for (int i = 0; i < nt; i++) {
        float t = (i - mid) / sps;
        if (fabsf(t) < 1e-10f) {
            h[i] = 1.0f;
        } else if (alpha > 0 && fabsf(fabsf(t) - 1.0f / (2.0f * alpha)) < 1e-6f) {
            h[i] = kPi / (4.0f) * sinc_pi(1.0f / (2.0f * alpha));
        } else {
            float c = cosf(kPi * alpha * t);
            float den = 1.0f - (2.0f * alpha * t) * (2.0f * alpha * t);
            h[i] = sinc_pi(t) * c / den;
        }
    }
    return h;
}

// ---- twiddles: W[k] = ((float)cos(2 pi k/N), (float)-sin(2 pi k/N)), exact at k=0 and N/4
static void twiddle_table(int n, std::vector<float2> &w) {
    int h = n / 2 > 0 ? n / 2 : 1;
    w.resize(h);
    for (int k = 0; k < h; k++) {
        double a = 2.0 * M_PI * (double)k / (double)n;
        w[k] = make_float2((float)cos(a), (float)(-sin(a)));
    }
    w[0] = make_float2(1.0f, 0.0f);
    if (n >= 4) w[n / 4] = make_float2(0.0f, -1.0f);
}

template <int L>
static std::vector<float2> tw_image_t() {
    constexpr int N = 1 << L;
    std::vector<float2> w;
    twiddle_table(N, w);
    std::vector<float2> img(fft_tw_elems<L>(), make_float2(0, 0));
    for (int k = 0; k < N / 2; k++) img[tw_pad(k)] = w[k];
    float2 *c = img.data() + fft_twfull_elems<L>();
    for (int s = FftPlan<L>::Q0; s < L; s++)
        for (int j = 0; j < (N >> (s + 1)); j++) c[twc_off<L>(s) + j] = w[(size_t)j << s];
    return img;
}

std::vector<float2> build_twiddle_image(int L) {
    switch (L) {
    case 9: return tw_image_t<9>();
    case 10: return tw_image_t<10>();
    case 11: return tw_image_t<11>();
    case 12: return tw_image_t<12>();
    case 13: return tw_image_t<13>();
    case 14: return tw_image_t<14>();
    default: return {};
    }
}

// Host copy of the transform (used once, for the two sync templates).
void host_fft(std::vector<float2> &x, bool inverse) {
    const int n = (int)x.size();
    std::vector<float2> w;
    twiddle_table(n, w);
    for (int half = n / 2, step = 1; half >= 1; half >>= 1, step <<= 1)
        for (int blk = 0; blk < n; blk += 2 * half)
            for (int j = 0; j < half; j++) {
                float2 a = x[blk + j], b = x[blk + j + half], t = w[(size_t)j * step];
                if (inverse) t.y = -t.y;
                float dr = a.x - b.x, di = a.y - b.y;
                x[blk + j] = make_float2(a.x + b.x, a.y + b.y);
                float p = di * t.y, q = di * t.x;
                x[blk + j + half] = make_float2(fmaf(dr, t.x, -p), fmaf(dr, t.y, q));
            }
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(x[i], x[j]);
    }
}

// ---- sync templates: generate_sync_word (burst_downmix.c:138-219)
static std::vector<float2> sync_template(const std::vector<float> &rc, const int *uw, bool uplink) {
    const int pre = 16, uwl = 12, sps = 10;
    const int nsym = pre + uwl;
    const int plen = nsym * sps - (sps - 1);
    const int half = ((int)rc.size() - 1) / 2;
    std::vector<float2> buf(plen + rc.size() - 1, make_float2(0, 0));
    for (int i = 0; i < nsym; i++) {
        bool s0;                       // s0 = 1+1j, s1 = -1-1j
        if (i < pre) s0 = uplink ? (i % 2 != 0) : true;
        else s0 = uw[i - pre] == 0;
        buf[half + i * sps] = s0 ? make_float2(1.0f, 1.0f) : make_float2(-1.0f, -1.0f);
    }
    // pulse shaping: one FMA chain per output (simd_avx2.c:28-44); plen % 4 == 3 outputs fall
    // into the remainder loop, whose as-compiled arithmetic is reproduced by tail_chain below.
    auto tail_chain = [&](const float *x) {
        const int nt = (int)rc.size();
        float a = 0;
        int k = 0, k8 = nt & ~7;
        for (; k < k8; k++) a = a + rc[k] * x[2 * k];
        if (nt - k >= 4) for (int e = k + 4; k < e; k++) a = a + rc[k] * x[2 * k];
        for (; k < nt; k++) a = fmaf(rc[k], x[2 * k], a);
        return a;
    };
    std::vector<float2> shaped(plen);
    const int body = plen & ~3;
    for (int i = 0; i < plen; i++) {
        if (i < body) {
            float ar = 0, ai = 0;
            for (size_t k = 0; k < rc.size(); k++) {
                ar = fmaf(rc[k], buf[i + k].x, ar);
                ai = fmaf(rc[k], buf[i + k].y, ai);
            }
            shaped[i] = make_float2(ar, ai);
        } else {
            const float *x = &buf[i].x;
            shaped[i] = make_float2(tail_chain(x), tail_chain(x + 1));
        }
    }
    std::vector<float2> tpl(IR_CORR_N, make_float2(0, 0));
    for (int i = 0; i < plen; i++)     // reversed and conjugated
        tpl[i] = make_float2(shaped[plen - 1 - i].x, -shaped[plen - 1 - i].y);
    host_fft(tpl, false);
    return tpl;
}

void build_host_tables(HostTables &t, int det_fft_size) {
    t.det_window.resize(det_fft_size);
    blackman_f(t.det_window.data(), det_fft_size);
    for (auto &v : t.det_window) v /= 0.42f;
    const float out_rate = (float)IR_OUT_RATE;
    t.h_input = lowpass(1.0f, 10000000.0f, IR_OUT_RATE * 0.4f, IR_OUT_RATE * 0.2f);
    t.h_noise = lowpass(1.0f, out_rate, 40000.0f / 2.0f, 40000.0f);
    t.h_box.assign(20, 1.0f / 20);
    t.h_rrc = root_raised_cosine(1.0f, out_rate, 25000.0f, 0.4f, 51);
    t.h_rc = raised_cosine(out_rate, 25000.0f, 0.4f, 51);
    t.cfo_window.resize(IR_CFO_N);
    blackman_f(t.cfo_window.data(), IR_CFO_N);
    static const int uw_dl[12] = {0, 2, 2, 2, 2, 0, 0, 0, 2, 0, 0, 2};   // iridium.h:30
    static const int uw_ul[12] = {2, 2, 0, 0, 0, 2, 0, 0, 2, 0, 2, 2};   // iridium.h:31
    t.sync_dl_fft = sync_template(t.h_rc, uw_dl, false);
    t.sync_ul_fft = sync_template(t.h_rc, uw_ul, true);
}

}  // namespace ir
