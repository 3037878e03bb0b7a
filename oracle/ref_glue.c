/*
 * Glue for oracle/_ref/libref_path.so -- TEST INFRASTRUCTURE ONLY.
 *
 * libref_path.so = the reference's own hot-path translation units
 * (burst_detect.c burst_downmix.c qpsk_demod.c fir_filter.c window_func.c
 * simd_generic.c simd_avx2.c) compiled UNMODIFIED from /root/reference, plus
 * oracle/shim/fftw_shim.c, plus this file.  The reference objects expect a few
 * globals that live in main.c (main.c:94-186); this file supplies them so the
 * stage functions can be called one at a time from the parity tests:
 *
 *   burst_detector_create / _feed / _feed_cf32   (burst_detect.h:67-84)
 *   burst_downmix_create / _process              (burst_downmix.h:64-72)
 *   qpsk_demod                                   (qpsk_demod.h:42)
 *
 * Compiled with -I/root/reference so that blocking_queue.h etc. are read
 * where they lie; no reference source is copied into this repository.
 */
#define _GNU_SOURCE
#include <signal.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define C_FEK_BLOCKING_QUEUE_IMPLEMENTATION
#define C_FEK_FAIR_LOCK_IMPLEMENTATION
#include "blocking_queue.h"
#include "burst_detect.h"
#include "burst_downmix.h"
#include "qpsk_demod.h"
#include "simd_kernels.h"

/* ---- globals the reference TUs declare extern ---- */
Blocking_Queue samples_queue;
Blocking_Queue burst_queue;
Blocking_Queue frame_queue;
volatile sig_atomic_t running = 1;
int verbose = 0;
atomic_ulong stat_n_detected = 0;
atomic_ulong stat_n_dropped = 0;
char *save_bursts_dir = NULL;
int use_gardner = 1;
pthread_mutex_t fftw_planner_mutex = PTHREAD_MUTEX_INITIALIZER;

/* ---- small control surface for the tests ---- */
void ref_init(int no_simd, int gardner, int verbose_flag) {
    simd_init(no_simd);
    use_gardner = gardner;
    verbose = verbose_flag;
}

void ref_set_gardner(int on) { use_gardner = on; }

/* Collect bursts from the detector callback into a growable array. */
typedef struct {
    burst_data_t **items;
    size_t n, cap;
} ref_burst_list_t;

static void collect_cb(burst_data_t *b, void *user) {
    ref_burst_list_t *l = (ref_burst_list_t *)user;
    if (l->n == l->cap) {
        l->cap = l->cap ? l->cap * 2 : 256;
        l->items = realloc(l->items, l->cap * sizeof(*l->items));
    }
    l->items[l->n++] = b;
}

ref_burst_list_t *ref_burst_list_new(void) { return calloc(1, sizeof(ref_burst_list_t)); }
size_t ref_burst_list_len(ref_burst_list_t *l) { return l->n; }
burst_data_t *ref_burst_list_get(ref_burst_list_t *l, size_t i) { return l->items[i]; }
void ref_burst_list_free(ref_burst_list_t *l) {
    for (size_t i = 0; i < l->n; i++) {
        free(l->items[i]->samples);
        free(l->items[i]);
    }
    free(l->items);
    free(l);
}

/* Feed a whole recording the way spewer_thread + burst_detector_thread do
 * (main.c:223-284, burst_detect.c:941-956): `block` samples per feed call. */
void ref_detect_cf32(burst_detector_t *d, const float *iq, size_t n, size_t block,
                     ref_burst_list_t *out) {
    for (size_t off = 0; off < n; off += block) {
        size_t m = n - off < block ? n - off : block;
        burst_detector_feed_cf32(d, iq + 2 * off, m, collect_cb, out);
    }
}

void ref_detect_ci8(burst_detector_t *d, const int8_t *iq, size_t n, size_t block,
                    ref_burst_list_t *out) {
    for (size_t off = 0; off < n; off += block) {
        size_t m = n - off < block ? n - off : block;
        burst_detector_feed(d, iq + 2 * off, m, collect_cb, out);
    }
}

void ref_free(void *p) { free(p); }
