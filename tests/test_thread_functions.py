"""burst_detector_thread / burst_downmix_thread (burst_detect.c:941-960, burst_downmix.c:801-828) bind to
the HOST PROGRAM's queues through weak symbols.  No GPU: (1) stand-alone (ctypes) the symbols are null and
the functions return at once; (2) linked into a small C program that defines the queues the way main.c
does, they take from the queue, see it closed, destroy their (null) context and return."""
import ctypes as C
import importlib
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r"""
#include <stdio.h>
typedef struct { int closed; } Blocking_Queue;            /* stands in for blocking_queue.h:140-165 */
Blocking_Queue samples_queue, burst_queue, frame_queue;  /* main.c:176-178 */
unsigned long stat_n_detected, stat_n_dropped;           /* main.c:181,185 */
static int takes;
int blocking_queue_take(Blocking_Queue *bq, void *e) { (void)bq; (void)e; takes++; return 1; }   /* closed */
int blocking_queue_put(Blocking_Queue *bq, void *e) { (void)bq; (void)e; return 0; }
int blocking_queue_add(Blocking_Queue *bq, void *e) { (void)bq; (void)e; return 0; }
void *burst_detector_thread(void *arg);
void *burst_downmix_thread(void *arg);
int main(void) {
    void *a = burst_detector_thread(NULL);
    void *b = burst_downmix_thread(NULL);
    printf("takes=%d\n", takes);
    return (takes == 2 && !a && !b) ? 0 : 1;
}
"""


def _lib():
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    if not os.path.exists(pl.LIB_PATH):
        pl.build_library()
    return pl, pl.load_library()


def test_thread_functions_return_at_once_without_the_host_queues():
    _, L = _lib()
    for name in ("burst_detector_thread", "burst_downmix_thread"):
        f = getattr(L, name)
        f.restype = C.c_void_p
        f.argtypes = [C.c_void_p]
        assert f(None) is None


def test_thread_functions_bind_to_the_host_programs_queues():
    pl, _ = _lib()
    libdir = os.path.dirname(pl.LIB_PATH)
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "h.c"), os.path.join(d, "h")
        open(src, "w").write(HARNESS)
        subprocess.run(["gcc", "-O1", "-rdynamic", "-o", exe, src, "-L" + libdir, "-liridium_b200",
                        "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, (r.stdout, r.stderr)
        assert "takes=2" in r.stdout
