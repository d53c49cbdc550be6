"""Two-pass streaming iteration at n: LDG sweeps against the TMA-staged sweeps (option stream_tma)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qn_b200 as qn
eng = qn.default_engine()
iters = int(os.environ.get("QN_STREAM_ITERS", "30"))
for n in [int(a) for a in sys.argv[1:]] or [8192]:
    for tma in (0, 1):
        eng.set_option("stream_tma", tma)
        ms, gbs = C.c_double(), C.c_double()
        qn._lib.check(eng._lib.qn_stream_bench_f64(eng._ctx, n, iters, 2, 0, C.byref(ms), C.byref(gbs)))
        print("n=%d two-pass stream_tma=%d: %.1f us/iteration, %.0f GB/s algorithmic" % (n, tma, ms.value * 1e3, gbs.value))
