"""Streaming path: sweeps with / without eviction hints (option l2_resident) for sizes around the L2 capacity."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qn_b200 as qn
eng = qn.default_engine()
iters = int(os.environ.get("QN_STREAM_ITERS", "40"))
for n in [int(a) for a in sys.argv[1:]] or [1024, 2048, 2560, 2816, 3072, 4096]:
    for res in (0, 1):
        eng.set_option("l2_resident", res)
        row = []
        for name, flags in (("two-pass", 0), ("one-pass", qn._lib.FLAG_ONE_PASS)):
            ms, gbs = C.c_double(), C.c_double()
            qn._lib.check(eng._lib.qn_stream_bench_f64(eng._ctx, n, iters, 2, flags, C.byref(ms), C.byref(gbs)))
            row.append("%s %.1f us/it" % (name, ms.value * 1e3))
        print("n=%d (B^-1 %.0f MB) l2_resident=%d: %s" % (n, n * n * 8 / 1e6, res, "; ".join(row)), flush=True)
