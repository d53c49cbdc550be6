"""Streaming path at n (default 8192): µs per iteration against the size of the L2-persisting window
(option l2_persist_mb) and the sweep order, both modes.  python scripts/exp_persist.py [n] [mb ...]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qn_b200 as qn
eng = qn.default_engine()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
mbs = [int(a) for a in sys.argv[2:]] or [0, 32, 64, 96, -1]
print("device: L2 %d MB, persisting max %d MB, access window max %d MB" % (
    eng.get_option("device_l2_mb"), eng.get_option("device_l2_persist_max_mb"), eng.get_option("device_access_window_max_mb")))
iters = int(os.environ.get("QN_STREAM_ITERS", "30"))
for order in (0, 2):
    eng.set_option("sweep_order", order)
    for mb in mbs:
        eng.set_option("l2_persist_mb", mb)
        row = []
        for name, flags in (("two-pass", 0), ("one-pass", qn._lib.FLAG_ONE_PASS)):
            ms, gbs = C.c_double(), C.c_double()
            qn._lib.check(eng._lib.qn_stream_bench_f64(eng._ctx, n, iters, 2, flags, C.byref(ms), C.byref(gbs)))
            row.append("%s %.1f us/it (%.0f GB/s algorithmic)" % (name, ms.value * 1e3, gbs.value))
        print("n=%d sweep_order=%d l2_persist_mb=%d (window %d MB): %s" % (n, order, mb, eng.get_option("last_l2_persist_mb"), "; ".join(row)))
