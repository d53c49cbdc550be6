"""Time the streaming path (n=8192 by default) and report achieved vs measured HBM bandwidth."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qn_b200 as qn
eng = qn.default_engine()
out = {}
for n in [int(a) for a in sys.argv[1:]] or [8192]:
    ms, gbs = C.c_double(), C.c_double()
    modes = (("two-pass (reference order)", 0, 1.5), ("one-pass (QN_FLAG_ONE_PASS)", qn._lib.FLAG_ONE_PASS, 1.0))
    sel = os.environ.get("QN_STREAM_MODE", "both")  # both | twopass | onepass
    for mode, flags, phys in [m for m in modes if sel == "both" or (sel == "onepass") == bool(m[1])]:
        qn._lib.check(eng._lib.qn_stream_bench_f64(eng._ctx, n, int(os.environ.get("QN_STREAM_ITERS", "22")), 2, flags, C.byref(ms), C.byref(gbs)))
        out[(n, mode)] = dict(ms_per_iter=ms.value, algorithmic_gbs=gbs.value, physical_gbs=gbs.value * phys)
        print(n, mode, out[(n, mode)])
hbm = C.c_double(); qn._lib.check(eng._lib.qn_microbench_hbm(eng._ctx, C.byref(hbm)))
print("measured copy bandwidth GB/s", hbm.value)
