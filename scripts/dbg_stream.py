import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import qn_b200 as qn
from oracle import binding as ob
from conftest import starts
eng = qn.default_engine()
for n, iters in ((161, 2), (161, 3), (161, 4), (161,40), (162, 40), (256, 5), (300, 6)):
    Theta = starts(n, 1, key=n)
    got = eng.optimize_batched(qn.Rosenbrock(), Theta, max_iter=iters, ntrace=iters)
    ref = ob.optimize_batch(ob.rosenbrock(), Theta, max_iter=iters, ntrace=iters)
    msg = []
    for k in ("x", "g", "trace", "iters", "fevals", "status"):
        same = np.array_equal(got[k], ref[k], equal_nan=True)
        msg.append("%s=%s" % (k, same))
    tr = np.abs(got["trace"][0] - ref["trace"][0]).max(axis=1)
    first_bad = int(np.argmax(tr > 0)) if (tr > 0).any() else -1
    gd = np.abs(got["g"][:, 0] - ref["g"][:, 0])
    print(n, iters, " ".join(msg), "first bad trace iter", first_bad, "g diff idx", np.nonzero(gd > 0)[0][:8], gd.max())
