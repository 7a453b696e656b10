"""Developer check: variants 0 and 3 at N = Nv in (24, 32), one step; prints where they differ."""
import sys
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as graft
pkg = graft.load_package()
from lpsolver_b200 import solver
for N in (24, 32):
    cfg = dict(Nx=2, Nv=N, N=N, Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
    U = solver.set_init_ld(cfg["Nx"], N, cfg["Lv"], cfg["Lx"], 0.5, 2 * np.pi / 4., True)
    out, mom = {}, {}
    for variant in (0, 3):
        g = pkg.LPGpu(computeq_variant=variant, **cfg)
        g.upload_U(U)
        m0 = g.moments()
        g.collide_step()
        oc = g.download_U()
        mc = g.moments()
        g.upload_U(U)
        g.step(1)
        out[variant], mom[variant] = g.download_U(), g.moments()
        g.close()
        print(N, variant, "m0", m0, "\n   after collide", mc, "\n   after step", mom[variant], " max|dU_coll| %.3e max|dU_step| %.3e" % (np.abs(oc - U).max(), np.abs(out[variant] - U).max()))
        out[(variant, "c")] = oc
    d = np.abs(out[0] - out[3]); dc = np.abs(out[(0, "c")] - out[(3, "c")])
    print(N, "step diff max %.3e at %d (coef %d); collide diff max %.3e at %d" % (d.max(), d.argmax(), d.argmax() % 6, dc.max(), dc.argmax()))
