"""Ad-hoc stage-by-stage GPU-vs-oracle comparison (run under gpurun; not a pytest module)."""
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, ".")
from mrg_slam_b200 import lib as B  # noqa: E402
from mrg_slam_b200 import synth  # noqa: E402
from tests import oraclelib as O  # noqa: E402


def section(name):
    print(f"\n=== {name} ===", flush=True)


def pose_err(Ta, Tb):
    d = np.linalg.inv(Ta) @ Tb
    return np.linalg.norm(d[:3, 3]), np.arccos(np.clip((np.trace(d[:3, :3]) - 1) / 2, -1, 1))


def run(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()


def main():
    sensor = synth.VLP16 if len(sys.argv) < 2 else int(sys.argv[1])
    raw_a, raw_b = synth.scan(sensor, 3), synth.scan(sensor, 4)
    reg = B.Registration(B.default_config(B.FAST_VGICP))
    print("raw", raw_a.shape, raw_b.shape)

    def filters():
        section("distance filter")
        o = O.distance_filter(raw_a, 0.1, 35.0)
        g = reg.distance_filter(raw_a, 0.1, 35.0)
        print("oracle", o.shape, "gpu", g.shape, "equal", o.shape == g.shape and np.array_equal(o, g))
        section("voxelgrid 0.1")
        t = time.time(); ov, _ = O.voxelgrid(o, 0.1, 1); t1 = time.time() - t
        t = time.time(); gv, ovf = reg.voxelgrid(o, 0.1, 1); t2 = time.time() - t
        print("oracle", ov.shape, "gpu", gv.shape, "overflow", ovf, "equal", ov.shape == gv.shape and np.array_equal(ov, gv), t1, t2)
        if ov.shape == gv.shape and not np.array_equal(ov, gv):
            bad = np.where((ov != gv).any(1))[0]
            print("mismatch rows", len(bad), bad[:5], ov[bad[:3]], gv[bad[:3]])
        section("radius 0.5/2")
        keep = O.radius_outlier(ov, 0.5, 2)
        gr = reg.radius_outlier(ov, 0.5, 2)
        print("oracle", keep.sum(), "gpu", gr.shape, "equal", np.array_equal(ov[keep], gr))
        section("statistical 30/1.2")
        keep, dist, thr = O.statistical_outlier(ov, 30, 1.2)
        gs = reg.statistical_outlier(ov, 30, 1.2)
        print("oracle", keep.sum(), "gpu", gs.shape, "equal", np.array_equal(ov[keep], gs))
        section("prefilter chain")
        gp = reg.prefilter(raw_a)
        print("chain", gp.shape, "equal", np.array_equal(gp, gr))

    run(filters)

    def pre(c):
        c = O.distance_filter(c, 0.1, 35.0)
        c, _ = O.voxelgrid(c, 0.1, 1)
        return c[O.radius_outlier(c, 0.5, 2)]

    A, Bc = pre(raw_a), pre(raw_b)
    gt = np.linalg.inv(synth.pose(3)) @ synth.pose(4)
    print("clouds", A.shape, Bc.shape, "gt t", gt[:3, 3])

    def knn():
        section("knn")
        cl = B.Cloud(reg, A)
        rng = np.random.default_rng(0)
        q = np.concatenate([A[rng.integers(0, len(A), 500)], (rng.normal(size=(200, 4)) * 20).astype(np.float32)])
        oi, od = O.knn(A, q, 20)
        gi, gd = reg.debug_knn(cl, q, 20)
        print("idx equal", np.array_equal(oi, gi), "d2 equal", np.array_equal(od, gd), "mismatch rows", (oi != gi).any(1).sum())

    run(knn)

    for method, name in ((B.FAST_VGICP, "FAST_VGICP"), (B.FAST_GICP, "FAST_GICP")):
        def lsq(method=method, name=name):
            section(name)
            g = B.Registration(B.default_config(method))
            o = O.Registration(O.default_params(method))
            g.setInputTarget(A); g.setInputSource(Bc)
            o.setInputTarget(A); o.setInputSource(Bc)
            ocov, oknn = O.knn_covariances(Bc, 20, want_idx=True)
            gcov, gknn = g.debug_covariances(0, want_knn=True)
            print("knn sets equal", np.array_equal(np.sort(oknn, 1), np.sort(gknn, 1)), "ordered equal", np.array_equal(oknn, gknn),
                  "cov max abs diff", np.abs(ocov - gcov).max())
            if method == B.FAST_VGICP:
                tcov = O.knn_covariances(A, 20)
                oc, on, om, ov = O.vgicp_voxelmap(A, tcov, 1.0)
                gc, gn, gm, gv = g.debug_voxelmap()
                print("voxels", len(oc), len(gc), "coords equal", np.array_equal(oc, gc), "npts equal", np.array_equal(on, gn))
                if len(oc) == len(gc):
                    print("mean maxdiff", np.abs(om - gm).max(), "cov maxdiff", np.abs(ov - gv).max())
            for T in (np.eye(4), gt):
                oe, oH, ob, ocorr, oval = o.linearize(T)
                ge, gH, gb, gcorr, gval = g.debug_linearize(T)
                print("lin err", oe, ge, "rel", abs(oe - ge) / abs(oe), "H rel", np.abs(oH - gH).max() / np.abs(oH).max(), "b rel",
                      np.abs(ob - gb).max() / np.abs(ob).max(), "valid equal", np.array_equal(oval, gval), "corr equal",
                      np.array_equal(ocorr[oval], gcorr[gval]) if np.array_equal(oval, gval) else None)
            guess = gt.copy(); guess[0, 3] -= 0.3
            t = time.time(); ro = o.align(guess); t_o = time.time() - t
            t = time.time(); rg = g.align(guess); t_g = time.time() - t
            t = time.time(); rg = g.align(guess); t_g2 = time.time() - t
            To, Tg = o.getFinalTransformation(), g.getFinalTransformation()
            print("oracle conv", ro.converged, ro.iterations, ro.lm_evals, "gpu conv", rg.converged, rg.iterations, rg.evals)
            print("T diff (m, rad)", pose_err(To, Tg), "vs gt oracle", pose_err(gt, To), "gpu", pose_err(gt, Tg))
            print("time oracle %.3f gpu first %.3f second %.4f" % (t_o, t_g, t_g2), g.last_timings())
            fo, fg = o.getFitnessScore(), g.getFitnessScore()
            print("fitness", fo, fg, "rel", abs(fo - fg) / fo)
            fo, fg = o.getFitnessScore(1.0), g.getFitnessScore(1.0)
            print("fitness(max_range=1)", fo, fg, "rel", abs(fo - fg) / fo)
            al = g.aligned_cloud()
            print("aligned equal", np.array_equal(al, O.transform_cloud(Bc, Tg)))

        run(lsq)

    def ndt():
        section("NDT_OMP")
        for res in (1.0, 0.5):
            g = B.Registration(B.default_config(B.NDT_OMP, resolution=res))
            o = O.Registration(O.default_params(O.NDT_OMP, resolution=res))
            g.setInputTarget(A); g.setInputSource(Bc)
            o.setInputTarget(A); o.setInputSource(Bc)
            oi, on, om, oic, omin, odiv = O.ndt_grid(A, res)
            gi, gn, gm, gic, gmin, gdiv = g.debug_ndt_grid()
            print("res", res, "leaves", len(oi), len(gi), "idx equal", np.array_equal(oi, gi), "npts equal", np.array_equal(on, gn), "min_b",
                  omin, gmin, "div_b", odiv, gdiv)
            if len(oi) == len(gi):
                use = on >= 6
                print("mean maxdiff", np.abs(om - gm).max(), "icov rel maxdiff", (np.abs(oic - gic)[use]).max() / np.abs(oic[use]).max())
            for p in (np.zeros(6), np.array([0.4, 0.01, 0.0, 0.001, -0.002, 0.01])):
                os_, og, oH, ohits = o.ndt_derivatives(p)
                gs_, gg, gH, ghits = g.debug_ndt_derivatives(p)
                print("score", os_, gs_, "g rel", np.abs(og - gg).max() / np.abs(og).max(), "H rel", np.abs(oH - gH).max() / np.abs(oH).max(),
                      "hits equal", np.array_equal(ohits, ghits))
            for guess in (np.eye(4), gt):
                t = time.time(); ro = o.align(guess); t_o = time.time() - t
                t = time.time(); rg = g.align(guess); t_g = time.time() - t
                To, Tg = o.getFinalTransformation(), g.getFinalTransformation()
                print("oracle conv", ro.converged, ro.iterations, ro.lm_evals, "gpu", rg.converged, rg.iterations, rg.evals, "T diff", pose_err(To, Tg),
                      "vs gt", pose_err(gt, Tg), "time %.3f %.3f" % (t_o, t_g), g.last_timings())

    run(ndt)

    def batch():
        section("batch")
        g = B.Registration(B.default_config(B.FAST_VGICP))
        scans = [pre(synth.scan(sensor, i)) for i in range(6)]
        clouds = [B.Cloud(g, s) for s in scans]
        src = [clouds[i + 1] for i in range(5)]
        tgt = [clouds[i] for i in range(5)]
        gts = [np.linalg.inv(synth.pose(i)) @ synth.pose(i + 1) for i in range(5)]
        t = time.time(); res = g.align_batch(src, tgt, gts, with_fitness=True); tb = time.time() - t
        print("batch time", tb, g.last_timings(), "launches", g.kernel_launches())
        for i, r in enumerate(res):
            o = O.Registration(O.default_params(O.FAST_VGICP))
            o.setInputTarget(scans[i]); o.setInputSource(scans[i + 1]); ro = o.align(gts[i])
            print(i, "conv", r.converged, ro.converged, "it", r.iterations, ro.iterations, "T diff", pose_err(o.getFinalTransformation(), B.from_colmajor(list(r.T))),
                  "fit", r.fitness, o.getFitnessScore())

    run(batch)


if __name__ == "__main__":
    main()
