"""Generates tests/golden/golden_v1.npz from the oracle on seeded synthetic scans.

The reference ships no golden vectors or fixtures and none of its numeric libraries can be built or imported in
this environment, so these vectors cannot come from the real pclomp / fast_gicp / PCL code: they are
the oracle's own outputs, frozen so that (1) any later change to the oracle is caught (`tests/test_golden.py`) and
(2) the GPU path can be checked against committed numbers.  Parity with the upstream binaries stays UNPINNED.

    python tests/golden/make_golden.py        # rewrites golden_v1.npz
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mrg_slam_b200 import synth  # noqa: E402
from tests import oraclelib as O  # noqa: E402

SCAN_A, SCAN_B = 3, 4
GENERATOR_VERSION = 1


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def inputs():
    raw_a, raw_b = synth.scan(synth.VLP16, SCAN_A), synth.scan(synth.VLP16, SCAN_B)
    return raw_a, raw_b


def prefilter(raw):
    d = O.distance_filter(raw, 0.1, 35.0)
    v, _, vidx = O.voxelgrid(d, 0.1, 1, want_index=True)
    keep = O.radius_outlier(v, 0.5, 2)
    keep_s, dist, thr = O.statistical_outlier(v, 30, 1.2)
    return d, v, vidx, keep, keep_s, dist, thr


def guesses(gt):
    g1 = gt.copy(); g1[0, 3] -= 0.3
    g2 = gt.copy(); g2[1, 3] += 0.15; g2[0, 3] += 0.1
    return [np.eye(4), g1, g2]


def build():
    raw_a, raw_b = inputs()
    out = {"generator_version": np.array(GENERATOR_VERSION), "raw_a_sha": np.array(sha(raw_a)), "raw_b_sha": np.array(sha(raw_b))}
    d, v, vidx, keep, keep_s, dist, thr = prefilter(raw_a)
    out.update(dist_n=np.array(len(d)), dist_sha=np.array(sha(d)), vg_n=np.array(len(v)), vg_sha=np.array(sha(v)), vg_idx_sha=np.array(sha(vidx)),
               rad_n=np.array(int(keep.sum())), rad_sha=np.array(sha(v[keep])), sor_n=np.array(int(keep_s.sum())), sor_sha=np.array(sha(v[keep_s])),
               sor_thr=np.array(thr))
    A = v[keep]
    _, vb, _, kb, _, _, _ = prefilter(raw_b)
    B = vb[kb]
    gt = np.linalg.inv(synth.pose(SCAN_A)) @ synth.pose(SCAN_B)
    cov, knn = O.knn_covariances(B, 20, want_idx=True)
    out.update(cov_head=cov[:256], knn_sorted_sha=np.array(sha(np.sort(knn, 1))))
    coords, npts, mean, vcov = O.vgicp_voxelmap(A, O.knn_covariances(A, 20), 1.0)
    out.update(vox_coords_sha=np.array(sha(coords)), vox_npts_sha=np.array(sha(npts)), vox_n=np.array(len(coords)), vox_mean_head=mean[:64],
               vox_cov_head=vcov[:64])
    for res in (1.0, 0.5):
        idx, n, m, icov, min_b, div_b = O.ndt_grid(A, res)
        tag = str(res).replace(".", "p")
        out.update({f"ndt_idx_sha_{tag}": np.array(sha(idx)), f"ndt_npts_sha_{tag}": np.array(sha(n)), f"ndt_n_{tag}": np.array(len(idx)),
                    f"ndt_min_b_{tag}": min_b, f"ndt_div_b_{tag}": div_b})
    for name, method in (("vgicp", O.FAST_VGICP), ("gicp", O.FAST_GICP), ("ndt", O.NDT_OMP)):
        r = O.Registration(O.default_params(method))
        r.setInputTarget(A); r.setInputSource(B)
        Ts, conv, its, fits = [], [], [], []
        for g in guesses(gt):
            res = r.align(g)
            Ts.append(np.array(list(res.T), dtype=np.float32)); conv.append(res.converged); its.append(res.iterations)
            fits.append(r.getFitnessScore())
        out.update({f"{name}_T": np.stack(Ts), f"{name}_conv": np.array(conv), f"{name}_iters": np.array(its), f"{name}_fitness": np.array(fits)})
        if method != O.NDT_OMP:
            err, H, b, corr, valid = r.linearize(gt)
            out.update({f"{name}_lin_err": np.array(err), f"{name}_lin_H": H, f"{name}_lin_b": b, f"{name}_corr_sha": np.array(sha(corr[valid])),
                        f"{name}_valid_sha": np.array(sha(valid))})
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path), "bytes")
