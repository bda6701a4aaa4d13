#!/usr/bin/env python
"""How far does a result move if an upstream library differs from this repository's restatement?

The reference pins none of PCL / ndt_omp / fast_gicp and vendors none of them, so the oracle (and with it the GPU engine) follows
"the surveyor's best knowledge of upstream master" (SURVEY.md Appendix A); the details most likely to differ between upstream
versions are marked there.  For each of them the oracle can be switched to the alternative reading (orc_set_variant); this script
runs the affected operation both ways on seeded synthetic scans and reports the movement — a numerical bound on the unpinned-parity
risk.  CPU only.   python tools/oracle_variant_sensitivity.py > profiles/r2/oracle_variant_sensitivity.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mrg_slam_b200 import synth  # noqa: E402
from tests import oraclelib as O  # noqa: E402


def pose_error(Ta, Tb):
    d = np.linalg.inv(Ta) @ Tb
    return float(np.linalg.norm(d[:3, 3])), float(np.arccos(np.clip((np.trace(d[:3, :3]) - 1) / 2, -1, 1)))


def prefilter(c):
    c = O.distance_filter(c, 0.1, 35.0)
    c, _ = O.voxelgrid(c, 0.1, 1)
    return c[O.radius_outlier(c, 0.5, 2)]


def align(method, a, b, guess, **over):
    o = O.Registration(O.default_params(method, **over))
    o.setInputTarget(a); o.setInputSource(b)
    r = o.align(guess)
    return dict(T=o.getFinalTransformation(), converged=bool(r.converged), iterations=int(r.iterations), evals=int(r.lm_evals),
                fitness=float(o.getFitnessScore()))


def registration_cases(scans=((3, 4), (20, 21), (40, 42))):
    out = []
    for i, j in scans:
        a, b = prefilter(synth.scan(synth.VLP16, i)), prefilter(synth.scan(synth.VLP16, j))
        gt = np.linalg.inv(synth.pose(i)) @ synth.pose(j)
        near = gt.copy(); near[0, 3] -= 0.25; near[1, 3] += 0.1
        far = gt.copy(); far[0, 3] += 0.8; far[1, 3] -= 0.6
        c, s = np.cos(0.08), np.sin(0.08)
        far[:3, :3] = far[:3, :3] @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        out += [(a, b, np.eye(4)), (a, b, near), (a, b, far)]
    return out


def measure_registration(name, method, cases, **over):
    dt, dr, dit, dfit, flips = [], [], [], [], 0
    for a, b, g in cases:
        base = align(method, a, b, g, **over)
        with O.variant(name):
            alt = align(method, a, b, g, **over)
        te, re = pose_error(base["T"], alt["T"])
        dt.append(te); dr.append(re)
        dit.append(abs(base["iterations"] - alt["iterations"]))
        dfit.append(abs(base["fitness"] - alt["fitness"]) / max(abs(base["fitness"]), 1e-300))
        flips += int(base["converged"] != alt["converged"])
    return {"cases": len(cases), "max_translation_m": max(dt), "max_rotation_rad": max(dr), "max_iteration_difference": max(dit),
            "max_fitness_relative": max(dfit), "converged_flag_flips": flips}


def run(scans=((3, 4), (20, 21), (40, 42))):
    cases = registration_cases(scans)
    res = {}
    res["vgicp_coord_no_half"] = dict(affects="FAST_VGICP voxel assignment (A.2: floor(x / res - 0.5))",
                                      **measure_registration("vgicp_coord_no_half", O.FAST_VGICP, cases))
    for name, what in (("ndt_angle_eps_1e5", "NDT small-angle cut-off (A.3: the literal 10e-5)"),
                       ("ndt_inner_double", "NDT updateDerivatives in float (current ndt_omp) vs double (PCL / older ndt_omp) (A.3)"),
                       ("mt_clamp_max_first", "More-Thuente trial step clamp order (A.3)"),
                       ("ndt_cov_newer_pcl", "leaf covariance: older single-pass form (ndt_omp) vs newer PCL (A.4)"),
                       ("ndt_lookup_mul", "neighbourhood lookup by division vs multiplication by the inverse leaf (A.4)"),
                       ("euler_no_fixup", "Matrix3f::eulerAngles(0,1,2) first-angle fix-up of Eigen >= 3.3 (A.3)")):
        res[name] = dict(affects=what, **measure_registration(name, O.NDT_OMP, cases))
        if name == "ndt_lookup_mul":  # exact for leaf 1.0 / 0.5; the inverse of 0.3 is inexact in float
            res[name]["leaf_0.3"] = measure_registration(name, O.NDT_OMP, cases[:3], resolution=0.3 * 1.0)
    # ---- filters
    raw = synth.scan(synth.VLP16, 9)
    d0 = O.distance_filter(raw, 0.1, 35.0)
    with O.variant("norm_left_to_right"):
        d1 = O.distance_filter(raw, 0.1, 35.0)
    res["norm_left_to_right"] = dict(affects="distance filter norm association (A.11)", points=len(raw),
                                     kept_differs=int(len(d0) != len(d1) or not np.array_equal(d0, d1)), kept=len(d0), kept_variant=len(d1))
    v0, _ = O.voxelgrid(d0, 0.1, 1)
    with O.variant("voxelgrid_descending"):
        v1, _ = O.voxelgrid(d0, 0.1, 1)
    res["voxelgrid_descending"] = dict(affects="VoxelGrid summation order inside a voxel: upstream's sort is unstable (A.6)", voxels=len(v0),
                                       same_voxels=bool(len(v0) == len(v1)), coordinates_differing=int((v0 != v1).sum()),
                                       max_abs_difference_m=float(np.abs(v0[:, :3] - v1[:, :3]).max()))
    k0 = O.radius_outlier(v0, 0.5, 2)
    with O.variant("radius_nonstrict"):
        k1 = O.radius_outlier(v0, 0.5, 2)
    res["radius_nonstrict"] = dict(affects="RadiusOutlierRemoval d2 < r^2 vs <= r^2 (A.7)", points=len(v0), keep_flags_differing=int((k0 != k1).sum()))
    a, b, g = cases[1]
    T = align(O.FAST_VGICP, a, b, g)["T"]
    t0 = O.transform_cloud(b, T)
    f0 = O.fitness_score(a, b, T)[0]
    with O.variant("transform_left_to_right"):
        t1 = O.transform_cloud(b, T)
        f1 = O.fitness_score(a, b, T)[0]
    res["transform_left_to_right"] = dict(affects="transformPointCloud association, SSE vs scalar path (A.10)", points=len(b),
                                          coordinates_differing=int((t0 != t1).sum()), max_abs_difference_m=float(np.abs(t0 - t1).max()),
                                          fitness_relative=float(abs(f0 - f1) / f0))
    return res


if __name__ == "__main__":
    print(json.dumps(run(), indent=1))
