"""Bounds on the unpinned-parity risk: for every detail SURVEY.md Appendix A marks as upstream-version
dependent the oracle is switched to the alternative reading and the movement of the result is measured
(tools/oracle_variant_sensitivity.py; the full table is committed as profiles/r2/oracle_variant_sensitivity.json).

What the numbers say: nine of the eleven variants (arithmetic association, float vs double inner math of NDT, the small-angle
cut-off, clamp order, lookup by division, the Euler fix-up, strict vs non-strict radius, in-voxel summation order) move a final
transform by less than the north star's own tolerance (1e-4 m / 1e-4 rad) or not at all.  Two are STRUCTURAL — they change the
cost function itself: the -0.5 in fast_gicp's voxel coordinate (millimetres to a centimetre on VGICP) and the older vs newer
PCL form of the NDT leaf covariance (a factor (n/(n-1))^2 on every covariance: different iteration counts, and a far guess can
settle in another basin).  Both would show up at once against a real fast_gicp / ndt_omp dump
(tests/golden/import_reference_dump.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import oracle_variant_sensitivity as S  # noqa: E402
from tests import oraclelib as O  # noqa: E402


def test_every_variant_switch_is_exercised_and_bounded():
    res = S.run(scans=((3, 4),))
    assert set(res) == set(O.VARIANTS)
    for v in O.VARIANTS.values():
        assert O.lib().orc_get_variant(v) == 0  # the context manager restored the documented choice
    # association-only variants: rounding-level effects
    assert res["transform_left_to_right"]["max_abs_difference_m"] < 1e-5 and res["transform_left_to_right"]["fitness_relative"] < 1e-5
    assert res["voxelgrid_descending"]["same_voxels"] and res["voxelgrid_descending"]["max_abs_difference_m"] < 1e-5
    assert res["norm_left_to_right"]["kept_differs"] in (0, 1) and abs(res["norm_left_to_right"]["kept"] - res["norm_left_to_right"]["kept_variant"]) <= 2
    assert res["radius_nonstrict"]["keep_flags_differing"] <= 2
    # every non-structural variant: inside the north star's tolerance (1e-4 m / 1e-4 rad), same iteration counts, same flags
    for name in ("ndt_angle_eps_1e5", "ndt_inner_double", "mt_clamp_max_first", "ndt_lookup_mul", "euler_no_fixup"):
        r = res[name]
        assert r["converged_flag_flips"] == 0 and r["max_iteration_difference"] == 0, (name, r)
        assert r["max_translation_m"] <= 1e-4 and r["max_rotation_rad"] <= 1e-4, (name, r)
    assert res["ndt_lookup_mul"]["max_translation_m"] < 1e-12  # leaf 1.0: division and multiplication by the inverse agree
    # the two structural variants are large enough to be caught by any real fast_gicp / ndt_omp output
    assert res["vgicp_coord_no_half"]["max_translation_m"] > 1e-4
    assert res["ndt_cov_newer_pcl"]["max_translation_m"] > 1e-4 or res["ndt_cov_newer_pcl"]["max_iteration_difference"] > 0
