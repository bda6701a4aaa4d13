"""Generates tests/golden/golden_v3.npz: LoopDetector::matching with the consistency check
(/root/reference/src/mrg_slam/loop_detector.cpp:97-303) on the scene of tests/test_loop_consistency.py, frozen from the
oracle (FAST_GICP) through mrg_slam_b200.loop_closure.match_keyframes: best candidate, score, relative pose, the identity-check
deltas and the decision for every edge-corruption case.  Same caveat as make_golden.py: the reference ships no vectors;
these freeze the checker and the host logic.

    python tests/golden/make_golden_v3.py     # rewrites golden_v3.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mrg_slam_b200 import loop_closure as LC  # noqa: E402
from tests import oraclelib as O  # noqa: E402
from tests.test_loop_consistency import OracleBatch, _scene  # noqa: E402

CASES = ["ok", "prev_bad", "both_bad", "no_prev", "first", "disabled"]


def case_links(links, name):
    shift = np.eye(4)
    shift[0, 3] = 1.0
    base = links[2]
    lk = dict(links)
    lk[2] = LC.KeyframeLinks(first_keyframe=name == "first",
                             prev=None if name == "no_prev" else base.prev,
                             rel_pose_to_prev=base.rel_pose_to_prev @ shift if name in ("prev_bad", "both_bad", "first", "disabled") else base.rel_pose_to_prev,
                             next=base.next,
                             rel_pose_from_next=base.rel_pose_from_next @ shift if name in ("both_bad", "first", "disabled") else base.rel_pose_from_next)
    return lk


def build(make_reg=None):
    make_reg = make_reg or (lambda: OracleBatch(O.FAST_GICP))
    clouds, est, pairs, guesses, links = _scene()
    out = {"generator_version": np.array(3), "points": np.array([len(c) for c in clouds])}
    for name in CASES:
        accepted, loops, details, table = LC.match_keyframes(make_reg(), clouds, pairs, guesses, est, case_links(links, name),
                                                             enable_loop_closure_consistency_check=name != "disabled")
        lp = loops[0]
        out[f"{name}_source"] = np.array(-1 if lp.source is None else lp.source)
        out[f"{name}_score"] = np.array(lp.best_score)
        out[f"{name}_pose"] = np.asarray(lp.relative_pose, dtype=np.float32)
        out[f"{name}_accepted"] = np.array(len(accepted))
        d = details[0]
        out[f"{name}_deltas"] = np.array([d.get("prev", {}).get("delta_trans", -1.0), d.get("prev", {}).get("delta_angle", -1.0),
                                          d.get("next", {}).get("delta_trans", -1.0), d.get("next", {}).get("delta_angle", -1.0)])
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v3.npz")
    np.savez_compressed(path, **build())
    print("wrote", path)
