"""Host mirror of LoopDetector::matching + perform_loop_closure_consistency_check
(/root/reference/src/mrg_slam/loop_detector.cpp:97-303): the batched implementation in mrg_slam_b200/loop_closure.py against a
literal, sequential restatement of the reference's control flow written here, both driven by the CPU oracle (CPU test) and by
the product (GPU test)."""
import numpy as np
import pytest

from mrg_slam_b200 import loop_closure as LC
from mrg_slam_b200 import synth
from tests import oraclelib as O


class _Res:
    def __init__(self, T, converged, iterations, error, evals, fitness):
        self.T, self.converged, self.iterations, self.error, self.evals, self.fitness = T, converged, iterations, error, evals, fitness


class OracleBatch:
    """align_batch surface over the oracle's single-pair registration (clouds are numpy arrays)."""

    def __init__(self, method):
        self.reg = O.Registration(O.default_params(method))
        self.aligns = 0

    def align_batch(self, sources, targets, guesses, with_fitness=False, fitness_max_range=LC.DBL_MAX):
        out = []
        for s, t, g in zip(sources, targets, guesses):
            self.reg.setInputTarget(t)
            self.reg.setInputSource(s)
            r = self.reg.align(g)
            self.aligns += 1
            fit = self.reg.getFitnessScore(fitness_max_range) if with_fitness else 0.0
            out.append(_Res(list(r.T), r.converged, r.iterations, r.error, getattr(r, "lm_evals", 0), fit))
        return out


def _scene():
    """One new keyframe (cloud 0) with three candidates (1..3); 4 and 5 are the previous / next keyframes of candidate 2."""
    idx = [30, 20, 21, 22, 19, 23]  # scan numbers; 21 is nearest to the truth of the matching below

    def prep(c):
        c = O.distance_filter(c, 0.5, 30.0)
        c, _ = O.voxelgrid(c, 0.3, 1)
        return c

    clouds = [prep(synth.scan(synth.VLP16, i)) for i in idx]
    # the "new keyframe" is scan 21 seen again (a revisit): its cloud is scan 21 displaced by a known transform
    poses = [synth.pose(i) for i in idx]
    rng = np.random.default_rng(7)
    D = np.eye(4)
    D[:3, 3] = [0.3, -0.2, 0.02]
    c, s = np.cos(0.04), np.sin(0.04)
    D[:3, :3] = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
    new = clouds[2].copy()
    new[:, :3] = (new[:, :3] - D[:3, 3]) @ D[:3, :3]  # points of scan 21 expressed in the new keyframe's frame: p_new = D^-1 p_21
    clouds[0] = new.astype(np.float32)
    poses[0] = poses[2] @ D  # world <- new
    est = {i: p.copy() for i, p in enumerate(poses)}
    for i in est:  # graph estimates: truth + small drift, rotation slightly denormalised (what normalize_estimate is for)
        est[i][:3, 3] += rng.normal(scale=0.03, size=3)
        est[i][:3, :3] *= 1.0 + 1e-7
    pairs = [(0, 1), (0, 2), (0, 3)]
    guesses = [LC.registration_guess(est[0], est[c]).astype(np.float64) for _, c in pairs]
    rel = lambda a, b: np.linalg.inv(poses[a]) @ poses[b]  # a <- b
    links = {c: LC.KeyframeLinks() for c in range(6)}
    links[2] = LC.KeyframeLinks(prev=4, rel_pose_to_prev=rel(2, 4), next=5, rel_pose_from_next=rel(5, 2))
    links[1] = LC.KeyframeLinks(prev=4, rel_pose_to_prev=rel(1, 4), next=2, rel_pose_from_next=rel(2, 1))
    links[3] = LC.KeyframeLinks(prev=2, rel_pose_to_prev=rel(3, 2), next=5, rel_pose_from_next=rel(5, 3))
    return clouds, est, pairs, guesses, links


def reference_matching(make_reg, clouds, est, pairs, links, thresh=1.25, enable=True, max_dt=0.3, max_da=0.0523599):
    """loop_detector.cpp:97-180 + :190-303 for ONE new keyframe, statement by statement, on a single registration object."""
    reg = make_reg()
    new = pairs[0][0]
    reg.setInputTarget(clouds[new])                                          # :104
    best_score, best, T_best = LC.DBL_MAX, None, None
    new_est = LC.normalize_estimate(est[new])                                # :124
    for _, cand in pairs:
        reg.setInputSource(clouds[cand])                                     # :127
        guess = (np.linalg.inv(new_est) @ LC.normalize_estimate(est[cand])).astype(np.float32)   # :129-130
        r = reg.align(guess.astype(np.float64))                              # :134
        score = reg.getFitnessScore(LC.DBL_MAX)                              # :137
        if not r.converged or score > best_score:                            # :138
            continue
        best_score, best, T_best = score, cand, reg.getFinalTransformation().astype(np.float32)
    log = {}

    def check():                                                             # :190-218
        if best is not None and (links[best].first_keyframe or links[best].static_keyframe):
            return True
        if best is None or not enable or best_score > thresh:
            return False
        lk = links[best]
        if lk.prev is not None:                                              # :225-262
            reg.setInputSource(clouds[lk.prev])
            g = (np.linalg.inv(new_est) @ LC.normalize_estimate(est[lk.prev])).astype(np.float32)
            reg.align(g.astype(np.float64))
            Tp = reg.getFinalTransformation().astype(np.float32)
            M = np.linalg.inv(Tp) @ T_best @ lk.rel_pose_to_prev.astype(np.float32)
            log["prev"] = LC._delta(M)
            if not (log["prev"][0] > max_dt or log["prev"][1] > max_da):
                return True
        if lk.next is None:                                                  # :269-303
            return False
        reg.setInputSource(clouds[lk.next])
        g = (np.linalg.inv(new_est) @ LC.normalize_estimate(est[lk.next])).astype(np.float32)
        reg.align(g.astype(np.float64))
        Tn = reg.getFinalTransformation().astype(np.float32)
        M = np.linalg.inv(T_best) @ Tn @ lk.rel_pose_from_next.astype(np.float32)
        log["next"] = LC._delta(M)
        return not (log["next"][0] > max_dt or log["next"][1] > max_da)

    passed = check()
    if best_score > thresh:                                                  # :156
        return None, best, best_score, passed, log
    if enable and best is not None and not links[best].first_keyframe and not passed:   # :162
        return None, best, best_score, passed, log
    return (new, best, T_best), best, best_score, passed, log


def _run_cases(make_batch, make_single, atol):
    clouds, est, pairs, guesses, links = _scene()
    shift = np.eye(4)
    shift[0, 3] = 1.0
    cases = {
        "consistent_via_prev": {},
        "bad_prev_good_next": {"prev_bad": True},
        "both_bad": {"prev_bad": True, "next_bad": True},
        "no_prev_edge": {"no_prev": True},
        "first_keyframe": {"first": True, "prev_bad": True, "next_bad": True},
        "disabled": {"enable": False, "prev_bad": True, "next_bad": True},
    }
    for name, c in cases.items():
        lk = dict(links)
        base = links[2]
        lk[2] = LC.KeyframeLinks(first_keyframe=c.get("first", False),
                                 prev=None if c.get("no_prev") else base.prev,
                                 rel_pose_to_prev=base.rel_pose_to_prev @ shift if c.get("prev_bad") else base.rel_pose_to_prev,
                                 next=base.next,
                                 rel_pose_from_next=base.rel_pose_from_next @ shift if c.get("next_bad") else base.rel_pose_from_next)
        enable = c.get("enable", True)
        want, best, score, passed, log = reference_matching(make_single, clouds, est, pairs, lk, enable=enable)
        reg = make_batch()
        accepted, loops, details, table = LC.match_keyframes(reg, clouds, pairs, guesses, est, lk,
                                                             enable_loop_closure_consistency_check=enable)
        assert best == 2 and loops[0].source == 2, name  # the revisit of scan 21 matches candidate 2
        assert abs(loops[0].best_score - score) <= 1e-9 * score
        assert (len(accepted) == 1) == (want is not None), name
        for k in ("prev", "next"):
            assert (k in details[0]) == (k in log), (name, k, details[0], log)
            if k in log:
                assert abs(details[0][k]["delta_trans"] - log[k][0]) <= atol and abs(details[0][k]["delta_angle"] - log[k][1]) <= atol
        if want is not None:
            assert np.allclose(accepted[0].relative_pose, want[2], atol=1e-6)
        expect = {"consistent_via_prev": True, "bad_prev_good_next": True, "both_bad": False, "no_prev_edge": True, "first_keyframe": True,
                  "disabled": True}[name]
        assert (want is not None) == expect, (name, log)


def test_matching_with_consistency_check_on_the_oracle():
    _run_cases(lambda: OracleBatch(O.FAST_GICP), lambda: O.Registration(O.default_params(O.FAST_GICP)), 1e-6)


@pytest.mark.gpu
def test_matching_with_consistency_check_on_the_gpu():
    from mrg_slam_b200 import lib as B

    class GpuBatch:
        """numpy clouds in, like OracleBatch: uploads per call (the test's clouds are tiny)."""

        def __init__(self):
            self.reg = B.Registration(B.default_config(B.FAST_GICP))

        def align_batch(self, sources, targets, guesses, with_fitness=False, fitness_max_range=LC.DBL_MAX):
            cache = {}

            def up(c):
                if id(c) not in cache:
                    cache[id(c)] = B.Cloud(self.reg, c)
                return cache[id(c)]

            out = self.reg.align_batch([up(s) for s in sources], [up(t) for t in targets], guesses, with_fitness=with_fitness,
                                       fitness_max_range=fitness_max_range)
            for c in cache.values():
                c.close()
            return out

    # product (batched mirror) against the oracle (sequential restatement): decisions equal, deltas within the pose tolerance
    _run_cases(GpuBatch, lambda: O.Registration(O.default_params(O.FAST_GICP)), 2e-4)
