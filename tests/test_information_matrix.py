"""InformationMatrixCalculator::calc_information_matrix (/root/reference/src/mrg_slam/information_matrix_calculator.cpp:14-44, weight() :83-88):
the caller of the fitness call.  CPU: the weighting with a stub engine against the closed form and the oracle's fitness;
GPU: the mirror on the engine against the same mirror on the oracle."""
import numpy as np
import pytest

from mrg_slam_b200 import loop_closure as LC
from tests import oraclelib as O


class _Fixed:
    def __init__(self, f):
        self.f = f

    def fitness_pair(self, a, b, T):
        return self.f


class _OracleFitness:
    def fitness_pair(self, a, b, T):
        return O.fitness_score(a, b, T)[0]


def test_weighting_matches_the_closed_form():
    for f in (0.0, 0.01, 0.3, 1.25, 5.0):
        inf = LC.calc_information_matrix(_Fixed(f), None, None, np.eye(4))
        y = (1.0 - np.exp(-2.0 * f)) / (1.0 - np.exp(-2.0 * 1.25))
        wx = 0.1 ** 2 + (0.75 ** 2 - 0.1 ** 2) * y
        wq = 0.05 ** 2 + (0.2 ** 2 - 0.05 ** 2) * y
        want = np.diag([1 / wx] * 3 + [1 / wq] * 3)
        np.testing.assert_allclose(inf, want, rtol=1e-14, atol=0)
    # a perfect fit gives the minimum variances, a fit at the threshold the maximum ones (weight() is monotone in the score)
    assert np.isclose(LC.calc_information_matrix(_Fixed(0.0), None, None, np.eye(4))[0, 0], 1 / 0.1 ** 2)
    assert np.isclose(LC.calc_information_matrix(_Fixed(1.25), None, None, np.eye(4))[5, 5], 1 / 0.2 ** 2)
    const = LC.calc_information_matrix(_Fixed(123.0), None, None, np.eye(4), use_const_inf_matrix=True)
    np.testing.assert_array_equal(const, np.diag([2.0] * 3 + [10.0] * 3))  # identity / stddev (:20-21), not / variance


def test_oracle_fitness_feeds_the_mirror(vlp16_pair):
    a, b, gt = vlp16_pair
    inf = LC.calc_information_matrix(_OracleFitness(), a, b, gt)
    f = O.fitness_score(a, b, gt)[0]
    assert 0 < f < 1.25 and inf[0, 0] == pytest.approx(1 / LC.information_weight(2.0, 1.25, 0.01, 0.5625, f), rel=1e-14)
    assert np.count_nonzero(inf - np.diag(np.diag(inf))) == 0


@pytest.mark.gpu
def test_gpu_information_matrix_matches_oracle(vlp16_pair):
    from mrg_slam_b200 import lib as B
    a, b, gt = vlp16_pair
    reg = B.Registration(B.default_config(B.FAST_VGICP))
    ca, cb = B.Cloud(reg, a), B.Cloud(reg, b)
    for T in (gt, np.eye(4)):
        got = LC.calc_information_matrix(reg, ca, cb, T)
        want = LC.calc_information_matrix(_OracleFitness(), a, b, T)
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=0)
    ca.close(); cb.close(); reg.close()
