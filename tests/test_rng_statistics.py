"""Statistical checks of the product's motion-noise generator (Philox4x32-10 + Box-Muller, common.cuh
philox_normals3; restated draw for draw in oracle/c/mcl_oracle.c orc_normals3, and the GPU kernel is tested equal to
that restatement attempt by attempt in test_gpu_parity.py).  Draw-for-draw equality with our own restatement says
nothing about the DISTRIBUTION, and the radius word of an attempt is assembled from two Philox blocks (its high half
comes from a block shared by eight consecutive attempts), so: moments, Kolmogorov-Smirnov against N(0,1), independence
of the three normals of an attempt, and serial correlation between attempts t, t+1 ... t+7 of one particle (inside one
shared block and across two).  Fixed seeds: the numbers are deterministic; the bounds are ~4.5 standard errors."""
import numpy as np
import pytest
from scipy import stats

from oracle import clib

N_ITEMS, N_ATT = 12000, 16


@pytest.fixture(scope="module")
def draws():
    z = np.array([[clib.normals3(2024, 9, i, a) for a in range(N_ATT)] for i in range(N_ITEMS)])
    return z                                                       # (items, attempts, 3)


def test_normals_moments_and_ks(draws):
    f = draws.reshape(-1, 3)
    n = len(f)
    se = 1.0 / np.sqrt(n)
    assert np.all(np.abs(f.mean(0)) < 4.5 * se)
    assert np.all(np.abs(f.var(0) - 1.0) < 4.5 * np.sqrt(2.0) * se)
    assert np.all(np.abs(stats.skew(f, axis=0)) < 4.5 * np.sqrt(6.0) * se)
    assert np.all(np.abs(stats.kurtosis(f, axis=0)) < 4.5 * np.sqrt(24.0) * se)
    for k in range(3):
        assert stats.kstest(f[:, k], "norm").pvalue > 1e-3, k
    # tails: Box-Muller on 32-bit uniforms reaches |z| ~ 6.6; the 4-sigma tail mass must be there
    tail = np.mean(np.abs(f) > 4.0)
    assert 0.3 * 6.33e-5 < tail < 3.0 * 6.33e-5


def test_normals_of_one_attempt_are_independent(draws):
    f = draws.reshape(-1, 3)
    se = 1.0 / np.sqrt(len(f))
    c = np.corrcoef(f.T)
    assert abs(c[0, 1]) < 4.5 * se and abs(c[0, 2]) < 4.5 * se and abs(c[1, 2]) < 4.5 * se
    # z0, z1 share a radius: uncorrelated but their squares are not independent of it -- the angle must be uniform
    ang = np.arctan2(f[:, 1], f[:, 0])
    assert stats.kstest((ang + np.pi) / (2 * np.pi), "uniform").pvalue > 1e-3
    r2 = f[:, 0] ** 2 + f[:, 1] ** 2                               # chi^2 with 2 degrees of freedom
    assert stats.kstest(r2, "chi2", args=(2,)).pvalue > 1e-3


def test_no_serial_correlation_between_attempts_sharing_a_radius_block(draws):
    z0 = draws[:, :, 0]
    rad = np.hypot(draws[:, :, 0], draws[:, :, 1])
    for d in range(1, 8):
        # pairs (t, t + d) inside the first shared block (attempts 0..7)
        a, b = z0[:, 0:8 - d].ravel(), z0[:, d:8].ravel()
        assert abs(np.corrcoef(a, b)[0, 1]) < 4.5 / np.sqrt(len(a)), d
        ra, rb = rad[:, 0:8 - d].ravel(), rad[:, d:8].ravel()
        assert abs(np.corrcoef(ra, rb)[0, 1]) < 4.5 / np.sqrt(len(ra)), d
    # across the block boundary (attempt 7 | 8) and between particles (item i | i + 1, same attempt)
    assert abs(np.corrcoef(rad[:, 7], rad[:, 8])[0, 1]) < 4.5 / np.sqrt(N_ITEMS)
    assert abs(np.corrcoef(z0[:-1].ravel(), z0[1:].ravel())[0, 1]) < 4.5 / np.sqrt(z0[:-1].size)
    # the shared HIGH half of the radius word: radii of the same block must not cluster (rank correlation)
    assert abs(stats.spearmanr(rad[:, 2], rad[:, 3]).statistic) < 4.5 / np.sqrt(N_ITEMS)


def test_mh_and_resampling_uniforms_are_uniform():
    u = np.array([clib.uniform53(7, 3, i, 0, clib.STREAM_MH) for i in range(100000)])
    assert stats.kstest(u, "uniform").pvalue > 1e-3
    assert abs(np.corrcoef(u[:-1], u[1:])[0, 1]) < 4.5 / np.sqrt(len(u))
    assert u.min() >= 0.0 and u.max() < 1.0
