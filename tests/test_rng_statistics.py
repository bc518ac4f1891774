"""Statistical checks of the product's motion-noise generator (Philox4x32-10 + Box-Muller, common.cuh
normals3_from_words; restated draw for draw in oracle/c/mcl_oracle.c, and the GPU kernel is tested equal to that
restatement attempt by attempt in test_gpu_parity.py).  Draw-for-draw equality with our own restatement says nothing
about the DISTRIBUTION, and the radius word of an attempt is assembled from up to three Philox blocks (its top nibble
comes from a block shared by 32 consecutive attempts; after a zero nibble the next 16 bits are dealt by rank from
another stream), so: moments, Kolmogorov-Smirnov against N(0,1), independence of the three normals of an attempt,
serial correlation between neighbouring attempts of one particle, and the radius distribution of the zero-nibble
attempts on their own (the far tail, which is what lets a particle leave a wall).  Fixed seeds: the numbers are
deterministic; the bounds are ~4.5 standard errors."""
import numpy as np
import pytest
from scipy import stats

from oracle import clib

N_ITEMS, N_ATT = 12000, 40


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


def test_radius_words_are_uniform_including_the_rank_dealt_tail():
    """u1 = (radius word + 1) / 2^32 recovered from the normals (R1^2 = z0^2 + z1^2 = -2 ln u1): uniform over all
    attempts >= 1, and uniform on (0, 1/16] for the attempts whose top nibble is zero (dealt by rank from MOTION_R2)."""
    z = np.array([clib.normals3_seq(77, 3, i, 200)[1:] for i in range(1500)]).reshape(-1, 3)
    u1 = np.exp(-0.5 * (z[:, 0] ** 2 + z[:, 1] ** 2))
    assert stats.kstest(u1, "uniform").pvalue > 1e-3
    tail = u1[u1 <= 1.0 / 16.0] * 16.0
    assert abs(len(tail) / len(u1) - 1.0 / 16.0) < 4.5 * np.sqrt(15.0 / 256.0 / len(u1))
    assert stats.kstest(tail, "uniform").pvalue > 1e-3
    far = np.mean(u1 <= 2.0 ** -12)                               # R1 >= 4.08: one attempt in 4096
    assert abs(far - 2.0 ** -12) < 4.5 * np.sqrt(2.0 ** -12 / len(u1))


def test_mh_and_resampling_uniforms_are_uniform():
    u = np.array([clib.uniform53(7, 3, i, 0, clib.STREAM_MH) for i in range(100000)])
    assert stats.kstest(u, "uniform").pvalue > 1e-3
    assert abs(np.corrcoef(u[:-1], u[1:])[0, 1]) < 4.5 / np.sqrt(len(u))
    assert u.min() >= 0.0 and u.max() < 1.0
