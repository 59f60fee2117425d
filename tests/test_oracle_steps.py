"""The oracle's restated sub-steps against the third-party routines the reference calls
(scipy.ndimage.gaussian_filter1d, scipy.signal.find_peaks, numpy mean/std) -- bit for bit."""
import itertools
import math

import numpy as np
import pytest
from scipy.ndimage import gaussian_filter1d
from scipy.signal import find_peaks

from oracle import segment_oracle as orc


def _signals(rng, n_trials, max_n=400):
    for _ in range(n_trials):
        n = int(rng.integers(2, max_n))
        y = np.zeros(n)
        k = int(rng.integers(0, max(1, n // 3)))
        idx = rng.integers(0, n, size=k)
        np.add.at(y, idx, rng.integers(1, 50, size=k).astype(float))
        yield y


@pytest.mark.parametrize("sigma", [1.0, 2.5, 5.0, 12.0, 50.0])
def test_gaussian_reflect_bit_exact(sigma):
    rng = np.random.default_rng(int(sigma * 10))
    w = orc.gaussian_weights(sigma, 4.0)
    for y in _signals(rng, 60):
        assert np.array_equal(orc.gaussian_filter(y, w, "reflect"), gaussian_filter1d(y, sigma, truncate=4.0))


@pytest.mark.parametrize("sigma", [1.0, 3.3, 5.0, 7.5, 20.0])
def test_gaussian_constant_bit_exact(sigma):
    rng = np.random.default_rng(int(sigma * 7))
    w = orc.gaussian_weights(sigma, 1.0)
    for y in _signals(rng, 60):
        ref = gaussian_filter1d(list(y), sigma, mode="constant", cval=0.0, truncate=1.0)
        assert np.array_equal(orc.gaussian_filter(y, w, "constant"), ref)


def test_pairwise_sum_matches_numpy():
    rng = np.random.default_rng(3)
    for n in list(range(0, 140)) + [255, 256, 257, 1000, 1023, 4097, 20011]:
        a = rng.random(n) * rng.integers(1, 1000)
        if n == 0:
            assert orc.pairwise_sum(a) == 0.0
            continue
        assert orc.pairwise_sum(a) == float(np.add.reduce(a)), n


def test_variance_threshold_matches_numpy():
    rng = np.random.default_rng(4)
    for trial in range(40):
        Y = [rng.random(int(rng.integers(1, 900))) * (rng.random() < 0.7) for _ in range(int(rng.integers(1, 6)))]
        Y = [np.where(rng.random(len(y)) < 0.4, 0.0, y) for y in Y]
        nz = np.array([v for y in Y for v in y if v > 0])
        vf = float(rng.uniform(0.1, 9))
        if len(nz) == 0:
            assert math.isnan(orc.variance_threshold(Y, vf))
            continue
        assert orc.variance_threshold(Y, vf) == nz.mean() + vf * nz.std()


def test_local_maxima_matches_scipy_including_plateaus():
    rng = np.random.default_rng(5)
    for _ in range(300):
        n = int(rng.integers(1, 120))
        x = rng.integers(0, 4, size=n).astype(float)  # few levels => many plateaus
        want = find_peaks(x)[0].tolist()
        assert orc.local_maxima(x) == want
        assert orc.local_maxima_np(x).tolist() == want


def test_plateau_midpoint_on_smoothed_equal_spikes():
    # equal spikes an odd distance apart: bit-equal centre samples, floor midpoint is chosen
    y = np.zeros(200)
    y[80] = y[89] = 12.0
    g = orc.gaussian_filter(y, orc.gaussian_weights(5.0, 4.0), "reflect")
    assert g[84] == g[85]
    assert orc.candidates(g) == [0, 84, 199]
    assert find_peaks(g)[0].tolist() == [84]


def test_select_by_distance_matches_scipy_without_ties():
    rng = np.random.default_rng(6)
    for _ in range(200):
        n = int(rng.integers(30, 400))
        x = rng.random(n)
        peaks = orc.local_maxima(x)
        got = orc.select_by_distance(peaks, [x[p] for p in peaks], 20)
        assert got == find_peaks(x, distance=20)[0].tolist()


def test_smooth_threshold_table_lengths():
    assert len(orc.smooth_threshold(0.9)) == 100
    assert len(orc.smooth_threshold(0.5)) == 6
    assert len(orc.smooth_threshold(1.0)) == 108
    assert orc.smooth_threshold(0.9)[-1] == 0.89 and orc.pair_thresholds(10 ** 6, orc.smooth_threshold(0.9), 0.9) == (
        0.9, 0.09999999999999998)


def test_coverage_closed_form_known_answer():
    # one rep with two intervals [2,5] and [9,12] (inclusive samples), candidates 0,4,10,14
    C = orc.coverage_matrix([[(2, 5), (9, 12)], []], [0, 4, 10, 14])
    assert C[:, 0].tolist() == [0, 2, 5, 8, 8]
    assert C[:, 1].tolist() == [0, 0, 0, 0, 0]


def _brute_best(n, cv, ins, out, lo):
    best = int(ins[0, n - 1])
    for r in range(2, n):
        for mid in itertools.combinations(range(1, n - 1), r - 1):
            chain = (0,) + mid + (n - 1,)
            if any(cv[b] - cv[a] < 5 for a, b in zip(chain[:-1], chain[1:])):
                continue
            sc = sum(int(ins[a, b]) for a, b in zip(chain[:-1], chain[1:]))
            ok = True
            for a, b, c in zip(chain[:-2], chain[1:-1], chain[2:]):
                if out[a, b, c] < lo:
                    ok = False
                    break
                sc += int(out[a, b, c])
            if ok:
                best = max(best, sc)
    return best


def test_dp_solve_is_optimal_on_random_tables():
    rng = np.random.default_rng(7)
    for _ in range(150):
        n = int(rng.integers(3, 9))
        cv = np.cumsum(rng.integers(1, 12, size=n)).tolist()
        ins = -rng.integers(0, 6, size=(n, n)).astype(np.int64)
        out = rng.integers(0, 9, size=(n, n, n)).astype(np.int64)
        lo = int(rng.integers(0, 5))
        chosen = orc.dp_solve(cv, 0, n - 1, ins, out, lo)
        # score of the returned chain
        chain = sorted(set([0, n - 1] + chosen))
        sc = sum(int(ins[a, b]) for a, b in zip(chain[:-1], chain[1:]))
        sc += sum(int(out[a, b, c]) for a, b, c in zip(chain[:-2], chain[1:-1], chain[2:]))
        assert sc == _brute_best(n, cv, ins, out, lo)


def test_thread_cigar_clips_insertions_too():
    cig = [(5, "M"), (4, "I"), (3, "D"), (6, "M")]
    assert orc.thread_cigar(cig, 105, 100, 10) == 15
    assert orc.thread_cigar(cig, 107, 100, 10) == 17  # I clipped to the 2 remaining target bases (quirk)
    assert orc.thread_cigar(cig, 108, 100, 10) == 18


def test_longest_poly_known_answers():
    seq = "A" * 25 + "C" + "A" * 4
    runs = list(orc.longest_poly([c == "A" for c in seq]))
    assert runs == [(0, 30, 29 / 30)]
    seq = "CC" + "A" * 22 + "GGGGGGGGGGGGG" + "A" * 3
    runs = list(orc.longest_poly([c == "A" for c in seq]))
    assert runs[0] == (2, 22, 1.0)
