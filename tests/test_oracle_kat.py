"""Known answers the reference states in its own docstrings."""
import numpy as np

import pfx_oracle as O


def test_sum_over_residuals_docstring():
    # ProfileOperations.py:213-221: P = [10,13,17,50,20,10,5] -> 32.5
    s = O.sinusoid_fittings(np.array([10, 13, 17, 50, 20, 10, 5]))
    assert s[3] == 32.5


def test_peak_indexes_docstring():
    # ProfileOperations.py:271-289: two peaks at indexes 3 and 11
    prof = [0, 0, 5, 10, 5, 0, 0, 0, 0, 0, 0, 15, 0, 0, 0, 0]
    assert O.count_peaks(prof) == [3, 11]


def test_diff_of_six_peaks():
    # ProfileOperations.py:340-347: six peaks -> five differences
    prof = np.zeros(128)
    for k in range(6):
        prof[10 + 20 * k] = 50
    assert len(O.count_peaks(prof)) == 6


def test_filter_score():
    # CandidateFileInterface.py:97-112
    assert O.filter_score(13, -1.0) == 0.0
    assert O.filter_score(13, -1e-6) == -1e-6
    assert O.filter_score(14, 12.5) == 12.5
    assert O.filter_score(18, -3.25) == 3.25
    assert O.filter_score(5, -3.25) == -3.25


def test_lyon_column_order_and_constant_input():
    # DataProcessor.py:918-926: mean, stdev, skewness, kurtosis
    x = np.array([2, 8, 0, 4, 1, 9, 9, 0])
    m = O.lyon4(x)
    assert abs(m[0] - 4.125) < 1e-15 and abs(m[2] - 0.2650554122698573) < 1e-12
    c = O.lyon4(np.full(64, 7))
    assert c[0] == 7 and c[1] == 0 and np.isnan(c[2]) and np.isnan(c[3])


def test_row_format_py2():
    # DataProcessor.py:321-325: join, then 'nan' and 'inf' become '0' over the whole line
    row = O.format_row("/data/cand_inf.phcx.gz", [1.0, float("nan"), float("-inf"), 0.1 + 0.2, 1e-5])
    assert row == "/data/cand_0.phcx.gz,1.0,0,-0,0.3,1e-05"
    arff = O.format_row("c", [2.5, float("inf")], arff=True)
    assert arff == "2.5,0,?%c"


def test_histogram_histogram_zero_bins_fails():
    r = O.bates22_phcx(np.full(128, 77), np.ones((16, 128), dtype=np.int64), np.arange(256) % 200,
                       [10, 100, 0.5, 0.05, 0, 300, 0, 0])
    assert r.status & O.ST_HIST_ZERO_BINS
