"""GPU suite, SURVEY section 8 row f3: the on-device posterior summary (csrc/summary.cu, b200nuts_summary /
b200nuts_summarize_draws) is the kernel the headline ESS of bench.py comes from, so it is pinned on hardware to

  * the unmodified reference's own `stansummary -s 12` table (C/src/cmdstan/stansummary_helper.hpp:397-428,
    S/analyze/mcmc/ess.hpp:28-103, split_rank_normalized_ess.hpp:28-55, rhat.hpp, mcse.hpp, chainset.hpp) through the
    committed goldens tests/golden/ref_stansummary_full.npz (generator: tests/golden/gen_golden.py), every cell at
    rtol 2e-9 -- even and odd draw counts, constant / discrete / tied columns;
  * the torch restatement of the metric (cmdstan_b200/analyze.py, itself pinned to the same goldens and to the live
    reference on the CPU suite) on draws of a c3-shaped run that never left HBM, and on thousands of pooled chains.
"""
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

FULL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_stansummary_full.npz")
RTOL = 2e-9   # the golden table is printed with 12 significant digits


@pytest.mark.parametrize("tag", ["bernoulli_even", "eight_schools_odd"])
def test_gpu_summary_vs_stansummary_golden(tag):
    import cmdstan_b200 as B
    z = np.load(FULL)
    ref = z[tag + "_table"]
    got = B.summarize_draws(z[tag + "_draws"])
    assert list(z[tag + "_stat_names"]) == ["Mean", "MCSE", "StdDev", "MAD", "5%", "50%", "95%", "ESS_bulk", "ESS_tail", "R_hat"]
    assert list(B.SUMMARY_COLUMNS) == ["Mean", "MCSE", "StdDev", "MAD", "5%", "50%", "95%", "ESS_bulk", "ESS_tail", "R_hat"]
    assert got.shape == ref.shape
    assert np.array_equal(np.isnan(got), np.isnan(ref)), tag
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-12, equal_nan=True, err_msg=tag)


def test_engine_summary_equals_restatement_on_c3_shaped_run():
    """e.summary() (draws read in place from HBM, write_array applied on the fly) == the pinned restatement applied to
    fetch_draws() of the same rows, for every column of a logistic run with c3's D = 101 (N reduced so the test takes
    seconds); and the bench's reduction of the table (min over lp__ and the model columns of ESS_bulk) == analyze.min_bulk_ess."""
    import torch
    import cmdstan_b200 as B
    from cmdstan_b200 import analyze as A
    data = cases.logistic_data(4000, 100)
    e = B.Engine("logistic", data, num_chains=96, random_seed=11, num_warmup=120, num_samples=75)   # odd draw count: split drops the middle draw
    e.init_chains()
    e.run(120 + 75)
    tab = e.summary(row0=0, nrows=75, constrained=True)
    draws = e.fetch_draws(row0=0, nrows=75, constrained=True)
    e.close()
    ref = A.stansummary_table(draws)
    assert tab.shape == ref.shape == (7 + 101, 10)
    assert np.array_equal(np.isnan(tab), np.isnan(ref))
    np.testing.assert_allclose(tab, ref, rtol=RTOL, atol=1e-12, equal_nan=True)
    k = B.SUMMARY_COLUMNS.index("ESS_bulk")
    cols = [0] + list(range(7, tab.shape[0]))
    mess, _ = A.min_bulk_ess(torch.tensor(draws))
    assert abs(np.nanmin(tab[cols, k]) - mess) <= RTOL * mess


def test_engine_summary_of_a_row_window():
    """a window in the middle of the stored rows (what bench.py asks for: row0 = W)"""
    import cmdstan_b200 as B
    from cmdstan_b200 import analyze as A
    data = cases.logistic_data(3000, 20)
    e = B.Engine("logistic", data, num_chains=32, random_seed=12, num_warmup=60, num_samples=50)
    e.init_chains()
    e.run(110)
    tab = e.summary(row0=7, nrows=40, constrained=True)
    ref = A.stansummary_table(e.fetch_draws(row0=7, nrows=40, constrained=True))
    e.close()
    np.testing.assert_allclose(tab, ref, rtol=RTOL, atol=1e-12, equal_nan=True)


def test_summary_many_chains_pooled():
    """thousands of pooled chains (the regime of the benchmark: 4096 chains x 50 draws of eight schools): same table as the
    restatement; ESS_bulk of the pooled draws grows with the number of chains."""
    import cmdstan_b200 as B
    from cmdstan_b200 import analyze as A
    e = B.Engine("eight_schools_nc", cases.EIGHT_SCHOOLS, num_chains=4096, random_seed=5, num_warmup=200, num_samples=50)
    e.init_chains()
    e.run(250)
    tab = e.summary(constrained=True)
    draws = e.fetch_draws(constrained=True)
    e.close()
    ref = A.stansummary_table(draws)
    np.testing.assert_allclose(tab, ref, rtol=RTOL, atol=1e-12, equal_nan=True)
    k = B.SUMMARY_COLUMNS.index("ESS_bulk")
    assert np.nanmin(tab[[0] + list(range(7, tab.shape[0])), k]) > 20000
