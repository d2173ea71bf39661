"""Posterior diagnostics used to compute the headline metric (min bulk-ESS / s): split rank-normalised ESS and R-hat.

Restatement of the reference's definitions for tensors of draws (so that thousands of chains can be analysed on the
GPU without going through text CSV -> stansummary):

    split_chains                  stan/src/stan/analyze/mcmc/split_chains.hpp:49-67
    rank_transform                stan/src/stan/analyze/mcmc/rank_normalization.hpp:23-57  (average ranks, (r - 3/8)/(S + 1/4))
    ess (Geyer, FFT autocov.)     stan/src/stan/analyze/mcmc/ess.hpp:28-103, autocovariance.hpp
    split_rank_normalized_ess     stan/src/stan/analyze/mcmc/split_rank_normalized_ess.hpp:28-55   (bulk)
    rhat / split_rank_normalized_rhat   stan/src/stan/analyze/mcmc/rhat.hpp, split_rank_normalized_rhat.hpp
    mcse_mean / mcse_sd           stan/src/stan/analyze/mcmc/mcse.hpp:23-63

Checked against the reference's own `stansummary` in tests/test_analyze.py (golden table under tests/golden/).
Works on CPU or CUDA torch tensors; this is measurement code, not part of the sampling hot path.
"""
import math

import torch


def split_chains(x):
    """x [chains, draws, ...] -> [2*chains, draws//2, ...] (head and tail halves; middle draw dropped when odd)."""
    n = x.shape[1]
    half = n // 2
    tail = (n + 1) // 2
    a, b = x[:, :half], x[:, tail:tail + half]
    return torch.stack([a, b], dim=1).reshape((2 * x.shape[0], half) + tuple(x.shape[2:]))


def rank_transform(x):
    """Pooled average ranks -> normal scores.  x [chains, draws] (one parameter)."""
    flat = x.reshape(-1)
    size = flat.numel()
    vals, order = torch.sort(flat)
    _, inverse, counts = torch.unique_consecutive(vals, return_inverse=True, return_counts=True)
    start = torch.cumsum(counts, 0) - counts                      # ranks start at start+1
    avg = start.to(torch.float64) + (counts.to(torch.float64) + 1.0) / 2.0
    ranks_sorted = avg[inverse]
    p = (ranks_sorted - 0.375) / (size + 0.25)
    z_sorted = torch.special.ndtri(p)
    z = torch.empty_like(z_sorted)
    z[order] = z_sorted
    return z.reshape(x.shape)


def autocovariance(x):
    """Biased (1/N) autocovariance along dim 1 via FFT.  x [chains, draws] -> [chains, draws]."""
    n = x.shape[1]
    xc = x - x.mean(dim=1, keepdim=True)
    m = 1
    while m < 2 * n:
        m *= 2
    f = torch.fft.rfft(xc, n=m, dim=1)
    ac = torch.fft.irfft(f * torch.conj(f), n=m, dim=1)[:, :n] / n
    # the reference normalises by ac(0) before scaling by the variance (autocovariance.hpp:52-53): a chain without
    # variance yields NaN at every lag (this is what makes ess() of a one-sided tail indicator come out as S/2)
    const = (xc == 0).all(dim=1)
    return torch.where(const[:, None], torch.full_like(ac, float("nan")), ac)


def ess(x):
    """stan::analyze::ess for one parameter.  x [chains, draws] (float64)."""
    num_chains, n = x.shape
    acov = autocovariance(x)
    chain_mean = x.mean(dim=1)
    chain_var = acov[:, 0] * n / (n - 1)
    w = chain_var.mean()
    var_plus = w * (n - 1) / n
    if num_chains > 1:
        var_plus = var_plus + chain_mean.var(unbiased=True)
    mean_acov = acov.mean(dim=0).cpu().numpy()
    w = float(w)
    var_plus = float(var_plus)
    rho = [0.0] * (n + 2)
    rho_even = 1.0
    rho[0] = rho_even
    rho_odd = 1 - (w - mean_acov[1]) / var_plus
    rho[1] = rho_odd
    t = 1
    while t < n - 4 and (rho_even + rho_odd > 0) and not math.isnan(rho_even + rho_odd):
        rho_even = 1 - (w - mean_acov[t + 1]) / var_plus
        rho_odd = 1 - (w - mean_acov[t + 2]) / var_plus
        if rho_even + rho_odd >= 0:
            rho[t + 1] = rho_even
            rho[t + 2] = rho_odd
        if rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]:
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2
            rho[t + 2] = rho[t + 1]
        t += 2
    max_t = t
    if rho_even > 0:
        rho[max_t + 1] = rho_even
    total = num_chains * n
    tau = -1 + 2 * sum(rho[:max_t]) + rho[max_t + 1]
    tau = max(tau, 1 / math.log10(total))
    return total / tau


def rhat(x):
    """stan::analyze::rhat for one parameter.  x [chains, draws]."""
    n = x.shape[1]
    chain_mean = x.mean(dim=1)
    chain_var = x.var(dim=1, unbiased=True)
    var_between = n * chain_mean.var(unbiased=True)
    var_within = chain_var.mean()
    return float(torch.sqrt((var_between / var_within + n - 1) / n))


def mcse_mean(draws):
    """stan::analyze::mcse_mean (stan/src/stan/analyze/mcmc/mcse.hpp:23-30): sqrt(sample variance / ess(chains)).  draws [chains, n]."""
    x = torch.as_tensor(draws).to(torch.float64)
    if x.shape[1] < 4 or not _ok(x):
        return float("nan")
    return math.sqrt(float(x.reshape(-1).var(unbiased=True)) / ess(x))


def mcse_sd(draws):
    """stan::analyze::mcse_sd (mcse.hpp:41-63): ESS of |x - mean| on split chains, Taylor expansion of sqrt(variance)."""
    x = torch.as_tensor(draws).to(torch.float64)
    if x.shape[1] < 4 or not _ok(x):
        return float("nan")
    ctr = (x - x.mean()).abs()
    ess_mean = ess(split_chains(ctr))
    evar = float((ctr ** 2).mean())
    varvar = (float((ctr ** 4).mean()) - evar ** 2) / ess_mean
    return math.sqrt(varvar / evar / 4.0)


def _approx_const(x, val):
    """Eigen isApproxToConstant(val, 1e-12): |x - val| <= min(|x|, |val|) * 1e-12 for every coefficient."""
    return bool(((x - val).abs() <= torch.minimum(x.abs(), val.abs()) * 1e-12).all())


def _ok(s):
    """is_finite_and_varies (stan/src/stan/analyze/mcmc/check_chains.hpp:27-48): every draw finite, no chain constant,
    and (more than one chain) the chains' first draws not all equal.  s [chains, draws]."""
    if not bool(torch.isfinite(s).all()):
        return False
    for c in range(s.shape[0]):
        if _approx_const(s[c], s[c, 0]):
            return False
    if s.shape[0] > 1 and _approx_const(s[:, 0], s[0, 0]):
        return False
    return True


def ess_bulk(draws):
    """draws [chains, draws] -> split rank-normalised bulk ESS (NaN for constant / non-finite input or < 4 split draws)."""
    s = split_chains(draws.to(torch.float64))
    if s.shape[1] < 4 or not _ok(s):
        return float("nan")
    return ess(rank_transform(s))


def ess_tail(draws):
    s = split_chains(draws.to(torch.float64))
    if s.shape[1] < 4 or not _ok(s):
        return float("nan")
    flat = s.reshape(-1)
    out = []
    for prob, op in ((0.05, torch.le), (0.95, torch.ge)):
        qv = torch.quantile(flat, prob)
        ind = op(s, qv).to(torch.float64)
        out.append(ess(ind))
    out = [v for v in out if not math.isnan(v)]
    return min(out) if out else float("nan")


def split_rhat(draws):
    """max of bulk and tail split rank-normalised R-hat (stansummary's R_hat column)."""
    s = split_chains(draws.to(torch.float64))
    if s.shape[1] < 4 or not _ok(s):
        return float("nan")
    bulk = rhat(rank_transform(s))
    med = torch.quantile(s.reshape(-1), 0.5)   # math::quantile (interpolating), not the lower median
    tail = rhat(rank_transform(torch.abs(s - med)))
    return max(bulk, tail)


def mad(draws):
    """chainset::med_abs_deviation (stan/src/stan/mcmc/chainset.hpp:266-276): 1.4826 * median(|x - median(x)|), pooled."""
    x = torch.as_tensor(draws).to(torch.float64).reshape(-1)
    return 1.4826 * float(torch.quantile((x - torch.quantile(x, 0.5)).abs(), 0.5))


def stansummary_table(draws):
    """draws [chains, rows, ncols] -> [ncols, 10] table in stansummary's column order (Mean, MCSE, StdDev, MAD, 5%, 50%,
    95%, ESS_bulk, ESS_tail, R_hat; src/cmdstan/stansummary_helper.hpp:405-428).  CPU/torch restatement used to check the
    on-device b200nuts_summary; pinned to the compiled reference by tests/golden/ref_stansummary_full.npz."""
    t = torch.as_tensor(draws).to(torch.float64)
    rows = []
    for k in range(t.shape[2]):
        x = t[:, :, k]
        flat = x.reshape(-1)
        q = torch.quantile(flat, torch.tensor([0.05, 0.5, 0.95], dtype=torch.float64))
        rows.append([float(flat.mean()), mcse_mean(x), float(flat.std(unbiased=True)), mad(x), float(q[0]), float(q[1]), float(q[2]),
                     ess_bulk(x), ess_tail(x), split_rhat(x)])
    return torch.tensor(rows, dtype=torch.float64).numpy()


def summarize(draws, columns=None, device=None):
    """draws [chains, draws, K] (numpy or torch) -> dict of per-column tensors: mean, sd, ess_bulk, rhat.
    Columns are processed one at a time on `device` (so an [8192, 1000, 1211] tensor never needs a second copy)."""
    t = torch.as_tensor(draws)
    K = t.shape[2]
    cols = range(K) if columns is None else columns
    res = {"mean": [], "sd": [], "ess_bulk": [], "rhat": []}
    for k in cols:
        x = t[:, :, k].to(device=device or t.device, dtype=torch.float64)
        res["mean"].append(float(x.mean()))
        res["sd"].append(float(x.std(unbiased=True)))
        res["ess_bulk"].append(ess_bulk(x))
        res["rhat"].append(split_rhat(x))
    return res


def min_bulk_ess(draws, device=None):
    """min over lp__ and all model columns (sampler columns 1..6 excluded) of the pooled bulk ESS -- the metric of
    BASELINE.json.  draws [chains, rows, 7 + K]."""
    t = torch.as_tensor(draws)
    cols = [0] + list(range(7, t.shape[2]))
    vals = []
    for k in cols:
        x = t[:, :, k].to(device=device or t.device, dtype=torch.float64)
        vals.append(ess_bulk(x))
    good = [v for v in vals if not math.isnan(v)]
    return (min(good) if good else float("nan")), vals
