"""GPU tests of the C++ host side (cmdstan_b200/host: b200stan = CmdStan's `model sample ...` surface on the B200 service).

The drop-in claim is checked against the UNMODIFIED reference executable (oracle/_ref/<model>, built from
/root/reference by oracle/build_ref.sh and shipped to the GPU box): same command line => same CSV (header, adaptation
block, draws at the printed precision), and the reference's own `stansummary` reads our files."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

import cases
from oracle import oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cmdstan_b200", "bin", "b200stan")


def _cli(model, args, cwd, env=None):
    import cmdstan_b200  # noqa: F401  builds the library + CLI if needed
    cmdstan_b200.load_library()
    r = subprocess.run([CLI, model] + args, cwd=cwd, capture_output=True, text=True, env=env)
    return r


def _bernoulli_json(tmp_path):
    p = tmp_path / "bernoulli.data.json"
    p.write_text(json.dumps({"N": 10, "y": [int(v) for v in cases.BERNOULLI["y"]]}))
    return str(p)


def _strip(lines):
    """header lines that legitimately differ: start time, output file name, compile info, elapsed times"""
    out = []
    for ln in lines:
        if ln.startswith("# start_datetime") or ln.startswith("#   file = ") or ln.startswith("# stanc_version") or "seconds (" in ln:
            continue
        out.append(ln)
    return out


@pytest.mark.skipif(not O.have_ref("bernoulli"), reason="oracle/_ref not built")
def test_cli_csv_identical_to_reference_bernoulli(tmp_path):
    """`bernoulli sample num_chains=4 save_warmup=1 ... seed=1234` through b200stan and through the unmodified reference:
    identical header / comment structure, the first 30 rows of every chain identical as TEXT (default 6 significant
    digits), all rows numerically equal to 1e-3 (device libm differs from glibc in the last ulp, which the dual-averaging
    recursion amplifies to ~2e-5 on the adapted step size after 100 warm-up transitions), same tree shapes throughout."""
    dj = _bernoulli_json(tmp_path)
    args = ["sample", "num_chains=4", "num_samples=100", "num_warmup=100", "save_warmup=1", "data", f"file={dj}", "random", "seed=1234", "output"]
    r = _cli("bernoulli", args + ["file=mine.csv"], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    env = dict(os.environ, STAN_NUM_THREADS="4")
    rr = subprocess.run([os.path.join(O.REF, "bernoulli")] + args + ["file=ref.csv", "num_threads=4"], cwd=tmp_path, capture_output=True, text=True, env=env)
    assert rr.returncode == 0, rr.stdout + rr.stderr
    for c in range(1, 5):
        mine = [ln for ln in _strip((tmp_path / f"mine_{c}.csv").read_text().splitlines()) if not ln.startswith("# num_threads")]
        ref = [ln for ln in _strip((tmp_path / f"ref_{c}.csv").read_text().splitlines()) if not ln.startswith("# num_threads")]
        assert len(mine) == len(ref)
        h = mine.index("lp__,accept_stat__,stepsize__,treedepth__,n_leapfrog__,divergent__,energy__,theta")
        assert mine[:h + 1] == ref[:h + 1]            # version lines, model name, full argument echo, column names
        assert mine[h + 1:h + 31] == ref[h + 1:h + 31]  # first 30 transitions: same text
        rows_m = np.array([ln.split(",") for ln in mine[h + 1:] if not ln.startswith("#")], dtype=float)
        rows_r = np.array([ln.split(",") for ln in ref[h + 1:] if not ln.startswith("#")], dtype=float)
        assert rows_m.shape == rows_r.shape == (200, 8)
        assert np.array_equal(rows_m[:, 3:6], rows_r[:, 3:6]) and np.allclose(rows_m, rows_r, rtol=1e-3)
        cm = [ln for ln in mine[h + 1:] if ln.startswith("#")]
        cr = [ln for ln in ref[h + 1:] if ln.startswith("#")]
        assert [x.split("=")[0] for x in cm if "Step size" not in x][:2] == [x.split("=")[0] for x in cr if "Step size" not in x][:2]
        assert len(cm) == len(cr) and cm[0] == cr[0] == "# Adaptation terminated" and cm[2] == cr[2]
        assert abs(float(cm[1].split("=")[1]) - float(cr[1].split("=")[1])) < 1e-3 and abs(float(cm[3][1:]) - float(cr[3][1:])) < 1e-3
    # console: identical argument echo and progress lines
    echo = lambda t, f: [ln for ln in t.split("\n\n")[0].replace(f, "X").splitlines() if not ln.startswith("num_threads")]
    assert echo(r.stdout, "mine.csv") == echo(rr.stdout, "ref.csv")
    assert "Chain [1] Iteration:   1 / 200 [  0%]  (Warmup)" in r.stdout and "Chain [4] Iteration: 200 / 200 [100%]  (Sampling)" in r.stdout
    # generate_transitions' progress lines (the reference's own stdout interleaves them across its chain threads)
    want = sorted(f"Chain [{c}] Iteration: {it:3d} / 200 [{100 * it // 200:3d}%]  ({ph})" for c in range(1, 5)
                  for it, ph in ((1, "Warmup"), (100, "Warmup"), (101, "Sampling"), (200, "Sampling")))
    assert sorted(ln for ln in r.stdout.splitlines() if ln.startswith("Chain [") and "Iteration" in ln) == want


@pytest.mark.skipif(not O.have_ref("bernoulli"), reason="oracle/_ref not built")
def test_reference_stansummary_reads_our_csv(tmp_path):
    dj = _bernoulli_json(tmp_path)
    r = _cli("bernoulli", ["sample", "num_chains=4", "data", f"file={dj}", "random", "seed=77", "output", "file=o.csv", "refresh=0"], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    files = [str(tmp_path / f"o_{c}.csv") for c in range(1, 5)]
    s = subprocess.run([os.path.join(O.REF, "stansummary")] + files, capture_output=True, text=True)
    assert s.returncode == 0, s.stdout + s.stderr
    row = [ln.split() for ln in s.stdout.splitlines() if ln.startswith("theta")][0]
    mean, rhat, ess_bulk = float(row[1]), float(row[-1]), float(row[-3])
    assert abs(mean - 0.25) < 0.02 and rhat < 1.01 and ess_bulk > 800   # posterior Beta(3, 9): mean 0.25
    # our own ESS restatement agrees with the reference's on the same files
    import torch
    from cmdstan_b200 import analyze as A
    draws = np.stack([O.read_stan_csv(f)[1] for f in files])
    assert abs(A.ess_bulk(torch.tensor(draws[:, :, 7])) - ess_bulk) <= 0.01 * ess_bulk + 1


def test_cli_options_save_warmup_thin_sigfigs_metric(tmp_path):
    d = tmp_path / "es.json"
    d.write_text(json.dumps({"J": 8, "y": cases.EIGHT_SCHOOLS["y"].tolist(), "sigma": cases.EIGHT_SCHOOLS["sigma"].tolist()}))
    r = _cli("eight_schools_nc", ["sample", "num_samples=50", "num_warmup=100", "save_warmup=1", "thin=3", "adapt", "delta=0.9", "save_metric=1",
                                  "algorithm=hmc", "engine=nuts", "max_depth=8", "id=3", "data", f"file={d}", "random", "seed=5", "output",
                                  "file=run.csv", "sig_figs=12", "refresh=25"], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    names, draws, stepsize, metric, timing = O.read_stan_csv(str(tmp_path / "run.csv"))
    assert names[:7] == ["lp__", "accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__"]
    assert names[7:] == ["mu", "tau"] + [f"theta_tilde.{i}" for i in range(1, 9)] + [f"theta.{i}" for i in range(1, 9)]
    assert draws.shape == ((100 + 2) // 3 + (50 + 2) // 3, 25)
    assert (draws[:, 3] <= 8).all() and len(metric) == 10 and set(timing) == {"Warm-up", "Sampling", "Total"}
    assert np.allclose(draws[:, 9 + 8:], draws[:, [7]] + draws[:, [8]] * draws[:, 9:17], rtol=1e-10)  # theta = mu + tau * theta_tilde
    txt = (tmp_path / "run.csv").read_text()
    assert "#       delta = 0.9\n" in txt and "#             max_depth = 8\n" in txt and "# id = 3\n" in txt and "#   sig_figs = 12\n" in txt
    assert re.search(r"^-?\d+\.\d{9,11},", txt.split("energy__")[1].split("\n")[1])  # 12 significant digits
    m = json.loads((tmp_path / "run_metric.json").read_text())
    assert m["metric_type"] == "diag_e" and len(m["inv_metric"]) == 10 and abs(m["stepsize"] - stepsize) < 1e-5 * stepsize
    # matched stream: the same run through the oracle (id=3 => chain id 3)
    ref = O.run_chain(O.Model("eight_schools_nc", cases.EIGHT_SCHOOLS), seed=5, chain_id=3, num_warmup=100, num_samples=50, save_warmup=1, num_thin=3,
                      delta=0.9, max_depth=8)
    assert np.array_equal(draws[:4, 3:6], ref["draws"][:4, 3:6]) and np.allclose(draws[:4, :7], ref["draws"][:4, :7], rtol=1e-6)


def test_cli_init_and_metric_files(tmp_path):
    dj = _bernoulli_json(tmp_path)
    (tmp_path / "init.json").write_text(json.dumps({"theta": 0.3}))
    (tmp_path / "metric.json").write_text(json.dumps({"inv_metric": [0.5]}))
    r = _cli("bernoulli", ["sample", "num_samples=30", "num_warmup=30", "adapt", "engaged=0", "algorithm=hmc", "metric_file=metric.json", "stepsize=0.7",
                           "data", f"file={dj}", "init=init.json", "random", "seed=9", "output", "file=fix.csv", "sig_figs=18"], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    names, draws, stepsize, metric, _ = O.read_stan_csv(str(tmp_path / "fix.csv"))
    # fixed step size and metric; util::run_sampler still writes the sampler-state block (run_sampler.hpp:66-67)
    assert stepsize == 0.7 and metric is not None and metric[0] == 0.5 and (draws[:, 2] == 0.7).all() and draws.shape[0] == 30
    ref = O.run_chain(O.Model("bernoulli", cases.BERNOULLI), q0=[np.log(0.3 / 0.7)], inv_metric=[0.5], seed=9, chain_id=1, num_warmup=30, num_samples=30,
                      adapt_engaged=0, stepsize=0.7)
    m = O.Model("bernoulli", cases.BERNOULLI)
    want = np.concatenate([ref["draws"][:, :7], m.write_array(ref["draws"][:, 7:])], axis=1)
    assert np.array_equal(draws[:, 3:6], want[:, 3:6]) and np.allclose(draws, want, rtol=1e-9)


@pytest.mark.skipif(not O.have_ref("bernoulli"), reason="oracle/_ref not built")
def test_cli_metric_unit_e_vs_reference(tmp_path):
    dj = _bernoulli_json(tmp_path)
    args = ["sample", "num_samples=40", "num_warmup=60", "save_warmup=1", "algorithm=hmc", "metric=unit_e", "data", f"file={dj}", "random", "seed=3",
            "output", "sig_figs=12"]
    r = _cli("bernoulli", args + ["file=u.csv"], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    rr = subprocess.run([os.path.join(O.REF, "bernoulli")] + args + ["file=ur.csv"], cwd=tmp_path, capture_output=True, text=True)
    assert rr.returncode == 0, rr.stdout + rr.stderr
    mine = _strip((tmp_path / "u.csv").read_text().splitlines())
    ref = _strip((tmp_path / "ur.csv").read_text().splitlines())
    assert len(mine) == len(ref) and "# No free parameters for unit metric" in mine and "#         metric = unit_e" in mine
    cm, cr = [x for x in mine if x.startswith("#")], [x for x in ref if x.startswith("#")]
    assert [x for x in cm if "Step size" not in x] == [x for x in cr if "Step size" not in x]
    dm = np.array([x.split(",") for x in mine if not x.startswith("#")][1:], dtype=float)
    dr = np.array([x.split(",") for x in ref if not x.startswith("#")][1:], dtype=float)
    assert np.array_equal(dm[:, 3:6], dr[:, 3:6]) and np.allclose(dm[:30], dr[:30], rtol=1e-9) and np.allclose(dm, dr, rtol=1e-4)


@pytest.mark.skipif(not O.have_ref("eight_schools_nc"), reason="oracle/_ref not built")
def test_cli_metric_dense_e_vs_reference(tmp_path):
    """`model sample algorithm=hmc metric=dense_e` (hmc_nuts_dense_e_adapt): same comment skeleton as the unmodified reference
    (config echo with metric = dense_e, "# Elements of inverse mass matrix:" followed by D rows of D values), first
    transitions equal, save_metric JSON with metric_type dense_e and a D x D inv_metric; then the written metric is fed back
    through metric_file= with adapt engaged=0 (hmc_nuts_dense_e: no adaptation, but the sampler state is still written)."""
    dj = tmp_path / "es.json"
    dj.write_text(json.dumps({"J": 8, "y": [28, 8, -3, 7, -1, 1, 18, 12], "sigma": [15, 10, 16, 11, 9, 11, 10, 18]}))
    args = ["sample", "num_samples=30", "num_warmup=150", "save_warmup=1", "adapt", "save_metric=1", "algorithm=hmc", "metric=dense_e", "data",
            f"file={dj}", "random", "seed=5", "output", "sig_figs=12"]
    r = _cli("eight_schools_nc", args + ["file=d.csv"], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    rr = subprocess.run([os.path.join(O.REF, "eight_schools_nc")] + args + ["file=dr.csv"], cwd=tmp_path, capture_output=True, text=True)
    assert rr.returncode == 0, rr.stdout + rr.stderr
    mine = _strip((tmp_path / "d.csv").read_text().splitlines())
    ref = _strip((tmp_path / "dr.csv").read_text().splitlines())
    assert len(mine) == len(ref) and "#         metric = dense_e" in mine
    i = mine.index("# Elements of inverse mass matrix:")
    assert ref.index("# Elements of inverse mass matrix:") == i
    M = np.array([ln[1:].split(",") for ln in mine[i + 1:i + 11]], dtype=float)
    assert M.shape == (10, 10) and np.allclose(M, M.T, rtol=1e-5) and np.all(np.linalg.eigvalsh(M) > 0)
    dm = np.array([x.split(",") for x in mine if not x.startswith("#")][1:], dtype=float)
    dr = np.array([x.split(",") for x in ref if not x.startswith("#")][1:], dtype=float)
    assert np.array_equal(dm[:10, 3:6], dr[:10, 3:6]) and np.allclose(dm[:10], dr[:10], rtol=1e-7, atol=1e-9)
    mj = json.loads((tmp_path / "d_metric.json").read_text())
    rj = json.loads((tmp_path / "dr_metric.json").read_text())
    assert mj["metric_type"] == rj["metric_type"] == "dense_e" and np.array(mj["inv_metric"]).shape == np.array(rj["inv_metric"]).shape == (10, 10)
    # the adapted metric read back from the JSON file, no adaptation: stepsize__ column constant, no adaptation block
    args2 = ["sample", "num_samples=40", "num_warmup=0", "adapt", "engaged=0", "algorithm=hmc", "metric=dense_e", f"metric_file={tmp_path / 'd_metric.json'}",
             "stepsize=0.3", "data", f"file={dj}", "random", "seed=6", "output", "file=f.csv"]
    r = _cli("eight_schools_nc", args2, tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = (tmp_path / "f.csv").read_text().splitlines()
    rows = np.array([x.split(",") for x in lines if x and not x.startswith("#")][1:], dtype=float)
    # util::run_sampler writes the sampler state too (run_sampler.hpp:66-67): the step size given and the metric read back
    assert rows.shape[0] == 40 and (rows[:, 2] == 0.3).all()
    k = lines.index("# Adaptation terminated")
    assert lines[k + 1] == "# Step size = 0.3" and lines[k + 2] == "# Elements of inverse mass matrix:"


def test_cli_errors(tmp_path):
    dj = _bernoulli_json(tmp_path)
    assert _cli("bernoulli", ["sample", "adapt", "delta=1.5"], tmp_path).returncode == 64
    assert _cli("bernoulli", ["sample", "num_warmup=0", "data", f"file={dj}"], tmp_path).returncode == 64
    assert _cli("bernoulli", ["optimize"], tmp_path).returncode == 64
    bad = tmp_path / "bad.json"
    bad.write_text(json.dumps({"N": 3, "y": [0, 1, 2]}))
    r = _cli("bernoulli", ["sample", "data", f"file={bad}"], tmp_path)
    assert r.returncode != 0 and "must be in the interval [0, 1]" in r.stderr
    r = _cli("bernoulli", ["sample", "data", "file=/nonexistent.json"], tmp_path)
    assert r.returncode != 0


def test_cli_plugin_model(tmp_path):
    """a model that is not built in, reached from the command line through a plug-in library (B200NUTS_MODEL_LIBRARY,
    include/b200nuts_model.h): Stan-CSV columns come from the model's output_name, the reference's stansummary reads the files."""
    from cmdstan_b200 import _build
    rng = np.random.default_rng(3)
    X = rng.normal(size=(400, 3))
    y = 1.0 + X @ np.array([0.5, -1.0, 2.0]) + 0.3 * rng.normal(size=400)
    dj = tmp_path / "linreg.json"
    dj.write_text(json.dumps({"N": 400, "K": 3, "X": X.tolist(), "y": y.tolist()}))
    env = dict(os.environ, B200NUTS_MODEL_LIBRARY=_build.build_example_plugins())
    r = _cli("linreg", ["sample", "num_samples=200", "num_warmup=200", "num_chains=4", "data", f"file={dj}", "random", "seed=11", "output", "file=lr.csv"],
             tmp_path, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    files = [tmp_path / f"lr_{i}.csv" for i in range(1, 5)]
    draws = []
    for f in files:
        lines = f.read_text().splitlines()
        assert "# model = linreg_model" in lines
        body = [x for x in lines if x and not x.startswith("#")]
        assert body[0].split(",") == ["lp__", "accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__", "alpha",
                                      "beta.1", "beta.2", "beta.3", "sigma"]
        draws.append(np.array([x.split(",") for x in body[1:]], dtype=float))
    d = np.stack(draws)
    assert d.shape == (4, 200, 12)
    assert np.all(np.abs(d[:, :, 7:].mean(axis=(0, 1)) - np.array([1.0, 0.5, -1.0, 2.0, 0.3])) < 0.08)
    if O.have_ref("bernoulli"):
        s = subprocess.run([os.path.join(O.REF, "stansummary")] + [str(f) for f in files], capture_output=True, text=True)
        assert s.returncode == 0 and "beta[3]" in s.stdout and "sigma" in s.stdout
    # unknown model name, with and without a library
    r = _cli("no_such_model", ["sample", "data", f"file={dj}"], tmp_path, env=env)
    assert r.returncode != 0 and "unknown model" in (r.stdout + r.stderr)
    # init=<file> for a plug-in model goes through the model's transform_inits (include/b200nuts_model.h, model_base.hpp:348): the run
    # with every parameter given starts where the Python host starts with the hand-computed unconstrained vector (identical files'
    # draws); a partial file works too; a negative scale is refused with lb_free's message
    import cmdstan_b200 as B
    B.load_model_library(_build.build_example_plugins())
    ini = tmp_path / "init.json"
    ini.write_text(json.dumps({"alpha": 0.9, "beta": [0.4, -0.8, 1.5], "sigma": 0.5}))
    args = ["sample", "num_samples=30", "num_warmup=40", "data", f"file={dj}", f"init={ini}", "random", "seed=5", "output", "file=ini.csv", "sig_figs=18"]
    r = _cli("linreg", args, tmp_path, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    body = [x for x in (tmp_path / "ini.csv").read_text().splitlines() if x and not x.startswith("#")]
    cli_draws = np.array([x.split(",") for x in body[1:]], dtype=float)
    rc, res = B.hmc_nuts_diag_e_adapt("linreg", {"X": X, "y": y}, num_chains=1, random_seed=5, num_warmup=40, num_samples=30,
                                      init=np.array([[0.9, 0.4, -0.8, 1.5, np.log(0.5)]]))
    assert rc == 0 and np.allclose(res["draws"][0], cli_draws, rtol=1e-15, atol=0)
    ini.write_text(json.dumps({"sigma": 0.5}))
    r = _cli("linreg", ["sample", "num_samples=5", "num_warmup=20", "data", f"file={dj}", f"init={ini}", "output", "file=p.csv"], tmp_path, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    ini.write_text(json.dumps({"alpha": 0.0, "beta": [0, 0, 0], "sigma": -1.0}))
    r = _cli("linreg", ["sample", "data", f"file={dj}", f"init={ini}"], tmp_path, env=env)
    assert r.returncode != 0 and "lb_free" in (r.stdout + r.stderr)


@pytest.mark.skipif(not O.have_ref("bernoulli"), reason="oracle/_ref not built")
def test_cli_diagnostic_file_vs_reference(tmp_path):
    """`output diagnostic_file=d.csv`: header (stan version, model, configuration echo; no date), column names
    (theta, p_theta, g_theta), rows (q, p, g of every saved draw) and timing footer as the unmodified reference writes them."""
    dj = _bernoulli_json(tmp_path)
    args = ["sample", "num_samples=30", "num_warmup=60", "save_warmup=1", "thin=2", "data", f"file={dj}", "random", "seed=12", "output", "sig_figs=12",
            "diagnostic_file=d.csv"]
    r = _cli("bernoulli", args + ["file=o.csv"], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    (tmp_path / "ref").mkdir()
    rr = subprocess.run([os.path.join(O.REF, "bernoulli")] + args + ["file=o.csv"], cwd=tmp_path / "ref", capture_output=True, text=True)
    assert rr.returncode == 0, rr.stdout + rr.stderr
    mine = _strip((tmp_path / "d.csv").read_text().splitlines())
    ref = _strip((tmp_path / "ref" / "d.csv").read_text().splitlines())
    assert len(mine) == len(ref)
    assert [x for x in mine if x.startswith("#")] == [x for x in ref if x.startswith("#")]
    bm, br = [x for x in mine if not x.startswith("#")], [x for x in ref if not x.startswith("#")]
    assert bm[0] == br[0] == "lp__,accept_stat__,stepsize__,treedepth__,n_leapfrog__,divergent__,energy__,theta,p_theta,g_theta"
    dm, dr = np.array([x.split(",") for x in bm[1:]], dtype=float), np.array([x.split(",") for x in br[1:]], dtype=float)
    assert dm.shape == dr.shape == (45, 10)
    assert np.array_equal(dm[:, 3:6], dr[:, 3:6]) and np.allclose(dm[:20], dr[:20], rtol=1e-9) and np.allclose(dm, dr, rtol=1e-4, atol=1e-6)


@pytest.mark.skipif(not O.have_ref("eight_schools_nc"), reason="oracle/_ref not built")
@pytest.mark.parametrize("jacobian", [1, 0])
def test_cli_log_prob_method_vs_reference(jacobian, tmp_path):
    """`method=log_prob` on the device against the reference's (reverse-mode AD of the real stan::math densities), with and
    without the Jacobian terms: gate 1 (1e-10) through the command line."""
    dj = tmp_path / "es.json"
    dj.write_text(json.dumps({"J": 8, "y": [28, 8, -3, 7, -1, 1, 18, 12], "sigma": [15, 10, 16, 11, 9, 11, 10, 18]}))
    pj = tmp_path / "p.json"
    pj.write_text(json.dumps({"params_r": np.random.default_rng(3).uniform(-2, 2, size=(6, 10)).tolist()}))
    args = ["log_prob", f"unconstrained_params={pj}", f"jacobian={jacobian}", "data", f"file={dj}", "output", "sig_figs=16"]
    r = _cli("eight_schools_nc", args + ["file=mine.csv"], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    rr = subprocess.run([os.path.join(O.REF, "eight_schools_nc")] + args + ["file=ref.csv"], cwd=tmp_path, capture_output=True, text=True)
    assert rr.returncode == 0, rr.stdout + rr.stderr
    rd = lambda f: [x for x in (tmp_path / f).read_text().splitlines() if x and not x.startswith("#")]
    a, b = rd("mine.csv"), rd("ref.csv")
    assert a[0] == b[0] == "lp__," + ",".join(["g_mu", "g_tau"] + [f"g_theta_tilde.{i}" for i in range(1, 9)])
    va, vb = np.array([x.split(",") for x in a[1:]], dtype=float), np.array([x.split(",") for x in b[1:]], dtype=float)
    assert va.shape == vb.shape == (6, 11)
    assert np.max(np.abs(va - vb) / np.maximum(1e-9, np.abs(vb))) < 1e-10
