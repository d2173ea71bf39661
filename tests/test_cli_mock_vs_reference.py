"""CPU suite: the whole C++ host side above the C ABI (cmdstan_b200/host: argument tree -> service shim -> Stan-CSV writer,
adaptation block, metric JSON, per-chain file naming, thinning, save_warmup, progress messages), run END TO END without a
GPU against a test-only mock of libb200nuts.so that serves the ABI from the CPU oracle (tests/host_emul/mock_b200nuts.cpp),
and compared with the files the UNMODIFIED reference writes for the same randomised command lines.  The oracle is bit-exact
for bernoulli, so every byte of every output file must agree except the lines that carry wall-clock times, the start date,
the output file name and the stanc version.  (On the GPU the same host code runs over the real engine: tests/test_gpu_cli.py.)"""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from test_cli_args_vs_reference import CLI, _random_args

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
MOCK_DIR = os.path.join(HERE, "host_emul", "mock")

pytestmark = [pytest.mark.ref, pytest.mark.skipif(not O.have_ref("bernoulli"), reason="oracle/_ref not built")]


@pytest.fixture(scope="module")
def mock_env():
    import cmdstan_b200
    cmdstan_b200.load_library()   # builds the CLI if needed
    O.lib()                       # builds liboracle.so if needed
    os.makedirs(MOCK_DIR, exist_ok=True)
    so = os.path.join(MOCK_DIR, "libb200nuts.so")
    src = os.path.join(HERE, "host_emul", "mock_b200nuts.cpp")
    deps = [src, os.path.join(ROOT, "include", "b200nuts.h"), os.path.join(ROOT, "oracle", "nuts_oracle.h"), os.path.join(ROOT, "oracle", "liboracle.so")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-ffp-contract=off", "-shared", "-o", so, src, "-L" + os.path.join(ROOT, "oracle"),
                               "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    return dict(os.environ, LD_LIBRARY_PATH=MOCK_DIR + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))


def _id(args):
    return next((int(a.split("=")[1]) for a in args if a.startswith("id=")), 1)


def _strip(text):
    keep = []
    for ln in text.splitlines():
        if "seconds (" in ln or ln.startswith("# start_datetime") or ln.startswith("#   file = ") or ln.startswith("# stanc_version") or ln.startswith("# stancflags"):
            continue
        keep.append(ln)
    return keep


@pytest.mark.parametrize("case", range(60))
def test_output_files_identical_to_reference(case, tmp_path, mock_env):
    rng = np.random.default_rng(700 + case)
    dj = tmp_path / "b.json"
    dj.write_text(json.dumps({"N": 10, "y": [0, 1, 0, 0, 0, 0, 0, 0, 0, 1]}))
    args = _random_args(rng, str(dj))
    if "random" not in args:   # a defaulted seed is time based
        k = args.index("output")
        args[k:k] = ["random", f"seed={int(rng.integers(0, 100000))}"]
    nch = int(rng.choice([1, 1, 2, 4]))
    if nch > 1:
        k = args.index("data") if "id=" not in " ".join(args) else next(i for i, a in enumerate(args) if a.startswith("id="))
        args[k:k] = [f"num_chains={nch}"]
    mine_dir, ref_dir = tmp_path / "mine", tmp_path / "ref"
    mine_dir.mkdir()
    ref_dir.mkdir()
    ref = subprocess.run([os.path.join(O.REF, "bernoulli")] + args + ["file=out.csv"], cwd=ref_dir, capture_output=True, text=True)
    mine = subprocess.run([CLI, "bernoulli"] + args + ["file=out.csv"], cwd=mine_dir, capture_output=True, text=True, env=mock_env)
    if "engine=static" in args and nch > 1:
        # the reference has single-chain services only for static HMC and refuses num_chains > 1; the batched engine runs them
        # like any other chains (an extension: chain c equals the single-chain run with id + c)
        assert ref.returncode != 0 and "unavailable for the 'static' HMC engine" in ref.stdout + ref.stderr
        assert mine.returncode == 0 and sorted(f for f in os.listdir(mine_dir) if f.startswith("out_") and f.endswith(".csv")) == [f"out_{k}.csv" for k in range(_id(args), _id(args) + nch)]
        return
    assert ref.returncode == 0, " ".join(args) + "\n" + ref.stdout[-800:] + ref.stderr[-800:]
    assert mine.returncode == 0, " ".join(args) + "\n" + mine.stdout[-800:] + mine.stderr[-800:]
    ref_files, mine_files = sorted(os.listdir(ref_dir)), sorted(os.listdir(mine_dir))
    assert mine_files == ref_files, " ".join(args)
    for f in ref_files:
        a, b = (mine_dir / f).read_text(), (ref_dir / f).read_text()
        if f.endswith("_config.json"):   # output save_cmdstan_config=1 (write_config(structured_writer&)): same text
            skip = lambda t: [ln for ln in t.splitlines() if "start_datetime" not in ln and "stanc_version" not in ln and '"file" : ' not in ln]
            assert skip(a) == skip(b), f + ": " + " ".join(args)
        elif f.endswith(".json"):
            ja, jb = json.loads(a), json.loads(b)
            assert ja.keys() == jb.keys() and ja["metric_type"] == jb["metric_type"], f
            assert np.allclose(np.array(ja["inv_metric"], dtype=float), np.array(jb["inv_metric"], dtype=float), rtol=1e-5) and abs(ja["stepsize"] - jb["stepsize"]) <= 1e-5 * jb["stepsize"], f
        else:
            assert _strip(a) == _strip(b), f + ": " + " ".join(args)
    # progress messages on the console (the reference's chains are threads: their lines interleave, so compare as a multiset)
    it = lambda s: sorted(ln.strip() for ln in s.splitlines() if "Iteration:" in ln)
    assert it(mine.stdout) == it(ref.stdout), " ".join(args)
    if nch == 1:   # one chain: the whole console text, in order
        con = lambda r: [ln for ln in (r.stdout + r.stderr).splitlines()
                         if ln.strip() and not any(t in ln for t in ("seconds", "Gradient evaluation", "leapfrog steps per", "Adjust your expectations", "file = ", "seed = "))]
        assert con(mine) == con(ref), " ".join(args)


@pytest.mark.parametrize("metric,adapt", [("diag_e", 1), ("diag_e", 0), ("dense_e", 1), ("dense_e", 0), ("unit_e", 0)])
def test_user_inits_and_metric_file_identical_to_reference(metric, adapt, tmp_path, mock_env):
    """init=<file> (constrained values -> transform_inits) and metric_file=<file> (read_diag_inv_metric / read_dense_inv_metric),
    with and without adaptation, for each metric: files identical to the reference's."""
    dj = tmp_path / "b.json"
    dj.write_text(json.dumps({"N": 10, "y": [0, 1, 0, 0, 0, 0, 0, 0, 0, 1]}))
    ini = tmp_path / "init.json"
    ini.write_text(json.dumps({"theta": 0.3}))
    args = ["sample", "num_samples=25", "num_warmup=40", "save_warmup=1", "adapt", f"engaged={adapt}", "algorithm=hmc", f"metric={metric}"]
    if metric != "unit_e":
        mf = tmp_path / "metric.json"
        mf.write_text(json.dumps({"inv_metric": [0.7] if metric == "diag_e" else [[0.7]]}))
        args.append(f"metric_file={mf}")
    args += ["stepsize=0.8", "data", f"file={dj}", f"init={ini}", "random", "seed=99", "output", "sig_figs=10"]
    mine_dir, ref_dir = tmp_path / "mine", tmp_path / "ref"
    mine_dir.mkdir()
    ref_dir.mkdir()
    ref = subprocess.run([os.path.join(O.REF, "bernoulli")] + args + ["file=out.csv"], cwd=ref_dir, capture_output=True, text=True)
    mine = subprocess.run([CLI, "bernoulli"] + args + ["file=out.csv"], cwd=mine_dir, capture_output=True, text=True, env=mock_env)
    assert ref.returncode == 0, ref.stdout[-800:] + ref.stderr[-800:]
    assert mine.returncode == 0, mine.stdout[-800:] + mine.stderr[-800:]
    a, b = _strip((mine_dir / "out.csv").read_text()), _strip((ref_dir / "out.csv").read_text())
    drop = lambda ls: [x for x in ls if not x.startswith("#     metric_file = ") and not x.startswith("# init = ")]
    assert drop(a) == drop(b)


def test_per_chain_init_and_metric_files(tmp_path, mock_env):
    """num_chains=2 with init=init.json / metric_file=met.json resolves to init_1.json, init_2.json, met_1.json, met_2.json
    (get_vec_var_context, command_helper.hpp): both chains' files identical to the reference's."""
    dj = tmp_path / "b.json"
    dj.write_text(json.dumps({"N": 10, "y": [0, 1, 0, 0, 0, 0, 0, 0, 0, 1]}))
    outs = {}
    for who in ("mine", "ref"):
        d = tmp_path / who
        d.mkdir()
        for c, (th, im) in enumerate([(0.2, 0.5), (0.6, 1.5)], start=1):
            (d / f"init_{c}.json").write_text(json.dumps({"theta": th}))
            (d / f"met_{c}.json").write_text(json.dumps({"inv_metric": [im]}))
        args = ["sample", "num_samples=8", "num_warmup=30", "algorithm=hmc", "metric_file=met.json", "num_chains=2", "data", f"file={dj}", "init=init.json",
                "random", "seed=3", "output", "sig_figs=9", "file=o.csv"]
        exe, env = ([CLI, "bernoulli"], mock_env) if who == "mine" else ([os.path.join(O.REF, "bernoulli")], None)
        r = subprocess.run(exe + args, cwd=d, capture_output=True, text=True, env=env)
        assert r.returncode == 0, who + r.stdout[-600:] + r.stderr[-600:]
        outs[who] = [_strip((d / f"o_{c}.csv").read_text()) for c in (1, 2)]
    assert outs["mine"] == outs["ref"]


@pytest.mark.parametrize("case", range(8))
def test_std_normal_output_files_identical_to_reference(case, tmp_path, mock_env):
    """a vector parameter (x.1, x.2, x.3): the 3-d standard normal is bit-exact in the oracle for diag_e / unit_e"""
    rng = np.random.default_rng(1200 + case)
    dj = tmp_path / "d.json"
    dj.write_text(json.dumps({"D": 3}))
    args = [a for a in _random_args(rng, str(dj)) if a != "metric=dense_e"]
    if "random" not in args:
        k = args.index("output")
        args[k:k] = ["random", f"seed={int(rng.integers(0, 100000))}"]
    mine_dir, ref_dir = tmp_path / "mine", tmp_path / "ref"
    mine_dir.mkdir()
    ref_dir.mkdir()
    ref = subprocess.run([os.path.join(O.REF, "std_normal")] + args + ["file=out.csv"], cwd=ref_dir, capture_output=True, text=True)
    mine = subprocess.run([CLI, "std_normal"] + args + ["file=out.csv"], cwd=mine_dir, capture_output=True, text=True, env=mock_env)
    assert ref.returncode == 0 and mine.returncode == 0, " ".join(args) + mine.stdout[-500:] + mine.stderr[-500:] + ref.stderr[-500:]
    assert _strip((mine_dir / "out.csv").read_text()) == _strip((ref_dir / "out.csv").read_text()), " ".join(args)


def test_failed_user_initialisation_like_reference(tmp_path, mock_env):
    """theta = 1 has no finite unconstrained image: util::initialize fails before any notice, header row or iteration message
    (hmc_nuts_diag_e_adapt.hpp:360-363) -- same exit code, same "Initialization failed." message, same (header-only) CSV file."""
    dj = tmp_path / "b.json"
    dj.write_text(json.dumps({"N": 10, "y": [0, 1, 0, 0, 0, 0, 0, 0, 0, 1]}))
    ini = tmp_path / "init.json"
    ini.write_text(json.dumps({"theta": 1.0}))
    args = ["sample", "num_samples=8", "num_warmup=30", "data", f"file={dj}", f"init={ini}", "random", "seed=3", "output", "file=o.csv"]
    out = {}
    for who in ("mine", "ref"):
        d = tmp_path / who
        d.mkdir()
        exe, env = ([CLI, "bernoulli"], mock_env) if who == "mine" else ([os.path.join(O.REF, "bernoulli")], None)
        r = subprocess.run(exe + args, cwd=d, capture_output=True, text=True, env=env)
        out[who] = (r.returncode, r.stdout + r.stderr, _strip((d / "o.csv").read_text()))
    assert out["mine"][0] == out["ref"][0] == 1
    assert "Initialization failed." in out["mine"][1] and "Initialization failed." in out["ref"][1]
    assert "Iteration:" not in out["mine"][1] and "WARNING" not in out["mine"][1]
    assert out["mine"][2] == out["ref"][2]


@pytest.mark.skipif(not O.have_ref("eight_schools_nc"), reason="oracle/_ref not built")
def test_eight_schools_transformed_parameters_columns(tmp_path, mock_env):
    """a model with transformed parameters (theta = mu + tau * theta_tilde written by write_array): same comment skeleton and
    header row as the reference, and -- the oracle's closed-form gradient differs from reverse-mode AD in the last ulp, so the
    funnel separates the two chains later -- the first 30 saved rows equal at 8 significant digits."""
    dj = tmp_path / "es.json"
    dj.write_text(json.dumps({"J": 8, "y": [28, 8, -3, 7, -1, 1, 18, 12], "sigma": [15, 10, 16, 11, 9, 11, 10, 18]}))
    args = ["sample", "num_samples=10", "num_warmup=50", "save_warmup=1", "data", f"file={dj}", "random", "seed=3", "output", "sig_figs=8", "file=o.csv"]
    out = {}
    for who in ("mine", "ref"):
        d = tmp_path / who
        d.mkdir()
        exe, env = ([CLI, "eight_schools_nc"], mock_env) if who == "mine" else ([os.path.join(O.REF, "eight_schools_nc")], None)
        r = subprocess.run(exe + args, cwd=d, capture_output=True, text=True, env=env)
        assert r.returncode == 0, who + r.stdout[-600:] + r.stderr[-600:]
        out[who] = _strip((d / "o.csv").read_text())
    a, b = out["mine"], out["ref"]
    assert len(a) == len(b)
    body = lambda ls: [x for x in ls if not x.startswith("#")]
    comments = lambda ls: [x for x in ls if x.startswith("#") and "Step size" not in x and not x[2:3].isdigit() and not x.startswith("# -")]
    assert comments(a) == comments(b)
    assert body(a)[0] == body(b)[0] and body(a)[0].endswith("theta_tilde.8,theta.1,theta.2,theta.3,theta.4,theta.5,theta.6,theta.7,theta.8")
    assert body(a)[1:31] == body(b)[1:31]


def test_chains_split_over_several_engines(tmp_path, mock_env):
    """B200_DEVICES=0,1,2 splits the chains contiguously over three engines (one per GPU in production; chain c of engine g runs
    stream id + first[g] + c): five chains with per-chain init and metric files give the reference's five files whatever the split."""
    dj = tmp_path / "b.json"
    dj.write_text(json.dumps({"N": 10, "y": [0, 1, 0, 0, 0, 0, 0, 0, 0, 1]}))
    args = ["sample", "num_samples=12", "num_warmup=40", "save_warmup=1", "thin=2", "adapt", "save_metric=1", "algorithm=hmc", "metric_file=met.json", "num_chains=5",
            "id=3", "data", f"file={dj}", "init=init.json", "random", "seed=8", "output", "sig_figs=10", "file=o.csv"]
    outs = {}
    for who, devs in (("ref", None), ("one", "0"), ("three", "0,1,2"), ("five", "0,1,2,3,4")):
        d = tmp_path / who
        d.mkdir()
        for c in range(3, 8):
            (d / f"init_{c}.json").write_text(json.dumps({"theta": 0.1 * (c - 1)}))
            (d / f"met_{c}.json").write_text(json.dumps({"inv_metric": [0.3 * c]}))
        if who == "ref":
            r = subprocess.run([os.path.join(O.REF, "bernoulli")] + args, cwd=d, capture_output=True, text=True)
        else:
            r = subprocess.run([CLI, "bernoulli"] + args, cwd=d, capture_output=True, text=True, env=dict(mock_env, B200_DEVICES=devs))
        assert r.returncode == 0, who + r.stdout[-600:] + r.stderr[-600:]
        outs[who] = {f: _strip((d / f).read_text()) for f in sorted(os.listdir(d)) if f.startswith("o_") and f.endswith(".csv")}
        assert sorted(outs[who]) == [f"o_{c}.csv" for c in range(3, 8)]
    assert outs["one"] == outs["ref"] and outs["three"] == outs["ref"] and outs["five"] == outs["ref"]


@pytest.mark.parametrize("fname", ["out", "out.txt", "sub/out.csv", "out.a.csv"])
@pytest.mark.parametrize("nch", [1, 2])
def test_output_file_naming_like_reference(fname, nch, tmp_path, mock_env):
    """file.hpp:92-212 (basename / suffix rules, _<chain id> tags, _metric.json, _config.json, diagnostic files): the same set of
    files as the reference for names without suffix, with another suffix, in a sub-directory, with inner dots."""
    dj = tmp_path / "b.json"
    dj.write_text(json.dumps({"N": 10, "y": [0, 1, 0, 0, 0, 0, 0, 0, 0, 1]}))
    args = ["sample", "num_samples=3", "num_warmup=20", f"num_chains={nch}", "adapt", "save_metric=1", "data", f"file={dj}", "random", "seed=3", "output",
            f"file={fname}", "diagnostic_file=dg", "save_cmdstan_config=1"]
    listing = {}
    for who in ("mine", "ref"):
        d = tmp_path / who
        (d / "sub").mkdir(parents=True)
        exe, env = ([CLI, "bernoulli"], mock_env) if who == "mine" else ([os.path.join(O.REF, "bernoulli")], None)
        r = subprocess.run(exe + args, cwd=d, capture_output=True, text=True, env=env)
        assert r.returncode == 0, who + r.stdout[-500:] + r.stderr[-500:]
        listing[who] = sorted(os.path.relpath(os.path.join(dp, f), d) for dp, _, fs in os.walk(d) for f in fs)
    assert listing["mine"] == listing["ref"]


EDGE = [
    "sample num_samples=0 num_warmup=30 data file=DATA random seed=3 output",
    "sample num_samples=5 num_warmup=0 adapt engaged=0 data file=DATA random seed=3 output",
    "sample num_samples=3 num_warmup=25 thin=7 save_warmup=1 data file=DATA random seed=3 output",
    "sample num_samples=0 num_warmup=0 adapt engaged=0 data file=DATA random seed=3 output",
    "sample num_samples=10 num_warmup=19 data file=DATA random seed=3 output",                       # no variance estimation below 20
    "sample num_samples=10 num_warmup=19 algorithm=hmc metric=dense_e data file=DATA random seed=3 output",   # ... "No covariance estimation" for dense_e
    "sample num_samples=10 num_warmup=40 data file=DATA random seed=3 output",                       # 15 % / 75 % / 10 % fallback windows
    "sample num_samples=10 num_warmup=30 algorithm=hmc engine=nuts max_depth=1 stepsize_jitter=1 data file=DATA init=0 random seed=3 output refresh=0",
    "sample num_samples=4 num_warmup=150 adapt term_buffer=0 init_buffer=0 window=150 data file=DATA random seed=4294967295 output sig_figs=18",
    "sample num_samples=4 num_warmup=30 adapt delta=0.1 data file=DATA random seed=0 id=0 output",
    # the order and spelling cmdstanpy uses: top-level arguments first, then method=sample ...
    "id=2 random seed=123 data file=DATA output refresh=10 sig_figs=6 method=sample num_samples=10 num_warmup=30 save_warmup=0 thin=1 algorithm=hmc engine=nuts max_depth=8 adapt engaged=1 delta=0.85",
]


@pytest.mark.parametrize("line", EDGE, ids=[str(i) for i in range(len(EDGE))])
def test_edge_case_command_lines_identical_to_reference(line, tmp_path, mock_env):
    """degenerate sizes (no draws, no warm-up, thin beyond the draws), window fallbacks and their console warnings, depth 1,
    full jitter, zero inits, the largest seed, id 0: output file AND console messages as the reference's."""
    dj = tmp_path / "b.json"
    dj.write_text(json.dumps({"N": 10, "y": [0, 1, 0, 0, 0, 0, 0, 0, 0, 1]}))
    args = [a.replace("DATA", str(dj)) for a in line.split()]
    k = args.index("output") + 1
    args[k:k] = ["file=o.csv"]
    out = {}
    for who in ("mine", "ref"):
        d = tmp_path / who
        d.mkdir()
        exe, env = ([CLI, "bernoulli"], mock_env) if who == "mine" else ([os.path.join(O.REF, "bernoulli")], None)
        r = subprocess.run(exe + args, cwd=d, capture_output=True, text=True, env=env)
        assert r.returncode == 0, who + r.stdout[-500:] + r.stderr[-500:]
        console = [ln for ln in (r.stdout + r.stderr).splitlines()
                   if ln.strip() and not any(t in ln for t in ("seconds", "Gradient evaluation", "leapfrog steps per", "Adjust your expectations", "file = "))]
        out[who] = (_strip((d / "o.csv").read_text()), console)
    assert out["mine"][0] == out["ref"][0]
    assert out["mine"][1] == out["ref"][1]


@pytest.mark.skipif(not O.have_ref("eight_schools_nc"), reason="oracle/_ref not built")
@pytest.mark.parametrize("init", [{"mu": 1.5}, {"tau": 0.7}, {"theta_tilde": [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]}, {"mu": -2, "tau": 3}],
                         ids=["mu", "tau", "theta_tilde", "mu+tau"])
def test_partial_user_inits_like_reference(init, tmp_path, mock_env):
    """partial inits (initialize.hpp:93-101, chained_var_context): the parameters the user leaves out take the random context's
    values -- drawn for ALL parameters in declaration order, constrained, and mapped back by transform_inits.  Same first rows
    as the reference at 8 significant digits (the funnel separates closed-form and AD gradients later)."""
    dj = tmp_path / "es.json"
    dj.write_text(json.dumps({"J": 8, "y": [28, 8, -3, 7, -1, 1, 18, 12], "sigma": [15, 10, 16, 11, 9, 11, 10, 18]}))
    ini = tmp_path / "i.json"
    ini.write_text(json.dumps(init))
    args = ["sample", "num_samples=5", "num_warmup=30", "save_warmup=1", "data", f"file={dj}", f"init={ini}", "random", "seed=3", "output", "sig_figs=8", "file=o.csv"]
    body = {}
    for who in ("mine", "ref"):
        d = tmp_path / who
        d.mkdir()
        exe, env = ([CLI, "eight_schools_nc"], mock_env) if who == "mine" else ([os.path.join(O.REF, "eight_schools_nc")], None)
        r = subprocess.run(exe + args, cwd=d, capture_output=True, text=True, env=env)
        assert r.returncode == 0, who + r.stdout[-500:] + r.stderr[-500:]
        body[who] = [x for x in (d / "o.csv").read_text().splitlines() if x and not x.startswith("#")]
    assert body["mine"][:12] == body["ref"][:12]


def _read_lp_csv(path):
    lines = [x for x in open(path).read().splitlines() if x and not x.startswith("#")]
    return lines[0], np.array([r.split(",") for r in lines[1:]], dtype=float)


@pytest.mark.skipif(not O.have_ref("eight_schools_nc"), reason="oracle/_ref not built")
@pytest.mark.parametrize("model,kind", [("bernoulli", "u2d"), ("bernoulli", "u1d"), ("bernoulli", "constrained"), ("eight_schools_nc", "u2d"),
                                        ("eight_schools_nc", "constrained")])
def test_log_prob_method_like_reference(model, kind, tmp_path, mock_env):
    """`method=log_prob` (command_helper.hpp:520-600): same configuration echo, header, column names (lp__, g_<name>) and values as
    the reference, for unconstrained sets given as a vector or a matrix (with the reference's column-major / row-major re-reading
    of the matrix) and for one set of constrained values."""
    if model == "bernoulli":
        data, D = {"N": 10, "y": [0, 1, 0, 0, 0, 0, 0, 0, 0, 1]}, 1
        cons = {"theta": 0.3}
    else:
        data, D = {"J": 8, "y": [28, 8, -3, 7, -1, 1, 18, 12], "sigma": [15, 10, 16, 11, 9, 11, 10, 18]}, 10
        cons = {"mu": 1.5, "tau": 2.5, "theta_tilde": [0.1, -0.2, 0.3, -0.4, 0.5, -0.6, 0.7, -0.8]}
    dj = tmp_path / "d.json"
    dj.write_text(json.dumps(data))
    pj = tmp_path / "p.json"
    rng = np.random.default_rng(5)
    if kind == "u2d":
        pj.write_text(json.dumps({"params_r": rng.uniform(-2, 2, size=(4, D)).tolist()}))
        arg = f"unconstrained_params={pj}"
    elif kind == "u1d":
        pj.write_text(json.dumps({"params_r": rng.uniform(-2, 2, size=D).tolist()}))
        arg = f"unconstrained_params={pj}"
    else:
        pj.write_text(json.dumps(cons))
        arg = f"constrained_params={pj}"
    args = ["log_prob", arg, "data", f"file={dj}", "output", "file=lp.csv", "sig_figs=14"]
    res = {}
    for who in ("mine", "ref"):
        d = tmp_path / who
        d.mkdir()
        exe, env = ([CLI, model], mock_env) if who == "mine" else ([os.path.join(O.REF, model)], None)
        r = subprocess.run(exe + args, cwd=d, capture_output=True, text=True, env=env)
        assert r.returncode == 0, who + r.stdout[-500:] + r.stderr[-500:]
        hdr, vals = _read_lp_csv(d / "lp.csv")
        comments = [x for x in (d / "lp.csv").read_text().splitlines() if x.startswith("#") and not any(t in x for t in ("start_datetime", "stanc", "seed =", "file = "))]
        res[who] = (hdr, vals, comments)
    assert res["mine"][0] == res["ref"][0] and res["mine"][2] == res["ref"][2]
    assert res["mine"][1].shape == res["ref"][1].shape
    np.testing.assert_allclose(res["mine"][1], res["ref"][1], rtol=1e-11, atol=1e-12)
