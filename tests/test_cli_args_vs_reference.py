"""CPU suite: the argument tree of the stand-alone front end (cmdstan_b200/host/arguments.hpp)
against the UNMODIFIED reference's parser (src/cmdstan/arguments/*, command.hpp:84-292), run live where oracle/_ref exists:
for randomised valid command lines both programs must print the same configuration echo (names, nesting, values, the
"(Default)" markers); for invalid ones both must refuse.  b200stan needs no GPU for this: it echoes the configuration before it
creates an engine (and then fails loudly without a device)."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cmdstan_b200", "bin", "b200stan")

pytestmark = [pytest.mark.ref, pytest.mark.skipif(not O.have_ref("bernoulli"), reason="oracle/_ref not built")]


def _echo(stdout):
    """the configuration echo: from 'method = ' to the 'num_threads' line"""
    lines = stdout.splitlines()
    i = next(k for k, ln in enumerate(lines) if ln.startswith("method = "))
    j = next(k for k, ln in enumerate(lines) if ln.startswith("num_threads"))
    out = [ln for ln in lines[i:j + 1] if not ln.strip().startswith("file = ")]
    # a defaulted seed is time based: keep the line's shape ("seed = <number> (Default)"), drop the number
    return ["  seed = <time> (Default)" if ln.startswith("  seed = ") and ln.endswith("(Default)") else ln for ln in out]


def _run(exe, args, cwd):
    return subprocess.run(exe + args, cwd=cwd, capture_output=True, text=True)


def _random_args(rng, dj):
    a = ["sample"]
    if rng.uniform() < 0.7:
        a.append(f"num_samples={int(rng.integers(1, 50))}")
    if rng.uniform() < 0.7:
        a.append(f"num_warmup={int(rng.integers(20, 60))}")
    if rng.uniform() < 0.4:
        a.append(f"save_warmup={int(rng.integers(0, 2))}")
    if rng.uniform() < 0.4:
        a.append(f"thin={int(rng.integers(1, 4))}")
    if rng.uniform() < 0.6:
        a.append("adapt")
        for name, vals in (("gamma", [0.05, 0.1]), ("delta", [0.8, 0.9, 0.65]), ("kappa", [0.75, 0.6]), ("t0", [10, 3.5]),
                           ("init_buffer", [75, 10]), ("term_buffer", [50, 5]), ("window", [25, 8]), ("save_metric", [0, 1])):
            if rng.uniform() < 0.35:
                a.append(f"{name}={rng.choice(vals)}")
    if rng.uniform() < 0.7:
        a.append("algorithm=hmc")
        u = rng.uniform()
        if u < 0.5:
            a += ["engine=nuts", f"max_depth={int(rng.integers(1, 12))}"]
        elif u < 0.7:
            a += ["engine=static", f"int_time={rng.choice([0.4, 2.5, 6.28319])}"]
        if rng.uniform() < 0.5:
            a.append(f"metric={rng.choice(['diag_e', 'unit_e', 'dense_e'])}")
        if rng.uniform() < 0.4:
            a.append(f"stepsize={rng.choice([1, 0.5, 0.03])}")
        if rng.uniform() < 0.3:
            a.append(f"stepsize_jitter={rng.choice([0, 0.2])}")
    if rng.uniform() < 0.3:
        a.append(f"id={int(rng.integers(1, 9))}")
    a += ["data", f"file={dj}"]
    if rng.uniform() < 0.3:
        a.append(f"init={rng.choice([0, 0.5, 2])}")
    if rng.uniform() < 0.7:
        a += ["random", f"seed={int(rng.integers(0, 100000))}"]
    a.append("output")
    if rng.uniform() < 0.4:
        a.append(f"refresh={int(rng.choice([100, 7]))}")
    if rng.uniform() < 0.4:
        a.append(f"sig_figs={int(rng.choice([-1, 4, 12]))}")
    if rng.uniform() < 0.3:
        a.append(f"save_cmdstan_config={int(rng.integers(0, 2))}")
    if rng.uniform() < 0.3 and "engine=static" not in a:   # the B200 engine keeps momenta for the NUTS engine only
        a.append("diagnostic_file=diag.csv")
    return a


@pytest.mark.parametrize("case", range(25))
def test_config_echo_matches_reference(case, tmp_path):
    import cmdstan_b200
    cmdstan_b200.load_library()   # builds the CLI if needed
    dj = tmp_path / "b.json"
    dj.write_text(json.dumps({"N": 10, "y": [0, 1, 0, 0, 0, 0, 0, 0, 0, 1]}))
    args = _random_args(np.random.default_rng(300 + case), str(dj))
    ref = _run([os.path.join(O.REF, "bernoulli")], args + ["file=ref.csv"], tmp_path)
    mine = _run([CLI, "bernoulli"], args + ["file=mine.csv"], tmp_path)
    assert ref.returncode == 0, " ".join(args) + "\n" + ref.stdout[-800:] + ref.stderr[-800:]
    assert _echo(mine.stdout) == _echo(ref.stdout), " ".join(args)


BAD = [["sample", "adapt", "delta=1.5"], ["sample", "num_samples=-3"], ["sample", "thin=0"], ["sample", "algorithm=hmc", "engine=nuts", "max_depth=0"],
       ["sample", "algorithm=hmc", "stepsize=-1"], ["sample", "algorithm=hmc", "stepsize_jitter=1.5"], ["sample", "bogus=1"], ["sample", "adapt", "gamma=0"],
       ["sample", "algorithm=hmc", "metric=banana"], ["random", "seed=-7"], ["sample", "num_warmup=abc"]]


@pytest.mark.parametrize("args", BAD, ids=[" ".join(b) for b in BAD])
def test_invalid_arguments_refused_like_reference(args, tmp_path):
    import cmdstan_b200
    cmdstan_b200.load_library()
    dj = tmp_path / "b.json"
    dj.write_text(json.dumps({"N": 10, "y": [0, 1, 0, 0, 0, 0, 0, 0, 0, 1]}))
    extra = ["data", f"file={dj}"] if args[0] == "sample" else []
    ref = _run([os.path.join(O.REF, "bernoulli")], args + extra, tmp_path)
    mine = _run([CLI, "bernoulli"], args + extra, tmp_path)
    assert ref.returncode != 0 and mine.returncode != 0
    assert "method = " not in mine.stdout   # refused before any configuration echo, as the reference does


DATA_CASES = {"out_of_range": {"N": 10, "y": [0, 1, 0, 0, 0, 0, 0, 0, 0, 2]}, "wrong_size": {"N": 10, "y": [0, 1, 0]}, "missing": {"N": 10},
              "real_for_int": {"N": 3, "y": [0, 1.5, 0]}, "broken_json": '{"N": 3, "y": [0,1,', "no_file": None}


@pytest.mark.parametrize("kind", list(DATA_CASES))
def test_bad_data_refused_like_reference(kind, tmp_path):
    """data errors surface before any engine exists (new_model / var_context, command.hpp:170-171): both programs exit non-zero"""
    import cmdstan_b200
    cmdstan_b200.load_library()
    dj = tmp_path / "d.json"
    v = DATA_CASES[kind]
    if isinstance(v, dict):
        dj.write_text(json.dumps(v))
    elif isinstance(v, str):
        dj.write_text(v)
    args = ["sample", "data", f"file={dj}", "output", "file=o.csv"]
    ref = _run([os.path.join(O.REF, "bernoulli")], args, tmp_path)
    mine = _run([CLI, "bernoulli"], args, tmp_path)
    assert ref.returncode != 0 and mine.returncode != 0
    assert "no CUDA device" not in mine.stderr + mine.stdout   # refused for the data, not for the missing GPU


def test_help_and_help_all_like_reference(tmp_path):
    """`help` / `help-all` exit 0 in both programs; every argument of the B200 front end's tree appears in the reference's
    help-all with the same name, in the same order (the reference's tree also holds the other methods)."""
    import re
    import cmdstan_b200
    cmdstan_b200.load_library()
    for h in ("help", "help-all"):
        assert _run([os.path.join(O.REF, "bernoulli")], [h], tmp_path).returncode == 0
        assert _run([CLI, "bernoulli"], [h], tmp_path).returncode == 0
    # argument names in tree order (the reference prints its own prose for descriptions and for some defaults)
    names = lambda out: [ln.strip().split("=<")[0] for ln in out.splitlines() if re.match(r"^\s*\S+=<.*>$", ln)]
    ref = names(_run([os.path.join(O.REF, "bernoulli")], ["help-all"], tmp_path).stdout)
    mine_out = _run([CLI, "bernoulli"], ["help-all"], tmp_path).stdout
    mine = names(mine_out)
    assert len(mine) >= 30
    it = iter(ref)
    missing = [n for n in mine if not any(n == r for r in it)]   # subsequence test
    assert not missing, missing[:5]
    for frag in ("num_samples=<int>", "Defaults to 1000", "delta=<double>", "Valid values: 0.1 <= delta <= 1", "Defaults to 0.8", "max_depth=<int>", "Defaults to 10",
                 "Valid values: unit_e, diag_e, dense_e", "Defaults to diag_e"):
        assert frag in mine_out
    assert _run([os.path.join(O.REF, "bernoulli")], [], tmp_path).returncode != 0 and _run([CLI, "bernoulli"], [], tmp_path).returncode != 0
