#!/usr/bin/env python
"""bench.py -- headline benchmark of the NUTS (diag_e, adaptive) sampling path.

Metric (BASELINE.json): min bulk-ESS / s (and gradient evaluations / s) of the SAMPLING phase, whole job over all GPUs.

  step        = one NUTS transition of every chain (after the reference's windowed adaptation has finished)
  workload    = BASELINE.json configs: c3 (default; the config the north-star target is quoted on) logistic regression
                N=100k, D=100, 1024 chains | c2 eight schools 4096 chains | c1 bernoulli 4 chains | c4 IRT | c5 poisson
  value       = min over (lp__, every model column) of the split rank-normalised bulk ESS of the K timed transitions,
                summed over ranks, / device time of those K transitions (max over ranks), inputs resident in HBM
  e2e         = same ESS / (host->device upload of the data set + wall time of the K transitions + device->host fetch of their
                draws), all through the C ABI with pinned host buffers; the two copy legs are timed on an engine whose device
                buffers exist (the copies, not cudaMalloc / stream creation) and after one untimed fetch of the same shape
  roofline    = dominant kernel family (gradient stage) timed live with CUDA events on the engine's stream; for the logistic
                link the INT8 digit-plane kernel on tcgen05 serves: `achieved` counts the algorithmic fp64 flops against the
                DGEMM rate measured in the run (a fraction above one is possible), `int8_*` what was issued to the tensor cores
  cpu_baseline= the UNMODIFIED reference (oracle/_ref/<model>, CmdStan 2.36 compiled from /root/reference) on the host
                cores, chains in parallel, on a bounded sample of the same workload
  --impl reference  times only that CPU arm and prints the same JSON line shape.

Launch: python bench.py --gpus N --steps K --warmup W     (N > 1: torchrun, one rank per GPU; chains shard, no collectives)
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import warnings

import numpy as np

warnings.filterwarnings("ignore")   # stdout carries exactly one JSON line

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    "c1": dict(model="bernoulli", chains=4, desc="bernoulli N=10, 4 chains"),
    "c2": dict(model="eight_schools_nc", chains=4096, desc="eight-schools non-centred, 4096 chains"),
    "c3": dict(model="logistic", chains=1024, N=100000, K=100, desc="logistic regression N=100k D=100, 1024 chains"),
    # c4 / c5 defaults are sized so that the line finishes in ~2 minutes on one B200 (profiles r01h / r01c)
    "c4": dict(model="irt_2pl", chains=8192, I=1000, J=100, desc="IRT 2PL 1000x100, 8192 chains", num_warmup=150, steps=30),
    "c5": dict(model="poisson_glm", chains=64, N=10000000, K=50, desc="poisson GLM N=10M D=50, 64 chains (rows sharded when N>1)", num_warmup=150, steps=30),
}


def make_data(w, scale=1.0):
    import cases
    m = w["model"]
    if m == "bernoulli":
        return cases.BERNOULLI
    if m == "eight_schools_nc":
        return cases.EIGHT_SCHOOLS
    if m == "logistic":
        return cases.logistic_data(int(w["N"] * scale), w["K"])
    if m == "poisson_glm":
        return cases.poisson_data(int(w["N"] * scale), w["K"])
    if m == "irt_2pl":
        return cases.irt_data(int(w["I"] * scale), w["J"])
    raise ValueError(m)


def data_bytes(data):
    return int(sum(np.asarray(v).nbytes for v in data.values()))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu=0):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([s.strip() for s in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def _pinned(shape, dtype=np.float64):
    """page-locked host buffer (numpy view of a pinned torch tensor): what the e2e legs copy from / to"""
    import torch
    tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32, np.dtype(np.int8): torch.int8, np.dtype(np.int64): torch.int64}[np.dtype(dtype)]
    return torch.empty(tuple(int(x) for x in shape), dtype=tdt, pin_memory=True).numpy()


def _pinned_copy(v):
    a = np.ascontiguousarray(v)
    if a.dtype not in (np.float64, np.int32, np.int8, np.int64) or a.ndim == 0:
        return v
    out = _pinned(a.shape, a.dtype)
    out[...] = a
    return out


def ncu_traffic(model):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/ncu_traffic.json);
    None when this workload has no capture."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(model)
        if not t:
            return None
        out = {"bytes": t["read_bytes"] + t["write_bytes"], "kernel": t["kernel"], "chains_in_launch": t["chains_in_launch"], "source": t["capture"]}
        for k in ("note", "cold_cache_bytes", "cold_cache_capture"):
            if k in t:
                out[k] = t[k]
        return out
    except Exception:
        return None


def algorithmic_work(w, data, model, D):
    """Per chain-gradient algorithmic flops and bytes (restated in DESIGN.md)."""
    if model in ("logistic", "poisson_glm"):
        N, K = data["X"].shape
        Dp = (K + 1 + 7) // 8 * 8
        return dict(flops=4.0 * N * Dp, bytes_per_batch=N * Dp * 8.0 + N * 4.0, kind="tensor")
    if model == "irt_2pl":
        I, J = data["y"].shape
        return dict(flops=25.0 * I * J, bytes_per_batch=None, bytes_per_chain=2 * D * 8.0, kind="fp64")
    return dict(flops={"bernoulli": 40.0, "eight_schools_nc": 150.0}.get(model, 4.0 * D), bytes_per_chain=2 * D * 8.0, kind="latency")


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference on the host cores
# ------------------------------------------------------------------------------------------------------------------
def run_reference(wname, w, data, cores, warm_state=None, budget_s=60.0, scale_note="", max_draws=400):
    """Bounded sample of the workload through oracle/_ref/<model> (CmdStan 2.36 `sample`), chains in parallel.
    Small models run the full adaptive path (long enough for the 1 ms timer); the GLMs are warm-started from an
    adapted state (stepsize, inverse metric, positions) with adapt engaged=0 so that only the sampling phase is timed."""
    from oracle import oracle as O
    from cmdstan_b200 import analyze as A
    import torch
    model = w["model"]
    if not O.have_ref(model):
        return None
    td = tempfile.mkdtemp(prefix="b200bench_")
    chains = 4 if wname == "c1" else cores
    if model in ("bernoulli", "eight_schools_nc", "std_normal"):
        ns = 20000
        res, wall, _ = O.ref_sample(model, data, num_chains=chains, seed=4321, num_warmup=1000, num_samples=ns, sig_figs=8, num_threads=cores, tmpdir=td)
        sample = f"{chains} chains x (1000 warmup + {ns} draws), full adaptive run, num_threads={cores}"
    else:
        # warm start (adaptation on this CPU would take hours for c3): per-chain inits + metric, common stepsize
        assert warm_state is not None
        D = warm_state["q"].shape[1]
        names = {"logistic": ["alpha", "beta"], "poisson_glm": ["alpha", "beta"]}.get(model)
        for c in range(chains):
            q = warm_state["q"][c % len(warm_state["q"])]
            im = warm_state["inv_metric"][c % len(warm_state["inv_metric"])]
            if model == "irt_2pl":
                I, J = data["y"].shape
                init = {"sigma_theta": float(np.exp(q[0])), "theta": q[1:1 + I].tolist(), "sigma_a": float(np.exp(q[1 + I])),
                        "a": np.exp(q[2 + I:2 + I + J]).tolist(), "mu_b": float(q[2 + I + J]), "sigma_b": float(np.exp(q[3 + I + J])),
                        "b": q[4 + I + J:].tolist()}
            else:
                init = {"alpha": float(q[0]), "beta": q[1:].tolist()}
            json.dump(init, open(os.path.join(td, f"init_{c + 1}.json"), "w"))
            json.dump({"inv_metric": im.tolist()}, open(os.path.join(td, f"metric_{c + 1}.json"), "w"))
        eps = float(np.median(warm_state["stepsize"]))
        # size the sample to the time budget with a 3-transition probe
        t0 = time.time()
        O.ref_sample(model, data, num_chains=chains, seed=4321, num_warmup=0, num_samples=3, sig_figs=8, num_threads=cores, tmpdir=td,
                     adapt_engaged=False, stepsize=eps, metric_file=os.path.join(td, "metric.json"), init=os.path.join(td, "init.json"))
        probe = time.time() - t0
        res_p, _, _ = O.ref_sample(model, data, num_chains=chains, seed=4321, num_warmup=0, num_samples=1, sig_figs=8, num_threads=cores, tmpdir=td,
                                   adapt_engaged=False, stepsize=eps, metric_file=os.path.join(td, "metric.json"), init=os.path.join(td, "init.json"))
        ns = int(max(8, min(max_draws, budget_s / max(probe / 3.0, 1e-3) * 0.8)))
        res, wall, _ = O.ref_sample(model, data, num_chains=chains, seed=4321, num_warmup=0, num_samples=ns, sig_figs=8, num_threads=cores, tmpdir=td,
                                    adapt_engaged=False, stepsize=eps, metric_file=os.path.join(td, "metric.json"), init=os.path.join(td, "init.json"))
        sample = f"{chains} chains x {ns} sampling transitions warm-started from the adapted state (adapt engaged=0, stepsize={eps:.4g}), num_threads={cores}{scale_note}"
    draws = np.stack([r[1] for r in res])
    samp_s = max(r[4].get("Sampling", 0.0) for r in res)
    if samp_s <= 0:
        samp_s = wall
    mess, _ = A.min_bulk_ess(torch.tensor(draws))
    nleap = float(draws[:, :, 4].sum())
    return dict(value=mess / samp_s, unit="min_bulk_ESS/s", cores=cores, kind="reference", sample=sample, sampling_s=samp_s, wall_s=wall,
                grad_evals_per_s=nleap / samp_s, min_bulk_ess=mess, draws=int(draws.shape[0] * draws.shape[1]), draws_per_chain=int(draws.shape[1]))


ESS_WINDOW = 200   # draws per chain the ESS-per-draw factor of `value` is estimated over (the K timed ones are its last K)


def common_config(wname, w, num_warmup, K, world, scale_note, chains_total, ess_window):
    """`config` of the JSON line: the SAME dict for the GPU arm and the reference arm (the driver compares them)."""
    model = w["model"]
    mode = "none" if world == 1 else ("rows + ncclAllReduce per leapfrog" if model == "poisson_glm" else "chains")
    return {"workload": wname + ": " + w["desc"] + scale_note, "num_warmup": num_warmup, "timed_transitions": K,
            "chains": chains_total, "ess_window_draws": ess_window,
            "inputs": "synthetic data of the named shape (SURVEY 8d seeds); c3/c5 data sets exceed L2 (80 MB / 4 GB), chain state resident in HBM",
            "sharding": mode,
            "gradient_count": "grad evals = sum of n_leapfrog__ (both arms); the reference's extra gradient at the start of every "
                              "transition is not executed by the GPU engine (cached) and not counted for either arm"}


class LineGuard:
    """The one JSON line must be printed whatever the optional legs (extra workloads, CPU baseline) do: after `limit_s` the
    guard prints the line as it stands and ends the process."""

    def __init__(self, limit_s):
        self.line, self.done, self.lock = None, False, threading.Lock()
        self.timer = threading.Timer(limit_s, self._expire)
        self.timer.daemon = True

    def arm(self, line):
        self.line = line
        self.timer.start()

    def _expire(self):
        with self.lock:
            if self.done or self.line is None:
                return
            self.done = True
            self.line.setdefault("notes", []).append("optional legs cut off by the time guard")
            print(json.dumps(self.line), flush=True)
        os._exit(0)

    def finish(self):
        with self.lock:
            if self.done:
                return
            self.done = True
            self.timer.cancel()
            print(json.dumps(self.line), flush=True)


def run_gpu_workload(wname, K, W, num_warmup, chains_per_gpu, scale, rank, world, local_rank, note, ess_window, primary=True,
                     strong_chains=False, dump_warm=""):
    """One workload on this process' GPU (all ranks call it together).  Returns (result dict on rank 0 else None, engine-side
    leftovers for the CPU leg).  K transitions are timed; the ESS-per-draw factor comes from the last `ess_window` draws."""
    import torch
    import torch.distributed as dist
    import cmdstan_b200 as B
    from cmdstan_b200 import sharding as S
    w = WORKLOADS[wname]
    model = w["model"]
    dev = torch.device("cuda", local_rank)
    data = make_data(w, scale)
    row_mode = model == "poisson_glm" and world > 1   # c5: rows data-sharded, one NCCL all-reduce per leapfrog (SURVEY 8e mode 2)
    C = chains_per_gpu
    ess_window = max(K, ess_window)
    Wx = max(W, ess_window - K)                         # untimed sampling transitions before the timed K
    num_samples = Wx + K

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_setup = time.time()
    if row_mode:
        e = S.row_sharded_engine(model, data, rank, world, transport=os.environ.get("B200_BENCH_TRANSPORT", "nccl"), device=local_rank,
                                 num_chains=C, random_seed=20241017, init_chain_id=1, num_warmup=num_warmup, num_samples=num_samples)
    else:   # chains shard: rank r runs stream ids 1 + r*C ... (no data-path collective)
        e = B.Engine(model, data, num_chains=C, random_seed=20241017, init_chain_id=1 + rank * C, num_warmup=num_warmup,
                     num_samples=num_samples, device=local_rank)
    e.init_chains()
    note(f"{wname}: engine ready, data uploaded")
    if os.environ.get("B200_BENCH_VERBOSE") and num_warmup > 20:   # adaptation in pieces, to see where the time goes
        done, adapt_s = 0, 0.0
        while done < num_warmup:
            n = min(max(10, num_warmup // 10), num_warmup - done)
            e.run(n)
            s_, l_, g_ = e.timing()
            adapt_s, done = adapt_s + s_, done + n
            note(f"{wname}: warm-up {done}/{num_warmup}: last {n} transitions {s_:.2f} s device, {l_} launches, gradient stage {g_:.2f} s")
    else:
        e.run(num_warmup)
        adapt_s, _, _ = e.timing()
    e.run(Wx)
    _, _, ng1 = e.chain_status()
    setup_wall = time.time() - t_setup
    # ---- timed region: K transitions of every chain
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    t0 = time.time()
    e.run(K)
    torch.cuda.synchronize()
    run_wall = time.time() - t0
    barrier()
    clk = clocks.stop()
    run_s, launches, grad_s = e.timing()
    note(f"{wname}: timed region: {run_s:.3f} s device, {launches} launches, gradient stage {grad_s:.3f} s")
    err, it, ng2 = e.chain_status()
    assert (err == 0).all(), f"chains stopped with errors: {np.unique(err)}"
    # ---- D2H of the timed draws through the C ABI (part of e2e), into pinned host memory allocated outside the timed leg
    draws = _pinned((C, K, 7 + e.K))
    e.fetch_draws(row0=Wx - K, nrows=K, constrained=True, out=draws) if Wx >= K else None   # untimed warm-up of the path (an earlier window
                                                                                          # of the same shape: staging buffer, first launch)
    t0 = time.time()
    e.fetch_draws(row0=Wx, nrows=K, constrained=True, out=draws)
    fetch_s = time.time() - t0
    # ---- H2D leg of e2e: a fresh engine uploads the same data set from pinned host memory
    # (the engine and its device buffers exist already, as they do in a process that samples more than once: the leg times the copies
    # of b200nuts_set_data_*, not cudaMalloc / stream creation, which cost 10 - 60 ms of host time on their own on the bench boxes)
    host_data = {k: _pinned_copy(v) for k, v in (S.shard_rows(data, rank, world) if row_mode else data).items()}
    e2 = B.Engine(model, host_data, num_chains=1, device=local_rank)
    torch.cuda.synchronize()
    t0 = time.time()
    e2._set_data(host_data)
    torch.cuda.synchronize()
    h2d_s = time.time() - t0
    e2.close()
    eps, im = e.adaptation()
    ar_us, ar_bytes = (e.measure_allreduce(200) if row_mode else (None, None))

    # stansummary's table computed on the device from the HBM-resident draws (b200nuts_summary, pinned to the reference's
    # stansummary goldens by tests/test_gpu_summary.py): bulk ESS and R-hat of lp__ and every model column over the ESS window,
    # pooled over this rank's chains
    r0 = Wx + K - ess_window
    tab = e.summary(row0=r0, nrows=ess_window, constrained=True)
    cols = [0] + list(range(7, tab.shape[0]))
    ess_cols = torch.tensor([v if v == v else float("inf") for v in tab[cols, B.SUMMARY_COLUMNS.index("ESS_bulk")]], dtype=torch.float64, device=dev)
    rhat_max = torch.tensor([float(np.nanmax(tab[cols, B.SUMMARY_COLUMNS.index("R_hat")]))], dtype=torch.float64, device=dev)
    if ess_window != K:   # the K timed draws alone (round 1's definition), reported beside
        tabK = e.summary(row0=Wx, nrows=K, constrained=True)
        essK = torch.tensor([v if v == v else float("inf") for v in tabK[cols, B.SUMMARY_COLUMNS.index("ESS_bulk")]], dtype=torch.float64, device=dev)
    else:
        essK = ess_cols.clone()
    ess_impl = "b200nuts_summary (CUDA; pinned to the stansummary goldens and to cmdstan_b200.analyze on hardware by tests/test_gpu_summary.py)"
    if primary and tab.shape[0] <= 400:
        # the headline uses ONE ESS implementation for both arms: cmdstan_b200.analyze (torch restatement of the reference's
        # split_rank_normalized_ess, pinned to stansummary), the same function run_reference applies to the reference's draws;
        # the device summary's value is kept beside it as a cross-check
        from cmdstan_b200 import analyze as A
        win = torch.tensor(e.fetch_draws(row0=r0, nrows=ess_window, constrained=True))
        _, per = A.min_bulk_ess(win, device=dev)
        dev_min = float(ess_cols.min())
        ess_cols = torch.tensor([v if v == v else float("inf") for v in per], dtype=torch.float64, device=dev)
        ess_impl = f"cmdstan_b200.analyze.min_bulk_ess (both arms); device summary of the same window gives {dev_min:.6g} on rank 0"
        del win
    note(f"{wname}: summary done")
    nleap = torch.tensor([float(draws[:, :, 4].sum()), float((ng2 - ng1).sum())], dtype=torch.float64, device=dev)
    times = torch.tensor([run_s, run_wall + fetch_s + h2d_s, grad_s, adapt_s, float(ar_us or 0.0)], dtype=torch.float64, device=dev)
    div = torch.tensor([float(draws[:, :, 5].sum()), float(draws[:, :, 3].mean()), float(np.max(ng2 - ng1))], dtype=torch.float64, device=dev)
    if world > 1:
        if not row_mode:   # row mode: every rank holds the same chains -- nothing to add up
            dist.all_reduce(ess_cols, op=dist.ReduceOp.SUM)   # ESS of independent chain sets adds
            dist.all_reduce(essK, op=dist.ReduceOp.SUM)
            dist.all_reduce(nleap, op=dist.ReduceOp.SUM)
            dist.all_reduce(div[:2], op=dist.ReduceOp.SUM)
        dist.all_reduce(rhat_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    min_ess_w, min_ess_K = float(ess_cols.min()), float(essK.min())
    run_s, e2e_s, grad_s, adapt_s, ar_us = [float(v) for v in times]
    n_chains_job = C * (1 if row_mode else world)
    ess_per_draw = min_ess_w / (n_chains_job * ess_window)
    ess_timed = ess_per_draw * n_chains_job * K            # ESS of the K timed transitions at the window's ESS-per-draw rate
    out = None
    leftovers = dict(engine=e, data=data, eps=eps, im=im, Wx=Wx, K=K, C=C)
    if rank == 0:
        pk, pk_src = peaks()
        work = algorithmic_work(w, data, model, e.D)
        n_grad_launch = max(1, int(div[2]))   # lock-step: one gradient-stage launch per leapfrog of the slowest chain
        n_chain_grads = float(nleap[1])       # chain-gradients executed in the timed region (all ranks)
        roof = None
        if work["kind"] == "tensor":
            # dominant kernel: fused GLM gradient; bound: FP64 tensor pipe; peak measured live with a DGEMM.  achieved = algorithmic
            # flops of the chain-gradients actually served (requesting chains are compacted) / time in the gradient stage
            fl = work["flops"] * n_chain_grads / world   # per rank: chains/world (chain mode) or rows/world (row mode)
            ach = fl / grad_s / 1e12 if grad_s > 0 else None
            peak = measure_fp64_peak(dev)
            # which formulation served: the INT8 digit-plane kernel on tcgen05 (variant 7) is the default for the logistic link
            glm_variant = int(os.environ.get("B200NUTS_GLM_VARIANT", "7" if model == "logistic" else "1"))
            i8 = glm_variant == 7 and model == "logistic"
            tr = ncu_traffic("logistic_i8" if i8 else model) if scale == 1.0 and world == 1 else None
            bytes_rank = work["bytes_per_batch"] / (world if row_mode else 1)
            roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s (fp64)", "frac": (ach / peak) if ach and peak else None,
                    "traffic": tr["bytes"] if tr else None, "traffic_detail": tr, "algorithmic_bytes_per_launch": bytes_rank,
                    "kernel": "glm_i8_kernel (+q planes, reduce, finish)" if i8 else "glm_dmma_kernel (+reduce, finish)", "launches": n_grad_launch,
                    "avg_launch_ms": 1e3 * grad_s / n_grad_launch, "avg_chains_per_launch": n_chain_grads / (1 if row_mode else world) / n_grad_launch,
                    "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no fp64 figure)",
                    "hbm_gbs_achieved": bytes_rank / (grad_s / n_grad_launch) / 1e9 if grad_s > 0 else None, "hbm_peak_gbs": pk.get("hbm_gbs"), "hbm_peak_source": pk_src}
            if i8:
                # the same algorithmic fp64 flops, executed as 21 + 21 exact INT8 digit-plane products per (128 rows x 32 chains) tile
                # padded to K = 128 design columns / M = 128 gradient rows: the tensor work actually issued, against the INT8 peak
                N_rows = data["X"].shape[0]
                macs = (n_chain_grads / world / 32.0) * math.ceil(N_rows / 128.0) * 2 * 21 * 128 * 32 * 128
                roof["formulation"] = ("fp64 contractions as exact INT8 digit-plane products (6 planes x 8 bit, 46-bit operands, diagonals below 2^-48 dropped) on "
                                       "tcgen05.mma kind::i8 with INT32 accumulators in tensor memory; `achieved` counts the ALGORITHMIC fp64 flops (4 N Dp per "
                                       "chain-gradient), so frac > 1 of the DGEMM rate is possible: the work left the FP64 pipe")
                roof["int8_top_s_issued"] = 2.0 * macs / grad_s / 1e12 if grad_s > 0 else None
                roof["int8_peak_top_s"] = 4550.0
                roof["int8_peak_source"] = "tools/microbench/umma_i8.cu on this pool (profiles/r02m_umma_i8.txt: 7 825 MAC/clk/SM x 148 SMs x 1.965 GHz); nominal 4 500"
                roof["int8_frac"] = roof["int8_top_s_issued"] / roof["int8_peak_top_s"] if roof["int8_top_s_issued"] else None
                # what this formulation has to move per launch: the six digit planes of X once (they are re-read per chain tile from L2),
                # the partial gradients of 37 row groups written once and read once by the reduction
                n_req = n_chain_grads / world / n_grad_launch
                roof["i8_algorithmic_bytes_per_launch"] = 6.0 * math.ceil(N_rows / 128.0) * 128 * 112 + 2.0 * 37 * n_req * 104 * 8
            if row_mode:
                roof["allreduce_us_per_leapfrog"] = ar_us
                roof["allreduce_payload_bytes"] = ar_bytes
                roof["allreduce_frac_of_gradient_stage"] = (ar_us * 1e-6 * n_grad_launch / grad_s) if grad_s > 0 else None
        elif work["kind"] == "fp64":
            fl = work["flops"] * n_chain_grads / world
            ach = fl / grad_s / 1e12 if grad_s > 0 else None
            peak = measure_fp64_peak(dev)
            tr = ncu_traffic(model) if scale == 1.0 else None
            try:   # the bound of this kernel is the scalar FP64 pipe: its own measured FMA rate beside the DGEMM rate
                fma_peak = B.measure_fp64_fma_peak(local_rank)
            except Exception:
                fma_peak = None
            roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s (fp64)", "frac": (ach / peak) if ach and peak else None,
                    "fp64_fma_peak": fma_peak, "frac_of_fma_peak": (ach / fma_peak) if ach and fma_peak else None,
                    "traffic": tr["bytes"] if tr else None, "traffic_detail": tr, "kernel": "irt_grad_kernel", "launches": n_grad_launch, "avg_launch_ms": 1e3 * grad_s / n_grad_launch,
                    "note": "bound is FP64 ALU + transcendental throughput; flops at SURVEY 8d's 25 per response; normalised by the measured DGEMM rate"}
        else:
            # fused chain kernel: latency/occupancy-bound by construction (tens of flops per gradient); report HBM traffic rate
            byt = float(nleap[1]) * (work["bytes_per_chain"] + 8 * e.D * 8.0)
            ach = byt / run_s / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": pk.get("hbm_gbs"), "unit": "GB/s", "frac": ach / pk.get("hbm_gbs"), "traffic": None,
                    "kernel": "fused_chain_kernel", "note": "latency-bound by construction (SURVEY.md 8d): state stays in L1/L2; fraction is small on purpose", "peak_source": pk_src}
        out = {
            "value": ess_timed / run_s, "unit": "ESS/s", "ms_per_step": 1e3 * run_s / K,
            "scaling": "strong" if (row_mode or strong_chains) else "weak",
            "grad_evals_per_sec": float(nleap[0]) / run_s, "grad_evals_executed_per_sec": float(nleap[1]) / run_s,
            "min_bulk_ess_window": min_ess_w, "ess_per_draw": ess_per_draw, "ess_window_draws": ess_window, "ess_implementation": ess_impl,
            "value_from_timed_draws_only": min_ess_K / run_s, "rhat_max": float(rhat_max[0]),
            "chains": n_chains_job, "chains_per_gpu": C, "timed_draws": int(n_chains_job * K),
            "divergences": float(div[0]), "mean_treedepth": float(div[1]) / (1 if row_mode else world),
            "adaptation_s": adapt_s, "setup_wall_s": setup_wall, "median_stepsize": float(np.median(eps)),
            "lockstep_launches_per_transition": n_grad_launch / K,
            # also per (warm-up + sampling) second -- the ESS rate scaled to the reference's 1000 sampling transitions,
            # over the device time of the adaptation actually run plus 1000 transitions at the timed rate
            "ess_per_sec_incl_warmup_1000_draws": (ess_per_draw * n_chains_job * 1000.0) / (adapt_s + run_s * 1000.0 / K) if K > 0 else None,
            "e2e": {"value": ess_timed / e2e_s, "unit": "ESS/s", "h2d_bytes_per_step": data_bytes(data) / (world if row_mode else 1) / K,
                    "d2h_bytes_per_step": int(draws.nbytes / K), "h2d_s": h2d_s, "run_wall_s": run_wall, "d2h_s": fetch_s},
            "gpu_launches": int(launches), "roofline": roof, "clocks": clk,
        }
        if dump_warm:
            nk = min(C, 64)
            np.savez(dump_warm, q=e.fetch_draws(chain0=0, nchains=nk, row0=Wx + K - 1, nrows=1, constrained=False)[:, 0, 7:],
                     inv_metric=im[:nk], stepsize=eps[:nk])
    return out, leftovers


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed transitions (default: 200; 30 for c4 / c5)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=os.environ.get("B200_BENCH_WORKLOAD", "c3"))
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the workload's)")
    ap.add_argument("--num-warmup", type=int, default=-1, help="NUTS adaptation transitions before the timed region (default: 1000; 150 for c4 / c5)")
    ap.add_argument("--scale", type=float, default=1.0, help="scale the data rows (testing only; invalidates the number)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=45.0)
    ap.add_argument("--ess-window", type=int, default=-1, help="draws per chain the ESS-per-draw factor is estimated over (default 200 for c1-c3, the timed K otherwise)")
    ap.add_argument("--extras", default=os.environ.get("B200_BENCH_EXTRAS", "auto"),
                    help="additional workloads measured after the headline and attached to the line as `multi_gpu_configs` (north_star's two "
                         "multi-GPU configurations): 'auto' = c4 (8192 chains split over the GPUs) and c5 (rows split over the GPUs, NCCL "
                         "all-reduce per leapfrog) when the headline workload is c3; 'none'; or a comma list")
    ap.add_argument("--dump-warm", default="", help="write the adapted state (q, inv_metric, stepsize of up to 64 chains) to this .npz: "
                    "the warm-start fixture of the reference arm (tests/golden/warm_<workload>.npz)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    t_start = time.time()

    def note(msg):   # progress on stderr (stdout carries exactly one JSON line)
        if os.environ.get("B200_BENCH_VERBOSE"):
            print(f"[bench +{time.time() - t_start:7.1f}s] {msg}", file=sys.stderr, flush=True)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[args.workload]
    if args.steps <= 0:
        args.steps = w.get("steps", 200)
    if args.num_warmup < 0:
        args.num_warmup = w.get("num_warmup", 1000)
    model = w["model"]
    cores = os.cpu_count() or 1
    K, W = max(args.steps, 4), max(args.warmup, 3)
    scale_note = "" if args.scale == 1.0 else f" [rows scaled x{args.scale}]"
    ess_window = args.ess_window if args.ess_window > 0 else (max(K, ESS_WINDOW) if args.workload in ("c1", "c2", "c3") else K)
    C = args.chains or w["chains"]
    chains_total = C if model == "poisson_glm" else C * world
    cfg_common = common_config(args.workload, w, args.num_warmup, K, world, scale_note, chains_total, ess_window)

    if args.impl == "reference":
        if rank != 0:
            return
        data = make_data(w, args.scale)
        warm = None
        fx = os.path.join(os.environ.get("B200_WARM_FIXTURE_DIR", os.path.join(ROOT, "tests", "golden")), f"warm_{args.workload}.npz")
        if model not in ("bernoulli", "eight_schools_nc") and os.path.exists(fx) and args.scale == 1.0:
            z = np.load(fx)
            warm = dict(q=z["q"], inv_metric=z["inv_metric"], stepsize=z["stepsize"])
        elif model not in ("bernoulli", "eight_schools_nc"):
            print(json.dumps({"impl": "reference", "unavailable": f"no warm-start fixture tests/golden/warm_{args.workload}.npz for this workload"}))
            return
        # bounded sample: enough CPU time for about ESS_WINDOW transitions per chain at c3 (0.2 s each on one core)
        r = run_reference(args.workload, w, data, cores, warm, budget_s=max(60.0, 12.0 * K / 10), scale_note=scale_note, max_draws=max(ess_window, 8))
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref binaries are missing (built from /root/reference by oracle/build_ref.sh)"}))
            return
        line = {"impl": "reference", "metric": "min_bulk_ESS_per_sec", "value": r["value"], "unit": "ESS/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * r["sampling_s"] / max(1, r["draws"] // max(1, (4 if args.workload == "c1" else cores))),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg_common,
                "grad_evals_per_sec": r["grad_evals_per_s"], "min_bulk_ess": r["min_bulk_ess"], "draws_per_chain": r["draws_per_chain"],
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "ESS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    res, left = run_gpu_workload(args.workload, K, W, args.num_warmup, C, args.scale, rank, world, local_rank, note, ess_window,
                                 primary=True, dump_warm=args.dump_warm)
    guard = LineGuard(float(os.environ.get("B200_BENCH_GUARD_S", "660")) - (time.time() - t_start))
    line = None
    if rank == 0:
        line = {"metric": "min_bulk_ESS_per_sec", "value": res["value"], "unit": "ESS/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": res["scaling"], "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": cfg_common}
        line.update({k: v for k, v in res.items() if k not in ("value", "unit", "ms_per_step", "scaling")})
        line["notes"] = ["value = (min bulk-ESS per draw over the last ess_window_draws draws of every chain, incl. the timed ones) x timed draws / "
                         "device time of the timed transitions; value_from_timed_draws_only is round 1's definition (ESS of the K timed draws alone, "
                         "optimistic for short K)"]
        guard.arm(line)
    e = left["engine"]
    # ---- CPU baseline (rank 0, N = 1 only: the reference arm of the driver covers the other N)
    if rank == 0 and not args.no_cpu and world == 1:
        try:
            eps, im, Wx = left["eps"], left["im"], left["Wx"]
            warm = dict(q=e.fetch_draws(chain0=0, nchains=min(C, cores), row0=Wx + K - 1, nrows=1, constrained=False)[:, 0, 7:],
                        inv_metric=im[:min(C, cores)], stepsize=eps)
            e.close()
            cb = run_reference(args.workload, w, left["data"], cores, warm, budget_s=args.cpu_budget, scale_note=scale_note, max_draws=max(ess_window, 8))
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")} if cb else {"value": None, "unit": "ESS/s", "cores": cores, "kind": "reference", "sample": "oracle/_ref missing"}
            if cb:
                line["cpu_baseline"]["grad_evals_per_sec"] = cb["grad_evals_per_s"]
        except Exception as ex:  # the CPU leg must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": "ESS/s", "cores": cores, "kind": "reference", "sample": f"failed: {ex}"[:300]}
    e.close()
    left.clear()
    # ---- north_star's two multi-GPU configurations, measured in the same run and attached to the line
    extras = args.extras
    if extras == "auto":
        extras = "c4,c5" if (args.workload == "c3" and args.scale == 1.0 and not args.chains) else "none"
    if extras != "none":
        attached = {}
        for xw in extras.split(","):
            if time.time() - t_start > float(os.environ.get("B200_BENCH_EXTRA_START_LIMIT_S", "420")):
                attached[xw] = {"skipped": "time budget of the bench run"}
                continue
            try:
                ww = WORKLOADS[xw]
                # c4: 8192 chains split over the GPUs (strong scaling, no collective); c5: all chains on every GPU, rows split
                xc = ww["chains"] if ww["model"] == "poisson_glm" else max(1, ww["chains"] // world)
                xr, xl = run_gpu_workload(xw, ww.get("steps", 30), 3, ww.get("num_warmup", 150), xc, 1.0, rank, world, local_rank, note,
                                          ww.get("steps", 30), primary=False, strong_chains=True)
                xl["engine"].close()
                xl.clear()
                if rank == 0:
                    xr["config"] = common_config(xw, ww, ww.get("num_warmup", 150), ww.get("steps", 30), world, "", xr["chains"], xr["ess_window_draws"])
                    attached[xw] = xr
            except Exception as ex:
                attached[xw] = {"failed": f"{type(ex).__name__}: {ex}"[:300]}
        if rank == 0:
            line["multi_gpu_configs"] = attached
    if rank == 0:
        guard.finish()
    if world > 1:
        dist.destroy_process_group()


def measure_fp64_peak(dev):
    """FP64 matmul rate of this GPU right now (cuBLAS DGEMM 8192^3, best of 5): the roofline denominator for the fp64
    tensor-core kernels -- MEASURED_PEAKS.json only carries bf16 and HBM figures."""
    import torch
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    best = 0.0
    for _ in range(5):
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        torch.matmul(a, b)
        t.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (s.elapsed_time(t) * 1e-3) / 1e12)
    return best


if __name__ == "__main__":
    main()
