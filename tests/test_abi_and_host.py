"""CPU suite: the C-ABI library loads without a GPU and exports every symbol include/b200nuts.h declares; the product
fails loudly without a CUDA device (no CPU fallback); the product package never references the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_header_symbols_are_exported():
    import cmdstan_b200 as B
    lib = B.load_library()
    hdr = open(os.path.join(ROOT, "include", "b200nuts.h")).read() + open(os.path.join(ROOT, "include", "b200nuts_model.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(b200nuts_[a-z0-9_]+)\s*\(", hdr)) - {"b200nuts_allreduce_fn", "b200nuts_model_entry", "b200nuts_model_entry_fn", "b200nuts_init_not_given"}  # the last: a static inline helper of the plug-in header
    assert declared == set(B.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_config_struct_layout_matches_header():
    import cmdstan_b200 as B
    lib = B.load_library()
    cfg = B.Config()
    lib.b200nuts_default_config(ctypes.byref(cfg))
    # reference defaults (src/cmdstan/arguments/arg_sample.hpp:18-29, arg_adapt.hpp:19-37, arg_nuts.hpp:15-16, arg_hmc.hpp:19-27)
    assert (cfg.num_chains, cfg.num_warmup, cfg.num_samples, cfg.num_thin, cfg.save_warmup) == (1, 1000, 1000, 1, 0)
    assert (cfg.delta, cfg.gamma, cfg.kappa, cfg.t0) == (0.8, 0.05, 0.75, 10.0)
    assert (cfg.init_buffer, cfg.term_buffer, cfg.window, cfg.max_depth) == (75, 50, 25, 10)
    assert (cfg.stepsize, cfg.stepsize_jitter, cfg.init_radius, cfg.init_chain_id, cfg.adapt_engaged) == (1.0, 0.0, 2.0, 1, 1)
    assert (cfg.row_shards, cfg.row_shard_rank, cfg.store_draws) == (1, 0, 1)


def test_usage_errors_before_touching_the_device():
    import cmdstan_b200 as B
    with pytest.raises(B.B200NutsError) as ei:
        B.Engine("bernoulli", {"y": np.array([0, 1])}, num_chains=0)
    assert ei.value.code == B.USAGE
    with pytest.raises(B.B200NutsError) as ei:
        B.Engine("bernoulli", {"y": np.array([0, 1])}, max_depth=40)
    assert ei.value.code == B.USAGE
    with pytest.raises(B.B200NutsError) as ei:
        B.Engine("bernoulli", {"y": np.array([0, 1])}, num_warmup=0)  # adapting with num_warmup = 0 (command.hpp:569-574)
    assert ei.value.code == B.USAGE


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cmdstan_b200 as B
    with pytest.raises(B.B200NutsError) as ei:
        B.Engine("bernoulli", {"y": np.array([0, 1])})
    assert ei.value.code == B.SOFTWARE and "no CUDA device" in str(ei.value)
    rc, res = B.hmc_nuts_diag_e_adapt("bernoulli", {"y": np.array([0, 1])})
    assert rc == B.SOFTWARE and res is None


def test_plugin_registry_without_a_gpu():
    """include/b200nuts_model.h: a plug-in library registers its models by name; ids start at 100; built-in names are kept;
    a malformed vtable is refused; constructing an engine for a plug-in model still fails loudly without a device."""
    import cmdstan_b200 as B
    from cmdstan_b200 import _build
    lib = B.load_library()
    first = B.load_model_library(_build.build_example_plugins())
    assert first >= 100 and B.find_model("linreg") == first and B.find_model("eight_schools_plugin") == first + 1
    assert B.load_model_library(_build.EXAMPLE_PLUGIN) == first          # idempotent
    assert [B.find_model(n) for n in B.MODELS] == list(B.MODELS.values())  # built-ins by name
    with pytest.raises(KeyError):
        B.find_model("no_such_model")
    with pytest.raises(B.B200NutsError) as ei:
        B.load_model_library("/nonexistent/libmodel.so")
    assert ei.value.code == B.NOINPUT
    with pytest.raises(B.B200NutsError) as ei:
        B.load_model_library(_build.LIB)   # a library without b200nuts_model_entry
    assert ei.value.code == B.USAGE
    bad = (ctypes.c_uint32 * 16)(99)  # wrong abi
    assert lib.b200nuts_register_model(ctypes.cast(bad, ctypes.c_void_p)) == -B.USAGE
    assert lib.b200nuts_register_model(None) == -B.USAGE
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(B.B200NutsError) as ei:
            B.Engine("linreg", {"X": np.zeros((3, 1)), "y": np.zeros(3)})
        assert ei.value.code == B.SOFTWARE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cmdstan_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("# oracle", "").lower() or f == "analyze.py" and "oracle" not in txt.lower(), f


def test_default_chain_kernels_stay_small():
    """The fused chain kernels of the small models are instruction-cache bound: carrying the dense-metric branches cost c2 15 %, the
    diagnostics bookkeeping 7 % (profiles r01i, r01l).  Optional features are compiled only into csrc/dense.cu's kernels; this
    guards the text size of the default ones (83.8 KB in the build that measured 1.04e7 ESS/s on c2)."""
    import shutil
    import subprocess
    import cmdstan_b200 as B
    from cmdstan_b200 import _build
    B.load_library()
    obj = os.path.join(_build.CSRC, "engine.o")
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(obj) or not os.path.exists(tool):
        pytest.skip("engine.o / cuobjdump not available")
    out = subprocess.run([tool, "-elf", obj], capture_output=True, text=True).stdout
    sizes = {}
    for ln in out.splitlines():
        p = ln.split()
        if len(p) >= 4 and p[-1].startswith(".text._ZN4b200") and ("fused_chain_kernel" in p[-1] or "advance_kernel" in p[-1]):
            sizes[p[-1]] = int(p[2], 16)
    fused = [v for k, v in sizes.items() if "fused_chain_kernelILb1ELi0" in k]
    assert fused and fused[0] <= 90 * 1024, sizes


def test_all_service_mirrors_compile():
    """The front end reaches the twelve service mirrors through two of them; the other ten are templates nobody instantiates.
    tests/host_emul/service_instantiation.cpp calls each with the front end's writer types and the reference's argument
    order (services/sample/hmc_{nuts,static}_{diag,unit,dense}_e[_adapt].hpp) -- a syntax-only compile keeps them honest."""
    import subprocess
    src = os.path.join(HERE, "host_emul", "service_instantiation.cpp")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_product_build_has_no_test_rig_code():
    """B200_SIMT_EMUL / B200_HOST_EMUL / B200_FASTMATH_HOST exist for the CPU test rigs only: the build recipe never defines them
    and nothing of the rigs' namespace ends up in the shipped library"""
    import subprocess
    from cmdstan_b200 import _build
    recipe = open(_build.__file__).read()
    for macro in ("B200_SIMT_EMUL", "B200_HOST_EMUL", "B200_FASTMATH_HOST"):
        assert macro not in recipe
    if not os.path.exists(_build.LIB):
        pytest.skip("library not built")
    syms = subprocess.run(["nm", "-C", _build.LIB], capture_output=True, text=True).stdout
    assert "simt::" not in syms and "emul_" not in syms and "oracle_" not in syms
