"""Builds libb200nuts.so (CUDA engine + C ABI) in-tree with nvcc for sm_100a.

The sampler translation unit is compiled with --fmad=false: the reference host build is x86-64 baseline (no FMA
contraction), and matched-stream parity of the trajectory arithmetic needs the same roundings.
The GLM gradient kernels keep FMA contraction: their reductions are re-associated anyway.  irt.cu holds the IRT gradient
AND the latency kernel that advances a chain (nuts.cuh state machine) in the same launch, so it is built without contraction
too; the gradient's fusions are explicit fma() calls.
summary.cu is also built without contraction so that the write_array arithmetic it applies on the fly (model_output_column)
rounds exactly like the write_array of b200nuts_fetch_draws (engine.cu).
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200nuts.so")
HOST = os.path.join(HERE, "host")
BIN = os.path.join(HERE, "bin")
CLI = os.path.join(BIN, "b200stan")
CLI_MODELS = ["std_normal", "bernoulli", "eight_schools_nc", "logistic", "irt_2pl", "poisson_glm"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
EXAMPLES = os.path.join(HERE, "..", "examples", "plugin_models")
EXAMPLE_PLUGIN = os.path.join(EXAMPLES, "libb200models_example.so")
UNITS = [("engine.cu", ["--fmad=false"]), ("advance_big.cu", ["--fmad=false"]), ("advance_cta.cu", ["--fmad=false"]), ("dense.cu", ["--fmad=false"]), ("glm.cu", []), ("glm_dmma.cu", []), ("glm_i8.cu", []), ("irt.cu", ["--fmad=false"]), ("summary.cu", ["--fmad=false"])]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def sources():
    inc = os.path.join(HERE, "..", "include")
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".o")] + [os.path.join(inc, "b200nuts.h"), os.path.join(inc, "b200nuts_model.h")]


def host_sources():
    return [os.path.join(HOST, f) for f in os.listdir(HOST)]


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(CLI):
        return True
    t = os.path.getmtime(LIB)
    if any(os.path.getmtime(s) > t for s in sources()):
        return True
    t = os.path.getmtime(CLI)
    return any(os.path.getmtime(s) > t for s in host_sources() + [LIB])


def build_cli():
    """C++ host side above the C ABI: cmdstan_b200/bin/b200stan (+ one link per model name, as CmdStan has one
    executable per model)."""
    os.makedirs(BIN, exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-o", CLI, os.path.join(HOST, "main.cpp"), "-L" + HERE, "-lb200nuts",
           "-Wl,-rpath,$ORIGIN/..", "-lpthread"]
    subprocess.check_call(cmd)
    for m in CLI_MODELS:
        link = os.path.join(BIN, m)
        if os.path.lexists(link):
            os.remove(link)
        os.symlink("b200stan", link)
    return CLI


def build_example_plugins(force=False):
    """examples/plugin_models/libb200models_example.so: user-written device models behind include/b200nuts_model.h
    (also the fixture of tests/test_gpu_plugin.py).  --fmad=false like the sampler TU (the eight-schools twin must round
    like the built-in)."""
    src = os.path.join(EXAMPLES, "example_models.cu")
    deps = [src, os.path.join(CSRC, "models.cuh"), os.path.join(CSRC, "warp.cuh"), os.path.join(HERE, "..", "include", "b200nuts_model.h")]
    if not force and os.path.exists(EXAMPLE_PLUGIN) and all(os.path.getmtime(d) <= os.path.getmtime(EXAMPLE_PLUGIN) for d in deps):
        return EXAMPLE_PLUGIN
    cmd = [_nvcc()] + ARCH + COMMON + ["--fmad=false", "-shared", src, "-o", EXAMPLE_PLUGIN]
    subprocess.check_call(cmd)
    return EXAMPLE_PLUGIN


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    for src, extra in UNITS:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [_nvcc()] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + out)
    cmd = [_nvcc()] + ARCH + ["-shared", "-o", LIB] + objs
    subprocess.check_call(cmd)
    build_cli()
    build_example_plugins(force)
    return LIB


if __name__ == "__main__":
    import sys
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
