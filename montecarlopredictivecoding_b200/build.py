"""Build libmcpc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m montecarlopredictivecoding_b200.build [--force] [--verbose]
"""
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libmcpc_b200.so")
STAMP = LIB + ".stamp"
PROBES_LIB = os.path.join(PKG, "libmcpc_b200_probes.so")      # validation-only known-answer tests (csrc/probes)
PROBES_DIR = os.path.join(CSRC, "probes")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def probe_sources():
    return sorted(os.path.join(PROBES_DIR, f) for f in os.listdir(PROBES_DIR) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files += probe_sources()
    files.append(os.path.join(ROOT, "include", "mcpc_b200.h"))
    files.append(os.path.join(ROOT, "include", "mcpc_b200_probes.h"))
    for f in files:
        h.update(os.path.relpath(f, ROOT).encode())      # relative: the same tree gives the same digest on every machine
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(a if not os.path.isabs(a) else os.path.relpath(a, ROOT) for a in NVCC_FLAGS).encode())
    return h.hexdigest()


def is_stale():
    """True when libmcpc_b200.so is missing or was built from other sources than the ones in the tree."""
    if not (os.path.exists(LIB) and os.path.exists(PROBES_LIB) and os.path.exists(STAMP)):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != _digest()


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force=False, verbose=False, debug=False):
    """Compile when sources changed (sha256 stamp).  Returns the path of the library.
    ``debug=True`` builds libmcpc_b200_debug.so with -DMCPC_DEBUG_BUILD (experiment knobs such as MCPC_WIDE_EPI_MODE;
    select it at run time with MCPC_NATIVE_LIB=<path>); the product library never contains them."""
    if debug:
        lib = LIB.replace(".so", "_debug.so")
        extra = os.environ.get("MCPC_EXTRA_NVCC_FLAGS", "").split()       # experiments: e.g. -DMCPC_UPD_NS=5 -DMCPC_UPD_STG=6144
        if os.environ.get("MCPC_DEBUG_LIB_SUFFIX"):
            lib = LIB.replace(".so", "_debug" + os.environ["MCPC_DEBUG_LIB_SUFFIX"] + ".so")
        cmd = [nvcc_path()] + NVCC_FLAGS + ["-DMCPC_DEBUG_BUILD"] + extra + ["-o", lib] + sources()
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed building the debug library:\n" + proc.stderr[-4000:])
        cmd = [nvcc_path()] + NVCC_FLAGS + ["-DMCPC_DEBUG_BUILD", "-o", PROBES_LIB.replace(".so", "_debug.so")] + probe_sources()
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed building the debug probes library:\n" + proc.stderr[-4000:])
        return lib
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(PROBES_LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    for lib, srcs in ((LIB, sources()), (PROBES_LIB, probe_sources())):
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", lib] + srcs
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed building {os.path.basename(lib)}:\n" + proc.stderr[-4000:])
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, debug="--debug" in sys.argv))
