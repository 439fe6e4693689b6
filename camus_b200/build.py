"""In-tree builds of the native libraries (no JIT cache: the .so files travel with the repo).

  libcamus_host.so  -- camus_b200/host/*.cpp, plain C++17 (parsing, flattening, DP, network)
  libcamus_b200.so  -- camus_b200/csrc/*.cu, hand-written sm_100a kernels + the C-ABI of
                       include/camus_b200.h (nvcc cross-compiles without a GPU)
  camus             -- the CLI (same flags as the reference's camus.go:142-197)

`python -m camus_b200.build [host|cuda|cli|all]`
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "_lib")

HOST_SO = os.path.join(LIB, "libcamus_host.so")
CUDA_SO = os.path.join(LIB, "libcamus_b200.so")
CLI_BIN = os.path.join(LIB, "camus")

CXXFLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-Wall", "-Wextra", "-pthread"]
NVCCFLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-Wall",
    "--expt-relaxed-constexpr",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    if r.stderr.strip():
        sys.stderr.write(r.stderr)


def _sources(sub: str, exts: tuple[str, ...]) -> list[str]:
    d = os.path.join(PKG, sub)
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def build_host(force: bool = False) -> str:
    os.makedirs(LIB, exist_ok=True)
    deps = _sources("host", (".cpp", ".hpp"))
    if force or not _newer(HOST_SO, deps):
        _run(["g++", *CXXFLAGS, "-shared", "-o", HOST_SO, os.path.join(PKG, "host", "host_api.cpp")])
    return HOST_SO


def nvcc_path() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB, exist_ok=True)
    srcs = _sources("csrc", (".cu",))
    deps = srcs + _sources("csrc", (".cuh", ".h")) + [os.path.join(ROOT, "include", "camus_b200.h")]
    if force or not _newer(CUDA_SO, deps):
        cmd = [nvcc_path(), *NVCCFLAGS, "-I", os.path.join(ROOT, "include"), "-shared", "-o", CUDA_SO, *srcs,
               "-cudart", "static", "-ldl"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        cmd += os.environ.get("CAMUS_NVCC_EXTRA", "").split()
        _run(cmd)
    return CUDA_SO


def build_cli(force: bool = False) -> str:
    build_cuda(force)
    deps = _sources("host", (".cpp", ".hpp"))
    if force or not _newer(CLI_BIN, deps + [CUDA_SO]):
        _run(["g++", *CXXFLAGS, "-I", os.path.join(ROOT, "include"), "-o", CLI_BIN,
              os.path.join(PKG, "host", "camus_main.cpp"), "-L", LIB, "-lcamus_b200",
              "-Wl,-rpath,$ORIGIN", "-pthread"])
    return CLI_BIN


def build_all(force: bool = False) -> None:
    build_host(force)
    build_cuda(force)
    if os.path.exists(os.path.join(PKG, "host", "camus_main.cpp")):
        build_cli(force)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    force = "--force" in sys.argv
    {"host": build_host, "cuda": build_cuda, "cli": build_cli, "all": build_all}[what](force)
    print("built:", what)
