"""Build recipe: nvcc for sm_100a, in-tree, no JIT cache.

  librt_b200.so  (this package dir)  the product: C ABI + kernels + host scene surface
  rt_cli         (this package dir)  command-line renderer (PPM out)
The checker (oracle/) has its own recipes in oracle/bindings.py; CMakeLists.txt builds the same two targets.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources(dirs: list[str]) -> list[str]:
    out = []
    for d in dirs:
        for base, _, files in os.walk(d):
            out += [os.path.join(base, f) for f in files if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h"))]
    return out


def _run(cmd: list[str]) -> None:
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def lib_path() -> str:
    return os.path.join(PKG, "librt_b200.so")


def build_product(force: bool = False, verbose_ptxas: bool = False, target: str | None = None,
                  defines: list[str] | None = None) -> str:
    """`target` / `defines`: an A/B build of the library with -D options into another file (load it with
    RT_B200_LIBRARY); the default builds the in-tree product and the CLI."""
    deps = _sources([CSRC, os.path.join(ROOT, "include")])
    variant = target is not None
    target = target or lib_path()
    if force or variant or not _newer(target, deps):
        extra = ["-Xptxas", "-v"] if verbose_ptxas else []
        extra += [f"-D{d}" for d in (defines or [])]
        _run(["nvcc", *NVCC_FLAGS, *extra, "-shared",
              os.path.join(CSRC, "rt_device.cu"), os.path.join(CSRC, "rt_host.cpp"),
              os.path.join(CSRC, "rt_jpeg.cpp"), os.path.join(CSRC, "rt_error.cpp"), "-o", target])
    if variant:
        return target
    cli = os.path.join(PKG, "rt_cli")
    if force or not _newer(cli, deps):
        _run(["nvcc", *NVCC_FLAGS, os.path.join(CSRC, "rt_cli.cpp"), "-o", cli, "-L", PKG, "-lrt_b200",
              "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"])
    return target


def build_all(force: bool = False) -> None:
    build_product(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
