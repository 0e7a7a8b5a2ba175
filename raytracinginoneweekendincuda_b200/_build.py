"""Build recipe: nvcc for sm_100a, in-tree, no JIT cache.

  librt_b200.so  (this package dir)  the product: C ABI + kernels + host scene surface
  rt_cli         (this package dir)  command-line renderer (PPM out)
  oracle/liboracle.so                the FP64 checker (tests / bench baseline only)
  oracle/_ref/*                      reference-derived checkers, only where /root/reference exists
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


def oracle_path() -> str:
    return os.path.join(ROOT, "oracle", "liboracle.so")


def build_product(force: bool = False, verbose_ptxas: bool = False) -> str:
    deps = _sources([CSRC, os.path.join(ROOT, "include")])
    target = lib_path()
    if force or not _newer(target, deps):
        extra = ["-Xptxas", "-v"] if verbose_ptxas else []
        _run(["nvcc", *NVCC_FLAGS, *extra, "-shared",
              os.path.join(CSRC, "rt_device.cu"), os.path.join(CSRC, "rt_host.cpp"),
              os.path.join(CSRC, "rt_error.cpp"), "-o", target])
    cli = os.path.join(PKG, "rt_cli")
    if force or not _newer(cli, deps):
        _run(["nvcc", *NVCC_FLAGS, os.path.join(CSRC, "rt_cli.cpp"), "-o", cli, "-L", PKG, "-lrt_b200",
              "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"])
    return target


def build_oracle(force: bool = False) -> str:
    src = os.path.join(ROOT, "oracle", "rt_oracle.cpp")
    target = oracle_path()
    if force or not _newer(target, [src, os.path.join(ROOT, "include", "rt_abi.h")]):
        _run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-pthread", "-Wall", "-Wextra",
              src, "-o", target])
    return target


def build_ref() -> None:
    """oracle/_ref from /root/reference, when it is there (build container only)."""
    if not os.path.isdir(os.environ.get("RT_REFERENCE_DIR", "/root/reference")):
        return
    out = os.path.join(ROOT, "oracle", "_ref")
    need = [os.path.join(out, "libref_stream.so"), os.path.join(out, "ref_gpu")]
    srcs = [os.path.join(ROOT, "oracle", f) for f in ("build_ref.py", "ref_stream_main.cpp", "ref_image.cpp")]
    srcs += _sources([os.path.join(ROOT, "oracle", "shim")])
    if all(_newer(t, srcs) for t in need):
        return
    _run([sys.executable, os.path.join(ROOT, "oracle", "build_ref.py")])


def build_all(force: bool = False) -> None:
    build_product(force)
    build_oracle(force)
    build_ref()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
