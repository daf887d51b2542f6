"""Build libgpp_b200.so (sm_100a only) in-tree with nvcc.  No JIT cache: the .so travels with the repo snapshot.

Usage:  python -m gpflowpilco_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libgpp_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", INCLUDE, "-I", CSRC,
]
LINK_LIBS = ["-lcusolver", "-lcublas", "-ldl"]


def _nvcc() -> str:
  nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
  if not os.path.exists(nvcc):
    raise RuntimeError("nvcc not found: libgpp_b200.so cannot be built (there is no CPU fallback)")
  return nvcc


def sources():
  return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_header_mtime() -> float:
  paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
  paths += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
  return max(os.path.getmtime(p) for p in paths)


def _compile(src: str, force: bool, verbose: bool) -> str:
  obj = os.path.join(OBJDIR, src[:-3] + ".o")
  spath = os.path.join(CSRC, src)
  if (not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(spath), _newest_header_mtime())):
    return obj
  cmd = [_nvcc(), *NVCC_FLAGS, "-c", spath, "-o", obj]
  res = subprocess.run(cmd, capture_output=True, text=True)
  with open(obj + ".log", "w") as f:
    f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
  if res.returncode != 0:
    raise RuntimeError(f"nvcc failed for {src}:\n{res.stderr[-4000:]}")
  if verbose:
    print(res.stderr)
  return obj


def build(force: bool = False, verbose: bool = False) -> str:
  os.makedirs(OBJDIR, exist_ok=True)
  with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
    objs = list(ex.map(lambda s: _compile(s, force, verbose), sources()))
  if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, *LINK_LIBS, "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
      raise RuntimeError(f"link failed:\n{res.stderr[-4000:]}")
  return LIB


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
