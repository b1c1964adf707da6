"""Build nvit_b200/libnvit_b200.so (sm_100a only) from nvit_b200/csrc/*.cu with nvcc.

    python -m nvit_b200.build [--force]

The library is built IN-TREE so that it travels with the repo snapshot to the GPU box; nvcc cross-compiles sm_100a
without a GPU.  No torch headers are involved: the boundary is a plain C ABI (include/nvit_b200.h).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(ROOT, "build", "obj")
LIB_PATH = os.path.join(HERE, "libnvit_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"),
] + os.environ.get("NVIT_EXTRA_NVCC_FLAGS", "").split()


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libnvit_b200.so")
    return exe


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, hooks: bool = False) -> str:
    """hooks=True builds nvit_b200/libnvit_b200_hooks.so with -DNVIT_BENCH_HOOKS (the measurement-only entry points of
    include/nvit_b200_tuning.h, section 2) into its own object directory; the product library never contains them."""
    global OBJ_DIR, LIB_PATH
    if hooks:
        saved = (OBJ_DIR, LIB_PATH, list(NVCC_FLAGS))
        OBJ_DIR, LIB_PATH = os.path.join(ROOT, "build", "obj_hooks"), os.path.join(HERE, "libnvit_b200_hooks.so")
        NVCC_FLAGS.append("-DNVIT_BENCH_HOOKS")
        try:
            return build(force=force, verbose=verbose)
        finally:
            OBJ_DIR, LIB_PATH = saved[0], saved[1]
            NVCC_FLAGS[:] = saved[2]
    os.makedirs(OBJ_DIR, exist_ok=True)
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(ROOT, "include", "nvit_b200.h"), os.path.join(ROOT, "include", "nvit_b200_tuning.h")]
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, hooks="--hooks" in sys.argv))
