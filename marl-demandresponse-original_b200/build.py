"""Builds csrc/libmdr_b200.so for sm_100a with nvcc (in-tree, so the .so travels to the GPU box)."""
import fcntl
import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(CSRC, "libmdr_b200.so")
SOURCES = ("mdr_kernels.cu", "mdr_abi.cu", "mdr_host.cu")
HEADERS = ("mdr_kernels.h", "mdr_device.cuh", "mdr_pipe.cuh", "mdr_pipe_split.cuh", "mdr_pipe_split_host.cuh", "mdr_fused.cuh", "mdr_populate.cuh", "mdr_big.cuh", "mdr_wide.cuh", "mdr_rollout.cuh", "mdr_compact.cuh", "mdr_expand.h",
           os.path.join("..", "..", "include", "mdr_b200.h"))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]
# (`--split-compile 0` builds 2.5x faster but measurably slower code: the fused kernel 13.6 vs 10.7 us per step, the
#  no-observation pipelined kernel 21.2 vs 18.9 us -- A/B on the same box, tools/ab_libs.sh; MDR_NVCC_EXTRA for dev builds)


class NvccMissing(RuntimeError):
    pass


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise NvccMissing("nvcc not found (set NVCC=/path/to/nvcc)")


def is_stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles into a temporary file and renames it over LIB_PATH, under an exclusive file lock: concurrent
    callers (one per torchrun rank) build once, and nobody can load a partially written library."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = find_nvcc()
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():  # another process built it while we waited for the lock
                return LIB_PATH
            extra = os.environ.get("MDR_NVCC_EXTRA", "").split()  # e.g. -DMDR_BLOCKS_256=4 for tuning experiments
            tmp = "%s.tmp.%d" % (LIB_PATH, os.getpid())
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + list(SOURCES)
            res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
            os.replace(tmp, LIB_PATH)
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
