"""Build libcer_mvs_b200.so in-tree with nvcc for sm_100a (no torch headers; a few seconds per file).

    python -m cer_mvs_b200.build [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "build")
LIB = os.path.join(HERE, "libcer_mvs_b200.so")
SOURCES = ["corr_ops.cu", "build_volume.cu", "build_volume_tc.cu", "update_hmma.cu", "update_tc.cu", "plan.cu", "io_ops.cu", "fusion_ops.cu", "encoder.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "cer_mvs_b200.h"))
    return hdrs


def _stale(target, deps):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _deps()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        r = subprocess.run(["nvcc"] + NVCC_FLAGS + ["-c", s, "-o", o], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        return s, r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        for s, log in ex.map(compile_one, jobs):
            with open(os.path.join(OBJ, os.path.basename(s) + ".ptxas.log"), "w") as f:
                f.write(log)
            if verbose:
                print(log)
    objs = [os.path.join(OBJ, src.replace(".cu", ".o")) for src in SOURCES]
    if force or jobs or _stale(LIB, objs):
        subprocess.check_call(["nvcc", "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                          "-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
