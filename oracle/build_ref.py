"""Compile the reference's own ``alt_cuda_corr`` (two source files, read where they lie under
/root/reference -- never copied) into ``oracle/_ref/alt_cuda_corr_ref*.so`` for sm_100.

TEST INFRASTRUCTURE ONLY.  The result is a torch/pybind11 extension exposing the reference's
``forward`` / ``backward``; GPU tests load it to check (a) the CPU restatement
``cer_oracle.corr_forward`` and (b) our drop-in kernel against the real reference kernel.
It does not run the reference's build system (setup.py); it is three direct compiler calls.
``oracle/_ref/`` is git-ignored but travels to the GPU box with gpurun.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CER_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")
NAME = "alt_cuda_corr_ref"


def so_path():
    return os.path.join(OUT, NAME + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force: bool = False) -> str:
    src_cpp = os.path.join(REF, "alt_cuda_corr", "correlation.cpp")
    src_cu = os.path.join(REF, "alt_cuda_corr", "correlation_kernel.cu")
    target = so_path()
    if not os.path.isfile(src_cu):
        if os.path.isfile(target):
            return target
        raise RuntimeError("reference sources not present and no prebuilt oracle/_ref")
    if os.path.isfile(target) and not force and \
            os.path.getmtime(target) > max(os.path.getmtime(src_cpp), os.path.getmtime(src_cu)):
        return target
    os.makedirs(OUT, exist_ok=True)
    import torch
    from torch.utils import cpp_extension as ce
    inc = []
    for p in ce.include_paths("cuda"):
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    defs = [f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    o_cpp = os.path.join(OUT, "correlation.o")
    o_cu = os.path.join(OUT, "correlation_kernel.o")
    cmds = [
        ["g++", "-O2", "-fPIC", "-std=c++17", "-c", src_cpp, "-o", o_cpp] + inc + defs,
        ["nvcc", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-gencode", "arch=compute_100,code=sm_100", "-c", src_cu, "-o", o_cu] + inc + defs,
    ]
    procs = [subprocess.Popen(c) for c in cmds]
    for p, c in zip(procs, cmds):
        if p.wait() != 0:
            raise RuntimeError("compile failed: " + " ".join(c))
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    link = ["g++", "-shared", o_cpp, o_cu, "-o", target, f"-L{libdir}", "-L/usr/local/cuda/lib64",
            "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
            f"-Wl,-rpath,{libdir}"]
    subprocess.check_call(link)
    os.remove(o_cpp)
    os.remove(o_cu)
    return target


def load():
    """Import the built extension (needs torch imported first)."""
    import importlib.util
    import torch  # noqa: F401
    path = so_path()
    if not os.path.isfile(path):
        raise RuntimeError("oracle/_ref not built")
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
