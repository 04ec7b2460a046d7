"""TEST INFRASTRUCTURE ONLY -- compiles the reference's OWN hash-grid encoder, unmodified and from where it lies under
/root/reference (s-nerfpp/zipnerf/gridencoder/src/{gridencoder.cu,bindings.cpp}), into oracle/_ref/_gridencoder_ref.so
for sm_100a with one explicit nvcc command (the reference's setup.py / JIT loader are not used; they pin -std=c++14,
which torch 2.11's headers reject).  Nothing is copied into the repo: oracle/_ref/ is git-ignored and holds only the
built module, which travels to the GPU box with the snapshot and is the live checker of the GPU parity tests
(tests/test_gpu_gridencoder.py) and the generator of tests/golden/grid_*.npz (oracle/make_golden_grid.py).

    python oracle/build_ref_gridencoder.py        # no-op when /root/reference is absent (e.g. on the GPU box)
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/s-nerfpp/zipnerf/gridencoder/src"
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "_gridencoder_ref"
OUT = os.path.join(OUT_DIR, NAME + ".so")


def build(verbose=False):
    srcs = [os.path.join(SRC, f) for f in ("gridencoder.cu", "bindings.cpp")]
    if not all(os.path.exists(s) for s in srcs):
        return None                       # reference not mounted: keep whatever was built earlier
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs + [__file__]):
        return OUT
    import torch
    from torch.utils import cpp_extension as ext
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = [f"-I{p}" for p in ext.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = (["nvcc", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-w",
            "-gencode", "arch=compute_100a,code=sm_100a",
            "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__",
            f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
           + inc + srcs
           + [f"-L{lib}", "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
              f"-Xlinker=-rpath={lib}", "-o", OUT])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the reference grid encoder failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(" ".join(cmd))
    return OUT


def load():
    """The built reference module (None when it has not been built)."""
    if not os.path.exists(OUT):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose=True))
    print(load())
