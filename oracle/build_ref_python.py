"""TEST / BASELINE INFRASTRUCTURE ONLY -- byte-compiles the reference's OWN hot-path modules, unmodified and from where they
lie under /root/reference (s-nerf/model/render.py, s-nerf/model/run_nerf_helpers.py), into oracle/_ref/snerf_ref_model/*.pyc.

Nothing is copied into the repo: oracle/_ref/ is git-ignored and holds only built artefacts (CPython 3.12 bytecode here, the
compiled grid encoder next to it), which travel to the GPU box with the snapshot.  There the unmodified reference is what
`bench.py --impl reference` and the `cpu_baseline` leg time on the host cores (`cpu_baseline.kind: "reference"`); without
the artefacts they fall back to the numpy / torch-CPU port (`kind: "port"`).

    python oracle/build_ref_python.py        # no-op when /root/reference is absent (e.g. on the GPU box)
"""
import hashlib
import json
import os
import py_compile
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/s-nerf/model"
PKG = "snerf_ref_model"
OUT_DIR = os.path.join(HERE, "_ref", PKG)
FILES = ("render.py", "run_nerf_helpers.py")


def build(verbose=False):
    srcs = [os.path.join(SRC, f) for f in FILES]
    if not all(os.path.exists(s) for s in srcs):
        return None                       # reference not mounted: keep whatever was built earlier
    os.makedirs(OUT_DIR, exist_ok=True)
    manifest = {"python": sys.version.split()[0], "files": {}}
    for s in srcs:
        out = os.path.join(OUT_DIR, os.path.basename(s) + "c")
        py_compile.compile(s, cfile=out, doraise=True, optimize=0)
        manifest["files"][os.path.basename(s)] = hashlib.sha256(open(s, "rb").read()).hexdigest()
        if verbose:
            print("compiled", s, "->", out)
    json.dump(manifest, open(os.path.join(OUT_DIR, "MANIFEST.json"), "w"), indent=1)
    return OUT_DIR


# ---- zip-NeRF proposal resampling (s-nerfpp/zipnerf/internal/stepfun.py + math.py): the composition
# tools/stepfun_bench.py times beside csrc/snerf_stepfun.cu on the GPU box (BASELINE configs[3])
ZIP_SRC = "/root/reference/s-nerfpp/zipnerf/internal"
ZIP_OUT = os.path.join(HERE, "_ref", "zipnerf_ref", "internal")
ZIP_FILES = ("stepfun.py", "math.py")


def build_zip(verbose=False):
    srcs = [os.path.join(ZIP_SRC, f) for f in ZIP_FILES]
    if not all(os.path.exists(s) for s in srcs):
        return None
    os.makedirs(ZIP_OUT, exist_ok=True)
    open(os.path.join(ZIP_OUT, "__init__.py"), "w").close()
    manifest = {"python": sys.version.split()[0], "files": {}}
    for s in srcs:
        out = os.path.join(ZIP_OUT, os.path.basename(s) + "c")
        py_compile.compile(s, cfile=out, doraise=True, optimize=0)
        manifest["files"][os.path.basename(s)] = hashlib.sha256(open(s, "rb").read()).hexdigest()
        if verbose:
            print("compiled", s, "->", out)
    json.dump(manifest, open(os.path.join(ZIP_OUT, "MANIFEST.json"), "w"), indent=1)
    return ZIP_OUT


def load_zip_stepfun():
    """The byte-compiled reference `internal.stepfun` (sourceless import), or None.  math.py:5 decorates one helper (erf,
    unused by the resampling step) with torch.jit.script, which needs the SOURCE file: the decorator is made the identity
    while the module is imported."""
    if not all(os.path.exists(os.path.join(ZIP_OUT, f + "c")) for f in ZIP_FILES):
        return None
    import importlib
    import torch
    root = os.path.dirname(ZIP_OUT)
    if root not in sys.path:
        sys.path.insert(0, root)
    script = torch.jit.script
    torch.jit.script = lambda f, *a, **k: f
    try:
        return importlib.import_module("internal.stepfun")
    finally:
        torch.jit.script = script


def available() -> bool:
    return all(os.path.exists(os.path.join(OUT_DIR, f + "c")) for f in FILES)


def load():
    """(render_module, helpers_module) of the byte-compiled reference (sourceless import), or None."""
    if not available():
        return None
    import torch
    for name in ("matplotlib", "matplotlib.pyplot"):       # run_nerf_helpers.py:12 imports pyplot for an unused plotting helper
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    root = os.path.join(HERE, "_ref")
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib
    helpers = importlib.import_module(PKG + ".run_nerf_helpers")
    render = importlib.import_module(PKG + ".render")
    torch.autograd.set_detect_anomaly(False)               # run_nerf_helpers.py:2 switches anomaly mode on globally
    # render.py:6 picks cuda whenever a GPU is visible and moves its own t_vals / t_rand there (render.py:330,344); this arm
    # runs the reference on the HOST cores, so its device global is set to what that line yields on a CPU-only machine
    render._DEVICE = torch.device("cpu")
    return render, helpers


if __name__ == "__main__":
    print(build(verbose=True))
    print(build_zip(verbose=True))
