"""Import the UNMODIFIED reference (`/root/reference/s-nerf/model`) in the build container.

TEST INFRASTRUCTURE ONLY.  `/root/reference` does not exist on the GPU box, so nothing
that runs there (`-m gpu` tests, smoke(), bench.py) may import this module; it is used
by `oracle/make_golden.py` and `oracle/check_against_reference.py` to pin the numpy
oracle and to generate the committed fixtures under `tests/golden/`.

The reference needs `matplotlib` (absent here) only for an unused plotting helper
(run_nerf_helpers.py:12), so an empty stub module is injected; importing the module
also switches autograd anomaly mode on globally (run_nerf_helpers.py:2) which we undo.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("SNERF_REFERENCE_ROOT", "/root/reference/s-nerf")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


def load():
    """Returns (render_module, helpers_module) of the reference."""
    if not available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    import torch
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from model import render as ref_render, run_nerf_helpers as ref_helpers  # type: ignore
    torch.autograd.set_detect_anomaly(False)
    return ref_render, ref_helpers


def build_reference_net(ref_helpers, params: dict, D=8, W=256, input_ch=63, input_ch_views=27):
    """Reference `NeRF` module carrying the given numpy state_dict."""
    import torch
    net = ref_helpers.NeRF(D=D, W=W, input_ch=input_ch, input_ch_views=input_ch_views,
                           output_ch=5, skips=[4], use_viewdirs=True)
    sd = {k: torch.from_numpy(v.copy()) for k, v in params.items() if not k.startswith("_")}
    net.load_state_dict(sd)
    return net.eval()


def reference_query_fn(ref_helpers, multires=10, multires_views=4, netchunk=1 << 16):
    """The closure `create_nerf` builds (render.py:215-218)."""
    embed_fn, _ = ref_helpers.get_embedder(multires, 0)
    embeddirs_fn, _ = ref_helpers.get_embedder(multires_views, 0)
    return lambda inputs, viewdirs, network_fn: ref_helpers.run_network(
        inputs, viewdirs, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=netchunk)
