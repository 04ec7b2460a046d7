"""Stage-wise check of the tensor-core training step (set_train_precision('bf16' | 'fp16')) on a GPU box.

    python tools/tc_train_check.py [n_rays] [precision]

1. forward outputs vs the fp32 training forward on the same draws;
2. the stores: every activation slot against a torch restatement that consumes the PREVIOUS slot (so each layer is judged
   on its own), d_raw -> dz slots likewise, and the weight-gradient GEMM against dz^T @ act computed by torch in fp32 from
   the very same stores (isolates layout / descriptor bugs from rounding);
3. parameter gradients vs the fp32 training path (cosine, relative L2 per tensor);
4. timing of the step's kernels (CUDA events).
Test infrastructure: the product never imports this file."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import snerf_b200                                          # noqa: E402
from snerf_b200 import autograd as A                       # noqa: E402
from snerf_b200 import make_query_fn                       # noqa: E402
from snerf_b200.render import _linspace01                  # noqa: E402
from tools import synth                                    # noqa: E402

NC, NF = 64, 128


def layout(n_rays, Nc=NC, Nf=NF):
    pairs, S = (n_rays + 1) // 2, Nc + Nf
    rows_c, rows_f = pairs * 2 * Nc, pairs * 2 * S
    off, L = 0, {}

    def take(name, b):
        nonlocal off
        L[name] = off
        off += (b + 1023) // 1024 * 1024
    take("act_c", rows_c * 10 * 512); take("act_f", rows_f * 10 * 512)
    take("dz_c", rows_c * 10 * 512); take("dz_f", rows_f * 10 * 512)
    take("bits_c", rows_c * 9 * 32); take("bits_f", rows_f * 9 * 32)
    take("draw_c", n_rays * Nc * 16); take("draw_f", n_rays * S * 16)
    take("raw_c", n_rays * Nc * 16); take("raw_f", n_rays * S * 16)
    take("z_c", n_rays * Nc * 4); take("z_f", n_rays * S * 4)
    L["rows_c"], L["rows_f"], L["total"] = rows_c, rows_f, off
    return L


def store(ws, off, rows, dtype):
    """[slot][rows][256] view of a store kept as 4 KiB blocks [slot][row / 32][channel / 64][8-channel chunk][row % 32][8]"""
    t = ws[off:off + rows * 10 * 512].view(dtype).view(10, rows // 32, 4, 8, 32, 8)
    return t.permute(0, 1, 4, 2, 3, 5).reshape(10, rows, 256)


def mask_bits(ws, off, rows):
    """[9][rows][256] bool view of the relu' bit store ([slot][row / 32][channel / 64][row % 32] 64-bit words; in each 32-bit
    half bit w = column 2 w, bit 16 + w = column 2 w + 1)"""
    w = ws[off:off + rows * 9 * 32].view(torch.int64).view(9, rows // 32, 4, 32)
    w = w.permute(0, 1, 3, 2).reshape(9, rows, 4)                      # [slot][row][cb]
    col = torch.arange(64, device=ws.device)
    bit = (col // 32) * 32 + torch.where(col % 2 == 0, (col % 32) // 2, 16 + (col % 32) // 2)
    return ((w[..., None] >> bit) & 1).bool().reshape(9, rows, 256)


def make_nets(dev):
    from snerf_b200 import NeRF
    nets = []
    for seed in (20, 21):
        p = synth.nerf_params(seed, trunk_gain=1.5, sigma_bias=1.0)
        m = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
        nets.append(m.to(dev))
    return nets


def run(prec, rb, nets, q, draws, tgt, keep=None):
    snerf_b200.set_train_precision(prec)
    A._DEBUG_KEEP = keep
    for n in nets:
        for p in n.parameters():
            p.grad = None
    call = A._Call(rb, nets[0], nets[1], q.multires, q.multires_views, NC, NF, False, False, _linspace01(NC, rb.device),
                   _linspace01(NF, rb.device), *draws)
    out = A.render_rays_train(call)
    loss = ((out["rgb_map"] - tgt) ** 2).mean() + ((out["rgb0"] - tgt) ** 2).mean() + 0.01 * out["depth_map"].mean() \
        + 0.05 * (out["disp_map"] - 0.1).abs().mean() + 1e-3 * (out["weights"] ** 2).sum(-1).mean()
    loss.backward()
    torch.cuda.synchronize()
    A._DEBUG_KEEP = None
    snerf_b200.set_train_precision("fp32")
    grads = {f"{i}.{k}": p.grad.detach().clone() for i, n in enumerate(nets) for k, p in n.named_parameters()}
    return {k: v.detach().clone() for k, v in out.items()}, grads, float(loss)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-300))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    nets = make_nets(dev)
    q, _, _ = make_query_fn()
    rs = np.random.RandomState(3)
    c2w = synth.camera(0)
    o, d = synth.pinhole_rays(900, 1600, 1266.4, c2w, [816.3, 491.5])
    idx = rs.choice(900 * 1600, n, replace=False)
    rb = torch.from_numpy(synth.ray_batch(o.reshape(-1, 3)[idx], d.reshape(-1, 3)[idx], 1.8, 110.0)).to(dev)
    S = NC + NF
    draws = [torch.rand(n, NC, device=dev), torch.rand(n, NF, device=dev), torch.randn(n, NC, device=dev), torch.randn(n, S, device=dev)]
    tgt = torch.rand(n, 3, device=dev)

    out32, g32, l32 = run("fp32", rb, nets, q, draws, tgt)
    keep = []
    out16, g16, l16 = run(prec, rb, nets, q, draws, tgt, keep)
    print(f"loss fp32 {l32:.6f}  {prec} {l16:.6f}")
    for k in ("rgb_map", "rgb0", "depth_map", "acc_map", "weights", "z_all"):
        if k in out32 and k in out16:
            print(f"  out {k:10s} max abs diff {float((out32[k] - out16[k]).abs().max()):.3e}   rel L2 {rel(out16[k], out32[k]):.3e}")

    # ---- the stores
    ws = keep[0]
    L = layout(n)
    dt = torch.bfloat16 if prec == "bf16" else torch.float16
    ok = True
    for tag, net, rows, X in (("c", nets[0], L["rows_c"], NC), ("f", nets[1], L["rows_f"], S)):
        act = store(ws, L["act_" + tag], rows, dt).float()
        dz = store(ws, L["dz_" + tag], rows, torch.bfloat16).float()
        mb = mask_bits(ws, L["bits_" + tag], rows)
        for k_ in range(8):
            ok &= bool(torch.equal(mb[k_], act[1 + k_] > 0))
        ok &= bool(torch.equal(mb[8][:, :128], act[0, :, 128:256] > 0))
        print(f"  [{tag}] relu' bit store matches the stored activations: {ok}")
        sd = {k: v.detach().float() for k, v in net.state_dict().items()}
        wq = lambda w: w.to(dt).float()          # weights as the tensor core sees them
        wb = lambda w: w.to(torch.bfloat16).float()
        enc = act[0, :, :64]
        # forward layers, each from the previous stored slot
        for l in range(8):
            w, b = sd[f"pts_linears.{l}.weight"], sd[f"pts_linears.{l}.bias"]
            x = enc[:, :63] if l == 0 else (torch.cat([enc[:, :63], act[l]], 1) if l == 5 else act[l])
            ref = torch.relu(x @ wq(w).T + b)
            e = rel(act[1 + l], ref)
            print(f"  [{tag}] act h{l}: rel L2 vs layer-wise restatement {e:.3e}")
            ok &= e < 2e-2
        feat = act[8] @ wq(sd["feature_linear.weight"]).T + sd["feature_linear.bias"]
        print(f"  [{tag}] act feature: {rel(act[9], feat):.3e}")
        # backward chain from the stores
        nv = n * X
        draw = ws[L["draw_" + tag]:L["draw_" + tag] + nv * 16].view(torch.float32).view(nv, 4)
        drawp = torch.zeros(rows, 4, device=dev); drawp[:nv] = draw
        v = act[0, :, 128:256]
        dvp = (drawp[:, :3] @ sd["rgb_linear.weight"]) * (v > 0)
        print(f"  [{tag}] dz dvp: {rel(dz[0, :, :128], dvp):.3e}   d_raw copy: {rel(dz[0, :, 128:132], drawp):.3e}")
        dfe = dz[0, :, :128] @ wb(sd["views_linears.0.weight"][:, :256])
        print(f"  [{tag}] dz dfeature: {rel(dz[9], dfe):.3e}")
        d7 = (dz[9] @ wb(sd["feature_linear.weight"]) + drawp[:, 3:4] * sd["alpha_linear.weight"]) * (act[8] > 0)
        print(f"  [{tag}] dz dz7: {rel(dz[8], d7):.3e}")
        for l in range(7, 0, -1):
            w = sd[f"pts_linears.{l}.weight"]
            w = w[:, 63:] if l == 5 else w
            ref = (dz[1 + l] @ wb(w)) * (act[l] > 0)
            e = rel(dz[l], ref)
            print(f"  [{tag}] dz dz{l - 1}: {e:.3e}")
            ok &= e < 2e-2
        # weight gradients from the stores (fp32 torch GEMMs on the stored 16-bit values)
        i = 0 if tag == "c" else 1
        chk = []
        chk.append(("pts_linears.0.weight", dz[1].T @ enc[:, :63]))
        for l in range(1, 8):
            x = torch.cat([enc[:, :63], act[l]], 1) if l == 5 else act[l]
            chk.append((f"pts_linears.{l}.weight", dz[1 + l].T @ x))
        for l in range(8):
            chk.append((f"pts_linears.{l}.bias", dz[1 + l].sum(0)))
        chk.append(("feature_linear.weight", dz[9].T @ act[8])); chk.append(("feature_linear.bias", dz[9].sum(0)))
        chk.append(("views_linears.0.weight", dz[0, :, :128].T @ torch.cat([act[9], act[0, :, 64:91]], 1)))
        chk.append(("views_linears.0.bias", dz[0, :, :128].sum(0)))
        chk.append(("rgb_linear.weight", dz[0, :, 128:131].T @ v)); chk.append(("rgb_linear.bias", dz[0, :, 128:131].sum(0)))
        chk.append(("alpha_linear.weight", dz[0, :, 131:132].T @ act[8])); chk.append(("alpha_linear.bias", dz[0, :, 131:132].sum(0)))
        for name, ref in chk:
            e = rel(g16[f"{i}.{name}"], ref)
            flag = "" if e < 2e-3 else "   <-- MISMATCH"
            ok &= e < 2e-3
            print(f"  [{tag}] dW {name:26s} vs stores: {e:.3e}{flag}")
    # ---- gradients vs the fp32 path
    worst = 0.0
    for k in g32:
        a, b = g16[k].double().flatten(), g32[k].double().flatten()
        cos = float((a @ b) / (a.norm() * b.norm() + 1e-300))
        r = rel(g16[k], g32[k])
        worst = max(worst, r)
        print(f"  grad {k:30s} cos {cos:.5f}  rel L2 {r:.3e}")
    print(f"worst rel L2 vs fp32 training path: {worst:.3e};   store checks {'OK' if ok else 'FAILED'}")

    # ---- timing at 512 rays
    n2 = 512
    idx = rs.choice(900 * 1600, n2, replace=False)
    rb2 = torch.from_numpy(synth.ray_batch(o.reshape(-1, 3)[idx], d.reshape(-1, 3)[idx], 1.8, 110.0)).to(dev)
    draws2 = [torch.rand(n2, NC, device=dev), torch.rand(n2, NF, device=dev), torch.randn(n2, NC, device=dev), torch.randn(n2, S, device=dev)]
    tgt2 = torch.rand(n2, 3, device=dev)
    for p_ in ("tf32", prec):
        for _ in range(3):
            run(p_, rb2, nets, q, draws2, tgt2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run(p_, rb2, nets, q, draws2, tgt2)
        e1.record(); torch.cuda.synchronize()
        print(f"512-ray fwd+loss+bwd ({p_}, incl. host overhead of this script): {e0.elapsed_time(e1) / 10:.3f} ms")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
