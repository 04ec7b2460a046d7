#!/usr/bin/env python
"""bench.py -- rays/sec of the render_rays hot path (64 coarse + 128 fine samples) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the fused renderer over one synthetic 1600x900 camera (1,440,000 rays,
BASELINE.json configs[1]: nuScenes-CAM_FRONT-like pinhole camera, near 1.8 m, far 110 m, NeRF 8x256 x2,
random-init weights), per GPU.  Rays shard embarrassingly: every rank renders its own camera, there
is no data-path collective (weak scaling).  Prints ONE JSON line on rank 0.

  value       rays/s with the ray batch resident in HBM (device-timed, CUDA events, max over ranks)
  e2e         same metric through the public API (snerf_b200.render) from pinned HOST buffers, with the
              host->device copy of the rays and the device->host read of the result inside the timed region
  roofline    achieved MLP TFLOP/s (303.83 MFLOP per ray) against the measured bf16
              tensor peak of MEASURED_PEAKS.json
  cpu_baseline  the numpy oracle (a port of the reference's PyTorch CPU path) timed on this box's host
              cores on a bounded sample of the same rays -- a reported baseline, not the target
`--impl reference` times that CPU port as its own arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, FOCAL, CX, CY = 900, 1600, 1266.4, 816.3, 491.5
NEAR, FAR = 1.8, 110.0
NC, NF = 64, 128
FLOP_PER_RAY = 303.83e6          # 256 samples x 1,186,816 FLOP, unpadded
HBM_BYTES_PER_RAY = 604          # 44 B in + 560 B out (full reference output dict)
METRIC = "rays/sec (64c+128f samples)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"burst": float(p["bf16_tflops"]), "sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "hbm_gbs": float(p["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def camera_rays_numpy(cam_index):
    """Synthetic camera `cam_index` (CAM_FRONT-like intrinsics, yawed by 60 deg per camera) + the workload helpers
    (tools/synth.py: plain numpy data generation; the oracle is imported by the CPU-baseline legs only)."""
    from tools import synth
    return synth.camera(cam_index), synth


def make_networks(dev):
    import torch
    from snerf_b200 import NeRF
    from tools import synth               # deterministic synthetic weights (seeded numpy)
    nets, params = [], []
    for seed in (20, 21):
        p = synth.nerf_params(seed, trunk_gain=1.5, sigma_bias=1.0)
        m = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
        nets.append(m.to(dev))
        params.append(p)
    return nets, params


def pick_cpu_threads(params, rb_sample):
    """Thread count at which the CPU port runs fastest on this host (more is not always faster on a
    many-core box); the chosen count is what `cores` reports."""
    from oracle import snerf_oracle as O
    best, best_v = 1, 0.0
    n = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, 128, n) if c <= n})
    for c in cands:
        O.set_backend("torch", threads=c)
        O.render_rays(rb_sample[:64], params[0], params[1], NC, NF)
        t0 = time.perf_counter()
        O.render_rays(rb_sample[:512], params[0], params[1], NC, NF)
        v = 512 / (time.perf_counter() - t0)
        if v > best_v:
            best, best_v = c, v
    return best


def cpu_port_rays_per_s(params, rb_sample, repeats=1):
    """The reference's CPU path as restated by the oracle (torch-CPU encode/MLP, best thread count)."""
    from oracle import snerf_oracle as O
    threads = pick_cpu_threads(params, rb_sample)
    O.set_backend("torch", threads=threads)
    O.render_rays(rb_sample[:256], params[0], params[1], NC, NF)  # warm the thread pools
    t0 = time.perf_counter()
    for _ in range(repeats):
        O.render_rays(rb_sample, params[0], params[1], NC, NF)
    dt = time.perf_counter() - t0
    O.set_backend("numpy")
    return rb_sample.shape[0] * repeats / dt, threads


def parity_numbers(got, ex, ref, inter, rb):
    """rgb L1 (mean abs), max-rel on rgb / depth / coarse weights (relative to |ref| + 1 % of the
    tensor's rms, the metric of tests/conftest.err_metric) and the end-to-end `inds` mismatch rate: the inverse-CDF bin each
    resampled depth fell into (recovered from the kernel's z_samples against the coarse mid-points) vs the oracle's
    `below = max(0, inds - 1)` (run_nerf_helpers.py:363-365)."""
    def max_rel(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-2 * (np.sqrt(np.mean(b * b)) + 1e-30))))
    z = ref["z_vals_map"]
    bins = 0.5 * (z[:, 1:] + z[:, :-1])
    mine = np.stack([np.searchsorted(bins[r], ex["z_samples"][r], side="right") for r in range(bins.shape[0])]) - 1
    mine = np.clip(mine, 0, bins.shape[1] - 2)
    theirs = np.clip(np.maximum(inter["inds"] - 1, 0), 0, bins.shape[1] - 2)
    return {"rgb_l1": float(np.mean(np.abs(got["rgb_map"] - ref["rgb_map"]))),
            "rgb_max_rel": max_rel(got["rgb_map"], ref["rgb_map"]),
            "depth_max_rel": max_rel(got["depth_map"], ref["depth_map"]),
            "weights_max_rel": max_rel(got["weights"], ref["weights"]),
            "depth_rel_l1": float(np.mean(np.abs(got["depth_map"] - ref["depth_map"])) / np.mean(np.abs(ref["depth_map"]))),
            "inds_mismatch_rate": float(np.mean(mine != theirs)), "rays": int(z.shape[0])}


def reference_cpu(params, rb_sample, threads, repeats=1):
    """The UNMODIFIED reference's render_rays (s-nerf/model/render.py:281-409, byte-compiled into oracle/_ref by
    oracle/build_ref_python.py) on the host cores: (rays/s, rgb_map of the sample), or None when it is not staged."""
    from oracle import build_ref_python
    mods = build_ref_python.load()
    if mods is None:
        return None
    import torch
    ref_render, ref_helpers = mods
    torch.set_num_threads(threads)
    nets = []
    for p in params:
        net = ref_helpers.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items() if not k.startswith("_")})
        nets.append(net.eval())
    embed_fn, _ = ref_helpers.get_embedder(10, 0)
    embeddirs_fn, _ = ref_helpers.get_embedder(4, 0)
    qfn = lambda inputs, viewdirs, network_fn: ref_helpers.run_network(inputs, viewdirs, network_fn, embed_fn=embed_fn,
                                                                      embeddirs_fn=embeddirs_fn, netchunk=1 << 16)
    rb = torch.from_numpy(rb_sample)

    def once(r):
        with torch.no_grad():      # eval path of the reference: chunks of 32768 rays (batchify_rays, render.py:8-19)
            return ref_render.batchify_rays(r, 1024 * 32, network_fn=nets[0], network_query_fn=qfn, N_samples=NC, N_importance=NF,
                                            network_fine=nets[1], perturb=0., raw_noise_std=0., white_bkgd=False, lindisp=False)
    once(rb[:256])
    t0 = time.perf_counter()
    for _ in range(repeats):
        out = once(rb)
    dt = time.perf_counter() - t0
    return rb.shape[0] * repeats / dt, out["rgb_map"].numpy()


TRAIN_RAYS = 512                 # rays per GPU per training step (config 3: 4096 rays over 8 GPUs)
# per sample: forward 593,408 MAC + dX 557,696 MAC (no input gradient) + dW 593,408 MAC; x2 FLOP x256 samples
TRAIN_FLOP_PER_RAY = (593408 + 557696 + 593408) * 2 * 256


def config3_loss(out, tgt, tdisp, conf):
    """Config-3 objective in torch, as the reference composes it (train.py:149-209): RgbLoss on the
    fine (+ coarse) colours + depth_lambda * calc_depth_loss = mean over rays with a LiDAR return of confidence *
    (|disp - t| + coarse_depth_mult * |disp_coarse - t|) (loss_factory.py:5-37, confidence.py:211-226), the L1 taken in
    disparity: `tdisp` is the LiDAR target as inverse depth (0 = no return) and the renderer's own `disp_map` / `disp0`
    (1 / max(1e-10, depth / acc), run_nerf_helpers.py:417) stand for 1 / pred so empty rays stay finite.  Used on the
    CPU arm; the GPU arm evaluates the same expression with the fused kernel (snerf_b200.losses.RgbDepthLoss)."""
    m = tdisp != 0
    depth_loss = (conf[m] * ((out["disp_map"][m] - tdisp[m]).abs() + COARSE_DEPTH_MULT * (out["disp0"][m] - tdisp[m]).abs())).mean()
    return ((out["rgb_map"] - tgt) ** 2).mean() + ((out["rgb0"] - tgt) ** 2).mean() + DEPTH_LAMBDA * depth_loss


DEPTH_LAMBDA, COARSE_DEPTH_MULT = 0.1, 0.2


def train_arm(dev, rank, world, steps, warmup, qfn, barrier, max_over_ranks):
    """BASELINE configs[2]: one training iteration of the reference (train.py:110-221) = fused forward (activations saved)
    + fused RgbDepthLoss + backward kernels + ONE in-place all-reduce of the flat 4.77 MB gradient buffer + one-kernel Adam,
    TRAIN_RAYS rays per GPU, perturb=1, raw_noise_std=1.  Timed as the product runs it: the whole iteration captured in one
    CUDA graph (snerf_b200.optim.GraphedTrainStep) and replayed per step; the same iteration launched eagerly from Python
    is reported beside it."""
    import torch
    import snerf_b200
    from snerf_b200 import render_rays
    from snerf_b200.losses import RgbDepthLoss
    from snerf_b200.optim import FlatAdam, GraphedTrainStep
    (net_c, net_f), params = make_networks(dev)
    total = steps + warmup
    rs = np.random.RandomState(1000 + rank)          # every rank draws its own rays
    c2w, O = camera_rays_numpy(rank)
    o_np, d_np = O.pinhole_rays(H, W, FOCAL, c2w, [CX, CY])
    batches = []
    for _ in range(total):
        idx = rs.choice(H * W, TRAIN_RAYS, replace=False)
        rb = O.ray_batch(o_np.reshape(-1, 3)[idx], d_np.reshape(-1, 3)[idx], NEAR, FAR)
        dep = (1.0 / rs.uniform(2, 100, TRAIN_RAYS)) * (rs.rand(TRAIN_RAYS) > 0.3)      # LiDAR target as disparity, 30 % missing
        # one flat buffer per step, plane by plane: [ray batch N x 11 | target rgb N x 3 | depth N | confidence N] -- one H2D copy,
        # and every plane is a contiguous view (column slices of an [N, 16] matrix would cost four strided-copy launches a step)
        b = np.concatenate([rb.ravel(), rs.rand(TRAIN_RAYS, 3).ravel(), dep, rs.rand(TRAIN_RAYS)]).astype(np.float32)
        batches.append(torch.from_numpy(b).pin_memory())
    resident = [b.to(dev) for b in batches]
    loss_host = torch.zeros(1).pin_memory()

    snerf_b200.set_mode("fp32")
    precision = os.environ.get("SNERF_BENCH_TRAIN_PRECISION", "bf16")
    snerf_b200.set_train_precision(precision)
    # one kernel forward + one backward; L1 between disparities (targets are stored as inverse depth, 0 = no return)
    criterion = RgbDepthLoss(DEPTH_LAMBDA, COARSE_DEPTH_MULT, disparity_depth=False, rgb0_weight=1.0)
    opt = FlatAdam([net_c, net_f], lr=5e-4, betas=(0.9, 0.999))     # parameters and gradients in one flat buffer each

    def loss_of(b):
        n = TRAIN_RAYS
        rb, tgt, dep, conf = b[:11 * n].view(n, 11), b[11 * n:14 * n].view(n, 3), b[14 * n:15 * n], b[15 * n:16 * n]
        out = render_rays(rb, net_c, qfn, NC, N_importance=NF, network_fine=net_f, perturb=1.0, raw_noise_std=1.0)
        return criterion(out["rgb_map"], tgt, out["disp_map"], out["disp0"], dep, conf, rgb_coarse=out["rgb0"])

    @torch.enable_grad()
    def eager_step(b):
        opt.grads.zero()
        loss = loss_of(b)
        loss.backward()
        opt.grads.all_reduce(average=True)
        opt.step()
        return loss.detach()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = {}
    for _ in range(warmup + 5):                       # the first steps also warm the caching allocator / NCCL
        eager_step(resident[0])
    barrier()
    e0.record()
    for i in range(warmup, total):
        loss = eager_step(resident[i])
    e1.record()
    barrier()
    res["eager"] = max_over_ranks(e0.elapsed_time(e1))

    graph_err = None
    try:
        gstep = GraphedTrainStep(resident[0], loss_of, opt)
    except Exception as e:                            # reported; the eager numbers stand in
        graph_err, gstep = repr(e)[:200], eager_step
    for arm in ("resident", "e2e"):
        for i in range(warmup):
            gstep(resident[i] if arm == "resident" else batches[i].to(dev, non_blocking=True))
        barrier()
        e0.record()
        for i in range(warmup, total):
            if arm == "resident":
                loss = gstep(resident[i])
            else:
                loss = gstep(batches[i].to(dev, non_blocking=True))
                loss_host.copy_(loss.reshape(1), non_blocking=True)
        e1.record()
        barrier()
        res[arm] = max_over_ranks(e0.elapsed_time(e1))
    loss_v = float(loss)
    # the collective alone: the in-place all-reduce of the flat 4.77 MB gradient buffer (0 on one GPU)
    ar_ms = 0.0
    if world > 1:
        for _ in range(3):
            opt.grads.all_reduce(average=True)
        barrier()
        e0.record()
        for _ in range(20):
            opt.grads.all_reduce(average=True)
        e1.record()
        barrier()
        ar_ms = max_over_ranks(e0.elapsed_time(e1)) / 20
    n_grad = int(opt.grads.flat.numel())
    opt.grads.release()
    snerf_b200.set_train_precision("fp32")
    v = world * TRAIN_RAYS * steps / (res["resident"] * 1e-3)
    ve = world * TRAIN_RAYS * steps / (res["e2e"] * 1e-3)
    arith = {"bf16": "bf16 operands and stores / f32 accumulate (tcgen05), f32 gradients and Adam",
             "tf32": "tf32 operands / f32 accumulate+storage (tcgen05)", "fp32": "f32 (FFMA)"}[precision]
    return {"metric": "train rays/s (fwd + loss + bwd + grad all-reduce + Adam)", "value": v, "unit": "rays/s",
            "ms_per_step": res["resident"] / steps, "rays_per_gpu_step": TRAIN_RAYS, "allreduce_ms": ar_ms,
            "gradient_bytes": n_grad * 4, "dtype": arith,
            "graph": {"captured": graph_err is None, "error": graph_err, "ms_per_step": res["resident"] / steps},
            "eager": {"ms_per_step": res["eager"] / steps, "value": world * TRAIN_RAYS * steps / (res["eager"] * 1e-3),
                      "note": "the same iteration launched kernel by kernel from Python (host-bound)"},
            "tflops_per_gpu": v / world * TRAIN_FLOP_PER_RAY / 1e12, "flop_per_ray": TRAIN_FLOP_PER_RAY,
            "e2e": {"value": ve, "unit": "rays/s", "ms_per_step": res["e2e"] / steps,
                    "h2d_bytes_per_step": TRAIN_RAYS * 16 * 4, "d2h_bytes_per_step": 4},
            "gpu_launches_per_step": ("1 fused forward (activation + relu' bit stores) + 1 composite-bwd + 1 fused dX chain + 1 grouped "
                                      "weight-gradient GEMM" if precision == "bf16" else
                                      "layer-batched GEMM chain" if precision == "tf32" else "1 fused forward + 4 backward kernels")
                                     + "; plus weight re-packing (3 launches per network), fused loss (1 + 1), RNG, Adam (2)",
            "collective": "one in-place all-reduce of 1,191,688 fp32 gradients per step" if world > 1 else "none (1 GPU)",
            "final_loss": loss_v, "config": "configs[2]: 512 rays/GPU/step, perturb=1, raw_noise_std=1, rgb MSE (fine + coarse) + 0.1 x masked, confidence-weighted depth L1 in disparity (coarse_depth_mult 0.2)"}, params


def frame6_arm(dev, rank, world, kw, barrier, max_over_ranks):
    """BASELINE configs[4]: one full 6-camera 1600x900 frame = 8.64 M rays, the 900 rows
    of EVERY camera split into contiguous blocks over the ranks (strong scaling; no data-path collective: each rank owns
    its rows of the six images).  Rays come from the library's get_rays kernel; the timed region is the six launches."""
    import torch
    from snerf_b200 import get_rays, render_rays
    from snerf_b200.parallel import shard_range
    a, b = shard_range(H, rank, world)
    batches = []
    for cam in range(6):
        c2w, _ = camera_rays_numpy(cam)
        ro, rd = get_rays(H, W, FOCAL, torch.from_numpy(c2w), ori_points=[CX, CY], device=dev)
        ro, rd = ro[a:b].reshape(-1, 3), rd[a:b].reshape(-1, 3)
        ones = torch.ones_like(rd[:, :1])
        batches.append(torch.cat([ro, rd, NEAR * ones, FAR * ones, rd / torch.norm(rd, dim=-1, keepdim=True)], -1).contiguous())
    out = render_rays(batches[0], **kw)          # warm-up
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for rb in batches:
        out = render_rays(rb, **kw)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    finite = bool(torch.isfinite(out["rgb_map"]).all().item())
    return {"workload": f"configs[4]: 6 cameras x {H}x{W} = {6 * H * W} rays per frame, rows [{a}:{b}) of every camera on rank 0 "
                        f"(contiguous row blocks over {world} rank(s)), NeRF 8x256 coarse+fine, 64c+128f, eval",
            "scaling": "strong", "rays_per_frame": 6 * H * W, "s_per_frame": ms * 1e-3, "value": 6 * H * W / (ms * 1e-3),
            "unit": "rays/s", "launches": 6, "collective": "none", "outputs_finite": finite}


def cpu_train_rays_per_s(params, threads, n_rays=128):
    """The reference's training step (torch-CPU autograd over the eager ops, via the differentiable oracle)."""
    import torch
    from oracle import snerf_oracle as O, snerf_oracle_grad as OG
    torch.set_num_threads(threads)
    rs = np.random.RandomState(5)
    c2w, synth = camera_rays_numpy(0)
    o_np, d_np = synth.pinhole_rays(H, W, FOCAL, c2w, [CX, CY])
    idx = rs.choice(H * W, n_rays, replace=False)
    rb = synth.ray_batch(o_np.reshape(-1, 3)[idx], d_np.reshape(-1, 3)[idx], NEAR, FAR)
    tgt, dep, conf = (torch.from_numpy(rs.rand(n_rays, 3).astype(np.float32)),
                      torch.from_numpy(((1.0 / rs.uniform(2, 100, n_rays)) * (rs.rand(n_rays) > 0.3)).astype(np.float32)),
                      torch.from_numpy(rs.rand(n_rays).astype(np.float32)))
    Pc, Pf = OG.params_to_torch(params[0]), OG.params_to_torch(params[1])
    dt = None
    for it in range(2):
        t0 = time.perf_counter()
        with torch.enable_grad():
            out = OG.render_rays(rb, Pc, Pf, NC, NF, t_rand=rs.rand(n_rays, NC).astype(np.float32),
                                 u=rs.rand(n_rays, NF).astype(np.float32), noise0=rs.rand(n_rays, NC).astype(np.float32),
                                 noise1=rs.rand(n_rays, NC + NF).astype(np.float32))
            config3_loss(out, tgt, dep, conf).backward()
        dt = time.perf_counter() - t0
    return n_rays / dt


def run_reference_arm(args):
    """--impl reference: the CPU port of the reference path on this box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import snerf_oracle as O          # this arm IS the CPU port of the reference path
    c2w, synth = camera_rays_numpy(0)
    o, d = synth.pinhole_rays(H, W, FOCAL, c2w, [CX, CY])
    idx = np.random.RandomState(0).choice(H * W, args.cpu_rays, replace=False)
    rb = synth.ray_batch(o.reshape(-1, 3)[idx], d.reshape(-1, 3)[idx], NEAR, FAR)
    params = [synth.nerf_params(s, trunk_gain=1.5, sigma_bias=1.0) for s in (20, 21)]
    threads = pick_cpu_threads(params, rb)
    kind = "port"
    ref = reference_cpu(params, rb, threads, repeats=args.steps)     # the unmodified reference, when staged (warm-up untimed)
    if ref is not None:
        kind = "reference"
        v = ref[0]
        dt = rb.shape[0] * args.steps / v
    else:
        O.set_backend("torch", threads=threads)
        for _ in range(args.warmup):
            O.render_rays(rb[:512], params[0], params[1], NC, NF)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.render_rays(rb, params[0], params[1], NC, NF)
        dt = time.perf_counter() - t0
        v = rb.shape[0] * args.steps / dt
    cores = threads
    what = ("the UNMODIFIED reference render_rays (byte-compiled s-nerf/model/render.py, oracle/_ref), torch-CPU" if kind == "reference"
            else "oracle port (numpy + torch-CPU encode/MLP on all host threads)")
    sample = f"host has {os.cpu_count()} logical cores, fastest thread count {threads} used; {args.cpu_rays} rays of camera 0 per step x {args.steps} steps, {what}, fp32"
    grid = None
    if not args.no_grid:
        # the reference's only native kernel family near this path: its own gridencoder.cu (oracle/_ref), on the GPU
        try:
            import torch
            if torch.cuda.is_available():
                from tools import grid_bench
                grid = grid_bench.run("reference", torch.device("cuda", 0))
        except Exception as e:  # the CPU arm above must still be reported
            grid = {"unavailable": repr(e)[:200]}
    print(json.dumps({
        **({"grid": grid} if grid is not None else {}),
        "impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: 1600x900 pinhole camera, NeRF 8x256 x2, 64c+128f, bounded CPU sample"},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp16", "fp16x3", "fp32"])
    ap.add_argument("--cpu-rays", type=int, default=4096, help="rays in the bounded CPU-baseline sample")
    ap.add_argument("--rays", type=int, default=H * W, help="rays per step per GPU (default: full 1600x900 image)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the config-3 training sub-benchmark")
    ap.add_argument("--no-parity-mode", action="store_true", help="skip the fp16x3 (fp32-class) arm")
    ap.add_argument("--train-steps", type=int, default=20)
    ap.add_argument("--no-frame6", action="store_true", help="skip the 6-camera full-frame (configs[4], strong scaling) sub-benchmark")
    ap.add_argument("--no-grid", action="store_true", help="skip the config-4 hash-grid encoder sub-benchmark")
    ap.add_argument("--no-mip", action="store_true", help="skip the mip-NeRF path (section 8 row f-2(i)) sub-benchmark")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import snerf_b200
    from snerf_b200 import get_rays, make_query_fn, render_rays
    from snerf_b200.render import render

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA sm_100 device; snerf_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    snerf_b200.set_mode(args.mode)
    torch.set_grad_enabled(False)     # rendering arms; train_arm() switches autograd back on for its steps
    (net_c, net_f), params = make_networks(dev)
    qfn, _, _ = make_query_fn()
    kw = dict(network_fn=net_c, network_query_fn=qfn, N_samples=NC, N_importance=NF, network_fine=net_f,
              perturb=0., raw_noise_std=0., white_bkgd=False, lindisp=False)

    # this rank's camera, generated on the device by the library's own get_rays kernel
    c2w, O = camera_rays_numpy(rank)
    rays_o, rays_d = get_rays(H, W, FOCAL, torch.from_numpy(c2w), ori_points=[CX, CY], device=dev)
    rays_o, rays_d = rays_o.reshape(-1, 3)[:args.rays].contiguous(), rays_d.reshape(-1, 3)[:args.rays].contiguous()
    n_rays = rays_o.shape[0]
    vd = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    ones = torch.ones_like(rays_d[:, :1])
    ray_batch = torch.cat([rays_o, rays_d, NEAR * ones, FAR * ones, vd], -1).contiguous()  # [N, 11], HBM resident

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident arm: `value` ----------------
    for _ in range(max(args.warmup, 3)):
        out = render_rays(ray_batch, **kw)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = render_rays(ray_batch, **kw)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    value = world * n_rays * args.steps / (ms * 1e-3)
    finite = bool(torch.isfinite(out["rgb_map"]).all().item())

    # ---------------- end-to-end arm: public API from pinned host buffers ----------------
    h_o, h_d = rays_o.cpu().pin_memory(), rays_d.cpu().pin_memory()
    h_out = {k: torch.empty(s, dtype=torch.float32).pin_memory() for k, s in
             (("rgb", (n_rays, 3)), ("disp", (n_rays,)), ("acc", (n_rays,)), ("depth", (n_rays,)))}
    h2d = h_o.numel() * 4 + h_d.numel() * 4
    d2h = sum(t.numel() * 4 for t in h_out.values())

    def e2e_step():
        o_d, d_d = h_o.to(dev, non_blocking=True), h_d.to(dev, non_blocking=True)
        rgb, disp, acc, depth, _ = render(H, W, FOCAL, chunk=None, rays=(o_d, d_d), ndc=False, near=NEAR, far=FAR,
                                          use_viewdirs=True, **kw)
        h_out["rgb"].copy_(rgb, non_blocking=True); h_out["disp"].copy_(disp, non_blocking=True)
        h_out["acc"].copy_(acc, non_blocking=True); h_out["depth"].copy_(depth, non_blocking=True)

    for _ in range(2):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * n_rays * args.steps / (ms_e2e * 1e-3)

    # ---------------- fp32-class tensor-core mode (fp16x3) on the same workload: the mode that meets the 1e-4 parity bar
    parity_mode = None
    if args.mode != "fp16x3" and not args.no_parity_mode:
        snerf_b200.set_mode("fp16x3")
        for _ in range(2):
            render_rays(ray_batch, **kw)
        barrier()
        e0.record()
        for _ in range(2):
            out3 = render_rays(ray_batch, **kw)
        e1.record()
        barrier()
        ms3 = max_over_ranks(e0.elapsed_time(e1))
        parity_mode = {"mode": "fp16x3", "value": world * n_rays * 2 / (ms3 * 1e-3), "unit": "rays/s", "steps": 2,
                       "ms_per_step": ms3 / 2,
                       "mma_tflops_per_gpu": n_rays * 2 / (ms3 * 1e-3) * 3 * FLOP_PER_RAY / 1e12,
                       "note": "same kernel, every operand split into fp16 hi + lo, 3 tcgen05 passes per k-block (3x the MMA "
                               "work): rgb/depth/weights within 1e-4 of the fp32 reference (tests/test_gpu_parity.py::"
                               "test_fused_fp16x3_config2)"}
        del out3
        snerf_b200.set_mode(args.mode)

    train = None
    if not args.no_train:
        try:
            train, _ = train_arm(dev, rank, world, args.train_steps, 3, qfn, barrier, max_over_ranks)
        except Exception as e:      # a reported sub-benchmark must not take the headline line down with it
            train = {"unavailable": repr(e)[:200]}
            torch.set_grad_enabled(False)
            snerf_b200.set_train_precision("fp32")
        snerf_b200.set_mode(args.mode)

    frame6 = None
    if not args.no_frame6:
        try:
            frame6 = frame6_arm(dev, rank, world, kw, barrier, max_over_ranks)
        except Exception as e:      # a reported sub-benchmark must not take the headline line down with it
            frame6 = {"unavailable": repr(e)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    per_gpu_tflops = value / world * FLOP_PER_RAY / 1e12
    roofline = {"bound": "tensor", "achieved": per_gpu_tflops, "peak": pk["sustained"], "unit": "TFLOP/s",
                "frac": per_gpu_tflops / pk["sustained"], "traffic": None,
                "peak_kind": "bf16 sustained, " + pk["source"], "frac_of_burst": per_gpu_tflops / pk["burst"],
                "hbm_side": {"achieved_gbs": value / world * HBM_BYTES_PER_RAY / 1e9, "peak_gbs": pk["hbm_gbs"],
                             "note": "ray I/O is 604 B/ray: HBM is not the limiter"}}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            tj = json.load(open(prof))
            roofline["traffic"] = tj.get("dram_bytes_per_launch")
            roofline["traffic_source"] = "profiles/traffic.json: ncu --set full capture of this kernel at this problem size, " + str(tj.get("captured", "round 1"))
        except Exception:
            pass

    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16": "bf16", "fp16": "f16", "fp16x3": "f16x3 (fp16 hi/lo split, fp32-class)", "fp32": "f32"}[args.mode], "data": "synthetic",
        "config": {"workload": f"configs[1]: {n_rays} rays/GPU/step (1600x900 pinhole camera per GPU), NeRF 8x256 coarse+fine, "
                               "64c+128f, eval (perturb=0), full reference output dict written to HBM",
                   "e2e_io": "H2D 24 B/ray (origins + directions from pinned host memory), D2H 24 B/ray (rgb, disp, acc, depth to pinned host "
                             "memory); the extras dict (z_vals, weights, rgb0, ... 536 B/ray) stays on the device as in the reference",
                   "l2": "per-step working set (63 MB rays + 806 MB outputs) exceeds the 126 MB L2; weights (2.4 MB) are meant to be L2-resident",
                   "mode": args.mode, "parallelism": f"ray-sharded x{world}, no collective"},
        "clocks": clocks, "gpu_launches": args.steps,
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "roofline": roofline, "outputs_finite": finite,
    }
    if world == 1 and not args.no_cpu_baseline:
        # ---- cpu_baseline leg: the only place of this arm that touches oracle/ (as the timed CPU port and as the checker)
        from oracle import snerf_oracle
        synth, O = O, snerf_oracle
        o_np, d_np = synth.pinhole_rays(H, W, FOCAL, c2w, [CX, CY])
        idx = np.random.RandomState(0).choice(H * W, args.cpu_rays, replace=False)
        rb = synth.ray_batch(o_np.reshape(-1, 3)[idx], d_np.reshape(-1, 3)[idx], NEAR, FAR)
        v, threads = cpu_port_rays_per_s(params, rb)
        kind, what = "port", "oracle port of the reference CPU path (numpy + torch-CPU encode/MLP on all host threads)"
        ref_cpu = reference_cpu(params, rb, threads)
        if ref_cpu is not None:
            line["cpu_port_rays_s"] = v
            v, kind = ref_cpu[0], "reference"
            what = "the UNMODIFIED reference render_rays (byte-compiled s-nerf/model/render.py from oracle/_ref), torch-CPU, chunk 32768"
        line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": threads, "kind": kind,
                                "sample": f"host has {os.cpu_count()} logical cores, fastest thread count {threads} used; {args.cpu_rays} rays of the same camera, {what}, fp32"}
        # parity of the timed configuration against the oracle on the same rays (rgb L1, max-rel on
        # rgb / depth / weights, inds mismatch rate), for the timed mode and for the fp32-class tensor-core mode
        sub = torch.from_numpy(rb[:1024]).to(dev)
        ref_all = O.render_rays(rb[:1024], params[0], params[1], NC, NF, return_intermediates=True)
        inter = ref_all.pop("_inter")

        def parity_of(mode_name):
            snerf_b200.set_mode(mode_name)
            o_ = render_rays(sub, _extras=True, **kw)
            ex_ = {k: v.cpu().numpy() for k, v in o_.pop("_extras").items()}
            snerf_b200.set_mode(args.mode)
            return parity_numbers({k: v.cpu().numpy() for k, v in o_.items()}, ex_, ref_all, inter, rb[:1024])

        parity = parity_of(args.mode)
        line["rgb_l1_vs_oracle"] = parity["rgb_l1"]
        line["parity_vs_oracle"] = parity
        if parity_mode is not None:
            parity_mode["parity_vs_oracle"] = parity_of("fp16x3")
            parity_mode["rgb_l1_vs_oracle"] = parity_mode["parity_vs_oracle"]["rgb_l1"]
            parity_mode["depth_rel_l1_vs_oracle"] = parity_mode["parity_vs_oracle"]["depth_rel_l1"]
        if train is not None and "value" in train:
            tv = cpu_train_rays_per_s(params, threads)
            train["cpu_baseline"] = {"value": tv, "unit": "rays/s", "cores": threads, "kind": "port",
                                     "sample": "128 rays, one fwd+bwd of the differentiable oracle (torch-CPU autograd), second of two runs"}
    if parity_mode is not None:
        line["parity_mode"] = parity_mode
    if train is not None:
        line["train"] = train
    if frame6 is not None:
        line["frame6"] = frame6
    if not args.no_grid:
        # BASELINE configs[3]: hash-grid encoder at zip-NeRF shapes (a parity-test configuration; reported, not the headline)
        try:
            from tools import grid_bench, stepfun_bench
            line["grid"] = grid_bench.run("ours", dev)
            line["grid"]["proposal_resample"] = stepfun_bench.run(dev)
        except Exception as e:      # reported, not the headline
            line["grid"] = {"unavailable": repr(e)[:200]}
    if not args.no_mip:
        try:   # the model train.py / eval.py run, at the shipped configuration
            from tools import mip_bench
            line["mip"] = mip_bench.run(dev)
        except Exception as e:      # reported, not the headline
            line["mip"] = {"unavailable": repr(e)[:200]}
    try:
        # the same frame through render() with the reference's default chunk (render.py:22-25:
        # 32768 rays per render_rays call, results concatenated) -- 44 launches + torch.cat per image instead of one launch
        def chunk_step():
            return render(H, W, FOCAL, chunk=1024 * 32, rays=(rays_o, rays_d), ndc=False, near=NEAR, far=FAR, use_viewdirs=True, **kw)
        chunk_step()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(2):
            chunk_step()
        e1.record()
        torch.cuda.synchronize()
        ms_c = e0.elapsed_time(e1)
        line["chunk32768"] = {"value": n_rays * 2 / (ms_c * 1e-3), "unit": "rays/s", "ms_per_image": ms_c / 2,
                              "launches_per_image": -(-n_rays // (1024 * 32)), "steps": 2,
                              "note": "render(chunk=32768): the reference's chunking (batchify_rays, render.py:8-19) on rank 0, rays resident"}
    except Exception as e:      # informational row, measured last: nothing it does can affect the numbers above
        line["chunk32768"] = {"unavailable": repr(e)[:200]}
    # ---- driver-kept summary of the sub-benchmarks (the driver keeps `config` whole): configs[2] training,
    # configs[4] 6-camera strong scaling, the fp32-class tensor-core rate and the parity numbers of SURVEY.md section 8(d)
    def pick(obj, *path):
        for k in path:
            if not isinstance(obj, dict) or k not in obj:
                return None
            obj = obj[k]
        return obj
    par, par3 = line.get("parity_vs_oracle"), pick(parity_mode, "parity_vs_oracle")
    line["config"]["sub"] = {
        "train_ms_per_step": pick(train, "ms_per_step"), "train_rays_s": pick(train, "value"),
        "train_e2e_rays_s": pick(train, "e2e", "value"), "train_allreduce_ms": pick(train, "allreduce_ms"),
        "train_arith": pick(train, "dtype"), "train_rays_per_gpu_step": pick(train, "rays_per_gpu_step"),
        "train_graph_captured": pick(train, "graph", "captured"), "train_eager_ms_per_step": pick(train, "eager", "ms_per_step"),
        "frame6_s": pick(frame6, "s_per_frame"), "frame6_rays_s": pick(frame6, "value"), "frame6_scaling": "strong",
        "fp16x3_rays_s": pick(parity_mode, "value"),
        "parity_" + args.mode: None if par is None else {k: par[k] for k in
                                 ("rgb_l1", "rgb_max_rel", "depth_max_rel", "weights_max_rel", "inds_mismatch_rate", "rays")},
        "parity_fp16x3": None if par3 is None else {k: par3[k] for k in
                                 ("rgb_l1", "rgb_max_rel", "depth_max_rel", "weights_max_rel", "inds_mismatch_rate", "rays")},
        "chunk32768_rays_s": pick(line.get("chunk32768"), "value"),
        "mip_rays_s": pick(line.get("mip"), "value"), "mip_tflops": pick(line.get("mip"), "tflops"),
    }
    # verbose sub-objects first, the contract keys last: the tail of stdout is what a truncating reader sees
    tail_keys = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                 "vs_baseline", "dtype", "data", "gpu_launches", "clocks", "cpu_baseline", "roofline", "e2e", "config"]
    line = {**{k: v for k, v in line.items() if k not in tail_keys}, **{k: line[k] for k in tail_keys if k in line}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
