#!/usr/bin/env python
"""Benchmark of the TEDM hot path on B200 (see BASELINE.json / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

Workload (config.workload = "tedm_seg_inference"): BASELINE.json configs[3] -- TEDM shared-weight
timestep-ensembled segmentation inference at config.py defaults: 1x128x128 images, S = 8 timesteps
[1,10,25,50,200,400,600,800], UNet dim 64 / mults (1,2,4,8), B = 16 images per GPU per step
(config.py:58).  One step = B images -> B masks = 8B UNet forwards + heads + ensemble.  Weights are
random-init (seeded), images synthetic.  Printed line: whole-job images/s (`value`, inputs resident
in HBM), the same through host buffers (`e2e`), the conv kernel's tensor roofline, and the CPU
baseline (the oracle port of the reference's path timed on this box's cores).
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STEPS_TEDM = [1, 10, 25, 50, 200, 400, 600, 800]
IMG = 128
GFLOP_UNET_FWD = 58.97          # per image, reference form
GFLOP_TEDM_REF = 505.0          # per image: 8 x (58.97 + 4.16)
GFLOP_TEDM_MIN = 436.1          # per image: 8 x (53.87 + 0.64)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.thr.join(timeout=2)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# synthetic model / data (identical on every arm)
# ------------------------------------------------------------------------------------------------
def synth_state(n_steps: int):
    from oracle import tedm_oracle as O          # shapes only (bench.py may use oracle/ for its CPU legs)
    from tests.golden.synth import synth_state_dict
    shapes = {**O.unet_param_shapes(prefix="diffusion_model.model."), **O.head_param_shapes(n_steps, True)}
    return synth_state_dict(shapes, 0)


def synth_batch(b: int, seed: int):
    import torch
    g = torch.Generator().manual_seed(1234 + seed)
    return torch.rand(b, 1, IMG, IMG, generator=g)


# ------------------------------------------------------------------------------------------------
# CPU legs: the oracle port of the reference's path
# ------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "oracle", "_ref")       # unmodified reference sources, vendored by oracle/make_ref.py


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "models", "datasetDM_model.py"))


_REF_MODEL = None


def reference_tedm_model():
    """The reference's own modules (oracle/_ref/models/*.py, unmodified) set up as its TEDM test script does
    (auxiliary/postprocessing/testing_shared_weights.py:55-75): DatasetDM + the shared 960-input head, eval mode."""
    global _REF_MODEL
    if _REF_MODEL is None:
        import torch
        from argparse import Namespace
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        from einops.layers.torch import Rearrange
        from models.datasetDM_model import DatasetDM as RefDatasetDM          # oracle/_ref
        from torch import nn
        S = len(STEPS_TEDM)
        m = RefDatasetDM(Namespace(normalize=True, saved_diffusion_model="/nonexistent", verbose=False, device="cpu",
                                   t_steps_to_save=STEPS_TEDM))
        m.classifier = nn.Sequential(Rearrange('b (step act) h w -> (b step) act h w', step=S), nn.Conv2d(960, 128, 1), nn.ReLU(),
                                     nn.BatchNorm2d(128), nn.Conv2d(128, 32, 1), nn.ReLU(), nn.BatchNorm2d(32), nn.Conv2d(32, 1, 1))
        missing = m.load_state_dict(synth_state(S), strict=False)
        assert not missing.unexpected_keys, missing.unexpected_keys
        _REF_MODEL = m.eval()
    return _REF_MODEL


def cpu_tedm_images_per_s(n_images: int, repeats: int = 1):
    """The reference's TEDM inference (model(x) -> sigmoid -> mean over steps -> > .5) on the host cores.  With oracle/_ref
    present it is the reference's OWN code (kind "reference"); otherwise the oracle port (kind "port").
    -> (images/s, seconds, kind)"""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    x0 = synth_batch(n_images, 99)
    S = len(STEPS_TEDM)
    best = None
    if reference_available():
        model, kind = reference_tedm_model(), "reference"
        with torch.no_grad():
            for _ in range(repeats):
                t0 = time.perf_counter()
                logits = model(x0)                                   # fresh noise per step inside (diffusion_model.py:193)
                prob = torch.sigmoid(logits).reshape(n_images, S, 1, IMG, IMG).mean(1)
                _mask = prob > 0.5
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
        return n_images / best, best, kind
    from oracle import tedm_oracle as O
    sd = synth_state(S)
    sd.update(O.schedule_tables())
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            noises = [torch.randn(n_images, 1, IMG, IMG) for _ in range(S)]
            O.tedm_segment(sd, x0, STEPS_TEDM, noises, shared=True)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return n_images / best, best, "port"


def gpu_eager_baselines(dev, n_images: int = 16):
    """the reference's path as plain PyTorch eager ops ON THE SAME GPU (cuDNN / cuBLAS kernels; the
    oracle restatement, since /root/reference is not on the box), fp32 and under autocast(bf16).  A reported baseline
    only -- never part of the product path."""
    import torch
    from oracle import tedm_oracle as O
    torch.backends.cudnn.benchmark = True
    sd = synth_state(len(STEPS_TEDM))
    sd.update(O.schedule_tables())
    sd = {k: v.to(dev) for k, v in sd.items()}
    x0 = synth_batch(n_images, 99).to(dev)
    noises = [torch.randn(n_images, 1, IMG, IMG, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + i))
              for i in range(len(STEPS_TEDM))]
    out = {"unit": "images/s", "sample": f"{n_images} images x 8 timesteps per pass (the batch of the headline run), 3 timed passes after 2 warm-ups",
           "kind": "port (oracle restatement as torch eager CUDA ops)"}
    for name, ctx in (("fp32", contextlib.nullcontext()), ("autocast_bf16", torch.autocast("cuda", dtype=torch.bfloat16))):
        try:
            with torch.no_grad(), ctx:
                for _ in range(2):
                    O.tedm_segment(sd, x0, STEPS_TEDM, noises, shared=True)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    O.tedm_segment(sd, x0, STEPS_TEDM, noises, shared=True)
                e1.record()
                torch.cuda.synchronize()
            out[name] = 3 * n_images / (e0.elapsed_time(e1) * 1e-3)
        except Exception as e:  # a baseline must never take the bench down
            out[name] = None
            out[name + "_error"] = f"{type(e).__name__}: {e}"[:200]
    return out


def sampler_leg(dev, batch: int = 64, steps: int = 20):
    """BASELINE configs[4]: DDPM ancestral sampling.  Times `steps` reverse steps (UNet forward + the one-kernel
    posterior update with the exact dynamic-threshold quantile) of a `batch`-image chain batch; a full chain is 1000."""
    import torch
    from argparse import Namespace
    from tedm_b200.models import DiffusionModel
    from tedm_b200.trainers.utils import GraphedSampler
    torch.manual_seed(7)
    m = DiffusionModel(Namespace(normalize=True)).to(dev).eval()
    gs = GraphedSampler(m, batch, 1, IMG)            # what sample_images / sample_plot_image run: one graph replay per step
    gs.x.normal_()
    for t in range(999, 996, -1):
        gs.step(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(996, 996 - steps, -1):
        img = gs.step(t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"workload": "ddpm_ancestral_sampling", "batch_per_gpu": batch, "timed_reverse_steps": steps, "ms_per_reverse_step": ms,
            "image_steps_per_s": batch / (ms * 1e-3), "images_per_s_full_1000_step_chain": batch / ms,
            "tflops_unet_fwd": batch * GFLOP_UNET_FWD / ms, "finite": bool(torch.isfinite(img).all())}


def head_train_leg(dev, batch: int = 16, steps: int = 5):
    """BASELINE configs[2]: LEDM (steps [50,150,250], 2880-input head) and TEDM (8 steps, shared 960-input head) head
    training on JSRT-shaped synthetic pairs: frozen-UNet features + head forward / BCE / backward / Adam per step, through
    the trainer's own loop body (tedm_b200/trainers/train_baseline.py)."""
    import torch
    from argparse import Namespace
    from tedm_b200.autograd import bce_with_logits_rows
    from tedm_b200.models import DatasetDM, tedm_classifier
    from tedm_b200.optim import FusedAdam
    out = {"workload": "ledm_tedm_head_training_step", "batch_per_gpu": batch, "unit": "images/s"}
    g = torch.Generator().manual_seed(11)
    x = torch.rand(batch, 1, IMG, IMG, generator=g).to(dev)
    y = (torch.rand(batch, 1, IMG, IMG, generator=g) > 0.5).float().to(dev)
    for name, t_steps, shared in (("LEDM", [50, 150, 250], False), ("TEDM", STEPS_TEDM, True)):
        torch.manual_seed(3)
        m = DatasetDM(Namespace(normalize=True, saved_diffusion_model="", t_steps_to_save=t_steps))
        if shared:
            m.classifier = tedm_classifier(len(t_steps))
        m = m.to(dev).train()
        m.diffusion_model.eval()
        opt = FusedAdam(m.classifier.parameters(), lr=1e-4)

        def step():
            opt.zero_grad()
            loss = bce_with_logits_rows(m(x), y).mean()
            loss.backward()
            opt.step()
            return loss

        for _ in range(5):
            step()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(max(steps, 5))]
        for a, b in evs:                   # per-step events, median: one-off allocator / GC hiccups do not count
            a.record()
            loss = step()
            b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
        out[name] = {"t_steps": len(t_steps), "ms_per_step": ms, "images_per_s": batch / (ms * 1e-3),
                     "unet_forwards_per_s": batch * len(t_steps) / (ms * 1e-3), "loss_last": float(loss)}
        del m, opt
        import gc
        gc.collect()
    torch.cuda.empty_cache()
    return out


def cpu_train_images_per_s(n_images: int):
    """The reference's training step (train_step + backward, trainers/train_CXR14.py:30-40) as the oracle port runs it
    on the host cores: fp32 torch autograd through oracle.ddpm_loss."""
    import torch
    from oracle import tedm_oracle as O
    from tests.golden.synth import synth_state_dict
    torch.set_num_threads(os.cpu_count() or 1)
    sd = {k: v.requires_grad_(True) for k, v in synth_state_dict(O.unet_param_shapes(prefix="model."), 0).items()}
    full = dict(sd)
    full.update(O.schedule_tables())
    x0 = synth_batch(n_images, 98)
    t = torch.randint(0, 1000, (n_images,), generator=torch.Generator().manual_seed(3))
    nz = torch.randn(n_images, 1, IMG, IMG, generator=torch.Generator().manual_seed(4))
    t0 = time.perf_counter()
    O.ddpm_loss(full, x0, t, nz).backward()
    dt = time.perf_counter() - t0
    return n_images / dt, dt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores, all threads, on OUR
    arm's config (B = 16 images x 8 timesteps per step).  The reference is unpackaged pure Python (nothing for pip to
    build); oracle/make_ref.py vendors its unmodified sources into oracle/_ref/ (git-ignored, shipped with the snapshot)
    and this arm runs THEM.  If oracle/_ref is absent the oracle port (pinned to the live reference by tests/golden) is
    timed instead and the line says kind "port".  A step takes ~10 s on 16 cores, so the number of timed steps is capped
    to keep the run within a few minutes; `steps` in the line is the count actually timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_img = args.batch
    if args.warmup:
        cpu_tedm_images_per_s(1)                        # one small pass: thread pool, allocator, module set-up
    t_all, n_steps, kind = 0.0, 0, "port"
    budget_s = float(os.environ.get("TEDM_BENCH_REF_BUDGET_S", "150"))
    for _ in range(args.steps):
        v, dt, kind = cpu_tedm_images_per_s(n_img)
        t_all += dt
        n_steps += 1
        if t_all + dt > budget_s:
            break
    args.steps = n_steps
    value = n_img * args.steps / t_all
    line = {"metric": "tedm_seg_images_per_s", "value": value, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": "tedm_seg_inference", "img_size": IMG, "t_steps": STEPS_TEDM, "unet_dim": 64,
                       "dim_mults": [1, 2, 4, 8], "batch_per_gpu_per_step": n_img, "global_batch": n_img},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": os.cpu_count(), "kind": kind,
                             "sample": f"{n_img} images x 8 timesteps per step, {args.steps} steps ("
                                       + ("the reference's own modules from oracle/_ref" if kind == "reference" else "oracle port")
                                       + ", torch CPU fp32, all host threads)"},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from argparse import Namespace
    from tedm_b200 import native as N
    from tedm_b200.models import DatasetDM, tedm_classifier

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device: tedm_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N.load()

    B, S = args.batch, len(STEPS_TEDM)
    model = DatasetDM(Namespace(normalize=True, saved_diffusion_model="", t_steps_to_save=STEPS_TEDM))
    model.classifier = tedm_classifier(S)
    model.load_state_dict(synth_state(S), strict=False)
    model = model.eval().to(dev)

    # inputs: a ring of distinct batches (activations per step ~ several GB >> 126 MB L2, so no L2 flush needed)
    n_ring = 4
    host = [synth_batch(B, rank * 100 + i).pin_memory() for i in range(n_ring)]
    resident = [h.to(dev) for h in host]
    noise = [torch.randn(B, 1, IMG, IMG, device=dev, generator=torch.Generator(device=dev).manual_seed(5 + i)) for i in range(n_ring)]
    mask_host = torch.empty(B, 1, IMG, IMG, dtype=torch.bool).pin_memory()

    # noise=None: every call draws its B x S noise images inside the (graph-replayed) call, where the reference draws them
    # (randn_like per step, diffusion_model.py:193) -- the Philox launch is part of the timed step
    def step_resident(i):
        return model.segment(resident[i % n_ring], None, graph=True)[0]

    def step_e2e(i):
        x = host[i % n_ring].to(dev, non_blocking=True)
        mask = model.segment(x, None, graph=True)[0]
        mask_host.copy_(mask, non_blocking=True)
        return mask

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for i in range(args.warmup):
        step_resident(i)
    l0, f0 = N.launches, N.conv_flops
    with ClockSampler(local) as clocks:
        ms = timed(step_resident, args.steps)
    launches = N.launches - l0
    conv_flops_step = (N.conv_flops - f0) / args.steps
    for i in range(min(args.warmup, 2)):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)

    value = world * B * args.steps / (ms * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (tcgen05 implicit-GEMM conv): CUDA events around every launch ----
    records = []

    @contextlib.contextmanager
    def conv_timer(flops, shape=None):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        yield
        b.record()
        records.append((flops, a, b, shape))

    elem_records = []

    @contextlib.contextmanager
    def elem_timer(name, nbytes):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        yield
        b.record()
        elem_records.append((name, nbytes, a, b))

    N.conv_timer = conv_timer
    N.elem_timer = elem_timer
    try:
        model.segment(resident[0], noise[0])          # eager: every conv / GroupNorm launch bracketed by its own event pair
    finally:
        N.conv_timer = None
        N.elem_timer = None
    torch.cuda.synchronize()
    gn_ms = sum(a.elapsed_time(b) for _, _, a, b in elem_records)
    gn_bytes = sum(nb for _, nb, _, _ in elem_records)
    conv_ms = sum(a.elapsed_time(b) for _, a, b, _ in records)
    conv_fl = sum(f for f, _, _, _ in records)
    if os.environ.get("TEDM_BENCH_CONV_TABLE") and rank == 0:
        agg = {}
        for f, a, b, shp in records:
            e = agg.setdefault(shp, [0, 0.0, 0.0])
            e[0] += 1
            e[1] += a.elapsed_time(b)
            e[2] += f
        with open(os.environ["TEDM_BENCH_CONV_TABLE"], "w") as fh:
            fh.write("mode B H W c0 c1 cout gn res | launches  ms_total  TFLOP/s  share_of_conv_time\n")
            for shp, (cnt, t_ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                fh.write(f"{shp} | {cnt:3d} {t_ms:9.3f} {fl / (t_ms * 1e-3) / 1e12:8.1f} {100 * t_ms / conv_ms:6.1f}%\n")
    pk = peaks()
    achieved = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    # DRAM traffic of the dominant conv launch (3x3 64->64 @128x128, 128 images): dram__bytes_read + write of ONE launch from
    # an `ncu --set full` capture of this very command, parsed into profiles/r02_conv_traffic.json by
    # scripts/conv_traffic.py (null when that file is absent -- never a literal)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_conv_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj
    roofline = {"bound": "tensor", "kernel": "conv_igemm_kernel + conv_ws4_kernel (all conv launches of one step)", "achieved": achieved,
                "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tflops_sustained"],
                "peak_source": pk["source"] + " (sustained cuBLAS bf16)",
                "traffic": traffic, "traffic_unit": "bytes/launch of the dominant launch (see traffic_source)",
                "traffic_source": traffic_src,
                "conv_launches_per_step": len(records), "conv_ms_per_step": conv_ms,
                "conv_share_of_step": conv_ms / (ms / args.steps) if ms > 0 else None,
                "conv_gflop_per_step": conv_fl / 1e9}

    gn_gbs = gn_bytes / (gn_ms * 1e-3) / 1e9 if gn_ms > 0 else 0.0
    roofline_hbm = {"bound": "hbm", "kernel": "gn_silu_kernel (all GroupNorm+SiLU launches of one step; the largest memory-bound kernel)",
                    "achieved": gn_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gn_gbs / pk["hbm_gbs"],
                    "peak_source": pk["source"] + " (copy bandwidth)", "algorithmic_bytes_per_step": gn_bytes,
                    "launches_per_step": len(elem_records), "ms_per_step": gn_ms,
                    "share_of_step": gn_ms / (ms / args.steps) if ms > 0 else None, "traffic": None}
    line = {"metric": "tedm_seg_images_per_s", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "tedm_seg_inference", "img_size": IMG, "t_steps": STEPS_TEDM, "unet_dim": 64,
                       "dim_mults": [1, 2, 4, 8], "batch_per_gpu_per_step": B, "global_batch": B * world,
                       "parallelism": f"dp{world} (image x timestep shards, no collective)",
                       "launch": "DatasetDM.segment(graph=True): the step's native launches replayed from one CUDA graph",
                       "l2": "working set per step (GBs of activations) >> 126 MB L2; 4 distinct input batches cycled"},
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": B * IMG * IMG * 4,
                    "d2h_bytes_per_step": B * IMG * IMG, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "tflops": {"unet_fwd_reference_form": value * GFLOP_TEDM_REF / 1e3, "minimal_form": value * GFLOP_TEDM_MIN / 1e3,
                       "executed_conv_gflop_per_image": conv_flops_step / B / 1e9,
                       "frac_of_sustained_peak_minimal_form": value * GFLOP_TEDM_MIN / 1e3 / (pk["tflops_sustained"] * world)},
            "roofline": roofline, "roofline_hbm": roofline_hbm, "clocks": clocks.summary()}
    # ---- the same workload in the fp32 precision mode (the reference's default arithmetic; 1e-4 parity, tests/test_gpu_fp32.py)
    if rank == 0 and world == 1 and not args.no_fp32:
        model.set_precision("fp32")
        try:
            for i in range(2):
                step_resident(i)
            n32 = max(2, args.steps // 2)
            ms32 = timed(step_resident, n32)
            line["fp32_mode"] = {"dtype": "f32", "value": B * n32 / (ms32 * 1e-3), "unit": "images/s", "ms_per_step": ms32 / n32,
                                 "steps": n32, "arithmetic": "fp32 storage; convs as split-bf16 (hi*hi + lo*hi + hi*lo) on tcgen05, "
                                 "fp32 accumulation; GroupNorm / LayerNorm / attention / head in fp32"}
        finally:
            model.set_precision("bf16")
    # ---- second half of BASELINE.json's metric: the DDPM pre-training step (UNet fwd+bwd TFLOP/s vs bf16 peak) ----
    del model, resident, noise
    torch.cuda.empty_cache()
    if not args.no_train:
        line["train"] = train_leg(args, dev, world, rank, pk, barrier)
        if rank == 0 and world == 1:
            line["sampler"] = sampler_leg(dev)
            line["head_train"] = head_train_leg(dev)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["gpu_eager_baseline"] = gpu_eager_baselines(dev)
            torch.cuda.empty_cache()
            if "train" in line:
                v, dt = cpu_train_images_per_s(8)
                line["train"]["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                                 "sample": f"one fwd+bwd on 8 images ({dt:.1f} s), oracle port, torch CPU fp32 autograd"}
            cpu_tedm_images_per_s(1)                                   # warm-up (thread pool, allocator)
            v, dt, kind = cpu_tedm_images_per_s(16)
            line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": os.cpu_count(), "kind": kind,
                                    "sample": f"16 images x 8 timesteps, one pass ({dt:.1f} s), "
                                              + ("the reference's own modules (oracle/_ref)" if kind == "reference" else "oracle port")
                                              + " on torch CPU fp32"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


GFLOP_TRAIN = 176.9             # per image: 3 x 58.97


def train_leg(args, dev, world, rank, pk, barrier):
    """BASELINE configs[1]: DDPM backbone pre-training, data-parallel.  One step = q_sample + UNet forward + L1/p2 loss +
    UNet backward + ONE gradient all-reduce (world > 1) + Adam, replayed from CUDA graphs (tedm_b200/train.py)."""
    import torch
    import torch.distributed as dist
    from argparse import Namespace
    from tedm_b200.models import DiffusionModel
    from tedm_b200.optim import FusedAdam
    from tedm_b200.train import GraphedTrainStep

    out = {"workload": "ddpm_pretraining_step", "unit": "images/s", "gflop_per_image_fwd_bwd": GFLOP_TRAIN,
           "step": "q_sample + UNet fwd + L1/p2 loss + UNet bwd + grad all-reduce (world>1) + fused Adam; CUDA-graph replay",
           "parallelism": f"dp{world} (one flat-arena NCCL all-reduce per step)" if world > 1 else "dp1"}
    torch.manual_seed(1234 + rank)
    model = DiffusionModel(Namespace(normalize=True)).to(dev).train()
    if world > 1:                       # replicas start from rank 0's weights
        for p in model.parameters():
            dist.broadcast(p.data, src=0)
    opt = FusedAdam(model.parameters(), lr=1e-4)
    for B in args.train_batches:
        host = [synth_batch(B, 500 + rank * 100 + i).pin_memory() for i in range(4)]
        resident = [h.to(dev) for h in host]
        step = GraphedTrainStep(model, opt, resident[0], warmup=max(3, args.warmup))

        def timed(fn, steps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                fn(i)
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = t.item()
            return ms

        losses = []

        def e2e(i):
            x = host[i % 4].to(dev, non_blocking=True)
            losses.append(float(step(x)))              # loss read back to the host every step

        for i in range(3):
            step(resident[i % 4])
        ms = timed(lambda i: step(resident[i % 4]), args.steps)
        ms_e2e = timed(e2e, args.steps)
        ips = world * B * args.steps / (ms * 1e-3)
        out[f"batch{B}"] = {
            "batch_per_gpu": B, "global_batch": B * world, "ms_per_step": ms / args.steps, "images_per_s": ips,
            "tflops_fwd_bwd": ips * GFLOP_TRAIN / 1e3, "frac_of_sustained_bf16_peak": ips * GFLOP_TRAIN / 1e3 / (pk["tflops_sustained"] * world),
            "e2e_images_per_s": world * B * args.steps / (ms_e2e * 1e-3), "h2d_bytes_per_step": B * IMG * IMG * 4,
            "d2h_bytes_per_step": 4, "native_calls_per_step": step.native_calls_per_step, "loss_last": losses[-1]}
        step.close()
        del step
        torch.cuda.empty_cache()
    best = max((k for k in out if k.startswith("batch")), key=lambda k: out[k]["tflops_fwd_bwd"])
    out["value"] = out[best]["images_per_s"]
    out["tflops_fwd_bwd"] = out[best]["tflops_fwd_bwd"]
    out["frac_of_sustained_bf16_peak"] = out[best]["frac_of_sustained_bf16_peak"]
    out["batch_per_gpu"] = out[best]["batch_per_gpu"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU per step (config.py:58 default 16)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the DDPM training-step leg")
    ap.add_argument("--no-fp32", action="store_true", help="skip the fp32-precision-mode leg")
    ap.add_argument("--train-batches", type=int, nargs="+", default=[16, 64, 128],
                    help="per-GPU batch sizes of the training leg (config.py:58 default is 16)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
