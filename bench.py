#!/usr/bin/env python
"""bench.py -- sliding-window 3D U-Net inference throughput (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (N = 1): BASELINE.json configs[1] -- msseg2 ModularUNet(2 -> 2, filters [40,40,80,80,120,120], depth 6,
residual blocks, BlurConv3d / BlurConvTranspose3d), synthetic 2-channel 256 x 256 x 192 volume, patch 96^3,
overlap 48, padding 'edge' (144 patches), bf16 tensor-core path.  One *step* = one whole volume through the hot
path: patch extraction -> 144 network forwards -> overlap-add -> divide/crop/argmax -> uint8 labels -> confusion
counts.  `value` times it with the volume resident in HBM; `e2e` times the reference-facing
PatchPredict.predict with a pinned HOST volume in and HOST probabilities out.  For N > 1 every rank processes its
own volume of the cohort (no data-path collective; weak scaling), timed as max over ranks.

`--impl reference` times the reference's CPU implementation of the same path (the oracle port -- the
reference is pure Python + ATen and cannot be imported on the GPU box) on the host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "segmentation-pipeline_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "3D U-Net sliding-window inference Mvoxels/s"
UNIT = "Mvoxel/s"
VOLUME = (2, 256, 256, 192)
PATCH, OVERLAP, PADDING = 96, 48, "edge"
FILTERS = [40, 40, 80, 80, 120, 120]
PATCH_BATCH = int(os.environ.get("B200SEG_PATCH_BATCH", "48"))
FLOP_PER_PATCH = 668.74e9          # 2 * MACs of the reference layer list at 96^3
N_PATCHES = 144
WORKLOAD = ("config2: msseg2 ModularUNet 2->2 filters [40,40,80,80,120,120] depth 6 residual blur-down/up, "
            "synthetic 2x256x256x192 volume, patch 96^3 overlap 48 padding edge (144 patches), bf16")


# ------------------------------------------------------------------------------------------------- synthetic inputs
def synthetic_volume(seed: int, with_mask: bool = False):
    """Lesion-like content (the reference's msseg2 data are FLAIR volumes with small bright MS lesions): a smooth
    low-amplitude background (low-pass filtered noise) + 0.05 * white noise, and ~60 compact
    ellipsoidal lesions (semi-axes 6..14 voxels) that carry a fixed two-channel signature.  ``with_mask`` also
    returns the lesion mask (uint8), which is the target the confident read-out head was fitted to and the target
    of the Dice / confusion step."""
    g = torch.Generator().manual_seed(1234 + seed)
    c, w, h, d = VOLUME
    coarse = torch.randn(1, c, w // 16, h // 16, d // 16, generator=g)
    vol = 0.15 * torch.nn.functional.interpolate(coarse, size=(w, h, d), mode="trilinear", align_corners=False)[0]
    vol += 0.05 * torch.randn(c, w, h, d, generator=g)
    mask = torch.zeros((w, h, d), dtype=torch.bool)
    n_lesions = 60
    centres = torch.rand(n_lesions, 3, generator=g) * torch.tensor([w, h, d], dtype=torch.float32)
    radii = 6 + 8 * torch.rand(n_lesions, 3, generator=g)
    for ctr, r in zip(centres.tolist(), radii.tolist()):
        lo = [max(int(ctr[a] - r[a]) - 1, 0) for a in range(3)]
        hi = [min(int(ctr[a] + r[a]) + 2, (w, h, d)[a]) for a in range(3)]
        zz, yy, xx = torch.meshgrid(*[torch.arange(lo[a], hi[a], dtype=torch.float32) for a in range(3)], indexing="ij")
        inside = ((zz - ctr[0]) / r[0]) ** 2 + ((yy - ctr[1]) / r[1]) ** 2 + ((xx - ctr[2]) / r[2]) ** 2 < 1
        mask[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] |= inside
    sig = torch.tensor([1.0, -0.6])[:c]
    vol += sig[:, None, None, None] * mask[None].float()
    vol = vol.contiguous()
    return (vol, mask.to(torch.uint8)) if with_mask else vol


def perturb_bn(model, seed: int = 1) -> None:
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm3d):
            m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
            m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
            m.weight.data.copy_(0.5 + torch.rand(m.weight.shape, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.bias.shape, generator=g))


def build_model():
    from segmentation_pipeline import models as M
    torch.manual_seed(0)
    model = M.ModularUNet(in_channels=2, out_channels=2, filters=list(FILTERS), depth=6,
                          block_params={'residual': True}, downsample_class=M.BlurConv3d,
                          downsample_params={'kernel_size': 3, 'stride': 2, 'padding': 1},
                          upsample_class=M.BlurConvTranspose3d,
                          upsample_params={'kernel_size': 3, 'stride': 2, 'padding': 1, 'output_padding': 0})
    perturb_bn(model)
    load_readout(model)
    return model.eval()


READOUT = os.path.join(ROOT, "tests", "golden", "readout_msseg2.npz")


def load_readout(model) -> bool:
    """Confident head: out_conv = the linear read-out fitted once on the CPU oracle's features
    of this very model / volume distribution by oracle/make_readout.py and committed as a small fixture -- so that
    the bench's label check counts every voxel instead of measuring the last bits of a p ~ 0.5 random head."""
    if not os.path.exists(READOUT):
        return False
    z = np.load(READOUT)
    with torch.no_grad():
        model.out_conv.weight.zero_()
        model.out_conv.weight[:, :, 1, 1, 1] = torch.from_numpy(z["weight"])
        model.out_conv.bias.copy_(torch.from_numpy(z["bias"]))
    return True


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- CPU reference arm
ORACLE_CFG = {"depth": 6, "filters": FILTERS, "block": {"residual": True}, "down": "blur", "up": "blur"}
SAMPLE_PATCHES = (5, 23, 41, 59, 77, 95, 113, 131)      # spread over the 144-patch grid (lesions in every one)


class CpuReference:
    """The reference's CPU path (oracle port: oracle/unet.py + oracle/grid.py -- the reference is pure Python over
    ATen + torchio and cannot be imported on the GPU box) for single patches of the config-2 volume:
    GridSampler crop + fp32 forward + GridAggregator '+='.  Used by the cpu_baseline leg and by --impl reference."""

    def __init__(self, threads: int):
        from oracle import grid as ogrid, unet
        torch.set_num_threads(threads)
        self.ogrid, self.unet = ogrid, unet
        model = build_model()
        self.sd = {k: v.clone() for k, v in model.state_dict().items()}
        vol = synthetic_volume(0).numpy()
        self.padded = ogrid.pad_volume(vol, OVERLAP, PADDING)
        self.loc = ogrid.grid_locations(self.padded.shape[1:], PATCH, OVERLAP)
        assert len(self.loc) == N_PATCHES
        self.out = np.zeros((2, *self.padded.shape[1:]), np.float32)
        self.cnt = np.zeros((2, *self.padded.shape[1:]), np.float32)
        self.outputs = {}

    def patch_input(self, i: int) -> np.ndarray:
        return self.ogrid.extract_patches(self.padded, self.loc[i % N_PATCHES][None])

    def run_patch(self, i: int, keep: bool = False) -> float:
        t0 = time.perf_counter()
        l = self.loc[i % N_PATCHES]
        patch = self.patch_input(i)
        with torch.no_grad():
            y = self.unet.modular_unet_forward(self.sd, torch.from_numpy(patch), ORACLE_CFG).numpy()
        i0, j0, k0, i1, j1, k1 = l
        self.out[:, i0:i1, j0:j1, k0:k1] += y[0]
        self.cnt[:, i0:i1, j0:j1, k0:k1] += 1
        dt = time.perf_counter() - t0
        if keep:
            self.outputs[i] = y[0]
        return dt

    def finalize_seconds(self) -> float:
        """divide + crop + argmax of the full padded volume (get_output_tensor + CustomArgMax), measured once."""
        out = torch.from_numpy(self.out)
        cnt = torch.from_numpy(np.maximum(self.cnt, 1))
        t0 = time.perf_counter()
        probs = torch.true_divide(out, cnt)                       # GridAggregator.get_output_tensor
        b = OVERLAP // 2
        probs = probs[:, b:-b, b:-b, b:-b]                        # torchio Crop of the padded border
        torch.argmax(probs, dim=0, keepdim=True)                  # CustomArgMax (custom_label_transforms.py:267)
        return time.perf_counter() - t0


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ref = CpuReference(threads)
    # bounded sample per step so that the whole --steps/--warmup run ends within a few minutes; at least 8 patches
    # are timed in total
    per_step = 8 if args.steps <= 2 else (3 if args.steps <= 8 else 1)
    per_step = max(per_step, -(-8 // max(args.steps, 1)))
    k = 0
    for _ in range(args.warmup):
        ref.run_patch(SAMPLE_PATCHES[k % len(SAMPLE_PATCHES)])
        k += 1
    times = []
    for _ in range(args.steps):
        for _ in range(per_step):
            times.append(ref.run_patch(SAMPLE_PATCHES[k % len(SAMPLE_PATCHES)] + k // len(SAMPLE_PATCHES)))
            k += 1
    t_fin = ref.finalize_seconds()
    t_patch = sum(times) / len(times)
    vox = VOLUME[1] * VOLUME[2] * VOLUME[3]
    t_volume = N_PATCHES * t_patch + t_fin
    value = vox / t_volume / 1e6
    sample = (f"{per_step} of {N_PATCHES} patches per step, {len(times)} timed in total (extract + fp32 forward + "
              f"overlap-add; {t_patch:.2f} s per patch, min {min(times):.2f} max {max(times):.2f}), extrapolated "
              f"x{N_PATCHES}; divide/crop/argmax of the padded volume measured once ({t_fin:.2f} s) and added")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_volume * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "arithmetic": "fp32 on the host CPU (the reference's own precision)",
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------- GPU arm
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p.get("bf16_tflops_sustained", 1351.6), p.get("hbm_gbs", 6554.9), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic():
    """DRAM bytes per launch of the dominant conv layer from the committed ncu --set full capture
    (profiles/roofline_traffic.json, written by tools/ncu_summary.py from the .ncu-rep); None when absent."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        t = json.load(f)
    return t.get("dram_bytes_per_launch"), t


class ConvTimer:
    """CUDA-event pair around every conv_tc launch INSIDE the timed steps (on the launching stream); an event record
    costs ~1 us of host time against ~700 us per launch."""

    def __init__(self, b200seg):
        self.lib = b200seg
        self.orig = b200seg.conv3d_tc
        self.pairs = []

    def __enter__(self):
        orig, pairs = self.orig, self.pairs

        def traced(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            orig(*a, **k)
            e.record()
            pairs.append((s, e))

        self.lib.conv3d_tc = traced
        return self

    def __exit__(self, *exc):
        self.lib.conv3d_tc = self.orig
        return False

    def total_ms(self) -> float:
        return sum(s.elapsed_time(e) for s, e in self.pairs)


def gpu_incumbent(ref: "CpuReference", device, n_batch: int = 8) -> dict:
    """The 'existing' GPU path (SURVEY.md section 2.1 'bar to beat'): the same network as plain ATen/cuDNN calls
    (oracle/unet.py's functional forward on CUDA tensors) under torch.autocast(bfloat16) with channels_last_3d input,
    on the same box, same weights, batches of 8 patches.  Part of the baseline leg (rank 0, N = 1 only)."""
    sd = {k: v.to(device) for k, v in ref.sd.items()}
    x = torch.from_numpy(np.concatenate([ref.patch_input(i) for i in SAMPLE_PATCHES[:n_batch]])).to(device)
    x = x.contiguous(memory_format=torch.channels_last_3d)
    try:
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            for _ in range(2):
                y = ref.unet.modular_unet_forward(sd, x, ORACLE_CFG)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record()
            for _ in range(reps):
                y = ref.unet.modular_unet_forward(sd, x, ORACLE_CFG)
            e1.record()
            torch.cuda.synchronize()
        ms_batch = e0.elapsed_time(e1) / reps
    except Exception as exc:  # noqa: BLE001  (cuDNN may not have a bf16 NDHWC kernel for every layer)
        return {"unavailable": f"{type(exc).__name__}: {str(exc)[:200]}"}
    vox = VOLUME[1] * VOLUME[2] * VOLUME[3]
    ms_volume = ms_batch * N_PATCHES / n_batch
    return {"value": vox / (ms_volume * 1e-3) / 1e6, "unit": UNIT, "ms_per_volume_forward_only": ms_volume,
            "tflops": FLOP_PER_PATCH * n_batch / (ms_batch * 1e-3) / 1e12,
            "what": f"ATen/cuDNN functional forward, autocast bf16 + channels_last_3d, {n_batch} patches per batch, "
                    f"{reps} timed batches, extrapolated x{N_PATCHES // n_batch}; forward only (no extraction / "
                    f"aggregation / argmax)",
            "out_dtype": str(y.dtype)}


def gpu_incumbent_train(ref: "CpuReference", device, steps: int = 3) -> dict:
    """The 'existing' GPU TRAINING step next to ``train_config5``: the same network, batch and loss as plain ATen / cuDNN
    autograd (oracle/unet.py's functional forward with batch-statistic BatchNorm on CUDA tensors, torch.autocast(bfloat16),
    channels_last_3d input), SGD on the same parameters.  Baseline leg (rank 0, N = 1 only)."""
    try:
        sd = {k: v.to(device).clone() for k, v in ref.sd.items()}
        params = [v.requires_grad_(True) for k, v in sd.items()
                  if v.is_floating_point() and "running" not in k and "kernel" not in k]
        opt = torch.optim.SGD(params, lr=1e-3, momentum=0.95)
        g = torch.Generator().manual_seed(100)
        x = torch.randn(TRAIN_BATCH, 2, PATCH, PATCH, PATCH, generator=g).to(device)
        x = x.contiguous(memory_format=torch.channels_last_3d)
        labels = (torch.rand(TRAIN_BATCH, PATCH, PATCH, PATCH, generator=g) < 0.05).long()
        y = torch.nn.functional.one_hot(labels, 2).movedim(-1, 1).float().to(device)
        cfg = dict(ORACLE_CFG, block=dict(ORACLE_CFG["block"], bn_training=True))
        class_weights = torch.tensor([1.0, 100.0], device=device)

        def torch_hybrid_loss(p, t, w, eps=1e-8):
            """criterions/hybrid_logistic_dice_loss.py:13-43 as tensor expressions on the device."""
            dims = (2, 3, 4)
            dice = 2 * (p * t).sum(dims) / ((t * t).sum(dims) + (p * p).sum(dims) + eps)
            logistic = (t * torch.log((p + eps) / (1 + eps))).mean(dims) * w[None]
            return 0.5 * (-logistic).mean() + 0.5 * (1 - dice).mean()

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                probs = ref.unet.modular_unet_forward(sd, x, cfg)
            loss = torch_hybrid_loss(probs.float(), y, class_weights)
            opt.zero_grad()
            loss.backward()
            opt.step()

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": f"{type(exc).__name__}: {str(exc)[:200]}"}
    return {"ms_per_step": ms, "patches_per_s": TRAIN_BATCH / (ms * 1e-3),
            "tflops": 3 * FLOP_PER_PATCH * TRAIN_BATCH / (ms * 1e-3) / 1e12,
            "what": f"ATen/cuDNN autograd of the functional network, autocast bf16 + channels_last_3d, batch {TRAIN_BATCH} x "
                    f"(2, {PATCH}^3), torch loss, SGD; {steps} timed steps after 2 warm-up steps"}


def label_check(ref: "CpuReference", model, device) -> dict:
    """Unfiltered argmax agreement of the CUDA bf16 path with the CPU fp32 oracle on the sample patches the baseline
    leg just ran (same weights incl. the fitted head), plus the oracle's top-2 margin histogram of the disagreeing
    voxels."""
    idx = sorted(ref.outputs)
    x = torch.from_numpy(np.concatenate([ref.patch_input(i) for i in idx])).to(device)
    with torch.no_grad():
        got = model(x).float().cpu()
    want = torch.from_numpy(np.stack([ref.outputs[i] for i in idx]))
    lw, lg = want.argmax(1), got.argmax(1)
    top2 = torch.topk(want, 2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    dis = margin[lw != lg]
    edges = [0.0, 1e-4, 1e-3, 1e-2, 5e-2, 1.0 + 1e-6]
    return {"patches": len(idx), "voxels": int(lw.numel()), "agreement": float((lw == lg).float().mean()),
            "class_fractions": [round(float((lw == c).float().mean()), 4) for c in range(want.shape[1])],
            "median_margin": float(margin.median()), "disagreeing": int(dis.numel()),
            "disagree_margin_hist": dict(zip(["<1e-4", "<1e-3", "<1e-2", "<5e-2", ">=5e-2"],
                                             [int(((dis >= lo) & (dis < hi)).sum()) for lo, hi in zip(edges[:-1], edges[1:])])),
            "max_abs_prob_err": float((want - got).abs().max()),
            "head": "fitted read-out (tests/golden/readout_msseg2.npz)" if os.path.exists(READOUT) else "random init"}


# ------------------------------------------------------------------------------------------------- z-slab (config 3)
SLAB_VOLUME = (2, 224, 224, 224)
SLAB_WORKLOAD = ("config3: qsm_deep_grey_matter-style NestedResUNet(2 -> 10, filters 40), synthetic 2x224^3 volume, patch 96^3 "
                 "overlap 48 padding edge (125 patches), bf16, ONE volume partitioned over the GPUs")
SLAB_FLOP_PER_PATCH = 1529078.0 * 96 ** 3


def build_slab_model():
    from segmentation_pipeline import models as M
    torch.manual_seed(3)
    model = M.NestedResUNet(2, 10, 40)
    perturb_bn(model, 4)
    return model.eval()


def slab_volume() -> torch.Tensor:
    g = torch.Generator().manual_seed(4321)
    c, w, h, d = SLAB_VOLUME
    coarse = torch.randn(1, c, w // 8, h // 8, d // 8, generator=g)
    vol = torch.nn.functional.interpolate(coarse, size=(w, h, d), mode="trilinear", align_corners=False)[0]
    return (vol + 0.1 * torch.randn(c, w, h, d, generator=g)).contiguous()


def run_slab_section(world: int, rank: int, device, steps: int = 3, warmup: int = 2, patch_batch: int = 16) -> dict:
    """Strong scaling of ONE config-3 volume over the ranks (segmentation_pipeline.distributed.slab_predict: balanced
    runs of the sorted patch list, one all-to-all of raw output blocks over NCCL, owner-side accumulation in sorted
    order, label all-gather), timed on the device as max over ranks -- and, on rank 0, the same volume on one GPU for
    the speed-up and for the bit-exactness check of labels and probabilities."""
    import torch.distributed as dist
    from segmentation_pipeline.distributed import CudaSlabOps, make_slab_plan, slab_predict
    from segmentation_pipeline.grid import PatchGrid
    from segmentation_pipeline.prediction import PatchPredict
    model = build_slab_model().to(device)
    vol = slab_volume().to(device)
    grid = PatchGrid(vol.shape[1:], PATCH, OVERLAP, PADDING)
    ops = CudaSlabOps(model, patch_batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(warmup):
            labels, probs = slab_predict(vol, grid, ops, gather_probs=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            labels, _ = slab_predict(vol, grid, ops, gather_probs=False)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_slab = float(ms.item())
        out = None
        if rank == 0:
            predictor = PatchPredict(patch_batch_size=patch_batch, patch_size=PATCH, patch_overlap=OVERLAP,
                                     padding_mode=PADDING)
            for _ in range(warmup):
                ref_probs, ref_labels = predictor.predict_volume(model, vol)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                predictor.predict_volume(model, vol, want_probs=False)
            e1.record()
            torch.cuda.synchronize()
            ms_one = e0.elapsed_time(e1) / steps
            plan = make_slab_plan(grid, world)
            send, _ = plan.split_sizes(0, 10)
            vox = SLAB_VOLUME[1] * SLAB_VOLUME[2] * SLAB_VOLUME[3]
            out = {"workload": SLAB_WORKLOAD, "n_gpus": world, "scaling": "strong", "patches": len(grid.locations),
                   "patches_per_rank": [b - a for a, b in plan.runs], "patch_batch": patch_batch,
                   "ms_per_volume": ms_slab, "ms_per_volume_1gpu": ms_one, "speedup": ms_one / ms_slab,
                   "efficiency": ms_one / ms_slab / world,
                   "value": vox / (ms_slab * 1e-3) / 1e6, "unit": UNIT,
                   "tflops": SLAB_FLOP_PER_PATCH * len(grid.locations) / (ms_slab * 1e-3) / 1e12,
                   "labels_bit_identical_to_1gpu": bool(torch.equal(labels, ref_labels)),
                   "probs_bit_identical_to_1gpu": bool(torch.equal(probs, ref_probs)),
                   "all_to_all_bytes_sent_by_rank0": 4 * (sum(send) - send[0]),
                   "collectives": "all_to_all_single (fp32 output blocks) + all_gather (uint8 label slabs), NCCL"}
        barrier()
    return out


# ------------------------------------------------------------------------------------------------- config 1
CONFIG1_WORKLOAD = ("config1: reference ModularUNet(1 -> 2, filters [40, 80, 120], depth 3, AvgPool / trilinear), synthetic "
                    "1x96^3 volume, patch 64^3 overlap 16 (8 patches)")


def run_config1_section(device, steps: int = 3, warmup: int = 2) -> dict:
    """BASELINE config 1 (the reference's CPU-runnable case) on the GPU in both precisions: the fp32 CUDA-core path
    (logits within 1e-5 of the reference) and the bf16 tensor-core path."""
    from segmentation_pipeline import models as M
    from segmentation_pipeline.models import set_precision
    from segmentation_pipeline.prediction import PatchPredict
    torch.manual_seed(0)
    model = M.ModularUNet(1, 2, [40, 80, 120], 3)
    perturb_bn(model, 1)
    model.eval().to(device)
    g = torch.Generator().manual_seed(3)
    vol = torch.randn(1, 96, 96, 96, generator=g).to(device)
    predictor = PatchPredict(patch_batch_size=8, patch_size=64, patch_overlap=16, padding_mode=None)
    out = {"workload": CONFIG1_WORKLOAD, "flop_per_volume": 0.953e12}
    labels = {}
    with torch.no_grad():
        for precision in ("fp32", "bf16"):
            set_precision(precision)
            for _ in range(warmup):
                predictor.predict_volume(model, vol, want_probs=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                _, labels[precision] = predictor.predict_volume(model, vol, want_probs=False)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[precision] = {"ms_per_volume": ms, "value": 96 ** 3 / (ms * 1e-3) / 1e6, "unit": UNIT,
                              "tflops": 0.953e12 / (ms * 1e-3) / 1e12}
    set_precision("bf16")
    out["label_agreement_bf16_vs_fp32_path"] = float((labels["fp32"] == labels["bf16"]).float().mean())
    return out


# ------------------------------------------------------------------------------------------------- cohort (config 4)
COHORT_WORKLOAD = ("config4: dmri_hippo-style NestedResUNet(3 -> 2, filters 40, dropout 0.2) eval, cohort of 64 synthetic "
                   "3x96x88x24 volumes, StandardPredict(sagittal_split=True), 64 / N volumes per GPU, Dice on device")
COHORT_SHAPE = (3, 96, 88, 24)
COHORT_SIZE = 64


def run_cohort_section(world: int, rank: int, device, steps: int = 3, warmup: int = 2, subjects_per_call: int = 8) -> dict:
    """BASELINE config 4: every rank runs its share of a 64-subject cohort through the reference-facing API --
    StandardPredict(sagittal_split=True).predict -> add_evaluation_labels -> the device confusion histogram -- and the
    int64 confusion matrices are all-reduced (the only collective of cohort mode; exact)."""
    import torch.distributed as dist
    import b200seg
    from segmentation_pipeline import _tio, models as M
    from segmentation_pipeline.distributed import all_reduce_confusion, shard_subjects
    from segmentation_pipeline.evaluators.segmentation_evaluator import counts_from_cm
    from segmentation_pipeline.prediction import StandardPredict, split_and_flip, reverse_split_and_flip
    torch.manual_seed(11)
    model = M.NestedResUNet(3, 2, 40, dropout_p=0.2)
    perturb_bn(model, 12)
    model.eval().to(device)
    mine = shard_subjects(COHORT_SIZE, rank, world)
    g = torch.Generator().manual_seed(500 + rank)
    vols = [torch.randn(COHORT_SHAPE, generator=g).pin_memory() for _ in mine]
    targets = [(torch.rand(COHORT_SHAPE[1:], generator=g) > 0.5).to(torch.uint8).to(device) for _ in mine]
    subjects = [_tio.Subject(X=_tio.ScalarImage(tensor=v), name=f"c{rank}_{i}") for i, v in zip(mine, vols)]
    predictor = StandardPredict(sagittal_split=True)
    cm = torch.zeros((2, 2), dtype=torch.int64, device=device)
    vox = COHORT_SHAPE[1] * COHORT_SHAPE[2] * COHORT_SHAPE[3]

    def step_api():
        """public API: pinned host volumes in, host probabilities (LabelMap) out, labels + counts on the device"""
        for a in range(0, len(subjects), subjects_per_call):
            chunk = subjects[a:a + subjects_per_call]
            _, batch = predictor.predict(model, device, chunk, {"label_values": {"left_whole": 1}})
            y = batch["y_pred"]
            labels = torch.empty((y.shape[0], *y.shape[2:]), dtype=torch.uint8, device=device)
            for i in range(y.shape[0]):
                b200seg.argmax(y[i].contiguous(), None, labels[i])
                b200seg.confusion(labels[i], targets[a + i], 2, cm)

    dev_batches = [torch.stack(vols[a:a + subjects_per_call]).to(device) for a in range(0, len(vols), subjects_per_call)]

    def step_resident():
        for bi, xb in enumerate(dev_batches):
            y = reverse_split_and_flip(model(split_and_flip(xb).contiguous()))
            for i in range(y.shape[0]):
                lab = torch.empty(y.shape[2:], dtype=torch.uint8, device=device)
                b200seg.argmax(y[i].contiguous(), None, lab)
                b200seg.confusion(lab, targets[bi * subjects_per_call + i], 2, cm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(warmup):
            fn()
        cm.zero_()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with torch.no_grad():
        ms_res = timed(step_resident)
        ms_api = timed(step_api)
        all_reduce_confusion(cm)
    tp, fp, tn, fn = counts_from_cm(cm.cpu(), 1)
    flop = 1516118.0 * vox * COHORT_SIZE
    return {"workload": COHORT_WORKLOAD, "n_gpus": world, "scaling": "strong (fixed cohort of 64)",
            "volumes_per_gpu": len(mine), "subjects_per_call": subjects_per_call,
            "ms_per_cohort": ms_res, "volumes_per_s": COHORT_SIZE / (ms_res * 1e-3),
            "value": COHORT_SIZE * vox / (ms_res * 1e-3) / 1e6, "unit": UNIT, "tflops": flop / (ms_res * 1e-3) / 1e12,
            "e2e": {"ms_per_cohort": ms_api, "value": COHORT_SIZE * vox / (ms_api * 1e-3) / 1e6, "unit": UNIT,
                    "api": "StandardPredict(sagittal_split=True).predict on pinned host subjects -> host LabelMaps; argmax "
                           "+ confusion on the device"},
            "confusion_total": int(cm.sum()), "dice_vs_random_target": 2 * tp / max(2 * tp + fp + fn, 1),
            "collective": "all_reduce(SUM) of the 2x2 int64 confusion matrix"}


# ------------------------------------------------------------------------------------------------- training (config 5)
TRAIN_BATCH = 4
TRAIN_WORKLOAD = ("config5: one trainer iteration (segmentation_trainer.py:162-180) of the msseg2 ModularUNet in training mode "
                  "(BatchNorm3d batch statistics), batch 4 x (2, 96^3) per GPU (research/msseg2/msseg2.py:94,153), fp32, "
                  "HybridLogisticDiceLoss([1, 100]), backward, SGD(lr 1e-3, momentum 0.95); N > 1: "
                  "DistributedDataParallel gradient all-reduce over NCCL")


class _StdoutToStderr:
    """fd-level redirect: NCCL writes its version banner to fd 1 when DistributedDataParallel creates its communicator;
    the bench contract is ONE JSON line on stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def run_train_section(world: int, rank: int, local: int, device, steps: int = 4, warmup: int = 3) -> dict:
    """BASELINE config 5 in both precisions of the training step: 'fp32' (the reference's arithmetic: CUDA-core
    convolutions) and 'bf16' (mixed precision: forward and dgrad on the tcgen05 engine, fp32 statistics / wgrad /
    parameters)."""
    out = {"workload": TRAIN_WORKLOAD}
    with _StdoutToStderr():
        for precision in ("fp32", "bf16"):
            out[precision] = _run_train(world, rank, local, device, precision, steps, warmup)
    return out


def _run_train(world: int, rank: int, local: int, device, precision: str, steps: int, warmup: int) -> dict:
    """Forward + loss + backward + optimizer step; every activation kernel from libb200seg."""
    import torch.distributed as dist
    from segmentation_pipeline.criterions.hybrid_logistic_dice_loss import HybridLogisticDiceLoss
    from segmentation_pipeline.models import set_precision
    set_precision(precision)
    torch.cuda.reset_peak_memory_stats(device)
    net = build_model().to(device).train()
    model = net
    if world > 1:
        # the blur convolutions' bias parameters are never used by the reference's forward (components.py:119)
        model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], find_unused_parameters=True)
    opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.95)
    criterion = HybridLogisticDiceLoss(logistic_class_weights=[1, 100])
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(TRAIN_BATCH, 2, PATCH, PATCH, PATCH, generator=g).to(device)
    labels = (torch.rand(TRAIN_BATCH, PATCH, PATCH, PATCH, generator=g) < 0.05).long()
    y = torch.nn.functional.one_hot(labels, 2).movedim(-1, 1).float().to(device)
    losses = []

    def step():
        out = criterion(model(x), y)
        opt.zero_grad()
        out["loss"].backward()
        opt.step()
        losses.append(out["loss"].detach())

    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    marks[0].record()
    for i in range(steps):
        step()
        marks[i + 1].record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    step_ms = [a.elapsed_time(b) for a, b in zip(marks[:-1], marks[1:])]
    ms = torch.tensor([marks[0].elapsed_time(marks[-1]) / steps], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    set_precision("bf16")
    patches = world * TRAIN_BATCH
    return {"n_gpus": world, "dtype": "f32" if precision == "fp32" else "bf16 activations / fp32 accumulation, statistics and weights",
            "scaling": "weak", "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "step_ms_rank0": step_ms, "patches_per_s": patches / (ms * 1e-3),
            "value": patches * PATCH ** 3 / (ms * 1e-3) / 1e6, "unit": "Mvoxel/s (patch voxels trained)",
            "tflops": 3 * FLOP_PER_PATCH * patches / (ms * 1e-3) / 1e12,
            "tflops_basis": "3 x the forward FLOPs (forward + dgrad + wgrad)",
            "loss": [float(v) for v in losses],
            "peak_mem_gib": torch.cuda.max_memory_allocated(device) / 2 ** 30,
            "collective": "DistributedDataParallel bucketed all-reduce of the fp32 gradients (NCCL)" if world > 1 else None}


def run_gpu_arm(args) -> None:
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)     # NCCL_DEBUG & co. are left exactly as the launcher set them

    import b200seg
    from segmentation_pipeline import _tio
    from segmentation_pipeline.evaluators.segmentation_evaluator import counts_from_cm
    from segmentation_pipeline.models import set_precision
    from segmentation_pipeline.prediction import PatchPredict
    b200seg.load_library()   # raises if the CUDA extension is missing: no fallback
    set_precision("bf16")
    model = build_model().to(device)
    predictor = PatchPredict(patch_batch_size=PATCH_BATCH, patch_size=PATCH, patch_overlap=OVERLAP,
                             padding_mode=PADDING, overlap_mode="average")
    vol, mask = synthetic_volume(rank, with_mask=True)
    vol_host = vol.pin_memory()
    vol_dev = vol_host.to(device)
    target = mask.to(device)
    cm = torch.zeros((2, 2), dtype=torch.int64, device=device)
    vox = VOLUME[1] * VOLUME[2] * VOLUME[3]

    def step_resident():
        _, labels = predictor.predict_volume(model, vol_dev, want_probs=False, want_labels=True)
        b200seg.confusion(labels, target, 2, cm)
        return labels

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    warmup = max(args.warmup, 3)
    with torch.no_grad():
        for _ in range(warmup):
            step_resident()
        cm.zero_()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        launches0 = b200seg.launches()
        with ConvTimer(b200seg) as conv_timer:             # events INSIDE the timed steps
            ms = timed(step_resident, args.steps)
        gpu_launches = b200seg.launches() - launches0
        clocks = sampler.stop() if rank == 0 else None
        conv_ms = conv_timer.total_ms() / args.steps
        n_conv = len(conv_timer.pairs) // args.steps
        # Dice of the lesion class from the accumulated device confusion matrix (cohort: summed over ranks, exact)
        if world > 1:
            dist.all_reduce(cm, op=dist.ReduceOp.SUM)
        tp, fp, tn, fn = counts_from_cm(cm.cpu(), 1)
        dice = 2 * tp / max(2 * tp + fp + fn, 1)

        # ---- end to end through the reference-facing API: pinned host volume in, host probabilities out
        subject = _tio.Subject(X=_tio.ScalarImage(tensor=vol_host), name=f"bench{rank}")

        def step_e2e():
            predictor.predict(model, device, [subject], {"label_values": {"lesion": 1}})

        for _ in range(warmup):     # also warms the pinned-host allocator (two 100 MB blocks alternate)
            step_e2e()
        e2e_steps = max(1, min(args.steps, 5))
        ms_e2e = timed(step_e2e, e2e_steps)

    # ---- patch_batch_size = 1, the reference inference script's setting (research/msseg2/competition/ms-inference.py:32):
    #      144 single-patch forwards per volume; the plan of one patch is replayed as a CUDA graph
    batch1 = None
    if rank == 0 and not args.main_only:
        from segmentation_pipeline import prediction as _pred
        p1 = PatchPredict(patch_batch_size=1, patch_size=PATCH, patch_overlap=OVERLAP, padding_mode=PADDING)
        batch1 = {"patch_batch_size": 1}
        with torch.no_grad():
            _, lab48 = predictor.predict_volume(model, vol_dev, want_probs=False)
            for name, device_batch in (("regrouped_on_device", 48), ("literal", 0)):
                _pred.set_device_batch(device_batch)
                for _ in range(2):
                    p1.predict_volume(model, vol_dev, want_probs=False)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    _, lab1 = p1.predict_volume(model, vol_dev, want_probs=False)
                e1.record()
                torch.cuda.synchronize()
                ms1 = e0.elapsed_time(e1) / 3
                batch1[name] = {"ms_per_volume": ms1, "value": vox / (ms1 * 1e-3) / 1e6, "unit": UNIT,
                                "slowdown_vs_patch_batch_%d" % PATCH_BATCH: ms1 / (ms / args.steps),
                                "labels_identical_to_batched": bool(torch.equal(lab1, lab48))}
            _pred.set_device_batch(48)
        batch1["how"] = ("regrouped_on_device (default): patch_batch_size is a memory knob, results do not depend on it, so "
                         "up to 48 patches share a launch sequence when the workspace fits; literal: one 96^3 patch per "
                         "forward, its plan replayed 144 times as a CUDA graph")
    # ---- z-slab mode: one config-3 volume over all ranks (strong scaling; the collective path)
    set_precision("bf16")
    slab = run_slab_section(world, rank, device) if ((world > 1 or args.slab) and not args.main_only) else None
    cohort4 = None if args.main_only else run_cohort_section(world, rank, device)
    train5 = None if args.main_only else run_train_section(world, rank, local, device)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_tf, peak_gbs, peak_src = measured_peaks()
    ms_per_step = ms / args.steps
    value = world * vox / (ms_per_step * 1e-3) / 1e6
    flop_step = FLOP_PER_PATCH * N_PATCHES
    achieved_tf = flop_step / (conv_ms * 1e-3) / 1e12
    e2e_ms = ms_e2e / e2e_steps
    vol_bytes = vol_host.numel() * 4
    traffic, traffic_meta = committed_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "patch_batch": PATCH_BATCH, "volumes_per_step": world,
                   "partition": "one volume per GPU (cohort), no data-path collective",
                   "l2": "working set per step (>= 9 GB of activations, 100 MB volume) is far larger than the 126 MB L2"},
        "e2e": {"value": world * vox / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": vol_bytes,
                "d2h_bytes_per_step": vol_bytes, "ms_per_step": e2e_ms,
                "api": "PatchPredict.predict(model, device, [subject]) with a pinned host volume; returns host probabilities"},
        "gpu_launches": gpu_launches,
        # SURVEY.md section 8(d) also asks for patch-voxels per second (network throughput incl. the 8x overlap)
        "patch_mvoxel_per_s": world * N_PATCHES * PATCH ** 3 / (ms_per_step * 1e-3) / 1e6,
        "dice_lesion": dice,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf,
                     "frac_of_spec_2250": achieved_tf / 2250.0,
                     "traffic": traffic, "traffic_source": traffic_meta,
                     "kernel": "conv_tc_kernel",
                     "launches_per_step": n_conv, "kernel_ms_per_step": conv_ms,
                     "timing": "sum of CUDA-event pairs around every conv_tc launch inside the timed steps, averaged",
                     "share_of_step": conv_ms / ms_per_step,
                     "algorithmic_flop_per_step": flop_step, "peak_source": peak_src},
    }
    if slab is not None:
        line["slab"] = slab
    line["cohort_config4"] = cohort4
    line["train_config5"] = train5
    if world == 1 and not args.main_only:
        line["config1"] = run_config1_section(device)
        set_precision("bf16")
    line["patch_batch_1"] = batch1
    if world == 1 and not args.main_only:
        threads = os.cpu_count() or 1
        ref = CpuReference(threads)
        ref.run_patch(SAMPLE_PATCHES[0])                                  # warm-up
        times = [ref.run_patch(i, keep=True) for i in SAMPLE_PATCHES]
        t_fin = ref.finalize_seconds()
        t_patch = sum(times) / len(times)
        t_volume = N_PATCHES * t_patch + t_fin
        line["cpu_baseline"] = {"value": vox / t_volume / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{len(times)} of {N_PATCHES} patches (extract + fp32 forward + overlap-add, "
                                          f"{t_patch:.2f} s each) after 1 warm-up, extrapolated x{N_PATCHES}, plus "
                                          f"divide/crop/argmax of the padded volume ({t_fin:.2f} s)"}
        line["label_check"] = label_check(ref, model, device)
        line["gpu_incumbent"] = gpu_incumbent(ref, device)
        line["gpu_incumbent_train"] = gpu_incumbent_train(ref, device)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--slab", action="store_true", help="also run the config-3 z-slab section at N = 1")
    ap.add_argument("--main-only", action="store_true",
                    help="profiling aid: only the headline workload (config 2), none of the extra sections")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if not torch.cuda.is_available():
        sys.exit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    run_gpu_arm(args)


if __name__ == "__main__":
    main()
