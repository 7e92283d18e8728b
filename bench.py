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
def synthetic_volume(seed: int) -> torch.Tensor:
    """Smooth-ish content in [-1, 1]: low-pass filtered noise + 0.1 * white noise."""
    g = torch.Generator().manual_seed(1234 + seed)
    c, w, h, d = VOLUME
    coarse = torch.randn(1, c, w // 4, h // 4, d // 4, generator=g)
    smooth = torch.nn.functional.interpolate(coarse, size=(w, h, d), mode="trilinear", align_corners=False)[0]
    vol = smooth + 0.1 * torch.randn(c, w, h, d, generator=g)
    vol = vol - vol.amin()
    return (vol / vol.amax() * 2 - 1).contiguous()


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
    return model.eval()


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
def cpu_reference_sample(threads: int, n_patches: int = 1):
    """The reference's CPU path (oracle port: oracle/unet.py + oracle/grid.py) on `n_patches` of the 144 patches:
    extraction + fp32 forward + overlap-add per patch.  Returns seconds per patch."""
    from oracle import grid as ogrid, unet
    torch.set_num_threads(threads)
    model = build_model()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    cfg = {"depth": 6, "filters": FILTERS, "block": {"residual": True}, "down": "blur", "up": "blur"}
    vol = synthetic_volume(0).numpy()
    padded = ogrid.pad_volume(vol, OVERLAP, PADDING)
    loc = ogrid.grid_locations(padded.shape[1:], PATCH, OVERLAP)
    assert len(loc) == N_PATCHES
    state = {"padded": padded, "loc": loc, "sd": sd, "cfg": cfg,
             "out": np.zeros((2, *padded.shape[1:]), np.float32), "cnt": np.zeros((2, *padded.shape[1:]), np.float32)}

    def run(first: int) -> float:
        t0 = time.perf_counter()
        for i in range(first, first + n_patches):
            l = loc[i % N_PATCHES]
            patch = ogrid.extract_patches(padded, l[None])
            with torch.no_grad():
                y = unet.modular_unet_forward(sd, torch.from_numpy(patch), cfg).numpy()
            i0, j0, k0, i1, j1, k1 = l
            state["out"][:, i0:i1, j0:j1, k0:k1] += y[0]
            state["cnt"][:, i0:i1, j0:j1, k0:k1] += 1
        return (time.perf_counter() - t0) / n_patches

    return run, state


def cpu_finalize_seconds(state) -> float:
    """divide + crop + argmax of the full padded volume on the CPU (the reference's get_output_tensor +
    CustomArgMax), measured once."""
    out = torch.from_numpy(state["out"])
    cnt = torch.from_numpy(np.maximum(state["cnt"], 1))
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
    run, state = cpu_reference_sample(threads)
    for w in range(args.warmup):
        run(w)
    times = [run(args.warmup + k) for k in range(args.steps)]
    t_fin = cpu_finalize_seconds(state)
    t_patch = sum(times) / len(times)
    vox = VOLUME[1] * VOLUME[2] * VOLUME[3]
    t_volume = N_PATCHES * t_patch + t_fin
    value = vox / t_volume / 1e6
    sample = (f"1 of {N_PATCHES} patches per step (extract + fp32 forward + overlap-add), extrapolated x{N_PATCHES}; "
              f"divide/crop/argmax of the padded volume measured once ({t_fin:.2f} s) and added")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_volume * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD.replace("bf16", "fp32 on host CPU"), "sample": sample},
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


def run_gpu_arm(args) -> None:
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner (and any NCCL_DEBUG output) to stdout
        # unless it is given a file
        os.environ["NCCL_DEBUG"] = os.environ.get("B200SEG_NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", os.path.join(ROOT, "gpurun_out", "nccl.%h.%p.log")
                              if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else os.devnull)
        dist.init_process_group("nccl", device_id=device)

    import b200seg
    from segmentation_pipeline import _tio
    from segmentation_pipeline.models import set_precision
    from segmentation_pipeline.prediction import PatchPredict
    b200seg.load_library()   # raises if the CUDA extension is missing: no fallback
    set_precision("bf16")
    model = build_model().to(device)
    predictor = PatchPredict(patch_batch_size=PATCH_BATCH, patch_size=PATCH, patch_overlap=OVERLAP,
                             padding_mode=PADDING, overlap_mode="average")
    vol_host = synthetic_volume(rank).pin_memory()
    vol_dev = vol_host.to(device)
    target = (torch.rand(VOLUME[1:], generator=torch.Generator().manual_seed(99 + rank)) > 0.5).to(torch.uint8).to(device)
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

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step_resident()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        launches0 = b200seg.launches()
        ms = timed(step_resident, args.steps)
        gpu_launches = b200seg.launches() - launches0
        clocks = sampler.stop() if rank == 0 else None

        # ---- end to end through the reference-facing API: pinned host volume in, host probabilities out
        subject = _tio.Subject(X=_tio.ScalarImage(tensor=vol_host), name=f"bench{rank}")

        def step_e2e():
            predictor.predict(model, device, [subject], {"label_values": {"lesion": 1}})

        for _ in range(max(args.warmup, 3)):     # also warms the pinned-host allocator (two 100 MB blocks alternate)
            step_e2e()
        ms_e2e = timed(step_e2e, max(1, min(args.steps, 3)))
        e2e_steps = max(1, min(args.steps, 3))

        # ---- roofline of the dominant kernel (conv_tc): per-launch CUDA events in a separate instrumented pass
        events = []
        orig = b200seg.conv3d_tc

        def traced(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            orig(*a, **k)
            e.record()
            events.append((s, e))

        b200seg.conv3d_tc = traced
        try:
            step_resident()
            torch.cuda.synchronize()
        finally:
            b200seg.conv3d_tc = orig
        conv_ms = sum(s.elapsed_time(e) for s, e in events)
        n_conv = len(events)

    if rank != 0:
        return
    peak_tf, peak_gbs, peak_src = measured_peaks()
    ms_per_step = ms / args.steps
    value = world * vox / (ms_per_step * 1e-3) / 1e6
    flop_step = FLOP_PER_PATCH * N_PATCHES
    achieved_tf = flop_step / (conv_ms * 1e-3) / 1e12
    e2e_ms = ms_e2e / e2e_steps
    vol_bytes = vol_host.numel() * 4
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
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
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf,
                     # dram__bytes_read.sum + dram__bytes_write.sum of the dominant launch (up_blocks.0 conv0||res_conv,
                     # 48 patches) from the committed ncu --set full capture profiles/r01_ncu_full_up0_conv0res_b48_raw.csv
                     "traffic": 14.721e9,
                     "traffic_note": "bytes per launch of the dominant conv layer (80->80 ch, 48 x 96^3); algorithmic "
                                     "bytes of that launch are 13.59e9 (activations in + out once)",
                     "kernel": "conv_tc_kernel",
                     "launches_per_step": n_conv, "kernel_ms_per_step": conv_ms,
                     "share_of_step": conv_ms / ms_per_step,
                     "algorithmic_flop_per_step": flop_step, "peak_source": peak_src},
    }
    if world == 1:
        threads = os.cpu_count() or 1
        run, state = cpu_reference_sample(threads)
        run(0)
        t_patch = (run(1) + run(2)) / 2
        t_fin = cpu_finalize_seconds(state)
        t_volume = N_PATCHES * t_patch + t_fin
        line["cpu_baseline"] = {"value": vox / t_volume / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"2 of {N_PATCHES} patches (extract + fp32 forward + overlap-add) after 1 warm-up, "
                                          f"extrapolated x{N_PATCHES}, plus divide/crop/argmax of the padded volume "
                                          f"({t_fin:.2f} s)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if not torch.cuda.is_available():
        sys.exit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    run_gpu_arm(args)


if __name__ == "__main__":
    main()
