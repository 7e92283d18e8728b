#!/usr/bin/env python
"""Device timing of the training-step kernels at the msseg2 level-0 size (batch 4 x 96^3, 40 channels)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segmentation-pipeline_b200")]

import torch  # noqa: E402

import b200seg as lib  # noqa: E402
from segmentation_pipeline.models import _train  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    e = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    dtype = torch.bfloat16 if (len(sys.argv) > 3 and sys.argv[3] == "bf16") else torch.float32
    dev = torch.device("cuda")
    lib.load_library()
    runner = _train._Runner({}, dev, "bf16" if dtype == torch.bfloat16 else "fp32")

    def buf(c, ext):
        b = lib.Blocked(n, (c + 7) // 8, ext, ext, ext, dtype, dev)
        b.tensor.normal_()
        return b

    vox = n * e ** 3
    for cin, cout in ((40, 40), (80, 40), (2, 40)):
        x, dz = buf(cin, e), buf(cout, e)
        ms = timed(lambda: lib.wgrad(dz.view(cout), x.view(cin), 3, 1, 1, dev))
        print(f"wgrad k3 {cin:3d}->{cout:3d} @{e}^3 x{n}: {ms:8.2f} ms  {2 * vox * 27 * cin * cout / ms / 1e9:7.1f} TFLOP/s")
        w = torch.randn(cout, cin, 3, 3, 3, device=dev)
        dx = buf(cin, e)
        wp = _train._pack(w.flip(2, 3, 4).permute(2, 3, 4, 0, 1))
        ms = timed(lambda: runner.conv(dz.view(cout), wp, cin, dx.view(cin)))
        print(f"dgrad k3 {cout:3d}->{cin:3d} @{e}^3 x{n}: {ms:8.2f} ms  {2 * vox * 27 * cin * cout / ms / 1e9:7.1f} TFLOP/s")
    x, dzc = buf(40, e), buf(40, e // 2)
    ms = timed(lambda: lib.wgrad(dzc.view(40), x.view(40), 4, 2, 1, dev))
    print(f"wgrad down 40->40 @{e // 2}^3 x{n}: {ms:8.2f} ms  {2 * vox / 8 * 64 * 1600 / ms / 1e9:7.1f} TFLOP/s")
    ms = timed(lambda: lib.wgrad(dzc.view(40), x.view(40), 4, 2, 1, dev))
    xc, dy = buf(40, e // 2), buf(40, e)
    ms = timed(lambda: lib.wgrad(xc.view(40), dy.view(40), 4, 2, 1, dev))
    print(f"wgrad up   40->40 @{e}^3 x{n}: {ms:8.2f} ms  {2 * vox / 8 * 64 * 1600 / ms / 1e9:7.1f} TFLOP/s")
    z, dy, dzb = buf(40, e), buf(40, e), buf(40, e)
    vec = lambda v: torch.full((40,), v, device=dev)
    bytes_t = vox * 40 * (2 if dtype == torch.bfloat16 else 4)
    ms = timed(lambda: lib.bn_backward(dy.view(40), z.view(40), vec(1.0), vec(0.0), vec(0.0), vec(0.0), vec(1.0), True,
                                       dzb.view(40), 40, dev))
    print(f"bn_backward 40ch: {ms:8.2f} ms  {5 * bytes_t / ms / 1e6:7.1f} GB/s (2 reads + 2 reads + 1 write)")
    ms = timed(lambda: lib.channel_moments(z.view(40), 40, dev))
    print(f"channel_moments 40ch: {ms:8.2f} ms  {bytes_t / ms / 1e6:7.1f} GB/s")
    ms = timed(lambda: lib.affine_act(z.view(40), vec(1.0), vec(0.0), vec(0.0), dzb.view(40), residual=dy.view(40)))
    print(f"affine_act+res 40ch: {ms:8.2f} ms  {3 * bytes_t / ms / 1e6:7.1f} GB/s")


if __name__ == "__main__":
    main()
