"""CPU check of the tensor-core engine's algorithm: host weight packer (models/_plan.py) x kernel data flow
(emulated byte-for-byte by tests/tc_emulator.py) == torch convolution.  Catches any disagreement between the
Python packer and the descriptor arithmetic of csrc/conv_tc.cu before a GPU is involved."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from segmentation_pipeline.models import _plan
from tc_emulator import DOWN, K3, K3T, UP, emulate_conv_tc


def _bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _blocked(x):
    n, c, z, y, xx = x.shape
    c8 = (c + 7) // 8
    out = torch.zeros(n, c8 * 8, z, y, xx)
    out[:, :c] = x
    return out.reshape(n, c8, 8, z, y, xx).permute(0, 1, 3, 4, 5, 2).contiguous().numpy().astype(np.float64)


CASES = [
    # mode, cin, cout, (z, y, x), TZ
    (K3, 24, 40, (5, 18, 10), None),     # 3 chunks -> pair + lone group, two tiles per axis, N spill (40 -> 48)
    (K3, 16, 16, (3, 7, 5), None),       # single pair, tile larger than the tensor
    (K3, 2, 8, (4, 17, 9), 3),           # lone chunk only (first layer), z split in two tiles
    (K3, 16, 80, (7, 16, 8), None),      # widest N (240), TZ limited to 6 -> two z tiles
    (DOWN, 16, 16, (8, 20, 12), None),
    (DOWN, 8, 24, (4, 34, 18), 1),       # lone chunk, several tiles, TZ = 1
    (UP, 16, 16, (3, 5, 9), None),
    (UP, 8, 40, (2, 17, 3), None),       # lone chunk, N spill
    (UP, 16, 80, (4, 4, 4), None),       # 4 planes x 80 > 256 -> column blocks of 2 planes
    (K3T, 40, 2, (5, 30, 14), None),     # out_conv shape: 2 pairs + lone chunk, 3 x 3 tiles of 14 x 6, 18 tap columns
    (K3T, 8, 1, (3, 15, 7), 2),          # lone chunk only, one column per tap, z split
    (K3T, 16, 4, (4, 14, 6), None),      # widest tap block (36 -> 40 columns), tile exactly full
]


@pytest.mark.parametrize("mode,cin,cout,ext,tz", CASES)
def test_engine_dataflow_matches_torch(mode, cin, cout, ext, tz):
    g = torch.Generator().manual_seed(1000 + cin * 7 + cout + mode)
    x = _bf16(torch.randn(1, cin, *ext, generator=g))
    if mode == UP:
        w = _bf16(torch.randn(cin, cout, 4, 4, 4, generator=g) * 0.1)
        ref = F.conv_transpose3d(x.double(), w.double(), stride=2, padding=1)
    elif mode == DOWN:
        w = _bf16(torch.randn(cout, cin, 4, 4, 4, generator=g) * 0.1)
        ref = F.conv3d(x.double(), w.double(), stride=2, padding=1)
    else:
        w = _bf16(torch.randn(cout, cin, 3, 3, 3, generator=g) * 0.1)
        ref = F.conv3d(x.double(), w.double(), padding=1)
    chunks = (cin + 7) // 8
    phys = _plan.physical_weight(w, mode == UP, [(0, cin)], chunks, 0, cout)
    packed = _plan.pack_tc_weight(mode, phys, chunks, cout)
    geo = _plan.tc_geometry(mode, chunks, cout)
    assert packed.numel() == geo["n_pass"] * geo["n_bimg"] * geo["bimg_elems"]
    out = emulate_conv_tc(mode, _blocked(x), packed.float().numpy().astype(np.float64), cout, tz)
    got = torch.from_numpy(out[:, :cout])
    assert not torch.isnan(got).any(), "engine read bytes it never wrote / accumulated before first touch"
    assert torch.allclose(got, ref, atol=1e-9, rtol=1e-9)
    # padded output channels must come out exactly zero (K3T returns the tap-summed channels only)
    assert np.all(out[:, cout:] == 0)


def test_concat_segments_map_to_physical_chunks():
    # cat([a (12 ch), b (8 ch)]): a occupies chunks 0-1 (4 pad channels), b chunk 2
    g = torch.Generator().manual_seed(5)
    a = _bf16(torch.randn(1, 12, 3, 6, 5, generator=g))
    b = _bf16(torch.randn(1, 8, 3, 6, 5, generator=g))
    w = _bf16(torch.randn(8, 20, 3, 3, 3, generator=g) * 0.1)
    ref = F.conv3d(torch.cat([a, b], 1).double(), w.double(), padding=1)
    xb = np.concatenate([_blocked(a), _blocked(b)], axis=1)
    phys = _plan.physical_weight(w, False, [(0, 12), (2, 8)], 3, 0, 8)
    packed = _plan.pack_tc_weight(K3, phys, 3, 8)
    out = emulate_conv_tc(K3, xb, packed.float().numpy().astype(np.float64), 8)
    assert torch.allclose(torch.from_numpy(out), ref, atol=1e-9, rtol=1e-9)


def test_direct_weight_layout():
    w = torch.arange(2 * 3 * 27, dtype=torch.float32).reshape(2, 3, 3, 3, 3)
    op = _plan.ConvOp(K3, _plan.Ref("in", 0, 3), [(0, 3)], w, np.ones(2, np.float32), np.zeros(2, np.float32),
                      np.ones(2, np.float32))
    packed = _plan.pack_direct_weight(op, 1)
    assert packed.shape == (27, 8, 8)
    assert packed[5, 2, 1] == w[1, 2].reshape(-1)[5]
    assert packed[:, 3:, :].abs().sum() == 0 and packed[:, :, 2:].abs().sum() == 0
