"""ctypes binding of libb200seg.so (include/b200seg.h) for PyTorch tensors.

PyTorch is plumbing here: it owns device memory and streams; every function below passes raw
``data_ptr()`` values and the current CUDA stream handle through the C ABI.  There is no fallback: if the
shared library is missing or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_void_p
from typing import Optional, Sequence

import torch

F32, BF16 = 0, 1
TC_K3, TC_DOWN, TC_UP = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200SEG_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libb200seg.so"))


class View(Structure):
    """Mirror of ``b200seg_view``: a chunk range of a blocked activation buffer [N][C8][Z][Y][X][8]."""
    _fields_ = [("data", c_void_p), ("dtype", c_int32), ("n", c_int32), ("c", c_int32), ("c8_total", c_int32),
                ("c8_off", c_int32), ("z", c_int32), ("y", c_int32), ("x", c_int32)]


class Epilogue(Structure):
    """Mirror of ``b200seg_epilogue``."""
    _fields_ = [("scale", c_void_p), ("shift", c_void_p), ("slope", c_void_p), ("dst0", View), ("dst1", View),
                ("split", c_int32), ("residual", View), ("out_ncdhw", c_void_p), ("softmax", c_int32),
                ("slope01", c_int32)]


_lib = None

_SIGNATURES = {
    "b200seg_last_error": (c_char_p, []),
    "b200seg_version": (c_int32, []),
    "b200seg_device_info": (c_int32, [POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    "b200seg_pack_ncdhw": (c_int32, [c_void_p, View, c_void_p]),
    "b200seg_unpack_ncdhw": (c_int32, [View, c_void_p, c_void_p]),
    "b200seg_conv3d_direct": (c_int32, [View, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                        POINTER(Epilogue), c_void_p]),
    "b200seg_conv3d_tc": (c_int32, [c_int32, View, c_void_p, c_int64, c_int32, POINTER(Epilogue), c_void_p]),
    "b200seg_conv3d_tc_wbytes": (c_int64, [c_int32, c_int32, c_int32]),
    "b200seg_instnorm_scratch_bytes": (c_int64, [View]),
    "b200seg_instnorm": (c_int32, [View, c_void_p, c_void_p, c_float, c_float, View, c_void_p, c_int64, c_void_p]),
    "b200seg_avgpool2": (c_int32, [View, View, c_void_p]),
    "b200seg_upsample_trilinear2": (c_int32, [View, View, c_void_p]),
    "b200seg_copy_view": (c_int32, [View, View, c_void_p]),
    "b200seg_pack_ncdhw_tta": (c_int32, [c_void_p, c_int32, c_int32, c_int32, POINTER(c_int32), POINTER(c_int32), View,
                                         c_void_p]),
    "b200seg_tta_accumulate": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, POINTER(c_int32),
                                         POINTER(c_int32), c_void_p, c_void_p, c_void_p]),
    "b200seg_tta_finalize": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_int32, c_void_p]),
    "b200seg_copy_planes": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                      c_void_p]),
    "b200seg_softmax_ncdhw": (c_int32, [c_void_p, c_int64, c_int32, c_int64, c_int32, c_float, c_void_p]),
    "b200seg_grid_extract": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, POINTER(c_int32), c_int32,
                                       POINTER(c_int32), c_int32, c_float, View, c_void_p]),
    "b200seg_overlap_add": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, POINTER(c_int32),
                                      c_int32, c_void_p]),
    "b200seg_overlap_crop": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, POINTER(c_int32),
                                       c_int32, POINTER(c_int32), c_int32, c_void_p]),
    "b200seg_finalize": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                   POINTER(c_int32), c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200seg_finalize_region": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                          POINTER(c_int32), POINTER(c_int32), c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200seg_argmax": (c_int32, [c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_void_p]),
    "b200seg_hybrid_loss_scratch_bytes": (c_int64, [c_int64, c_int32]),
    "b200seg_hybrid_loss_forward": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int64, c_float, c_void_p, c_int32,
                                              c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "b200seg_hybrid_loss_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_float, c_void_p,
                                               c_int32, c_void_p, c_void_p, c_void_p]),
    "b200seg_ccl3d_roots": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                      c_int32, c_void_p, c_void_p]),
    "b200seg_train_scratch_bytes": (c_int64, [c_int32]),
    "b200seg_channel_moments": (c_int32, [View, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200seg_affine_act": (c_int32, [View, c_void_p, c_void_p, c_void_p, View, View, c_void_p]),
    "b200seg_bn_backward": (c_int32, [View, View, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p,
                                      c_void_p, c_void_p, View, c_void_p]),
    "b200seg_softmax_backward": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, View, c_void_p]),
    "b200seg_wgrad_scratch_floats": (c_int64, [c_int32, c_int32, c_int32]),
    "b200seg_wgrad": (c_int32, [View, View, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "b200seg_channel_scale": (c_int32, [View, c_void_p, View, c_void_p]),
    "b200seg_avgpool2_backward": (c_int32, [View, View, View, c_void_p]),
    "b200seg_upsample_trilinear2_backward": (c_int32, [View, View, c_void_p]),
    "b200seg_window_patches": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                         c_void_p, c_void_p]),
    "b200seg_divide_separable": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                           c_void_p]),
    "b200seg_relabel_lut": (c_int32, [c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_void_p]),
    "b200seg_dilate_cross": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "b200seg_dilate_where": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "b200seg_relabel_masked": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_int64, c_void_p,
                                         c_void_p]),
    "b200seg_label_equals": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "b200seg_mask_assign": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    "b200seg_ccl3d_relabel": (c_int32, [c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_void_p]),
    "b200seg_overlap_histogram": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "b200seg_confusion": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_int32, c_void_p, c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """Loads the C-ABI library and declares every prototype.  Raises if the library is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"libb200seg.so not found at {p}: build it with "
                           f"`python segmentation-pipeline_b200/build.py` (there is no CPU / PyTorch fallback)")
    lib = ctypes.CDLL(p)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().b200seg_last_error().decode()
        raise RuntimeError(f"b200seg {what} failed ({rc}): {msg}")


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


class on_device:
    """``with on_device(t_or_device):`` makes that CUDA device current for the block, so the stream handed to the
    library, the scratch allocations and cudaGetDevice() inside it all refer to the device the data lives on.  The
    reference API is ``predict(model, device, subjects)``: the caller's *current* device may be another GPU."""

    def __init__(self, where):
        dev = where.device if isinstance(where, torch.Tensor) else torch.device(where)
        if dev.type != "cuda":
            raise RuntimeError("b200seg operates on CUDA devices only (no CPU fallback)")
        self._ctx = torch.cuda.device(dev)

    def __enter__(self):
        return self._ctx.__enter__()

    def __exit__(self, *exc):
        return self._ctx.__exit__(*exc)


def _ptr(t: Optional[torch.Tensor]) -> c_void_p:
    return c_void_p(0 if t is None else t.data_ptr())


def _require_cuda(*tensors: torch.Tensor) -> None:
    """Every tensor of a call must live on the CURRENT CUDA device: kernels are launched on its stream, so a pointer
    from another GPU would fault or silently run through peer access (use ``on_device``)."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("b200seg operates on CUDA tensors only (no CPU fallback)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError(f"b200seg: tensor on cuda:{t.device.index} but the current device is cuda:{cur}; "
                               f"wrap the call in `with b200seg.on_device(tensor):`")


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return F32
    if dtype == torch.bfloat16:
        return BF16
    raise RuntimeError(f"unsupported activation dtype {dtype}")


def launches() -> int:
    """Number of kernel launches issued through this binding so far (for bench.py's gpu_launches)."""
    return _LAUNCHES[0]


_LAUNCHES = [0]


class Blocked:
    """A blocked activation buffer [N][C8][Z][Y][X][8] held in a torch tensor, plus view construction."""

    def __init__(self, n: int, c8_total: int, z: int, y: int, x: int, dtype: torch.dtype, device) -> None:
        self.n, self.c8_total, self.z, self.y, self.x = n, c8_total, z, y, x
        self.dtype = dtype
        self.tensor = torch.empty((n, c8_total, z, y, x, 8), dtype=dtype, device=device)

    def view(self, c: int, c8_off: int = 0, n: Optional[int] = None) -> View:
        if c8_off + (c + 7) // 8 > self.c8_total:
            raise RuntimeError("view exceeds buffer")
        return View(self.tensor.data_ptr(), dtype_code(self.dtype), self.n if n is None else n, c, self.c8_total,
                    c8_off, self.z, self.y, self.x)


NULL_VIEW = View(None, 0, 0, 0, 0, 0, 0, 0, 0)


def device_info():
    sm, major, minor = c_int32(), c_int32(), c_int32()
    _check(load_library().b200seg_device_info(sm, major, minor), "device_info")
    return sm.value, major.value, minor.value


def pack_ncdhw(src: torch.Tensor, dst: View) -> None:
    _require_cuda(src)
    assert src.dtype == torch.float32 and src.is_contiguous()
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_pack_ncdhw(_ptr(src), dst, _stream()), "pack_ncdhw")


def unpack_ncdhw(src: View, dst: torch.Tensor) -> None:
    _require_cuda(dst)
    assert dst.dtype == torch.float32 and dst.is_contiguous()
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_unpack_ncdhw(src, _ptr(dst), _stream()), "unpack_ncdhw")


def make_epilogue(scale: torch.Tensor, shift: torch.Tensor, slope: torch.Tensor, dst0: View = NULL_VIEW,
                  dst1: View = NULL_VIEW, split: int = 0, residual: View = NULL_VIEW,
                  out_ncdhw: Optional[torch.Tensor] = None, softmax: bool = False,
                  slope01: int = 0) -> Epilogue:
    """``slope01``: 1 = the caller guarantees every slope is in [0, 1] (enables the cheaper max(v, v*slope) form);
    2 = identity epilogue (scale 1, shift 0, slope 1 for every channel)."""
    _require_cuda(scale, shift, slope, out_ncdhw)
    return Epilogue(scale.data_ptr(), shift.data_ptr(), slope.data_ptr(), dst0, dst1, split, residual,
                    0 if out_ncdhw is None else out_ncdhw.data_ptr(), 1 if softmax else 0, int(slope01))


def conv3d_direct(inp: View, weight: torch.Tensor, cout: int, ksize: int, stride: int, pad: int, transposed: bool,
                  epi: Epilogue) -> None:
    _require_cuda(weight)
    assert weight.dtype == torch.float32 and weight.is_contiguous()
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_conv3d_direct(inp, _ptr(weight), cout, ksize, stride, pad, 1 if transposed else 0,
                                                ctypes.byref(epi), _stream()), "conv3d_direct")


def conv3d_tc(mode: int, inp: View, wpacked: torch.Tensor, cout: int, epi: Epilogue) -> None:
    _require_cuda(wpacked)
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_conv3d_tc(mode, inp, _ptr(wpacked), wpacked.numel() * wpacked.element_size(), cout,
                                            ctypes.byref(epi), _stream()), "conv3d_tc")


def conv3d_tc_wbytes(mode: int, cin_chunks: int, cout: int) -> int:
    return int(load_library().b200seg_conv3d_tc_wbytes(mode, cin_chunks, cout))


def instnorm(x: View, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor], eps: float, slope: float,
             residual: View = NULL_VIEW) -> None:
    """In-place InstanceNorm3d + activation (+ residual) on a blocked view (two streaming launches)."""
    lib = load_library()
    need = int(lib.b200seg_instnorm_scratch_bytes(x))
    scratch = torch.empty(max(need // 4, 1), dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
    _LAUNCHES[0] += 2
    _check(lib.b200seg_instnorm(x, _ptr(gamma), _ptr(beta), float(eps), float(slope), residual, _ptr(scratch), need,
                                _stream()), "instnorm")


def avgpool2(inp: View, out: View) -> None:
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_avgpool2(inp, out, _stream()), "avgpool2")


def upsample_trilinear2(inp: View, out: View) -> None:
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_upsample_trilinear2(inp, out, _stream()), "upsample_trilinear2")


def copy_view(inp: View, out: View) -> None:
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_copy_view(inp, out, _stream()), "copy_view")


def pack_ncdhw_tta(src: torch.Tensor, perm: Sequence[int], flip: Sequence[bool], dst: View) -> None:
    """src fp32 (N, C, w0, w1, w2) -> blocked ``dst`` holding src.permute(0, 1, *perm + 2).flip(axes with flip[k])."""
    _require_cuda(src)
    assert src.dtype == torch.float32 and src.is_contiguous() and src.dim() == 5
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_pack_ncdhw_tta(_ptr(src), src.shape[2], src.shape[3], src.shape[4], _i32(perm),
                                                 _i32([1 if f else 0 for f in flip]), dst, _stream()), "pack_ncdhw_tta")


def tta_accumulate(member: torch.Tensor, perm: Sequence[int], flip: Sequence[bool], acc: Optional[torch.Tensor],
                   votes: Optional[torch.Tensor]) -> None:
    """member: fp32 (N, C, S0, S1, S2) in the member's space; acc fp32 / votes uint8: (N, C, w0, w1, w2) original space."""
    _require_cuda(member, acc, votes)
    target = acc if acc is not None else votes
    assert member.dtype == torch.float32 and member.is_contiguous() and target.is_contiguous()
    assert (acc is None or acc.dtype == torch.float32) and (votes is None or votes.dtype == torch.uint8)
    n, c, w0, w1, w2 = target.shape
    assert member.shape[:2] == (n, c) and tuple(member.shape[2:]) == tuple(target.shape[2 + p] for p in perm)
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_tta_accumulate(_ptr(member), n, c, w0, w1, w2, _i32(perm),
                                                 _i32([1 if f else 0 for f in flip]), _ptr(acc), _ptr(votes), _stream()),
           "tta_accumulate")


def tta_finalize(acc: Optional[torch.Tensor], votes: Optional[torch.Tensor], onehot: Optional[torch.Tensor],
                 members: int) -> None:
    _require_cuda(acc, votes, onehot)
    target = acc if acc is not None else votes
    n, c = target.shape[:2]
    vox = target[0, 0].numel()
    assert onehot is None or (onehot.dtype == torch.int64 and onehot.is_contiguous() and onehot.shape == votes.shape)
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_tta_finalize(_ptr(acc), _ptr(votes), _ptr(onehot), n, c, vox, members, _stream()),
           "tta_finalize")


def copy_planes(patch: torch.Tensor, plane_lo: int, plane_hi: int, dst: torch.Tensor) -> None:
    """patch: fp32 (C, p0, p1, p2) contiguous; dst: fp32 buffer with room for (C, plane_hi - plane_lo, p1, p2)."""
    _require_cuda(patch, dst)
    assert patch.dtype == torch.float32 and dst.dtype == torch.float32 and patch.is_contiguous() and dst.is_contiguous()
    c, p0, p1, p2 = patch.shape
    assert dst.numel() >= c * (plane_hi - plane_lo) * p1 * p2
    _check(load_library().b200seg_copy_planes(_ptr(patch), c, p0, p1, p2, plane_lo, plane_hi, _ptr(dst), _stream()),
           "copy_planes")


def softmax_ncdhw(data: torch.Tensor, sm_channels: int = 0, diag_bias: float = 0.0) -> None:
    """In-place channel softmax on fp32 (N, C, ...) data; ``sm_channels`` = C of a StochasticMatrix head."""
    _require_cuda(data)
    assert data.dtype == torch.float32 and data.is_contiguous()
    n, c = data.shape[:2]
    vox = data[0, 0].numel()
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_softmax_ncdhw(_ptr(data), n, c, vox, sm_channels, float(diag_bias), _stream()),
           "softmax_ncdhw")


def _i32(values: Sequence[int]):
    arr = (c_int32 * len(values))(*[int(v) for v in values])
    return arr


def grid_extract(volume: torch.Tensor, locations: Sequence[Sequence[int]], border: Sequence[int], pad_mode: int,
                 pad_value: float, dst: View) -> None:
    """volume: fp32 (C, W, H, D) on the device; locations: host rows [i0,j0,k0,i1,j1,k1] in padded coordinates."""
    _require_cuda(volume)
    assert volume.dtype == torch.float32 and volume.is_contiguous() and volume.dim() == 4
    c, w, h, d = volume.shape
    flat = [int(v) for row in locations for v in row]
    _LAUNCHES[0] += (len(locations) + 63) // 64
    _check(load_library().b200seg_grid_extract(_ptr(volume), c, w, h, d, _i32(flat), len(locations), _i32(border),
                                               pad_mode, float(pad_value), dst, _stream()), "grid_extract")


def overlap_add(out: torch.Tensor, patches: torch.Tensor, locations: Sequence[Sequence[int]]) -> None:
    """out: fp32 (C, PW, PH, PD) accumulated in place; patches: fp32 (B, C, p0, p1, p2); host locations."""
    _require_cuda(out, patches)
    assert out.dtype == torch.float32 and patches.dtype == torch.float32
    assert out.is_contiguous() and patches.is_contiguous()
    c, pw, ph, pd = out.shape
    assert patches.shape[1] == c and patches.shape[0] >= len(locations)
    flat = [int(v) for row in locations for v in row]
    _LAUNCHES[0] += (len(locations) + 63) // 64
    _check(load_library().b200seg_overlap_add(_ptr(out), c, pw, ph, pd, _ptr(patches), _i32(flat), len(locations),
                                              _stream()), "overlap_add")


def window_patches(patches: torch.Tensor, windows: Sequence[torch.Tensor], count: Optional[int] = None) -> None:
    """patches (B, C, p0, p1, p2) fp32 *= outer product of the three per-axis fp32 windows, in place ('hann' mode)."""
    _require_cuda(patches, *windows)
    assert patches.dtype == torch.float32 and patches.is_contiguous() and patches.dim() == 5
    b, c, p0, p1, p2 = patches.shape
    for w, size in zip(windows, (p0, p1, p2)):
        assert w.dtype == torch.float32 and w.is_contiguous() and w.numel() == size
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_window_patches(_ptr(patches), b if count is None else count, c, p0, p1, p2,
                                                 _ptr(windows[0]), _ptr(windows[1]), _ptr(windows[2]), _stream()),
           "window_patches")


def divide_separable(out: torch.Tensor, sums: Sequence[torch.Tensor]) -> None:
    """out (C, PW, PH, PD) fp32 /= outer product of the three per-axis fp32 vectors, in place ('hann' mode)."""
    _require_cuda(out, *sums)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.dim() == 4
    for v, size in zip(sums, out.shape[1:]):
        assert v.dtype == torch.float32 and v.is_contiguous() and v.numel() == size
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_divide_separable(_ptr(out), *out.shape, _ptr(sums[0]), _ptr(sums[1]), _ptr(sums[2]),
                                                   _stream()), "divide_separable")


def overlap_crop(out: torch.Tensor, patches: torch.Tensor, locations: Sequence[Sequence[int]],
                 border: Sequence[int], volume_padded: bool) -> None:
    _require_cuda(out, patches)
    assert out.dtype == torch.float32 and patches.dtype == torch.float32
    assert out.is_contiguous() and patches.is_contiguous()
    c, pw, ph, pd = out.shape
    flat = [int(v) for row in locations for v in row]
    _LAUNCHES[0] += len(locations)
    _check(load_library().b200seg_overlap_crop(_ptr(out), c, pw, ph, pd, _ptr(patches), _i32(flat), len(locations),
                                               _i32(border), 1 if volume_padded else 0, _stream()), "overlap_crop")


def finalize(out: torch.Tensor, counts, border: Sequence[int], probs: Optional[torch.Tensor],
             labels_i64: Optional[torch.Tensor], labels_u8: Optional[torch.Tensor]) -> None:
    """counts: None (crop mode) or three int32 device tensors with the per-axis coverage of the padded grid."""
    _require_cuda(out, probs, labels_i64, labels_u8)
    c, pw, ph, pd = out.shape
    cw = ch = cd = None
    if counts is not None:
        cw, ch, cd = counts
        for t in (cw, ch, cd):
            assert t.dtype == torch.int32 and t.is_cuda
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_finalize(_ptr(out), c, pw, ph, pd, _ptr(cw), _ptr(ch), _ptr(cd), _i32(border),
                                           _ptr(probs), _ptr(labels_i64), _ptr(labels_u8), _stream()), "finalize")


def finalize_region(out: torch.Tensor, counts, offset: Sequence[int], extent: Sequence[int],
                    probs: Optional[torch.Tensor], labels_i64: Optional[torch.Tensor],
                    labels_u8: Optional[torch.Tensor], count_offset: int = 0) -> None:
    """finalize() for the sub-region [offset, offset + extent) of the accumulator; ``count_offset`` is the global
    padded coordinate of the accumulator's first plane along axis 0 (z-slab mode)."""
    _require_cuda(out, probs, labels_i64, labels_u8)
    c, pw, ph, pd = out.shape
    cw = ch = cd = None
    cw_ptr = c_void_p(0)
    if counts is not None:
        cw, ch, cd = counts
        for t in (cw, ch, cd):
            assert t.dtype == torch.int32 and t.is_cuda
        cw_ptr = c_void_p(cw.data_ptr() + 4 * int(count_offset))
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_finalize_region(_ptr(out), c, pw, ph, pd, cw_ptr, _ptr(ch), _ptr(cd), _i32(offset),
                                                  _i32(extent), _ptr(probs), _ptr(labels_i64), _ptr(labels_u8),
                                                  _stream()), "finalize_region")


def argmax(probs: torch.Tensor, labels_i64: Optional[torch.Tensor] = None,
           labels_u8: Optional[torch.Tensor] = None) -> None:
    """probs: fp32 (C, ...) contiguous; labels: int64 / uint8 tensors with prod(...) elements."""
    _require_cuda(probs, labels_i64, labels_u8)
    assert probs.dtype == torch.float32 and probs.is_contiguous()
    c = probs.shape[0]
    vox = probs[0].numel()
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_argmax(_ptr(probs), c, vox, _ptr(labels_i64), _ptr(labels_u8), _stream()), "argmax")


def confusion(pred: torch.Tensor, target: torch.Tensor, num_classes: int, cm: torch.Tensor) -> None:
    """Accumulates the (num_classes x num_classes) int64 joint histogram cm[target][pred]."""
    _require_cuda(pred, target, cm)
    assert pred.dtype == target.dtype and pred.dtype in (torch.uint8, torch.int64)
    assert pred.is_contiguous() and target.is_contiguous() and pred.numel() == target.numel()
    assert cm.dtype == torch.int64 and cm.numel() == num_classes * num_classes and cm.is_contiguous()
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_confusion(_ptr(pred), _ptr(target), pred.element_size(), pred.numel(), num_classes,
                                            _ptr(cm), _stream()), "confusion")


def hybrid_loss_forward(prediction: torch.Tensor, target: torch.Tensor, dice_weight: float,
                        class_weights: Optional[torch.Tensor], square_dice: bool):
    """-> (out3 fp32 [loss, dice_loss, logistic_loss], sums fp32 (N*C, 4)) on the device."""
    _require_cuda(prediction, target, class_weights)
    assert prediction.dtype == torch.float32 and target.dtype == torch.float32
    assert prediction.is_contiguous() and target.is_contiguous() and prediction.shape == target.shape
    n, c = prediction.shape[:2]
    vox = prediction[0, 0].numel()
    lib = load_library()
    need = int(lib.b200seg_hybrid_loss_scratch_bytes(n, c))
    scratch = torch.empty(need // 8, dtype=torch.float64, device=prediction.device)
    sums = torch.empty((n * c, 4), dtype=torch.float32, device=prediction.device)
    out3 = torch.empty(3, dtype=torch.float32, device=prediction.device)
    _LAUNCHES[0] += 2
    _check(lib.b200seg_hybrid_loss_forward(_ptr(prediction), _ptr(target), n, c, vox, float(dice_weight),
                                           _ptr(class_weights), 1 if square_dice else 0, _ptr(scratch), need, _ptr(sums),
                                           _ptr(out3), _stream()), "hybrid_loss_forward")
    return out3, sums


def hybrid_loss_backward(prediction: torch.Tensor, target: torch.Tensor, sums: torch.Tensor, dice_weight: float,
                         class_weights: Optional[torch.Tensor], square_dice: bool, grad_loss: torch.Tensor) -> torch.Tensor:
    _require_cuda(prediction, target, sums, grad_loss, class_weights)
    n, c = prediction.shape[:2]
    vox = prediction[0, 0].numel()
    grad = torch.empty_like(prediction)
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_hybrid_loss_backward(_ptr(prediction), _ptr(target), _ptr(sums), n, c, vox,
                                                       float(dice_weight), _ptr(class_weights), 1 if square_dice else 0,
                                                       _ptr(grad_loss), _ptr(grad), _stream()), "hybrid_loss_backward")
    return grad


def connected_components(mask: torch.Tensor, connectivity: int = 2, max_components: int = 1 << 20,
                         by_value: int = 0):
    """mask: (W, H, D) uint8 / int32 / int64 on the device (foreground = values > 0) -> (labels int32 (W, H, D) with
    components numbered 1..N in raster-scan order like skimage.morphology.label, N).  ``by_value``: label an integer
    image like skimage.measure.label (background 0, only equal-valued neighbours connect); ``by_value=2``: label the
    inverted mask (values <= 0)."""
    _require_cuda(mask)
    assert mask.dim() == 3 and mask.is_contiguous() and mask.dtype in (torch.uint8, torch.int32, torch.int64)
    w, h, d = mask.shape
    dev = mask.device
    parent = torch.empty((w, h, d), dtype=torch.int32, device=dev)
    roots = torch.empty(max_components, dtype=torch.int32, device=dev)
    n_roots = torch.zeros(1, dtype=torch.int32, device=dev)
    lib = load_library()
    _LAUNCHES[0] += 3
    _check(lib.b200seg_ccl3d_roots(_ptr(mask), mask.element_size(), w, h, d, int(connectivity), int(by_value),
                                   _ptr(parent), _ptr(roots), max_components, _ptr(n_roots), _stream()), "ccl3d_roots")
    n = int(n_roots.item())
    if n > max_components:
        raise RuntimeError(f"connected_components: {n} components exceed max_components = {max_components}")
    sorted_roots = torch.sort(roots[:n]).values.contiguous()        # a few hundred integers: plumbing
    labels = torch.empty((w, h, d), dtype=torch.int32, device=dev)
    _LAUNCHES[0] += 1
    _check(lib.b200seg_ccl3d_relabel(_ptr(parent), parent.numel(), _ptr(sorted_roots), n, _ptr(labels), _stream()),
           "ccl3d_relabel")
    return labels, n


def overlap_histogram(target: torch.Tensor, pred: Optional[torch.Tensor], n_target: int, n_pred: int) -> torch.Tensor:
    """int64 (n_target + 1, n_pred + 1) table of voxel counts per (target component, predicted component) pair;
    ``pred=None, n_pred=0``: the voxel count of every label 0..n_target."""
    _require_cuda(target, pred)
    assert target.dtype == torch.int32 and target.is_contiguous()
    assert pred is None or (pred.dtype == torch.int32 and pred.is_contiguous() and target.numel() == pred.numel())
    hist = torch.zeros((n_target + 1, n_pred + 1), dtype=torch.int64, device=target.device)
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_overlap_histogram(_ptr(target), _ptr(pred), target.numel(), n_target, n_pred,
                                                    _ptr(hist), _stream()), "overlap_histogram")
    return hist


def relabel_lut(src: torch.Tensor, lut: torch.Tensor) -> torch.Tensor:
    """int32 labels -> lut[labels] (int32 lookup table on the device)."""
    _require_cuda(src, lut)
    assert src.dtype == torch.int32 and lut.dtype == torch.int32 and src.is_contiguous() and lut.is_contiguous()
    dst = torch.empty_like(src)
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_relabel_lut(_ptr(src), src.numel(), _ptr(lut), lut.numel(), _ptr(dst), _stream()),
           "relabel_lut")
    return dst


def dilate_cross(src: torch.Tensor) -> torch.Tensor:
    """Grey dilation of an int32 (W, H, D) volume with the 3-D cross (skimage.morphology.dilation's default)."""
    _require_cuda(src)
    assert src.dtype == torch.int32 and src.dim() == 3 and src.is_contiguous()
    dst = torch.empty_like(src)
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_dilate_cross(_ptr(src), *src.shape, _ptr(dst), _stream()), "dilate_cross")
    return dst


def _require_i32(*tensors):
    _require_cuda(*tensors)
    for t in tensors:
        assert t.dtype == torch.int32 and t.is_contiguous()


def dilate_where(dil_src: torch.Tensor, mask: torch.Tensor, pass_src: torch.Tensor) -> torch.Tensor:
    """where(mask & (D != dil_src), D, pass_src) with D = dilate_cross(dil_src), one kernel."""
    _require_i32(dil_src, mask, pass_src)
    assert dil_src.dim() == 3 and dil_src.shape == mask.shape == pass_src.shape
    dst = torch.empty_like(dil_src)
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_dilate_where(_ptr(dil_src), _ptr(mask), _ptr(pass_src), *dil_src.shape, _ptr(dst),
                                               _stream()), "dilate_where")
    return dst


def relabel_masked(img: torch.Tensor, lut: torch.Tensor, comp: torch.Tensor, keep_lut: torch.Tensor) -> torch.Tensor:
    """where(keep_lut[comp], lut[img], 0)."""
    _require_i32(img, lut, comp, keep_lut)
    assert img.shape == comp.shape
    dst = torch.empty_like(img)
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_relabel_masked(_ptr(img), _ptr(lut), lut.numel(), _ptr(comp), _ptr(keep_lut),
                                                 keep_lut.numel(), img.numel(), _ptr(dst), _stream()), "relabel_masked")
    return dst


def label_equals(src: torch.Tensor, value: int) -> torch.Tensor:
    _require_i32(src)
    dst = torch.empty_like(src)
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_label_equals(_ptr(src), src.numel(), int(value), _ptr(dst), _stream()), "label_equals")
    return dst


def mask_assign(dst: torch.Tensor, mask: torch.Tensor, value: int) -> torch.Tensor:
    _require_i32(dst, mask)
    assert dst.shape == mask.shape
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_mask_assign(_ptr(dst), _ptr(mask), dst.numel(), int(value), _stream()), "mask_assign")
    return dst


# ------------------------------------------------------------------------------------------------- training step
def _train_scratch(channels: int, device) -> torch.Tensor:
    return torch.empty(load_library().b200seg_train_scratch_bytes(int(channels)) // 8, dtype=torch.float64, device=device)


def channel_moments(x: View, channels: int, device):
    """(mean, biased var) fp32 vectors of length round_up(channels, 8) over (N, Z, Y, X) of a blocked fp32 view."""
    cpad = (channels + 7) // 8 * 8
    mean = torch.empty(cpad, dtype=torch.float32, device=device)
    var = torch.empty(cpad, dtype=torch.float32, device=device)
    scratch = _train_scratch(channels, device)
    _LAUNCHES[0] += 2
    _check(load_library().b200seg_channel_moments(x, _ptr(scratch), _ptr(mean), _ptr(var), _stream()), "channel_moments")
    return mean, var


def affine_act(src: View, scale: torch.Tensor, shift: torch.Tensor, slope: torch.Tensor, dst: View,
               residual: View = NULL_VIEW) -> None:
    _require_cuda(scale, shift, slope)
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_affine_act(src, _ptr(scale), _ptr(shift), _ptr(slope), residual, dst, _stream()),
           "affine_act")


def bn_backward(dy: View, z: View, scale, shift, slope, mean, rstd, has_norm: bool, dz: View, channels: int, device):
    """-> (sum_g, sum_gx) fp32 vectors (d beta, d gamma); writes dz."""
    _require_cuda(scale, shift, slope, mean, rstd)
    cpad = (channels + 7) // 8 * 8
    sum_g = torch.empty(cpad, dtype=torch.float32, device=device)
    sum_gx = torch.empty(cpad, dtype=torch.float32, device=device)
    scratch = _train_scratch(channels, device)
    _LAUNCHES[0] += 3
    _check(load_library().b200seg_bn_backward(dy, z, _ptr(scale), _ptr(shift), _ptr(slope), _ptr(mean), _ptr(rstd),
                                              1 if has_norm else 0, _ptr(scratch), _ptr(sum_g), _ptr(sum_gx), dz,
                                              _stream()), "bn_backward")
    return sum_g, sum_gx


def softmax_backward(probs: torch.Tensor, dprobs: torch.Tensor, softmax: bool, dst: View) -> None:
    _require_cuda(probs, dprobs)
    assert probs.dtype == torch.float32 and dprobs.dtype == torch.float32 and probs.shape == dprobs.shape
    assert probs.is_contiguous() and dprobs.is_contiguous()
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_softmax_backward(_ptr(probs), _ptr(dprobs), probs.shape[0], probs.shape[1],
                                                   1 if softmax else 0, dst, _stream()), "softmax_backward")


def wgrad(a: View, b: View, ksize: int, stride: int, pad: int, device) -> torch.Tensor:
    """-> fp32 (k^3, round_up(a.c, 8), round_up(b.c, 8)): sum over positions of A[a](pos) * B[b](stride*pos + tap - pad)."""
    lib = load_library()
    scratch = torch.empty(lib.b200seg_wgrad_scratch_floats(a.c, b.c, ksize), dtype=torch.float32, device=device)
    grad = torch.empty((ksize ** 3, (a.c + 7) // 8 * 8, (b.c + 7) // 8 * 8), dtype=torch.float32, device=device)
    _LAUNCHES[0] += 2
    _check(lib.b200seg_wgrad(a, b, ksize, stride, pad, _ptr(scratch), _ptr(grad), _stream()), "wgrad")
    return grad


def avgpool2_backward(dy: View, dx: View, add: View = NULL_VIEW) -> None:
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_avgpool2_backward(dy, add, dx, _stream()), "avgpool2_backward")


def upsample_trilinear2_backward(dy: View, dx: View) -> None:
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_upsample_trilinear2_backward(dy, dx, _stream()), "upsample_trilinear2_backward")


def channel_scale(src: View, mask: torch.Tensor, dst: View) -> None:
    """dst = src * mask[n][c]; mask fp32 (N, round_up(C, 8)) on the device (Dropout3d)."""
    _require_cuda(mask)
    assert mask.dtype == torch.float32 and mask.is_contiguous() and mask.dim() == 2
    _LAUNCHES[0] += 1
    _check(load_library().b200seg_channel_scale(src, _ptr(mask), dst, _stream()), "channel_scale")
