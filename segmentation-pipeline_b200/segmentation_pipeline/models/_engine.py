"""Execution engine of the native plans: packs parameters for the device once, owns the activation
workspaces, and replays the op list through libb200seg (no ATen compute, no CPU fallback).

Precision
    'fp32'  direct CUDA-core kernels, fp32 activations (logits within 1e-5 of the reference)
    'bf16'  tcgen05 tensor-core engine, bf16 activations, fp32 accumulation (the throughput path)
    'auto'  (default) bf16 when the module's parameters or the input are bf16, or under
            ``torch.autocast('cuda', dtype=torch.bfloat16)``; fp32 otherwise.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _plan
from ._plan import DOWN, K3, UP, ConvOp, Plan, PoolOp, Ref, SoftmaxOp, UpsampleOp, c8

_PRECISION = [os.environ.get("B200SEG_PRECISION", "auto")]
TRACE = None    # set to a list to record (call, extent, (start, end) CUDA events) per op (tools/profile_layers.py)
TC_MAX_COUT = 80           # output channels per launch on the large levels (3 z-planes per accumulator set)
TC_MAX_COUT_DEEP = 120     # deep levels (>= 3): one launch for 120 channels -- they are launch-bound, not MMA-bound
if os.environ.get("B200SEG_TC_VARIANT") == "1":
    TC_MAX_COUT_DEEP = 80  # the two-CTAs-per-SM test variant has half the shared memory: a 120-wide weight image does not fit


_NO_K3T = os.environ.get("B200SEG_TC_NO_K3T", "0") == "1"     # profiling hook: final layer through the plain K3 mode

# CUDA graphs: the op list of a plan is a fixed sequence of ~55-160 launches over fixed buffers.  On small inputs
# (patch_batch_size = 1 as in the reference's ms-inference.py:32, the 48 x 88 x 24 half volumes of dmri_hippo) a launch
# lasts a few microseconds while Python + ctypes + tensor-map encoding cost tens per call, so the host is the bottleneck:
# there the whole list is captured once per workspace and replayed.  Large batches are GPU-bound and stay eager (their
# launches can then be timed individually).  B200SEG_GRAPHS=0 disables, =all forces graphs for every size.
_GRAPHS = os.environ.get("B200SEG_GRAPHS", "1")
GRAPH_MAX_VOXELS = 4 * 96 ** 3


def set_precision(mode: str) -> None:
    if mode not in ("auto", "fp32", "bf16"):
        raise ValueError("precision must be 'auto', 'fp32' or 'bf16'")
    _PRECISION[0] = mode


def get_precision() -> str:
    return _PRECISION[0]


def _b200seg():
    import b200seg  # the ctypes binding; raises if libb200seg.so is missing
    b200seg.load_library()
    return b200seg


def _resolve_precision(module: nn.Module, x: torch.Tensor) -> str:
    mode = _PRECISION[0]
    if mode != "auto":
        return mode
    if x.dtype == torch.bfloat16:
        return "bf16"
    first = next(module.parameters(), None)
    if first is not None and first.dtype == torch.bfloat16:
        return "bf16"
    try:
        autocast_bf16 = torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16
    except TypeError:  # older signature
        autocast_bf16 = torch.is_autocast_enabled() and torch.get_autocast_gpu_dtype() == torch.bfloat16
    if autocast_bf16:
        return "bf16"
    return "fp32"


def lower(module: nn.Module) -> Plan:
    from .modular_unet import ModularUNet
    from .nested_residual_unet import NestedResUNet
    if isinstance(module, ModularUNet):
        return _plan.lower_modular_unet(module)
    if isinstance(module, NestedResUNet):
        return _plan.lower_nested_res_unet(module)
    if isinstance(module, NestedResUNet.Block):
        cin, cout = module.conv1.in_channels, module.out_ch
        plan = Plan(cin, cout)
        plan.add_buffer("in", cin, 0)
        out = Ref(plan.add_buffer("out", cout, 0), 0, cout)
        _plan._lower_block(plan, 0, "block", Ref("in", 0, cin), [(0, cin)], [module.conv1, module.conv2],
                           [module.bn1, module.bn2], [module.activation1, module.activation2],
                           module.res_conv if module.residual else None, out)
        return plan
    return _plan.lower_single(module)


class _ConvCall:
    """One kernel launch of a ConvOp (a ConvOp wider than the engine's N limit becomes several)."""
    __slots__ = ("tc", "mode", "src", "weight", "cout", "scale", "shift", "slope", "dst0", "dst1", "split",
                 "residual", "final", "softmax", "ksize", "stride", "pad", "transposed", "name", "slope01")


class CompiledPlan:
    def __init__(self, plan: Plan, precision: str, device: torch.device):
        self.plan = plan
        self.precision = precision
        self.device = device
        self.act_dtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.calls: List[object] = []
        self.workspaces: Dict[Tuple[int, int, int, int], Dict[str, object]] = {}
        self._compile()

    # ------------------------------------------------------------------ compile: pack weights, split wide convs
    def _dev(self, array) -> torch.Tensor:
        return torch.as_tensor(array, dtype=torch.float32).contiguous().to(self.device)

    def _padded(self, array, n):
        out = torch.zeros(n, dtype=torch.float32)
        out[:len(array)] = torch.as_tensor(array, dtype=torch.float32)
        return out

    def _compile(self) -> None:
        for op in self.plan.ops:
            if isinstance(op, ConvOp):
                self._compile_conv(op)
            else:
                self.calls.append(op)

    def _compile_conv(self, op: ConvOp) -> None:
        n_chunks = c8(op.src.c)
        use_tc = self.precision == "bf16"
        geom = {K3: (3, 1, 1, False), DOWN: (4, 2, 1, False), UP: (4, 2, 1, True)}[op.mode]
        if not use_tc:
            call = self._new_call(op, False, geom)
            call.weight = _plan.pack_direct_weight(op, n_chunks).to(self.device)
            call.cout = op.cout
            cpad = c8(op.cout) * 8
            call.scale, call.shift, call.slope = (self._dev(self._padded(a, cpad)) for a in
                                                   (op.scale, op.shift, _slope_pad(op.slope, cpad)))
            call.dst0, call.dst1, call.split, call.residual = op.dst0, op.dst1, op.split, op.residual
            self.calls.append(call)
            return
        # tensor-core path: pieces of <= TC_MAX_COUT output channels; a fused two-destination conv stays one
        # launch when it fits, otherwise each destination gets its own launches
        pieces = []
        level = self.plan.buffers[op.src.buf][1]
        TC_MAX_COUT = TC_MAX_COUT_DEEP if level >= 3 else globals()["TC_MAX_COUT"]
        if op.dst1 is not None and op.cout <= TC_MAX_COUT:
            pieces.append((0, op.cout, op.dst0, op.dst1, op.split, op.residual))
        else:
            parts = [(0, op.cout, op.dst0, op.residual)] if op.dst1 is None else \
                [(0, op.split, op.dst0, op.residual), (op.split, op.cout, op.dst1, None)]
            for lo, hi, dst, res in parts:
                for p_lo in range(lo, hi, TC_MAX_COUT):
                    p_hi = min(p_lo + TC_MAX_COUT, hi)
                    d = None if dst is None else Ref(dst.buf, dst.off + (p_lo - lo) // 8, p_hi - p_lo)
                    r = None if res is None else Ref(res.buf, res.off + (p_lo - lo) // 8, p_hi - p_lo)
                    pieces.append((p_lo, p_hi, d, None, 0, r))
        if op.final and len(pieces) != 1:
            raise _plan.UnsupportedModule("final convolution wider than the tensor-core N limit")
        for lo, hi, d0, d1, split, res in pieces:
            call = self._new_call(op, True, geom)
            phys = _plan.physical_weight(op.weight, op.mode == UP, op.segments, n_chunks, lo, hi)
            if op.final and op.mode == K3 and hi - lo <= _plan.K3T_MAX_COUT and n_chunks <= 12 and not _NO_K3T:
                call.mode = _plan.K3T      # tiny-Cout final layer: taps as accumulator columns (see conv_tc.cu)
            call.weight = _plan.pack_tc_weight(call.mode, phys, n_chunks, hi - lo).to(self.device)
            call.cout = hi - lo
            cpad = c8(hi - lo) * 8
            call.scale, call.shift, call.slope = (self._dev(self._padded(a[lo:hi], cpad)) for a in
                                                   (op.scale, op.shift, op.slope))
            call.slope = self._dev(_slope_pad(op.slope[lo:hi], cpad))
            call.dst0, call.dst1, call.split, call.residual = d0, d1, split, res
            self.calls.append(call)

    def _norm_params(self, op):
        cache = self.__dict__.setdefault("_norm_cache", {})
        hit = cache.get(id(op))
        if hit is None:
            hit = (None if op.gamma is None else self._dev(op.gamma), None if op.beta is None else self._dev(op.beta))
            cache[id(op)] = hit
        return hit

    def _new_call(self, op: ConvOp, tc: bool, geom) -> _ConvCall:
        call = _ConvCall()
        call.tc, call.mode, call.src = tc, op.mode, op.src
        call.ksize, call.stride, call.pad, call.transposed = geom
        call.final, call.softmax, call.name = op.final, op.softmax, op.name
        call.slope01 = int(bool(((op.slope >= 0.0) & (op.slope <= 1.0)).all()))   # ReLU / LeakyReLU / none
        if bool((op.scale == 1.0).all() and (op.shift == 0.0).all() and (op.slope == 1.0).all()):
            call.slope01 = 2                                                   # identity epilogue (blur convolutions)
        return call

    # ------------------------------------------------------------------ workspaces
    def _workspace(self, n: int, z: int, y: int, x: int) -> Dict[str, object]:
        key = (n, z, y, x)
        ws = self.workspaces.get(key)
        if ws is None:
            lib = _b200seg()
            div = 1 << (self.plan.levels - 1)
            if z % div or y % div or x % div:
                raise RuntimeError(f"spatial extent {(z, y, x)} must be divisible by {div} for this network")
            ws = {}
            for name, (chunks, level) in self.plan.buffers.items():
                if level >= 0:
                    ext = (z >> level, y >> level, x >> level)
                else:
                    ext = (z << -level, y << -level, x << -level)
                buf = lib.Blocked(n, chunks, *ext, self.act_dtype, self.device)
                # padding channels / never-written chunks must read as zero (they multiply zero weights, but
                # uninitialised memory may hold NaN bit patterns)
                buf.tensor.zero_()
                ws[name] = buf
            if len(self.workspaces) >= 8:       # 6 orientations of a non-cubic patch + ragged last batches
                self.workspaces.pop(next(iter(self.workspaces)))
            self.workspaces[key] = ws
        return ws

    def workspace_bytes(self, n: int, z: int, y: int, x: int) -> int:
        """Device memory of the activation workspace of one (n, z, y, x) input (+ its fp32 output)."""
        esize = 2 if self.act_dtype == torch.bfloat16 else 4
        total = 0
        for chunks, level in self.plan.buffers.values():
            vox = (z * y * x) >> (3 * level) if level >= 0 else (z * y * x) << (-3 * level)
            total += chunks * 8 * vox * esize
        return n * (total + self.plan.out_channels * z * y * x * 4)

    def input_buffer(self, n: int, z: int, y: int, x: int):
        """The blocked input buffer of a workspace (so that the grid sampler can extract patches into it)."""
        return self._workspace(n, z, y, x)["in"]

    # ------------------------------------------------------------------ run
    def run_blocked(self, n: int, z: int, y: int, x: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Runs the plan on the workspace whose 'in' buffer has already been filled; returns fp32 NCDHW.

        Without ``out`` the result lives in a buffer owned by the workspace (valid until the next run on it) -- the
        sliding-window predictor consumes it immediately; small inputs replay a captured CUDA graph."""
        ws = self._workspace(n, z, y, x)
        s = self.plan.out_scale
        oz, oy, ox = ((z << s, y << s, x << s) if s >= 0 else (z >> -s, y >> -s, x >> -s))
        if out is not None:
            self._launch_all(ws, out, (n, z, y, x))
            return out
        static = ws.get("__out__")
        if static is None:
            static = ws["__out__"] = torch.empty((n, self.plan.out_channels, oz, oy, ox), dtype=torch.float32,
                                                 device=self.device)
        graph_ok = _GRAPHS != "0" and TRACE is None and (_GRAPHS == "all" or n * z * y * x <= GRAPH_MAX_VOXELS) \
            and not torch.cuda.is_current_stream_capturing()
        if not graph_ok:
            self._launch_all(ws, static, (n, z, y, x))
            return static
        graph = ws.get("__graph__")
        if graph is None:
            # first call: eager (loads kernels, sets function attributes); second call: capture
            if not ws.get("__warm__"):
                ws["__warm__"] = True
                self._launch_all(ws, static, (n, z, y, x))
                return static
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._launch_all(ws, static, (n, z, y, x))
            ws["__graph__"] = graph
        graph.replay()
        return static

    def _launch_all(self, ws, out: torch.Tensor, extent) -> None:
        lib = _b200seg()
        n, z, y, x = extent
        wrote_final = False

        def view(ref: Optional[Ref]):
            if ref is None:
                return lib.NULL_VIEW
            return ws[ref.buf].view(ref.c, ref.off)

        for call in self.calls:
            if TRACE is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
                TRACE.append((call, (n, z, y, x), ev))
            if isinstance(call, _ConvCall):
                epi = lib.make_epilogue(call.scale, call.shift, call.slope, view(call.dst0), view(call.dst1),
                                        call.split, view(call.residual), out if call.final else None, call.softmax,
                                        call.slope01)
                if call.tc:
                    lib.conv3d_tc(call.mode, view(call.src), call.weight, call.cout, epi)
                else:
                    lib.conv3d_direct(view(call.src), call.weight, call.cout, call.ksize, call.stride, call.pad,
                                      call.transposed, epi)
                wrote_final = wrote_final or call.final
            elif isinstance(call, PoolOp):
                lib.avgpool2(view(call.src), view(call.dst))
            elif isinstance(call, UpsampleOp):
                lib.upsample_trilinear2(view(call.src), view(call.dst))
            elif isinstance(call, _plan.UnpackOp):
                lib.unpack_ncdhw(ws["out"].view(self.plan.out_channels), out)
                wrote_final = True
            elif isinstance(call, SoftmaxOp):
                lib.softmax_ncdhw(out, call.sm_channels, call.diag_bias)
            elif isinstance(call, _plan.InstNormOp):
                gamma, beta = self._norm_params(call)
                lib.instnorm(view(call.ref), gamma, beta, call.eps, call.slope, view(call.residual))
            else:  # pragma: no cover
                raise RuntimeError(f"unknown op {call}")
            if TRACE is not None:
                ev[1].record()
        if not wrote_final:
            lib.unpack_ncdhw(ws["out"].view(self.plan.out_channels), out)

    def run(self, x: torch.Tensor) -> torch.Tensor:
        lib = _b200seg()
        if x.dim() != 5 or x.shape[1] != self.plan.in_channels:
            raise RuntimeError(f"expected input (N, {self.plan.in_channels}, W, H, D), got {tuple(x.shape)}")
        n, _, z, y, xx = x.shape
        src = x.detach().to(torch.float32).contiguous()
        ws = self._workspace(n, z, y, xx)
        lib.pack_ncdhw(src, ws["in"].view(self.plan.in_channels))
        return self.run_blocked(n, z, y, xx).clone()      # the workspace keeps its own output buffer


def _slope_pad(slope, cpad):
    """Padding channels keep the identity activation so that they stay exactly zero."""
    out = torch.ones(cpad, dtype=torch.float32)
    out[:len(slope)] = torch.as_tensor(slope, dtype=torch.float32)
    return out


def _fingerprint(module: nn.Module):
    return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))


def compiled_for(module: nn.Module, precision: str, device: torch.device) -> CompiledPlan:
    cache = module.__dict__.setdefault("_b200_cache", {})
    key = (precision, str(device))
    fp = _fingerprint(module)
    hit = cache.get(key)
    if hit is not None and hit[0] == fp:
        return hit[1]
    compiled = CompiledPlan(lower(module), precision, device)
    cache[key] = (fp, compiled)
    return compiled


def forward_native(module: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """The ``forward`` of every mirrored module.  ``model.eval()``: the inference plan.  ``model.train()``: a
    ``ModularUNet`` / ``NestedResUNet`` runs the device training step (``_train.py``: batch-statistic BatchNorm, autograd); every other
    module with BatchNorm / Dropout raises instead of silently computing something else."""
    from .components import StochasticMatrix
    from .modular_unet import ModularUNet
    from .nested_residual_unet import NestedResUNet
    if not x.is_cuda:
        raise RuntimeError("segmentation_pipeline.models (b200) runs on CUDA tensors only: there is no CPU "
                           "fallback on the product path (the CPU oracle lives under oracle/ for tests)")
    if module.training and isinstance(module, (ModularUNet, NestedResUNet)):
        from . import _train
        return _train.forward_train(module, x)
    if module.training and any(isinstance(m, (nn.modules.batchnorm._BatchNorm, nn.Dropout3d))
                               for m in module.modules()):
        raise NotImplementedError("training-mode forward (batch statistics / dropout / autograd) of this module is not "
                                  "lowered (ModularUNet and NestedResUNet are); call model.eval()")
    lib = _b200seg()
    # the caller's current device may be another GPU: launch on the stream of the device that holds x
    with lib.on_device(x):
        if isinstance(module, StochasticMatrix):
            out = x.detach().to(torch.float32).contiguous().clone()
            if out.shape[1] != module.channels ** 2:
                raise RuntimeError("Expected dim 1 of input tensor to be the square of the number of out channels")
            lib.softmax_ncdhw(out, module.channels, float(module.diag_bias or 0.0))
            return out
        first = next(module.parameters(), None)
        if first is not None and first.device != x.device:
            raise RuntimeError(f"module parameters live on {first.device} but the input is on {x.device}")
        precision = _resolve_precision(module, x)
        y = compiled_for(module, precision, x.device).run(x)
        return y if x.dtype == torch.float32 else y.to(x.dtype)
