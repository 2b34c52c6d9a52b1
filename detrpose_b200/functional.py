"""Host-side mirror of the reference's sampling core over the C-ABI kernels.

``ms_deform_attn_core`` has the argument meaning of the reference's
``ms_deform_attn_core_pytorch(value, value_spatial_shapes, sampling_locations,
attention_weights)`` (/root/reference/src/models/detrpose/ms_deform_attn.py:145-193):
``value`` is the list of per-level tensors ``(N*H, Dh, H_l*W_l)`` with
arbitrary strides that ``Transformer.forward`` builds (transformer.py:1285-1286),
the result is ``(N, Lq, H*Dh)``, and it is differentiable w.r.t. the value
list, the locations and the attention weights.

PyTorch is used here for device memory, streams and autograd plumbing only;
all arithmetic runs in ``libmsda_b200.so``.  There is no fallback path.
"""
from __future__ import annotations

import weakref
from typing import List, Sequence, Tuple

import torch

from . import _lib

__all__ = [
    "ms_deform_attn_core", "sample_indices", "level_start_index", "locations_and_weights",
    "pack_value", "clear_repack_cache", "set_default_coord_mode", "get_default_coord_mode", "ValueList",
    "ms_deform_attn_fused",
]

_DTYPE_CODE = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}

# Rounding chain for the pixel coordinate (see include/msda_b200.h).  On a GPU the reference runs
# ATen's CUDA grid sampler, which nvcc compiles with "(g+1)*size - 1" contracted into one FMA
# (tools/probe_coord_mode.py: 11 of 11 ambiguous samples in 3.2e8 agree with the FMA chain, see
# profiles/r01_coord_chain_probe.json), so FMA is the default.  COORD_UNFUSED reproduces the
# one-rounding-per-op chain of elementwise torch ops / the CPU oracle bit for bit.
_default_coord_mode = _lib.COORD_FMA


def set_default_coord_mode(mode: int) -> None:
    global _default_coord_mode
    if mode not in (_lib.COORD_UNFUSED, _lib.COORD_FMA):
        raise ValueError(f"unknown coord mode {mode}")
    _default_coord_mode = mode


def get_default_coord_mode() -> int:
    return _default_coord_mode


def level_start_index(spatial_shapes: Sequence[Sequence[int]]) -> List[int]:
    """Offsets of each level in the flattened pyramid (exclusive running sum of H_l*W_l)."""
    starts, acc = [], 0
    for h, w in spatial_shapes:
        starts.append(acc)
        acc += int(h) * int(w)
    return starts


def _shapes_tuple(spatial_shapes) -> Tuple[Tuple[int, int], ...]:
    if isinstance(spatial_shapes, torch.Tensor):
        spatial_shapes = spatial_shapes.tolist()
    return tuple((int(h), int(w)) for h, w in spatial_shapes)


def _code(dtype: torch.dtype) -> int:
    try:
        return _DTYPE_CODE[dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {dtype}: the kernels take float32 or bfloat16") from None


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: detrpose_b200 has no CPU path")


# --------------------------------------------------------------------------
# value layout: list of strided per-level views  ->  channel-last pyramid
# --------------------------------------------------------------------------
class ValueList:
    """Lazy stand-in for the reference's value list (transformer.py:1285-1286) that remembers ``memory``.

    ``patch.install_value_producer`` makes ``Transformer.forward`` hand this to the decoder instead of the
    permuted copy: the kernels read ``memory (N, S, C)`` directly (row f2: no repack, no un-repack), while
    any other consumer that indexes / iterates it gets the reference's own per-level tensors, built on
    first use by the reference's own expression.
    """

    def __init__(self, memory: torch.Tensor, n_heads: int, split_sizes, dim=-1):
        self.memory = memory
        self.n_heads = int(n_heads)
        self.split_sizes = [int(s) for s in split_sizes]
        self._levels = None

    def levels(self):
        if self._levels is None:
            v = self.memory.unflatten(2, (self.n_heads, -1)).permute(0, 2, 3, 1).flatten(0, 1)
            self._levels = v.split(self.split_sizes, dim=-1)
        return self._levels

    def __len__(self):
        return len(self.split_sizes)

    def __getitem__(self, i):
        return self.levels()[i]

    def __iter__(self):
        return iter(self.levels())


class _ValueHub:
    """State shared by every core call on one value (list) inside one autograd graph: the channel-last
    pyramid (repacked once) and ONE fp32 gradient buffer that the backward launches of all decoder layers
    accumulate into (C ABI ``accumulate=1``), handed to autograd -- and, for the reference's strided list,
    un-repacked -- exactly once, by the token node's backward."""

    __slots__ = ("pyramid", "shapes", "n_heads", "is_list", "value_meta", "buffer", "token", "tok_grad",
                 "spent", "__weakref__")

    def __init__(self, pyramid, shapes, n_heads, is_list, value_meta):
        self.pyramid, self.shapes, self.n_heads = pyramid, shapes, n_heads
        self.is_list, self.value_meta = is_list, value_meta
        self.buffer = None
        self.token = None
        self.tok_grad = None
        self.spent = False


class _HubToken(torch.autograd.Function):
    """1-element tensor that ties the core calls on one value to the value's autograd node."""

    @staticmethod
    def forward(ctx, hub, *value):
        # weak: hub -> token -> this node -> hub would be a reference cycle, and the pyramid / gradient buffer
        # of every pass would wait for the cycle collector instead of being freed with the graph.  The hub is
        # kept alive by the core nodes of the same graph, which run before this node does.
        ctx.hub_ref = weakref.ref(hub)
        ctx.n = len(value)
        return torch.zeros(1, dtype=torch.float32, device=value[0].device)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, _):
        hub = ctx.hub_ref()
        if hub is None:
            return (None,) * (1 + ctx.n)
        gv, hub.buffer = hub.buffer, None
        hub.spent = True                       # a later forward on the same value starts a new hub
        if gv is None:
            return (None,) * (1 + ctx.n)
        if hub.is_list:
            dtype = hub.value_meta[0][1]
            grads = _unpack_grad(gv, hub.shapes, hub.n_heads, dtype if dtype in _DTYPE_CODE else torch.float32)
            grads = [g.to(m[1]) for g, m in zip(grads, hub.value_meta)]
        else:
            shape, dtype = hub.value_meta[0]
            grads = [gv.reshape(shape).to(dtype)]
        stats["grad_handover"] += 1
        return (None, *grads)


class _HubCache:
    """One-entry cache so that the decoder layers sharing one value (list) share one hub.

    Keyed on the identity of the value object(s) (held weakly; the entry dies with them), their version
    counters, the autograd mode and the CUDA stream, so a new forward pass -- new view objects, even at
    the same address -- an in-place update, or a consumer on another stream never hits a stale entry."""

    def __init__(self):
        self._ref = None
        self._sig = None
        self._hub = None

    def get(self, tensors, grad):
        if self._ref is None or self._ref() is not tensors[0]:
            return None
        if self._sig != self._signature(tensors, grad) or self._hub.spent:
            return None
        return self._hub

    def put(self, tensors, grad, hub):
        me = weakref.ref(self)

        def _gone(_):
            c = me()
            if c is not None and c._ref is not None and c._ref() is None:
                c.clear()
        self._ref = weakref.ref(tensors[0], _gone)
        self._sig = self._signature(tensors, grad)
        self._hub = hub

    def clear(self):
        self._ref = self._sig = self._hub = None

    @staticmethod
    def _signature(tensors, grad):
        stream = torch.cuda.current_stream(tensors[0].device).cuda_stream if tensors[0].is_cuda else 0
        return (grad, stream) + tuple((id(v), v._version, v.data_ptr(), tuple(v.shape), tuple(v.stride()))
                                      for v in tensors)


_hub_cache = _HubCache()
# launch counters (tests, bench)
stats = {"repack_launches": 0, "forward_launches": 0, "backward_launches": 0, "grad_handover": 0,
         "unpack_launches": 0, "fused_forward_launches": 0}


def clear_repack_cache() -> None:
    _hub_cache.clear()


def _get_hub(value, shapes, n_heads) -> _ValueHub:
    """Hub of ``value`` (a tensor, a ``ValueList`` or the reference's list of per-level tensors)."""
    if isinstance(value, ValueList):
        value = value.memory
    is_list = not isinstance(value, torch.Tensor)
    tensors = list(value) if is_list else [value]
    grad = torch.is_grad_enabled() and any(t.requires_grad for t in tensors)
    hub = _hub_cache.get(tensors, grad)
    if hub is None:
        pyramid = pack_value(tensors if is_list else value, shapes, n_heads)
        hub = _ValueHub(pyramid, shapes, n_heads, is_list, [(tuple(v.shape), v.dtype) for v in tensors])
        if grad:
            hub.token = _HubToken.apply(hub, *tensors)
            hub.tok_grad = torch.zeros(1, dtype=torch.float32, device=pyramid.device)
        _hub_cache.put(tensors, grad, hub)
    return hub


def _zero_copy_view(levels, shapes, n_heads):
    """Return a (N, S, H, Dh) view if the level list already is one channel-last pyramid."""
    v0 = levels[0]
    nh, dh, _ = v0.shape
    if v0.stride(1) != 1 or nh % n_heads:
        return None
    n = nh // n_heads
    es = v0.element_size()
    s_nh, s_s = v0.stride(0), v0.stride(2)
    base = v0.untyped_storage().data_ptr()
    acc = 0
    for v, (h, w) in zip(levels, shapes):
        if (v.stride(0), v.stride(1), v.stride(2)) != (s_nh, 1, s_s):
            return None
        if v.untyped_storage().data_ptr() != base or v.data_ptr() != v0.data_ptr() + acc * s_s * es:
            return None
        acc += h * w
    if (s_nh * es) % 16 or (s_s * es) % 16 or v0.data_ptr() % 16:
        return None
    return torch.as_strided(v0, (n, acc, n_heads, dh), (n_heads * s_nh, s_s, s_nh, 1), v0.storage_offset())


def pack_value(value, spatial_shapes, n_heads: int) -> torch.Tensor:
    """Bring ``value`` into the kernel-native layout ``(N, S, H, Dh)`` (channel stride 1).

    Accepts the reference's list of ``(N*H, Dh, H_l*W_l)`` views (zero-copy when
    they already are channel-innermost views of one buffer -- the N == 1 case of
    transformer.py:1285-1286 -- otherwise one repack kernel), or a single tensor
    ``(N, S, C)`` / ``(N, S, H, Dh)`` such as the reference's ``memory``.
    """
    shapes = _shapes_tuple(spatial_shapes)
    total = sum(h * w for h, w in shapes)
    if isinstance(value, ValueList):
        value = value.memory
    if isinstance(value, torch.Tensor):
        _require_cuda(value, "value")
        if value.dtype == torch.float16:         # the reference's AMP is fp16 autocast: its sampler runs fp32
            value = value.float()
        if value.dim() == 3:
            value = value.unflatten(2, (n_heads, -1))
        if value.dim() != 4 or value.shape[1] != total or value.shape[2] != n_heads:
            raise ValueError(f"value tensor must be (N, S={total}, C) or (N, S, H={n_heads}, Dh), got {tuple(value.shape)}")
        if value.stride(3) != 1 or any((s * value.element_size()) % 16 for s in value.stride()[:3]) \
                or value.data_ptr() % 16:
            value = value.contiguous()
        return value
    levels = list(value)
    if len(levels) != len(shapes):
        raise ValueError(f"{len(levels)} value levels but {len(shapes)} spatial shapes")
    for v, (h, w) in zip(levels, shapes):
        _require_cuda(v, "value")
        if v.dim() != 3 or v.shape[2] != h * w or v.shape[:2] != levels[0].shape[:2]:
            raise ValueError(f"value level of shape {tuple(v.shape)} does not match (N*H, Dh, {h}*{w})")
    view = _zero_copy_view(levels, shapes, n_heads) if levels[0].dtype in _DTYPE_CODE else None
    if view is not None:
        return view
    nh, dh, _ = levels[0].shape
    if nh % n_heads:
        raise ValueError(f"value leading dim {nh} is not a multiple of n_heads={n_heads}")
    n = nh // n_heads
    dtype = levels[0].dtype
    if dtype == torch.float16:                   # the reference's AMP (fp16 autocast): its sampler runs fp32
        levels = [v.float() for v in levels]
        dtype = torch.float32
    _code(dtype)                                 # anything else but fp32 / bf16: TypeError
    pyramid = torch.empty((n, total, n_heads, dh), dtype=dtype, device=levels[0].device)
    lib = _lib.load()
    with torch.cuda.device(pyramid.device):
        rc = lib.msda_b200_repack(
            _lib.ptr_array([v.data_ptr() for v in levels]),
            _lib.i64_array([s for v in levels for s in v.stride()]),
            _code(dtype), _lib.i32_array([d for hw in shapes for d in hw]),
            pyramid.data_ptr(), _code(dtype), n, n_heads, dh, len(shapes), _stream_ptr(pyramid.device))
    _lib.check(rc, "msda_b200_repack")
    stats["repack_launches"] += 1
    return pyramid


# --------------------------------------------------------------------------
# raw launches (no autograd)
# --------------------------------------------------------------------------
def _forward_raw(pyramid, shapes, loc, attn, out_dtype, coord_mode):
    n, _, n_heads, dh = pyramid.shape
    _, lq, _, n_levels, n_points, _ = loc.shape
    out = torch.empty((n, lq, n_heads * dh), dtype=out_dtype, device=pyramid.device)
    lib = _lib.load()
    with torch.cuda.device(pyramid.device):
        rc = lib.msda_b200_forward(
            pyramid.data_ptr(), _code(pyramid.dtype), _lib.i64_array(pyramid.stride()[:3]),
            _lib.i32_array([d for hw in shapes for d in hw]),
            loc.data_ptr(), attn.data_ptr(), out.data_ptr(), _code(out_dtype),
            n, lq, n_heads, dh, n_levels, n_points, coord_mode, _stream_ptr(pyramid.device))
    _lib.check(rc, "msda_b200_forward")
    stats["forward_launches"] += 1
    return out


def _backward_raw(pyramid, shapes, loc, attn, grad_out, need_value, need_small, coord_mode, into=None,
                  accumulate=None):
    """``into``: optional fp32 (N, S, H, Dh) buffer for the value gradient; the launch ADDS to it
    (``accumulate``, default when ``into`` is given: layers sharing a value) or overwrites it."""
    if accumulate is None:
        accumulate = into is not None
    n, total, n_heads, dh = pyramid.shape
    _, lq, _, n_levels, n_points, _ = loc.shape
    dev = pyramid.device
    grad_value = None
    if need_value:
        grad_value = into if into is not None else torch.empty((n, total, n_heads, dh), dtype=torch.float32,
                                                               device=dev)
    grad_loc = torch.empty_like(loc) if need_small else None
    grad_attn = torch.empty_like(attn) if need_small else None
    lib = _lib.load()
    with torch.cuda.device(dev):
        rc = lib.msda_b200_backward(
            pyramid.data_ptr(), _code(pyramid.dtype), _lib.i64_array(pyramid.stride()[:3]),
            _lib.i32_array([d for hw in shapes for d in hw]),
            loc.data_ptr(), attn.data_ptr(), grad_out.data_ptr(), _code(grad_out.dtype),
            grad_value.data_ptr() if need_value else None, 1 if (need_value and accumulate) else 0,
            grad_loc.data_ptr() if need_small else None,
            grad_attn.data_ptr() if need_small else None,
            n, lq, n_heads, dh, n_levels, n_points, coord_mode, _stream_ptr(dev))
    _lib.check(rc, "msda_b200_backward")
    stats["backward_launches"] += 1
    return grad_value, grad_loc, grad_attn


def _unpack_grad(grad_value, shapes, n_heads, dtype):
    """fp32 channel-last grad pyramid -> list of (N*H, Dh, H_l*W_l) grads (views of one buffer)."""
    n, total, _, dh = grad_value.shape
    buf = torch.empty((n * n_heads, dh, total), dtype=dtype, device=grad_value.device)
    levels = list(buf.split([h * w for h, w in shapes], dim=-1))
    lib = _lib.load()
    with torch.cuda.device(buf.device):
        rc = lib.msda_b200_unpack_grad(
            grad_value.data_ptr(), _lib.i32_array([d for hw in shapes for d in hw]),
            _lib.ptr_array([v.data_ptr() for v in levels]),
            _lib.i64_array([s for v in levels for s in v.stride()]),
            _code(dtype), n, n_heads, dh, len(shapes), _stream_ptr(buf.device))
    _lib.check(rc, "msda_b200_unpack_grad")
    stats["unpack_launches"] += 1
    return levels


def _as_f32_contig(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _MSDACore(torch.autograd.Function):
    """out = core(loc, attn, token); the value pyramid and its shared gradient buffer live in the hub."""

    @staticmethod
    def forward(ctx, loc, attn, meta, hub, token):
        shapes, n_heads, coord_mode, out_dtype = meta
        pyramid = hub.pyramid
        loc_c, attn_c = _as_f32_contig(loc), _as_f32_contig(attn)
        out = _forward_raw(pyramid, shapes, loc_c, attn_c, out_dtype or pyramid.dtype, coord_mode)
        ctx.meta = meta
        ctx.hub = hub
        ctx.small_dtypes = (loc.dtype, attn.dtype)
        # the pyramid goes through save_for_backward as well: an in-place change of value / memory between
        # forward and backward then trips autograd's version check instead of giving silently wrong gradients
        ctx.save_for_backward(loc_c, attn_c, pyramid)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        shapes, n_heads, coord_mode, _ = ctx.meta
        loc_c, attn_c, pyramid = ctx.saved_tensors
        hub = ctx.hub
        need_small = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        need_value = ctx.needs_input_grad[4]
        if grad_out.dtype not in _DTYPE_CODE:
            grad_out = grad_out.float()
        grad_out = grad_out.contiguous()
        into, first = hub.buffer, hub.buffer is None
        if need_value and first:
            n, total, _, dh = pyramid.shape
            into = torch.empty((n, total, n_heads, dh), dtype=torch.float32, device=pyramid.device)
        _, gl, ga = _backward_raw(pyramid, shapes, loc_c, attn_c, grad_out, need_value, need_small,
                                  coord_mode, into=into if need_value else None, accumulate=not first)
        tok = None
        if need_value:
            hub.buffer = into
            tok = hub.tok_grad                  # a non-None gradient, so that the token node runs
        if gl is not None:
            gl, ga = gl.to(ctx.small_dtypes[0]), ga.to(ctx.small_dtypes[1])
        return gl, ga, None, None, tok


def ms_deform_attn_core(value, value_spatial_shapes, sampling_locations, attention_weights,
                        *, n_heads: int = None, out_dtype: torch.dtype = None,
                        coord_mode: int = None) -> torch.Tensor:
    """Multi-scale deformable attention sampling core on B200 (forward + autograd).

    Same argument meaning and result as the reference's
    ``ms_deform_attn_core_pytorch`` (ms_deform_attn.py:145-193) with its optional
    flags off.  ``value`` may also be one ``(N, S, C)`` / ``(N, S, H, Dh)`` tensor
    (the reference's ``memory``), which is consumed zero-copy.
    """
    _require_cuda(sampling_locations, "sampling_locations")
    if sampling_locations.dim() != 6 or sampling_locations.shape[-1] != 2:
        raise ValueError("sampling_locations must be (N, Lq, H, L, P, 2)")
    n, lq, heads, n_levels, n_points, _ = sampling_locations.shape
    if tuple(attention_weights.shape) != (n, lq, heads, n_levels, n_points):
        raise ValueError("attention_weights must be (N, Lq, H, L, P) matching sampling_locations")
    if n_heads is not None and n_heads != heads:
        raise ValueError(f"n_heads={n_heads} but sampling_locations has {heads} heads")
    shapes = _shapes_tuple(value_spatial_shapes)
    if len(shapes) != n_levels:
        raise ValueError(f"{len(shapes)} spatial shapes for {n_levels} levels")
    if n_levels > _lib.MAX_LEVELS or n_points > _lib.MAX_POINTS:
        raise ValueError(f"at most {_lib.MAX_LEVELS} levels and {_lib.MAX_POINTS} points are supported")
    meta = (shapes, heads, _default_coord_mode if coord_mode is None else coord_mode, out_dtype)
    hub = _get_hub(value, shapes, heads)
    return _MSDACore.apply(sampling_locations, attention_weights, meta, hub, hub.token)


def sample_indices(sampling_locations: torch.Tensor, spatial_shapes, coord_mode: int = None):
    """Integer corner indices ``(y0, x0)`` per sample and the level offsets, from the kernels.

    Returns ``(idx int32 (N, Lq, H, L, P, 2), level_start int32 (L,))``.
    """
    _require_cuda(sampling_locations, "sampling_locations")
    shapes = _shapes_tuple(spatial_shapes)
    loc = _as_f32_contig(sampling_locations)
    n, lq, heads, n_levels, n_points, _ = loc.shape
    idx = torch.empty(loc.shape, dtype=torch.int32, device=loc.device)
    starts = torch.empty((n_levels,), dtype=torch.int32, device=loc.device)
    lib = _lib.load()
    with torch.cuda.device(loc.device):
        rc = lib.msda_b200_sample_indices(
            _lib.i32_array([d for hw in shapes for d in hw]), loc.data_ptr(), idx.data_ptr(), starts.data_ptr(),
            n, lq, heads, n_levels, n_points,
            _default_coord_mode if coord_mode is None else coord_mode, _stream_ptr(loc.device))
    _lib.check(rc, "msda_b200_sample_indices")
    return idx, starts


_normalizer_cache: dict = {}


def level_normalizer(shapes, device) -> torch.Tensor:
    """fp32 ``(L, 2)`` tensor of ``(W_l, H_l)`` on ``device``, created once per (shapes, device): building it
    from a Python list on every call (ms_deform_attn.py:414) is a blocking host-to-device copy."""
    key = (tuple(tuple(int(d) for d in hw) for hw in shapes), str(device))
    t = _normalizer_cache.get(key)
    if t is None:
        t = torch.tensor([[w, h] for h, w in key[0]], dtype=torch.float32, device=device)
        _normalizer_cache[key] = t
    return t


class _Prologue(torch.autograd.Function):
    """Fused softmax + location kernel with an analytic backward (elementwise torch ops)."""

    @staticmethod
    def forward(ctx, offsets, logits, ref_points, meta):
        shapes, n_heads, n_levels, n_points = meta
        loc, att = _locations_raw(offsets, logits, ref_points, shapes, n_heads, n_levels, n_points)
        ctx.meta = meta
        ctx.shapes_in = (offsets.shape, logits.shape, ref_points.shape)
        ctx.save_for_backward(att)
        return loc, att

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loc, grad_att):
        shapes, n_heads, n_levels, n_points = ctx.meta
        (att,) = ctx.saved_tensors
        off_shape, lg_shape, ref_shape = ctx.shapes_in
        n, lq = att.shape[:2]
        g_off = g_lg = g_ref = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[2]:
            grad_loc = grad_loc.float()
            if ctx.needs_input_grad[0]:
                norm = level_normalizer(shapes, att.device)
                g_off = (grad_loc / norm.view(1, 1, 1, n_levels, 1, 2)).reshape(off_shape)
            if ctx.needs_input_grad[2]:
                g_ref = grad_loc.sum(dim=(2, 4))                       # over heads and points -> (N, Lq, L, 2)
                if ref_shape[2] == 1:
                    g_ref = g_ref.sum(dim=2, keepdim=True)
                g_ref = g_ref.reshape(ref_shape)
        if ctx.needs_input_grad[1]:
            a = att.view(n, lq, n_heads, n_levels * n_points)
            g = grad_att.float().reshape(n, lq, n_heads, n_levels * n_points)
            g_lg = (a * (g - (a * g).sum(-1, keepdim=True))).reshape(lg_shape)
        return g_off, g_lg, g_ref, None


def locations_and_weights(offsets: torch.Tensor, logits: torch.Tensor, ref_points: torch.Tensor,
                          spatial_shapes, n_heads: int, n_levels: int, n_points: int):
    """Fused prologue, differentiable: softmax over L*P and ``ref + offsets / (W_l, H_l)``.

    ms_deform_attn.py:392-393 and :412-416.  ``offsets`` ``(N, Lq, H*L*P*2)``, ``logits``
    ``(N, Lq, H*L*P)``, ``ref_points`` ``(N, Lq, 1|L, 2)``, all fp32.  Returns fp32
    ``(locations (N, Lq, H, L, P, 2), attention (N, Lq, H, L, P))``.
    """
    shapes = _shapes_tuple(spatial_shapes)
    return _Prologue.apply(offsets, logits, ref_points, (shapes, n_heads, n_levels, n_points))


def _locations_raw(offsets: torch.Tensor, logits: torch.Tensor, ref_points: torch.Tensor,
                   spatial_shapes, n_heads: int, n_levels: int, n_points: int):
    """Fused prologue launch (no autograd): softmax over L*P and ``ref + offsets / (W_l, H_l)``.

    ms_deform_attn.py:392-393 and :412-416.  ``offsets`` ``(N, Lq, H*L*P*2)``, ``logits``
    ``(N, Lq, H*L*P)``, ``ref_points`` ``(N, Lq, 1|L, 2)``.  Returns fp32
    ``(locations (N, Lq, H, L, P, 2), attention (N, Lq, H, L, P))``.
    """
    _require_cuda(offsets, "offsets")
    shapes = _shapes_tuple(spatial_shapes)
    n, lq = offsets.shape[:2]
    off, lg, ref = _as_f32_contig(offsets), _as_f32_contig(logits), _as_f32_contig(ref_points)
    ref_levels = ref.shape[2]
    loc = torch.empty((n, lq, n_heads, n_levels, n_points, 2), dtype=torch.float32, device=off.device)
    att = torch.empty((n, lq, n_heads, n_levels, n_points), dtype=torch.float32, device=off.device)
    lib = _lib.load()
    with torch.cuda.device(off.device):
        rc = lib.msda_b200_locations(off.data_ptr(), lg.data_ptr(), ref.data_ptr(), ref_levels,
                                     _lib.i32_array([d for hw in shapes for d in hw]),
                                     loc.data_ptr(), att.data_ptr(), n, lq, n_heads, n_levels, n_points,
                                     _stream_ptr(off.device))
    _lib.check(rc, "msda_b200_locations")
    return loc, att


# --------------------------------------------------------------------------
# row f1: sampler with the prologue fused in (ms_deform_attn.py:392-393, :412-416 + core :145-193)
# --------------------------------------------------------------------------
_fused_ok_cache: dict = {}


def _fused_supported(pyramid, shapes, lq, n_levels, n_points) -> bool:
    n, _, n_heads, dh = pyramid.shape
    key = (pyramid.dtype, tuple(pyramid.stride()[:3]), shapes, n, lq, n_heads, dh, n_levels, n_points)
    ok = _fused_ok_cache.get(key)
    if ok is None:
        lib = _lib.load()
        ok = bool(lib.msda_b200_fused_supported(_code(pyramid.dtype), _lib.i64_array(pyramid.stride()[:3]),
                                                _lib.i32_array([d for hw in shapes for d in hw]),
                                                n, lq, n_heads, dh, n_levels, n_points))
        _fused_ok_cache[key] = ok
    return ok


class _MSDAFused(torch.autograd.Function):
    """out = sampler(offsets, logits, ref, token): softmax, locations and sampling in one launch per direction."""

    @staticmethod
    def forward(ctx, offsets, logits, ref, meta, hub, token):
        shapes, n_heads, n_levels, n_points, coord_mode, out_dtype = meta
        pyramid = hub.pyramid
        off, lg, rf = _as_f32_contig(offsets), _as_f32_contig(logits), _as_f32_contig(ref)
        n, lq = off.shape[:2]
        dh = pyramid.shape[3]
        dev = pyramid.device
        need_bwd = any(ctx.needs_input_grad[:3]) or ctx.needs_input_grad[5]
        attn = torch.empty((n, lq, n_heads, n_levels, n_points), dtype=torch.float32, device=dev) if need_bwd else None
        odt = out_dtype or pyramid.dtype
        out = torch.empty((n, lq, n_heads * dh), dtype=odt, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            rc = lib.msda_b200_forward_fused(
                pyramid.data_ptr(), _code(pyramid.dtype), _lib.i64_array(pyramid.stride()[:3]),
                _lib.i32_array([d for hw in shapes for d in hw]),
                off.data_ptr(), lg.data_ptr(), rf.data_ptr(), rf.shape[2],
                out.data_ptr(), _code(odt), attn.data_ptr() if need_bwd else None,
                n, lq, n_heads, dh, n_levels, n_points, coord_mode, _stream_ptr(dev))
        _lib.check(rc, "msda_b200_forward_fused")
        stats["forward_launches"] += 1
        stats["fused_forward_launches"] += 1
        ctx.meta, ctx.hub = meta, hub
        ctx.in_meta = (offsets.shape, offsets.dtype, logits.shape, logits.dtype, ref.shape, ref.dtype)
        if need_bwd:
            ctx.save_for_backward(off, rf, attn, pyramid)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        shapes, n_heads, n_levels, n_points, coord_mode, _ = ctx.meta
        off, rf, attn, pyramid = ctx.saved_tensors
        hub = ctx.hub
        off_shape, off_dtype, lg_shape, lg_dtype, ref_shape, ref_dtype = ctx.in_meta
        need_small = any(ctx.needs_input_grad[:3])
        need_value = ctx.needs_input_grad[5]
        n, total, _, dh = pyramid.shape
        lq = off.shape[1]
        dev = pyramid.device
        if grad_out.dtype != pyramid.dtype:
            grad_out = grad_out.to(pyramid.dtype)
        grad_out = grad_out.contiguous()
        into, first = hub.buffer, hub.buffer is None
        if need_value and first:
            into = torch.empty((n, total, n_heads, dh), dtype=torch.float32, device=dev)
        g_off = torch.empty((n, lq, n_heads, n_levels, n_points, 2), dtype=torch.float32, device=dev) if need_small else None
        g_att = torch.empty_like(attn) if need_small else None
        lib = _lib.load()
        with torch.cuda.device(dev):
            rc = lib.msda_b200_backward_fused(
                pyramid.data_ptr(), _code(pyramid.dtype), _lib.i64_array(pyramid.stride()[:3]),
                _lib.i32_array([d for hw in shapes for d in hw]),
                off.data_ptr(), rf.data_ptr(), rf.shape[2], attn.data_ptr(),
                grad_out.data_ptr(), _code(grad_out.dtype),
                into.data_ptr() if need_value else None, 0 if first else 1,
                g_off.data_ptr() if need_small else None, g_att.data_ptr() if need_small else None,
                n, lq, n_heads, dh, n_levels, n_points, coord_mode, _stream_ptr(dev))
            _lib.check(rc, "msda_b200_backward_fused")
            stats["backward_launches"] += 1
            g_lg = g_ref = None
            if need_small and ctx.needs_input_grad[1]:
                rc = lib.msda_b200_softmax_backward(attn.data_ptr(), g_att.data_ptr(), g_att.data_ptr(),
                                                    n * lq * n_heads, n_levels * n_points, _stream_ptr(dev))
                _lib.check(rc, "msda_b200_softmax_backward")
                g_lg = g_att.reshape(lg_shape).to(lg_dtype)
        if need_small and ctx.needs_input_grad[2]:
            # location = ref + offset / (W_l, H_l): d/d ref = sum over heads and points (and levels when the
            # reference point is shared) of grad_offset * (W_l, H_l)
            norm = level_normalizer(shapes, dev).view(1, 1, 1, n_levels, 1, 2)
            g_ref = (g_off * norm).sum(dim=(2, 4))
            if ref_shape[2] == 1:
                g_ref = g_ref.sum(dim=2, keepdim=True)
            g_ref = g_ref.reshape(ref_shape).to(ref_dtype)
        tok = None
        if need_value:
            hub.buffer = into
            tok = hub.tok_grad
        g_off_out = g_off.reshape(off_shape).to(off_dtype) if (need_small and ctx.needs_input_grad[0]) else None
        return g_off_out, g_lg, g_ref, None, None, tok


def ms_deform_attn_fused(value, value_spatial_shapes, offsets, logits, ref_points, *, n_heads: int,
                         n_levels: int, n_points: int, out_dtype: torch.dtype = None,
                         coord_mode: int = None) -> torch.Tensor:
    """``MSDeformAttn.forward`` after its two Linears, 2-D reference points (ms_deform_attn.py:392-393,
    :412-416, :440 -> :145-193): softmax over L*P, ``ref + offsets / (W_l, H_l)`` and the sampling core in
    ONE launch per direction; locations and weights never reach HBM.

    ``offsets`` ``(N, Lq, H*L*P*2)``, ``logits`` ``(N, Lq, H*L*P)`` (the Linear outputs, any float dtype:
    the arithmetic is fp32), ``ref_points`` ``(N, Lq, 1|L, 2)``; ``value`` as for ``ms_deform_attn_core``.
    Differentiable w.r.t. offsets, logits, reference points and value.  Shapes outside the fused kernels run
    the two-step path (``locations_and_weights`` + ``ms_deform_attn_core``) with the same result."""
    _require_cuda(offsets, "offsets")
    shapes = _shapes_tuple(value_spatial_shapes)
    if len(shapes) != n_levels:
        raise ValueError(f"{len(shapes)} spatial shapes for {n_levels} levels")
    if n_levels > _lib.MAX_LEVELS or n_points > _lib.MAX_POINTS:
        raise ValueError(f"at most {_lib.MAX_LEVELS} levels and {_lib.MAX_POINTS} points are supported")
    n, lq = offsets.shape[:2]
    if offsets.numel() != n * lq * n_heads * n_levels * n_points * 2 or logits.numel() * 2 != offsets.numel():
        raise ValueError("offsets must be (N, Lq, H*L*P*2) and logits (N, Lq, H*L*P)")
    if ref_points.dim() != 4 or ref_points.shape[-1] != 2 or ref_points.shape[2] not in (1, n_levels):
        raise ValueError("ref_points must be (N, Lq, 1|L, 2)")
    hub = _get_hub(value, shapes, n_heads)
    cm = _default_coord_mode if coord_mode is None else coord_mode
    if not _fused_supported(hub.pyramid, shapes, lq, n_levels, n_points):
        loc, att = locations_and_weights(offsets.float(), logits.float(), ref_points.float(), shapes, n_heads,
                                         n_levels, n_points)
        return _MSDACore.apply(loc, att, (shapes, n_heads, cm, out_dtype), hub, hub.token)
    meta = (shapes, n_heads, n_levels, n_points, cm, out_dtype)
    return _MSDAFused.apply(offsets, logits, ref_points, meta, hub, hub.token)
