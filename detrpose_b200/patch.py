"""Drop the sm_100a sampling core into an unmodified DETRPose checkout.

The reference model resolves the core by module-global name
(``ms_deform_attn_core_pytorch`` inside
``src/models/detrpose/ms_deform_attn.py``, called at :440 and :496), so
replacing that one global accelerates every ``MSDeformAttn`` instance while
keeping the reference class -- its ``isinstance`` re-init hook
(transformer.py:1154-1156), parameter names, ``state_dict`` and EMA copies -- exactly
as it is.  Configurations outside the kernel contract (sampling modulation,
region pooling, energy sampler) are routed, explicitly, to the reference's own
function.  CPU tensors raise: this package has no CPU path (pass
``cpu_to_reference=True`` to hand them back to the reference's code instead,
e.g. for ONNX export of a patched checkout).
"""
from __future__ import annotations

import functools

from . import functional as MF

__all__ = ["install", "uninstall", "make_core"]

_ORIGINAL_ATTR = "_detrpose_b200_original_core"


def make_core(reference_core, cpu_to_reference: bool = False):
    """Wrap the reference core: baseline configuration -> kernels, optional branches -> reference."""

    @functools.wraps(reference_core)
    def core(value, value_spatial_shapes, sampling_locations, attention_weights,
             sampling_modulation=None, region_kernel_size=1, is_energy=False):
        baseline = (sampling_modulation is None and (region_kernel_size is None or region_kernel_size <= 1)
                    and not is_energy)
        if not baseline or (cpu_to_reference and not sampling_locations.is_cuda):
            return reference_core(value, value_spatial_shapes, sampling_locations, attention_weights,
                                  sampling_modulation=sampling_modulation,
                                  region_kernel_size=region_kernel_size, is_energy=is_energy)
        return MF.ms_deform_attn_core(value, value_spatial_shapes, sampling_locations, attention_weights)

    return core


def install(reference_module, cpu_to_reference: bool = False) -> None:
    """``reference_module``: the imported ``src.models.detrpose.ms_deform_attn`` module object."""
    if hasattr(reference_module, _ORIGINAL_ATTR):
        return
    original = reference_module.ms_deform_attn_core_pytorch
    setattr(reference_module, _ORIGINAL_ATTR, original)
    reference_module.ms_deform_attn_core_pytorch = make_core(original, cpu_to_reference)


def uninstall(reference_module) -> None:
    original = getattr(reference_module, _ORIGINAL_ATTR, None)
    if original is not None:
        reference_module.ms_deform_attn_core_pytorch = original
        delattr(reference_module, _ORIGINAL_ATTR)
