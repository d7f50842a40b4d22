"""GPU counterpart of ``/root/reference/tools/temporal_ir.py`` (pattern
extraction only; the PNG I/O of ``:57-90, 117-122`` stays with the caller).

``extract_temporal_ir_pattern(frames)`` maps a stack of T real IR frames
(emitter ramp; T = 7 in the reference, ``:64-70``) to the binary IR-dot pattern
of ``:93-114`` + ``get_smoothed_ir_pattern`` (``:35-40``)."""
import torch

from .. import ops


def extract_temporal_ir_pattern(frames: torch.Tensor, ks: int = 11, threshold: float = 0.005) -> torch.Tensor:
    """frames: CUDA uint8 [T,H,W] or [B,T,H,W] -> float32 {0,1} pattern [H,W] / [B,H,W]."""
    return ops.temporal_ir_pattern(frames, ks=ks, threshold=threshold)
