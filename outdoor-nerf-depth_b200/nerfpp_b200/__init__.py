"""B200-native NeRF++ ray-marching hot path (sampling -> 8x256 MLP -> composite -> depth-prior loss).

The product is ``libnerfpp_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/nerfpp_b200.h``); this package is the thin torch-facing binding.  The drop-in modules
that mirror the reference's import surface (``ddp_model``, ``depth_loss``) live one directory up.
"""
from ._lib import FIELD_SIMT, FIELD_TC, FIELD_TC_SPLIT, NerfppError, lib  # noqa: F401
from . import ops  # noqa: F401
from .render import render_rays, cascade_forward, render_single_image  # noqa: F401
from .graph import GraphedRenderStep, GraphedTrainStep, PipelinedRenderStep  # noqa: F401

__all__ = ["ops", "lib", "NerfppError", "FIELD_TC", "FIELD_SIMT", "FIELD_TC_SPLIT", "render_rays", "cascade_forward", "render_single_image", "GraphedRenderStep", "GraphedTrainStep", "PipelinedRenderStep"]
