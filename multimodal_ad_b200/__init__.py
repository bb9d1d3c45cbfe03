"""multimodal_ad_b200 — B200-native hot path of dongzj56/Multimodal_AD.

Host-side mirrors of the reference's interfaces for the accelerated path,
over the C-ABI CUDA library in csrc/ (include/mmad_b200.h).  No CPU fallback.
"""
from . import _lib  # noqa: F401
from .models.ROI_pol import ROIPool, RoiPlan, roi_pool  # noqa: F401

__all__ = ["ROIPool", "RoiPlan", "roi_pool"]
