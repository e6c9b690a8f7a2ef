"""B200-native CLIP-Event training loss head (similarity + InfoNCE, IPOT graph alignment).

Public surface mirrors the reference's ``model_clip`` / ``model_ot`` names for this path; all
compute runs in hand-written sm_100a CUDA kernels behind the C ABI in include/clip_event_b200.h.
"""
from .model_clip import (ClipEventHead, CriterionAlignment, CriterionContrastive, LazyLogits, LossHeadStep,  # noqa: F401
                         ProjectionTail)
from .model_ot import cost_matrix_cosine, ipot, optimal_transport_dist, trace  # noqa: F401

__all__ = ["ClipEventHead", "CriterionAlignment", "CriterionContrastive", "LazyLogits", "LossHeadStep", "ProjectionTail",
           "cost_matrix_cosine", "ipot", "optimal_transport_dist", "trace"]
