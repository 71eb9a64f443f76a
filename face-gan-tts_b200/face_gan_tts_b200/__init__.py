"""face_gan_tts_b200 -- B200-native (sm_100a) log-prior + Monotonic Alignment Search.

Drop-in for the alignment hot path of CognitiveModeling/Face-GAN-TTS:
`model/monotonic_align` (Cython, CPU) and the log-prior block of
`FaceTTS.compute_loss` (model/face_tts.py:165-174).  All compute runs in
libmas_b200.so (hand-written CUDA behind a C ABI, include/mas_b200.h); this
package is the thin PyTorch-facing host layer.  No CPU fallback.
"""
from . import losses, monotonic_align, sharding  # noqa: F401
from .alignment import (  # noqa: F401
    AlignmentPlan,
    AlignmentResult,
    align,
    durations_to_logw,
    expand_durations,
    generate_path,
    log_prior,
    log_prior_maximum_path,
    pack_batch,
    upload_batch,
    upload_packed_batch,
    unpack_batch,
)
from .install import install, uninstall  # noqa: F401
from .losses import (  # noqa: F401
    AlignmentLosses,
    alignment_losses,
    crop_frames,
    duration_loss,
    gather_mu_y,
    prior_loss,
    sequence_mask,
)

__all__ = [
    "monotonic_align", "AlignmentPlan", "AlignmentResult", "align", "log_prior", "log_prior_maximum_path", "generate_path",
    "durations_to_logw", "expand_durations", "upload_batch", "pack_batch", "upload_packed_batch", "unpack_batch", "install", "uninstall", "losses", "AlignmentLosses", "alignment_losses", "crop_frames",
    "duration_loss", "gather_mu_y", "prior_loss", "sequence_mask",
]
