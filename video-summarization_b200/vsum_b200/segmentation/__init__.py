"""Kernel temporal segmentation with the reference's call surface
(`src/data/preprocess/segmentations/kts/*`, `create_segments.py:24-52`) on the GPU kernels behind
`vsum_kts_gram` / `vsum_kts_dp`."""
from .kts import cpd_nonlin, kts_seg, kts_segmentation

__all__ = ["cpd_nonlin", "kts_seg", "kts_segmentation"]
