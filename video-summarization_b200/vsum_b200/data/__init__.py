"""Input side of the hot path without h5py: a packed on-disk format, a
memory-mapped (or page-locked) native reader, and prefetching loaders whose batches are already in the packed
(padding-free) layout the scorer and the evaluation kernels consume.  Mirrors the records of `src/data/dataset.py`."""
from .packed import PackedDataset, PackedLoader, UserSummaries, convert_h5, write_pack
from .eval_loader import EvalBatch, PackedEvalLoader

__all__ = ["PackedDataset", "PackedLoader", "PackedEvalLoader", "EvalBatch", "UserSummaries", "convert_h5", "write_pack"]
