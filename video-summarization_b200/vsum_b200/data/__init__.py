"""Input side of the hot path without h5py: a packed on-disk format, a
memory-mapped native reader, and a prefetching loader whose batches are already in the packed
(padding-free) layout the scorer consumes.  Mirrors the records of `src/data/dataset.py`."""
from .packed import PackedDataset, PackedLoader, UserSummaries, convert_h5, write_pack

__all__ = ["PackedDataset", "PackedLoader", "UserSummaries", "convert_h5", "write_pack"]
