"""Drop-in for the reference's `src/model` package: put `video-summarization_b200/dropin`
(instead of the reference's `src`) on sys.path and `from model import SimNet, PretrainModel`
(train.py:13, pretrain.py:8) resolves to the B200 implementation."""
from vsum_b200.model import SimNet, PretrainModel  # noqa: F401
