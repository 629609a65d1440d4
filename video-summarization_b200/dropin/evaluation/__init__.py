"""Drop-in for the reference's `src/evaluation` package (train.py:16,
generate_summary_image.py:18): same module and function names, B200 kernels underneath."""
from .compute_metrics import eval_metrics  # noqa: F401  (mirrors src/evaluation/__init__.py:2)
