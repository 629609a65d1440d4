from .compute_metrics import eval_metrics, eval_fscores, upsample
from .generate_summary import generate_summary
from .knapsack_implementation import knapSack
from .evaluation_metrics import evaluate_summary
