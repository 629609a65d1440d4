"""Rank correlations against each user's frame scores.  Out of the GPU scope of this round
(SURVEY.md section 8(f) rank 1): like the reference (`src/evaluation/compute_correlation.py:4-15`)
this calls scipy on the host, so `eval_metrics` keeps returning its 3-tuple."""
from scipy import stats


def evaluate_scores(predicted_summary, user_scores):
    taus, rhos = [], []
    pred_rank = stats.rankdata(-predicted_summary)
    for row in user_scores:
        user_rank = stats.rankdata(-row)
        rhos.append(stats.spearmanr(pred_rank, user_rank)[0])
        taus.append(stats.kendalltau(pred_rank, user_rank)[0])
    return sum(taus) / len(taus), sum(rhos) / len(rhos)
