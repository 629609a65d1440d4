/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked or loaded by the product library.
 *
 * Plain-C CPU restatement of the evaluation half of the hot path of
 * BerserkerMother/Video-Summarization (paths relative to the reference root):
 *   - upsample + per-shot float32 mean : src/evaluation/generate_summary.py:25-42
 *     (numpy 2.3.5 float32 pairwise summation)
 *   - 15 % capacity                     : src/evaluation/generate_summary.py:45-46
 *   - 0/1 knapsack DP + back-track      : src/evaluation/knapsack_implementation.py:11-28
 *   - int8 summary mask                 : src/evaluation/generate_summary.py:51-53
 *   - per-user overlap P/R/F, avg|max   : src/evaluation/evaluation_metrics.py:12-33
 *
 * It exists so parity tests can check the CUDA kernels on thousands of videos in seconds;
 * `oracle/ref_port.py` is the same algorithm in the reference's own execution model (pure
 * Python loops) and both are pinned against the imported reference (tests/golden/).
 * Build: `make -C oracle` -> oracle/_build/libvsum_oracle.so.  Compile with -ffp-contract=off:
 * the reference rounds after every float32 add.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* numpy `pairwise_sum` for contiguous float32 (generate_summary.py:42 -> ndarray.mean). */
static float pairwise_sum_f32(const float *a, int64_t n) {
    if (n < 8) {
        float res = -0.0f;
        for (int64_t i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_sum_f32(a, n2) + pairwise_sum_f32(a + n2, n - n2);
}

float vsum_oracle_pairwise_sum_f32(const float *a, int64_t n) { return pairwise_sum_f32(a, n); }

/* generate_summary.py:25-35.  `picks` is already the integer position list; the reference
 * appends n_frames when the last pick differs from it. out has n_frames entries. */
void vsum_oracle_upsample(const float *scores, int64_t n_scores, const int32_t *picks,
                          int64_t n_picks, int64_t n_frames, float *out) {
    memset(out, 0, (size_t)n_frames * sizeof(float));
    int64_t n_pos = n_picks + ((int64_t)picks[n_picks - 1] != n_frames ? 1 : 0);
    for (int64_t i = 0; i + 1 < n_pos; ++i) {
        int64_t lo = picks[i];
        int64_t hi = (i + 1 < n_picks) ? picks[i + 1] : n_frames;
        float v = (i == n_scores) ? 0.0f : scores[i];
        if (lo < 0) lo = 0;
        if (hi > n_frames) hi = n_frames;
        for (int64_t f = lo; f < hi; ++f) out[f] = v;
    }
}

/* generate_summary.py:38-42: shot length and float32 mean widened to fp64. */
void vsum_oracle_shot_means(const float *frame_scores, const int32_t *cps, int64_t n_shots,
                            double *val_out, int32_t *wt_out) {
    for (int64_t s = 0; s < n_shots; ++s) {
        int64_t lo = cps[2 * s], hi = cps[2 * s + 1];
        int64_t n = hi - lo + 1;
        wt_out[s] = (int32_t)n;
        float total = 0.0f + pairwise_sum_f32(frame_scores + lo, n);
        val_out[s] = (double)(total / (float)n);
    }
}

int32_t vsum_oracle_capacity(int32_t last_end) { return (int32_t)((double)(last_end + 1) * 0.15); }

/* knapsack_implementation.py:11-28.  selected_out[i] = 1 iff shot i is chosen. */
void vsum_oracle_knapsack(int32_t cap, const int32_t *wt, const double *val, int64_t n,
                          uint8_t *selected_out) {
    if (n <= 0 || cap < 0) return;
    int64_t width = (int64_t)cap + 1;
    double *row = (double *)calloc((size_t)width, sizeof(double));
    uint8_t *take = (uint8_t *)calloc((size_t)(n * width), 1);
    for (int64_t i = 0; i < n; ++i) {
        int64_t w_i = wt[i];
        double v = val[i];
        uint8_t *t = take + i * width;
        for (int64_t w = cap; w >= 1; --w) {          /* high -> low keeps row[w - w_i] old */
            if (w_i <= w) {
                double a = v + row[w - w_i], b = row[w];
                double m = (b > a) ? b : a;            /* Python max(a, b) */
                t[w] = (m != b);                       /* line 26: K[i][w] != K[i-1][w] */
                row[w] = m;
            }
        }
    }
    memset(selected_out, 0, (size_t)n);
    int64_t w = cap;
    for (int64_t i = n - 1; i >= 0; --i) {
        if (take[i * width + w]) { selected_out[i] = 1; w -= wt[i]; }
    }
    free(row);
    free(take);
}

/* generate_summary.py:51-53. out has last_end+1 entries. */
void vsum_oracle_summary_mask(const int32_t *cps, int64_t n_shots, const uint8_t *selected,
                              int8_t *out) {
    int64_t len = (int64_t)cps[2 * (n_shots - 1) + 1] + 1;
    memset(out, 0, (size_t)len);
    for (int64_t s = 0; s < n_shots; ++s)
        if (selected[s])
            for (int64_t f = cps[2 * s]; f <= cps[2 * s + 1]; ++f) out[f] = 1;
}

/* evaluation_metrics.py:12-33.  method: 0 = avg, 1 = max.  per_user_out may be NULL. */
double vsum_oracle_fscore(const int8_t *summary, int64_t sum_len, const float *user_summary,
                          int64_t n_users, int64_t n_cols, int method, double *per_user_out) {
    int64_t s_cnt = 0;
    for (int64_t f = 0; f < sum_len; ++f) s_cnt += summary[f];
    double acc = 0.0, best = 0.0;
    for (int64_t u = 0; u < n_users; ++u) {
        const float *g = user_summary + u * n_cols;
        int64_t o_cnt = 0, g_cnt = 0;
        for (int64_t f = 0; f < n_cols; ++f) {
            int64_t gi = (int64_t)g[f];                 /* float -> int truncation (line 20) */
            g_cnt += gi;
            if (f < sum_len) o_cnt += ((int64_t)summary[f]) & gi;
        }
        double p = (double)o_cnt / (double)s_cnt;
        double r = (double)o_cnt / (double)g_cnt;
        double fs = (p + r == 0.0) ? 0.0 : ((2.0 * p) * r * 100.0) / (p + r);
        if (per_user_out) per_user_out[u] = fs;
        if (u == 0) { best = fs; acc = 0.0 + fs; }
        else {
            if (fs > best) best = fs;                  /* Python max(list): first maximal, NaN-sticky */
            acc = acc + fs;
        }
    }
    return method == 1 ? best : acc / (double)n_users;
}

/* One video end to end (generate_summary.py:17-55 + evaluation_metrics.py).  Scratch is
 * allocated internally.  Returns the reduced F; optional outputs may be NULL. */
double vsum_oracle_video(const float *scores, int64_t n_scores, const int32_t *picks,
                         int64_t n_picks, int64_t n_frames, const int32_t *cps, int64_t n_shots,
                         const float *user_summary, int64_t n_users, int64_t n_cols, int method,
                         double *val_out, int32_t *wt_out, int32_t *cap_out,
                         uint8_t *selected_out, int8_t *summary_out) {
    float *fs = (float *)malloc((size_t)n_frames * sizeof(float));
    double *val = val_out ? val_out : (double *)malloc((size_t)n_shots * sizeof(double));
    int32_t *wt = wt_out ? wt_out : (int32_t *)malloc((size_t)n_shots * sizeof(int32_t));
    uint8_t *sel = selected_out ? selected_out : (uint8_t *)malloc((size_t)n_shots);
    int64_t sum_len = (int64_t)cps[2 * (n_shots - 1) + 1] + 1;
    int8_t *summ = summary_out ? summary_out : (int8_t *)malloc((size_t)sum_len);
    vsum_oracle_upsample(scores, n_scores, picks, n_picks, n_frames, fs);
    vsum_oracle_shot_means(fs, cps, n_shots, val, wt);
    int32_t cap = vsum_oracle_capacity((int32_t)(sum_len - 1));
    if (cap_out) *cap_out = cap;
    vsum_oracle_knapsack(cap, wt, val, n_shots, sel);
    vsum_oracle_summary_mask(cps, n_shots, sel, summ);
    double f = vsum_oracle_fscore(summ, sum_len, user_summary, n_users, n_cols, method, NULL);
    free(fs);
    if (!val_out) free(val);
    if (!wt_out) free(wt);
    if (!selected_out) free(sel);
    if (!summary_out) free(summ);
    return f;
}
