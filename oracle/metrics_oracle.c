/* Plain-C restatement of the reference's per-class average precision.
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Reference: step_recognition/utils/metrics.py:25-62 (perframe_average_precision, metrics == 'AP'), whose
 * arithmetic is scikit-learn's average_precision_score (requirements.txt:9, not vendored; this image: 1.9.0):
 *   sort by score descending; one (precision, recall) point per DISTINCT score, at the last index of the run;
 *   precision = tp / (tp + fp), recall = tp / positives; AP = sum_n (R_n - R_{n-1}) * P_n, in double.
 * An independent second restatement next to oracle/metrics_np.py: tests/test_oracle_rank4.py checks both against
 * the golden values the reference itself produced (tests/golden/map_cases.npz).
 *
 * Build: see oracle/Makefile (gcc -O2 -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct { float score; int32_t pos; } item_t;

static int cmp_desc(const void* a, const void* b) {
    const float x = ((const item_t*)a)->score, y = ((const item_t*)b)->score;
    return (x < y) - (x > y);  /* ties: any order -- only run ends contribute */
}

/* scores[n], positive[n] (0 / non-zero).  Returns AP, or NaN when there is no positive (the reference skips such
 * classes, metrics.py:54) or on allocation failure. */
double oracle_average_precision(const float* scores, const int32_t* positive, int64_t n) {
    item_t* it = (item_t*)malloc(sizeof(item_t) * (size_t)(n > 0 ? n : 1));
    if (!it) return NAN;
    int64_t total = 0;
    for (int64_t i = 0; i < n; ++i) {
        it[i].score = scores[i];
        it[i].pos = positive[i] != 0;
        total += it[i].pos;
    }
    if (total == 0) { free(it); return NAN; }
    qsort(it, (size_t)n, sizeof(item_t), cmp_desc);
    double ap = 0.0, prev_recall = 0.0;
    int64_t tp = 0;
    for (int64_t i = 0; i < n; ++i) {
        tp += it[i].pos;
        if (i == n - 1 || it[i + 1].score != it[i].score) {  /* last index of a run of equal scores */
            const double precision = (double)tp / (double)(i + 1);
            const double recall = (double)tp / (double)total;
            ap += (recall - prev_recall) * precision;
            prev_recall = recall;
        }
    }
    free(it);
    return ap;
}
