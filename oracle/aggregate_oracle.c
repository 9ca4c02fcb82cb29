/* Plain-C restatement of the reference frame->step collapse.
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Follows /root/reference/utils/aggregate.py:
 *   :55,65-72  200-frame window mode vote, lowest label wins ties
 *              (np.argmax(np.bincount(block)))
 *   :26-43     find_changes  (i>=1 with a[i]!=a[i-1], then len(a))
 *   :7-23      eliminate_consecutive_duplicates (run values)
 * Pinned by tests/test_oracle.py against the reference's golden pair
 * output_miniRoad/output_miniROAD.json -> data/output/aggregated_data.json.
 *
 * Build: see oracle/Makefile (gcc -O2 -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Window mode vote. labels[T] in [0, num_labels). out[T]. Returns 0, or -1 on bad label. */
int oracle_window_mode(const int32_t* labels, int64_t T, int32_t window,
                       int32_t num_labels, int32_t* out) {
    int64_t* counts = (int64_t*)malloc(sizeof(int64_t) * (size_t)num_labels);
    if (!counts) return -2;
    for (int64_t s = 0; s < T; s += window) {
        int64_t e = s + window < T ? s + window : T;
        memset(counts, 0, sizeof(int64_t) * (size_t)num_labels);
        for (int64_t i = s; i < e; ++i) {
            if (labels[i] < 0 || labels[i] >= num_labels) { free(counts); return -1; }
            counts[labels[i]]++;
        }
        int32_t best = 0;
        for (int32_t k = 1; k < num_labels; ++k)
            if (counts[k] > counts[best]) best = k; /* strict > : lowest label wins ties */
        for (int64_t i = s; i < e; ++i) out[i] = best;
    }
    free(counts);
    return 0;
}

/* Run-length collapse. vals/changes must hold up to T entries.
 * Returns the number of runs (>=1), or -1 when T == 0 (the reference raises IndexError). */
int64_t oracle_rle(const int32_t* a, int64_t T, int32_t* vals, int64_t* changes) {
    if (T <= 0) return -1;
    int64_t n = 0;
    vals[0] = a[0];
    for (int64_t i = 1; i < T; ++i) {
        if (a[i] != a[i - 1]) {
            changes[n] = i;
            ++n;
            vals[n] = a[i];
        }
    }
    changes[n] = T;
    return n + 1;
}

/* pred path: window vote then RLE, for one video. */
int64_t oracle_aggregate_pred(const int32_t* labels, int64_t T, int32_t window,
                              int32_t num_labels, int32_t* vals, int64_t* changes) {
    if (T <= 0) return -1;
    int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (size_t)T);
    if (!tmp) return -2;
    int rc = oracle_window_mode(labels, T, window, num_labels, tmp);
    if (rc != 0) { free(tmp); return -3; }
    int64_t n = oracle_rle(tmp, T, vals, changes);
    free(tmp);
    return n;
}
