/* CPU oracle for atlas ROI pooling, plain C.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this.  The product path never links or calls it.
 *
 * Restates /root/reference/image_features.py:80-82,111-114 (one-hot mask x
 * feature sum / clamp_min(count, 1e-6)) as a single pass over the label map,
 * plus the max/argmax extension defined in oracle/roi_oracle.py.  Sums are
 * accumulated in double.  Parity status: pinned through oracle/roi_oracle.py
 * (tests/test_roi_oracle.py checks this file against it and against the
 * golden fixtures).
 *
 * Build: make -C oracle   ->  oracle/_build/libroi_oracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <math.h>

/* feats: n_vols x n_voxels float32, labels: n_voxels int32 in [0, n_rois].
 * mean/max: n_vols x n_rois float32, argmax: n_vols x n_rois int32,
 * counts: n_rois int64.  Returns 0, or -1 on a label outside [0, n_rois]. */
int roi_pool_oracle_c(const float* feats, int64_t n_vols, int64_t n_voxels,
                      const int32_t* labels, int32_t n_rois,
                      float* mean, float* max, int32_t* argmax, int64_t* counts)
{
    double* sum = (double*)malloc(sizeof(double) * (size_t)(n_rois + 1));
    float* mx = (float*)malloc(sizeof(float) * (size_t)(n_rois + 1));
    int32_t* am = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_rois + 1));
    if (!sum || !mx || !am) { free(sum); free(mx); free(am); return -2; }
    for (int32_t r = 0; r < n_rois; ++r) counts[r] = 0;
    for (int64_t v = 0; v < n_voxels; ++v) {
        int32_t l = labels[v];
        if (l < 0 || l > n_rois) { free(sum); free(mx); free(am); return -1; }
        if (l) counts[l - 1]++;
    }
    for (int64_t n = 0; n < n_vols; ++n) {
        const float* f = feats + n * n_voxels;
        for (int32_t r = 0; r <= n_rois; ++r) { sum[r] = 0.0; mx[r] = -INFINITY; am[r] = -1; }
        for (int64_t v = 0; v < n_voxels; ++v) {
            int32_t l = labels[v];
            if (!l) continue;
            float x = f[v];
            sum[l] += (double)x;
            if (am[l] < 0 || x > mx[l]) { mx[l] = x; am[l] = (int32_t)v; }   /* first occurrence wins */
        }
        for (int32_t r = 1; r <= n_rois; ++r) {
            float den = (float)counts[r - 1];
            if (den < 1e-6f) den = 1e-6f;
            mean[n * n_rois + r - 1] = (float)sum[r] / den;
            max[n * n_rois + r - 1] = counts[r - 1] ? mx[r] : 0.0f;
            argmax[n * n_rois + r - 1] = counts[r - 1] ? am[r] : -1;
        }
    }
    free(sum); free(mx); free(am);
    return 0;
}
