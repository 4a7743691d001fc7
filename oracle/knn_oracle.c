/* TEST INFRASTRUCTURE ONLY.  Brute-force CPU restatement of simple-knn's results.
 *
 * Reference: /root/reference Reconstruct/submodules/simple-knn/simple_knn.cu
 *   :150-186 boxMeanDist  -> mean of the 3 smallest squared distances to OTHER points (self excluded by
 *                            position, so duplicates at distance 0 count), best[] kept ascending (:134-148),
 *                            result (best[0]+best[1]+best[2])/3.0f (:185).
 *   :391-421 updateKBest2 / :423-466 boxKnn2 -> exact K nearest other points (ids + squared distances),
 *                            slot order unspecified; unfilled slots keep FLT_MAX / -1 (spatial.cu:36-37).
 * The distance is d.x*d.x + d.y*d.y + d.z*d.z (:136,:400); nvcc's default -fmad=true contracts it to
 * fma(dz,dz, fma(dx,dx, dy*dy)) (SASS of the expression compiled for sm_100: FMUL dy,dy; FFMA dx,dx; FFMA dz,dz),
 * reproduced here with fmaf so distCUDA2 can be compared bit-for-bit (verified on the GPU against oracle/_ref).
 * PARITY UNPINNED by reference tests (none exist for simple-knn); pinned on the GPU box against the
 * reference extension rebuilt into oracle/_ref (tests/test_knn.py).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>

static inline float dist2(const float* a, const float* b) {
    float dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

void adbo_knn_mean3(int P, const float* pts, float* out) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
        for (int j = 0; j < P; ++j) {
            if (j == i) continue;
            float d = dist2(pts + 3 * i, pts + 3 * j);
            for (int k = 0; k < 3; ++k)
                if (best[k] > d) { float t = best[k]; best[k] = d; d = t; }
        }
        out[i] = (best[0] + best[1] + best[2]) / 3.0f;
    }
}

/* Exact K nearest (self excluded) sorted by (distance, id); unfilled -> FLT_MAX / -1.
 * query_idx nullable (then Q==P and query i is point i); candidate mask nullable. */
void adbo_knn_index(int P, const float* pts, int K, int Q, const int32_t* query_idx, const uint8_t* is_candidate,
                    float* dists, int32_t* ids) {
#pragma omp parallel for schedule(static)
    for (int qi = 0; qi < Q; ++qi) {
        int i = query_idx ? query_idx[qi] : qi;
        float* bd = dists + (int64_t)qi * K;
        int32_t* bi = ids + (int64_t)qi * K;
        for (int k = 0; k < K; ++k) { bd[k] = FLT_MAX; bi[k] = -1; }
        for (int j = 0; j < P; ++j) {
            if (j == i) continue;
            if (is_candidate && !is_candidate[j]) continue;
            float d = dist2(pts + 3 * i, pts + 3 * j);
            int32_t id = j;
            for (int k = 0; k < K; ++k) {
                if (d < bd[k] || (d == bd[k] && (bi[k] < 0 || id < bi[k]))) {
                    float td = bd[k]; int32_t ti = bi[k];
                    bd[k] = d; bi[k] = id; d = td; id = ti;
                }
            }
        }
    }
}
