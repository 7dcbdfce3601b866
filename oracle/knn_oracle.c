/* TEST INFRASTRUCTURE ONLY — CPU oracle for SURVEY.md §8 (f3), never on the product path.
 *
 * Plain-C brute-force restatement of what `simple_knn._C.distCUDA2` returns
 * (gaussiansplatting/submodules/simple-knn/simple_knn.cu): for every point the mean of the squared
 * distances to its three nearest OTHER points (self is excluded by index, duplicates count with
 * distance 0 — boxMeanDist :155-186), with
 *   - the distance rounded as the reference's GPU code rounds it: d = other - self,
 *     dist = fma(dz,dz, fma(dx,dx, dy*dy))  (updateKBest :133-146 as nvcc contracts `dx*dx + dy*dy + dz*dz`:
 *     FMUL dy*dy, FFMA dx*dx+., FFMA dz*dz+. in the SASS of oracle/_ref/libsimple_knn_ref.so);
 *   - the three best kept by the same compare-and-swap chain (:138-145), initial value FLT_MAX (:161);
 *   - the mean as ((b0 + b1) + b2) / 3.0f (:185).
 * The reference's Morton sort and box culling (:186-221) only prune the search; they do not change the
 * result, so the restatement is an exhaustive O(P^2) scan.
 *
 * PINNED: tests/test_gpu_knn.py checks this file bit-for-bit against the reference itself
 * (oracle/_ref/libsimple_knn_ref.so, built from the reference's sources by oracle/Makefile) on the GPU box.
 */
#include <float.h>
#include <math.h>
#include <stddef.h>

#include <pthread.h>

typedef struct { long long P, begin, end; const float* pts; float* out; } Job;

#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target_clones("fma", "default")))
#endif
static void scan_rows(long long P, long long begin, long long end, const float* pts, float* mean_dist2) {
  for (long long i = begin; i < end; ++i) {
    const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    for (long long j = 0; j < P; ++j) {
      if (j == i) continue;
      const float dx = pts[3 * j] - x, dy = pts[3 * j + 1] - y, dz = pts[3 * j + 2] - z;
      float dist = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
      for (int k = 0; k < 3; ++k) {
        if (best[k] > dist) { const float t = best[k]; best[k] = dist; dist = t; }
      }
    }
    mean_dist2[i] = ((best[0] + best[1]) + best[2]) / 3.0f;
  }
}

static void* worker(void* arg) {
  Job* j = (Job*)arg;
  scan_rows(j->P, j->begin, j->end, j->pts, j->out);
  return NULL;
}

/* threads <= 1 runs inline; rows are independent, so the result does not depend on the thread count. */
int knn_oracle_dist2(long long P, const float* pts, float* mean_dist2, int threads) {
  if (P < 0 || (P > 0 && (!pts || !mean_dist2))) return 1;
  if (threads > 64) threads = 64;
  if (threads <= 1 || P < 256) { scan_rows(P, 0, P, pts, mean_dist2); return 0; }
  pthread_t tid[64];
  Job job[64];
  const long long per = (P + threads - 1) / threads;
  int started = 0;
  for (int t = 0; t < threads; ++t) {
    long long b = t * per, e = b + per < P ? b + per : P;
    if (b >= e) break;
    job[t] = (Job){P, b, e, pts, mean_dist2};
    if (pthread_create(&tid[t], NULL, worker, &job[t]) != 0) { scan_rows(P, b, e, pts, mean_dist2); tid[t] = 0; }
    started = t + 1;
  }
  for (int t = 0; t < started; ++t) if (tid[t]) pthread_join(tid[t], NULL);
  return 0;
}
