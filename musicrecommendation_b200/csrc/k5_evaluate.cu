// k5_evaluate.cu — the reference's evaluation metric on the GPU.
//
// MusicRecommender.scala:521-639: scores are min-max normalised over the WHOLE model (MR:524-529); for each threshold t in
// 0.0, 0.1, ... (10 values, MR:590; distributed.scala:395 uses 11) the predictions are { (u, s) : normalised score > t }; for each
// distinct label song c the confusion counts over test users (MR:541-553) give precision / recall (MR:561-579) and
//   AP(c) = sum_{i < n-2} (R_i - R_{i+1}) P_i + R_{n-2} P_{n-2} + 0        (MR:600-610, List.sum = left fold from 0.0)
// mAP = (sum_c AP(c)) / |label songs|  (MR:626, summed on the host in ascending song order = the oracle's canonical order).
// The reference spends 62-75 s per model on this at 2000 / 100 users (README:940-944).  All fp64 operations use explicit
// round-to-nearest intrinsics, so the result is bit-identical to the CPU restatement.
#include "mr_common.cuh"
#include "mr_kernels.h"

namespace mr {

// min / max over the emitted (non-NaN) scores; scores are >= 0 so their bit patterns order like unsigned integers
__global__ void eval_minmax_kernel(const double* __restrict__ scores, long long n, unsigned long long* __restrict__ mn,
                                   unsigned long long* __restrict__ mx) {
  unsigned long long lo = ~0ULL, hi = 0ULL;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double x = scores[i];
    if (x == x) {
      const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(x));
      lo = b < lo ? b : lo; hi = b > hi ? b : hi;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo, o), c = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = a < lo ? a : lo; hi = c > hi ? c : hi;
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(mn, lo); atomicMax(mx, hi); }
}

// One warp per label song: lanes stride over the test users, count TP / FP / FN per threshold, lane 0 finishes the AP.
__global__ void __launch_bounds__(128)
eval_ap_kernel(const double* __restrict__ scores, int n_users, int n_songs, const long long* __restrict__ lab_ptr,
               const int* __restrict__ lab_col, const int* __restrict__ new_songs, int n_new, int n_thresholds,
               const unsigned long long* __restrict__ mn_bits, const unsigned long long* __restrict__ mx_bits, double* __restrict__ ap_out) {
  const double TH[11] = {0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0};
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= n_new) return;
  const int song = new_songs[c];
  const double mn = __longlong_as_double(static_cast<long long>(*mn_bits)), mx = __longlong_as_double(static_cast<long long>(*mx_bits));
  const double range = __dsub_rn(mx, mn);
  int tp[11], fp[11], fn[11];
#pragma unroll
  for (int t = 0; t < 11; ++t) { tp[t] = 0; fp[t] = 0; fn[t] = 0; }
  for (int u = lane; u < n_users; u += 32) {
    // testLabels(user).contains(song): label rows are ascending
    long long lo = lab_ptr[u], hi = lab_ptr[u + 1];
    const long long end = hi;
    while (lo < hi) { const long long m = (lo + hi) >> 1; if (lab_col[m] < song) lo = m + 1; else hi = m; }
    const bool labelled = lo < end && lab_col[lo] == song;
    double norm = __longlong_as_double(0x7ff8000000000000LL);
    if (song < n_songs) {
      const double x = scores[static_cast<long long>(u) * n_songs + song];
      if (x == x) norm = __ddiv_rn(__dsub_rn(x, mn), range);                  // (score - min) / (max - min), MR:529
    }
#pragma unroll
    for (int t = 0; t < 11; ++t) {
      if (t < n_thresholds) {
        const bool predicted = norm > TH[t];                                   // NaN compares false
        tp[t] += predicted && labelled; fp[t] += predicted && !labelled; fn[t] += !predicted && labelled;
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 11; ++t)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tp[t] += __shfl_xor_sync(0xffffffffu, tp[t], o); fp[t] += __shfl_xor_sync(0xffffffffu, fp[t], o); fn[t] += __shfl_xor_sync(0xffffffffu, fn[t], o);
    }
  if (lane == 0) {
    double prec[11], rec[11];
#pragma unroll
    for (int t = 0; t < 11; ++t) {
      prec[t] = (tp[t] + fp[t] > 0) ? __ddiv_rn(static_cast<double>(tp[t]), static_cast<double>(tp[t] + fp[t])) : 0.0;   // MR:561-566
      rec[t] = (tp[t] + fn[t] > 0) ? __ddiv_rn(static_cast<double>(tp[t]), static_cast<double>(tp[t] + fn[t])) : 0.0;    // MR:574-579
    }
    double ap = 0.0;
#pragma unroll
    for (int t = 0; t < 11; ++t) {
      if (t < n_thresholds) {
        double term;
        if (t == n_thresholds - 1) term = 0.0;
        else if (t == n_thresholds - 2) term = __dmul_rn(__dsub_rn(rec[t], 0.0), prec[t]);
        else term = __dmul_rn(__dsub_rn(rec[t], rec[t + 1]), prec[t]);
        ap = __dadd_rn(ap, term);
      }
    }
    ap_out[c] = ap;
  }
}

int launch_evaluate(const double* scores, int n_users, int n_songs, const long long* lab_ptr, const int* lab_col, const int* new_songs,
                    int n_new, int n_thresholds, unsigned long long* minmax, double* ap_out, int num_sms, cudaStream_t st) {
  if (n_new <= 0 || n_thresholds < 2 || n_thresholds > 11) return -2;
  const unsigned long long init[2] = {~0ULL, 0ULL};
  if (cudaMemcpyAsync(minmax, init, sizeof init, cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
  eval_minmax_kernel<<<num_sms * 8, 256, 0, st>>>(scores, static_cast<long long>(n_users) * n_songs, minmax, minmax + 1);
  eval_ap_kernel<<<(n_new + 3) / 4, 128, 0, st>>>(scores, n_users, n_songs, lab_ptr, lab_col, new_songs, n_new, n_thresholds, minmax, minmax + 1, ap_out);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------------------------------
// mAP@k of the ranked lists (north_star's "mAP@500"; the reference has no ranking, so this is the Million Song Dataset Challenge
// definition, stated in include/mrscore.h at mr_map_at_k):  AP@k(u) = (sum_i [r_i in L_u] * hits_i / i) / min(|L_u|, k).
// One warp per test user: the lanes test 32 ranks at a time against the user's ascending label row (binary search), lane 0 adds the
// terms of the relevant ranks in rank order with explicit round-to-nearest operations, so the per-user AP is bit-identical to the CPU
// restatement; the mean over users is taken on the host in ascending user order.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
map_at_k_kernel(const int* __restrict__ top_song, const int* __restrict__ top_len, int n_users, int k, const long long* __restrict__ lab_ptr,
                const int* __restrict__ lab_col, double* __restrict__ ap_out) {
  const int lane = threadIdx.x & 31;
  const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= n_users) return;
  const long long l0 = lab_ptr[u], l1 = lab_ptr[u + 1];
  const long long nl = l1 - l0;
  if (nl <= 0) { if (lane == 0) ap_out[u] = 0.0; return; }
  const int n = min(top_len[u], k);
  const int* row = top_song + static_cast<long long>(u) * k;
  int hits = 0;
  double ap = 0.0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    bool rel = false;
    if (i < n) {
      const int s = row[i];
      long long lo = l0, hi = l1;
      while (lo < hi) { const long long m = (lo + hi) >> 1; if (lab_col[m] < s) lo = m + 1; else hi = m; }
      rel = lo < l1 && lab_col[lo] == s;
    }
    uint32_t m = __ballot_sync(0xffffffffu, rel);
    if (lane == 0) {
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        ++hits;
        ap = __dadd_rn(ap, __ddiv_rn(static_cast<double>(hits), static_cast<double>(base + b + 1)));
      }
    }
  }
  if (lane == 0) ap_out[u] = __ddiv_rn(ap, static_cast<double>(nl < k ? nl : static_cast<long long>(k)));
}

int launch_map_at_k(const int* top_song, const int* top_len, int n_users, int k, const long long* lab_ptr, const int* lab_col,
                    double* ap_out, cudaStream_t st) {
  if (n_users <= 0) return 0;
  map_at_k_kernel<<<(n_users + 3) / 4, 128, 0, st>>>(top_song, top_len, n_users, k, lab_ptr, lab_col, ap_out);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace mr
