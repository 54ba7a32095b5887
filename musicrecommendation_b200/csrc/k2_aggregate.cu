// k2_aggregate.cu — kernel K2: similarity-weighted score aggregation (HBM-bound gather).
//
// UBM  Sint[b][s] = sum_{v in U_s^train} Ct[v][b] * qv[v]              (MusicRecommender.scala:159-166: rank = sum of
//                                                                     cosineSimilarity(user, u2) over train users who heard s)
// IBM  Sint[b][s] = sum_{j in I_u, j != s} G[row(j)][s] * qd[j]        (MusicRecommender.scala:249-257)
//
// All accumulation is exact 64-bit integer arithmetic (canonical fixed point), so the result does not depend
// on the summation order, on how a song's listeners are split across warps, or on how test users are sharded across GPUs.
#include "mr_common.cuh"
#include "mr_kernels.h"

namespace mr {

// One warp per work item (a song, or a <= kSplitLen slice of a popular song's listener list).  The 32 lanes hold the 128
// test users of the batch, 4 per lane: every gathered panel row is one coalesced read (UBM: 4 x u16 = 8 bytes per lane,
// 256-byte rows; IBM user space: 4 x u32 = 16 bytes per lane, 512-byte rows).  Listener ids and their q factors are fetched
// 32 at a time (coalesced) and broadcast with shuffles; 8 row gathers are kept in flight per warp.
template <bool kUbm>
__global__ void __launch_bounds__(256)
aggregate_panel_kernel(AggItems items, const int* __restrict__ csc_idx, const uint32_t* __restrict__ qv,
                       const void* __restrict__ panel, long long* __restrict__ sint, long long spitch) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const uint2* __restrict__ ct2 = reinterpret_cast<const uint2*>(panel);   // UBM: row = 32 uint2
  const uint4* __restrict__ wi4 = reinterpret_cast<const uint4*>(panel);   // IBM: row = 32 uint4
  for (int it = blockIdx.x * warps_per_block + (threadIdx.x >> 5); it < items.n_items; it += gridDim.x * warps_per_block) {
    const int song = items.song[it];
    const long long beg = items.begin[it];
    const int len = items.len[it];
    unsigned long long a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int k = 0; k < len; k += 32) {
      const int n = min(32, len - k);
      int v = 0; uint32_t q = 0;
      if (lane < n) { v = __ldg(csc_idx + beg + k + lane); if (kUbm) q = __ldg(qv + v); }
      int j = 0;
      for (; j + 8 <= n; j += 8) {
        if (kUbm) {
          uint2 c[8]; uint32_t qq[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int vj = __shfl_sync(0xffffffffu, v, j + t);
            qq[t] = __shfl_sync(0xffffffffu, q, j + t);
            c[t] = __ldg(ct2 + static_cast<long long>(vj) * 32 + lane);
          }
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            a0 += static_cast<unsigned long long>(c[t].x & 0xffffu) * qq[t];
            a1 += static_cast<unsigned long long>(c[t].x >> 16) * qq[t];
            a2 += static_cast<unsigned long long>(c[t].y & 0xffffu) * qq[t];
            a3 += static_cast<unsigned long long>(c[t].y >> 16) * qq[t];
          }
        } else {
          uint4 c[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int vj = __shfl_sync(0xffffffffu, v, j + t);
            c[t] = __ldg(wi4 + static_cast<long long>(vj) * 32 + lane);
          }
#pragma unroll
          for (int t = 0; t < 8; ++t) { a0 += c[t].x; a1 += c[t].y; a2 += c[t].z; a3 += c[t].w; }
        }
      }
      for (; j < n; ++j) {
        const int vj = __shfl_sync(0xffffffffu, v, j);
        if (kUbm) {
          const uint32_t qj = __shfl_sync(0xffffffffu, q, j);
          const uint2 c = __ldg(ct2 + static_cast<long long>(vj) * 32 + lane);
          a0 += static_cast<unsigned long long>(c.x & 0xffffu) * qj;
          a1 += static_cast<unsigned long long>(c.x >> 16) * qj;
          a2 += static_cast<unsigned long long>(c.y & 0xffffu) * qj;
          a3 += static_cast<unsigned long long>(c.y >> 16) * qj;
        } else {
          const uint4 c = __ldg(wi4 + static_cast<long long>(vj) * 32 + lane);
          a0 += c.x; a1 += c.y; a2 += c.z; a3 += c.w;
        }
      }
    }
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(sint) + static_cast<long long>(4 * lane) * spitch + song;
    if (items.split[it]) {
      atomicAdd(dst, a0); atomicAdd(dst + spitch, a1); atomicAdd(dst + 2 * spitch, a2); atomicAdd(dst + 3 * spitch, a3);
    } else {
      dst[0] = a0; dst[spitch] = a1; dst[2 * spitch] = a2; dst[3 * spitch] = a3;
    }
  }
}

template <bool kUbm>
static int launch_panel(const AggItems& items, const int* csc_idx, const uint32_t* qv, const void* panel, long long* sint,
                        long long spitch, int num_sms, cudaStream_t st) {
  if (items.n_items <= 0) return 0;
  const int threads = 256, wpb = threads / 32;
  long long blocks = (static_cast<long long>(items.n_items) + wpb - 1) / wpb;
  const long long cap = static_cast<long long>(num_sms) * 8;   // 8 resident CTAs of 256 threads per SM (64 warps)
  if (blocks > cap) blocks = cap;
  aggregate_panel_kernel<kUbm><<<static_cast<int>(blocks), threads, 0, st>>>(items, csc_idx, qv, panel, sint, spitch);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_aggregate_ubm(const AggItems& items, const int* csc_idx, const uint32_t* qv, const uint16_t* ct, long long n_train,
                         long long* sint, long long spitch, int num_sms, cudaStream_t st) {
  (void)n_train;
  return launch_panel<true>(items, csc_idx, qv, ct, sint, spitch, num_sms, st);
}

int launch_aggregate_w32(const AggItems& items, const int* csc_idx, const uint32_t* wi, long long n_train, long long* sint,
                         long long spitch, int num_sms, cudaStream_t st) {
  (void)n_train;
  return launch_panel<false>(items, csc_idx, nullptr, wi, sint, spitch, num_sms, st);
}

// IBM: one thread per (user, song); the |I_u| Gram rows of the user are read coalesced along the song axis.
__global__ void __launch_bounds__(256)
aggregate_ibm_kernel(const long long* __restrict__ te_ptr, const int* __restrict__ te_col, const int* __restrict__ te_grow,
                     const uint32_t* __restrict__ qd, int u0, const int32_t* __restrict__ g, long long ldg, int n_songs,
                     long long* __restrict__ sint, long long spitch) {
  const int b = blockIdx.y;
  const long long beg = te_ptr[u0 + b], end = te_ptr[u0 + b + 1];
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_songs; s += gridDim.x * blockDim.x) {
    unsigned long long acc = 0;
    for (long long i = beg; i < end; ++i) {
      const int j = __ldg(te_col + i);
      const int32_t cnt = __ldg(g + static_cast<long long>(__ldg(te_grow + i)) * ldg + s);
      if (j != s) acc += static_cast<unsigned long long>(static_cast<uint32_t>(cnt)) * __ldg(qd + j);   // s2 != song, MR:252
    }
    sint[static_cast<long long>(b) * spitch + s] = static_cast<long long>(acc);
  }
}

int launch_aggregate_ibm(const long long* te_ptr, const int* te_col, const int* te_grow, const uint32_t* qd, int u0, int n_users,
                         const int32_t* g, long long ldg, int n_songs, long long* sint, long long spitch, cudaStream_t st) {
  if (n_users <= 0 || n_songs <= 0) return 0;
  int gx = (n_songs + 255) / 256;
  if (gx > 1024) gx = 1024;
  aggregate_ibm_kernel<<<dim3(gx, n_users), 256, 0, st>>>(te_ptr, te_col, te_grow, qd, u0, g, ldg, n_songs, sint, spitch);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace mr
