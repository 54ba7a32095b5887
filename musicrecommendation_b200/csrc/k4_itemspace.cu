// k4_itemspace.cu — the item-space engine: both models as rows of (weighted) song-song co-occurrence matrices.
//
//   IBM  Sint_i[u][s] = sum_{j in I_u, j != s} qd[j] * G[j][s]      G[j][s]  = |U_j ∩ U_s| over train users   (MusicRecommender.scala:232-235, 249-257)
//   UBM  Sint_u[u][s] = sum_{j in I_u}         Gq[j][s]             Gq[j][s] = sum_{v in U_j ∩ U_s} qv[v]      (MusicRecommender.scala:142-148, 159-166:
//        sum_v |I_u ∩ I_v| qv[v] [s in I_v] regrouped by the shared song j — the same integers, summed in another order)
//
// Popular ("head") songs are shared by many test users, so their rows G[j][:], Gq[j][:] are computed once per train set
// (gram_head_* below, or the tcgen05 count GEMM for dense-friendly shapes) and kept in HBM as dense rows of 6 bytes per entry
// (u16 count + u32 weighted sum, overflowing entries in an exact exception list); a test user's score
// row is then a sum of |I_u ∩ head| coalesced, streaming row reads (head_rowsum_kernel — HBM-bandwidth bound).  The long tail
// of rarely heard songs is expanded on the fly through the inverted index with exact 64-bit integer atomics
// (tail_scatter_kernel).  Every entry is an exact integer, so the result equals the user-space engine and the oracle bit for bit.
#include "mr_common.cuh"
#include "mr_kernels.h"

#include <algorithm>

namespace mr {

// ---------------------------------------------------------------------------------------------------------------------
// Precompute of the head rows from the inverted index: one warp per (head song h, listener v of that song), lanes over I_v.
//   G[h][s] += 1, Gq[h][s] += qv[v]   for every s in I_v
// Work is the flattened list of (h, listener) pairs (lst_ptr = exclusive prefix of the head songs' train degrees) so that
// the 80k-listener rows and the 400-listener rows balance.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gram_head_scatter_kernel(const int* __restrict__ head_song, const long long* __restrict__ lst_ptr, int r0, int r1,
                         const long long* __restrict__ csc_ptr, const int* __restrict__ csc_idx,
                         const long long* __restrict__ tr_ptr, const int* __restrict__ tr_col, const uint32_t* __restrict__ qv,
                         uint32_t* __restrict__ g, unsigned long long* __restrict__ gq, long long pitch) {
  const int lane = threadIdx.x & 31;
  const long long w0 = lst_ptr[r0], w1 = lst_ptr[r1];
  const long long n_warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long w = w0 + static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); w < w1; w += n_warps) {
    int lo = r0, hi = r1;                          // row = upper_bound(lst_ptr, w) - 1 within the chunk [r0, r1)
    while (lo < hi) { const int m = (lo + hi) >> 1; if (lst_ptr[m + 1] <= w) lo = m + 1; else hi = m; }
    const int h = lo;
    const int j = head_song[h];
    const int v = csc_idx[csc_ptr[j] + (w - lst_ptr[h])];
    const unsigned long long q = qv[v];
    const long long b = tr_ptr[v], e = tr_ptr[v + 1];
    for (long long m = b + lane; m < e; m += 32) {
      const int s = __ldg(tr_col + m);
      atomicAdd(g + static_cast<long long>(h - r0) * pitch + s, 1u);
      atomicAdd(gq + static_cast<long long>(h - r0) * pitch + s, q);
    }
  }
}

int launch_gram_head_scatter(const int* head_song, const long long* lst_ptr, int r0, int r1, const long long* csc_ptr, const int* csc_idx,
                             const long long* tr_ptr, const int* tr_col, const uint32_t* qv, uint32_t* g, unsigned long long* gq,
                             long long pitch, int num_sms, cudaStream_t st) {
  if (r1 <= r0) return 0;
  cudaError_t e = cudaMemsetAsync(g, 0, static_cast<size_t>(r1 - r0) * pitch * sizeof(uint32_t), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(gq, 0, static_cast<size_t>(r1 - r0) * pitch * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return -1;
  gram_head_scatter_kernel<<<num_sms * 8, 256, 0, st>>>(head_song, lst_ptr, r0, r1, csc_ptr, csc_idx, tr_ptr, tr_col, qv, g, gq, pitch);   // exits at once where a chunk has little work
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------------------------------
// Pack a chunk of staged rows (u32 counts, u64 weighted sums) into the resident 6-byte-per-entry form: the low 16 bits of G
// and the low 32 bits of Gq.  The few entries that do not fit (pairs of very popular songs) keep their high parts exactly in
// an exception list (row, song, G - low, Gq - low) that head_fixup_kernel adds back.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_head_rows_kernel(const uint32_t* __restrict__ g, const unsigned long long* __restrict__ gq, int r0, int n_rows, long long pitch,
                      uint16_t* __restrict__ g16, uint32_t* __restrict__ gq32, HeadExceptions ex) {
  const long long n = static_cast<long long>(n_rows) * pitch;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint32_t a = g[i];
    const unsigned long long b = gq[i];
    const long long o = static_cast<long long>(r0) * pitch + i;
    g16[o] = static_cast<uint16_t>(a & 0xffffu);
    gq32[o] = static_cast<uint32_t>(b & 0xffffffffULL);
    if ((a >> 16) | (b >> 32)) {
      const unsigned int pos = atomicAdd(ex.count, 1u);
      if (pos < ex.capacity) {
        ex.row[pos] = r0 + static_cast<int>(i / pitch);
        ex.song[pos] = static_cast<int>(i % pitch);
        ex.g_extra[pos] = a & 0xffff0000u;
        ex.gq_extra[pos] = b & 0xffffffff00000000ULL;
      }
    }
  }
}

int launch_pack_head_rows(const uint32_t* g, const unsigned long long* gq, int r0, int n_rows, long long pitch, uint16_t* g16,
                          uint32_t* gq32, HeadExceptions ex, int num_sms, cudaStream_t st) {
  if (n_rows <= 0) return 0;
  const long long n = static_cast<long long>(n_rows) * pitch;
  const int grid = static_cast<int>(std::min<long long>(num_sms * 16LL, (n + 255) / 256));
  pack_head_rows_kernel<<<grid, 256, 0, st>>>(g, gq, r0, n_rows, pitch, g16, gq32, ex);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------------------------------
// Head part of a batch: Sint[b][s] = sum over the user's head entries of the precomputed (packed) rows.  Thread = (user b,
// 4 songs); every row read is a coalesced 8-byte (4 x u16 of G) or 16-byte (4 x u32 of Gq) vector load along the song axis;
// 4 rows are kept in flight.  kModels: 1 = UBM only, 2 = IBM only, 3 = both.
// ---------------------------------------------------------------------------------------------------------------------
template <int kModels>
__global__ void __launch_bounds__(256)
head_rowsum_kernel(const long long* __restrict__ hu_ptr, const int* __restrict__ hu_row, const int* __restrict__ hu_song,
                   const uint32_t* __restrict__ hu_q, int u0, const uint16_t* __restrict__ g16, const uint32_t* __restrict__ gq32,
                   long long pitch, int n_songs, long long* __restrict__ sint_u, long long* __restrict__ sint_i, long long spitch) {
  // blockIdx.x = user, blockIdx.y = song tile: CTAs that are resident together work on the SAME song tile for different users, so
  // the row tiles of popular songs (shared by many users of the batch) are fetched from HBM once and hit in L2 for the others
  const int b = blockIdx.x;
  const long long beg = hu_ptr[u0 + b], end = hu_ptr[u0 + b + 1];
  const int s = 4 * (blockIdx.y * blockDim.x + threadIdx.x);
  if (s >= n_songs) return;
  unsigned long long ua[4] = {0, 0, 0, 0}, ia[4] = {0, 0, 0, 0};
  long long i = beg;
  for (; i + 4 <= end; i += 4) {
    uint4 cu[4]; uint2 ci[4]; uint32_t q[4]; int js[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const long long row = static_cast<long long>(__ldg(hu_row + i + t)) * pitch + s;
      if (kModels & 1) cu[t] = __ldg(reinterpret_cast<const uint4*>(gq32 + row));
      if (kModels & 2) { ci[t] = __ldg(reinterpret_cast<const uint2*>(g16 + row)); q[t] = __ldg(hu_q + i + t); js[t] = __ldg(hu_song + i + t); }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (kModels & 1) { ua[0] += cu[t].x; ua[1] += cu[t].y; ua[2] += cu[t].z; ua[3] += cu[t].w; }
      if (kModels & 2) {
        const uint32_t c0 = ci[t].x & 0xffffu, c1 = ci[t].x >> 16, c2 = ci[t].y & 0xffffu, c3 = ci[t].y >> 16;
        const int d = js[t] - s;                                                          // s2 != song, MR:252
        if (d != 0) ia[0] += static_cast<unsigned long long>(c0) * q[t];
        if (d != 1) ia[1] += static_cast<unsigned long long>(c1) * q[t];
        if (d != 2) ia[2] += static_cast<unsigned long long>(c2) * q[t];
        if (d != 3) ia[3] += static_cast<unsigned long long>(c3) * q[t];
      }
    }
  }
  for (; i < end; ++i) {
    const long long row = static_cast<long long>(__ldg(hu_row + i)) * pitch + s;
    if (kModels & 1) { const uint4 c = __ldg(reinterpret_cast<const uint4*>(gq32 + row)); ua[0] += c.x; ua[1] += c.y; ua[2] += c.z; ua[3] += c.w; }
    if (kModels & 2) {
      const uint2 c = __ldg(reinterpret_cast<const uint2*>(g16 + row));
      const uint32_t q = __ldg(hu_q + i); const int d = __ldg(hu_song + i) - s;
      if (d != 0) ia[0] += static_cast<unsigned long long>(c.x & 0xffffu) * q;
      if (d != 1) ia[1] += static_cast<unsigned long long>(c.x >> 16) * q;
      if (d != 2) ia[2] += static_cast<unsigned long long>(c.y & 0xffffu) * q;
      if (d != 3) ia[3] += static_cast<unsigned long long>(c.y >> 16) * q;
    }
  }
  const long long o = static_cast<long long>(b) * spitch + s;
  if (kModels & 1) {
    *reinterpret_cast<ulonglong2*>(sint_u + o) = make_ulonglong2(ua[0], ua[1]);
    *reinterpret_cast<ulonglong2*>(sint_u + o + 2) = make_ulonglong2(ua[2], ua[3]);
  }
  if (kModels & 2) {
    *reinterpret_cast<ulonglong2*>(sint_i + o) = make_ulonglong2(ia[0], ia[1]);
    *reinterpret_cast<ulonglong2*>(sint_i + o + 2) = make_ulonglong2(ia[2], ia[3]);
  }
}

// IBM-only pass: the packed count rows are 2 bytes per song, so a thread takes 8 songs (one 16-byte load per row) to keep as
// many bytes in flight per thread as the UBM pass does.
__global__ void __launch_bounds__(256)
head_rowsum_ibm8_kernel(const long long* __restrict__ hu_ptr, const int* __restrict__ hu_row, const int* __restrict__ hu_song,
                        const uint32_t* __restrict__ hu_q, int u0, const uint16_t* __restrict__ g16, long long pitch, int n_songs,
                        long long* __restrict__ sint_i, long long spitch) {
  const int b = blockIdx.x;
  const long long beg = hu_ptr[u0 + b], end = hu_ptr[u0 + b + 1];
  const int s = 8 * (blockIdx.y * blockDim.x + threadIdx.x);
  if (s >= n_songs) return;
  unsigned long long ia[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto add_row = [&](const uint4& c, uint32_t q, int j) {
    const uint32_t w[4] = {c.x, c.y, c.z, c.w};
    const int d = j - s;                                                              // s2 != song, MR:252
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (d != 2 * t) ia[2 * t] += static_cast<unsigned long long>(w[t] & 0xffffu) * q;
      if (d != 2 * t + 1) ia[2 * t + 1] += static_cast<unsigned long long>(w[t] >> 16) * q;
    }
  };
  long long i = beg;
  for (; i + 4 <= end; i += 4) {
    uint4 c[4]; uint32_t q[4]; int js[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      c[t] = __ldg(reinterpret_cast<const uint4*>(g16 + static_cast<long long>(__ldg(hu_row + i + t)) * pitch + s));
      q[t] = __ldg(hu_q + i + t); js[t] = __ldg(hu_song + i + t);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) add_row(c[t], q[t], js[t]);
  }
  for (; i < end; ++i)
    add_row(__ldg(reinterpret_cast<const uint4*>(g16 + static_cast<long long>(__ldg(hu_row + i)) * pitch + s)), __ldg(hu_q + i), __ldg(hu_song + i));
  unsigned long long* o = reinterpret_cast<unsigned long long*>(sint_i) + static_cast<long long>(b) * spitch + s;
#pragma unroll
  for (int t = 0; t < 8; t += 2) *reinterpret_cast<ulonglong2*>(o + t) = make_ulonglong2(ia[t], ia[t + 1]);
}

int launch_head_rowsum(int models, const long long* hu_ptr, const int* hu_row, const int* hu_song, const uint32_t* hu_q, int u0,
                       int n_users, const uint16_t* g16, const uint32_t* gq32, long long pitch, int n_songs, long long* sint_u,
                       long long* sint_i, long long spitch, cudaStream_t st) {
  if (n_users <= 0 || n_songs <= 0) return 0;
  const dim3 grid(n_users, (n_songs + 1023) / 1024);
  if (models == 1) head_rowsum_kernel<1><<<grid, 256, 0, st>>>(hu_ptr, hu_row, hu_song, hu_q, u0, g16, gq32, pitch, n_songs, sint_u, sint_i, spitch);
  else if (models == 2 && pitch % 8 == 0 && spitch % 8 == 0)
    head_rowsum_ibm8_kernel<<<dim3(n_users, (n_songs + 2047) / 2048), 256, 0, st>>>(hu_ptr, hu_row, hu_song, hu_q, u0, g16, pitch, n_songs, sint_i, spitch);
  else if (models == 2) head_rowsum_kernel<2><<<grid, 256, 0, st>>>(hu_ptr, hu_row, hu_song, hu_q, u0, g16, gq32, pitch, n_songs, sint_u, sint_i, spitch);
  else head_rowsum_kernel<3><<<grid, 256, 0, st>>>(hu_ptr, hu_row, hu_song, hu_q, u0, g16, gq32, pitch, n_songs, sint_u, sint_i, spitch);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// The packed rows' exception entries: one CTA per user, threads over (head entry, exception of that entry's row).
template <int kModels>
__global__ void __launch_bounds__(128)
head_fixup_kernel(const long long* __restrict__ hu_ptr, const int* __restrict__ hu_row, const int* __restrict__ hu_song,
                  const uint32_t* __restrict__ hu_q, int u0, const long long* __restrict__ ex_ptr, const int* __restrict__ ex_song,
                  const uint32_t* __restrict__ ex_g, const unsigned long long* __restrict__ ex_gq, long long* __restrict__ sint_u,
                  long long* __restrict__ sint_i, long long spitch) {
  const int b = blockIdx.x;
  unsigned long long* su = reinterpret_cast<unsigned long long*>(sint_u) + static_cast<long long>(b) * spitch;
  unsigned long long* si = reinterpret_cast<unsigned long long*>(sint_i) + static_cast<long long>(b) * spitch;
  for (long long i = hu_ptr[u0 + b]; i < hu_ptr[u0 + b + 1]; ++i) {
    const int row = hu_row[i];
    const long long xb = ex_ptr[row], xe = ex_ptr[row + 1];
    if (xb == xe) continue;
    const int j = hu_song[i];
    const unsigned long long q = hu_q[i];
    for (long long x = xb + threadIdx.x; x < xe; x += blockDim.x) {
      const int s = ex_song[x];
      if ((kModels & 1) && ex_gq[x]) atomicAdd(su + s, ex_gq[x]);
      if ((kModels & 2) && s != j && ex_g[x]) atomicAdd(si + s, static_cast<unsigned long long>(ex_g[x]) * q);
    }
  }
}

int launch_head_fixup(int models, const long long* hu_ptr, const int* hu_row, const int* hu_song, const uint32_t* hu_q, int u0, int n_users,
                      const long long* ex_ptr, const int* ex_song, const uint32_t* ex_g, const unsigned long long* ex_gq,
                      long long* sint_u, long long* sint_i, long long spitch, cudaStream_t st) {
  if (n_users <= 0) return 0;
  if (models == 1) head_fixup_kernel<1><<<n_users, 128, 0, st>>>(hu_ptr, hu_row, hu_song, hu_q, u0, ex_ptr, ex_song, ex_g, ex_gq, sint_u, sint_i, spitch);
  else if (models == 2) head_fixup_kernel<2><<<n_users, 128, 0, st>>>(hu_ptr, hu_row, hu_song, hu_q, u0, ex_ptr, ex_song, ex_g, ex_gq, sint_u, sint_i, spitch);
  else head_fixup_kernel<3><<<n_users, 128, 0, st>>>(hu_ptr, hu_row, hu_song, hu_q, u0, ex_ptr, ex_song, ex_g, ex_gq, sint_u, sint_i, spitch);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------------------------------
// Tail part of a batch: for every (user b, visible tail song j): for v in U_j^train, for s in I_v, s != j:
//   Sint_u[b][s] += qv[v],  Sint_i[b][s] += qd[j]
// Work is the flattened list of (tail entry, listener) pairs (tu_lptr = exclusive prefix of the entries' train degrees): one
// warp per pair, lanes over I_v, exact u64 atomics (fire-and-forget).  Flattening keeps all 64 warp slots of every SM busy
// although entries have between 1 and a few hundred listeners.
// ---------------------------------------------------------------------------------------------------------------------
template <int kModels>
__global__ void __launch_bounds__(256)
tail_scatter_kernel(const int* __restrict__ tu_user, const int* __restrict__ tu_song, const long long* __restrict__ tu_lptr,
                    long long e0, long long e1, const long long* __restrict__ csc_ptr, const int* __restrict__ csc_idx,
                    const long long* __restrict__ tr_ptr, const int* __restrict__ tr_col, const uint32_t* __restrict__ qv,
                    const uint32_t* __restrict__ qd, int u0, long long* __restrict__ sint_u, long long* __restrict__ sint_i,
                    long long spitch) {
  const int lane = threadIdx.x & 31;
  const long long w0 = tu_lptr[e0], w1 = tu_lptr[e1];
  const long long n_warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long w = w0 + static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); w < w1; w += n_warps) {
    long long lo = e0, hi = e1;                      // e = upper_bound(tu_lptr, w) - 1 within [e0, e1)
    while (lo < hi) { const long long m = (lo + hi) >> 1; if (tu_lptr[m + 1] <= w) lo = m + 1; else hi = m; }
    const long long e = lo;
    const int b = tu_user[e] - u0;
    const int j = tu_song[e];
    const int v = csc_idx[csc_ptr[j] + (w - tu_lptr[e])];
    const unsigned long long q = qv[v], qj = qd[j];
    unsigned long long* su = reinterpret_cast<unsigned long long*>(sint_u) + static_cast<long long>(b) * spitch;
    unsigned long long* si = reinterpret_cast<unsigned long long*>(sint_i) + static_cast<long long>(b) * spitch;
    const long long rb = tr_ptr[v], re = tr_ptr[v + 1];
    for (long long m = rb + lane; m < re; m += 32) {
      const int s = __ldg(tr_col + m);
      if (s == j) continue;
      if (kModels & 1) atomicAdd(su + s, q);
      if (kModels & 2) atomicAdd(si + s, qj);
    }
  }
}

int launch_tail_scatter(int models, const int* tu_user, const int* tu_song, const long long* tu_lptr, long long e0, long long e1,
                        const long long* csc_ptr, const int* csc_idx, const long long* tr_ptr, const int* tr_col, const uint32_t* qv,
                        const uint32_t* qd, int u0, long long* sint_u, long long* sint_i, long long spitch, int num_sms, cudaStream_t st) {
  if (e1 <= e0) return 0;
  const int grid = num_sms * 8;
  if (models == 1) tail_scatter_kernel<1><<<grid, 256, 0, st>>>(tu_user, tu_song, tu_lptr, e0, e1, csc_ptr, csc_idx, tr_ptr, tr_col, qv, qd, u0, sint_u, sint_i, spitch);
  else if (models == 2) tail_scatter_kernel<2><<<grid, 256, 0, st>>>(tu_user, tu_song, tu_lptr, e0, e1, csc_ptr, csc_idx, tr_ptr, tr_col, qv, qd, u0, sint_u, sint_i, spitch);
  else tail_scatter_kernel<3><<<grid, 256, 0, st>>>(tu_user, tu_song, tu_lptr, e0, e1, csc_ptr, csc_idx, tr_ptr, tr_col, qv, qd, u0, sint_u, sint_i, spitch);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace mr
