// k4_itemspace.cu — the item-space engine: both models as rows of (weighted) song-song co-occurrence matrices.
//
//   IBM  Sint_i[u][s] = sum_{j in I_u, j != s} qd[j] * G[j][s]      G[j][s]  = |U_j ∩ U_s| over train users   (MusicRecommender.scala:232-235, 249-257)
//   UBM  Sint_u[u][s] = sum_{j in I_u}         Gq[j][s]             Gq[j][s] = sum_{v in U_j ∩ U_s} qv[v]      (MusicRecommender.scala:142-148, 159-166:
//        sum_v |I_u ∩ I_v| qv[v] [s in I_v] regrouped by the shared song j — the same integers, summed in another order)
//
// Popular ("head") songs are shared by many test users, so their rows G[j][:], Gq[j][:] are computed once per train set
// (gram_head_* below, or the tcgen05 count GEMM for dense-friendly shapes) and kept in HBM as dense rows of 6 bytes per entry
// (u16 count + u32 weighted sum, overflowing entries in an exact exception list); a test user's score
// row is then a sum of |I_u ∩ head| coalesced row reads (head_rowsum_kernel — each distinct row crosses HBM once per batch, the rest is
// served from L2; bound by the L2 -> SM fabric).  The long tail
// of rarely heard songs is expanded on the fly through the inverted index with exact 64-bit integer atomics
// (tail_scatter_kernel).  Every entry is an exact integer, so the result equals the user-space engine and the oracle bit for bit.
#include "mr_common.cuh"
#include "mr_kernels.h"

#include <algorithm>
#include <cstdlib>

namespace mr {

// ---------------------------------------------------------------------------------------------------------------------
// Precompute of the head rows from the inverted index: one warp per (head song h, listener v of that song), lanes over I_v.
//   G[h][s] += 1, Gq[h][s] += qv[v]   for every s in I_v
// Work is the flattened list of (h, listener) pairs (lst_ptr = exclusive prefix of the head songs' train degrees) so that
// the 80k-listener rows and the 400-listener rows balance.
// ---------------------------------------------------------------------------------------------------------------------
// kPacked: one 64-bit atomic per event instead of two — the count rides in bits 44..63 of the weighted-sum accumulator (the caller
// checked that no weighted sum can reach 2^44 and no count 2^20); halves the L2 atomic traffic the precompute is bound by.
constexpr int kPackShift = 44;

template <bool kPacked>
__global__ void __launch_bounds__(256)
gram_head_scatter_kernel(const int* __restrict__ head_song, const long long* __restrict__ lst_ptr, int r0, int r1,
                         const long long* __restrict__ csc_ptr, const int* __restrict__ csc_idx,
                         const long long* __restrict__ tr_ptr, const long long* __restrict__ tr_end, const int* __restrict__ tr_col,
                         const uint32_t* __restrict__ qv, uint32_t* __restrict__ g, unsigned long long* __restrict__ gq, long long pitch) {
  const int lane = threadIdx.x & 31;
  const long long w0 = lst_ptr[r0], w1 = lst_ptr[r1];
  const long long n_warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long w = w0 + static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); w < w1; w += n_warps) {
    int lo = r0, hi = r1;                          // row = upper_bound(lst_ptr, w) - 1 within the chunk [r0, r1)
    while (lo < hi) { const int m = (lo + hi) >> 1; if (lst_ptr[m + 1] <= w) lo = m + 1; else hi = m; }
    const int h = lo;
    const int j = head_song[h];
    const int v = csc_idx[csc_ptr[j] + (w - lst_ptr[h])];
    const unsigned long long q = kPacked ? (static_cast<unsigned long long>(qv[v]) + (1ULL << kPackShift)) : qv[v];
    const long long b = tr_ptr[v], e = tr_end[v];
    for (long long m = b + lane; m < e; m += 32) {
      const int s = __ldg(tr_col + m);
      if (!kPacked) atomicAdd(g + static_cast<long long>(h - r0) * pitch + s, 1u);
      atomicAdd(gq + static_cast<long long>(h - r0) * pitch + s, q);
    }
  }
}

int launch_gram_head_scatter(const int* head_song, const long long* lst_ptr, int r0, int r1, const long long* csc_ptr, const int* csc_idx,
                             const long long* tr_ptr, const long long* tr_end, const int* tr_col, const uint32_t* qv, uint32_t* g,
                             unsigned long long* gq, long long pitch, int packed, int num_sms, cudaStream_t st) {
  if (r1 <= r0) return 0;
  cudaError_t e = cudaSuccess;
  if (!packed) e = cudaMemsetAsync(g, 0, static_cast<size_t>(r1 - r0) * pitch * sizeof(uint32_t), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(gq, 0, static_cast<size_t>(r1 - r0) * pitch * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return -1;
  // exits at once where a chunk has little work
  if (packed) gram_head_scatter_kernel<true><<<num_sms * 8, 256, 0, st>>>(head_song, lst_ptr, r0, r1, csc_ptr, csc_idx, tr_ptr, tr_end, tr_col, qv, g, gq, pitch);
  else gram_head_scatter_kernel<false><<<num_sms * 8, 256, 0, st>>>(head_song, lst_ptr, r0, r1, csc_ptr, csc_idx, tr_ptr, tr_end, tr_col, qv, g, gq, pitch);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------------------------------
// The same events added straight into the FINAL packed rows, for head songs none of whose entries can overflow (weighted listener
// sum < 2^32, train degree < 2^16 — checked per row on the host): Gq32[h][s] += qv[v] and G16[h][s] += 1 as two 32-bit L2 atomics
// (the u16 count is bumped through the aligned 32-bit word that holds it: a count < 2^16 never carries into its neighbour).  The
// caller hands over an L2-sized chunk of rows: they are zeroed here (the lines stay dirty in L2), receive their events as L2 hits
// and drain to HBM once — 6 bytes of DRAM traffic per entry, no staging area, no pack pass.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gram_head_direct_kernel(const int* __restrict__ head_song, const long long* __restrict__ lst_ptr, int r0, int r1,
                        const long long* __restrict__ csc_ptr, const int* __restrict__ csc_idx,
                        const long long* __restrict__ tr_ptr, const long long* __restrict__ tr_end, const int* __restrict__ tr_col,
                        const uint32_t* __restrict__ qv, uint16_t* __restrict__ g16, uint32_t* __restrict__ gq32, long long pitch) {
  const int lane = threadIdx.x & 31;
  const long long w0 = lst_ptr[r0], w1 = lst_ptr[r1];
  const long long n_warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  uint32_t* g_words = reinterpret_cast<uint32_t*>(g16);
  for (long long w = w0 + static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); w < w1; w += n_warps) {
    int lo = r0, hi = r1;                          // row = upper_bound(lst_ptr, w) - 1 within the chunk [r0, r1)
    while (lo < hi) { const int m = (lo + hi) >> 1; if (lst_ptr[m + 1] <= w) lo = m + 1; else hi = m; }
    const int h = lo;
    const int j = head_song[h];
    const int v = csc_idx[csc_ptr[j] + (w - lst_ptr[h])];
    const uint32_t q = qv[v];
    const long long row = static_cast<long long>(h) * pitch;   // pitch is a multiple of 32: the parity of row + s is the parity of s
    const long long b = tr_ptr[v], e = tr_end[v];
    for (long long m = b + lane; m < e; m += 32) {
      const int s = __ldg(tr_col + m);
      atomicAdd(gq32 + row + s, q);
      atomicAdd(g_words + ((row + s) >> 1), (s & 1) ? 0x10000u : 1u);
    }
  }
}

int launch_gram_head_direct(const int* head_song, const long long* lst_ptr, int r0, int r1, const long long* csc_ptr, const int* csc_idx,
                            const long long* tr_ptr, const long long* tr_end, const int* tr_col, const uint32_t* qv, uint16_t* g16,
                            uint32_t* gq32, long long pitch, int num_sms, cudaStream_t st) {
  if (r1 <= r0) return 0;
  cudaError_t e = cudaMemsetAsync(g16 + static_cast<long long>(r0) * pitch, 0, static_cast<size_t>(r1 - r0) * pitch * sizeof(uint16_t), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(gq32 + static_cast<long long>(r0) * pitch, 0, static_cast<size_t>(r1 - r0) * pitch * sizeof(uint32_t), st);
  if (e != cudaSuccess) return -1;
  gram_head_direct_kernel<<<num_sms * 8, 256, 0, st>>>(head_song, lst_ptr, r0, r1, csc_ptr, csc_idx, tr_ptr, tr_end, tr_col, qv, g16, gq32, pitch);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------------------------------
// Pack a chunk of staged rows (u32 counts, u64 weighted sums) into the resident 6-byte-per-entry form: the low 16 bits of G
// and the low 32 bits of Gq.  The few entries that do not fit (pairs of very popular songs) keep their high parts exactly in
// an exception list (row, song, G - low, Gq - low) that head_fixup_kernel adds back.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_head_rows_kernel(const uint32_t* __restrict__ g, const unsigned long long* __restrict__ gq, int packed, int r0, int n_rows, long long pitch,
                      int n_songs, uint16_t* __restrict__ g16, uint32_t* __restrict__ gq32, HeadExceptions ex) {
  const long long n = static_cast<long long>(n_rows) * pitch;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned long long raw = gq[i];
    const uint32_t a = packed ? static_cast<uint32_t>(raw >> kPackShift) : g[i];
    const unsigned long long b = packed ? (raw & ((1ULL << kPackShift) - 1)) : raw;
    const long long o = static_cast<long long>(r0) * pitch + i;
    g16[o] = static_cast<uint16_t>(a & 0xffffu);
    gq32[o] = static_cast<uint32_t>(b & 0xffffffffULL);
    if (((a >> 16) | (b >> 32)) && static_cast<int>(i % pitch) < n_songs) {   // pad columns hold zeros; never let them reach the list
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

int launch_pack_head_rows(const uint32_t* g, const unsigned long long* gq, int packed, int r0, int n_rows, long long pitch, int n_songs,
                          uint16_t* g16, uint32_t* gq32, HeadExceptions ex, int num_sms, cudaStream_t st) {
  if (n_rows <= 0) return 0;
  const long long n = static_cast<long long>(n_rows) * pitch;
  const int grid = static_cast<int>(std::min<long long>(num_sms * 16LL, (n + 255) / 256));
  pack_head_rows_kernel<<<grid, 256, 0, st>>>(g, gq, packed, r0, n_rows, pitch, n_songs, g16, gq32, ex);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------------------------------
// Head part of a batch: Sint[b][s] = sum over the user's head entries of the precomputed (packed) rows.
//
// HBM-bandwidth bound, and most of the gathered bytes are rows of popular songs that many users of the batch share, so the kernel
// is organised for L2 reuse: the grid is (group, song tile) with the group index fastest and exactly as many groups as CTAs fit on
// the GPU, i.e. one wave of CTAs = one song tile.  A group is a list of work segments (user, range of its head entries) that the
// host balanced to equal length (longest-processing-time packing; users with more entries than a group's share are split into
// several segments that accumulate with atomics into a pre-zeroed row).  Equal work keeps all resident CTAs on the same song tile
// at the same time, so a row tile is fetched from HBM once and served from L2 to every other user that needs it — the DRAM traffic
// of a pass approaches "each distinct head row of the batch once + the Sint rows written".
// Thread = kVec consecutive songs; every row read is one coalesced vector load along the song axis (4 x u32 of Gq32 / 8 x u16 of
// G16 at the widest), several rows in flight per thread.  Sint is written with streaming stores (it is re-read only by top-k, long
// after it has left L2).
// ---------------------------------------------------------------------------------------------------------------------
template <int W> struct RowWords { uint32_t w[W]; };
template <int W> __device__ __forceinline__ RowWords<W> ld_row_words(const void* p);
template <> __device__ __forceinline__ RowWords<1> ld_row_words<1>(const void* p) { RowWords<1> r; r.w[0] = __ldg(static_cast<const uint32_t*>(p)); return r; }
template <> __device__ __forceinline__ RowWords<2> ld_row_words<2>(const void* p) {
  const uint2 v = __ldg(static_cast<const uint2*>(p)); RowWords<2> r; r.w[0] = v.x; r.w[1] = v.y; return r;
}
template <> __device__ __forceinline__ RowWords<4> ld_row_words<4>(const void* p) {
  const uint4 v = __ldg(static_cast<const uint4*>(p)); RowWords<4> r; r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w; return r;
}

// kIbm = false: UBM pass over Gq32 (W songs per thread).  kIbm = true: IBM pass over G16 (2 W songs per thread, weight qd[j] per row).
// The reference's `s2 != song` (MusicRecommender.scala:252) needs no test here: s == j happens only at s in I_u, a listened pair that
// getModel never emits (MR:109) — mask_listened_kernel overwrites those entries before anything reads them.
// A group's segments and entries (head row index, weight) are contiguous in the group-ordered arrays the host built; the CTA stages
// them in shared memory with one coalesced round of loads, so the inner loop's only global accesses are the row loads themselves
// (no dependent index -> row chain per round).
template <bool kIbm, int W>
__global__ void __launch_bounds__(256, kHeadCtasPerSm)
head_rowsum_kernel(const int4* __restrict__ grp_hdr, const int4* __restrict__ seg, const int* __restrict__ ge_row,
                   const uint32_t* __restrict__ ge_q, int seg_cap, int ent_cap, const uint16_t* __restrict__ g16,
                   const uint32_t* __restrict__ gq32, long long pitch, int n_songs, long long* __restrict__ sint, long long spitch) {
  constexpr int kVec = kIbm ? 2 * W : W;
  constexpr int kRows = kIbm ? (W == 1 ? 4 : 2) : (W == 4 ? 4 : 8);   // rows in flight per thread
  extern __shared__ int4 s_dyn[];
  int4* s_seg = s_dyn;                                                // [seg_cap]
  uint32_t* s_row = reinterpret_cast<uint32_t*>(s_dyn + seg_cap);     // [ent_cap]
  const uint32_t rb64 = static_cast<uint32_t>(pitch * (kIbm ? 2 : 4) / 64);
  uint32_t* s_q = reinterpret_cast<uint32_t*>(s_row + ent_cap);       // [ent_cap] (IBM only)
  const int4 hd = __ldg(grp_hdr + blockIdx.x);                        // x: first segment, y: segments, z: first entry, w: entries
  for (int i = threadIdx.x; i < hd.y; i += blockDim.x) s_seg[i] = __ldg(seg + hd.x + i);
  for (int i = threadIdx.x; i < hd.w; i += blockDim.x) {
    s_row[i] = __ldg(ge_row + hd.z + i) * rb64;                        // row offset in 64-byte units (rows are multiples of 64 bytes; < 2^32 for any HBM size)
    if (kIbm) s_q[i] = __ldg(ge_q + hd.z + i);
  }
  __syncthreads();
  const int s = kVec * (blockIdx.y * blockDim.x + threadIdx.x);
  if (s >= n_songs) return;
  const char* base = kIbm ? reinterpret_cast<const char*>(g16 + s) : reinterpret_cast<const char*>(gq32 + s);
  for (int gi = 0; gi < hd.y; ++gi) {
    const int4 sg = s_seg[gi];                                        // x: Sint row, y..z: staged entries, w: accumulate into a pre-zeroed row
    unsigned long long acc[kVec];
#pragma unroll
    for (int t = 0; t < kVec; ++t) acc[t] = 0;
    auto add_row = [&](const RowWords<W>& c, uint32_t q) {
#pragma unroll
      for (int t = 0; t < W; ++t) {
        if (!kIbm) acc[t] += c.w[t];
        else { acc[2 * t] += static_cast<unsigned long long>(c.w[t] & 0xffffu) * q; acc[2 * t + 1] += static_cast<unsigned long long>(c.w[t] >> 16) * q; }
      }
    };
    int i = sg.y;
    for (; i + kRows <= sg.z; i += kRows) {
      RowWords<W> c[kRows];
#pragma unroll
      for (int t = 0; t < kRows; ++t) c[t] = ld_row_words<W>(base + static_cast<unsigned long long>(s_row[i + t]) * 64u);
#pragma unroll
      for (int t = 0; t < kRows; ++t) add_row(c[t], kIbm ? s_q[i + t] : 0u);
    }
    for (; i < sg.z; ++i) add_row(ld_row_words<W>(base + static_cast<unsigned long long>(s_row[i]) * 64u), kIbm ? s_q[i] : 0u);
    unsigned long long* o = reinterpret_cast<unsigned long long*>(sint) + static_cast<long long>(sg.x) * spitch + s;
    if (sg.w) {
#pragma unroll
      for (int t = 0; t < kVec; ++t) if (acc[t]) atomicAdd(o + t, acc[t]);
    } else if (kVec == 1) {
      __stcs(o, acc[0]);
    } else {
#pragma unroll
      for (int t = 0; t < kVec; t += 2) __stcs(reinterpret_cast<ulonglong2*>(o + t), make_ulonglong2(acc[t], acc[t + 1]));
    }
  }
}

int launch_head_rowsum(int model, int words, int threads, const int4* grp_hdr, int n_groups, const int4* seg, const int* ge_row,
                       const uint32_t* ge_q, int seg_cap, int ent_cap, const uint16_t* g16, const uint32_t* gq32, long long pitch,
                       int n_songs, long long* sint, long long spitch, cudaStream_t st) {
  if (n_groups <= 0 || n_songs <= 0) return 0;
  if (threads < 32 || threads > 256 || threads % 32) return -2;
  const int tile = threads * (model == 2 ? 2 * words : words);   // songs per CTA
  const dim3 grid(n_groups, (n_songs + tile - 1) / tile);
  const size_t smem = static_cast<size_t>(seg_cap) * sizeof(int4) + static_cast<size_t>(ent_cap) * 8;
  if (smem > 48 * 1024) return -3;
#define MR_HR(IBM, W) head_rowsum_kernel<IBM, W><<<grid, threads, smem, st>>>(grp_hdr, seg, ge_row, ge_q, seg_cap, ent_cap, g16, gq32, pitch, n_songs, sint, spitch)
  if (model == 1) { if (words == 4) MR_HR(false, 4); else if (words == 2) MR_HR(false, 2); else if (words == 1) MR_HR(false, 1); else return -2; }
  else if (model == 2) { if (words == 4) MR_HR(true, 4); else if (words == 2) MR_HR(true, 2); else if (words == 1) MR_HR(true, 1); else return -2; }
  else return -2;
#undef MR_HR
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// rows of users whose head entries were split over several segments start from zero (the segments accumulate atomically)
__global__ void __launch_bounds__(256)
zero_rows_kernel(const int* __restrict__ rows, long long* __restrict__ sint, long long spitch) {
  ulonglong2* o = reinterpret_cast<ulonglong2*>(sint + static_cast<long long>(rows[blockIdx.x]) * spitch);
  for (long long i = blockIdx.y * blockDim.x + threadIdx.x; i < spitch / 2; i += static_cast<long long>(gridDim.y) * blockDim.x) o[i] = make_ulonglong2(0, 0);
}

int launch_zero_rows(const int* rows, int n_rows, long long* sint, long long spitch, cudaStream_t st) {
  if (n_rows <= 0) return 0;
  zero_rows_kernel<<<dim3(n_rows, 64), 256, 0, st>>>(rows, sint, spitch);
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
// kLanes: lanes that share one pair.  32 for whole rows; with a song window only the in-window prefix of I_v is walked (a few songs
// of ~47 when the window is 1/8 of the songs), so a pair gets 8 or 16 lanes and a warp works on 4 or 2 pairs at a time.
// ---------------------------------------------------------------------------------------------------------------------
template <int kModels, int kLanes>
__global__ void __launch_bounds__(256)
tail_scatter_kernel(const int* __restrict__ tu_user, const int* __restrict__ tu_song, const long long* __restrict__ tu_lptr,
                    long long e0, long long e1, const long long* __restrict__ csc_ptr, const int* __restrict__ csc_idx,
                    const long long* __restrict__ tr_ptr, const long long* __restrict__ tr_end, const int* __restrict__ tr_col,
                    const uint32_t* __restrict__ qv, const uint32_t* __restrict__ qd, int u0, long long* __restrict__ sint_u,
                    long long* __restrict__ sint_i, long long spitch) {
  const int lane = threadIdx.x & (kLanes - 1);
  const long long w0 = tu_lptr[e0], w1 = tu_lptr[e1];
  const long long n_warps = static_cast<long long>(gridDim.x) * (blockDim.x / kLanes);
  for (long long w = w0 + static_cast<long long>(blockIdx.x) * (blockDim.x / kLanes) + (threadIdx.x / kLanes); w < w1; w += n_warps) {
    long long lo = e0, hi = e1;                      // e = upper_bound(tu_lptr, w) - 1 within [e0, e1)
    while (lo < hi) { const long long m = (lo + hi) >> 1; if (tu_lptr[m + 1] <= w) lo = m + 1; else hi = m; }
    const long long e = lo;
    const int b = tu_user[e] - u0;
    const int j = tu_song[e];
    const int v = csc_idx[csc_ptr[j] + (w - tu_lptr[e])];
    const unsigned long long q = qv[v], qj = qd[j];
    unsigned long long* su = reinterpret_cast<unsigned long long*>(sint_u) + static_cast<long long>(b) * spitch;
    unsigned long long* si = reinterpret_cast<unsigned long long*>(sint_i) + static_cast<long long>(b) * spitch;
    const long long rb = tr_ptr[v], re = tr_end[v];
    for (long long m = rb + lane; m < re; m += kLanes) {
      const int s = __ldg(tr_col + m);
      if (s == j) continue;
      if (kModels & 1) atomicAdd(su + s, q);
      if (kModels & 2) atomicAdd(si + s, qj);
    }
  }
}

int launch_tail_scatter(int models, const int* tu_user, const int* tu_song, const long long* tu_lptr, long long e0, long long e1,
                        const long long* csc_ptr, const int* csc_idx, const long long* tr_ptr, const long long* tr_end, const int* tr_col,
                        const uint32_t* qv, const uint32_t* qd, int u0, long long* sint_u, long long* sint_i, long long spitch, long long n_pairs,
                        int lanes, cudaStream_t st) {
  if (e1 <= e0 || n_pairs <= 0) return 0;
  if (lanes != 8 && lanes != 16 && lanes != 32) return -2;
  // one warp (or 8 / 16 lanes of one) per (entry, listener) pair in short-lived CTAs (measured 25 % faster than a persistent
  // grid-stride grid: the block scheduler balances the 1..4000-song listeners better than a static stride does)
  const int per_cta = 256 / lanes;
  const int grid = static_cast<int>(std::min<long long>((n_pairs + per_cta - 1) / per_cta, 1LL << 30));
#define MR_TS(M, L) tail_scatter_kernel<M, L><<<grid, 256, 0, st>>>(tu_user, tu_song, tu_lptr, e0, e1, csc_ptr, csc_idx, tr_ptr, tr_end, tr_col, qv, qd, u0, sint_u, sint_i, spitch)
#define MR_TS_L(M) do { if (lanes == 32) MR_TS(M, 32); else if (lanes == 16) MR_TS(M, 16); else MR_TS(M, 8); } while (0)
  if (models == 1) MR_TS_L(1); else if (models == 2) MR_TS_L(2); else MR_TS_L(3);
#undef MR_TS_L
#undef MR_TS
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace mr
