// k7_testlists.cu — the per-shard work lists of the item-space engine, built on the device.
//
// mr_set_test_users used to walk the shard's test entries on the host (one random table lookup per entry, a few ms per 350 k
// entries); that walk sat inside every end-to-end step.  Here the entries are classified, compacted and tagged by three small kernels
// and two scans, and the host only sees the per-user prefixes it needs for planning:
//   head entries  (song has a precomputed row)  -> hu_row (row index), hu_song, hu_q (q_26(d_j), MusicRecommender.scala:237) and the
//                                                   per-user prefix hu_ptr
//   tail entries  (song expanded on the fly)    -> tu_user, tu_song, the exclusive prefix tu_lptr of the entries' train listener counts
//                                                   (the flattened (entry, listener) pair index space of tail_scatter_kernel) and tu_ptr
// Entry order is preserved (user-major, songs ascending), exactly what the host loop produced.
#include "mr_common.cuh"
#include "mr_kernels.h"

#include <cub/cub.cuh>

#include <algorithm>

namespace mr {

// flag[e] = 1 for a head entry, deg[e] = train listeners of a tail entry's song (0 for head entries); both arrays have nnz + 1 items,
// the last one 0, so that the exclusive scans deliver the totals at index nnz
__global__ void __launch_bounds__(256)
classify_entries_kernel(const int* __restrict__ te_col, long long nnz, const int2* __restrict__ song_info, int* __restrict__ flag,
                        long long* __restrict__ deg) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e > nnz) return;
  if (e == nnz) { flag[e] = 0; deg[e] = 0; return; }
  const int2 si = __ldg(song_info + te_col[e]);
  flag[e] = si.x >= 0 ? 1 : 0;
  deg[e] = si.x >= 0 ? 0 : static_cast<long long>(static_cast<uint32_t>(si.y));
}

// one warp per test user: scatter its entries to their compacted positions, publish the user's prefixes
__global__ void __launch_bounds__(256)
scatter_entries_kernel(const long long* __restrict__ te_ptr, const int* __restrict__ te_col, int n_users, long long nnz,
                       const int2* __restrict__ song_info, const int* __restrict__ head_pos, const long long* __restrict__ lsum,
                       int* __restrict__ hu_row, int* __restrict__ hu_song, uint32_t* __restrict__ hu_q, long long* __restrict__ hu_ptr,
                       int* __restrict__ tu_user, int* __restrict__ tu_song, long long* __restrict__ tu_lptr, long long* __restrict__ tu_ptr) {
  const int lane = threadIdx.x & 31;
  const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u > n_users) return;
  if (u == n_users) {   // totals
    if (lane == 0) {
      const long long n_head = head_pos[nnz];
      hu_ptr[u] = n_head; tu_ptr[u] = nnz - n_head; tu_lptr[nnz - n_head] = lsum[nnz];
    }
    return;
  }
  const long long b = te_ptr[u], e1 = te_ptr[u + 1];
  if (lane == 0) { const long long hp = head_pos[b]; hu_ptr[u] = hp; tu_ptr[u] = b - hp; }
  for (long long e = b + lane; e < e1; e += 32) {
    const int j = te_col[e];
    const int2 si = __ldg(song_info + j);
    const int hp = head_pos[e];
    if (si.x >= 0) { hu_row[hp] = si.x; hu_song[hp] = j; hu_q[hp] = static_cast<uint32_t>(si.y); }
    else { const long long tp = e - hp; tu_user[tp] = u; tu_song[tp] = j; tu_lptr[tp] = lsum[e]; }
  }
}

size_t test_lists_temp_bytes(long long nnz) {
  size_t a = 0, b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, a, static_cast<const int*>(nullptr), static_cast<int*>(nullptr), static_cast<int>(nnz + 1));
  cub::DeviceScan::ExclusiveSum(nullptr, b, static_cast<const long long*>(nullptr), static_cast<long long*>(nullptr), static_cast<int>(nnz + 1));
  return std::max(a, b);
}

int launch_build_test_lists(const long long* te_ptr, const int* te_col, int n_users, long long nnz, const int2* song_info, int* flag,
                            int* head_pos, long long* deg, long long* lsum, void* cub_tmp, size_t cub_tmp_bytes, int* hu_row, int* hu_song,
                            uint32_t* hu_q, long long* hu_ptr, int* tu_user, int* tu_song, long long* tu_lptr, long long* tu_ptr,
                            cudaStream_t st) {
  if (nnz + 1 >= (1LL << 31)) return -2;
  const int n1 = static_cast<int>(nnz + 1);
  classify_entries_kernel<<<(n1 + 255) / 256, 256, 0, st>>>(te_col, nnz, song_info, flag, deg);
  size_t need = cub_tmp_bytes;
  if (cub::DeviceScan::ExclusiveSum(cub_tmp, need, flag, head_pos, n1, st) != cudaSuccess) return -1;
  need = cub_tmp_bytes;
  if (cub::DeviceScan::ExclusiveSum(cub_tmp, need, deg, lsum, n1, st) != cudaSuccess) return -1;
  const int warps = n_users + 1;
  scatter_entries_kernel<<<(warps + 7) / 8, 256, 0, st>>>(te_ptr, te_col, n_users, nnz, song_info, head_pos, lsum, hu_row, hu_song, hu_q, hu_ptr,
                                                         tu_user, tu_song, tu_lptr, tu_ptr);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// group-ordered copies of the head entries for head_rowsum_kernel's shared-memory staging: desc = (first head entry, absolute
// destination, count, unused); one warp per segment
__global__ void __launch_bounds__(256)
gather_group_entries_kernel(const int4* __restrict__ desc, int n_desc, const int* __restrict__ hu_row, const uint32_t* __restrict__ hu_q,
                            int* __restrict__ ge_row, uint32_t* __restrict__ ge_q) {
  const int lane = threadIdx.x & 31;
  const int d = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (d >= n_desc) return;
  const int4 c = __ldg(desc + d);
  for (int i = lane; i < c.z; i += 32) { ge_row[c.y + i] = hu_row[c.x + i]; ge_q[c.y + i] = hu_q[c.x + i]; }
}

int launch_gather_group_entries(const int4* desc, int n_desc, const int* hu_row, const uint32_t* hu_q, int* ge_row, uint32_t* ge_q,
                                cudaStream_t st) {
  if (n_desc <= 0) return 0;
  gather_group_entries_kernel<<<(n_desc + 7) / 8, 256, 0, st>>>(desc, n_desc, hu_row, hu_q, ge_row, ge_q);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace mr
