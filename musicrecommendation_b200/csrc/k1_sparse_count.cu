// k1_sparse_count.cu — kernel K1s: the same intersection counts as K1, computed from the inverted index.
//
// Used for shapes whose dense 0/1 operands do not fit in HBM (MSD scale: T x S bytes = 385 GB) and for the long popularity
// tail where a dense contraction would multiply zeros.  Produces bit-identical counts to the tensor-core kernel (integers).
//   UBM  Ct[v][b] += 1 for every (b, j in I_u(b), v in U_j^train)              (MusicRecommender.scala:142-145)
//   IBM  G[r][s]  += 1 for every (r, v in U_{rows[r]}^train, s in I_v)          (MusicRecommender.scala:232-235)
#include "mr_common.cuh"
#include "mr_kernels.h"

namespace mr {

// One warp per (batch user b, visible song j): lanes stride over the listeners of j (coalesced index reads) and bump the
// packed u16 counter of (v, b) with a 32-bit atomic on the containing word.  A counter cannot carry into its neighbour
// because a count is bounded by |I_u| <= 65535 (checked at load time).
__global__ void __launch_bounds__(256)
sparse_count_u16t_kernel(const long long* __restrict__ te_ptr, const int* __restrict__ te_col, int u0, int n_users,
                         const long long* __restrict__ csc_ptr, const int* __restrict__ csc_idx, unsigned int* __restrict__ ct32,
                         long long n_train) {
  const int lane = threadIdx.x & 31;
  const long long e0 = te_ptr[u0], e1 = te_ptr[u0 + n_users];
  const long long n_warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long e = e0 + static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); e < e1; e += n_warps) {
    // batch-local user of entry e: upper_bound over the <= 129 row pointers of the batch
    int lo = 0, hi = n_users;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (te_ptr[u0 + m + 1] <= e) lo = m + 1; else hi = m; }
    const int b = lo;
    const int j = te_col[e];
    const long long beg = csc_ptr[j], end = csc_ptr[j + 1];
    const unsigned int inc = 1u << (16 * (b & 1));
    for (long long i = beg + lane; i < end; i += 32) {
      const int v = __ldg(csc_idx + i);
      atomicAdd(ct32 + (ct_index(n_train, v, b) >> 1), inc);
    }
  }
}

int launch_sparse_count_u16t(const long long* te_ptr, const int* te_col, int u0, int n_users, const long long* csc_ptr,
                             const int* csc_idx, uint16_t* ct, long long n_train, cudaStream_t st) {
  if (n_users <= 0) return 0;
  cudaError_t e = cudaMemsetAsync(ct, 0, static_cast<size_t>(n_train) * kUserBatch * sizeof(uint16_t), st);
  if (e != cudaSuccess) return -1;
  sparse_count_u16t_kernel<<<148 * 8, 256, 0, st>>>(te_ptr, te_col, u0, n_users, csc_ptr, csc_idx,
                                                     reinterpret_cast<unsigned int*>(ct), n_train);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// IBM in "user space" (SURVEY.md A.3 restructuring): Wi[v][b] = sum_{j in I_u(b) ∩ I_v} qd[j].  Then
// Sint_i[b][s] = sum_{v in U_s^train} Wi[v][b] equals sum_{j in I_u} |U_s ∩ U_j| * qd[j] for every pair with s not in I_u
// (for those no j equals s, so the reference's `s2 != song` filter MR:252 removes nothing); listened pairs are masked.
// The panel holds u32 entries; qd <= 2^26 so an entry only wraps past ~64 shared songs.  A wrap is detected from the value
// the atomic returns and recorded as a carry event (v, b); launch_carry_fixup adds the missing 2^32 exactly.
__global__ void __launch_bounds__(256)
sparse_wcount_u32_kernel(const long long* __restrict__ te_ptr, const int* __restrict__ te_col, int u0, int n_users,
                         const long long* __restrict__ csc_ptr, const int* __restrict__ csc_idx, const uint32_t* __restrict__ qd,
                         uint32_t* __restrict__ wi, long long n_train, CarryList carry) {
  const int lane = threadIdx.x & 31;
  const long long e0 = te_ptr[u0], e1 = te_ptr[u0 + n_users];
  const long long n_warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long e = e0 + static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); e < e1; e += n_warps) {
    int lo = 0, hi = n_users;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (te_ptr[u0 + m + 1] <= e) lo = m + 1; else hi = m; }
    const int b = lo;
    const int j = te_col[e];
    const uint32_t q = qd[j];
    const long long beg = csc_ptr[j], end = csc_ptr[j + 1];
    for (long long i = beg + lane; i < end; i += 32) {
      const int v = __ldg(csc_idx + i);
      const uint32_t old = atomicAdd(wi + wi_index(n_train, v, b), q);
      if (old > 0xffffffffu - q) {                       // wrapped: remember that (v, b) is short by 2^32
        const unsigned int pos = atomicAdd(carry.count, 1u);
        if (pos < carry.capacity) carry.events[pos] = make_uint2(static_cast<unsigned int>(v), static_cast<unsigned int>(b));
      }
    }
  }
}

int launch_sparse_wcount_u32(const long long* te_ptr, const int* te_col, int u0, int n_users, const long long* csc_ptr,
                             const int* csc_idx, const uint32_t* qd, uint32_t* wi, long long n_train, CarryList carry, cudaStream_t st) {
  if (n_users <= 0) return 0;
  cudaError_t e = cudaMemsetAsync(wi, 0, static_cast<size_t>(n_train) * kUserBatch * sizeof(uint32_t), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(carry.count, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return -1;
  sparse_wcount_u32_kernel<<<148 * 8, 256, 0, st>>>(te_ptr, te_col, u0, n_users, csc_ptr, csc_idx, qd, wi, n_train, carry);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Every carry event (v, b) means Wi[v][b] is 2^32 larger than the panel says: add 2^32 to Sint[b][s] for each s in I_v.
__global__ void carry_fixup_kernel(CarryList carry, const long long* __restrict__ tr_ptr, const int* __restrict__ tr_col,
                                   long long* __restrict__ sint, long long spitch) {
  const unsigned int n = min(*carry.count, carry.capacity);
  for (unsigned int i = blockIdx.x; i < n; i += gridDim.x) {
    const uint2 ev = carry.events[i];
    const long long b = tr_ptr[ev.x], e = tr_ptr[ev.x + 1];
    for (long long m = b + threadIdx.x; m < e; m += blockDim.x)
      atomicAdd(reinterpret_cast<unsigned long long*>(sint) + static_cast<long long>(ev.y) * spitch + tr_col[m], 1ULL << 32);
  }
}

int launch_carry_fixup(CarryList carry, const long long* tr_ptr, const int* tr_col, long long* sint, long long spitch, cudaStream_t st) {
  carry_fixup_kernel<<<148, 128, 0, st>>>(carry, tr_ptr, tr_col, sint, spitch);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// One warp per (Gram row r, listener v of song rows[r]): lanes stride over I_v.
__global__ void __launch_bounds__(256)
sparse_gram_rows_kernel(const int* __restrict__ rows, int n_rows, const long long* __restrict__ csc_ptr,
                        const int* __restrict__ csc_idx, const long long* __restrict__ tr_ptr, const int* __restrict__ tr_col,
                        int32_t* __restrict__ g, long long ldg) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.y;
  if (r >= n_rows) return;
  const int song = rows[r];
  const long long beg = csc_ptr[song], end = csc_ptr[song + 1];
  const long long n_warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long i = beg + static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); i < end; i += n_warps) {
    const int v = csc_idx[i];
    const long long b = tr_ptr[v], e = tr_ptr[v + 1];
    for (long long m = b + lane; m < e; m += 32) atomicAdd(g + static_cast<long long>(r) * ldg + __ldg(tr_col + m), 1);
  }
}

int launch_sparse_gram_rows(const int* rows, int n_rows, const long long* csc_ptr, const int* csc_idx, const long long* tr_ptr,
                            const int* tr_col, int32_t* g, long long ldg, cudaStream_t st) {
  if (n_rows <= 0) return 0;
  cudaError_t e = cudaMemsetAsync(g, 0, static_cast<size_t>(n_rows) * ldg * sizeof(int32_t), st);
  if (e != cudaSuccess) return -1;
  sparse_gram_rows_kernel<<<dim3(4, n_rows), 256, 0, st>>>(rows, n_rows, csc_ptr, csc_idx, tr_ptr, tr_col, g, ldg);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace mr
