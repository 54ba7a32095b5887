// k6_ingest.cu — TSV triplets -> int-id CSR on the GPU: the data model `new MusicRecommender(train, test, labels)` builds
// (MusicRecommender.scala:26-91; twin distributed.scala:91-125), SURVEY.md §8f N3.
//
//   line      `user \t song \t count`, third field ignored (MR:35); a line without exactly 3 fields after Java's String.split dropped
//             the trailing empty ones is a scala.MatchError (MR:34-35) -> MR_ERR_BAD_ARG naming the line
//   users     train users and test users are separate id spaces; songs = every song of the train and the test file (MR:38, 51, 58);
//             ids are assigned in ascending String.compareTo order (== byte order for the ASCII ids of the data set), so that int order
//             is the order of the alignment sort main.scala:57-59
//   degrees   the `.length` of the reference's NON-deduplicated per-user / per-song lists (MR:40-41, 147, 237): duplicate rows inflate
//             them, while the CSR rows the numerators iterate are sorted and unique
//   labels    rows of users that are not test users are dropped; label songs that occur nowhere else get ids >= n_songs
//
// Pipeline (one device buffer holds the three files back to back):
//   1  newline positions            cub::DeviceSelect::If over a counting iterator
//   2  parse_lines_kernel           thread per line: field boundaries + a 64-bit hash of the user and the song field
//   3  per id space: radix sort (hash, element), run heads -> dense "unique" index, exact string comparison of every element with
//      its run's representative (a 64-bit collision is reported, never silently merged), representatives' strings gathered
//   4  host: the ~1.4 M unique strings are sorted (std::sort, memcmp) -> rank per unique; device: element -> id
//   5  per matrix: keys (row << 32 | col) radix-sorted, duplicates dropped by an adjacent compare + scan, degrees by atomics
// CUB's device-wide sort / scan / select are library primitives; everything specific to the format is hand-written here.  Ingest is
// not on the scoring hot path (it runs once per data set).
#include "../../include/mrscore.h"

#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

struct mr_ingest {
  std::string err;
  int32_t n_train = 0, n_test = 0, n_songs = 0, n_label_only = 0;
  std::vector<int64_t> ptr[3];        // train, test, labels
  std::vector<int32_t> col[3];
  std::vector<int32_t> deg_train, deg_test, deg_song;
  std::vector<char> chars[3];         // id tables: train users, test users, songs (+ label-only songs)
  std::vector<int64_t> off[3];
  double timing[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // ms: h2d, lines+parse, unique users, unique songs, host sort, csr, d2h, total
};

namespace {

int ing_fail(mr_ingest* g, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (g) g->err = buf;
  return code;
}

#define ING_CUDA(expr)                                                                                              \
  do {                                                                                                              \
    cudaError_t e__ = (expr);                                                                                       \
    if (e__ != cudaSuccess)                                                                                         \
      return ing_fail(g, e__ == cudaErrorMemoryAllocation ? MR_ERR_OOM : MR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, \
                      cudaGetErrorString(e__), __FILE__, __LINE__);                                                 \
  } while (0)

template <class T>
struct DVec {   // device array freed on scope exit
  T* p = nullptr; size_t n = 0;
  DVec() = default;
  DVec(const DVec&) = delete;
  DVec& operator=(const DVec&) = delete;
  ~DVec() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t count) {
    if (p) { cudaFree(p); p = nullptr; }
    n = count;
    return cudaMalloc(reinterpret_cast<void**>(&p), std::max<size_t>(count, 1) * sizeof(T));
  }
};

struct Temp {   // grow-only CUB temporary storage
  void* p = nullptr; size_t bytes = 0;
  ~Temp() { if (p) cudaFree(p); }
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, need);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
};

__device__ __forceinline__ uint64_t hash_bytes(const unsigned char* p, int len) {
  uint64_t h = 0xcbf29ce484222325ULL;                       // FNV-1a, then a splitmix64 finaliser
  for (int i = 0; i < len; ++i) { h ^= p[i]; h *= 0x100000001b3ULL; }
  h ^= static_cast<uint64_t>(len) << 56;
  h ^= h >> 30; h *= 0xbf58476d1ce4e5b9ULL; h ^= h >> 27; h *= 0x94d049bb133111ebULL; h ^= h >> 31;
  return h;
}

__global__ void count_newlines_kernel(const char* __restrict__ buf, long long begin, long long end, unsigned long long* __restrict__ out) {
  unsigned long long c = 0;
  for (long long i = begin + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < end; i += static_cast<long long>(gridDim.x) * blockDim.x)
    c += buf[i] == '\n';
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

struct IsNewline {
  const char* buf;
  __device__ bool operator()(long long i) const { return buf[i] == '\n'; }
};

// one thread per line of one file: [begin, end) of the line from the newline positions, trailing '\r' and trailing tabs dropped
// (Source.getLines / String.split), exactly two tabs must remain
__global__ void parse_lines_kernel(const char* __restrict__ buf, long long file_begin, long long file_end, const long long* __restrict__ nl,
                                   long long n_nl, long long n_lines, long long line0, long long* __restrict__ u_off, int* __restrict__ u_len,
                                   long long* __restrict__ s_off, int* __restrict__ s_len, uint64_t* __restrict__ u_hash,
                                   uint64_t* __restrict__ s_hash, unsigned long long* __restrict__ first_bad) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_lines) return;
  const long long b = i == 0 ? file_begin : nl[i - 1] + 1;
  long long e = i < n_nl ? nl[i] : file_end;
  if (e > b && buf[e - 1] == '\r') --e;
  while (e > b && buf[e - 1] == '\t') --e;
  long long t1 = -1, t2 = -1; int tabs = 0;
  for (long long p = b; p < e; ++p)
    if (buf[p] == '\t') { if (tabs == 0) t1 = p; else if (tabs == 1) t2 = p; ++tabs; }
  const long long g = line0 + i;
  if (tabs != 2) {
    atomicMin(first_bad, static_cast<unsigned long long>(g));
    u_off[g] = b; u_len[g] = 0; s_off[g] = b; s_len[g] = 0; u_hash[g] = 0; s_hash[g] = 0;
    return;
  }
  u_off[g] = b; u_len[g] = static_cast<int>(t1 - b);
  s_off[g] = t1 + 1; s_len[g] = static_cast<int>(t2 - t1 - 1);
  u_hash[g] = hash_bytes(reinterpret_cast<const unsigned char*>(buf + b), static_cast<int>(t1 - b));
  s_hash[g] = hash_bytes(reinterpret_cast<const unsigned char*>(buf + t1 + 1), static_cast<int>(t2 - t1 - 1));
}

__global__ void line_flags_kernel(int* flag, long long n, long long first_label_line) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = i < first_label_line ? 1 : 2;
}

// first 8 bytes of a unique's string as a big-endian integer, zero padded: ordering by it never contradicts the byte order of the full
// strings (ties are resolved on the host), so the device sorts the uniques and the host only orders the runs of equal prefixes
__global__ void prefix_keys_kernel(const char* __restrict__ blob, const long long* __restrict__ str_off, int n_uniq, uint64_t* __restrict__ key) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_uniq) return;
  const long long b = str_off[u], len = str_off[u + 1] - b;
  uint64_t k = 0;
  for (int i = 0; i < 8; ++i) k = (k << 8) | (i < len ? static_cast<unsigned char>(blob[b + i]) : 0u);
  key[u] = k;
}

__global__ void iota_u32_kernel(uint32_t* p, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = static_cast<uint32_t>(i);
}

__global__ void run_heads_kernel(const uint64_t* __restrict__ key, long long n, int* __restrict__ head) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}

// sorted position i -> its element gets the run's unique index; run heads publish their element as the representative
__global__ void assign_unique_kernel(const uint32_t* __restrict__ idx, const int* __restrict__ head, const int* __restrict__ incl, long long n,
                                     const int* __restrict__ flag, int* __restrict__ elem_uniq, int* __restrict__ rep_elem,
                                     int* __restrict__ uniq_flag) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int u = incl[i] - 1;
  const uint32_t e = idx[i];
  elem_uniq[e] = u;
  if (head[i]) rep_elem[u] = static_cast<int>(e);
  if (flag[e]) atomicOr(uniq_flag + u, flag[e]);
}

__global__ void check_collisions_kernel(const char* __restrict__ buf, const long long* __restrict__ off, const int* __restrict__ len,
                                        const int* __restrict__ elem_uniq, const int* __restrict__ rep_elem, long long n,
                                        int* __restrict__ collision) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int r = rep_elem[elem_uniq[e]];
  if (r == e) return;
  bool same = len[e] == len[r];
  for (int k = 0; same && k < len[e]; ++k) same = buf[off[e] + k] == buf[off[r] + k];
  if (!same) *collision = 1;
}

__global__ void rep_lengths_kernel(const int* __restrict__ rep_elem, const int* __restrict__ len, int n_uniq, long long* __restrict__ out) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < n_uniq) out[u] = len[rep_elem[u]];
}

__global__ void gather_strings_kernel(const char* __restrict__ buf, const long long* __restrict__ off, const int* __restrict__ len,
                                      const int* __restrict__ rep_elem, const long long* __restrict__ str_off, int n_uniq,
                                      char* __restrict__ blob) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_uniq) return;
  const int e = rep_elem[u];
  const char* src = buf + off[e];
  char* dst = blob + str_off[u];
  for (int k = 0; k < len[e]; ++k) dst[k] = src[k];
}

__global__ void apply_rank_kernel(const int* __restrict__ elem_uniq, const int* __restrict__ rank, long long n, int* __restrict__ id) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e < n) id[e] = rank[elem_uniq[e]];
}

// (row, col) of the elements [e0, e0 + n) -> sort keys; rows < 0 (label rows of unknown users) sort last and are counted
__global__ void make_keys_kernel(const int* __restrict__ row, const int* __restrict__ col, long long n, unsigned long long* __restrict__ key,
                                 unsigned long long* __restrict__ n_dropped) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (row[i] < 0 || col[i] < 0) { key[i] = ~0ULL; atomicAdd(n_dropped, 1ULL); }
  else key[i] = (static_cast<unsigned long long>(static_cast<uint32_t>(row[i])) << 32) | static_cast<uint32_t>(col[i]);
}

__global__ void csr_flags_kernel(const unsigned long long* __restrict__ key, long long n, int* __restrict__ uniq, int* __restrict__ deg,
                                 long long* __restrict__ row_count) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = static_cast<int>(key[i] >> 32);
  const bool first = i == 0 || key[i] != key[i - 1];
  uniq[i] = first ? 1 : 0;
  atomicAdd(deg + r, 1);                                         // `.length` counts duplicate rows (MR:40, 147)
  if (first) atomicAdd(reinterpret_cast<unsigned long long*>(row_count + r), 1ULL);
}

__global__ void csr_cols_kernel(const unsigned long long* __restrict__ key, const int* __restrict__ uniq, const int* __restrict__ pos, long long n,
                                int* __restrict__ col) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n && uniq[i]) col[pos[i]] = static_cast<int>(key[i] & 0xffffffffULL);
}

__global__ void bincount_kernel(const int* __restrict__ id, long long n, int n_bins, int* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n && id[i] >= 0 && id[i] < n_bins) atomicAdd(out + id[i], 1);
}

inline unsigned int blocks_for(long long n, int threads = 256) { return static_cast<unsigned int>(std::max<long long>(1, (n + threads - 1) / threads)); }

struct UniqueSet {            // result of step 3 for one id space
  int n_uniq = 0;
  DVec<int> elem_uniq;        // [n] unique index (hash order) of every element
  std::vector<char> blob;     // representatives' strings, concatenated in unique-index order
  std::vector<long long> str_off;   // [n_uniq + 1]
  std::vector<int> flag;      // [n_uniq] OR of the elements' flags
  std::vector<uint32_t> order;      // [n_uniq] unique indices sorted by the 8-byte string prefix (device radix sort)
  std::vector<uint64_t> order_key;  // [n_uniq] the prefixes in that order
};

// elements = fields [off[e], off[e] + len[e]) of buf with hash[e], e in [0, n)
int unique_fields(mr_ingest* g, const char* d_buf, const uint64_t* d_hash, const long long* d_off, const int* d_len, const int* d_flag, long long n,
                  Temp& tmp, UniqueSet& out) {
  out.n_uniq = 0;
  ING_CUDA(out.elem_uniq.alloc(static_cast<size_t>(n)));
  out.blob.clear(); out.str_off.assign(1, 0); out.flag.clear(); out.order.clear(); out.order_key.clear();
  if (n == 0) return MR_OK;
  DVec<uint64_t> key_sorted; DVec<uint32_t> idx, idx_sorted; DVec<int> head, incl, rep_elem, uniq_flag, collision;
  ING_CUDA(key_sorted.alloc(n)); ING_CUDA(idx.alloc(n)); ING_CUDA(idx_sorted.alloc(n)); ING_CUDA(head.alloc(n)); ING_CUDA(incl.alloc(n));
  iota_u32_kernel<<<blocks_for(n), 256>>>(idx.p, n);
  size_t need = 0;
  ING_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, d_hash, key_sorted.p, idx.p, idx_sorted.p, static_cast<int>(n)));
  ING_CUDA(tmp.reserve(need));
  ING_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, need, d_hash, key_sorted.p, idx.p, idx_sorted.p, static_cast<int>(n)));
  run_heads_kernel<<<blocks_for(n), 256>>>(key_sorted.p, n, head.p);
  ING_CUDA(cub::DeviceScan::InclusiveSum(nullptr, need, head.p, incl.p, static_cast<int>(n)));
  ING_CUDA(tmp.reserve(need));
  ING_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, need, head.p, incl.p, static_cast<int>(n)));
  int n_uniq = 0;
  ING_CUDA(cudaMemcpy(&n_uniq, incl.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost));
  out.n_uniq = n_uniq;
  ING_CUDA(rep_elem.alloc(n_uniq)); ING_CUDA(uniq_flag.alloc(n_uniq)); ING_CUDA(collision.alloc(1));
  ING_CUDA(cudaMemset(uniq_flag.p, 0, static_cast<size_t>(n_uniq) * sizeof(int)));
  ING_CUDA(cudaMemset(collision.p, 0, sizeof(int)));
  assign_unique_kernel<<<blocks_for(n), 256>>>(idx_sorted.p, head.p, incl.p, n, d_flag, out.elem_uniq.p, rep_elem.p, uniq_flag.p);
  check_collisions_kernel<<<blocks_for(n), 256>>>(d_buf, d_off, d_len, out.elem_uniq.p, rep_elem.p, n, collision.p);
  int coll = 0;
  ING_CUDA(cudaMemcpy(&coll, collision.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (coll) return ing_fail(g, MR_ERR_STATE, "two different ids share a 64-bit hash; use the host ingest for this data set");
  // representatives' strings -> host
  DVec<long long> rep_len, str_off; DVec<char> blob;
  ING_CUDA(rep_len.alloc(static_cast<size_t>(n_uniq) + 1)); ING_CUDA(str_off.alloc(static_cast<size_t>(n_uniq) + 1));
  ING_CUDA(cudaMemset(rep_len.p, 0, (static_cast<size_t>(n_uniq) + 1) * sizeof(long long)));
  rep_lengths_kernel<<<blocks_for(n_uniq), 256>>>(rep_elem.p, d_len, n_uniq, rep_len.p);
  ING_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, rep_len.p, str_off.p, n_uniq + 1));
  ING_CUDA(tmp.reserve(need));
  ING_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, need, rep_len.p, str_off.p, n_uniq + 1));
  out.str_off.resize(static_cast<size_t>(n_uniq) + 1);
  ING_CUDA(cudaMemcpy(out.str_off.data(), str_off.p, out.str_off.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  const long long total = out.str_off.back();
  ING_CUDA(blob.alloc(static_cast<size_t>(total)));
  gather_strings_kernel<<<blocks_for(n_uniq), 256>>>(d_buf, d_off, d_len, rep_elem.p, str_off.p, n_uniq, blob.p);
  out.blob.resize(static_cast<size_t>(total));
  if (total) ING_CUDA(cudaMemcpy(out.blob.data(), blob.p, static_cast<size_t>(total), cudaMemcpyDeviceToHost));
  out.flag.resize(n_uniq);
  ING_CUDA(cudaMemcpy(out.flag.data(), uniq_flag.p, static_cast<size_t>(n_uniq) * sizeof(int), cudaMemcpyDeviceToHost));
  {  // order of the uniques by string prefix
    DVec<uint64_t> pk, pk_sorted; DVec<uint32_t> ui, ui_sorted;
    ING_CUDA(pk.alloc(n_uniq)); ING_CUDA(pk_sorted.alloc(n_uniq)); ING_CUDA(ui.alloc(n_uniq)); ING_CUDA(ui_sorted.alloc(n_uniq));
    prefix_keys_kernel<<<blocks_for(n_uniq), 256>>>(blob.p, str_off.p, n_uniq, pk.p);
    iota_u32_kernel<<<blocks_for(n_uniq), 256>>>(ui.p, n_uniq);
    ING_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, pk.p, pk_sorted.p, ui.p, ui_sorted.p, n_uniq));
    ING_CUDA(tmp.reserve(need));
    ING_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, need, pk.p, pk_sorted.p, ui.p, ui_sorted.p, n_uniq));
    out.order.resize(n_uniq); out.order_key.resize(n_uniq);
    ING_CUDA(cudaMemcpy(out.order.data(), ui_sorted.p, static_cast<size_t>(n_uniq) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    ING_CUDA(cudaMemcpy(out.order_key.data(), pk_sorted.p, static_cast<size_t>(n_uniq) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  }
  ING_CUDA(cudaGetLastError());
  return MR_OK;
}

// Rank the uniques: those whose flag has `primary_bit` get 0 .. n_primary-1 in ascending string order (String.compareTo on ASCII ==
// byte order, shorter prefix first); the others n_primary .. in ascending order if `keep_rest`, else -1.  Fills the id table.
int rank_uniques(const UniqueSet& u, int primary_bit, bool keep_rest, std::vector<int>& rank, std::vector<char>& chars, std::vector<int64_t>& off,
                 int* n_primary_out) {
  const int n = u.n_uniq;
  std::vector<uint32_t> sorted(u.order);                    // by 8-byte prefix (device); finish runs of equal prefixes with the full comparison
  auto less = [&](uint32_t a, uint32_t b) {
    const long long la = u.str_off[a + 1] - u.str_off[a], lb = u.str_off[b + 1] - u.str_off[b];
    const int c = memcmp(u.blob.data() + u.str_off[a], u.blob.data() + u.str_off[b], static_cast<size_t>(std::min(la, lb)));
    return c != 0 ? c < 0 : la < lb;
  };
  for (int i = 0; i < n;) {
    int j = i + 1;
    while (j < n && u.order_key[j] == u.order_key[i]) ++j;
    if (j - i > 1) std::sort(sorted.begin() + i, sorted.begin() + j, less);
    i = j;
  }
  // primary uniques first (ids 0 .. n_primary-1), the others after them, each group in string order
  std::vector<uint32_t> order; order.reserve(n);
  for (int pass = 0; pass < 2; ++pass)
    for (int i = 0; i < n; ++i)
      if (((u.flag[sorted[i]] & primary_bit) != 0) == (pass == 0)) order.push_back(sorted[i]);
  rank.assign(n, -1);
  chars.clear(); off.assign(1, 0);
  int n_primary = 0;
  for (int i = 0; i < n; ++i) {
    const int x = static_cast<int>(order[i]);
    const bool primary = (u.flag[x] & primary_bit) != 0;
    if (primary) ++n_primary;
    if (!primary && !keep_rest) continue;
    rank[x] = i;
    chars.insert(chars.end(), u.blob.begin() + u.str_off[x], u.blob.begin() + u.str_off[x + 1]);
    off.push_back(static_cast<int64_t>(chars.size()));
  }
  *n_primary_out = n_primary;
  return MR_OK;
}

int build_csr(mr_ingest* g, const int* d_row, const int* d_col, long long n, int n_rows, Temp& tmp, std::vector<int64_t>& ptr, std::vector<int32_t>& col,
              std::vector<int32_t>* deg_out) {
  ptr.assign(static_cast<size_t>(n_rows) + 1, 0);
  col.clear();
  if (deg_out) deg_out->assign(n_rows, 0);
  if (n == 0 || n_rows == 0) return MR_OK;
  DVec<unsigned long long> key, key_sorted, dropped; DVec<int> uniq, pos, deg, d_colv; DVec<long long> row_count, row_ptr;
  ING_CUDA(key.alloc(n)); ING_CUDA(key_sorted.alloc(n)); ING_CUDA(dropped.alloc(1));
  ING_CUDA(cudaMemset(dropped.p, 0, sizeof(unsigned long long)));
  make_keys_kernel<<<blocks_for(n), 256>>>(d_row, d_col, n, key.p, dropped.p);
  size_t need = 0;
  ING_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, need, key.p, key_sorted.p, static_cast<int>(n)));
  ING_CUDA(tmp.reserve(need));
  ING_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, need, key.p, key_sorted.p, static_cast<int>(n)));
  unsigned long long n_drop = 0;
  ING_CUDA(cudaMemcpy(&n_drop, dropped.p, sizeof n_drop, cudaMemcpyDeviceToHost));
  const long long m = n - static_cast<long long>(n_drop);       // dropped keys are ~0 and sort last
  if (m == 0) return MR_OK;
  ING_CUDA(uniq.alloc(m)); ING_CUDA(pos.alloc(m)); ING_CUDA(deg.alloc(n_rows)); ING_CUDA(row_count.alloc(static_cast<size_t>(n_rows) + 1));
  ING_CUDA(row_ptr.alloc(static_cast<size_t>(n_rows) + 1));
  ING_CUDA(cudaMemset(deg.p, 0, static_cast<size_t>(n_rows) * sizeof(int)));
  ING_CUDA(cudaMemset(row_count.p, 0, (static_cast<size_t>(n_rows) + 1) * sizeof(long long)));
  csr_flags_kernel<<<blocks_for(m), 256>>>(key_sorted.p, m, uniq.p, deg.p, row_count.p);
  ING_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, uniq.p, pos.p, static_cast<int>(m)));
  ING_CUDA(tmp.reserve(need));
  ING_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, need, uniq.p, pos.p, static_cast<int>(m)));
  ING_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, row_count.p, row_ptr.p, n_rows + 1));
  ING_CUDA(tmp.reserve(need));
  ING_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, need, row_count.p, row_ptr.p, n_rows + 1));
  ING_CUDA(cudaMemcpy(ptr.data(), row_ptr.p, ptr.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
  const long long nnz = ptr.back();
  ING_CUDA(d_colv.alloc(static_cast<size_t>(nnz)));
  csr_cols_kernel<<<blocks_for(m), 256>>>(key_sorted.p, uniq.p, pos.p, m, d_colv.p);
  col.resize(static_cast<size_t>(nnz));
  if (nnz) ING_CUDA(cudaMemcpy(col.data(), d_colv.p, static_cast<size_t>(nnz) * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (deg_out) ING_CUDA(cudaMemcpy(deg_out->data(), deg.p, static_cast<size_t>(n_rows) * sizeof(int32_t), cudaMemcpyDeviceToHost));
  ING_CUDA(cudaGetLastError());
  return MR_OK;
}

double ms_since(std::chrono::steady_clock::time_point t) {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
}

int ingest_impl(mr_ingest* g, int device, const char* const bufs[3], const uint64_t lens[3]) {
  static const char* const kFileName[3] = {"train", "test", "labels"};
  auto t_total = std::chrono::steady_clock::now();
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0) return ing_fail(g, MR_ERR_CUDA, "no CUDA device: %s — the native ingest has no CPU fallback", cudaGetErrorString(ce));
  if (device < 0 || device >= count) return ing_fail(g, MR_ERR_BAD_ARG, "device id %d out of range (%d devices)", device, count);
  ING_CUDA(cudaSetDevice(device));
  for (int f = 0; f < 3; ++f)
    if (lens[f] && !bufs[f]) return ing_fail(g, MR_ERR_BAD_ARG, "%s buffer is null", kFileName[f]);
  const uint64_t total_bytes = lens[0] + lens[1] + lens[2];
  if (total_bytes >= (1ULL << 40)) return ing_fail(g, MR_ERR_BAD_ARG, "input too large");
  Temp tmp;
  const bool dbg = getenv("MRSCORE_DEBUG_TIMING") != nullptr;
  auto t_lap = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!dbg) return;
    cudaDeviceSynchronize();
    fprintf(stderr, "[mrscore] ingest: %s %.1f ms\n", what, ms_since(t_lap));
    t_lap = std::chrono::steady_clock::now();
  };
  // ---- 0: the three files back to back in one device buffer
  auto t0 = std::chrono::steady_clock::now();
  DVec<char> buf;
  ING_CUDA(buf.alloc(static_cast<size_t>(total_bytes)));
  long long fbeg[4] = {0, static_cast<long long>(lens[0]), static_cast<long long>(lens[0] + lens[1]), static_cast<long long>(total_bytes)};
  for (int f = 0; f < 3; ++f)
    if (lens[f]) ING_CUDA(cudaMemcpy(buf.p + fbeg[f], bufs[f], static_cast<size_t>(lens[f]), cudaMemcpyHostToDevice));
  g->timing[0] = ms_since(t0);
  lap("alloc + h2d");
  // ---- 1: newline positions per file
  t0 = std::chrono::steady_clock::now();
  DVec<long long> nl[3]; DVec<long long> n_sel;
  ING_CUDA(n_sel.alloc(1));
  long long n_nl[3] = {0, 0, 0}, n_lines[3] = {0, 0, 0}, line0[4] = {0, 0, 0, 0};
  DVec<unsigned long long> nl_count;
  ING_CUDA(nl_count.alloc(1));
  for (int f = 0; f < 3; ++f) {
    if (!lens[f]) continue;
    ING_CUDA(cudaMemset(nl_count.p, 0, sizeof(unsigned long long)));
    count_newlines_kernel<<<148 * 8, 256>>>(buf.p, fbeg[f], fbeg[f + 1], nl_count.p);
    unsigned long long total_nl = 0;
    ING_CUDA(cudaMemcpy(&total_nl, nl_count.p, sizeof total_nl, cudaMemcpyDeviceToHost));
    ING_CUDA(nl[f].alloc(static_cast<size_t>(total_nl)));
    const long long piece = 1LL << 30;                    // cub's selection takes an int item count
    long long done = 0;
    for (long long p0 = fbeg[f]; p0 < fbeg[f + 1]; p0 += piece) {
      const int cnt = static_cast<int>(std::min<long long>(piece, fbeg[f + 1] - p0));
      thrust::counting_iterator<long long> it(p0);
      size_t need = 0;
      ING_CUDA(cub::DeviceSelect::If(nullptr, need, it, nl[f].p + done, n_sel.p, cnt, IsNewline{buf.p}));
      ING_CUDA(tmp.reserve(need));
      ING_CUDA(cub::DeviceSelect::If(tmp.p, need, it, nl[f].p + done, n_sel.p, cnt, IsNewline{buf.p}));
      long long c = 0;
      ING_CUDA(cudaMemcpy(&c, n_sel.p, sizeof c, cudaMemcpyDeviceToHost));
      done += c;
    }
    if (done != static_cast<long long>(total_nl)) return ing_fail(g, MR_ERR_STATE, "newline count mismatch (%lld vs %llu)", done, total_nl);
    n_nl[f] = done;
    char last = 0;
    ING_CUDA(cudaMemcpy(&last, buf.p + fbeg[f + 1] - 1, 1, cudaMemcpyDeviceToHost));
    n_lines[f] = done + (last != '\n' ? 1 : 0);
  }
  lap("newline count + select");
  for (int f = 0; f < 3; ++f) line0[f + 1] = line0[f] + n_lines[f];
  const long long L = line0[3];
  if (L >= (1LL << 31)) return ing_fail(g, MR_ERR_BAD_ARG, "more than 2^31 lines");
  // ---- 2: parse
  DVec<long long> u_off, s_off; DVec<int> u_len, s_len; DVec<uint64_t> u_hash, s_hash; DVec<unsigned long long> first_bad;
  ING_CUDA(u_off.alloc(L)); ING_CUDA(s_off.alloc(L)); ING_CUDA(u_len.alloc(L)); ING_CUDA(s_len.alloc(L)); ING_CUDA(u_hash.alloc(L)); ING_CUDA(s_hash.alloc(L));
  ING_CUDA(first_bad.alloc(1));
  ING_CUDA(cudaMemset(first_bad.p, 0xff, sizeof(unsigned long long)));
  lap("per-line arrays alloc");
  for (int f = 0; f < 3; ++f)
    if (n_lines[f])
      parse_lines_kernel<<<blocks_for(n_lines[f]), 256>>>(buf.p, fbeg[f], fbeg[f + 1], nl[f].p, n_nl[f], n_lines[f], line0[f], u_off.p, u_len.p, s_off.p,
                                                          s_len.p, u_hash.p, s_hash.p, first_bad.p);
  unsigned long long bad = 0;
  ING_CUDA(cudaMemcpy(&bad, first_bad.p, sizeof bad, cudaMemcpyDeviceToHost));
  if (bad != ~0ULL) {
    int f = 0;
    while (f < 2 && static_cast<long long>(bad) >= line0[f + 1]) ++f;
    return ing_fail(g, MR_ERR_BAD_ARG, "scala.MatchError: line %lld of the %s file does not have 3 tab-separated fields", static_cast<long long>(bad) - line0[f] + 1,
                    kFileName[f]);
  }
  lap("parse");
  g->timing[1] = ms_since(t0);
  // per-element flags: users 1 = train / test line, 2 = label line; songs 1 = train or test line, 2 = label line
  DVec<int> flag;
  ING_CUDA(flag.alloc(L));
  if (L) line_flags_kernel<<<blocks_for(L), 256>>>(flag.p, L, line0[2]);
  lap("flags");
  // ---- 3/4: id spaces
  t0 = std::chrono::steady_clock::now();
  DVec<int> user_id, song_id;
  ING_CUDA(user_id.alloc(L)); ING_CUDA(song_id.alloc(L));
  double host_sort_ms = 0;
  int rc;
  {  // train users: elements [0, line0[1])
    UniqueSet us;
    if ((rc = unique_fields(g, buf.p, u_hash.p, u_off.p, u_len.p, flag.p, line0[1], tmp, us))) return rc;
    auto th = std::chrono::steady_clock::now();
    std::vector<int> rank; int n_primary = 0;
    rank_uniques(us, 1, false, rank, g->chars[0], g->off[0], &n_primary);
    host_sort_ms += ms_since(th);
    g->n_train = n_primary;
    DVec<int> d_rank;
    ING_CUDA(d_rank.alloc(rank.size()));
    if (!rank.empty()) ING_CUDA(cudaMemcpy(d_rank.p, rank.data(), rank.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (line0[1]) apply_rank_kernel<<<blocks_for(line0[1]), 256>>>(us.elem_uniq.p, d_rank.p, line0[1], user_id.p);
    ING_CUDA(cudaDeviceSynchronize());
  }
  {  // test users ∪ label users: elements [line0[1], L); only users of the test file get ids
    UniqueSet us;
    const long long e0 = line0[1], n = L - e0;
    if ((rc = unique_fields(g, buf.p, u_hash.p + e0, u_off.p + e0, u_len.p + e0, flag.p + e0, n, tmp, us))) return rc;
    auto th = std::chrono::steady_clock::now();
    std::vector<int> rank; int n_primary = 0;
    rank_uniques(us, 1, false, rank, g->chars[1], g->off[1], &n_primary);
    host_sort_ms += ms_since(th);
    g->n_test = n_primary;
    DVec<int> d_rank;
    ING_CUDA(d_rank.alloc(rank.size()));
    if (!rank.empty()) ING_CUDA(cudaMemcpy(d_rank.p, rank.data(), rank.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (n) apply_rank_kernel<<<blocks_for(n), 256>>>(us.elem_uniq.p, d_rank.p, n, user_id.p + e0);
    ING_CUDA(cudaDeviceSynchronize());
  }
  g->timing[2] = ms_since(t0) - host_sort_ms;
  t0 = std::chrono::steady_clock::now();
  double host_sort_songs = 0;
  {  // songs: every element; songs of the train / test files first (MR:38, 51, 58), label-only songs after them
    UniqueSet us;
    if ((rc = unique_fields(g, buf.p, s_hash.p, s_off.p, s_len.p, flag.p, L, tmp, us))) return rc;
    auto th = std::chrono::steady_clock::now();
    std::vector<int> rank; int n_primary = 0;
    rank_uniques(us, 1, true, rank, g->chars[2], g->off[2], &n_primary);
    host_sort_songs = ms_since(th);
    g->n_songs = n_primary;
    g->n_label_only = us.n_uniq - n_primary;
    DVec<int> d_rank;
    ING_CUDA(d_rank.alloc(rank.size()));
    if (!rank.empty()) ING_CUDA(cudaMemcpy(d_rank.p, rank.data(), rank.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (L) apply_rank_kernel<<<blocks_for(L), 256>>>(us.elem_uniq.p, d_rank.p, L, song_id.p);
    ING_CUDA(cudaDeviceSynchronize());
  }
  g->timing[3] = ms_since(t0) - host_sort_songs;
  g->timing[4] = host_sort_ms + host_sort_songs;
  // ---- 5: matrices
  t0 = std::chrono::steady_clock::now();
  if ((rc = build_csr(g, user_id.p, song_id.p, n_lines[0], g->n_train, tmp, g->ptr[0], g->col[0], &g->deg_train))) return rc;
  if ((rc = build_csr(g, user_id.p + line0[1], song_id.p + line0[1], n_lines[1], g->n_test, tmp, g->ptr[1], g->col[1], &g->deg_test))) return rc;
  if ((rc = build_csr(g, user_id.p + line0[2], song_id.p + line0[2], n_lines[2], g->n_test, tmp, g->ptr[2], g->col[2], nullptr))) return rc;
  {  // |songsToUsersMap(s)| over train AND test-visible rows, duplicates included (MR:41, 53, 237)
    DVec<int> deg;
    ING_CUDA(deg.alloc(std::max(g->n_songs, 1)));
    ING_CUDA(cudaMemset(deg.p, 0, static_cast<size_t>(std::max(g->n_songs, 1)) * sizeof(int)));
    if (line0[2]) bincount_kernel<<<blocks_for(line0[2]), 256>>>(song_id.p, line0[2], g->n_songs, deg.p);
    g->deg_song.assign(g->n_songs, 0);
    if (g->n_songs) ING_CUDA(cudaMemcpy(g->deg_song.data(), deg.p, static_cast<size_t>(g->n_songs) * sizeof(int32_t), cudaMemcpyDeviceToHost));
  }
  g->timing[5] = ms_since(t0);
  g->timing[7] = ms_since(t_total);
  return MR_OK;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int mr_ingest_tsv(int device, const char* train, uint64_t train_len, const char* test, uint64_t test_len, const char* labels,
                  uint64_t labels_len, mr_ingest** out) {
  if (!out) return MR_ERR_BAD_ARG;
  mr_ingest* g = new mr_ingest();
  *out = g;   // returned even on failure so that mr_ingest_error works; the caller still calls mr_ingest_free
  const char* bufs[3] = {train, test, labels};
  const uint64_t lens[3] = {train_len, test_len, labels_len};
  return ingest_impl(g, device, bufs, lens);
}

const char* mr_ingest_error(const mr_ingest* g) { return g ? g->err.c_str() : "null ingest result"; }

int mr_ingest_dims(const mr_ingest* g, int32_t* n_train, int32_t* n_test, int32_t* n_songs, int32_t* n_label_only_songs) {
  if (!g) return MR_ERR_BAD_ARG;
  if (n_train) *n_train = g->n_train;
  if (n_test) *n_test = g->n_test;
  if (n_songs) *n_songs = g->n_songs;
  if (n_label_only_songs) *n_label_only_songs = g->n_label_only;
  return MR_OK;
}

int mr_ingest_get(const mr_ingest* g, int which, const void** ptr, int64_t* n_elems) {
  if (!g || !ptr || !n_elems) return MR_ERR_BAD_ARG;
  auto set = [&](const void* p, size_t n) { *ptr = p; *n_elems = static_cast<int64_t>(n); return MR_OK; };
  switch (which) {
    case MR_ING_TR_PTR: return set(g->ptr[0].data(), g->ptr[0].size());
    case MR_ING_TR_COL: return set(g->col[0].data(), g->col[0].size());
    case MR_ING_TE_PTR: return set(g->ptr[1].data(), g->ptr[1].size());
    case MR_ING_TE_COL: return set(g->col[1].data(), g->col[1].size());
    case MR_ING_LAB_PTR: return set(g->ptr[2].data(), g->ptr[2].size());
    case MR_ING_LAB_COL: return set(g->col[2].data(), g->col[2].size());
    case MR_ING_DEG_TRAIN: return set(g->deg_train.data(), g->deg_train.size());
    case MR_ING_DEG_TEST: return set(g->deg_test.data(), g->deg_test.size());
    case MR_ING_DEG_SONG: return set(g->deg_song.data(), g->deg_song.size());
    case MR_ING_TRAIN_USER_CHARS: return set(g->chars[0].data(), g->chars[0].size());
    case MR_ING_TRAIN_USER_OFF: return set(g->off[0].data(), g->off[0].size());
    case MR_ING_TEST_USER_CHARS: return set(g->chars[1].data(), g->chars[1].size());
    case MR_ING_TEST_USER_OFF: return set(g->off[1].data(), g->off[1].size());
    case MR_ING_SONG_CHARS: return set(g->chars[2].data(), g->chars[2].size());
    case MR_ING_SONG_OFF: return set(g->off[2].data(), g->off[2].size());
    case MR_ING_TIMING_MS: return set(g->timing, 8);
    default: return MR_ERR_BAD_ARG;
  }
}

void mr_ingest_free(mr_ingest* g) { delete g; }

}  // extern "C"
#pragma GCC visibility pop
