// mrscore.cu — the C-ABI of include/mrscore.h: handle, data residency in HBM, and the batch pipeline K1 -> K2 -> K3.
//
// HBM layout:
//   train CSR  (user -> songs)   tr_ptr i64[T+1], tr_col i32[nnz]        MusicRecommender.scala:55  trainUsersToSongsMap
//   train CSC  (song -> users)   csc_ptr i64[S+1], csc_idx i32[nnz]      MusicRecommender.scala:60  songsToUsersMap (train part)
//   qv u32[T], qd u32[S]         rint(2^24 / sqrt(deg)), rint(2^26 / sqrt(deg))   cosine denominators MR:147 / MR:237 in fixed point
//   rsa f64[U], rsd f64[S]       2^-24 / sqrt(deg), 2^-26 / sqrt(deg)              the other half of each denominator
//   A_tr  u8[T][pitchS], A_trT u8[S][pitchT]   dense 0/1 operands of the tensor-core count GEMM (tensor engine only)
//   user space, per batch of 128 test users: Ct u16[T][128], Sint_u / Sint_i i64[128][spitch], G i32[rows][ldg], select bits
//   item space: head rows G16 / Gq32 (once per train set), Sint_u / Sint_i i64[batch][spitch] with batch = as many test users as fit
//   (plan_item_batches), balanced work groups of the head pass, per-slice tail scatter -> mask -> top-k -> streamed copy of the result
#include "../../include/mrscore.h"
#include "mr_common.cuh"
#include "mr_kernels.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <queue>
#include <mutex>
#include <thread>
#include <vector>

using namespace mr;

namespace {

constexpr int kSplitLen = 4096;   // listeners per K2 work item
// largest degrees whose fixed-point cosine factor stays within 1e-5 relative: 0.5 * sqrt(deg) / 2^k <= 1e-5  <=>  deg <= (2e-5 * 2^k)^2
constexpr int kMaxDegUbm = 112589;      // (2e-5 * 2^24)^2 = 112 589.99
constexpr int kPreStreams = 4;          // streams of the in-place head-row build (start_head_rows): 113.8 / 82.9 / 75.8 / 73.4 ms with 1 / 2 / 3 / 4
constexpr int kPreExtraStreams = 3;     // at most 1 + 3 (MRSCORE_PRE_STREAMS)
constexpr int kMaxDegIbm = 1801439;     // (2e-5 * 2^26)^2 = 1 801 439.85

}  // namespace

struct mr_handle {
  int device = 0; int num_sms = 148; unsigned flags = 0; int engine = MR_ENGINE_AUTO;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr; cudaEvent_t ev_slice = nullptr;   // mr_topk streams finished slices of the result to the host beside the compute
  // Batch pipeline of the item-space top-k (run_batches): the head pass of batch b + 1 is issued on `stream` while the tail scatter, mask
  // and select of batch b run on slice_stream (higher priority); the two batches use the two Sint panels (a pure model needs only one).
  // ev_head[p] / ev_done[p]: head pass / slices of the batch in panel p.
  cudaStream_t slice_stream = nullptr; cudaEvent_t ev_head[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  cudaStream_t slice_stream2 = nullptr; cudaEvent_t ev_join = nullptr;   // single-batch calls alternate their slices between the two slice streams
  std::string err;
  long long launches = 0; size_t dev_bytes = 0;
  std::vector<void*> allocs;          // everything freed in mr_destroy
  // train
  int T = 0, S = 0; long long nnz_tr = 0; bool loaded = false;
  long long *d_tr_ptr = nullptr, *d_csc_ptr = nullptr; int *d_tr_col = nullptr, *d_csc_idx = nullptr;
  // Song window (MR_OPT_SONG_WINDOW_LO / _HI, the song partition of distributed.scala:459-461): only the songs [win_lo, win_hi) are scored.
  // Internally song ids are ROTATED by win_lo (id' = id - win_lo mod S), so that the scored columns are [0, n_cols) of every array and
  // kernel while histories still refer to all S songs; CSR rows stay ascending (their in-window entries become a prefix that ends at
  // d_tr_end[v] / d_te_end[u]); only the ranked ids handed back are shifted by win_lo again.  Without a window n_cols = S, *_end = *_ptr + 1.
  int win_lo = 0, win_hi = 0, n_cols = 0; bool windowed = false;
  const long long *d_tr_end = nullptr, *d_te_end = nullptr;
  // asynchronous head-row build (mr_prepare_async): the kernels run on pre_stream, ev_pre marks their end; head_pending until the
  // exception list has been read back and uploaded (finish_head_rows, by the first call that needs the rows)
  cudaStream_t pre_stream = nullptr; cudaEvent_t ev_pre = nullptr; bool head_pending = false; unsigned int* h_n_ex = nullptr;
  cudaStream_t pre_extra[kPreExtraStreams] = {}; cudaEvent_t ev_pre_extra[kPreExtraStreams] = {};   // extra streams of the in-place row build (chunks rotate over them)
  HeadExceptions pend_ex{};
  unsigned int* d_topk_stats = nullptr;   // [3] counters of the top-k select since mr_load (BlendParams::stats)
  uint32_t *d_qv = nullptr, *d_qd = nullptr; double* d_rsd = nullptr; float *d_rsv_f = nullptr, *d_rsd_f = nullptr, *d_rsd_up = nullptr;
  std::vector<int32_t> deg_song, deg_song_train;
  struct SongInfo { int head; uint32_t v; };   // head row or -1; v = q_26(d_s) of a head song, train listeners of a tail song
  int2* d_song_info = nullptr;                 // the same records on the device (k7_testlists.cu)
  std::vector<SongInfo> song_info;             // per song, 8 bytes: the one random access per test entry in mr_set_test_users
  uint32_t max_qv = 0;                         // largest UBM weight q_24(|I_v|) of a train user (number of byte planes of the tensor precompute)
  unsigned long long max_qsum = 0;             // max over songs of song_qsum
  std::vector<unsigned long long> song_qsum;   // per song: sum of qv over its train listeners = upper bound of any Gq entry of its row
  bool ubm_int_ok = false;                     // every UBM numerator of the current shard is provably < 2^52 (top-k may rank the integers)
  int *d_item_song = nullptr, *d_item_len = nullptr; long long* d_item_begin = nullptr; uint8_t* d_item_split = nullptr; int n_items = 0;
  long long pitchS = 0, pitchT = 0, spitch = 0, ldg = 0; size_t dense_bytes = 0;
  uint8_t *d_Atr = nullptr, *d_AtrT = nullptr;
  // item-space engine: head songs and their precomputed rows
  int space_flag = MR_SPACE_AUTO; int space = MR_SPACE_USER; int n_head = 0; bool head_ready = false;
  int item_batch_cap = 0;              // MRSCORE_ITEM_BATCH: upper bound on test users per item-space batch (0 = as many as fit in HBM)
  int batch_rows = kUserBatch;         // test users per batch of the current shard = rows of the Sint panels
  int head_words_u = 4, head_words_i = 4, head_threads = 128, n_groups = 0;   // head_rowsum: 32-bit words per row load (UBM / IBM pass), groups per batch
  int4 *d_seg = nullptr, *d_grp_hdr = nullptr; int *d_ge_row = nullptr, *d_split_rows = nullptr; uint32_t* d_ge_q = nullptr; int seg_cap = 1, ent_cap = 1;
  std::vector<int> h_split_ptr;   // balanced work groups per batch
  std::vector<int> head_index;         // song -> head row or -1
  int* d_head_song = nullptr; long long* d_head_lst_ptr = nullptr; uint16_t* d_g16 = nullptr; uint32_t* d_gq32 = nullptr;
  size_t head_rows_cap = 0;            // entries the packed-row allocations hold (kept across mr_invalidate_prepared)
  int n_head_staged = 0;               // rows [0, n_head_staged) are built in a 64-bit staging area and packed; the rest directly in place
  long long opt_head_min_deg = 0;      // mr_set_option(MR_OPT_HEAD_MIN_DEG): 0 = default rule
  long long* d_ex_ptr = nullptr; int* d_ex_song = nullptr; uint32_t* d_ex_g = nullptr; unsigned long long* d_ex_gq = nullptr; long long n_ex = 0;
  long long *d_hu_ptr = nullptr; int *d_hu_row = nullptr, *d_hu_song = nullptr; uint32_t* d_hu_q = nullptr;
  int *d_tu_user = nullptr, *d_tu_song = nullptr; long long* d_tu_lptr = nullptr; std::vector<long long> h_tu_ptr, h_tu_lptr; long long n_head_entries = 0, n_tail_entries = 0;
  // test shard (freed / reallocated by mr_set_test_users)
  int U = 0; long long nnz_te = 0; bool have_test = false;
  // grow-only device buffers of the test shard and its results: steady-state mr_set_test_users / mr_topk calls do no cudaMalloc
  enum { SL_TE_PTR, SL_TE_COL, SL_TE_GROW, SL_RSA, SL_RSA_F, SL_PAIR_BASE, SL_ROWS, SL_HU_PTR, SL_HU_ROW, SL_HU_SONG, SL_HU_Q, SL_TU_USER,
         SL_TU_SONG, SL_TU_LPTR, SL_TU_PTR, SL_EX_PTR, SL_EX_SONG, SL_EX_G, SL_EX_GQ, SL_L_FLAG, SL_L_HEADPOS, SL_L_DEG, SL_L_LSUM, SL_L_TMP, SL_COPY_DESC, SL_SEG, SL_GRP_HDR, SL_GE_ROW, SL_GE_Q, SL_SPLIT_ROWS, SL_SINT_U, SL_SINT_I, SL_SEL, SL_GRAM_IDS, SL_CNT, SL_SIMF, SL_DENSE, SL_OUT_PACK, SL_TE_END, SL_PRE_G, SL_PRE_GQ, SL_PRE_EXC, SL_N };
  void* slot_p[SL_N] = {}; size_t slot_cap[SL_N] = {};
  long long *d_te_ptr = nullptr, *d_pair_base = nullptr; int *d_te_col = nullptr, *d_te_grow = nullptr; double* d_rsa = nullptr; float* d_rsa_f = nullptr;
  std::vector<long long> h_te_ptr; std::vector<int> h_te_col;
  std::vector<long long> batch_row_off; int* d_rows = nullptr; int max_batch_rows = 0;
  long long pair_index_base = 0, n_pairs_total = 0;
  // workspaces
  uint8_t *d_Ate = nullptr, *d_Aj = nullptr; uint16_t* d_ct = nullptr; long long *d_sint_u = nullptr, *d_sint_i = nullptr;
  uint32_t* d_wi = nullptr;             // IBM user-space weighted counts u32[T][128] (sparse engine)
  unsigned int* d_carry_count = nullptr; uint2* d_carry_events = nullptr; unsigned int carry_cap = 1u << 20; unsigned int* h_carry_seen = nullptr;
  uint64_t* d_sel = nullptr; long long sel_pitch = 0; int32_t* d_g = nullptr; size_t g_bytes = 0; size_t aj_bytes = 0;
  double* d_dense = nullptr; int32_t* d_cnt = nullptr; float* d_simf = nullptr;
  // results
  int *d_out_song = nullptr, *d_out_len = nullptr; double* d_out_score = nullptr; int out_k = 0; bool have_topk = false; size_t out_pack_bytes = 0;
  // profiling
  cudaEvent_t ev[2] = {nullptr, nullptr}; double t_ms[MR_T_N] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};

namespace {

int fail(mr_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (h) h->err = buf;
  return code;
}

#define MR_CUDA(h, expr)                                                                                       \
  do {                                                                                                         \
    cudaError_t e__ = (expr);                                                                                  \
    if (e__ != cudaSuccess)                                                                                    \
      return fail(h, e__ == cudaErrorMemoryAllocation ? MR_ERR_OOM : MR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, \
                  cudaGetErrorString(e__), __FILE__, __LINE__);                                                \
  } while (0)

#define MR_LAUNCH(h, expr)                                                                     \
  do {                                                                                         \
    int rc__ = (expr);                                                                         \
    (h)->launches++;                                                                           \
    if (rc__ != 0) {                                                                           \
      cudaError_t e__ = cudaGetLastError();                                                    \
      return fail(h, MR_ERR_CUDA, "%s failed: rc=%d (%s)", #expr, rc__, cudaGetErrorString(e__)); \
    }                                                                                          \
  } while (0)

template <class Tp>
int dev_alloc(mr_handle* h, Tp** out, size_t count, std::vector<void*>& owner) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(count, 1) * sizeof(Tp);
  MR_CUDA(h, cudaMalloc(&p, bytes));
  owner.push_back(p);
  h->dev_bytes += bytes;
  *out = static_cast<Tp*>(p);
  return MR_OK;
}

template <class Tp>
int slot_alloc(mr_handle* h, int slot, Tp** out, size_t count, bool exact = false) {
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(Tp);
  if (h->slot_cap[slot] < bytes) {
    if (h->slot_p[slot]) { cudaFree(h->slot_p[slot]); h->dev_bytes -= h->slot_cap[slot]; h->slot_p[slot] = nullptr; h->slot_cap[slot] = 0; }
    const size_t cap = exact ? bytes : bytes + bytes / 4 + 256;
    void* p = nullptr;
    MR_CUDA(h, cudaMalloc(&p, cap));
    h->slot_p[slot] = p; h->slot_cap[slot] = cap; h->dev_bytes += cap;
  }
  *out = static_cast<Tp*>(h->slot_p[slot]);
  return MR_OK;
}

template <class Tp>
int slot_upload(mr_handle* h, int slot, Tp** out, const Tp* src, size_t count) {
  int rc = slot_alloc(h, slot, out, count);
  if (rc) return rc;
  if (count) MR_CUDA(h, cudaMemcpyAsync(*out, src, count * sizeof(Tp), cudaMemcpyHostToDevice, h->stream));
  return MR_OK;
}

template <class Tp>
int dev_upload(mr_handle* h, Tp** out, const Tp* src, size_t count, std::vector<void*>& owner) {
  int rc = dev_alloc(h, out, count, owner);
  if (rc) return rc;
  if (count) MR_CUDA(h, cudaMemcpyAsync(*out, src, count * sizeof(Tp), cudaMemcpyHostToDevice, h->stream));
  return MR_OK;
}

inline uint32_t q_of(int32_t deg, double scale) { return deg <= 0 ? 0u : static_cast<uint32_t>(llrint(scale / std::sqrt(static_cast<double>(deg)))); }
inline double rs_of(int32_t deg, double inv) { return deg <= 0 ? 0.0 : inv / std::sqrt(static_cast<double>(deg)); }
inline float rsf_of(int32_t deg) { return deg <= 0 ? 0.0f : static_cast<float>(1.0 / std::sqrt(static_cast<double>(deg))); }
inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

// f(lo, hi) over disjoint row ranges on a few host threads.  The per-shard host work of mr_set_test_users (CSR validation, id rotation,
// cosine factors: O(test entries)) sits on the calling thread inside every end-to-end step; a rank of an 8-GPU job shares the host with
// seven others, hence at most hardware threads / 8 (and at most 8) workers.
template <class F>
void parallel_rows(int n, F&& f) {
  const int hw = static_cast<int>(std::thread::hardware_concurrency());
  const int t = std::max(1, std::min(8, hw / 8));
  if (n < 16384 || t == 1) { f(0, n); return; }
  const int chunk = (n + t - 1) / t;
  std::vector<std::thread> workers;
  for (int i = 1; i < t; ++i) workers.emplace_back([&f, i, chunk, n] { f(std::min(n, i * chunk), std::min(n, (i + 1) * chunk)); });
  f(0, std::min(n, chunk));
  for (auto& w : workers) w.join();
}

int check_csr(mr_handle* h, const char* what, int n_rows, int n_cols, const int64_t* ptr, const int32_t* col) {
  if (!ptr || (!col && ptr[n_rows] > 0)) return fail(h, MR_ERR_BAD_ARG, "%s: null CSR arrays", what);
  if (ptr[0] != 0) return fail(h, MR_ERR_BAD_ARG, "%s: rowptr[0] != 0", what);
  for (int r = 0; r < n_rows; ++r)
    if (ptr[r + 1] < ptr[r]) return fail(h, MR_ERR_BAD_ARG, "%s: rowptr not monotone at row %d", what, r);
  // the first offending row (smallest index) is reported, whichever thread finds it
  struct Bad { int row = -1; int kind = 0; int col = 0; };
  Bad best; std::mutex mu;
  parallel_rows(n_rows, [&](int lo, int hi) {
    Bad b;
    for (int r = lo; r < hi && b.row < 0; ++r) {
      for (int64_t i = ptr[r]; i < ptr[r + 1]; ++i) {
        if (col[i] < 0 || col[i] >= n_cols) { b.row = r; b.kind = 1; b.col = col[i]; break; }
        if (i > ptr[r] && col[i] <= col[i - 1]) { b.row = r; b.kind = 2; break; }
      }
    }
    if (b.row >= 0) {
      std::lock_guard<std::mutex> g(mu);
      if (best.row < 0 || b.row < best.row) best = b;
    }
  });
  const Bad& b = best;
  if (b.row >= 0 && b.kind == 1) return fail(h, MR_ERR_BAD_ARG, "%s: column id %d out of range at row %d", what, b.col, b.row);
  if (b.row >= 0) return fail(h, MR_ERR_BAD_ARG, "%s: row %d not ascending/unique", what, b.row);
  return MR_OK;
}

struct PhaseTimer {   // CUDA events on the library stream (or the given one) around one phase (only with MR_PROFILE)
  mr_handle* h; int phase; bool on; cudaStream_t st;
  PhaseTimer(mr_handle* hh, int ph, cudaStream_t s = nullptr) : h(hh), phase(ph), on((hh->flags & MR_PROFILE) != 0), st(s ? s : hh->stream) {
    if (on) cudaEventRecord(h->ev[0], st);
  }
  ~PhaseTimer() {
    if (!on) return;
    cudaEventRecord(h->ev[1], st);
    cudaEventSynchronize(h->ev[1]);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) h->t_ms[phase] += ms;
  }
};

__global__ void ct_to_i32_kernel(const uint16_t* __restrict__ ct, int n_train, int nb, int32_t* __restrict__ out) {
  // out[b][v] = Ct[v][b]
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < static_cast<long long>(nb) * n_train;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / n_train), v = static_cast<int>(i % n_train);
    out[i] = ct[ct_index(n_train, v, b)];
  }
}
__global__ void ct_to_cos_kernel(const uint16_t* __restrict__ ct, int n_train, int nb, const float* __restrict__ rsa,
                                 const float* __restrict__ rsv, float* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < static_cast<long long>(nb) * n_train;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / n_train), v = static_cast<int>(i % n_train);
    out[i] = static_cast<float>(ct[ct_index(n_train, v, b)]) * (rsa[b] * rsv[v]);
  }
}
__global__ void gram_to_cos_kernel(const int32_t* __restrict__ g, long long ldg, const int* __restrict__ rows, int n_rows, int n_songs,
                                   const float* __restrict__ rsd, float* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < static_cast<long long>(n_rows) * n_songs;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / n_songs), s = static_cast<int>(i % n_songs);
    out[i] = static_cast<float>(g[r * ldg + s]) * (rsd[rows[r]] * rsd[s]);
  }
}
__global__ void iota_kernel(int* p, int start, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = start + i;
}

void free_list(std::vector<void*>& v) {
  for (void* p : v) cudaFree(p);
  v.clear();
}

// ---- K1 for one batch of test users -> Ct u16[T][128]
int count_ubm_batch(mr_handle* h, int b0, int nb) {
  if (h->engine == MR_ENGINE_TENSOR) {
    {
      PhaseTimer t(h, MR_T_EXPAND);
      MR_LAUNCH(h, launch_expand_rows(h->d_te_ptr, h->d_te_col, nullptr, b0, nb, kUserBatch, h->pitchS, h->d_Ate, h->stream));
    }
    PhaseTimer t(h, MR_T_COUNT);
    MR_LAUNCH(h, launch_count_gemm(h->d_Ate, kUserBatch, h->d_Atr, h->T, h->pitchS, kUserBatch, h->T, EPI_U16_T, h->d_ct, kUserBatch,
                                   nullptr, nullptr, h->num_sms, h->stream));
  } else {
    PhaseTimer t(h, MR_T_COUNT);
    MR_LAUNCH(h, launch_sparse_count_u16t(h->d_te_ptr, h->d_te_col, b0, nb, h->d_csc_ptr, h->d_csc_idx, h->d_ct, h->T, h->stream));
  }
  return MR_OK;
}

// ---- K1 for a list of Gram rows (device array of song ids) -> G i32[n_rows][ldg]
int gram_rows(mr_handle* h, const int* d_rows, int n_rows) {
  if (h->engine == MR_ENGINE_TENSOR) {
    const int rows_pad = static_cast<int>(round_up(n_rows, 128));
    {
      PhaseTimer t(h, MR_T_EXPAND);
      MR_LAUNCH(h, launch_expand_rows(h->d_csc_ptr, h->d_csc_idx, d_rows, 0, n_rows, rows_pad, h->pitchT, h->d_Aj, h->stream));
    }
    PhaseTimer t(h, MR_T_COUNT);
    MR_LAUNCH(h, launch_count_gemm(h->d_Aj, rows_pad, h->d_AtrT, h->S, h->pitchT, n_rows, h->S, EPI_I32, h->d_g, h->ldg, nullptr,
                                   nullptr, h->num_sms, h->stream));
  } else {
    PhaseTimer t(h, MR_T_COUNT);
    for (int r0 = 0; r0 < n_rows; r0 += 32768) {
      const int n = std::min(32768, n_rows - r0);
      MR_LAUNCH(h, launch_sparse_gram_rows(d_rows + r0, n, h->d_csc_ptr, h->d_csc_idx, h->d_tr_ptr, h->d_tr_col,
                                           h->d_g + static_cast<long long>(r0) * h->ldg, h->ldg, h->stream));
    }
  }
  return MR_OK;
}

int ensure_gram_ws(mr_handle* h, int n_rows) {
  const size_t need_g = static_cast<size_t>(round_up(std::max(n_rows, 1), 128)) * h->ldg * sizeof(int32_t);
  if (need_g > h->g_bytes) {
    int32_t* p; int rc = dev_alloc(h, &p, need_g / sizeof(int32_t), h->allocs);
    if (rc) return rc;
    h->d_g = p; h->g_bytes = need_g;
  }
  if (h->engine == MR_ENGINE_TENSOR) {
    const size_t need_a = static_cast<size_t>(round_up(std::max(n_rows, 1), 128)) * h->pitchT;
    if (need_a > h->aj_bytes) {
      uint8_t* p; int rc = dev_alloc(h, &p, need_a, h->allocs);
      if (rc) return rc;
      h->d_Aj = p; h->aj_bytes = need_a;
    }
  }
  return MR_OK;
}

// Item-space: compute the rows G[h][:], Gq[h][:] of the head songs once per train set (lazily, on first use, or mr_prepare).
// Head rows are ordered by train degree, descending.  Two construction paths, both from the inverted index:
//   staged rows [0, n_head_staged): the most popular songs, whose entries can exceed 16 / 32 bits — one packed 64-bit accumulator per
//     entry in an L2-resident staging chunk, then pack_head_rows_kernel writes the 6-byte form + the exact exception list;
//   direct rows [n_head_staged, n_head): every song whose weighted listener sum is < 2^32 and whose train degree is < 2^16 — no entry
//     of such a row can overflow, so the events are added straight into the final u16 / u32 arrays (32-bit L2 atomics), an L2-sized
//     chunk of rows at a time: zero the chunk (it stays dirty in L2), scatter into it, let it drain to HBM once.  This path moves
//     6 bytes per entry instead of 8 (zero) + 8 (read) + 6 (write) and has no pack pass: it is what made the sparse 85 % of the head
//     rows cheap (they carry few events but paid the full dense staging traffic).
// On the tensor engine (dense-friendly shapes) every row is staged and computed by K1: one 0/1 count GEMM + byte-plane GEMMs of the weights.
int finish_head_rows(mr_handle* h);

// Launch the whole build on stream st (the library stream, or pre_stream for mr_prepare_async) and leave the handle in the
// head_pending state; nothing here waits for the device except the tensor engine's temporary operands.
int start_head_rows(mr_handle* h, cudaStream_t st) {
  if (h->head_ready || h->head_pending) return MR_OK;
  int rc;
  if (st != h->stream) {   // the build overwrites rows that earlier work on the library stream may still be reading
    MR_CUDA(h, cudaEventRecord(h->ev_pre, h->stream));
    MR_CUDA(h, cudaStreamWaitEvent(st, h->ev_pre, 0));
  }
  const bool dbg = getenv("MRSCORE_DEBUG_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
  auto t_start = now();
  const size_t n = static_cast<size_t>(std::max(h->n_head, 1)) * h->spitch;
  if (h->head_rows_cap < n) {   // allocated once per handle; a failed or invalidated precompute reuses it
    if (h->d_g16) { cudaFree(h->d_g16); h->d_g16 = nullptr; }
    if (h->d_gq32) { cudaFree(h->d_gq32); h->d_gq32 = nullptr; }
    h->dev_bytes -= h->head_rows_cap * 6; h->head_rows_cap = 0;
    MR_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->d_g16), n * sizeof(uint16_t)));
    cudaError_t em = cudaMalloc(reinterpret_cast<void**>(&h->d_gq32), n * sizeof(uint32_t));
    if (em != cudaSuccess) { cudaFree(h->d_g16); h->d_g16 = nullptr; return fail(h, MR_ERR_OOM, "head rows: %s", cudaGetErrorString(em)); }
    h->head_rows_cap = n; h->dev_bytes += n * 6;
  }
  const bool tensor = h->engine == MR_ENGINE_TENSOR && h->n_head > 0 && h->T < (1 << 23);
  const int n_staged = tensor ? h->n_head : h->n_head_staged;
  // number of byte planes the weights q_24(|I_v|) need (3 when every train user has >= 2 songs: q < 2^24)
  int n_planes = 1;
  while (n_planes < 4 && (static_cast<unsigned long long>(h->max_qv) >> (8 * n_planes)) != 0) ++n_planes;
  // staging chunk.  Tensor engine: <= 8 GiB of (u32 + u64) rows, a multiple of the 128-row GEMM tile.  Scatter path: small enough
  // (<= 64 MiB) that the rows being built stay L2-resident, so the integer atomics are resolved in L2 instead of as DRAM read-modify-writes.
  long long chunk = 1; bool packed = false;
  if (tensor) {
    chunk = (8LL << 30) / (h->spitch * 12) / 128 * 128;
    chunk = std::max<long long>(128, std::min<long long>(chunk, 4096));
  } else if (n_staged > 0) {
    // one packed 64-bit accumulator per entry when no weighted sum can reach 2^44 and no count 2^20 (any realistic data set)
    packed = h->max_qsum < (1ULL << 44) && *std::max_element(h->deg_song_train.begin(), h->deg_song_train.end()) < (1 << 20) &&
             !getenv("MRSCORE_PRECOMPUTE_UNPACKED");
    chunk = std::max<long long>(4, std::min<long long>((64LL << 20) / (h->spitch * (packed ? 8 : 12)), 4096));
    if (const char* e = getenv("MRSCORE_PRECOMPUTE_CHUNK")) chunk = std::max(1LL, atoll(e));
  }
  // The staging rows and the exception list live in grow-only slots of the handle: a rebuild (mr_invalidate_prepared + mr_prepare, i.e.
  // every step of a job that counts the model build) does no cudaMalloc / cudaFree — next to 100+ GB of live allocations those cost
  // tens to hundreds of milliseconds, more than the kernels of a song partition's precompute.  Only the tensor engine's dense operands
  // (small shapes) are temporary.
  std::vector<void*> tmp;
  uint32_t* g_stage = nullptr; unsigned long long* gq_stage = nullptr;
  HeadExceptions ex{};
  ex.capacity = n_staged > 0 ? (1u << 24) : 1u;
  const size_t stage_entries = n_staged > 0 ? static_cast<size_t>(chunk) * h->spitch : 1;
  {
    char* xb = nullptr;                 // [count (16 B)] [gq_extra u64] [row i32] [song i32] [g_extra u32]
    const size_t cap = ex.capacity;
    if ((rc = slot_alloc(h, mr_handle::SL_PRE_G, &g_stage, packed ? 1 : stage_entries, true)) ||
        (rc = slot_alloc(h, mr_handle::SL_PRE_GQ, &gq_stage, stage_entries, true)) || (rc = slot_alloc(h, mr_handle::SL_PRE_EXC, &xb, 16 + cap * 20, true))) return rc;
    ex.count = reinterpret_cast<unsigned int*>(xb);
    ex.gq_extra = reinterpret_cast<unsigned long long*>(xb + 16);
    ex.row = reinterpret_cast<int*>(xb + 16 + cap * 8);
    ex.song = reinterpret_cast<int*>(xb + 16 + cap * 12);
    ex.g_extra = reinterpret_cast<uint32_t*>(xb + 16 + cap * 16);
  }
  auto bail = [&](int code) { cudaStreamSynchronize(st); free_list(tmp); return code; };
  cudaError_t e = cudaMemsetAsync(ex.count, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return bail(fail(h, MR_ERR_CUDA, "memset: %s", cudaGetErrorString(e)));
  uint8_t *a_rows = nullptr, *b_plane = nullptr;
  if (tensor) {
    if ((rc = dev_alloc(h, &a_rows, static_cast<size_t>(chunk) * h->pitchT, tmp)) ||
        (rc = dev_alloc(h, &b_plane, static_cast<size_t>(h->S) * h->pitchT * n_planes, tmp))) { free_list(tmp); return rc; }
    // the GEMM epilogues write only the columns < S: the pad columns of the staging rows must not hold garbage (pack reads the full pitch)
    e = cudaMemsetAsync(g_stage, 0, stage_entries * sizeof(uint32_t), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(gq_stage, 0, stage_entries * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return bail(fail(h, MR_ERR_CUDA, "memset: %s", cudaGetErrorString(e)));
    // the byte planes of q_24(|I_v|) as weighted B operands (built once, reused by every chunk)
    PhaseTimer t(h, MR_T_EXPAND, st);
    for (int plane = 0; plane < n_planes; ++plane) {
      int lrc = launch_expand_rows_weighted(h->d_csc_ptr, h->d_csc_idx, h->d_qv, plane, h->S, h->pitchT,
                                            b_plane + static_cast<size_t>(plane) * h->S * h->pitchT, st);
      h->launches++;
      if (lrc) return bail(fail(h, MR_ERR_CUDA, "launch_expand_rows_weighted failed"));
    }
  }
  if (dbg) { cudaStreamSynchronize(st); fprintf(stderr, "[mrscore] precompute: allocations %.1f ms (staged rows %d in chunks of %lld, direct rows %d)\n", ms_since(t_start), n_staged, chunk, h->n_head - n_staged); }
  auto t_loop = now();
  for (int r0 = 0; r0 < n_staged; r0 += static_cast<int>(chunk)) {
    const int nr = std::min<int>(static_cast<int>(chunk), n_staged - r0);
    if (tensor) {
      // G = A_head · A_trT^T as one 0/1 count GEMM; Gq as byte-plane GEMMs (255 * T < 2^31 keeps each plane exact in int32)
      const int n_pad = static_cast<int>(round_up(nr, 128));
      {
        PhaseTimer t(h, MR_T_EXPAND, st);
        int lrc = launch_expand_rows(h->d_csc_ptr, h->d_csc_idx, h->d_head_song + r0, 0, nr, n_pad, h->pitchT, a_rows, st);
        h->launches++;
        if (lrc) return bail(fail(h, MR_ERR_CUDA, "launch_expand_rows failed"));
      }
      PhaseTimer t(h, MR_T_COUNT, st);
      int lrc = launch_count_gemm(a_rows, n_pad, h->d_AtrT, h->S, h->pitchT, nr, h->S, EPI_I32, g_stage, h->spitch, nullptr, nullptr, h->num_sms, st);
      h->launches++;
      for (int plane = 0; plane < n_planes && !lrc; ++plane) {
        lrc = launch_count_gemm(a_rows, n_pad, b_plane + static_cast<size_t>(plane) * h->S * h->pitchT, h->S, h->pitchT, nr, h->S, EPI_ACC_U64,
                                gq_stage, h->spitch, nullptr, nullptr, h->num_sms, st, 8 * plane, plane > 0);
        h->launches++;
      }
      if (lrc) return bail(fail(h, MR_ERR_CUDA, "launch_count_gemm failed: rc=%d (%s)", lrc, cudaGetErrorString(cudaGetLastError())));
    } else {
      PhaseTimer t(h, MR_T_PRECOMPUTE, st);
      int lrc = launch_gram_head_scatter(h->d_head_song, h->d_head_lst_ptr, r0, r0 + nr, h->d_csc_ptr, h->d_csc_idx, h->d_tr_ptr, h->d_tr_end,
                                         h->d_tr_col, h->d_qv, g_stage, gq_stage, h->spitch, packed ? 1 : 0, h->num_sms, st);
      h->launches++;
      if (lrc) return bail(fail(h, MR_ERR_CUDA, "launch_gram_head_scatter failed"));
    }
    PhaseTimer t(h, MR_T_PRECOMPUTE, st);
    int lrc = launch_pack_head_rows(g_stage, gq_stage, packed ? 1 : 0, r0, nr, h->spitch, h->n_cols, h->d_g16, h->d_gq32, ex, h->num_sms, st);
    h->launches++;
    if (lrc) return bail(fail(h, MR_ERR_CUDA, "launch_pack_head_rows failed"));
  }
  if (dbg) { cudaStreamSynchronize(st); fprintf(stderr, "[mrscore] precompute: staged rows %.1f ms\n", ms_since(t_loop)); }
  if (n_staged < h->n_head) {
    // direct rows: an L2-sized chunk of final rows at a time
    long long mb = 48;
    if (const char* ev = getenv("MRSCORE_DIRECT_CHUNK_MB")) mb = std::max(1LL, atoll(ev));
    const long long dchunk = std::max<long long>(1, (mb << 20) / (h->spitch * 6));
    PhaseTimer t(h, MR_T_PRECOMPUTE, st);
    // Several streams (kPreStreams): the chunks are independent (disjoint rows) and each one is three short enqueues (two memsets + a
    // kernel of ~1.3 waves), so consecutive chunks rotate over `st` and the extra build streams and the next chunk's CTAs start while
    // the previous grid drains: 113.8 -> 73.4 ms for the 32 k rows of the MSD-shaped set with four streams (profiles/r02_summary.md §10).
    // Not under MR_PROFILE (the phase timer brackets `st` only).
    int n_streams = kPreStreams;
    if (const char* ev = getenv("MRSCORE_PRE_STREAMS")) n_streams = std::max(1, std::min(atoi(ev), 1 + kPreExtraStreams));
    if (h->flags & MR_PROFILE) n_streams = 1;
    auto sync_extra = [&] { for (int i = 0; i < kPreExtraStreams; ++i) cudaStreamSynchronize(h->pre_extra[i]); };
    for (int i = 0; i + 1 < n_streams; ++i) {
      if ((e = cudaEventRecord(h->ev_pre_extra[i], st)) != cudaSuccess || (e = cudaStreamWaitEvent(h->pre_extra[i], h->ev_pre_extra[i], 0)) != cudaSuccess)
        return bail(fail(h, MR_ERR_CUDA, "head-row precompute: %s", cudaGetErrorString(e)));
    }
    int ci = 0;
    for (int r0 = n_staged; r0 < h->n_head; r0 += static_cast<int>(dchunk), ++ci) {
      const int nr = std::min<int>(static_cast<int>(dchunk), h->n_head - r0);
      const int si = ci % n_streams;
      int lrc = launch_gram_head_direct(h->d_head_song, h->d_head_lst_ptr, r0, r0 + nr, h->d_csc_ptr, h->d_csc_idx, h->d_tr_ptr, h->d_tr_end,
                                        h->d_tr_col, h->d_qv, h->d_g16, h->d_gq32, h->spitch, h->num_sms, si ? h->pre_extra[si - 1] : st);
      h->launches++;
      if (lrc) { sync_extra(); return bail(fail(h, MR_ERR_CUDA, "launch_gram_head_direct failed")); }
    }
    for (int i = 0; i + 1 < n_streams; ++i) {   // `st` joins the extra streams
      if ((e = cudaEventRecord(h->ev_pre_extra[i], h->pre_extra[i])) != cudaSuccess || (e = cudaStreamWaitEvent(st, h->ev_pre_extra[i], 0)) != cudaSuccess) {
        sync_extra();
        return bail(fail(h, MR_ERR_CUDA, "head-row precompute: %s", cudaGetErrorString(e)));
      }
    }
  }
  // the exception count travels to pinned host memory behind the kernels; finish_head_rows picks it up
  e = cudaMemcpyAsync(h->h_n_ex, ex.count, sizeof(unsigned int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaEventRecord(h->ev_pre, st);
  if (e != cudaSuccess) return bail(fail(h, MR_ERR_CUDA, "head-row precompute: %s", cudaGetErrorString(e)));
  if (!tmp.empty()) { cudaStreamSynchronize(st); free_list(tmp); }     // tensor engine only: its dense operands are temporary
  if (dbg) fprintf(stderr, "[mrscore] precompute: launched after %.1f ms\n", ms_since(t_start));
  h->pend_ex = ex;
  h->head_pending = true;
  return MR_OK;
}

// Complete a started build: wait for its kernels, turn the exception list into a CSR by head row (host sort; it is a few thousand
// entries on MSD-shaped data) and order the library stream behind the build.
int finish_head_rows(mr_handle* h) {
  if (h->head_ready) return MR_OK;
  if (!h->head_pending) return fail(h, MR_ERR_STATE, "no head-row build in flight");
  int rc;
  h->head_pending = false;
  const HeadExceptions ex = h->pend_ex;
  cudaError_t e = cudaEventSynchronize(h->ev_pre);
  if (e != cudaSuccess) return fail(h, MR_ERR_CUDA, "head-row precompute: %s", cudaGetErrorString(e));
  const unsigned int n_ex = *h->h_n_ex;
  if (n_ex > ex.capacity) return fail(h, MR_ERR_OOM, "head-row exception list overflowed (%u entries)", n_ex);
  std::vector<int> xr(n_ex), xs(n_ex); std::vector<uint32_t> xg(n_ex); std::vector<unsigned long long> xq(n_ex);
  if (n_ex) {
    e = cudaMemcpy(xr.data(), ex.row, n_ex * sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(xs.data(), ex.song, n_ex * sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(xg.data(), ex.g_extra, n_ex * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(xq.data(), ex.gq_extra, n_ex * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(h, MR_ERR_CUDA, "head-row exception list: %s", cudaGetErrorString(e));
  }
  std::vector<unsigned int> order(n_ex);
  for (unsigned int i = 0; i < n_ex; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](unsigned int a, unsigned int b) { return xr[a] != xr[b] ? xr[a] < xr[b] : xs[a] < xs[b]; });
  std::vector<long long> ex_ptr(static_cast<size_t>(h->n_head) + 1, 0);
  std::vector<int> es(n_ex); std::vector<uint32_t> eg(n_ex); std::vector<unsigned long long> eq(n_ex);
  for (unsigned int i = 0; i < n_ex; ++i) { const unsigned int o = order[i]; ex_ptr[xr[o] + 1]++; es[i] = xs[o]; eg[i] = xg[o]; eq[i] = xq[o]; }
  for (int r = 0; r < h->n_head; ++r) ex_ptr[r + 1] += ex_ptr[r];
  if ((rc = slot_upload(h, mr_handle::SL_EX_PTR, &h->d_ex_ptr, ex_ptr.data(), ex_ptr.size()))) return rc;
  if ((rc = slot_upload(h, mr_handle::SL_EX_SONG, &h->d_ex_song, es.data(), es.size()))) return rc;
  if ((rc = slot_upload(h, mr_handle::SL_EX_G, &h->d_ex_g, eg.data(), eg.size()))) return rc;
  if ((rc = slot_upload(h, mr_handle::SL_EX_GQ, &h->d_ex_gq, eq.data(), eq.size()))) return rc;
  MR_CUDA(h, cudaStreamSynchronize(h->stream));
  h->n_ex = n_ex;
  h->head_ready = true;
  return MR_OK;
}

int ensure_head_rows(mr_handle* h) {
  if (h->head_ready) return MR_OK;
  int rc = start_head_rows(h, h->stream);
  if (rc) return rc;
  return finish_head_rows(h);
}

// Batch plan of the current test shard.  User space: 128-user batches (UMMA M).  Item space: as many test users per batch as the Sint
// panels (two models, 8 bytes per (user, song)) fit in HBM next to the head rows — a large batch shares the row tiles of popular
// songs among more users (head_rowsum_kernel) — split evenly, plus per batch the balanced work groups of head_rowsum_kernel:
// segments (Sint row, head-entry range) packed longest-first into n_groups bins of equal total length.
constexpr int kPlanTooLarge = -100;   // a work group exceeds the shared-memory staging area: the caller retries with smaller batches

int plan_item_batches_once(mr_handle* h, const std::vector<long long>& hu_ptr, int rows_cap) {
  int rc;
  const int U = h->U;
  h->batch_rows = kUserBatch;
  const size_t panel_rows_have = std::min(h->slot_cap[mr_handle::SL_SINT_U], h->slot_cap[mr_handle::SL_SINT_I]) / (static_cast<size_t>(h->spitch) * 8);
  if (h->space == MR_SPACE_ITEM && static_cast<size_t>(U) <= panel_rows_have && (h->item_batch_cap <= 0 || U <= h->item_batch_cap) &&
      (rows_cap <= 0 || U <= rows_cap) && h->slot_cap[mr_handle::SL_SEL] >= static_cast<size_t>(U) * h->sel_pitch * 8) {
    h->batch_rows = std::max(kUserBatch, U);   // the panels of an earlier shard already hold this one as a single batch: no memory query
  } else if (h->space == MR_SPACE_ITEM) {
    size_t free_b = 0, total_b = 0;
    MR_CUDA(h, cudaMemGetInfo(&free_b, &total_b));
    const size_t per_row = static_cast<size_t>(h->spitch) * 16 + static_cast<size_t>(h->sel_pitch) * 8;
    size_t avail = free_b + h->slot_cap[mr_handle::SL_SINT_U] + h->slot_cap[mr_handle::SL_SINT_I] + h->slot_cap[mr_handle::SL_SEL];
    size_t reserve = 6ULL << 30;   // left for the caller's context (torch, NCCL buffers, gathered top-k blocks) and small workspaces
    if (!h->head_ready && h->head_rows_cap == 0) reserve += static_cast<size_t>(std::max(h->n_head, 1)) * h->spitch * 6 + (1ULL << 30);   // head rows + staging come later
    long long max_rows = avail > reserve ? static_cast<long long>((avail - reserve) / per_row) : 0;
    max_rows = std::max<long long>(max_rows, kUserBatch);
    if (h->item_batch_cap > 0) max_rows = std::min<long long>(max_rows, h->item_batch_cap);
    if (rows_cap > 0) max_rows = std::min<long long>(max_rows, std::max(rows_cap, kUserBatch));
    const int n_batches = static_cast<int>((U + max_rows - 1) / max_rows);
    h->batch_rows = std::max(kUserBatch, (U + n_batches - 1) / n_batches);
  }
  const size_t rows = static_cast<size_t>(h->batch_rows);
  if ((rc = slot_alloc(h, mr_handle::SL_SINT_U, &h->d_sint_u, rows * h->spitch, true))) return rc;
  if ((rc = slot_alloc(h, mr_handle::SL_SINT_I, &h->d_sint_i, rows * h->spitch, true))) return rc;
  if ((rc = slot_alloc(h, mr_handle::SL_SEL, &h->d_sel, rows * h->sel_pitch, true))) return rc;
  if (h->space != MR_SPACE_ITEM) return MR_OK;

  const int G = h->n_groups, B = h->batch_rows;
  const int n_batches = (U + B - 1) / B;
  std::vector<int4> seg, grp_hdr, copy_desc; std::vector<int> split_rows;
  long long ge_size = 0;   // entries of the group-ordered copies (segments padded to multiples of 4)
  h->h_split_ptr.assign(1, 0);
  h->seg_cap = 1; h->ent_cap = 1;
  struct Item { int row, e0, e1, acc; };
  std::vector<Item> items; std::vector<std::vector<int>> bins(G);
  for (int bi = 0; bi < n_batches; ++bi) {
    const int b0 = bi * B, nb = std::min(B, U - b0);
    const long long entries = hu_ptr[b0 + nb] - hu_ptr[b0];
    // a segment costs its entries plus ~3 entry-equivalents for the Sint row it writes (8 B/song against 4 or 2 B/song per entry)
    constexpr int kRowCost = 3;
    const long long share = (entries + static_cast<long long>(kRowCost) * nb + G - 1) / G;
    const int cap = static_cast<int>(std::max<long long>(32, share));
    items.clear();
    for (int b = 0; b < nb; ++b) {
      const int e0 = static_cast<int>(hu_ptr[b0 + b]), e1 = static_cast<int>(hu_ptr[b0 + b + 1]), n = e1 - e0;
      if (n <= cap) { items.push_back({b, e0, e1, 0}); continue; }
      const int parts = (n + cap - 1) / cap;
      for (int p = 0; p < parts; ++p) items.push_back({b, e0 + static_cast<int>(static_cast<long long>(n) * p / parts), e0 + static_cast<int>(static_cast<long long>(n) * (p + 1) / parts), 1});
      split_rows.push_back(b);
    }
    h->h_split_ptr.push_back(static_cast<int>(split_rows.size()));
    // Longest-processing-time packing into G bins of equal cost: every item goes to the currently lightest bin.  Costs are small
    // integers and the minimum load never decreases, so the priority queue is a bucket queue (bins listed by load, a cursor that only
    // moves up): O(items + largest load) instead of a heap operation per user, which at 110 000 users per shard was half of
    // mr_set_test_users.  (A plain sequential fill balances the cost model just as well but measured 30 % slower in head_rowsum: LPT
    // mixes long and short segments in every bin, which averages out what the model gets wrong.)
    {   // counting sort by length, descending, stable
      int max_n = 0;
      for (const Item& it : items) max_n = std::max(max_n, it.e1 - it.e0);
      std::vector<int> start(static_cast<size_t>(max_n) + 2, 0);
      for (const Item& it : items) start[max_n - (it.e1 - it.e0) + 1]++;
      for (int i = 0; i <= max_n; ++i) start[i + 1] += start[i];
      std::vector<Item> sorted(items.size());
      for (const Item& it : items) sorted[start[max_n - (it.e1 - it.e0)]++] = it;
      items.swap(sorted);
    }
    for (auto& v : bins) v.clear();
    {
      std::vector<std::vector<int>> by_load(static_cast<size_t>(2 * (share + cap) + 64));
      by_load[0].resize(G);
      for (int g = 0; g < G; ++g) by_load[0][g] = G - 1 - g;      // pop_back hands out bin 0 first
      size_t cur = 0;
      for (size_t i = 0; i < items.size(); ++i) {
        while (by_load[cur].empty()) ++cur;
        const int g = by_load[cur].back(); by_load[cur].pop_back();
        bins[g].push_back(static_cast<int>(i));
        const size_t nl = cur + static_cast<size_t>(items[i].e1 - items[i].e0 + kRowCost);
        if (nl >= by_load.size()) by_load.resize(nl + nl / 2 + 1);
        by_load[nl].push_back(g);
      }
    }
    for (int g = 0; g < G; ++g) {   // the group's segments and entries, contiguous: the CTA stages them in shared memory
      const int seg0 = static_cast<int>(seg.size());
      const long long ent0 = ge_size;
      for (int i : bins[g]) {
        const int local = static_cast<int>(ge_size - ent0), n = items[i].e1 - items[i].e0;
        seg.push_back(make_int4(items[i].row, local, local + n, items[i].acc));
        if (n > 0) copy_desc.push_back(make_int4(items[i].e0, static_cast<int>(ge_size), n, 0));   // gathered on the device
        ge_size += n;
      }
      const int n_seg = static_cast<int>(seg.size()) - seg0, n_ent = static_cast<int>(ge_size - ent0);
      grp_hdr.push_back(make_int4(seg0, n_seg, static_cast<int>(ent0), n_ent));
      h->seg_cap = std::max(h->seg_cap, n_seg); h->ent_cap = std::max(h->ent_cap, n_ent);
    }
  }
  if (ge_size >= (1LL << 31)) return fail(h, MR_ERR_BAD_ARG, "too many head entries in one shard");
  if (static_cast<size_t>(h->seg_cap) * 16 + static_cast<size_t>(h->ent_cap) * 8 > 48 * 1024) {
    if (h->batch_rows > kUserBatch) return kPlanTooLarge;
    return fail(h, MR_ERR_BAD_ARG, "head_rowsum work group too large for shared-memory staging (%d segments, %d entries)", h->seg_cap, h->ent_cap);
  }
  if ((rc = slot_upload(h, mr_handle::SL_SEG, &h->d_seg, seg.data(), seg.size()))) return rc;
  if ((rc = slot_upload(h, mr_handle::SL_GRP_HDR, &h->d_grp_hdr, grp_hdr.data(), grp_hdr.size()))) return rc;
  int4* d_desc = nullptr;
  if ((rc = slot_upload(h, mr_handle::SL_COPY_DESC, &d_desc, copy_desc.data(), copy_desc.size()))) return rc;
  if ((rc = slot_alloc(h, mr_handle::SL_GE_ROW, &h->d_ge_row, static_cast<size_t>(ge_size)))) return rc;
  if ((rc = slot_alloc(h, mr_handle::SL_GE_Q, &h->d_ge_q, static_cast<size_t>(ge_size)))) return rc;
  MR_LAUNCH(h, launch_gather_group_entries(d_desc, static_cast<int>(copy_desc.size()), h->d_hu_row, h->d_hu_q, h->d_ge_row, h->d_ge_q, h->stream));
  if ((rc = slot_upload(h, mr_handle::SL_SPLIT_ROWS, &h->d_split_rows, split_rows.data(), split_rows.size()))) return rc;
  MR_CUDA(h, cudaStreamSynchronize(h->stream));   // the staging vectors above are pageable and go out of scope
  return MR_OK;
}

// A shard with very many head entries per work group (small S with many users per batch, or few groups) is split into more batches
// until a group's segments and entries fit the 48 KB staging area of head_rowsum_kernel.
int plan_item_batches(mr_handle* h, const std::vector<long long>& hu_ptr) {
  int rows_cap = 0;
  for (;;) {
    const int rc = plan_item_batches_once(h, hu_ptr, rows_cap);
    if (rc != kPlanTooLarge) return rc;
    rows_cap = std::max(kUserBatch, h->batch_rows / 2);
  }
}

enum RunMode { RUN_TOPK, RUN_DENSE, RUN_COUNTS_UBM, RUN_SIM_UBM };
struct TopkHostOut { int32_t* song; double* score; int32_t* len; };   // RUN_TOPK: caller buffers the finished slices are copied into
// RUN_DENSE: where the fp64 rows go.  ALL: out[u][s] for every test user (mr_score_dense).  USERS: out[i][s] for the listed users
// (DIST getRanks1 granularity).  SONGS: d_out[i][u] for the listed songs, gathered on the device (DIST getRanks2 granularity).
struct DenseOut { enum Kind { ALL, USERS, SONGS } kind; double* out; const int32_t* ids; int n_ids; const int* d_ids; double* d_out; };

int run_batches(mr_handle* h, int model, const BlendParams& bp, int k, RunMode mode, void* host_out) {
  const bool need_ubm = model != MODEL_IBM;
  const bool need_ibm = model != MODEL_UBM && mode != RUN_COUNTS_UBM && mode != RUN_SIM_UBM;
  const bool item_space = h->space == MR_SPACE_ITEM && (mode == RUN_TOPK || mode == RUN_DENSE);
  if (item_space) { int rc = ensure_head_rows(h); if (rc) return rc; }
  const int batch = item_space ? h->batch_rows : kUserBatch;
  // Batch pipeline (item-space top-k of a pure model over >= 2 batches; not under MR_PROFILE, whose phase timers serialise the
  // stream): the head pass of batch b + 1 is issued on the library stream while the tail scatter, mask and select of batch b run on
  // slice_stream.  A pure model leaves the other model's Sint panel idle, so consecutive batches alternate between the two panels: no
  // extra memory.  Head passes stay in order on h->stream, slices in order on slice_stream; ev_head[p] releases the slices of the batch
  // in panel p, ev_done[p] releases panel p for the head pass two batches later; the last batch joins slice_stream back into h->stream.
  // What it buys is the kernel boundaries (the next kernel's CTAs start while the previous grid drains): 727 -> 701 ms for the two
  // models of the one-GPU job.  Making the kernels truly share the SMs was measured and is WORSE (profiles/r02_summary.md §8): the head
  // pass lives on every resident CTA working on the same song tile out of L2, and a select streaming 48 GB through L2 beside it breaks that.
  const bool pipelined = item_space && mode == RUN_TOPK && (model == MODEL_UBM || model == MODEL_IBM) && !(h->flags & MR_PROFILE) &&
                         h->U > batch && !getenv("MRSCORE_NO_PIPELINE");
  // Every other item-space top-k (one batch, e.g. a GPU's share of a song-partitioned job, or a blend) puts consecutive slices on two
  // alternating streams instead: same idea, the boundaries between the slices' kernels (107.6 -> 105.1 ms for both models of a 1/8 song
  // partition; next to the batch pipeline it adds nothing).  The two streams' kernels have different shared-memory carve-outs and never
  // share an SM; asking for matched carve-outs makes them do so and is slower (profiles/r02_summary.md §8).
  const bool alternating = item_space && mode == RUN_TOPK && !pipelined && !(h->flags & MR_PROFILE) && !getenv("MRSCORE_NO_PIPELINE");
  const bool streamed = pipelined || alternating;   // the slices run off the library stream and are joined back into it
  cudaStream_t const hs = h->stream;
  cudaStream_t const sl[2] = {streamed ? h->slice_stream : h->stream, alternating ? h->slice_stream2 : (streamed ? h->slice_stream : h->stream)};
  for (int b0 = 0; b0 < h->U; b0 += batch) {
    const int nb = std::min(batch, h->U - b0);
    if (item_space) {
      const int models = (need_ubm ? 1 : 0) | (need_ibm ? 2 : 0);
      const int par = pipelined ? ((b0 / batch) & 1) ^ (model == MODEL_IBM ? 1 : 0) : 0;   // panel of this batch: 0 = d_sint_u, 1 = d_sint_i
      // the panels of this batch by role (pipelined: the pure model's panel alternates between the two allocations)
      long long* const panel = par ? h->d_sint_i : h->d_sint_u;
      long long* const pu = pipelined ? (model == MODEL_UBM ? panel : nullptr) : h->d_sint_u;
      long long* const pi = pipelined ? (model == MODEL_IBM ? panel : nullptr) : h->d_sint_i;
      {
        PhaseTimer t(h, MR_T_HEAD_ROWSUM);
        const int bi = b0 / batch;
        const int4* grp = h->d_grp_hdr + static_cast<long long>(bi) * h->n_groups;
        const int sp0 = h->h_split_ptr[bi], n_split = h->h_split_ptr[bi + 1] - sp0;
        // the slices that last used this panel (two batches ago in the batch pipeline, else the previous batch) have left it
        if (streamed && b0 >= (pipelined ? 2 : 1) * batch) MR_CUDA(h, cudaStreamWaitEvent(hs, h->ev_done[par], 0));
        if (need_ubm) {
          if (n_split) MR_LAUNCH(h, launch_zero_rows(h->d_split_rows + sp0, n_split, pu, h->spitch, hs));
          MR_LAUNCH(h, launch_head_rowsum(1, h->head_words_u, h->head_threads, grp, h->n_groups, h->d_seg, h->d_ge_row, h->d_ge_q, h->seg_cap, h->ent_cap, h->d_g16, h->d_gq32,
                                          h->spitch, h->n_cols, pu, h->spitch, hs));
        }
        if (need_ibm) {
          if (n_split) MR_LAUNCH(h, launch_zero_rows(h->d_split_rows + sp0, n_split, pi, h->spitch, hs));
          MR_LAUNCH(h, launch_head_rowsum(2, h->head_words_i, h->head_threads, grp, h->n_groups, h->d_seg, h->d_ge_row, h->d_ge_q, h->seg_cap, h->ent_cap, h->d_g16, h->d_gq32,
                                          h->spitch, h->n_cols, pi, h->spitch, hs));
        }
        if (h->n_ex > 0)
          MR_LAUNCH(h, launch_head_fixup(models, h->d_hu_ptr, h->d_hu_row, h->d_hu_song, h->d_hu_q, b0, nb, h->d_ex_ptr, h->d_ex_song, h->d_ex_g,
                                         h->d_ex_gq, pu, pi, h->spitch, hs));
        if (streamed) {
          MR_CUDA(h, cudaEventRecord(h->ev_head[par], hs));
          MR_CUDA(h, cudaStreamWaitEvent(sl[0], h->ev_head[par], 0));
          if (alternating) MR_CUDA(h, cudaStreamWaitEvent(sl[1], h->ev_head[par], 0));
        }
      }
      // per slice of users so that the atomics of one launch stay within a few GB of the Sint panels (measured optimum: 300-600 users);
      // in top-k mode the slice is masked and selected right away, and its rows of the result go to the caller's buffers on the copy
      // stream while the next slice computes
      // (kTailSubBatch users of whole rows = ~1.8 GB of a panel; the rows of a song partition are shorter, so a slice holds more of them
      // — fewer, larger launches of the tail / mask / select kernels)
      static const int tail_sub_env = getenv("MRSCORE_TAIL_SUB") ? std::max(8, atoi(getenv("MRSCORE_TAIL_SUB"))) : 0;
      const long long sub_fit = static_cast<long long>(kTailSubBatch) * round_up(h->S, 32) / h->spitch / (2 * h->num_sms) * (2 * h->num_sms);
      const int tail_sub = tail_sub_env ? tail_sub_env : static_cast<int>(std::max<long long>(kTailSubBatch, std::min<long long>(sub_fit, 1 << 20)));
      const int tail_lanes = 2LL * h->n_cols >= h->S ? 32 : (4LL * h->n_cols >= h->S ? 16 : 8);
      const bool sel_needed = model == MODEL_AGG || model == MODEL_STOCH;
      const TopkHostOut* ho = mode == RUN_TOPK ? static_cast<const TopkHostOut*>(host_out) : nullptr;
      for (int s0 = 0; s0 < nb; s0 += tail_sub) {
        const int sn = std::min(tail_sub, nb - s0);
        cudaStream_t const ts = sl[(s0 / tail_sub) & 1];
        {
          PhaseTimer t(h, MR_T_TAIL_SCATTER);
          const long long e0 = h->h_tu_ptr[b0 + s0], e1 = h->h_tu_ptr[b0 + s0 + sn];
          MR_LAUNCH(h, launch_tail_scatter(models, h->d_tu_user, h->d_tu_song, h->d_tu_lptr, e0, e1, h->d_csc_ptr, h->d_csc_idx, h->d_tr_ptr,
                                           h->d_tr_end, h->d_tr_col, h->d_qv, h->d_qd, b0, pu, pi, h->spitch, h->h_tu_lptr[e1] - h->h_tu_lptr[e0], tail_lanes, ts));
        }
        if (mode != RUN_TOPK) continue;
        PhaseTimer t(h, MR_T_TOPK);
        long long* su = need_ubm ? pu + static_cast<long long>(s0) * h->spitch : nullptr;
        long long* si = need_ibm ? pi + static_cast<long long>(s0) * h->spitch : nullptr;
        uint64_t* sel = sel_needed ? h->d_sel + static_cast<long long>(s0) * h->sel_pitch : nullptr;
        const int u0 = b0 + s0;
        MR_LAUNCH(h, launch_mask_listened(h->d_te_ptr, h->d_te_end, h->d_te_col, u0, sn, su, si, h->spitch, ts));
        if (sel_needed) MR_LAUNCH(h, launch_select_bits(bp, h->d_te_ptr, h->d_te_col, u0, sn, h->n_cols, sel, h->sel_pitch, ts));
        MR_LAUNCH(h, launch_topk(bp, h->d_te_ptr, su, si, h->spitch, sel, h->sel_pitch, u0, sn, h->n_cols, h->d_rsa, h->d_rsd, k, h->d_out_song,
                                 h->d_out_score, h->d_out_len, ts));
        if (ho) {
          MR_CUDA(h, cudaEventRecord(h->ev_slice, ts));
          MR_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_slice, 0));
          const size_t o = static_cast<size_t>(u0) * k, n = static_cast<size_t>(sn) * k;
          MR_CUDA(h, cudaMemcpyAsync(ho->song + o, h->d_out_song + o, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->copy_stream));
          MR_CUDA(h, cudaMemcpyAsync(ho->score + o, h->d_out_score + o, n * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
          MR_CUDA(h, cudaMemcpyAsync(ho->len + u0, h->d_out_len + u0, static_cast<size_t>(sn) * sizeof(int32_t), cudaMemcpyDeviceToHost, h->copy_stream));
        }
      }
      if (streamed) {
        if (alternating) {   // the batch's slices end on sl[0]
          MR_CUDA(h, cudaEventRecord(h->ev_join, sl[1]));
          MR_CUDA(h, cudaStreamWaitEvent(sl[0], h->ev_join, 0));
        }
        MR_CUDA(h, cudaEventRecord(h->ev_done[par], sl[0]));
        if (b0 + batch >= h->U) {   // last batch: the library stream joins the slices (of the last two batches in the batch pipeline)
          MR_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_done[par], 0));
          if (pipelined) MR_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_done[par ^ 1], 0));
        }
      }
      if (mode == RUN_TOPK) continue;   // the batch is finished; RUN_DENSE continues below on the whole batch
    } else if (need_ubm) {
      int rc = count_ubm_batch(h, b0, nb);
      if (rc) return rc;
      if (mode == RUN_COUNTS_UBM || mode == RUN_SIM_UBM) {
        PhaseTimer t(h, MR_T_OTHER);
        if (mode == RUN_COUNTS_UBM) {
          ct_to_i32_kernel<<<h->num_sms * 4, 256, 0, h->stream>>>(h->d_ct, h->T, nb, h->d_cnt);
          h->launches++;
          MR_CUDA(h, cudaMemcpyAsync(static_cast<int32_t*>(host_out) + static_cast<long long>(b0) * h->T, h->d_cnt,
                                     static_cast<size_t>(nb) * h->T * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        } else {
          ct_to_cos_kernel<<<h->num_sms * 4, 256, 0, h->stream>>>(h->d_ct, h->T, nb, h->d_rsa_f + b0, h->d_rsv_f, h->d_simf);
          h->launches++;
          MR_CUDA(h, cudaMemcpyAsync(static_cast<float*>(host_out) + static_cast<long long>(b0) * h->T, h->d_simf,
                                     static_cast<size_t>(nb) * h->T * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        }
        MR_CUDA(h, cudaStreamSynchronize(h->stream));
        continue;
      }
      PhaseTimer t(h, MR_T_AGG_UBM);
      MR_CUDA(h, cudaMemsetAsync(h->d_sint_u, 0, static_cast<size_t>(kUserBatch) * h->spitch * sizeof(long long), h->stream));
      AggItems items{h->d_item_song, h->d_item_begin, h->d_item_len, h->d_item_split, h->n_items};
      MR_LAUNCH(h, launch_aggregate_ubm(items, h->d_csc_idx, h->d_qv, h->d_ct, h->T, h->d_sint_u, h->spitch, h->num_sms, h->stream));
    }
    if (item_space) {
      // both models were produced above
    } else if (need_ibm && h->engine == MR_ENGINE_SPARSE) {
      // user-space formulation: weighted intersection counts, then the same inverted-index gather as UBM
      CarryList carry{h->d_carry_count, h->d_carry_events, h->carry_cap};
      {
        PhaseTimer t(h, MR_T_COUNT);
        MR_LAUNCH(h, launch_sparse_wcount_u32(h->d_te_ptr, h->d_te_col, b0, nb, h->d_csc_ptr, h->d_csc_idx, h->d_qd, h->d_wi, h->T, carry, h->stream));
      }
      PhaseTimer t(h, MR_T_AGG_IBM);
      MR_CUDA(h, cudaMemsetAsync(h->d_sint_i, 0, static_cast<size_t>(kUserBatch) * h->spitch * sizeof(long long), h->stream));
      AggItems items{h->d_item_song, h->d_item_begin, h->d_item_len, h->d_item_split, h->n_items};
      MR_LAUNCH(h, launch_aggregate_w32(items, h->d_csc_idx, h->d_wi, h->T, h->d_sint_i, h->spitch, h->num_sms, h->stream));
      MR_LAUNCH(h, launch_carry_fixup(carry, h->d_tr_ptr, h->d_tr_col, h->d_sint_i, h->spitch, h->stream));
      MR_CUDA(h, cudaMemcpyAsync(h->h_carry_seen + (b0 / kUserBatch) % 4096, h->d_carry_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    } else if (need_ibm) {
      const int bi = b0 / kUserBatch;
      const long long r_off = h->batch_row_off[bi];
      const int n_rows = static_cast<int>(h->batch_row_off[bi + 1] - r_off);
      if (n_rows > 0) {
        int rc = gram_rows(h, h->d_rows + r_off, n_rows);
        if (rc) return rc;
      }
      PhaseTimer t(h, MR_T_AGG_IBM);
      MR_LAUNCH(h, launch_aggregate_ibm(h->d_te_ptr, h->d_te_col, h->d_te_grow, h->d_qd, b0, nb, h->d_g, h->ldg, h->S, h->d_sint_i,
                                        h->spitch, h->stream));
    }
    PhaseTimer t(h, MR_T_TOPK);
    MR_LAUNCH(h, launch_mask_listened(h->d_te_ptr, h->d_te_end, h->d_te_col, b0, nb, need_ubm ? h->d_sint_u : nullptr,
                                      need_ibm ? h->d_sint_i : nullptr, h->spitch, h->stream));
    if (mode == RUN_DENSE) {
      for (int c0 = 0; c0 < nb; c0 += kDenseChunk) {
        const int cn = std::min(kDenseChunk, nb - c0);
        MR_LAUNCH(h, launch_dense_scores(model, (need_ubm ? h->d_sint_u : h->d_sint_i) + static_cast<long long>(c0) * h->spitch, h->spitch, b0 + c0, cn,
                                         h->n_cols, h->d_rsa, h->d_rsd, h->d_dense, h->stream));
        const DenseOut* dout = static_cast<const DenseOut*>(host_out);
        if (dout->kind == DenseOut::ALL) {
          MR_CUDA(h, cudaMemcpyAsync(dout->out + static_cast<long long>(b0 + c0) * h->n_cols, h->d_dense,
                                     static_cast<size_t>(cn) * h->n_cols * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        } else if (dout->kind == DenseOut::USERS) {
          for (int i = 0; i < dout->n_ids; ++i) {
            const int r = dout->ids[i] - (b0 + c0);
            if (r < 0 || r >= cn) continue;
            MR_CUDA(h, cudaMemcpyAsync(dout->out + static_cast<long long>(i) * h->n_cols, h->d_dense + static_cast<long long>(r) * h->n_cols,
                                       static_cast<size_t>(h->n_cols) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
          }
        } else {
          MR_LAUNCH(h, launch_gather_columns(h->d_dense, cn, h->n_cols, dout->d_ids, dout->n_ids, dout->d_out, h->U, b0 + c0, h->stream));
        }
        MR_CUDA(h, cudaStreamSynchronize(h->stream));
      }
    } else {
      if (model == MODEL_AGG || model == MODEL_STOCH)
        MR_LAUNCH(h, launch_select_bits(bp, h->d_te_ptr, h->d_te_col, b0, nb, h->n_cols, h->d_sel, h->sel_pitch, h->stream));
      MR_LAUNCH(h, launch_topk(bp, h->d_te_ptr, need_ubm ? h->d_sint_u : nullptr, need_ibm ? h->d_sint_i : nullptr, h->spitch,
                               (model == MODEL_AGG || model == MODEL_STOCH) ? h->d_sel : nullptr, h->sel_pitch, b0, nb, h->n_cols, h->d_rsa,
                               h->d_rsd, k, h->d_out_song, h->d_out_score, h->d_out_len, h->stream));
    }
  }
  if (need_ibm && !item_space && h->engine == MR_ENGINE_SPARSE) {
    // the carry list of the u32 weighted-count panel must not have overflowed (it never does on real data: an event needs
    // more than ~64 songs shared between one test user and one train user)
    MR_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 4096; ++i)
      if (h->h_carry_seen[i] > h->carry_cap)
        return fail(h, MR_ERR_OOM, "IBM weighted-count carry list overflowed (%u events > capacity %u)", h->h_carry_seen[i], h->carry_cap);
  }
  return MR_OK;
}

int make_blend_params(mr_handle* h, int model, double param, uint64_t seed, long long n_total, BlendParams* bp) {
  memset(bp, 0, sizeof *bp);
  bp->model = model;
  bp->ubm_int_ok = h->ubm_int_ok ? 1 : 0;
  bp->rsd_up = h->d_rsd_up;
  bp->te_end = h->d_te_end; bp->song_off = h->win_lo; bp->stats = h->d_topk_stats;
  if (model < MR_UBM || model > MR_STOCH) return fail(h, MR_ERR_BAD_ARG, "unknown model selector %d", model);
  if (model == MR_LC) { bp->alpha = param; bp->one_minus_alpha = 1 - param; }   // rank1 * alpha + rank2 * (1 - alpha), MR:328
  if (model == MR_AGG) {
    if (!(param >= 0 && param <= 1)) return fail(h, MR_ERR_PARAM_RANGE, "Percentage must be between 0 and 1");   // MR:366-369 (NaN fails too)
    bp->agg_threshold = static_cast<long long>(param * static_cast<double>(n_total));                        // MR:372 .toInt
  }
  if (model == MR_STOCH) {
    if (!(param >= 0 && param <= 1)) return fail(h, MR_ERR_PARAM_RANGE, "Probability must be between 0 and 1");   // MR:434-437
    bp->prob = param; bp->seed = seed;
  }
  return MR_OK;
}

}  // namespace

// =====================================================================================================================
#pragma GCC visibility push(default)
extern "C" {

int mr_create(mr_handle** out, const int* device_ids, int n_devices, unsigned flags) {
  if (!out) return MR_ERR_BAD_ARG;
  *out = nullptr;
  mr_handle* h = new mr_handle();
  *out = h;   // returned even on failure so that mr_last_error works; caller still calls mr_destroy
  if (n_devices != 1 || !device_ids) return fail(h, MR_ERR_BAD_ARG, "n_devices must be 1 (one process per GPU), got %d", n_devices);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(h, MR_ERR_CUDA, "no CUDA device: %s — libmrscore has no CPU fallback", cudaGetErrorString(e));
  h->device = device_ids[0];
  if (h->device < 0 || h->device >= count) return fail(h, MR_ERR_BAD_ARG, "device id %d out of range (%d devices)", h->device, count);
  MR_CUDA(h, cudaSetDevice(h->device));
  cudaDeviceProp prop;
  MR_CUDA(h, cudaGetDeviceProperties(&prop, h->device));
  if (prop.major != 10) return fail(h, MR_ERR_CUDA, "device %d is sm_%d%d; libmrscore is built for sm_100a only", h->device, prop.major, prop.minor);
  h->num_sms = prop.multiProcessorCount;
  h->flags = flags;
  h->engine = flags & MR_ENGINE_MASK;
  h->space_flag = flags & MR_SPACE_MASK;
  MR_CUDA(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  MR_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  MR_CUDA(h, cudaEventCreateWithFlags(&h->ev_slice, cudaEventDisableTiming));
  MR_CUDA(h, cudaStreamCreateWithFlags(&h->pre_stream, cudaStreamNonBlocking));
  MR_CUDA(h, cudaEventCreateWithFlags(&h->ev_pre, cudaEventDisableTiming));
  for (int i = 0; i < kPreExtraStreams; ++i) {
    MR_CUDA(h, cudaStreamCreateWithFlags(&h->pre_extra[i], cudaStreamNonBlocking));
    MR_CUDA(h, cudaEventCreateWithFlags(&h->ev_pre_extra[i], cudaEventDisableTiming));
  }
  {
    int prio_lo = 0, prio_hi = 0;
    MR_CUDA(h, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));   // numerically lower = higher priority; `stream` has the default (lowest)
    MR_CUDA(h, cudaStreamCreateWithPriority(&h->slice_stream, cudaStreamNonBlocking, prio_hi));
    MR_CUDA(h, cudaStreamCreateWithPriority(&h->slice_stream2, cudaStreamNonBlocking, prio_hi));
    MR_CUDA(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
      MR_CUDA(h, cudaEventCreateWithFlags(&h->ev_head[i], cudaEventDisableTiming));
      MR_CUDA(h, cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
    }
  }
  MR_CUDA(h, cudaMallocHost(reinterpret_cast<void**>(&h->h_n_ex), sizeof(unsigned int)));
  *h->h_n_ex = 0;
  MR_CUDA(h, cudaEventCreate(&h->ev[0]));
  MR_CUDA(h, cudaEventCreate(&h->ev[1]));
  return MR_OK;
}

void mr_destroy(mr_handle* h) {
  if (!h) return;
  if (h->stream) { cudaSetDevice(h->device); cudaStreamSynchronize(h->stream); }
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  if (h->ev_slice) cudaEventDestroy(h->ev_slice);
  if (h->pre_stream) { cudaStreamSynchronize(h->pre_stream); cudaStreamDestroy(h->pre_stream); }
  if (h->ev_pre) cudaEventDestroy(h->ev_pre);
  for (int i = 0; i < kPreExtraStreams; ++i) {
    if (h->pre_extra[i]) { cudaStreamSynchronize(h->pre_extra[i]); cudaStreamDestroy(h->pre_extra[i]); }
    if (h->ev_pre_extra[i]) cudaEventDestroy(h->ev_pre_extra[i]);
  }
  if (h->slice_stream) { cudaStreamSynchronize(h->slice_stream); cudaStreamDestroy(h->slice_stream); }
  if (h->slice_stream2) { cudaStreamSynchronize(h->slice_stream2); cudaStreamDestroy(h->slice_stream2); }
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  for (int i = 0; i < 2; ++i) { if (h->ev_head[i]) cudaEventDestroy(h->ev_head[i]); if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]); }
  if (h->h_n_ex) cudaFreeHost(h->h_n_ex);
  for (int i = 0; i < mr_handle::SL_N; ++i) if (h->slot_p[i]) cudaFree(h->slot_p[i]);
  if (h->d_g16) cudaFree(h->d_g16);
  if (h->d_gq32) cudaFree(h->d_gq32);
  free_list(h->allocs);
  if (h->h_carry_seen) cudaFreeHost(h->h_carry_seen);
  if (h->ev[0]) cudaEventDestroy(h->ev[0]);
  if (h->ev[1]) cudaEventDestroy(h->ev[1]);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* mr_last_error(const mr_handle* h) { return h ? h->err.c_str() : "null handle"; }

int mr_load(mr_handle* h, int n_train, int n_test, int n_songs, const int64_t* tr_rowptr, const int32_t* tr_col,
            const int64_t* te_rowptr, const int32_t* te_col, const int32_t* deg_train, const int32_t* deg_test,
            const int32_t* deg_song_all) {
  if (!h || !h->stream) return MR_ERR_STATE;
  if (h->loaded) return fail(h, MR_ERR_STATE, "mr_load called twice on one handle");
  if (n_train <= 0 || n_songs <= 0 || n_test < 0 || !deg_train || !deg_song_all) return fail(h, MR_ERR_BAD_ARG, "bad sizes / null degree arrays");
  int rc = check_csr(h, "train", n_train, n_songs, tr_rowptr, tr_col);
  if (rc) return rc;
  // The fixed-point cosine factors q_k(deg) = rint(2^k / sqrt(deg)) carry a relative rounding error of up to 0.5 * sqrt(deg) / 2^k
  // (DESIGN.md §3); beyond these degrees it would exceed the 1e-5 the scores are specified to (north_star): refuse instead of degrading.
  for (int v = 0; v < n_train; ++v)
    if (deg_train[v] > kMaxDegUbm) return fail(h, MR_ERR_BAD_ARG, "train user %d has %d songs: above %d the 2^24 fixed-point UBM weight leaves the 1e-5 score tolerance", v, deg_train[v], kMaxDegUbm);
  for (int s = 0; s < n_songs; ++s)
    if (deg_song_all[s] > kMaxDegIbm) return fail(h, MR_ERR_BAD_ARG, "song %d has %d listeners: above %d the 2^26 fixed-point IBM weight leaves the 1e-5 score tolerance", s, deg_song_all[s], kMaxDegIbm);
  MR_CUDA(h, cudaSetDevice(h->device));
  const int T = n_train, S = n_songs;
  const long long nnz = tr_rowptr[T];
  h->T = T; h->S = S; h->nnz_tr = nnz;
  // song window: rotate the song ids so that the window becomes the columns [0, n_cols) (see mr_handle)
  if (h->win_hi > S || h->win_lo >= (h->win_hi ? h->win_hi : S)) return fail(h, MR_ERR_BAD_ARG, "song window [%d,%d) outside [0,%d)", h->win_lo, h->win_hi, S);
  if (h->win_hi == 0) h->win_hi = S;
  h->windowed = h->win_lo > 0 || h->win_hi < S;
  h->n_cols = h->win_hi - h->win_lo;
  std::vector<int32_t> rot_col, rot_deg; std::vector<long long> tr_wend;
  if (h->windowed) {
    // the windowed path exists in the item-space formulation with inverted-index counts only
    if (h->engine == MR_ENGINE_TENSOR || h->space_flag == MR_SPACE_USER)
      return fail(h, MR_ERR_BAD_ARG, "a song window needs MR_ENGINE_SPARSE / MR_SPACE_ITEM (or AUTO)");
    h->engine = MR_ENGINE_SPARSE; h->space_flag = MR_SPACE_ITEM;
    const int lo = h->win_lo, up = S - lo;
    rot_col.resize(static_cast<size_t>(std::max<long long>(nnz, 1))); tr_wend.resize(T);
    for (int v = 0; v < T; ++v) {
      const int32_t* b = tr_col + tr_rowptr[v]; const int32_t* e = tr_col + tr_rowptr[v + 1];
      const int32_t* m = std::lower_bound(b, e, lo);             // [b, m) < lo <= [m, e)
      int32_t* o = rot_col.data() + tr_rowptr[v];
      for (const int32_t* p = m; p < e; ++p) *o++ = *p - lo;
      for (const int32_t* p = b; p < m; ++p) *o++ = *p + up;
      tr_wend[v] = tr_rowptr[v] + (std::lower_bound(m, e, h->win_hi) - m);
    }
    rot_deg.resize(S);
    for (int s = 0; s < S; ++s) rot_deg[s < lo ? s + up : s - lo] = deg_song_all[s];
    tr_col = rot_col.data(); deg_song_all = rot_deg.data();
  }
  h->pitchS = round_up(S, 128); h->pitchT = round_up(T, 128);
  h->spitch = round_up(h->n_cols, 32); h->ldg = round_up(S, 32);
  h->deg_song.assign(deg_song_all, deg_song_all + S);
  h->deg_song_train.assign(S, 0);
  for (long long i = 0; i < nnz; ++i) h->deg_song_train[tr_col[i]]++;

  // inverted index (counting sort; listeners of a song ascending because rows are visited in order)
  std::vector<long long> csc_ptr(static_cast<size_t>(S) + 1, 0);
  for (long long i = 0; i < nnz; ++i) csc_ptr[tr_col[i] + 1]++;
  for (int s = 0; s < S; ++s) csc_ptr[s + 1] += csc_ptr[s];
  std::vector<int> csc_idx(static_cast<size_t>(std::max<long long>(nnz, 1)));
  {
    std::vector<long long> fill(csc_ptr.begin(), csc_ptr.end() - 1);
    for (int v = 0; v < T; ++v)
      for (long long i = tr_rowptr[v]; i < tr_rowptr[v + 1]; ++i) csc_idx[fill[tr_col[i]]++] = v;
  }
  // fixed-point cosine factors
  std::vector<uint32_t> qv(T), qd(S);
  std::vector<double> rsd(S);
  std::vector<float> rsvf(T), rsdf(S);
  for (int v = 0; v < T; ++v) { qv[v] = q_of(deg_train[v], kQScaleUbm); rsvf[v] = rsf_of(deg_train[v]); h->max_qv = std::max(h->max_qv, qv[v]); }
  for (int s = 0; s < S; ++s) { qd[s] = q_of(deg_song_all[s], kQScaleIbm); rsd[s] = rs_of(deg_song_all[s], kQInvIbm); rsdf[s] = rsf_of(deg_song_all[s]); }
  h->song_qsum.assign(S, 0);
  for (int s = 0; s < S; ++s)
    for (long long i = csc_ptr[s]; i < csc_ptr[s + 1]; ++i) h->song_qsum[s] += qv[csc_idx[i]];
  // K2 work items: <= kSplitLen listeners each, longest first
  struct Item { int song; long long begin; int len; uint8_t split; };
  std::vector<Item> items;
  for (int s = 0; s < S; ++s) {
    const long long len = csc_ptr[s + 1] - csc_ptr[s];
    for (long long o = 0; o < len; o += kSplitLen)
      items.push_back({s, csc_ptr[s] + o, static_cast<int>(std::min<long long>(kSplitLen, len - o)), static_cast<uint8_t>(len > kSplitLen)});
  }
  std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.len > b.len; });
  h->n_items = static_cast<int>(items.size());
  std::vector<int> it_song(items.size()), it_len(items.size());
  std::vector<long long> it_begin(items.size());
  std::vector<uint8_t> it_split(items.size());
  for (size_t i = 0; i < items.size(); ++i) { it_song[i] = items[i].song; it_begin[i] = items[i].begin; it_len[i] = items[i].len; it_split[i] = items[i].split; }

  std::vector<long long> trp(tr_rowptr, tr_rowptr + T + 1);
  if ((rc = dev_upload(h, &h->d_tr_ptr, trp.data(), trp.size(), h->allocs))) return rc;
  if ((rc = dev_upload(h, &h->d_tr_col, tr_col, static_cast<size_t>(nnz), h->allocs))) return rc;
  if ((rc = dev_alloc(h, &h->d_topk_stats, 4, h->allocs))) return rc;
  MR_CUDA(h, cudaMemsetAsync(h->d_topk_stats, 0, 4 * sizeof(unsigned int), h->stream));
  h->d_tr_end = h->d_tr_ptr + 1;
  if (h->windowed) {
    long long* d_end = nullptr;
    if ((rc = dev_upload(h, &d_end, tr_wend.data(), tr_wend.size(), h->allocs))) return rc;
    h->d_tr_end = d_end;
  }
  if ((rc = dev_upload(h, &h->d_csc_ptr, csc_ptr.data(), csc_ptr.size(), h->allocs))) return rc;
  if ((rc = dev_upload(h, &h->d_csc_idx, csc_idx.data(), static_cast<size_t>(nnz), h->allocs))) return rc;
  if ((rc = dev_upload(h, &h->d_qv, qv.data(), qv.size(), h->allocs))) return rc;
  if ((rc = dev_upload(h, &h->d_qd, qd.data(), qd.size(), h->allocs))) return rc;
  if ((rc = dev_upload(h, &h->d_rsd, rsd.data(), rsd.size(), h->allocs))) return rc;
  {  // rsd rounded UP to fp32: the top-k select bounds IBM scores from above with it
    std::vector<float> up(S);
    for (int s = 0; s < S; ++s) { float f = static_cast<float>(rsd[s]); if (static_cast<double>(f) < rsd[s]) f = std::nextafterf(f, INFINITY); up[s] = f; }
    if ((rc = dev_upload(h, &h->d_rsd_up, up.data(), up.size(), h->allocs))) return rc;
    MR_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  if ((rc = dev_upload(h, &h->d_rsv_f, rsvf.data(), rsvf.size(), h->allocs))) return rc;
  if ((rc = dev_upload(h, &h->d_rsd_f, rsdf.data(), rsdf.size(), h->allocs))) return rc;
  if ((rc = dev_upload(h, &h->d_item_song, it_song.data(), it_song.size(), h->allocs))) return rc;
  if ((rc = dev_upload(h, &h->d_item_begin, it_begin.data(), it_begin.size(), h->allocs))) return rc;
  if ((rc = dev_upload(h, &h->d_item_len, it_len.data(), it_len.size(), h->allocs))) return rc;
  if ((rc = dev_upload(h, &h->d_item_split, it_split.data(), it_split.size(), h->allocs))) return rc;
  MR_CUDA(h, cudaStreamSynchronize(h->stream));   // host staging vectors go out of scope below

  // engine choice: dense operands A_tr (T x pitchS) and A_tr^T (S x pitchT) must fit comfortably
  h->dense_bytes = static_cast<size_t>(T) * h->pitchS + static_cast<size_t>(S) * h->pitchT;
  if (h->engine == MR_ENGINE_AUTO) {
    size_t free_b = 0, total_b = 0;
    MR_CUDA(h, cudaMemGetInfo(&free_b, &total_b));
    h->engine = (h->dense_bytes < free_b / 3) ? MR_ENGINE_TENSOR : MR_ENGINE_SPARSE;
  }
  if (h->engine == MR_ENGINE_TENSOR) {
    if ((rc = dev_alloc(h, &h->d_Atr, static_cast<size_t>(T) * h->pitchS, h->allocs))) return rc;
    if ((rc = dev_alloc(h, &h->d_AtrT, static_cast<size_t>(S) * h->pitchT, h->allocs))) return rc;
    if ((rc = dev_alloc(h, &h->d_Ate, static_cast<size_t>(kUserBatch) * h->pitchS, h->allocs))) return rc;
    PhaseTimer t(h, MR_T_EXPAND);
    MR_LAUNCH(h, launch_expand_rows(h->d_tr_ptr, h->d_tr_col, nullptr, 0, T, T, h->pitchS, h->d_Atr, h->stream));
    MR_LAUNCH(h, launch_expand_rows(h->d_csc_ptr, h->d_csc_idx, nullptr, 0, S, S, h->pitchT, h->d_AtrT, h->stream));
  } else {
    h->dense_bytes = 0;
  }
  // per-batch workspaces
  if ((rc = dev_alloc(h, &h->d_ct, static_cast<size_t>(T) * kUserBatch, h->allocs))) return rc;
  if (h->engine == MR_ENGINE_SPARSE) {
    if ((rc = dev_alloc(h, &h->d_wi, static_cast<size_t>(T) * kUserBatch, h->allocs))) return rc;
    if ((rc = dev_alloc(h, &h->d_carry_count, 1, h->allocs))) return rc;
    if ((rc = dev_alloc(h, &h->d_carry_events, h->carry_cap, h->allocs))) return rc;
    MR_CUDA(h, cudaMallocHost(reinterpret_cast<void**>(&h->h_carry_seen), 4096 * sizeof(unsigned int)));
    memset(h->h_carry_seen, 0, 4096 * sizeof(unsigned int));
  }
  if (const char* e = getenv("MRSCORE_ITEM_BATCH")) h->item_batch_cap = std::max(128, atoi(e));   // developer override of MR_OPT_ITEM_BATCH
  if (const char* e = getenv("MRSCORE_HEAD_WORDS_U")) { const int w = atoi(e); if (w == 1 || w == 2 || w == 4) h->head_words_u = w; }
  if (const char* e = getenv("MRSCORE_HEAD_WORDS_I")) { const int w = atoi(e); if (w == 1 || w == 2 || w == 4) h->head_words_i = w; }
  if (const char* e = getenv("MRSCORE_HEAD_THREADS")) { const int t = atoi(e); if (t == 32 || t == 64 || t == 128 || t == 256) h->head_threads = t; }
  h->n_groups = h->num_sms * std::min(32, kHeadCtasPerSm * 256 / h->head_threads);   // one wave of resident CTAs = one song tile for the whole batch
  if (const char* e = getenv("MRSCORE_HEAD_GROUPS")) h->n_groups = std::max(1, atoi(e));   // test aid: few groups -> several users per group on small shards
  h->sel_pitch = (h->n_cols + 63) / 64;
  // item-space head: songs with enough train listeners that a dense precomputed row beats expanding them per test user
  {
    long long min_deg = std::max<long long>(2, S / 6000);   // below ~64 listeners expanding a song on the fly is cheaper than streaming its row
    if (h->opt_head_min_deg > 0) min_deg = h->opt_head_min_deg;
    if (const char* e = getenv("MRSCORE_HEAD_MIN_DEG")) min_deg = std::max(1LL, atoll(e));
    size_t free_b = 0, total_b = 0;
    MR_CUDA(h, cudaMemGetInfo(&free_b, &total_b));
    const long long max_rows = static_cast<long long>((free_b / 100 * 47) / (static_cast<size_t>(h->spitch) * 6));   // packed rows: 6 B per entry; the rest holds the Sint panels
    std::vector<int> order(S);
    for (int s = 0; s < S; ++s) order[s] = s;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return csc_ptr[a + 1] - csc_ptr[a] > csc_ptr[b + 1] - csc_ptr[b]; });
    h->head_index.assign(S, -1);
    std::vector<int> head_song; std::vector<long long> lst_ptr(1, 0);
    for (int r = 0; r < S && r < max_rows; ++r) {
      const int s = order[r]; const long long d = csc_ptr[s + 1] - csc_ptr[s];
      if (d < min_deg) break;
      h->head_index[s] = static_cast<int>(head_song.size());
      head_song.push_back(s); lst_ptr.push_back(lst_ptr.back() + d);
    }
    h->n_head = static_cast<int>(head_song.size());
    // rows from n_head_staged on can be built in place: no entry of theirs can overflow 16 / 32 bits (ensure_head_rows)
    h->n_head_staged = h->n_head;
    while (h->n_head_staged > 0) {
      const int s = head_song[h->n_head_staged - 1];
      if (h->song_qsum[s] >= (1ULL << 32) || csc_ptr[s + 1] - csc_ptr[s] >= 65536) break;
      --h->n_head_staged;
    }
    if (getenv("MRSCORE_PRECOMPUTE_STAGE_ALL")) h->n_head_staged = h->n_head;
    h->song_info.resize(S);
    for (int s = 0; s < S; ++s) h->song_info[s] = {h->head_index[s], h->head_index[s] >= 0 ? qd[s] : static_cast<uint32_t>(h->deg_song_train[s])};
    h->max_qsum = *std::max_element(h->song_qsum.begin(), h->song_qsum.end());
    static_assert(sizeof(mr_handle::SongInfo) == sizeof(int2), "SongInfo is uploaded as int2");
    if ((rc = dev_upload(h, &h->d_song_info, reinterpret_cast<const int2*>(h->song_info.data()), h->song_info.size(), h->allocs))) return rc;
    if ((rc = dev_upload(h, &h->d_head_song, head_song.data(), head_song.size(), h->allocs))) return rc;
    if ((rc = dev_upload(h, &h->d_head_lst_ptr, lst_ptr.data(), lst_ptr.size(), h->allocs))) return rc;
  }
  MR_CUDA(h, cudaStreamSynchronize(h->stream));
  h->loaded = true;
  if (n_test > 0) return mr_set_test_users(h, n_test, te_rowptr, te_col, deg_test, 0, 0);
  return MR_OK;
}

int mr_set_test_users(mr_handle* h, int n_test, const int64_t* te_rowptr, const int32_t* te_col, const int32_t* deg_test,
                      int64_t pair_index_base, int64_t n_pairs_total) {
  if (!h || !h->loaded) return fail(h, MR_ERR_STATE, "mr_load has not succeeded on this handle");
  if (n_test <= 0 || !deg_test) return fail(h, MR_ERR_BAD_ARG, "n_test must be > 0 and deg_test non-null");
  int rc = check_csr(h, "test", n_test, h->S, te_rowptr, te_col);
  if (rc) return rc;
  for (int u = 0; u < n_test; ++u)
    if (te_rowptr[u + 1] - te_rowptr[u] > 65535) return fail(h, MR_ERR_BAD_ARG, "test user %d has more than 65535 visible songs (u16 count panel)", u);
  const bool dbg = getenv("MRSCORE_DEBUG_TIMING") != nullptr;
  auto t_dbg = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!dbg) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[mrscore] set_test_users: %s %.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_dbg).count());
    t_dbg = now;
  };
  MR_CUDA(h, cudaSetDevice(h->device));
  MR_CUDA(h, cudaStreamSynchronize(h->stream));
  lap("check_csr + sync");
  h->have_test = false; h->have_topk = false; h->out_k = 0;
  const int U = n_test; const long long nnz = te_rowptr[U];
  h->U = U; h->nnz_te = nnz;
  h->h_te_ptr.assign(te_rowptr, te_rowptr + U + 1);
  if (h->windowed) h->h_te_col.resize(static_cast<size_t>(nnz)); else h->h_te_col.assign(te_col, te_col + nnz);   // a window rewrites every row below
  std::vector<double> rsa(U); std::vector<float> rsaf(U);
  std::vector<long long> pair_base(static_cast<size_t>(U) + 1);
  pair_base[0] = pair_index_base;
  parallel_rows(U, [&](int lo, int hi) { for (int u = lo; u < hi; ++u) { rsa[u] = rs_of(deg_test[u], kQInvUbm); rsaf[u] = rsf_of(deg_test[u]); } });
  for (int u = 0; u < U; ++u) pair_base[u + 1] = pair_base[u] + (h->S - (te_rowptr[u + 1] - te_rowptr[u]));   // unlistened songs of u (MR:109)
  h->pair_index_base = pair_index_base;
  h->n_pairs_total = n_pairs_total > 0 ? n_pairs_total : pair_base[U] - pair_index_base;
  std::vector<long long> te_wend;
  if (h->windowed) {
    // rotate the rows like the train rows (mr_load); the pair index of column c of user u in MAIN:57-59 order is
    // pair_base[u] + (c + win_lo) - (listened songs below win_lo) - (listened songs of the window below c): fold the constants into pair_base
    const int lo = h->win_lo, up = h->S - lo;
    te_wend.resize(U);
    int32_t* const rot = h->h_te_col.data();
    const int win_hi = h->win_hi;
    parallel_rows(U, [&](int u_lo, int u_hi) {
      for (int u = u_lo; u < u_hi; ++u) {
        const int32_t* b = te_col + te_rowptr[u]; const int32_t* e = te_col + te_rowptr[u + 1];
        const int32_t* m = std::lower_bound(b, e, lo);
        int32_t* o = rot + te_rowptr[u];
        for (const int32_t* p = m; p < e; ++p) *o++ = *p - lo;
        for (const int32_t* p = b; p < m; ++p) *o++ = *p + up;
        te_wend[u] = te_rowptr[u] + (std::lower_bound(m, e, win_hi) - m);
        pair_base[u] += lo - (m - b);
      }
    });
    te_col = h->h_te_col.data();
  }
  // per batch: sorted union of the visible songs (the Gram rows the batch needs) and each entry's row index in it — only the
  // tensor engine's user-space IBM path consumes them
  const int n_batches = (U + kUserBatch - 1) / kUserBatch;
  std::vector<int> rows_all; std::vector<int> grow(h->engine == MR_ENGINE_TENSOR ? static_cast<size_t>(std::max<long long>(nnz, 1)) : 1);
  h->batch_row_off.assign(static_cast<size_t>(n_batches) + 1, 0);
  h->max_batch_rows = 0;
  if (h->engine == MR_ENGINE_TENSOR) {
    for (int b = 0; b < n_batches; ++b) {
      const long long e0 = te_rowptr[b * kUserBatch], e1 = te_rowptr[std::min(U, (b + 1) * kUserBatch)];
      std::vector<int> uni(te_col + e0, te_col + e1);
      std::sort(uni.begin(), uni.end());
      uni.erase(std::unique(uni.begin(), uni.end()), uni.end());
      for (long long e = e0; e < e1; ++e) grow[e] = static_cast<int>(std::lower_bound(uni.begin(), uni.end(), te_col[e]) - uni.begin());
      rows_all.insert(rows_all.end(), uni.begin(), uni.end());
      h->batch_row_off[b + 1] = static_cast<long long>(rows_all.size());
      h->max_batch_rows = std::max<int>(h->max_batch_rows, static_cast<int>(uni.size()));
    }
  }
  lap("host arrays");
  if ((rc = slot_upload(h, mr_handle::SL_TE_PTR, &h->d_te_ptr, h->h_te_ptr.data(), h->h_te_ptr.size()))) return rc;
  h->d_te_end = h->d_te_ptr + 1;
  if (h->windowed) {
    long long* d_end = nullptr;
    if ((rc = slot_upload(h, mr_handle::SL_TE_END, &d_end, te_wend.data(), te_wend.size()))) return rc;
    h->d_te_end = d_end;
  }
  if ((rc = slot_upload(h, mr_handle::SL_TE_COL, &h->d_te_col, h->h_te_col.data(), static_cast<size_t>(nnz)))) return rc;
  if ((rc = slot_upload(h, mr_handle::SL_TE_GROW, &h->d_te_grow, grow.data(), grow.size()))) return rc;
  if ((rc = slot_upload(h, mr_handle::SL_RSA, &h->d_rsa, rsa.data(), rsa.size()))) return rc;
  if ((rc = slot_upload(h, mr_handle::SL_RSA_F, &h->d_rsa_f, rsaf.data(), rsaf.size()))) return rc;
  if ((rc = slot_upload(h, mr_handle::SL_PAIR_BASE, &h->d_pair_base, pair_base.data(), pair_base.size()))) return rc;
  if ((rc = slot_upload(h, mr_handle::SL_ROWS, &h->d_rows, rows_all.data(), rows_all.size()))) return rc;
  lap("uploads 1");
  {  // item-space work lists (k7_testlists.cu): per user the precomputed head rows it sums, and its tail songs expanded on the fly
    std::vector<long long> hu_ptr(static_cast<size_t>(U) + 1, 0);
    h->h_tu_ptr.assign(static_cast<size_t>(U) + 1, 0);
    const size_t n1 = static_cast<size_t>(nnz) + 1;
    int *d_flag = nullptr, *d_head_pos = nullptr; long long *d_deg = nullptr, *d_lsum = nullptr, *d_tu_ptr = nullptr; char* d_tmp = nullptr;
    const size_t tmp_bytes = test_lists_temp_bytes(nnz);
    if ((rc = slot_alloc(h, mr_handle::SL_L_FLAG, &d_flag, n1)) || (rc = slot_alloc(h, mr_handle::SL_L_HEADPOS, &d_head_pos, n1)) ||
        (rc = slot_alloc(h, mr_handle::SL_L_DEG, &d_deg, n1)) || (rc = slot_alloc(h, mr_handle::SL_L_LSUM, &d_lsum, n1)) ||
        (rc = slot_alloc(h, mr_handle::SL_L_TMP, &d_tmp, tmp_bytes)) || (rc = slot_alloc(h, mr_handle::SL_TU_PTR, &d_tu_ptr, hu_ptr.size())) ||
        (rc = slot_alloc(h, mr_handle::SL_HU_PTR, &h->d_hu_ptr, hu_ptr.size())) || (rc = slot_alloc(h, mr_handle::SL_HU_ROW, &h->d_hu_row, n1)) ||
        (rc = slot_alloc(h, mr_handle::SL_HU_SONG, &h->d_hu_song, n1)) || (rc = slot_alloc(h, mr_handle::SL_HU_Q, &h->d_hu_q, n1)) ||
        (rc = slot_alloc(h, mr_handle::SL_TU_USER, &h->d_tu_user, n1)) || (rc = slot_alloc(h, mr_handle::SL_TU_SONG, &h->d_tu_song, n1)) ||
        (rc = slot_alloc(h, mr_handle::SL_TU_LPTR, &h->d_tu_lptr, n1)))
      return rc;
    MR_LAUNCH(h, launch_build_test_lists(h->d_te_ptr, h->d_te_col, U, nnz, h->d_song_info, d_flag, d_head_pos, d_deg, d_lsum, d_tmp, tmp_bytes, h->d_hu_row,
                                         h->d_hu_song, h->d_hu_q, h->d_hu_ptr, h->d_tu_user, h->d_tu_song, h->d_tu_lptr, d_tu_ptr, h->stream));
    MR_CUDA(h, cudaMemcpyAsync(hu_ptr.data(), h->d_hu_ptr, hu_ptr.size() * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    MR_CUDA(h, cudaMemcpyAsync(h->h_tu_ptr.data(), d_tu_ptr, h->h_tu_ptr.size() * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    long long longest = 0;   // overlaps the device work
    for (int u = 0; u < U; ++u) longest = std::max<long long>(longest, te_rowptr[u + 1] - te_rowptr[u]);
    // Sint_u[u][s] = sum_{j in I_u} Gq[j][s] <= sum_{j in I_u} qsum[j]: below 2^52 the top-k select may rank the UBM integers.  The cheap
    // bound (longest row x largest qsum) decides almost always; otherwise the exact per-user sums do.
    h->ubm_int_ok = static_cast<long double>(h->max_qsum) * static_cast<long double>(longest) < 4503599627370496.0L;
    if (!h->ubm_int_ok) {
      unsigned long long worst = 0;
      for (int u = 0; u < U; ++u) {
        unsigned long long bound = 0;
        for (long long e = te_rowptr[u]; e < te_rowptr[u + 1]; ++e) bound += h->song_qsum[te_col[e]];
        worst = std::max(worst, bound);
      }
      h->ubm_int_ok = worst < (1ULL << 52);
    }
    MR_CUDA(h, cudaStreamSynchronize(h->stream));
    h->n_head_entries = hu_ptr[U]; h->n_tail_entries = h->h_tu_ptr[U];
    h->h_tu_lptr.resize(static_cast<size_t>(h->n_tail_entries) + 1);
    MR_CUDA(h, cudaMemcpyAsync(h->h_tu_lptr.data(), h->d_tu_lptr, h->h_tu_lptr.size() * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    lap("device work lists");
    h->space = h->space_flag == MR_SPACE_AUTO ? (U >= 1024 ? MR_SPACE_ITEM : MR_SPACE_USER) : h->space_flag;
    if ((rc = plan_item_batches(h, hu_ptr))) return rc;
    lap("plan_item_batches");
  }
  MR_CUDA(h, cudaStreamSynchronize(h->stream));
  h->have_test = true;
  return MR_OK;
}

static int require_test(mr_handle* h) {
  if (!h) return MR_ERR_STATE;
  if (!h->loaded || !h->have_test) return fail(h, MR_ERR_STATE, "no data loaded (mr_load / mr_set_test_users)");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return fail(h, MR_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  return MR_OK;
}

int mr_counts_ubm(mr_handle* h, int32_t* out_UxT) {
  int rc = require_test(h);
  if (rc) return rc;
  if (!out_UxT) return fail(h, MR_ERR_BAD_ARG, "null output");
  if ((rc = slot_alloc(h, mr_handle::SL_CNT, &h->d_cnt, static_cast<size_t>(kUserBatch) * h->T))) return rc;
  BlendParams bp; memset(&bp, 0, sizeof bp);
  return run_batches(h, MODEL_UBM, bp, 0, RUN_COUNTS_UBM, out_UxT);
}

int mr_similarity_ubm(mr_handle* h, float* out_UxT) {
  int rc = require_test(h);
  if (rc) return rc;
  if (!out_UxT) return fail(h, MR_ERR_BAD_ARG, "null output");
  if (h->engine == MR_ENGINE_TENSOR) {
    // cosine normalisation fused into the GEMM epilogue: out[b][v] = c / (sqrt|I_u| * sqrt|I_v|)   (MR:147-148)
    if ((rc = slot_alloc(h, mr_handle::SL_SIMF, &h->d_simf, static_cast<size_t>(kUserBatch) * h->T))) return rc;
    for (int b0 = 0; b0 < h->U; b0 += kUserBatch) {
      const int nb = std::min(kUserBatch, h->U - b0);
      MR_LAUNCH(h, launch_expand_rows(h->d_te_ptr, h->d_te_col, nullptr, b0, nb, kUserBatch, h->pitchS, h->d_Ate, h->stream));
      MR_LAUNCH(h, launch_count_gemm(h->d_Ate, kUserBatch, h->d_Atr, h->T, h->pitchS, nb, h->T, EPI_COS_F32, h->d_simf, h->T,
                                     h->d_rsa_f + b0, h->d_rsv_f, h->num_sms, h->stream));
      MR_CUDA(h, cudaMemcpyAsync(out_UxT + static_cast<long long>(b0) * h->T, h->d_simf, static_cast<size_t>(nb) * h->T * sizeof(float),
                                 cudaMemcpyDeviceToHost, h->stream));
      MR_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return MR_OK;
  }
  if ((rc = slot_alloc(h, mr_handle::SL_SIMF, &h->d_simf, static_cast<size_t>(kUserBatch) * h->T))) return rc;
  BlendParams bp; memset(&bp, 0, sizeof bp);
  return run_batches(h, MODEL_UBM, bp, 0, RUN_SIM_UBM, out_UxT);
}

static int gram_range(mr_handle* h, int s0, int s1, int32_t* out_i32, float* out_f32) {
  int rc = require_test(h);
  if (rc == MR_ERR_STATE && h && h->loaded) rc = MR_OK;   // Gram rows need only the train replica
  if (rc) return rc;
  MR_CUDA(h, cudaSetDevice(h->device));
  if (h->windowed) return fail(h, MR_ERR_STATE, "Gram rows are not available on a handle with a song window (its song ids are rotated)");
  if (s0 < 0 || s1 > h->S || s0 > s1) return fail(h, MR_ERR_BAD_ARG, "song range [%d,%d) outside [0,%d)", s0, s1, h->S);
  const int chunk = 1024;
  int* d_ids = nullptr; float* d_f = nullptr;
  std::vector<void*> tmp;
  if ((rc = dev_alloc(h, &d_ids, chunk, tmp))) return rc;
  if (out_f32 && (rc = dev_alloc(h, &d_f, static_cast<size_t>(chunk) * h->S, tmp))) { free_list(tmp); return rc; }
  if ((rc = ensure_gram_ws(h, chunk))) { free_list(tmp); return rc; }
  for (int r0 = s0; r0 < s1; r0 += chunk) {
    const int n = std::min(chunk, s1 - r0);
    iota_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(d_ids, r0, n);
    h->launches++;
    if ((rc = gram_rows(h, d_ids, n))) { free_list(tmp); return rc; }
    cudaError_t e;
    if (out_i32) {
      e = cudaMemcpy2DAsync(out_i32 + static_cast<long long>(r0 - s0) * h->S, static_cast<size_t>(h->S) * 4, h->d_g,
                            static_cast<size_t>(h->ldg) * 4, static_cast<size_t>(h->S) * 4, n, cudaMemcpyDeviceToHost, h->stream);
    } else {
      gram_to_cos_kernel<<<h->num_sms * 4, 256, 0, h->stream>>>(h->d_g, h->ldg, d_ids, n, h->S, h->d_rsd_f, d_f);
      h->launches++;
      e = cudaMemcpyAsync(out_f32 + static_cast<long long>(r0 - s0) * h->S, d_f, static_cast<size_t>(n) * h->S * 4,
                          cudaMemcpyDeviceToHost, h->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { free_list(tmp); return fail(h, MR_ERR_CUDA, "gram rows copy-out: %s", cudaGetErrorString(e)); }
  }
  free_list(tmp);
  return MR_OK;
}

int mr_gram_rows_device(mr_handle* h, int s0, int s1, int32_t** dev_out, int64_t* ld) {
  if (!h || !h->loaded) return fail(h, MR_ERR_STATE, "mr_load has not succeeded on this handle");
  if (!dev_out || !ld) return fail(h, MR_ERR_BAD_ARG, "null output");
  MR_CUDA(h, cudaSetDevice(h->device));
  if (h->windowed) return fail(h, MR_ERR_STATE, "Gram rows are not available on a handle with a song window (its song ids are rotated)");
  if (s0 < 0 || s1 > h->S || s0 >= s1) return fail(h, MR_ERR_BAD_ARG, "song range [%d,%d) outside [0,%d)", s0, s1, h->S);
  const int n = s1 - s0;
  int rc = ensure_gram_ws(h, n);
  if (rc) return rc;
  int* d_ids = nullptr;
  if ((rc = slot_alloc(h, mr_handle::SL_GRAM_IDS, &d_ids, static_cast<size_t>(n)))) return rc;
  iota_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(d_ids, s0, n);
  h->launches++;
  if ((rc = gram_rows(h, d_ids, n))) return rc;
  MR_CUDA(h, cudaStreamSynchronize(h->stream));
  *dev_out = h->d_g; *ld = h->ldg;
  return MR_OK;
}

int mr_peer_alloc(mr_handle* h, uint64_t bytes, void** dev_ptr, unsigned char* handle_out_64) {
  if (!h || !h->stream) return MR_ERR_STATE;
  if (!dev_ptr || !handle_out_64 || bytes == 0) return fail(h, MR_ERR_BAD_ARG, "null output / zero size");
  MR_CUDA(h, cudaSetDevice(h->device));
  void* p = nullptr;
  MR_CUDA(h, cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) { cudaFree(p); return fail(h, MR_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle_out_64, &hd, 64);
  MR_CUDA(h, cudaMemsetAsync(p, 0, bytes, h->stream));
  MR_CUDA(h, cudaStreamSynchronize(h->stream));
  h->allocs.push_back(p);
  h->dev_bytes += bytes;
  *dev_ptr = p;
  return MR_OK;
}

int mr_peer_open(mr_handle* h, const unsigned char* handle_64, void** dev_ptr) {
  if (!h || !h->stream) return MR_ERR_STATE;
  if (!dev_ptr || !handle_64) return fail(h, MR_ERR_BAD_ARG, "null argument");
  MR_CUDA(h, cudaSetDevice(h->device));
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle_64, 64);
  MR_CUDA(h, cudaIpcOpenMemHandle(dev_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
  return MR_OK;
}

int mr_peer_close(mr_handle* h, void* dev_ptr) {
  if (!h || !h->stream) return MR_ERR_STATE;
  MR_CUDA(h, cudaSetDevice(h->device));
  MR_CUDA(h, cudaIpcCloseMemHandle(dev_ptr));
  return MR_OK;
}

static int gram_rows_scatter_impl(mr_handle* h, int s0, int s1, void* const* slot_ptrs, int n_owners, int rows_per_owner, int64_t ld, bool sync);
int mr_gram_rows_scatter(mr_handle* h, int s0, int s1, void* const* slot_ptrs, int n_owners, int rows_per_owner, int64_t ld) {
  return gram_rows_scatter_impl(h, s0, s1, slot_ptrs, n_owners, rows_per_owner, ld, true);
}
int mr_gram_rows_scatter_async(mr_handle* h, int s0, int s1, void* const* slot_ptrs, int n_owners, int rows_per_owner, int64_t ld) {
  return gram_rows_scatter_impl(h, s0, s1, slot_ptrs, n_owners, rows_per_owner, ld, false);
}

int mr_peer_signal(mr_handle* h, void* const* flag_ptrs, int n, uint64_t value) {
  if (!h || !h->stream) return MR_ERR_STATE;
  if (!flag_ptrs || n < 1 || n > 8) return fail(h, MR_ERR_BAD_ARG, "1..8 flag pointers expected");
  MR_CUDA(h, cudaSetDevice(h->device));
  PeerFlags f{};
  for (int i = 0; i < n; ++i) f.p[i] = static_cast<unsigned long long*>(flag_ptrs[i]);
  MR_LAUNCH(h, launch_peer_signal(f, n, value, h->stream));
  return MR_OK;
}

int mr_peer_wait(mr_handle* h, const void* flags, int n, uint64_t value) {
  if (!h || !h->stream) return MR_ERR_STATE;
  if (!flags || n < 1 || n > 8) return fail(h, MR_ERR_BAD_ARG, "1..8 flags expected");
  MR_CUDA(h, cudaSetDevice(h->device));
  MR_LAUNCH(h, launch_peer_wait(static_cast<const unsigned long long*>(flags), n, value, h->stream));
  return MR_OK;
}

int mr_sync(mr_handle* h) {
  if (!h || !h->stream) return MR_ERR_STATE;
  MR_CUDA(h, cudaSetDevice(h->device));
  MR_CUDA(h, cudaStreamSynchronize(h->stream));
  return MR_OK;
}

static int gram_rows_scatter_impl(mr_handle* h, int s0, int s1, void* const* slot_ptrs, int n_owners, int rows_per_owner, int64_t ld, bool sync) {
  if (!h || !h->loaded) return fail(h, MR_ERR_STATE, "mr_load has not succeeded on this handle");
  if (h->engine != MR_ENGINE_TENSOR) return fail(h, MR_ERR_STATE, "mr_gram_rows_scatter needs the tensor engine (the scatter is the GEMM epilogue)");
  MR_CUDA(h, cudaSetDevice(h->device));
  if (h->windowed) return fail(h, MR_ERR_STATE, "Gram rows are not available on a handle with a song window (its song ids are rotated)");
  if (s0 < 0 || s1 > h->S || s0 >= s1) return fail(h, MR_ERR_BAD_ARG, "song range [%d,%d) outside [0,%d)", s0, s1, h->S);
  const int n = s1 - s0;
  if (!slot_ptrs || n_owners < 1 || n_owners > 8 || rows_per_owner < 1 || static_cast<long long>(n_owners) * rows_per_owner < n || ld < h->S)
    return fail(h, MR_ERR_BAD_ARG, "bad slot table (owners %d, rows per owner %d, ld %lld)", n_owners, rows_per_owner, static_cast<long long>(ld));
  const int rows_pad = static_cast<int>(round_up(n, 128));
  const size_t need_a = static_cast<size_t>(rows_pad) * h->pitchT;
  if (need_a > h->aj_bytes) {
    uint8_t* pa; int rc = dev_alloc(h, &pa, need_a, h->allocs);
    if (rc) return rc;
    h->d_Aj = pa; h->aj_bytes = need_a;
  }
  int* d_ids = nullptr;
  int rc;
  if ((rc = slot_alloc(h, mr_handle::SL_GRAM_IDS, &d_ids, static_cast<size_t>(n)))) return rc;
  iota_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(d_ids, s0, n);
  h->launches++;
  {
    PhaseTimer t(h, MR_T_EXPAND);
    MR_LAUNCH(h, launch_expand_rows(h->d_csc_ptr, h->d_csc_idx, d_ids, 0, n, rows_pad, h->pitchT, h->d_Aj, h->stream));
  }
  int32_t* slots[8];
  for (int i = 0; i < n_owners; ++i) slots[i] = static_cast<int32_t*>(slot_ptrs[i]);
  PhaseTimer t(h, MR_T_COUNT);
  MR_LAUNCH(h, launch_count_gemm(h->d_Aj, rows_pad, h->d_AtrT, h->S, h->pitchT, n, h->S, EPI_I32_SCATTER, nullptr, ld, nullptr, nullptr, h->num_sms,
                                 h->stream, 0, 0, slots, n_owners, rows_per_owner));
  if (sync) MR_CUDA(h, cudaStreamSynchronize(h->stream));
  return MR_OK;
}

int mr_counts_ibm(mr_handle* h, int s0, int s1, int32_t* out_rows) {
  if (!out_rows) return fail(h, MR_ERR_BAD_ARG, "null output");
  return gram_range(h, s0, s1, out_rows, nullptr);
}
int mr_similarity_ibm(mr_handle* h, int s0, int s1, float* out_rows) {
  if (!out_rows) return fail(h, MR_ERR_BAD_ARG, "null output");
  return gram_range(h, s0, s1, nullptr, out_rows);
}

int mr_score_dense(mr_handle* h, int model, double* out_UxS) {
  int rc = require_test(h);
  if (rc) return rc;
  if (model != MR_UBM && model != MR_IBM) return fail(h, MR_ERR_BAD_ARG, "mr_score_dense: model must be MR_UBM or MR_IBM");
  if (!out_UxS) return fail(h, MR_ERR_BAD_ARG, "null output");
  if ((rc = slot_alloc(h, mr_handle::SL_DENSE, &h->d_dense, static_cast<size_t>(std::min(h->batch_rows, kDenseChunk)) * h->n_cols))) return rc;
  if (model == MR_IBM && h->engine != MR_ENGINE_SPARSE && (rc = ensure_gram_ws(h, h->max_batch_rows))) return rc;
  BlendParams bp; memset(&bp, 0, sizeof bp); bp.model = model; bp.te_end = h->d_te_end; bp.song_off = h->win_lo;
  DenseOut dout{DenseOut::ALL, out_UxS, nullptr, 0, nullptr, nullptr};
  return run_batches(h, model, bp, 0, RUN_DENSE, &dout);
}

static int score_subset(mr_handle* h, int model, const int32_t* ids, int n, double* out, bool songs) {
  int rc = require_test(h);
  if (rc) return rc;
  if (model != MR_UBM && model != MR_IBM) return fail(h, MR_ERR_BAD_ARG, "model must be MR_UBM or MR_IBM");
  if (n < 0 || (n > 0 && (!ids || !out))) return fail(h, MR_ERR_BAD_ARG, "null id list / output");
  const int first = songs ? h->win_lo : 0, limit = songs ? h->win_lo + h->n_cols : h->U;   // songs: the scored columns (the window, if one is set)
  for (int i = 0; i < n; ++i)
    if (ids[i] < first || ids[i] >= limit) return fail(h, MR_ERR_BAD_ARG, "%s id %d out of range [%d,%d)", songs ? "song" : "test user", ids[i], first, limit);
  if (n == 0) return MR_OK;
  if ((rc = slot_alloc(h, mr_handle::SL_DENSE, &h->d_dense, static_cast<size_t>(std::min(h->batch_rows, kDenseChunk)) * h->n_cols))) return rc;
  if (model == MR_IBM && h->engine != MR_ENGINE_SPARSE && (rc = ensure_gram_ws(h, h->max_batch_rows))) return rc;
  BlendParams bp; memset(&bp, 0, sizeof bp); bp.model = model; bp.te_end = h->d_te_end; bp.song_off = h->win_lo;
  std::vector<void*> tmp;
  DenseOut dout{songs ? DenseOut::SONGS : DenseOut::USERS, out, ids, n, nullptr, nullptr};
  if (songs) {
    int* d_ids = nullptr;
    std::vector<int> cols(ids, ids + n);
    for (int& c : cols) c -= h->win_lo;
    if ((rc = dev_upload(h, &d_ids, cols.data(), static_cast<size_t>(n), tmp)) || (rc = dev_alloc(h, &dout.d_out, static_cast<size_t>(n) * h->U, tmp))) { free_list(tmp); return rc; }
    dout.d_ids = d_ids;
  }
  rc = run_batches(h, model, bp, 0, RUN_DENSE, &dout);
  if (!rc && songs) {
    cudaError_t e = cudaMemcpyAsync(out, dout.d_out, static_cast<size_t>(n) * h->U * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) rc = fail(h, MR_ERR_CUDA, "mr_score_songs copy-out: %s", cudaGetErrorString(e));
  }
  free_list(tmp);
  return rc;
}

int mr_score_users(mr_handle* h, int model, const int32_t* user_idx, int n, double* out_nxS) { return score_subset(h, model, user_idx, n, out_nxS, false); }
int mr_score_songs(mr_handle* h, int model, const int32_t* song_ids, int n, double* out_nxU) { return score_subset(h, model, song_ids, n, out_nxU, true); }

int mr_map_at_k(mr_handle* h, int k, const int32_t* top_song, const int32_t* top_len, int n_users, const int64_t* lab_rowptr, const int32_t* lab_col,
                double* out_map, double* out_ap) {
  if (!h || !h->stream) return MR_ERR_STATE;
  if (k < 1 || n_users <= 0 || !lab_rowptr || !out_map) return fail(h, MR_ERR_BAD_ARG, "null / empty arguments");
  if ((top_song == nullptr) != (top_len == nullptr)) return fail(h, MR_ERR_BAD_ARG, "top_song and top_len must both be given or both be null");
  const long long n_lab = lab_rowptr[n_users];
  if (lab_rowptr[0] != 0 || n_lab < 0 || (n_lab > 0 && !lab_col)) return fail(h, MR_ERR_BAD_ARG, "bad label CSR");
  for (int u = 0; u < n_users; ++u) {
    if (lab_rowptr[u + 1] < lab_rowptr[u]) return fail(h, MR_ERR_BAD_ARG, "label rowptr not monotone at row %d", u);
    for (long long i = lab_rowptr[u] + 1; i < lab_rowptr[u + 1]; ++i)
      if (lab_col[i] <= lab_col[i - 1]) return fail(h, MR_ERR_BAD_ARG, "label row %d not ascending/unique", u);
  }
  MR_CUDA(h, cudaSetDevice(h->device));
  const int* d_song = nullptr; const int* d_len = nullptr;
  std::vector<void*> tmp;
  int rc;
  if (!top_song) {   // rank the device-resident result of the last mr_topk_device / mr_topk
    if (!h->have_topk || h->out_k != k || n_users != h->U) return fail(h, MR_ERR_STATE, "no top-%d result of %d users on the device", k, n_users);
    d_song = h->d_out_song; d_len = h->d_out_len;
  } else {
    int *ds = nullptr, *dl = nullptr;
    if ((rc = dev_upload(h, &ds, top_song, static_cast<size_t>(n_users) * k, tmp)) || (rc = dev_upload(h, &dl, top_len, static_cast<size_t>(n_users), tmp))) { free_list(tmp); return rc; }
    d_song = ds; d_len = dl;
  }
  std::vector<long long> lp(lab_rowptr, lab_rowptr + n_users + 1);
  long long* d_lp = nullptr; int* d_lc = nullptr; double* d_ap = nullptr;
  if ((rc = dev_upload(h, &d_lp, lp.data(), lp.size(), tmp)) || (rc = dev_upload(h, &d_lc, lab_col, static_cast<size_t>(n_lab), tmp)) ||
      (rc = dev_alloc(h, &d_ap, static_cast<size_t>(n_users), tmp))) { free_list(tmp); return rc; }
  int lrc;
  {
    PhaseTimer t(h, MR_T_OTHER);
    lrc = launch_map_at_k(d_song, d_len, n_users, k, d_lp, d_lc, d_ap, h->stream);
    h->launches++;
  }
  std::vector<double> ap(n_users);
  cudaError_t e = lrc ? cudaErrorLaunchFailure : cudaMemcpyAsync(ap.data(), d_ap, static_cast<size_t>(n_users) * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  free_list(tmp);
  if (e != cudaSuccess) return fail(h, MR_ERR_CUDA, "mr_map_at_k: %s", cudaGetErrorString(e));
  double total = 0.0; long long n_eval = 0;
  for (int u = 0; u < n_users; ++u) {
    if (lab_rowptr[u + 1] > lab_rowptr[u]) { total += ap[u]; ++n_eval; }   // left fold in ascending user order, users without labels skipped
    if (out_ap) out_ap[u] = ap[u];
  }
  *out_map = n_eval > 0 ? total / static_cast<double>(n_eval) : 0.0;
  return MR_OK;
}

int mr_set_option(mr_handle* h, int option, int64_t value) {
  if (!h) return MR_ERR_STATE;
  switch (option) {
    case MR_OPT_HEAD_MIN_DEG:
      if (h->loaded) return fail(h, MR_ERR_STATE, "MR_OPT_HEAD_MIN_DEG must be set before mr_load (it selects the head songs)");
      if (value < 0) return fail(h, MR_ERR_BAD_ARG, "MR_OPT_HEAD_MIN_DEG must be >= 0");
      h->opt_head_min_deg = value;
      return MR_OK;
    case MR_OPT_SONG_WINDOW_LO:
    case MR_OPT_SONG_WINDOW_HI:
      if (h->loaded) return fail(h, MR_ERR_STATE, "the song window must be set before mr_load (ids are rotated and the head rows sized by it)");
      if (value < 0 || value > 0x7fffffff) return fail(h, MR_ERR_BAD_ARG, "song window bound out of range");
      (option == MR_OPT_SONG_WINDOW_LO ? h->win_lo : h->win_hi) = static_cast<int>(value);
      return MR_OK;
    case MR_OPT_ITEM_BATCH:
      if (value < 0) return fail(h, MR_ERR_BAD_ARG, "MR_OPT_ITEM_BATCH must be >= 0");
      h->item_batch_cap = value > 0 ? static_cast<int>(std::max<int64_t>(128, std::min<int64_t>(value, 1 << 30))) : 0;
      return MR_OK;
    default:
      return fail(h, MR_ERR_BAD_ARG, "unknown option %d", option);
  }
}

int mr_invalidate_prepared(mr_handle* h) {
  if (!h || !h->loaded) return fail(h, MR_ERR_STATE, "mr_load has not succeeded on this handle");
  cudaError_t e = cudaSetDevice(h->device);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(h, MR_ERR_CUDA, "mr_invalidate_prepared: %s", cudaGetErrorString(e));
  if (h->head_pending) { cudaEventSynchronize(h->ev_pre); h->head_pending = false; }   // a build in flight is abandoned once it has drained
  h->head_ready = false; h->n_ex = 0;
  return MR_OK;
}

int mr_blend_dense(mr_handle* h, int kind, double param, uint64_t seed, const double* ubm, const double* ibm, double* out,
                   int64_t n_pairs, int64_t first_index, int64_t n_total) {
  if (!h || !h->stream) return MR_ERR_STATE;
  if (kind != MR_LC && kind != MR_AGG && kind != MR_STOCH) return fail(h, MR_ERR_BAD_ARG, "mr_blend_dense: kind must be MR_LC, MR_AGG or MR_STOCH");
  if (n_pairs < 0 || (n_pairs > 0 && (!ubm || !ibm || !out))) return fail(h, MR_ERR_BAD_ARG, "null arrays");
  BlendParams bp;
  int rc = make_blend_params(h, kind, param, seed, n_total > 0 ? n_total : n_pairs, &bp);
  if (rc) return rc;
  if (n_pairs == 0) return MR_OK;
  MR_CUDA(h, cudaSetDevice(h->device));
  std::vector<void*> tmp;
  double *du = nullptr, *di = nullptr, *dout = nullptr;
  if ((rc = dev_alloc(h, &du, n_pairs, tmp)) || (rc = dev_alloc(h, &di, n_pairs, tmp)) || (rc = dev_alloc(h, &dout, n_pairs, tmp))) { free_list(tmp); return rc; }
  cudaError_t e = cudaMemcpyAsync(du, ubm, n_pairs * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(di, ibm, n_pairs * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    PhaseTimer t(h, MR_T_OTHER);
    int lrc = launch_blend_arrays(bp, du, di, dout, n_pairs, first_index, h->stream);
    h->launches++;
    if (lrc) e = cudaErrorLaunchFailure;
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, n_pairs * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  free_list(tmp);
  if (e != cudaSuccess) return fail(h, MR_ERR_CUDA, "mr_blend_dense: %s", cudaGetErrorString(e));
  return MR_OK;
}

int mr_evaluate_dense(mr_handle* h, const double* scores_UxS, int n_users, int n_songs, const int64_t* lab_rowptr, const int32_t* lab_col,
                      int n_thresholds, double* out_map) {
  if (!h || !h->stream) return MR_ERR_STATE;
  if (!scores_UxS || !lab_rowptr || !out_map || n_users <= 0 || n_songs <= 0) return fail(h, MR_ERR_BAD_ARG, "null / empty arguments");
  if (n_thresholds != 10 && n_thresholds != 11) return fail(h, MR_ERR_BAD_ARG, "n_thresholds must be 10 (MR:590) or 11 (DIST:395)");
  const long long n_lab = lab_rowptr[n_users];
  if (n_lab <= 0 || !lab_col) return fail(h, MR_ERR_BAD_ARG, "empty label set");
  for (int u = 0; u < n_users; ++u)
    for (long long i = lab_rowptr[u] + 1; i < lab_rowptr[u + 1]; ++i)
      if (lab_col[i] <= lab_col[i - 1]) return fail(h, MR_ERR_BAD_ARG, "label row %d not ascending/unique", u);
  MR_CUDA(h, cudaSetDevice(h->device));
  std::vector<int> new_songs(lab_col, lab_col + n_lab);                       // newSongs = distinct label songs (MR:72, 79)
  std::sort(new_songs.begin(), new_songs.end());
  new_songs.erase(std::unique(new_songs.begin(), new_songs.end()), new_songs.end());
  const int n_new = static_cast<int>(new_songs.size());
  std::vector<long long> lp(lab_rowptr, lab_rowptr + n_users + 1);
  std::vector<void*> tmp;
  double *d_scores = nullptr, *d_ap = nullptr; long long* d_lp = nullptr; int *d_lc = nullptr, *d_new = nullptr; unsigned long long* d_mm = nullptr;
  int rc;
  if ((rc = dev_upload(h, &d_scores, scores_UxS, static_cast<size_t>(n_users) * n_songs, tmp)) || (rc = dev_upload(h, &d_lp, lp.data(), lp.size(), tmp)) ||
      (rc = dev_upload(h, &d_lc, lab_col, static_cast<size_t>(n_lab), tmp)) || (rc = dev_upload(h, &d_new, new_songs.data(), new_songs.size(), tmp)) ||
      (rc = dev_alloc(h, &d_ap, static_cast<size_t>(n_new), tmp)) || (rc = dev_alloc(h, &d_mm, 2, tmp))) { free_list(tmp); return rc; }
  int lrc;
  {
    PhaseTimer t(h, MR_T_OTHER);
    lrc = launch_evaluate(d_scores, n_users, n_songs, d_lp, d_lc, d_new, n_new, n_thresholds, d_mm, d_ap, h->num_sms, h->stream);
    h->launches += 2;
  }
  std::vector<double> ap(n_new);
  cudaError_t e = lrc ? cudaErrorLaunchFailure : cudaMemcpyAsync(ap.data(), d_ap, n_new * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  free_list(tmp);
  if (e != cudaSuccess) return fail(h, MR_ERR_CUDA, "mr_evaluate_dense: %s", cudaGetErrorString(e));
  double total = 0.0;
  for (int c = 0; c < n_new; ++c) total += ap[c];                              // foldLeft(0.0)(_+_), MR:626
  *out_map = total / n_new;
  return MR_OK;
}

static int topk_impl(mr_handle* h, int model, double param, uint64_t seed, int k, const TopkHostOut* ho) {
  int rc = require_test(h);
  if (rc) return rc;
  if (k < 1 || k > 1024) return fail(h, MR_ERR_BAD_ARG, "k must be in [1,1024], got %d", k);
  // The select ranks IEEE bit patterns of non-negative scores.  The reference's linear combination accepts any alpha (MR:317-330 has
  // no range check) and mr_blend_dense follows it, but outside [0,1] a blended score can be negative, which has no place in a ranking
  // of cosine sums: getTopK rejects it (documented deviation).
  if (model == MR_LC && !(param >= 0 && param <= 1))
    return fail(h, MR_ERR_PARAM_RANGE, "getTopK: alpha of the linear combination must be between 0 and 1");
  BlendParams bp;
  if ((rc = make_blend_params(h, model, param, seed, h->n_pairs_total, &bp))) return rc;
  bp.pair_base = h->d_pair_base;
  {  // one packed block (song | score | len) so that a caller can ship the whole result with a single transfer (mr_topk_packed)
    const size_t song_b = static_cast<size_t>(round_up(static_cast<long long>(h->U) * k * 4, 16)), score_b = static_cast<size_t>(h->U) * k * 8,
                 len_b = static_cast<size_t>(round_up(static_cast<long long>(h->U) * 4, 16));
    char* base = nullptr;
    if ((rc = slot_alloc(h, mr_handle::SL_OUT_PACK, &base, song_b + score_b + len_b))) return rc;
    h->d_out_song = reinterpret_cast<int*>(base);
    h->d_out_score = reinterpret_cast<double*>(base + song_b);
    h->d_out_len = reinterpret_cast<int*>(base + song_b + score_b);
    h->out_pack_bytes = song_b + score_b + len_b;
  }
  h->out_k = k;
  if (model != MR_UBM && h->engine != MR_ENGINE_SPARSE && (rc = ensure_gram_ws(h, h->max_batch_rows))) return rc;
  h->have_topk = false;
  const bool streamed = ho && h->space == MR_SPACE_ITEM;          // the item-space pipeline copies slice by slice
  rc = run_batches(h, model, bp, k, RUN_TOPK, streamed ? const_cast<TopkHostOut*>(ho) : nullptr);
  if (streamed) { cudaError_t e = cudaStreamSynchronize(h->copy_stream); if (!rc && e != cudaSuccess) rc = fail(h, MR_ERR_CUDA, "result copy: %s", cudaGetErrorString(e)); }
  if (rc) return rc;
  h->have_topk = true;
  if (ho && !streamed) return mr_topk_fetch(h, k, ho->song, ho->score, ho->len);
  if (ho) MR_CUDA(h, cudaStreamSynchronize(h->stream));
  return MR_OK;
}

int mr_topk_device(mr_handle* h, int model, double param, uint64_t seed, int k) { return topk_impl(h, model, param, seed, k, nullptr); }

int mr_topk_fetch(mr_handle* h, int k, int32_t* out_song, double* out_score, int32_t* out_len) {
  int rc = require_test(h);
  if (rc) return rc;
  if (!h->have_topk || h->out_k != k) return fail(h, MR_ERR_STATE, "no top-%d result on the device (call mr_topk_device first)", k);
  if (!out_song || !out_score || !out_len) return fail(h, MR_ERR_BAD_ARG, "null output");
  MR_CUDA(h, cudaMemcpyAsync(out_song, h->d_out_song, static_cast<size_t>(h->U) * k * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  MR_CUDA(h, cudaMemcpyAsync(out_score, h->d_out_score, static_cast<size_t>(h->U) * k * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  MR_CUDA(h, cudaMemcpyAsync(out_len, h->d_out_len, static_cast<size_t>(h->U) * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  MR_CUDA(h, cudaStreamSynchronize(h->stream));
  return MR_OK;
}

int mr_topk_device_ptrs(mr_handle* h, int k, void** song, void** score, void** len) {
  int rc = require_test(h);
  if (rc) return rc;
  if (!h->have_topk || h->out_k != k) return fail(h, MR_ERR_STATE, "no top-%d result on the device (call mr_topk_device first)", k);
  if (song) *song = h->d_out_song;
  if (score) *score = h->d_out_score;
  if (len) *len = h->d_out_len;
  return MR_OK;
}

int mr_topk_packed(mr_handle* h, int k, void** base, uint64_t* bytes, uint64_t* score_offset, uint64_t* len_offset) {
  int rc = require_test(h);
  if (rc) return rc;
  if (!h->have_topk || h->out_k != k) return fail(h, MR_ERR_STATE, "no top-%d result on the device (call mr_topk_device first)", k);
  if (base) *base = h->d_out_song;
  if (bytes) *bytes = h->out_pack_bytes;
  if (score_offset) *score_offset = static_cast<uint64_t>(reinterpret_cast<char*>(h->d_out_score) - reinterpret_cast<char*>(h->d_out_song));
  if (len_offset) *len_offset = static_cast<uint64_t>(reinterpret_cast<char*>(h->d_out_len) - reinterpret_cast<char*>(h->d_out_song));
  return MR_OK;
}

int mr_topk_merge(mr_handle* h, int k, int n_parts, int n_users, const int32_t* const* part_song, const double* const* part_score,
                  const int32_t* const* part_len, int32_t* out_song, double* out_score, int32_t* out_len) {
  if (!h || !h->stream) return MR_ERR_STATE;
  if (k < 1 || k > 1024 || n_users < 0 || !part_song || !part_score || !part_len || !out_song || !out_score || !out_len)
    return fail(h, MR_ERR_BAD_ARG, "mr_topk_merge: null argument or k outside [1,1024]");
  if (n_parts < 1 || n_parts > kMergeMaxParts) return fail(h, MR_ERR_BAD_ARG, "mr_topk_merge: %d partitions (1..%d)", n_parts, kMergeMaxParts);
  MergeParts p; memset(&p, 0, sizeof p);
  p.n = n_parts;
  for (int i = 0; i < n_parts; ++i) {
    if (!part_song[i] || !part_score[i] || !part_len[i]) return fail(h, MR_ERR_BAD_ARG, "mr_topk_merge: partition %d has a null array", i);
    p.song[i] = part_song[i]; p.score[i] = part_score[i]; p.len[i] = part_len[i];
  }
  MR_CUDA(h, cudaSetDevice(h->device));
  PhaseTimer t(h, MR_T_TOPK);
  MR_LAUNCH(h, launch_merge_topk(p, k, n_users, out_song, out_score, out_len, h->stream));
  return MR_OK;
}

int mr_topk(mr_handle* h, int model, double param, uint64_t seed, int k, int32_t* out_song, double* out_score, int32_t* out_len) {
  if (!out_song || !out_score || !out_len) return fail(h, MR_ERR_BAD_ARG, "null output");
  const TopkHostOut ho{out_song, out_score, out_len};
  return topk_impl(h, model, param, seed, k, &ho);
}

int mr_get_timing(mr_handle* h, double* ms_out, int n) {
  if (!h || !ms_out) return MR_ERR_BAD_ARG;
  for (int i = 0; i < n && i < MR_T_N; ++i) ms_out[i] = h->t_ms[i];
  return MR_OK;
}
int mr_reset_timing(mr_handle* h) {
  if (!h) return MR_ERR_BAD_ARG;
  for (double& t : h->t_ms) t = 0;
  return MR_OK;
}
int mr_get_info(mr_handle* h, int64_t* out, int n) {
  if (!h || !out) return MR_ERR_BAD_ARG;
  unsigned int st[4] = {0, 0, 0, 0};
  if (n > 16 && h->d_topk_stats && cudaMemcpy(st, h->d_topk_stats, sizeof st, cudaMemcpyDeviceToHost) != cudaSuccess) return fail(h, MR_ERR_CUDA, "mr_get_info: reading the select counters failed");
  const int64_t v[19] = {h->engine, h->launches, static_cast<int64_t>(h->dense_bytes), h->n_items, h->num_sms, static_cast<int64_t>(h->dev_bytes), h->space, h->n_head,
                         h->n_head_entries, h->n_tail_entries, h->n_ex, h->batch_rows, h->n_groups,
                         h->h_split_ptr.empty() ? 0 : h->h_split_ptr.back(), h->n_cols, h->win_lo, st[0], st[1], st[2]};
  for (int i = 0; i < n && i < 19; ++i) out[i] = v[i];
  return MR_OK;
}
int mr_prepare_async(mr_handle* h) {
  if (!h || !h->loaded) return fail(h, MR_ERR_STATE, "mr_load has not succeeded on this handle");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return fail(h, MR_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  return start_head_rows(h, h->pre_stream);
}

int mr_prepare(mr_handle* h) {
  if (!h || !h->loaded) return fail(h, MR_ERR_STATE, "mr_load has not succeeded on this handle");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return fail(h, MR_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  int rc = ensure_head_rows(h);
  if (rc) return rc;
  MR_CUDA(h, cudaStreamSynchronize(h->stream));
  return MR_OK;
}
int mr_set_profile(mr_handle* h, int on) {
  if (!h) return MR_ERR_BAD_ARG;
  if (on) h->flags |= MR_PROFILE; else h->flags &= ~static_cast<unsigned>(MR_PROFILE);
  return MR_OK;
}
void* mr_stream(mr_handle* h) { return h ? static_cast<void*>(h->stream) : nullptr; }

}  // extern "C"
#pragma GCC visibility pop
