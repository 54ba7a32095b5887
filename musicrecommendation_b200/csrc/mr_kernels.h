// mr_kernels.h — launcher prototypes shared between the kernel translation units and the C-ABI (mrscore.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mr {

constexpr int kUserBatch = 128;   // test users per batch of the user-space engine = UMMA M
constexpr int kHeadCtasPerSm = 8;  // head_rowsum: resident 256-thread CTAs per SM the UBM pass is compiled for (32 registers per thread)
constexpr int kDenseChunk = 1184;  // rows per device->host chunk of mr_score_dense
constexpr int kTailSubBatch = 592; // tail scatter runs per 592 users so its atomics stay within a ~1.8 GB slice of a Sint panel
// Count panels K1 hands to K2, train-user-major so that one gathered row serves the whole 128-user batch in one coalesced read:
//   UBM  u16 Ct[T][128]   256-byte rows     IBM (user space)  u32 Wi[T][128]   512-byte rows
// (A sub-panel-major variant with 32-byte, L2-resident rows was measured 1.7-1.9x slower on B200: random 32-byte sector
//  gathers out of L2 lose more than the saved HBM traffic gains — profiles/r01_notes.md.)
__host__ __device__ inline long long ct_index(long long T, int v, int b) { (void)T; return static_cast<long long>(v) * kUserBatch + b; }
__host__ __device__ inline long long wi_index(long long T, int v, int b) { (void)T; return static_cast<long long>(v) * kUserBatch + b; }

enum { EPI_I32 = 0, EPI_U16_T = 1, EPI_COS_F32 = 2, EPI_ACC_U64 = 3, EPI_I32_SCATTER = 4 };
enum { MODEL_UBM = 0, MODEL_IBM = 1, MODEL_LC = 2, MODEL_AGG = 3, MODEL_STOCH = 4 };

// ---- K1 (k1_count_gemm.cu)
int launch_count_gemm(const uint8_t* A, long long a_rows, const uint8_t* B, long long b_rows, long long pitch, int M, int N,
                      int epi, void* out, long long ld, const float* rsa, const float* rsb, int num_sms, cudaStream_t st,
                      int shift = 0, int accumulate = 0, int32_t* const* slots = nullptr, int n_slots = 0, int rows_per_owner = 0);
int launch_expand_rows_weighted(const long long* ptr, const int* idx, const uint32_t* weight, int plane, int n_rows, long long pitch,
                                uint8_t* out, cudaStream_t st);
int launch_expand_rows(const long long* ptr, const int* idx, const int* rows, int row0, int n_rows, int n_rows_pad,
                       long long pitch, uint8_t* out, cudaStream_t st);

// stream-ordered completion flags of the fused K-split exchange (k1_count_gemm.cu): a sender raises a 64-bit counter in every owner's
// memory (local or NVLink peer) after its tiles have landed; the owner spins on its own flags before it sums its slots
struct PeerFlags { unsigned long long* p[8]; };
int launch_peer_signal(PeerFlags flags, int n, unsigned long long value, cudaStream_t st);
int launch_peer_wait(const unsigned long long* flags, int n, unsigned long long value, cudaStream_t st);

// ---- K1s (k1_sparse_count.cu): the same counts from the inverted index, for shapes whose dense operands do not fit
int launch_sparse_count_u16t(const long long* te_ptr, const int* te_col, int u0, int n_users, const long long* csc_ptr,
                             const int* csc_idx, uint16_t* ct, long long n_train, cudaStream_t st);
struct CarryList {            // wrap-arounds of the u32 weighted-count panel: (v, b) pairs, repaired by launch_carry_fixup
  unsigned int* count;        // [1]
  uint2* events;              // [capacity]
  unsigned int capacity;
};
int launch_sparse_wcount_u32(const long long* te_ptr, const int* te_col, int u0, int n_users, const long long* csc_ptr,
                             const int* csc_idx, const uint32_t* qd, uint32_t* wi, long long n_train, CarryList carry, cudaStream_t st);
int launch_carry_fixup(CarryList carry, const long long* tr_ptr, const int* tr_col, long long* sint, long long spitch, cudaStream_t st);
int launch_sparse_gram_rows(const int* rows, int n_rows, const long long* csc_ptr, const int* csc_idx, const long long* tr_ptr,
                            const int* tr_col, int32_t* g, long long ldg, cudaStream_t st);

// ---- K2 (k2_aggregate.cu)
struct AggItems {            // work list over the train inverted index: one warp per item
  const int* song;           // [n_items]
  const long long* begin;    // [n_items] offset into csc_idx
  const int* len;            // [n_items]
  const uint8_t* split;      // [n_items] 1 -> the song is covered by several items (accumulate atomically)
  int n_items;
};
int launch_aggregate_ubm(const AggItems& items, const int* csc_idx, const uint32_t* qv, const uint16_t* ct, long long n_train,
                         long long* sint, long long spitch, int num_sms, cudaStream_t st);
int launch_aggregate_w32(const AggItems& items, const int* csc_idx, const uint32_t* wi, long long n_train, long long* sint,
                         long long spitch, int num_sms, cudaStream_t st);
int launch_aggregate_ibm(const long long* te_ptr, const int* te_col, const int* te_grow, const uint32_t* qd, int u0, int n_users,
                         const int32_t* g, long long ldg, int n_songs, long long* sint, long long spitch, cudaStream_t st);

// ---- item-space engine (k4_itemspace.cu)
struct HeadExceptions {      // entries of the packed head rows whose count exceeds 16 bits or whose weighted sum exceeds 32 bits
  unsigned int* count; unsigned int capacity;
  int* row; int* song; uint32_t* g_extra; unsigned long long* gq_extra;
};
int launch_gram_head_scatter(const int* head_song, const long long* lst_ptr, int r0, int r1, const long long* csc_ptr, const int* csc_idx,
                             const long long* tr_ptr, const long long* tr_end, const int* tr_col, const uint32_t* qv, uint32_t* g,
                             unsigned long long* gq, long long pitch, int packed, int num_sms, cudaStream_t st);
int launch_gram_head_direct(const int* head_song, const long long* lst_ptr, int r0, int r1, const long long* csc_ptr, const int* csc_idx,
                            const long long* tr_ptr, const long long* tr_end, const int* tr_col, const uint32_t* qv, uint16_t* g16,
                            uint32_t* gq32, long long pitch, int num_sms, cudaStream_t st);
int launch_pack_head_rows(const uint32_t* g, const unsigned long long* gq, int packed, int r0, int n_rows, long long pitch, int n_songs,
                          uint16_t* g16, uint32_t* gq32, HeadExceptions ex, int num_sms, cudaStream_t st);
// model: 1 = UBM pass over Gq32, 2 = IBM pass over G16; words = 32-bit words per row load (1, 2 or 4); threads per CTA (a CTA covers
// threads * songs-per-thread songs); groups / segments: see k4_itemspace.cu
int launch_head_rowsum(int model, int words, int threads, const int4* grp_hdr, int n_groups, const int4* seg, const int* ge_row,
                       const uint32_t* ge_q, int seg_cap, int ent_cap, const uint16_t* g16, const uint32_t* gq32, long long pitch,
                       int n_songs, long long* sint, long long spitch, cudaStream_t st);
int launch_zero_rows(const int* rows, int n_rows, long long* sint, long long spitch, cudaStream_t st);
int launch_head_fixup(int models, const long long* hu_ptr, const int* hu_row, const int* hu_song, const uint32_t* hu_q, int u0, int n_users,
                      const long long* ex_ptr, const int* ex_song, const uint32_t* ex_g, const unsigned long long* ex_gq,
                      long long* sint_u, long long* sint_i, long long spitch, cudaStream_t st);
int launch_tail_scatter(int models, const int* tu_user, const int* tu_song, const long long* tu_lptr, long long e0, long long e1,
                        const long long* csc_ptr, const int* csc_idx, const long long* tr_ptr, const long long* tr_end, const int* tr_col,
                        const uint32_t* qv, const uint32_t* qd, int u0, long long* sint_u, long long* sint_i, long long spitch, long long n_pairs,
                        int lanes, cudaStream_t st);

// ---- per-shard work lists of the item-space engine, built on the device (k7_testlists.cu)
size_t test_lists_temp_bytes(long long nnz);
int launch_build_test_lists(const long long* te_ptr, const int* te_col, int n_users, long long nnz, const int2* song_info, int* flag,
                            int* head_pos, long long* deg, long long* lsum, void* cub_tmp, size_t cub_tmp_bytes, int* hu_row, int* hu_song,
                            uint32_t* hu_q, long long* hu_ptr, int* tu_user, int* tu_song, long long* tu_lptr, long long* tu_ptr,
                            cudaStream_t st);
int launch_gather_group_entries(const int4* desc, int n_desc, const int* hu_row, const uint32_t* hu_q, int* ge_row, uint32_t* ge_q,
                                cudaStream_t st);

// ---- K3 (k3_topk.cu)
int launch_mask_listened(const long long* te_ptr, const long long* te_end, const int* te_col, int u0, int n_users, long long* sint_u,
                         long long* sint_i, long long spitch, cudaStream_t st);
int launch_dense_scores(int model, const long long* sint, long long spitch, int u0, int n_users, int n_songs, const double* rsa,
                        const double* rsd, double* out, cudaStream_t st);
struct BlendParams {
  int model;                 // MODEL_*
  double alpha, one_minus_alpha;   // LC (MusicRecommender.scala:328)
  long long agg_threshold;   // AGG: (pct * N).toInt (MR:372)
  double prob;               // STOCH (MR:447)
  unsigned long long seed;   // STOCH: java.util.Random seed
  const long long* pair_base;  // [U+1] exclusive prefix of per-user scored-pair counts (index in MAIN:57-59 order)
  const float* rsd_up;       // [S] rsd rounded up to fp32 (upper bounds for the IBM pre-filter of the top-k select), or null
  int ubm_int_ok;            // every UBM numerator of the shard is < 2^52: (double)Sint * rsu is strictly monotone in Sint, top-k may compare integers
  const long long* te_end;   // [U] end of each test row's scored columns: te_ptr + 1, or the in-window prefix ends (song window)
  int song_off;              // added to the ranked column ids: first song of the window (0 without one)
  unsigned int* stats;       // [3] or null: rows the fast path of the select handed to the exact path, short rows (exact path by design), degenerate rows (radix select)
};
int launch_select_bits(const BlendParams& bp, const long long* te_ptr, const int* te_col, int u0, int n_users, int n_songs,
                       uint64_t* sel, long long sel_pitch_words, cudaStream_t st);
int launch_topk(const BlendParams& bp, const long long* te_ptr, const long long* sint_u, const long long* sint_i, long long spitch, const uint64_t* sel,
                long long sel_pitch_words, int u0, int n_users, int n_songs, const double* rsa, const double* rsd, int k,
                int* out_song, double* out_score, int* out_len, cudaStream_t st);
constexpr int kMergeMaxParts = 16;   // song partitions whose ranked lists one launch_merge_topk joins
struct MergeParts { const int* song[kMergeMaxParts]; const double* score[kMergeMaxParts]; const int* len[kMergeMaxParts]; int n; };
int launch_merge_topk(const MergeParts& p, int k, int n_users, int* out_song, double* out_score, int* out_len, cudaStream_t st);
int launch_gather_columns(const double* dense, int n_rows, int n_songs, const int* songs, int n_sel, double* out, long long out_ld, int u_base,
                          cudaStream_t st);
int launch_blend_arrays(const BlendParams& bp, const double* ubm, const double* ibm, double* out, long long n, long long first_index,
                        cudaStream_t st);

// ---- evaluation (k5_evaluate.cu)
int launch_evaluate(const double* scores, int n_users, int n_songs, const long long* lab_ptr, const int* lab_col, const int* new_songs,
                    int n_new, int n_thresholds, unsigned long long* minmax, double* ap_out, int num_sms, cudaStream_t st);

int launch_map_at_k(const int* top_song, const int* top_len, int n_users, int k, const long long* lab_ptr, const int* lab_col,
                    double* ap_out, cudaStream_t st);

}  // namespace mr
