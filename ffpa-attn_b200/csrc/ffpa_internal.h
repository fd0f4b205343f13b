// Internal declarations shared by the C-ABI translation unit and the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/ffpa_b200.h"

namespace ffpa {

// kernel-side view of ffpa_fwd_params (tensor maps carry Q/K/V)
struct FwdKernelParams {
  void* o;
  float* lse;
  int64_t lse_bh_stride;   // elements between the LSE rows of consecutive (b, h) pairs (dense layout)
  const void* bias;
  int64_t o_stride[3];     // (b, h, n) in elements
  int64_t bias_stride[4];  // (b, h, q, k) in elements
  int batch, heads_q, heads_kv, seqlen_q, seqlen_kv, head_dim;
  int causal, bias_kind;
  float scale_log2;  // softmax_scale * log2(e)
  float dropout_p;
  uint64_t philox_seed, philox_offset;
  int n_mtiles, n_items;
  // optional work schedule (causal load balancing): sched[cluster * sched_stride + k] = item id,
  // -1 terminates; nullptr -> static round robin item = cluster + k * nclusters
  const int* sched;
  int sched_stride;
  // KV splits (decode-like shapes): fp32 partials [S, B, Hq, Nq, D] / [S, B, Hq, Nq]; kv_splits == 1: off
  int kv_splits;
  float* part_o;
  float* part_lse;
  // packed variable-length mode (cu_q != nullptr): token offsets per sequence, LSE is [Hq, total_q]
  const int* cu_q;
  const int* cu_k;
  int total_q, total_k;
  // Replay path for head dims > 768 (two O slabs): instead of recomputing S for the second slab, pass 0 stores
  // its P tiles (16-bit, tile-major [B Hq][q tile][64-key block][128][64]), the O rescale factor of every
  // (row, KV tile) and 1 / rowsum; ffpa_fwd_replay_kernel then computes O[:, 512:] = sum P V[:, 512:] as a
  // GEMM that replays the rescales. n_pass = slab passes the items enumerate (1 when the replay path is on).
  void* stash_p;
  float* stash_f;    // [B Hq][q tiles (even)][KV tiles][128]
  float* stash_inv;  // [B Hq][q tiles (even) * 128]
  int nk_pad, n_mt_even, n_pass;
  int o_tma;   // map_o is valid: the epilogue may TMA-store O tiles staged in the (then idle) P buffers
};

// second-slab forward GEMM over stashed P tiles (256 query rows per 2-CTA cluster, M = 256)
struct FwdReplayParams {
  void* o;
  int64_t o_stride[3];
  const float* stash_f;
  const float* stash_inv;
  int batch, heads_q, heads_kv, seqlen_q, seqlen_kv, head_dim;
  int causal;
  int nk_pad, n_mt_even;
  int n_qblocks, n_items;   // 256-row blocks per (b, h)
  const int* sched;
  int sched_stride;
};

namespace bwd {
// kernel-side parameters of the backward kernels (tensor maps carry the operands)
struct BwdKernelParams {
  void* out;               // dQ / dK / dV
  int64_t out_stride[3];   // (b, h, n) elements
  const float* lse2;       // [B, Hq, Nq_pad]  LSE * log2(e); +inf for rows without keys / padding
  const float* delta;      // [B, Hq, Nq_pad]  rowsum(dO * O); 0 in the padding
  int nq_pad;
  int batch, heads_q, heads_kv, seqlen_q, seqlen_kv, head_dim;
  int causal;
  float scale_log2;        // softmax_scale * log2(e)
  float scale;             // softmax_scale
  int n_rtiles, n_items;   // row tiles (128 stationary rows) per (b, head); total items
  // GENERAL variants: additive bias, dropout replay, dBias (fp32 [B, Hq, Nq, Nkv], dQ kind writes it)
  const void* bias;
  int64_t bias_stride[4];
  int bias_kind;
  float dropout_p;
  uint64_t philox_seed, philox_offset;
  float* dbias;             // fp32, bias-shaped (dbias_stride 0 on the dims the bias broadcasts over)
  int64_t dbias_stride[4];
  // scheduling: optional balanced table (as in the forward) and chunked items with fp32 atomics
  const int* sched;
  int sched_stride;
  int n_chunks, chunk_len;
  int n_pass;     // output-column slab passes per item (2 when head_dim > 512, else 1)
  // packed variable-length mode (cu_q != nullptr): seqlen_q / seqlen_kv are the maxima, batch = sequences
  const int* cu_q;
  const int* cu_k;
  int total_q, total_k;
  // stash path (dQ kind, large head dims): P_drop and dS tiles as 16-bit [B, Hq, nq_pad, nk_pad] for the
  // GEMM-only dK / dV kernel (ffpa_bwd_gemm_sm100.cuh); nullptr = off
  void* stash_p;
  void* stash_ds;
  int nk_pad;
  float* out32;   // non-null: accumulate into fp32 [B, H, out_rows, D] instead of storing `out`
  int out_rows;
  int zero_single;  // 1: rows that see at most one key get the analytic dS = 0 (only valid without a dLSE input)
};
// GEMM-only dK / dV kernel over stashed score tiles (256 keys per 2-CTA cluster)
struct BwdGemmParams {
  void* out;               // dK or dV
  int64_t out_stride[3];
  int batch, heads_q, heads_kv, seqlen_q, seqlen_kv, head_dim;
  int causal;
  float mul;               // softmax_scale for dK, 1 for dV
  int n_kblocks, n_items;  // 256-key blocks per (b, kv head); items = n_kblocks * B * Hkv
  int nk_pad;              // padded key count of the stash (multiple of 256)
  int n_pass;              // 512-wide output slabs per item (2 when head_dim > 512)
  const int* sched;
  int sched_stride;
};
}  // namespace bwd

// balanced (greedy LPT) item schedule, cached per shape on the device; returns nullptr on failure
const int* get_schedule(const int* cost, int n_items, int nclusters, int* stride_out, cudaStream_t stream);

int set_error(int code, const char* fmt, ...);
void count_launch();
int sm_count();

int launch_fwd_sm100(const ffpa_fwd_params& a, cudaStream_t stream);
int launch_bwd_sm100(const ffpa_bwd_params& a, cudaStream_t stream);
// fp8_bits: 1 enable | 2 smooth-K | 4 smooth-V | 8 per-channel V scales
int launch_fwd_fp8_sm100(const ffpa_fwd_params& a, int fp8_bits, cudaStream_t stream);
int fwd_kv_splits(int batch, int heads_q, int seqlen_q, int seqlen_kv, int head_dim);  // 1 = no split
// optional forward scratch not above cap_bytes: replay stash (head dims > 768, chunked when needed) or KV-split partials
uint64_t fwd_split_workspace_bytes(int batch, int heads_q, int heads_kv, int seqlen_q, int seqlen_kv, int head_dim, uint64_t cap_bytes);
uint64_t fwd_fp8_workspace_bytes(int batch, int heads_q, int heads_kv, int seqlen_q, int seqlen_kv, int head_dim);
// recommended backward scratch not above max(cap, minimum): adds the stash buffers (chunked when needed)
uint64_t bwd_workspace_bytes(int batch, int heads_q, int heads_kv, int seqlen_q, int seqlen_kv,
                             int head_dim, uint64_t cap_bytes);
double env_gb(const char* name, double dflt);   // cached getenv (read once per process)
bool env_off(const char* name);                 // cached: variable set to a value starting with '0'
uint64_t bwd_workspace_bytes_min(int batch, int heads_q, int heads_kv, int seqlen_q, int seqlen_kv,
                                 int head_dim);

}  // namespace ffpa
