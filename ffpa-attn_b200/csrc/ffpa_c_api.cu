// C ABI (include/ffpa_b200.h): argument validation mirroring the reference launcher's TORCH_CHECK
// contract (/root/reference/csrc/cuffpa/launch.cuh:79-129, ffpa_api.cc:53-62,180-205), then dispatch
// to the sm_100a kernels. No torch types, no host synchronisation, no CPU fallback.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <utility>
#include <vector>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include "ffpa_internal.h"

namespace ffpa {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
// backend hint: thread-local (the reference's process-global atomic, backend.h:16-25, races between threads
// that use different backends); a call with params.impl != AUTO never reads it
static thread_local int32_t g_impl_hint = 0;

// getenv is read once per variable and process (not on every launch)
struct EnvCache { std::mutex mu; std::vector<std::pair<std::string, std::string>> kv; };
static EnvCache& env_cache() { static EnvCache c; return c; }
static std::string env_get(const char* name) {   // by value: the cache may be refreshed concurrently
  EnvCache& c = env_cache();
  std::lock_guard<std::mutex> lk(c.mu);
  for (auto& e : c.kv) if (e.first == name) return e.second;
  const char* v = getenv(name);
  c.kv.emplace_back(name, v ? v : "");
  return c.kv.back().second;
}
void env_refresh() { EnvCache& c = env_cache(); std::lock_guard<std::mutex> lk(c.mu); c.kv.clear(); }
double env_gb(const char* name, double dflt) { const std::string v = env_get(name); return v.empty() ? dflt : atof(v.c_str()); }
bool env_off(const char* name) { const std::string v = env_get(name); return !v.empty() && v[0] == '0'; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

// ---------------------------------------------------------------------------------------------
// balanced schedules: greedy list scheduling (longest processing time first within each group of
// n_items/groups consecutive items keeps the head-major order, hence the L2 reuse of K/V).
// ---------------------------------------------------------------------------------------------
// Cache: keyed by an FNV hash of (device, nclusters, costs) with a full compare on hit; lookup comes first, the
// size cap only matters on insert, where the least recently used table is RETIRED -- freed once an event
// recorded (at eviction) on the stream of its last launch has completed, because kernels in flight may still
// read it; a table that was ever handed out during graph capture is pinned for the life of the process (the
// graph may be replayed at any time). A miss uploads the
// table with a synchronous cudaMalloc + cudaMemcpy (once per shape), which is not legal while the stream is
// being captured into a CUDA graph: a cold shape under capture runs with the static round robin instead
// (correct, only unbalanced) and the table is built on the next eager call.
struct SchedEntry {
  uint64_t hash; std::vector<int> cost; int nclusters; int stride; int* dev; int device; uint64_t last_use;
  cudaStream_t last_stream; bool pinned; cudaEvent_t done;
};
static std::mutex g_sched_mu;
static std::vector<SchedEntry> g_sched;
static std::vector<SchedEntry> g_retired;
static uint64_t g_sched_clock = 0;
static const size_t kSchedCap = 256;

static uint64_t fnv1a(const int* v, int n, int a, int b) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&h](uint32_t x) { for (int i = 0; i < 4; ++i) { h ^= (x >> (8 * i)) & 255u; h *= 1099511628211ull; } };
  mix((uint32_t)a); mix((uint32_t)b); mix((uint32_t)n);
  for (int i = 0; i < n; ++i) mix((uint32_t)v[i]);
  return h;
}

static void reap_retired() {
  for (size_t i = 0; i < g_retired.size();) {
    if (cudaEventQuery(g_retired[i].done) == cudaSuccess) {
      cudaFree(g_retired[i].dev);
      cudaEventDestroy(g_retired[i].done);
      g_retired[i] = std::move(g_retired.back());
      g_retired.pop_back();
    } else {
      ++i;
    }
  }
  cudaGetLastError();   // cudaErrorNotReady from the query is not an error
}

const int* get_schedule(const int* cost, int n_items, int nclusters, int* stride_out, cudaStream_t stream) {
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t h = fnv1a(cost, n_items, dev, nclusters);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  const bool capturing = cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone;
  if (capturing) cudaGetLastError();
  std::lock_guard<std::mutex> lk(g_sched_mu);
  for (auto& e : g_sched)
    if (e.hash == h && e.device == dev && e.nclusters == nclusters && (int)e.cost.size() == n_items &&
        std::memcmp(e.cost.data(), cost, sizeof(int) * n_items) == 0) {
      e.last_use = ++g_sched_clock;
      e.last_stream = stream;
      if (capturing) e.pinned = true;
      *stride_out = e.stride;
      return e.dev;
    }
  if (capturing) return nullptr;   // cold shape under graph capture: static schedule (see above)
  // order: keep the caller's grouping (items arrive head-major, m-tile fastest); inside each run of
  // increasing cost take the longest first
  std::vector<int> order(n_items);
  for (int i = 0; i < n_items; ++i) order[i] = i;
  int run = 0;
  while (run < n_items) {
    int end = run + 1;
    while (end < n_items && cost[end] >= cost[end - 1]) ++end;
    std::reverse(order.begin() + run, order.begin() + end);
    run = end;
  }
  std::vector<long long> load(nclusters, 0);
  std::vector<std::vector<int>> lists(nclusters);
  for (int idx : order) {
    int best = 0;
    for (int c = 1; c < nclusters; ++c)
      if (load[c] < load[best]) best = c;
    load[best] += cost[idx];
    lists[best].push_back(idx);
  }
  size_t stride = 0;
  for (auto& l : lists) stride = l.size() > stride ? l.size() : stride;
  stride += 1;  // room for the -1 terminator
  std::vector<int> table((size_t)nclusters * stride, -1);
  for (int c = 0; c < nclusters; ++c)
    for (size_t k = 0; k < lists[c].size(); ++k) table[(size_t)c * stride + k] = lists[c][k];
  reap_retired();
  if (g_sched.size() >= kSchedCap) {
    size_t lru = g_sched.size();
    for (size_t i = 0; i < g_sched.size(); ++i)
      if (!g_sched[i].pinned && (lru == g_sched.size() || g_sched[i].last_use < g_sched[lru].last_use)) lru = i;
    if (lru == g_sched.size()) return nullptr;   // every table belongs to a captured graph: static schedule
    SchedEntry victim = std::move(g_sched[lru]);
    g_sched[lru] = std::move(g_sched.back());
    g_sched.pop_back();
    // if the record fails (stream already destroyed) the table is leaked rather than freed under a running kernel
    if (cudaEventCreateWithFlags(&victim.done, cudaEventDisableTiming) == cudaSuccess &&
        cudaEventRecord(victim.done, victim.last_stream) == cudaSuccess)
      g_retired.push_back(std::move(victim));
    else
      cudaGetLastError();
  }
  int* d = nullptr;
  if (cudaMalloc(&d, table.size() * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  // synchronous copy: happens once per shape; the table is immutable afterwards
  if (cudaMemcpy(d, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(d);
    cudaGetLastError();
    return nullptr;
  }
  g_sched.push_back(SchedEntry{h, std::vector<int>(cost, cost + n_items), nclusters, (int)stride, d, dev, ++g_sched_clock,
                               stream, false, nullptr});
  *stride_out = (int)stride;
  return d;
}

static int check_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return set_error(FFPA_ERR_NO_DEVICE, "no CUDA device");
  static int cc[64] = {0};
  if (dev >= 0 && dev < 64 && cc[dev] == 0) {
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cc[dev] = major * 10 + minor;
  }
  const int c = (dev >= 0 && dev < 64) ? cc[dev] : 0;
  if (c / 10 != 10)
    return set_error(FFPA_ERR_NO_DEVICE, "ffpa_b200 needs an sm_100-class GPU (found sm_%d); there is no fallback path", c);
  return FFPA_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int check_strides(const char* name, const int64_t* s, const int32_t* dims) {
  if (s[3] != 1) return set_error(FFPA_ERR_INVALID_ARGUMENT, "%s: last-dim stride must be 1, got %lld", name, (long long)s[3]);
  for (int i = 0; i < 3; ++i)
    if (dims[i] > 1 && (s[i] % 8 != 0 || s[i] <= 0))
      return set_error(FFPA_ERR_INVALID_ARGUMENT, "%s: stride[%d]=%lld must be a positive multiple of 8 elements (16 bytes)", name, i, (long long)s[i]);
  return FFPA_OK;
}

// ---------------------------------------------------------------------------------------------
// forward: kernel-family resolution, FP8 knob contract, hybrid staging, workspace planning
// ---------------------------------------------------------------------------------------------
// Resolves which kernel family serves this call. *fp8_bits: 0 = fp16/bf16 kernel, else 1 | 2 smooth-K |
// 4 smooth-V | 8 per-channel V. Knobs that select sm_120 mma.sync variants the tcgen05 kernel does not have are
// refused by name instead of being dropped (reference codes: functional.py:46-67; checks: launch.cuh:303-310).
static int resolve_impl(const ffpa_fwd_params& p, int* fp8_bits) {
  *fp8_bits = 0;
  const int impl = p.impl != FFPA_IMPL_AUTO ? p.impl : g_impl_hint;
  if (impl < FFPA_IMPL_AUTO || impl > FFPA_IMPL_CUTE_TMA_FP4)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "backend impl %d out of range", impl);
  if (impl == FFPA_IMPL_CUTE_TMA_FP4) return set_error(FFPA_ERR_UNSUPPORTED, "FP4 path is not implemented on sm_100a");
  if (impl != FFPA_IMPL_CUTE_TMA_FP8) return FFPA_OK;
  if (p.fp8_q_quant_method == FFPA_QUANT_PER_THREAD || p.fp8_k_quant_method == FFPA_QUANT_PER_THREAD)
    return set_error(FFPA_ERR_UNSUPPORTED, "fp8_q_quant_method / fp8_k_quant_method = 'per_thread' is not implemented on sm_100a "
                     "(per-thread scales follow the mma.sync m16n8k32 fragment layout; tcgen05 takes whole tiles): use 'per_block'");
  if (p.fp8_q_quant_method != FFPA_QUANT_PER_BLOCK || p.fp8_k_quant_method != FFPA_QUANT_PER_BLOCK)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "ffpa_attn: Q/K quant method must be both per_block or both per_thread");
  if (p.fp8_v_quant_method != FFPA_QUANT_PER_BLOCK && p.fp8_v_quant_method != FFPA_QUANT_PER_CHANNEL)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "fp8_v_quant_method must be per_block (0) or per_channel (1), got %d", p.fp8_v_quant_method);
  if (p.fp8_qk_mm_type == FFPA_QK_MM_INT8)
    return set_error(FFPA_ERR_UNSUPPORTED, "fp8_qk_mm_type = 'int8' is not implemented on sm_100a (tcgen05 kind::i8 is not used by this build): use 'fp8'");
  if (p.fp8_qk_mm_type != FFPA_QK_MM_FP8) return set_error(FFPA_ERR_INVALID_ARGUMENT, "fp8_qk_mm_type must be 0 (fp8) or 1 (int8)");
  if (p.fp8_pv_acc_type == FFPA_PV_ACC_F16)
    return set_error(FFPA_ERR_UNSUPPORTED, "fp8_pv_acc_type = 'f16' is not implemented on sm_100a (TMEM accumulators are fp32): use 'f32'");
  if (p.fp8_pv_acc_type != FFPA_PV_ACC_F32) return set_error(FFPA_ERR_INVALID_ARGUMENT, "fp8_pv_acc_type must be 0 (f16) or 1 (f32)");
  if (p.fp8_smooth_v && p.fp8_v_quant_method != FFPA_QUANT_PER_CHANNEL)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "fp8_smooth_v requires fp8_v_quant_method='per_channel'");
  *fp8_bits = 1 | (p.fp8_smooth_k ? 2 : 0) | (p.fp8_smooth_v ? 4 : 0) | (p.fp8_v_quant_method == FFPA_QUANT_PER_CHANNEL ? 8 : 0);
  return FFPA_OK;
}

static int check_fwd(const ffpa_fwd_params* p) {
  if (!p->q || !p->k || !p->v || !p->o) return set_error(FFPA_ERR_INVALID_ARGUMENT, "q/k/v/o must be non-NULL device pointers");
  if (p->dtype != FFPA_DTYPE_F16 && p->dtype != FFPA_DTYPE_BF16)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "dtype must be fp16 or bf16");
  if (p->batch <= 0 || p->heads_q <= 0 || p->heads_kv <= 0 || p->seqlen_q <= 0 || p->seqlen_kv <= 0 || p->head_dim <= 0)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "all sizes must be positive (B=%d Hq=%d Hkv=%d Nq=%d Nkv=%d D=%d)",
                     p->batch, p->heads_q, p->heads_kv, p->seqlen_q, p->seqlen_kv, p->head_dim);
  if (p->heads_q % p->heads_kv != 0)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "Q heads (%d) must be an integer multiple of KV heads (%d)", p->heads_q, p->heads_kv);
  if (p->head_dim % 8 != 0) return set_error(FFPA_ERR_INVALID_ARGUMENT, "head_dim must be a multiple of 8, got %d", p->head_dim);
  if (p->head_dim > 1024) return set_error(FFPA_ERR_INVALID_ARGUMENT, "head_dim must be <= 1024, got %d", p->head_dim);
  const bool varlen = p->cu_seqlens_q != nullptr;
  if (varlen) {
    if (!p->cu_seqlens_k) return set_error(FFPA_ERR_INVALID_ARGUMENT, "cu_seqlens_k must be set together with cu_seqlens_q");
    if (p->total_q <= 0 || p->total_k <= 0) return set_error(FFPA_ERR_INVALID_ARGUMENT, "total_q / total_k must be positive in packed mode");
  }
  if (p->causal && !varlen && p->seqlen_kv < p->seqlen_q)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "causal attention requires Nkv >= Nq (got Nq=%d, Nkv=%d)", p->seqlen_q, p->seqlen_kv);
  if (p->causal && p->bias_kind != FFPA_BIAS_NONE)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "attn bias and causal masking are mutually exclusive");
  if (p->bias_kind != FFPA_BIAS_NONE) {
    if (p->bias_kind != FFPA_BIAS_F32 && p->bias_kind != FFPA_BIAS_QDTYPE)
      return set_error(FFPA_ERR_INVALID_ARGUMENT, "attn bias dtype must be fp32 or match Q");
    if (!p->bias) return set_error(FFPA_ERR_INVALID_ARGUMENT, "bias_kind set but bias pointer is NULL");
    if (p->bias_stride[3] != 1 && !(p->bias_stride[3] == 0))
      return set_error(FFPA_ERR_INVALID_ARGUMENT, "attn bias last dim must be contiguous (stride 1, or 0 for a broadcast key dim)");
  }
  if (!(p->dropout_p >= 0.f && p->dropout_p < 1.f))
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "dropout_p must be in [0, 1), got %f", (double)p->dropout_p);
  const int32_t qd[3] = {varlen ? 1 : p->batch, p->heads_q, varlen ? p->total_q : p->seqlen_q};
  const int32_t kd[3] = {varlen ? 1 : p->batch, p->heads_kv, varlen ? p->total_k : p->seqlen_kv};
  if (int e = check_strides("Q", p->q_stride, qd)) return e;
  if (int e = check_strides("K", p->k_stride, kd)) return e;
  if (int e = check_strides("V", p->v_stride, kd)) return e;
  if (int e = check_strides("O", p->o_stride, qd)) return e;
  if (!aligned16(p->q) || !aligned16(p->k) || !aligned16(p->v) || !aligned16(p->o))
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "q/k/v/o base pointers must be 16-byte aligned");
  if (p->lse_bh_stride != 0 && p->lse_bh_stride < p->seqlen_q)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "lse_bh_stride (%lld) must be 0 or >= seqlen_q", (long long)p->lse_bh_stride);
  return FFPA_OK;
}

// scratch of ONE kernel-family stage over these sizes
static uint64_t fwd_stage_workspace(const ffpa_fwd_params& p, int fp8_bits, uint64_t cap_bytes) {
  if (p.cu_seqlens_q) return 0;
  if (fp8_bits) return fwd_fp8_workspace_bytes(p.batch, p.heads_q, p.heads_kv, p.seqlen_q, p.seqlen_kv, p.head_dim);   // required
  return fwd_split_workspace_bytes(p.batch, p.heads_q, p.heads_kv, p.seqlen_q, p.seqlen_kv, p.head_dim, cap_bytes);   // optional
}

// FP8 hybrid (/root/reference/csrc/cuffpa/launch.cuh:30-58, 341-374): stage 1 = rows [0, n_early) on the fp16/bf16
// kernel against the keys they can see (all keys when not causal), stage 2 = rows [n_early, Nq) on the FP8 kernel
// against all keys. Bottom-right causal alignment makes both stages ordinary calls on zero-copy row views (the
// reference slices, pads and copies); LSE rows land in place through lse_bh_stride.
static bool hybrid_applies(const ffpa_fwd_params& p, int fp8_bits) {
  return fp8_bits && p.fp8_hybrid && p.fp8_hybrid_n_early > 0 && p.seqlen_q >= p.fp8_hybrid_n_early && !p.cu_seqlens_q;
}
static void hybrid_stages(const ffpa_fwd_params& p, ffpa_fwd_params* early, ffpa_fwd_params* late) {
  const int n_early = p.fp8_hybrid_n_early;
  const int64_t lse_bh = p.lse_bh_stride > 0 ? p.lse_bh_stride : p.seqlen_q;
  *early = p;
  early->impl = FFPA_IMPL_NATIVE;
  early->fp8_hybrid = 0;
  early->seqlen_q = n_early;
  if (p.causal) early->seqlen_kv = n_early + (p.seqlen_kv - p.seqlen_q);
  early->lse_bh_stride = lse_bh;
  *late = p;
  late->impl = FFPA_IMPL_CUTE_TMA_FP8;
  late->fp8_hybrid = 0;
  late->seqlen_q = p.seqlen_q - n_early;
  late->q = static_cast<const uint8_t*>(p.q) + 2 * (int64_t)n_early * p.q_stride[2];
  late->o = static_cast<uint8_t*>(p.o) + 2 * (int64_t)n_early * p.o_stride[2];
  late->lse = p.lse ? p.lse + n_early : nullptr;
  late->lse_bh_stride = lse_bh;
}

static int fwd_one(const ffpa_fwd_params& p, int fp8_bits, cudaStream_t stream) {
  if (p.cu_seqlens_q && (p.bias_kind != FFPA_BIAS_NONE || p.dropout_p > 0.f || fp8_bits))
    return set_error(FFPA_ERR_UNSUPPORTED, "packed variable-length mode supports neither attn bias, dropout nor fp8");
  if (fp8_bits) return launch_fwd_fp8_sm100(p, fp8_bits, stream);
  return launch_fwd_sm100(p, stream);
}

}  // namespace ffpa

using namespace ffpa;

extern "C" {

int ffpa_b200_fwd(const ffpa_fwd_params* p, void* stream) {
  if (!p) return set_error(FFPA_ERR_INVALID_ARGUMENT, "params is NULL");
  if (int e = check_device()) return e;
  if (int e = check_fwd(p)) return e;
  int fp8_bits = 0;
  if (int e = resolve_impl(*p, &fp8_bits)) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!hybrid_applies(*p, fp8_bits)) return fwd_one(*p, fp8_bits, st);
  if (p->fp8_hybrid_n_early % 128 != 0)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "ffpa_attn: fp8_hybrid_n_early must be multiple of 128");
  ffpa_fwd_params early, late;
  hybrid_stages(*p, &early, &late);
  if (int e = fwd_one(early, 0, st)) return e;   // the stages share the scratch: stream order serialises them
  if (late.seqlen_q <= 0) return FFPA_OK;
  return fwd_one(late, fp8_bits, st);
}

uint64_t ffpa_b200_fwd_workspace_bytes_p(const ffpa_fwd_params* p, uint64_t cap_bytes) {
  if (!p || p->batch <= 0 || p->heads_q <= 0 || p->heads_kv <= 0 || p->seqlen_q <= 0 || p->seqlen_kv <= 0 || p->head_dim <= 0) return 0;
  int fp8_bits = 0;
  if (resolve_impl(*p, &fp8_bits)) return 0;
  if (!hybrid_applies(*p, fp8_bits)) return fwd_stage_workspace(*p, fp8_bits, cap_bytes);
  ffpa_fwd_params early, late;
  hybrid_stages(*p, &early, &late);
  const uint64_t a = fwd_stage_workspace(early, 0, cap_bytes), b = late.seqlen_q > 0 ? fwd_stage_workspace(late, fp8_bits, cap_bytes) : 0;
  return a > b ? a : b;
}

static int check_bwd(const ffpa_bwd_params* p) {
  if (!p->q || !p->k || !p->v || !p->o || !p->lse || !p->d_o || !p->dq || !p->dk || !p->dv)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "q/k/v/o/lse/dO/dQ/dK/dV must be non-NULL device pointers");
  if (p->dtype != FFPA_DTYPE_F16 && p->dtype != FFPA_DTYPE_BF16)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "dtype must be fp16 or bf16");
  if (p->batch <= 0 || p->heads_q <= 0 || p->heads_kv <= 0 || p->seqlen_q <= 0 || p->seqlen_kv <= 0 || p->head_dim <= 0)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "all sizes must be positive");
  if (p->heads_q % p->heads_kv != 0)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "Q heads (%d) must be an integer multiple of KV heads (%d)", p->heads_q, p->heads_kv);
  if (p->head_dim % 8 != 0) return set_error(FFPA_ERR_INVALID_ARGUMENT, "head_dim must be a multiple of 8, got %d", p->head_dim);
  const bool varlen = p->cu_seqlens_q != nullptr;
  if (varlen) {
    if (!p->cu_seqlens_k) return set_error(FFPA_ERR_INVALID_ARGUMENT, "cu_seqlens_k must be set together with cu_seqlens_q");
    if (p->total_q <= 0 || p->total_k <= 0) return set_error(FFPA_ERR_INVALID_ARGUMENT, "total_q / total_k must be positive in packed mode");
    if (p->bias_kind != FFPA_BIAS_NONE || p->dropout_p > 0.f || p->d_bias)
      return set_error(FFPA_ERR_UNSUPPORTED, "packed variable-length mode supports neither attn bias nor dropout");
  }
  if (p->causal && !varlen && p->seqlen_kv < p->seqlen_q)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "causal attention requires Nkv >= Nq");
  if (p->head_dim > 1024)
    return set_error(FFPA_ERR_UNSUPPORTED, "backward kernels support head_dim <= 1024 (got %d)", p->head_dim);
  if (p->bias_kind != FFPA_BIAS_NONE) {
    if (p->causal) return set_error(FFPA_ERR_INVALID_ARGUMENT, "attn bias and causal masking are mutually exclusive");
    if (p->bias_kind != FFPA_BIAS_F32 && p->bias_kind != FFPA_BIAS_QDTYPE)
      return set_error(FFPA_ERR_INVALID_ARGUMENT, "attn bias dtype must be fp32 or match Q");
    if (!p->bias || (p->bias_stride[3] != 1 && p->bias_stride[3] != 0))
      return set_error(FFPA_ERR_INVALID_ARGUMENT, "attn bias must be non-NULL with a contiguous (or broadcast) last dim");
  }
  if (p->d_bias) {
    if (p->bias_kind == FFPA_BIAS_NONE) return set_error(FFPA_ERR_INVALID_ARGUMENT, "d_bias requested without an attn bias");
    for (int i = 0; i < 4; ++i)
      if ((p->bias_stride[i] == 0) != (p->d_bias_stride[i] == 0))
        return set_error(FFPA_ERR_INVALID_ARGUMENT, "d_bias must broadcast over exactly the dims the bias broadcasts over (dim %d)", i);
  }
  if (!(p->dropout_p >= 0.f && p->dropout_p < 1.f))
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "dropout_p must be in [0, 1), got %f", (double)p->dropout_p);
  return FFPA_OK;
}

int ffpa_b200_bwd(const ffpa_bwd_params* p, void* stream) {
  if (!p) return set_error(FFPA_ERR_INVALID_ARGUMENT, "params is NULL");
  if (int e = check_device()) return e;
  if (int e = check_bwd(p)) return e;
  const bool varlen = p->cu_seqlens_q != nullptr;
  const int32_t qd[3] = {varlen ? 1 : p->batch, p->heads_q, varlen ? p->total_q : p->seqlen_q};
  const int32_t kd[3] = {varlen ? 1 : p->batch, p->heads_kv, varlen ? p->total_k : p->seqlen_kv};
  if (int e = check_strides("Q", p->q_stride, qd)) return e;
  if (int e = check_strides("K", p->k_stride, kd)) return e;
  if (int e = check_strides("V", p->v_stride, kd)) return e;
  if (int e = check_strides("O", p->o_stride, qd)) return e;
  if (int e = check_strides("dO", p->do_stride, qd)) return e;
  if (int e = check_strides("dQ", p->dq_stride, qd)) return e;
  if (int e = check_strides("dK", p->dk_stride, kd)) return e;
  if (int e = check_strides("dV", p->dv_stride, kd)) return e;
  if (!aligned16(p->q) || !aligned16(p->k) || !aligned16(p->v) || !aligned16(p->o) || !aligned16(p->d_o) ||
      !aligned16(p->dq) || !aligned16(p->dk) || !aligned16(p->dv) || !aligned16(p->workspace))
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "all tensor base pointers must be 16-byte aligned");
  return launch_bwd_sm100(*p, static_cast<cudaStream_t>(stream));
}

uint64_t ffpa_b200_bwd_workspace_bytes_p(const ffpa_bwd_params* p, uint64_t cap_bytes) {
  if (!p || p->batch <= 0 || p->heads_q <= 0 || p->heads_kv <= 0 || p->seqlen_q <= 0 || p->seqlen_kv <= 0 || p->head_dim <= 0) return 0;
  if (p->cu_seqlens_q) return bwd_workspace_bytes_min(p->batch, p->heads_q, p->heads_kv, p->seqlen_q, p->seqlen_kv, p->head_dim);
  return bwd_workspace_bytes(p->batch, p->heads_q, p->heads_kv, p->seqlen_q, p->seqlen_kv, p->head_dim, cap_bytes);
}

uint64_t ffpa_b200_bwd_workspace_bytes_min_p(const ffpa_bwd_params* p) {
  if (!p || p->batch <= 0 || p->heads_q <= 0 || p->heads_kv <= 0 || p->seqlen_q <= 0 || p->seqlen_kv <= 0 || p->head_dim <= 0) return 0;
  return bwd_workspace_bytes_min(p->batch, p->heads_q, p->heads_kv, p->seqlen_q, p->seqlen_kv, p->head_dim);
}

int ffpa_b200_set_backend_impl(int32_t impl) {
  if (impl < FFPA_IMPL_AUTO || impl > FFPA_IMPL_CUTE_TMA_FP4)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "backend impl hint %d out of range", impl);
  g_impl_hint = impl;
  return FFPA_OK;
}
int32_t ffpa_b200_get_backend_impl(void) { return g_impl_hint; }
int32_t ffpa_b200_fwd_available(void) { return 1; }
int32_t ffpa_b200_bwd_available(void) { return 1; }
int32_t ffpa_b200_abi_version(void) { return FFPA_B200_ABI_VERSION; }
uint64_t ffpa_b200_launch_count(void) { return g_launches.load(); }
const char* ffpa_b200_last_error(void) { return g_err; }
void ffpa_b200_refresh_env(void) { env_refresh(); }

}  // extern "C"
