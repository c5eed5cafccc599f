// doc2tex_b200 engine: host orchestration + C ABI (include/doc2tex_b200.h).
//
// One engine per (process, device).  Weights arrive by their reference state_dict keys
// (SURVEY.md Appendix C), are folded / repacked once, and every API call enqueues hand-written
// sm_100a kernels on the caller's stream.  There is no CPU fallback anywhere in this file.
#include "../../include/doc2tex_b200.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "decoder_kernels.cuh"
#include "encoder_kernels.cuh"
#include "gemm_simt.cuh"
#include "lstm_kernels.cuh"
#include "gemm_tc.cuh"
#include "gemm_tc3.cuh"
#include "gemm_tc5.cuh"
#include "preprocess_kernels.cuh"

using namespace d2t;

namespace {

thread_local std::string g_create_error;

struct HostTensor {
  std::vector<float> f;
  std::vector<int64_t> shape;
};

struct ConvW {
  float* w = nullptr;      // [Cout][KH][KW][Cin]
  float* scale = nullptr;  // folded BN scale (or nullptr)
  float* shift = nullptr;  // folded BN shift / bias
  int cout = 0, cin = 0, kh = 0, kw = 0;
};

struct Fmap {  // NHWC activation (+ optional bf16 hi/lo planes for the cp.async-fed tensor-core convolutions)
  float* p = nullptr;
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  int B = 0, H = 0, W = 0, C = 0;
  size_t numel() const { return (size_t)B * H * W * C; }
};

// Caching slot allocator: cudaMalloc only the first time a size class is needed (warm-up), then reuse.
struct SlotPool {
  struct Slot { void* p; size_t cap; bool used; };
  std::vector<Slot> slots;
  size_t total = 0;
  void* get(size_t bytes, cudaError_t* err) {
    bytes = (bytes + 255) & ~(size_t)255;
    int best = -1;
    for (int i = 0; i < (int)slots.size(); ++i)
      if (!slots[i].used && slots[i].cap >= bytes && (best < 0 || slots[i].cap < slots[best].cap)) best = i;
    if (best >= 0 && slots[best].cap <= bytes * 2 + (1 << 20)) { slots[best].used = true; return slots[best].p; }
    void* p = nullptr;
    *err = cudaMalloc(&p, bytes);
    if (*err != cudaSuccess) return nullptr;
    slots.push_back({p, bytes, true});
    total += bytes;
    return p;
  }
  void release(void* p) {
    for (auto& s : slots) if (s.p == p) { s.used = false; return; }
  }
  void release_all() { for (auto& s : slots) s.used = false; }
  void destroy() { for (auto& s : slots) cudaFree(s.p); slots.clear(); total = 0; }
};

struct Tap { Fmap a; bool tokens; };

}  // namespace

constexpr int D2T_MAX_GROUPS = 8;

struct d2t_engine {
  d2t_config cfg{};
  int device = 0;
  int num_sms = 148;
  int enc_sms = 148;     // SMs the encoder's persistent kernels may occupy (the rest stay free for a concurrent decode)
  int active_sms = 148;  // grid-sizing budget of the API call being enqueued
  std::string err;
  bool finalized = false;
  int64_t launches = 0;

  std::unordered_map<std::string, HostTensor> host;
  std::vector<void*> owned;  // device weight allocations
  std::map<std::string, ConvW> conv;
  std::map<std::string, float*> dev;  // linear weights, biases, LN params, embeddings (by reference key)
  std::map<const float*, TcWeight> tcw;  // tensor-core operand planes + TMA maps, keyed by the fp32 weight matrix
  std::map<const float*, Tc3Maps> tc3;   // 64-byte-row weight maps of the cp.async-fed stem-convolution kernel

  SlotPool enc_pool, dec_pool;
  bool keep_taps = false;
  bool use_pdl = true;   // option "pdl": programmatic dependent launch in the decode step
  int pdl_max_rows = 2048;   // option "pdl_max_rows": decode calls with more rows launch without it (2 560-row calls: 6 084 vs 5 990 formulas/s)
  int use_pair = 2;      // option "pair": CTA-pair (cta_group::2) kernel for the wide stem convolutions (see run_contraction)
  bool use_tc3 = true;   // option "tc3": stem convolutions fed from bf16 activation planes by cp.async
  // ViTEncoder (fix_embed: False, interpolate_embed: True): pos_embed is resampled bicubically to the grid of each image
  // size (vit_encoder.py:58-95) — options "pos_interpolate", "pos_grid_h", "pos_grid_w"; tables cached per grid
  bool pos_interpolate = false;
  int pos_grid_h = 0, pos_grid_w = 0;
  std::map<std::pair<int, int>, float*> pos_tables;
  bool time_conv = false;   // option "time_conv": bracket layer3.1.conv1 with events (d2t_debug_conv_time)
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> conv_events;
  double conv_flops = 0.0;
  // option "time_decode": the decode loop runs eagerly and brackets the memory-bound launches of decoder layer 1 with events
  struct DecEvent { int kind; double bytes; cudaEvent_t a, b; };   // kind: 0 self-attn, 1 cross-attn, 2 beam step, 3 greedy pick
  std::vector<DecEvent> dec_events;
  float* beam_runner_up = nullptr;   // d2t_debug_beam_runner_up: caller-owned (B, max_steps) buffer filled by d2t_decode_beam
  int cur_step = 0;   // host copy of the step being enqueued (eager mode only; the device reads its own counter)
  bool fuse_pick = true;    // option "fuse_pick": 0 = separate embed / advance launches in the greedy decode step
  bool lean_acts = true;    // option "lean_acts": 0 = every stem layer writes fp32 AND operand planes, read or not
  int split_k = 1;       // option "split_k": 0 = never, 1 = auto split-K of the LayerNorm-fed decode projections
  int attn_kpi = 4;         // option "attn_kpi": keys in flight per quarter warp of the decode attention walk (2, 4, 8)
  int attn_split = 0;       // option "attn_split": warps per (row, head) of the per-row decode attention (0 = auto)
  bool stack_mma = true;    // option "stack_mma": bf16x3 decode projections issue 2 MMAs per k-step against [W_hi ; W_lo]
  int steps_per_graph = 8;  // option "steps_per_graph": decode steps captured per CUDA graph (= the early-exit poll interval)
  bool kv_bf16 = true;      // option "kv_bf16": bf16 KV caches in the single-pass bf16 mode (fp32-parity modes keep fp32)
  bool wide_decode = true;  // option "wide_decode": decode projections with more tiles than SMs run on the stem's persistent kernels
  int vit_planes = 1;       // option "vit_planes": the ViT blocks' Linears read bf16 operand planes by TMA (1: auto — CTA-pair kernel
                            // in bf16x3, single-CTA kernel in bf16 —, 2: single-CTA kernel only, 3: CTA pair wherever it applies)
  bool fuse_pool = true;    // option "fuse_pool": 2x2 max-pools 1 and 2 fused into the producing convolution's epilogue
  bool time_decode = false; // option "time_decode": bracket the decode-attention / beam-step launches with events
  bool dbg_decode = false, dbg_timeline = false;   // options "dbg_decode" / "dbg_timeline": phase / per-launch timestamps
  std::map<std::string, Tap> taps;

  // decode graph cache
  struct GraphEntry { std::vector<long long> key; cudaGraphExec_t exec = nullptr; int nodes = 0; };
  std::vector<GraphEntry> graphs;
  int* h_counters = nullptr;  // pinned [D2T_MAX_GROUPS][4]
  int decode_groups = 0;      // option "decode_groups": concurrent row groups of a decode call (0 = auto)
  cudaStream_t side[D2T_MAX_GROUPS] = {};   // side[g], g >= 1: stream of row group g (group 0 runs on `work`)
  cudaEvent_t ev_fork = nullptr, ev_join[D2T_MAX_GROUPS] = {};
  cudaEvent_t ev_poll = nullptr;   // early-exit poll of the decode loop (checked one poll late, see tfm_decode)
  cudaStream_t work = nullptr;   // engine-owned stream for the decode loop (the legacy default stream cannot be captured)
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;

  int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    err = buf;
    return code;
  }
};

// Captured decode-step graphs bake in the kernel selection: dropped whenever weights or a decode-affecting option change.
static void drop_graphs(d2t_engine* e) {
  for (auto& g : e->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
}

int finalize_attn_extras(d2t_engine* e);

#define CUDA_TRY(e, call)                                                                          \
  do {                                                                                             \
    cudaError_t _s = (call);                                                                       \
    if (_s != cudaSuccess) return (e)->fail(D2T_ERR_CUDA, "%s failed: %s (%s:%d)", #call,          \
                                            cudaGetErrorString(_s), __FILE__, __LINE__);          \
  } while (0)

// Decode loops run on the engine's own stream, ordered after the caller's stream on entry and
// joined back on exit (RAII), so that a step can be captured into a CUDA graph even when the caller
// is on the legacy default stream.
struct WorkStream {
  d2t_engine* e; cudaStream_t caller;
  WorkStream(d2t_engine* e_, cudaStream_t c) : e(e_), caller(c) {
    cudaEventRecord(e->ev_in, caller);
    cudaStreamWaitEvent(e->work, e->ev_in, 0);
  }
  ~WorkStream() {
    cudaEventRecord(e->ev_out, e->work);
    cudaStreamWaitEvent(caller, e->ev_out, 0);
  }
  cudaStream_t get() const { return e->work; }
};

namespace {

const std::string SEQ = "seqmodeler.SequenceModeling.";
const std::string NET = SEQ + "patch_embed.backbone.ConvNet.";
const std::string PRED = "predicter.Prediction.";

int upload(d2t_engine* e, const float* src, size_t n, float** out) {
  float* p = nullptr;
  CUDA_TRY(e, cudaMalloc(&p, n * sizeof(float)));
  e->owned.push_back(p);
  CUDA_TRY(e, cudaMemcpy(p, src, n * sizeof(float), cudaMemcpyHostToDevice));
  *out = p;
  return 0;
}

const HostTensor* find(d2t_engine* e, const std::string& key) {
  auto it = e->host.find(key);
  return it == e->host.end() ? nullptr : &it->second;
}

int need(d2t_engine* e, const std::string& key, const HostTensor** t, std::vector<int64_t> shape = {}) {
  *t = find(e, key);
  if (!*t) return e->fail(D2T_ERR_MISSING, "state_dict tensor '%s' was not loaded", key.c_str());
  if (!shape.empty() && (*t)->shape != shape) {
    std::string got, want;
    for (auto v : (*t)->shape) got += std::to_string(v) + ",";
    for (auto v : shape) want += std::to_string(v) + ",";
    return e->fail(D2T_ERR_INVALID, "tensor '%s' has shape (%s) expected (%s)", key.c_str(), got.c_str(), want.c_str());
  }
  return 0;
}

int upload_key(d2t_engine* e, const std::string& key, std::vector<int64_t> shape = {}) {
  const HostTensor* t;
  if (int rc = need(e, key, &t, shape)) return rc;
  float* p;
  if (int rc = upload(e, t->f.data(), t->f.size(), &p)) return rc;
  e->dev[key] = p;
  return 0;
}

// Conv2d(bias=False) + BatchNorm2d(eval): y = conv(x) * alpha + beta with alpha = gamma / sqrt(var + eps),
// beta = bias - mean * alpha (what torch's CPU inference kernel evaluates; resnet.py:32-40, quirk Q1).
int make_conv(d2t_engine* e, const std::string& cname, const std::string& bname, int cout, int cin, int kh, int kw) {
  const HostTensor *w, *g, *b, *m, *v;
  if (int rc = need(e, NET + cname + ".weight", &w, {cout, cin, kh, kw})) return rc;
  if (int rc = need(e, NET + bname + ".weight", &g, {cout})) return rc;
  if (int rc = need(e, NET + bname + ".bias", &b, {cout})) return rc;
  if (int rc = need(e, NET + bname + ".running_mean", &m, {cout})) return rc;
  if (int rc = need(e, NET + bname + ".running_var", &v, {cout})) return rc;
  std::vector<float> packed((size_t)cout * kh * kw * cin), alpha(cout), beta(cout);
  for (int o = 0; o < cout; ++o) {
    for (int i = 0; i < cin; ++i)
      for (int y = 0; y < kh; ++y)
        for (int x = 0; x < kw; ++x)
          packed[(((size_t)o * kh + y) * kw + x) * cin + i] = w->f[(((size_t)o * cin + i) * kh + y) * kw + x];
    const float invstd = 1.0f / sqrtf(v->f[o] + 1e-5f);
    alpha[o] = g->f[o] * invstd;
    beta[o] = b->f[o] - m->f[o] * alpha[o];
  }
  ConvW c;
  c.cout = cout; c.cin = cin; c.kh = kh; c.kw = kw;
  if (int rc = upload(e, packed.data(), packed.size(), &c.w)) return rc;
  if (int rc = upload(e, alpha.data(), cout, &c.scale)) return rc;
  if (int rc = upload(e, beta.data(), cout, &c.shift)) return rc;
  e->conv[cname] = c;
  return 0;
}

int make_layer(d2t_engine* e, const std::string& name, int cin, int planes, int blocks) {
  for (int b = 0; b < blocks; ++b) {
    const int ci = b == 0 ? cin : planes;
    const std::string p = name + "." + std::to_string(b);
    if (int rc = make_conv(e, p + ".conv1", p + ".bn1", planes, ci, 3, 3)) return rc;
    if (int rc = make_conv(e, p + ".conv2", p + ".bn2", planes, planes, 3, 3)) return rc;
    if (b == 0 && ci != planes)
      if (int rc = make_conv(e, p + ".downsample.0", p + ".downsample.1", planes, ci, 1, 1)) return rc;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------
inline int grid_for(long long total, int block, int num_sms) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)num_sms * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int run_contraction(d2t_engine* e, const ConvGemm& p, const TcWeight* tcw, int precision, cudaStream_t s) {
  if (p.K % 16 != 0 || p.C % 16 != 0)
    return e->fail(D2T_ERR_UNSUPPORTED, "contraction needs K and C multiples of 16 (K=%d C=%d)", p.K, p.C);
  if (precision != D2T_PREC_FP32 && tcw == nullptr) {
    auto it = e->tcw.find(p.w);
    if (it != e->tcw.end()) tcw = &it->second;
  }
  if (precision != D2T_PREC_FP32 && tcw != nullptr && tcw->ready && tcw->N == p.N && tcw->K == p.K && tc_supported(p)) {
    if (e->use_tc3 && tc3_supported(p, precision)) {
      auto m3 = e->tc3.find(p.w);
      if (m3 != e->tc3.end() && m3->second.ready) {
        // large 256-multiple-wide convolutions: CTA pair with both operands by TMA (option "pair": 0 never, 1 single-pass bf16
        // only — where shared-memory bandwidth bounds the single-CTA kernel —, 2 also the 3-pass parity mode)
        // short-K 1x1 problems (ViT Linears, downsample branches) are epilogue-bound: measured (tools/encoder_ab.py vit_planes=0,1,2,
        // profiles/r02d_encoder_ab_vit_planes.txt) the pair kernel wins in the 3-pass mode, the 128-wide single-CTA tile in bf16
        const bool vit_single = (e->vit_planes == 2 || (e->vit_planes == 1 && precision == D2T_PREC_BF16)) && p.KH == 1 && p.KW == 1 && p.K <= 1024;
        if (e->use_pair && !vit_single && !p.single_cta && (precision == D2T_PREC_BF16 || e->use_pair >= 2) && tc3_use_a_tma() && tc5_supported(p, precision, e->active_sms)) {
          cudaError_t st5 = launch_conv_gemm_tc5(p, *tcw, precision, s, e->active_sms);
          if (st5 != cudaSuccess) return e->fail(D2T_ERR_CUDA, "CTA-pair contraction launch failed: %s", cudaGetErrorString(st5));
          e->launches += 1;
          return 0;
        }
        cudaError_t st3 = launch_conv_gemm_tc3(p, m3->second, precision, s, e->active_sms);
        if (st3 != cudaSuccess) return e->fail(D2T_ERR_CUDA, "cp.async-fed contraction launch failed: %s", cudaGetErrorString(st3));
        e->launches += 1;
        return 0;
      }
    }
    cudaError_t st = launch_conv_gemm_tc(p, *tcw, precision, s, e->active_sms);
    if (st != cudaSuccess) return e->fail(D2T_ERR_CUDA, "tcgen05 contraction launch failed: %s", cudaGetErrorString(st));
    e->launches += 1;
    return 0;
  }
  cudaError_t st = launch_conv_gemm_simt(p, s, e->active_sms);
  if (st != cudaSuccess) return e->fail(D2T_ERR_CUDA, "contraction launch failed: %s", cudaGetErrorString(st));
  e->launches += 1;
  return 0;
}

ConvGemm linear_params(const float* x, const float* w, const float* bias, float* out, int M, int N, int K) {
  ConvGemm p{};
  p.x = x; p.w = w; p.scale = nullptr; p.shift = bias; p.res = nullptr; p.out = out; p.out2 = nullptr;
  p.dyn = nullptr; p.dyn_mul2 = 0; p.ldc = N; p.ldc2 = 0; p.ldr = N; p.n_split = N;
  p.B = M; p.H = 1; p.W = 1; p.C = K; p.KH = 1; p.KW = 1; p.SH = 1; p.SW = 1; p.PH = 0; p.PW = 0; p.OH = 1; p.OW = 1;
  p.M = M; p.N = N; p.K = K; p.act = ACT_NONE;
  return p;
}

int alloc_act(d2t_engine* e, SlotPool& pool, Fmap* a, int B, int H, int W, int C, bool planes = false, bool f32 = true) {
  cudaError_t st = cudaSuccess;
  a->B = B; a->H = H; a->W = W; a->C = C;
  a->hi = a->lo = nullptr;
  a->p = nullptr;
  if (f32) {
    a->p = (float*)pool.get(a->numel() * sizeof(float), &st);
    if (!a->p) return e->fail(D2T_ERR_CUDA, "cudaMalloc of %zu bytes failed: %s", a->numel() * 4, cudaGetErrorString(st));
  }
  if (planes) {
    a->hi = (__nv_bfloat16*)pool.get(a->numel() * 2, &st);
    if (a->hi && e->cfg.precision == D2T_PREC_BF16X3) a->lo = (__nv_bfloat16*)pool.get(a->numel() * 2, &st);
    if (!a->hi || (e->cfg.precision == D2T_PREC_BF16X3 && !a->lo))
      return e->fail(D2T_ERR_CUDA, "cudaMalloc of activation planes failed: %s", cudaGetErrorString(st));
  }
  return 0;
}

void free_act(d2t_engine* e, SlotPool& pool, Fmap& a) {
  if (!e->keep_taps && a.p) pool.release(a.p);
  if (a.hi) pool.release(a.hi);
  if (a.lo) pool.release(a.lo);
  a.p = nullptr; a.hi = a.lo = nullptr;
}

// the stem convolutions read their input from bf16 planes when the cp.async-fed kernel is on
inline bool stem_planes(const d2t_engine* e) {
  return e->use_tc3 && (e->cfg.precision == D2T_PREC_BF16X3 || e->cfg.precision == D2T_PREC_BF16);
}

// Which representations of a layer's output its consumers read: the fp32 tensor (residual adds, max-pools, taps) and /
// or the bf16 operand planes (the next tensor-core convolution).  The early stem layers are bound by activation traffic
// (ncu: conv0_2 moves 2.6 GB in 1.12 ms at 19 % tensor-pipe activity), so a representation nobody reads is not written.
enum OutNeed { NEED_F32 = 1, NEED_PLANES = 2, NEED_BOTH = 3 };

// true when run_contraction will route this problem to the cp.async-fed kernel, whose epilogue accepts out == nullptr
bool routes_to_tc3(d2t_engine* e, const ConvGemm& p) {
  const int prec = e->cfg.precision;
  if (prec != D2T_PREC_BF16X3 && prec != D2T_PREC_BF16) return false;
  if (!e->use_tc3 || !tc3_supported(p, prec) || !tc_supported(p)) return false;
  auto w = e->tcw.find(p.w);
  auto m = e->tc3.find(p.w);
  return w != e->tcw.end() && w->second.ready && w->second.N == p.N && w->second.K == p.K && m != e->tc3.end() && m->second.ready;
}

// pool = true: the 2x2 / stride-2 max-pool that follows this convolution is fused into its epilogue (tensor-core plane path
// only; *pooled reports whether it was — the caller runs the stand-alone pool kernel otherwise).  y is then the POOLED map.
int conv_layer(d2t_engine* e, const std::string& name, const Fmap& x, Fmap* y, int sh, int sw, int ph, int pw,
               const Fmap* res, int act, cudaStream_t s, int oh_override = -1, int ow_override = -1, int need = NEED_BOTH,
               bool pool = false, bool* pooled = nullptr) {
  auto it = e->conv.find(name);
  if (it == e->conv.end()) return e->fail(D2T_ERR_STATE, "conv '%s' not finalized", name.c_str());
  const ConvW& c = it->second;
  const int OH = oh_override > 0 ? oh_override : (x.H + 2 * ph - c.kh) / sh + 1;
  const int OW = ow_override > 0 ? ow_override : (x.W + 2 * pw - c.kw) / sw + 1;
  const bool planes_mode = stem_planes(e) && name != "patch_embed.proj";
  const bool want_planes = planes_mode && ((need & NEED_PLANES) || !e->lean_acts);
  ConvGemm p{};
  p.x = x.p; p.x_hi = x.hi; p.x_lo = x.lo;
  p.w = c.w; p.scale = c.scale; p.shift = c.shift; p.res = res ? res->p : nullptr;
  p.out2 = nullptr; p.dyn = nullptr; p.dyn_mul2 = 0;
  p.ldc = c.cout; p.ldc2 = 0; p.ldr = c.cout; p.n_split = c.cout;
  p.B = x.B; p.H = x.H; p.W = x.W; p.C = x.C; p.KH = c.kh; p.KW = c.kw; p.SH = sh; p.SW = sw; p.PH = ph; p.PW = pw;
  p.OH = OH; p.OW = OW; p.M = x.B * OH * OW; p.N = c.cout; p.K = c.kh * c.kw * c.cin; p.act = act;
  if (pooled) *pooled = false;
  if (pool && e->fuse_pool && !e->keep_taps && planes_mode && res == nullptr && OH % 2 == 0 && OW % 2 == 0) {
    p.pool = 1;
    if (routes_to_tc3(e, p)) {
      // the pooled map feeds tensor-core convolutions only (block conv1 and the 1x1 downsample read the operand planes)
      if (int rc = alloc_act(e, e->enc_pool, y, x.B, OH / 2, OW / 2, c.cout, true, false)) return rc;
      p.out_hi = y->hi; p.out_lo = y->lo; p.out = y->p;
      if (pooled) *pooled = true;
      return run_contraction(e, p, nullptr, e->cfg.precision, s);
    }
    p.pool = 0;
  }
  if (p.x == nullptr && !routes_to_tc3(e, p))
    return e->fail(D2T_ERR_STATE, "internal: conv '%s' reads an activation that exists as operand planes only", name.c_str());
  // the fp32 copy is dropped only when nobody reads it AND the kernel that will run tolerates its absence
  const bool want_f32 = (need & NEED_F32) || e->keep_taps || !want_planes || !e->lean_acts || !routes_to_tc3(e, p);
  if (int rc = alloc_act(e, e->enc_pool, y, x.B, OH, OW, c.cout, want_planes, want_f32)) return rc;
  p.out_hi = y->hi; p.out_lo = y->lo;
  p.out = y->p;
  if (e->time_conv && name == "layer3.1.conv1" && e->conv_events.size() < 4096) {
    cudaEvent_t a, b;
    CUDA_TRY(e, cudaEventCreate(&a));
    CUDA_TRY(e, cudaEventCreate(&b));
    CUDA_TRY(e, cudaEventRecord(a, s));
    const int rc = run_contraction(e, p, nullptr, e->cfg.precision, s);
    CUDA_TRY(e, cudaEventRecord(b, s));
    e->conv_events.emplace_back(a, b);
    e->conv_flops = 2.0 * p.M * p.N * p.K;
    return rc;
  }
  return run_contraction(e, p, nullptr, e->cfg.precision, s);
}

int pool_layer(d2t_engine* e, const Fmap& x, Fmap* y, int sh, int sw, int ph, int pw, cudaStream_t s) {
  const int OH = (x.H + 2 * ph - 2) / sh + 1, OW = (x.W + 2 * pw - 2) / sw + 1;
  if (int rc = alloc_act(e, e->enc_pool, y, x.B, OH, OW, x.C, stem_planes(e))) return rc;
  const long long total = (long long)y->numel() / 4;
  maxpool2x2_nhwc_kernel<<<grid_for(total, 256, e->active_sms), 256, 0, s>>>(x.p, y->p, x.B, x.H, x.W, x.C, OH, OW, sh, sw, ph, pw,
                                                                           y->hi, y->lo);
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  return 0;
}

void tap(d2t_engine* e, const std::string& name, const Fmap& a, bool tokens = false) {
  if (e->keep_taps) e->taps[name] = Tap{a, tokens};
}

int basic_block(d2t_engine* e, const std::string& name, Fmap& x, cudaStream_t s) {
  // BasicBlock.forward (resnet.py:32-48): conv-bn-relu, conv-bn, (+1x1 conv-bn downsample), add, relu
  Fmap t, ds, o;
  // t feeds conv2 only (operand planes); the downsample branch is read only as the fp32 residual
  if (int rc = conv_layer(e, name + ".conv1", x, &t, 1, 1, 1, 1, nullptr, ACT_RELU, s, -1, -1, NEED_PLANES)) return rc;
  const Fmap* res = &x;
  if (e->conv.count(name + ".downsample.0")) {
    if (int rc = conv_layer(e, name + ".downsample.0", x, &ds, 1, 1, 0, 0, nullptr, ACT_NONE, s, -1, -1, NEED_F32)) return rc;
    res = &ds;
  }
  if (int rc = conv_layer(e, name + ".conv2", t, &o, 1, 1, 1, 1, res, ACT_RELU, s)) return rc;
  free_act(e, e->enc_pool, t);
  if (ds.p || ds.hi) free_act(e, e->enc_pool, ds);
  free_act(e, e->enc_pool, x);
  x = o;
  return 0;
}

int layernorm(d2t_engine* e, const float* x, const float* w, const float* b, float* y, int rows, int D, float eps,
              cudaStream_t s, __nv_bfloat16* y_hi = nullptr, __nv_bfloat16* y_lo = nullptr, int nparts = 1,
              long long part_stride = 0) {
  const int threads = 256, wpb = threads / 32;
  const int grid = (rows + wpb - 1) / wpb;
  cudaError_t st;
  switch (D / 128) {
    case 1: st = launch_kernel(layernorm_kernel<1>, dim3(grid), dim3(threads), 0, s, x, w, b, y, rows, eps, y_hi, y_lo, nparts, part_stride); break;
    case 2: st = launch_kernel(layernorm_kernel<2>, dim3(grid), dim3(threads), 0, s, x, w, b, y, rows, eps, y_hi, y_lo, nparts, part_stride); break;
    case 4: st = launch_kernel(layernorm_kernel<4>, dim3(grid), dim3(threads), 0, s, x, w, b, y, rows, eps, y_hi, y_lo, nparts, part_stride); break;
    case 8: st = launch_kernel(layernorm_kernel<8>, dim3(grid), dim3(threads), 0, s, x, w, b, y, rows, eps, y_hi, y_lo, nparts, part_stride); break;
    default: return e->fail(D2T_ERR_UNSUPPORTED, "LayerNorm width %d unsupported (need 128/256/512/1024)", D);
  }
  if (st != cudaSuccess) return e->fail(D2T_ERR_CUDA, "layernorm launch: %s", cudaGetErrorString(st));
  if (D % 128) return e->fail(D2T_ERR_UNSUPPORTED, "LayerNorm width %d unsupported", D);
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  return 0;
}

// Operand planes of a Linear on the tensor-core plane path (option "vit_planes"): input planes written by the producing
// kernel, output planes for the next Linear; either side may be absent.
struct LinPlanes {
  const __nv_bfloat16 *x_hi = nullptr, *x_lo = nullptr;
  __nv_bfloat16 *out_hi = nullptr, *out_lo = nullptr;
};

int linear(d2t_engine* e, const float* x, const std::string& wkey, const std::string& bkey, float* out, int M, int N,
           int K, int act, const float* res, cudaStream_t s, int w_row_off = 0, const LinPlanes* pl = nullptr) {
  auto wi = e->dev.find(wkey);
  if (wi == e->dev.end()) return e->fail(D2T_ERR_STATE, "weight '%s' not finalized", wkey.c_str());
  const float* bias = nullptr;
  if (!bkey.empty()) {
    auto bi = e->dev.find(bkey);
    if (bi == e->dev.end()) return e->fail(D2T_ERR_STATE, "bias '%s' not finalized", bkey.c_str());
    bias = bi->second + w_row_off;
  }
  ConvGemm p = linear_params(x, wi->second + (size_t)w_row_off * K, bias, out, M, N, K);
  p.act = act; p.res = res; p.ldr = N;
  if (pl && pl->x_hi) {
    // rows as the pixels of ONE image row (B = 1, H = 1, W = M): the TMA im2col load of a 1x1 "convolution" walks 128 rows
    p.x_hi = pl->x_hi; p.x_lo = pl->x_lo; p.out_hi = pl->out_hi; p.out_lo = pl->out_lo;
    p.B = 1; p.W = M; p.OW = M;
    // only the plane kernels tolerate a missing fp32 input / output: refuse instead of dereferencing a null pointer
    if ((x == nullptr || out == nullptr) && !routes_to_tc3(e, p))
      return e->fail(D2T_ERR_STATE, "internal: Linear '%s' on operand planes cannot run on the plane kernels", wkey.c_str());
  }
  return run_contraction(e, p, nullptr, e->cfg.precision, s);
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* d2t_version(void) { return "doc2tex_b200 0.1 (sm_100a)"; }

const char* d2t_last_error(const d2t_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int64_t d2t_launch_count(const d2t_engine* e) { return e ? e->launches : 0; }

int d2t_create(const d2t_config* cfg, int device, d2t_engine** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return D2T_ERR_INVALID; }
  if (cfg->struct_size != (int32_t)sizeof(d2t_config)) {
    g_create_error = "d2t_config.struct_size mismatch (ABI)";
    return D2T_ERR_INVALID;
  }
  int ndev = 0;
  cudaError_t st = cudaGetDeviceCount(&ndev);
  if (st != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(st) + " (this engine has no CPU fallback)";
    return D2T_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return D2T_ERR_INVALID; }
  cudaDeviceProp prop;
  st = cudaGetDeviceProperties(&prop, device);
  if (st != cudaSuccess) { g_create_error = cudaGetErrorString(st); return D2T_ERR_CUDA; }
  if (prop.major != 10) {
    g_create_error = "doc2tex_b200 is built for sm_100a (B200) only; found compute capability " +
                     std::to_string(prop.major) + "." + std::to_string(prop.minor);
    return D2T_ERR_UNSUPPORTED;
  }
  if (cfg->in_channels != 1) { g_create_error = "only input_channel == 1 (grayscale) is supported"; return D2T_ERR_UNSUPPORTED; }
  if (cfg->hidden % 128 || cfg->hidden / cfg->heads != 32) {
    g_create_error = "hidden must be a multiple of 128 with head_dim 32";
    return D2T_ERR_UNSUPPORTED;
  }
  if (cfg->head == D2T_HEAD_TFM && (cfg->dec_heads <= 0 || cfg->hidden / cfg->dec_heads != 32)) {
    g_create_error = "TFM head needs d_model/nhead == 32";
    return D2T_ERR_UNSUPPORTED;
  }
  auto* e = new d2t_engine();
  e->cfg = *cfg;
  e->device = device;
  e->num_sms = prop.multiProcessorCount;
  e->enc_sms = e->active_sms = e->num_sms;
  cudaSetDevice(device);
  if (cudaMallocHost(&e->h_counters, 4 * D2T_MAX_GROUPS * sizeof(int)) != cudaSuccess) {
    g_create_error = "cudaMallocHost failed";
    delete e;
    return D2T_ERR_CUDA;
  }
  int prio_lo = 0, prio_hi = 0;   // decode chain = highest priority: its small kernels get the free SMs first
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (cudaStreamCreateWithPriority(&e->work, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming) != cudaSuccess) {
    g_create_error = "stream/event creation failed";
    delete e;
    return D2T_ERR_CUDA;
  }
  bool side_ok = cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&e->ev_poll, cudaEventDisableTiming) == cudaSuccess;
  for (int g = 1; g < D2T_MAX_GROUPS && side_ok; ++g)
    side_ok = cudaStreamCreateWithPriority(&e->side[g], cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
              cudaEventCreateWithFlags(&e->ev_join[g], cudaEventDisableTiming) == cudaSuccess;
  if (!side_ok) {
    g_create_error = "side stream/event creation failed";
    d2t_destroy(e);
    return D2T_ERR_CUDA;
  }
  // decode attention may need > 48 KB of dynamic shared memory for long encoder memories
  cudaFuncSetAttribute(beam_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  *out = e;
  return D2T_OK;
}

int d2t_destroy(d2t_engine* e) {
  if (!e) return D2T_OK;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (auto& g : e->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  for (auto& ev : e->conv_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  for (auto& ev : e->dec_events) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
  for (void* p : e->owned) cudaFree(p);
  e->enc_pool.destroy();
  e->dec_pool.destroy();
  if (e->h_counters) cudaFreeHost(e->h_counters);
  if (e->work) cudaStreamDestroy(e->work);
  if (e->ev_in) cudaEventDestroy(e->ev_in);
  if (e->ev_out) cudaEventDestroy(e->ev_out);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_poll) cudaEventDestroy(e->ev_poll);
  for (int g = 0; g < D2T_MAX_GROUPS; ++g) {
    if (e->side[g]) cudaStreamDestroy(e->side[g]);
    if (e->ev_join[g]) cudaEventDestroy(e->ev_join[g]);
  }
  delete e;
  return D2T_OK;
}

int d2t_load_tensor(d2t_engine* e, const char* key, const void* data, const int64_t* shape, int ndim, int dtype) {
  if (!e || !key || !data || ndim < 0 || ndim > 8) return e ? e->fail(D2T_ERR_INVALID, "bad load_tensor arguments") : D2T_ERR_INVALID;
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= (size_t)shape[i]; }
  t.f.resize(n);
  if (dtype == D2T_F32) {
    memcpy(t.f.data(), data, n * sizeof(float));
  } else if (dtype == D2T_I64) {
    const int64_t* s = (const int64_t*)data;
    for (size_t i = 0; i < n; ++i) t.f[i] = (float)s[i];
  } else {
    return e->fail(D2T_ERR_INVALID, "unsupported dtype %d for '%s'", dtype, key);
  }
  e->host[key] = std::move(t);
  e->finalized = false;
  return D2T_OK;
}

int d2t_finalize_weights(d2t_engine* e) {
  if (!e) return D2T_ERR_INVALID;
  CUDA_TRY(e, cudaSetDevice(e->device));
  CUDA_TRY(e, cudaDeviceSynchronize());
  for (void* p : e->owned) cudaFree(p);
  e->owned.clear(); e->conv.clear(); e->dev.clear(); e->tcw.clear(); e->tc3.clear();
  e->pos_tables.clear();   // the cached bicubic tables live in `owned` and were resampled from the OLD pos_embed
  drop_graphs(e);
  for (auto& ev : e->conv_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  e->conv_events.clear();
  const d2t_config& c = e->cfg;
  const int C = c.stem_channels, D = c.hidden;
  const int c16 = C / 16, c8 = C / 8, c4 = C / 4, c2 = C / 2;
  int rc;
  // --- ResNet stem (resnet.py:51-156) ---
  if ((rc = make_conv(e, "conv0_1", "bn0_1", c16, c.in_channels, 3, 3))) return rc;
  if ((rc = make_conv(e, "conv0_2", "bn0_2", c8, c16, 3, 3))) return rc;
  if ((rc = make_layer(e, "layer1", c8, c4, 1))) return rc;
  if ((rc = make_conv(e, "conv1", "bn1", c4, c4, 3, 3))) return rc;
  if ((rc = make_layer(e, "layer2", c4, c2, 2))) return rc;
  if ((rc = make_conv(e, "conv2", "bn2", c2, c2, 3, 3))) return rc;
  if ((rc = make_layer(e, "layer3", c2, C, 5))) return rc;
  if ((rc = make_conv(e, "conv3", "bn3", C, C, 3, 3))) return rc;
  if ((rc = make_layer(e, "layer4", C, C, 3))) return rc;
  if ((rc = make_conv(e, "conv4_1", "bn4_1", C, C, 2, 2))) return rc;
  if ((rc = make_conv(e, "conv4_2", "bn4_2", C, C, 2, 2))) return rc;
  // --- patch embed (patchembed.py:111-113): Conv2d(C -> D, k2, s2, bias) ---
  {
    const HostTensor *w, *b;
    if ((rc = need(e, SEQ + "patch_embed.proj.weight", &w, {D, C, 2, 2}))) return rc;
    if ((rc = need(e, SEQ + "patch_embed.proj.bias", &b, {D}))) return rc;
    std::vector<float> packed((size_t)D * 4 * C);
    for (int o = 0; o < D; ++o)
      for (int i = 0; i < C; ++i)
        for (int y = 0; y < 2; ++y)
          for (int x = 0; x < 2; ++x)
            packed[(((size_t)o * 2 + y) * 2 + x) * C + i] = w->f[(((size_t)o * C + i) * 2 + y) * 2 + x];
    ConvW cw; cw.cout = D; cw.cin = C; cw.kh = 2; cw.kw = 2;
    if ((rc = upload(e, packed.data(), packed.size(), &cw.w))) return rc;
    if ((rc = upload(e, b->f.data(), D, &cw.shift))) return rc;
    e->conv["patch_embed.proj"] = cw;
  }
  // --- ViT (vit_encoder.py:229-268, vision_transformer.py:119-122) ---
  if ((rc = upload_key(e, SEQ + "cls_token", {1, 1, D}))) return rc;
  if ((rc = upload_key(e, SEQ + "pos_embed", {1, c.max_tokens, D}))) return rc;
  for (int i = 0; i < c.depth; ++i) {
    const std::string p = SEQ + "blocks." + std::to_string(i) + ".";
    const char* vec256[] = {"norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias", "attn.proj.bias", "mlp.fc2.bias"};
    for (const char* k : vec256) if ((rc = upload_key(e, p + k, {D}))) return rc;
    if ((rc = upload_key(e, p + "attn.qkv.weight", {3 * D, D}))) return rc;
    if ((rc = upload_key(e, p + "attn.qkv.bias", {3 * D}))) return rc;
    if ((rc = upload_key(e, p + "attn.proj.weight", {D, D}))) return rc;
    if ((rc = upload_key(e, p + "mlp.fc1.weight", {4 * D, D}))) return rc;
    if ((rc = upload_key(e, p + "mlp.fc1.bias", {4 * D}))) return rc;
    if ((rc = upload_key(e, p + "mlp.fc2.weight", {D, 4 * D}))) return rc;
  }
  if ((rc = upload_key(e, SEQ + "norm.weight", {D}))) return rc;
  if ((rc = upload_key(e, SEQ + "norm.bias", {D}))) return rc;
  // --- prediction head ---
  if (c.head == D2T_HEAD_TFM) {
    const int V = c.vocab, F = c.dec_ff;
    if ((rc = upload_key(e, PRED + "word_embed.weight", {V, D}))) return rc;
    {
      const HostTensor* pe;
      if ((rc = need(e, PRED + "pos_enc.pe", &pe))) return rc;
      if (pe->shape.size() != 2 || pe->shape[1] != D || pe->shape[0] < c.max_seq_len + 2)
        return e->fail(D2T_ERR_INVALID, "pos_enc.pe must be (>=max_seq_len+2, %d)", D);
      if ((rc = upload_key(e, PRED + "pos_enc.pe"))) return rc;
    }
    for (int l = 0; l < c.dec_layers; ++l) {
      const std::string p = PRED + "model.layers." + std::to_string(l) + ".";
      for (const char* a : {"self_attn.", "multihead_attn."}) {
        if ((rc = upload_key(e, p + a + "in_proj_weight", {3 * D, D}))) return rc;
        if ((rc = upload_key(e, p + a + "in_proj_bias", {3 * D}))) return rc;
        if ((rc = upload_key(e, p + a + "out_proj.weight", {D, D}))) return rc;
        if ((rc = upload_key(e, p + a + "out_proj.bias", {D}))) return rc;
      }
      if ((rc = upload_key(e, p + "linear1.weight", {F, D}))) return rc;
      if ((rc = upload_key(e, p + "linear1.bias", {F}))) return rc;
      if ((rc = upload_key(e, p + "linear2.weight", {D, F}))) return rc;
      if ((rc = upload_key(e, p + "linear2.bias", {D}))) return rc;
      for (const char* n : {"norm1", "norm2", "norm3"}) {
        if ((rc = upload_key(e, p + n + ".weight", {D}))) return rc;
        if ((rc = upload_key(e, p + n + ".bias", {D}))) return rc;
      }
    }
    if ((rc = upload_key(e, PRED + "proj.weight", {V, D}))) return rc;
    if ((rc = upload_key(e, PRED + "proj.bias", {V}))) return rc;
  } else if (c.head == D2T_HEAD_ATTNV2 || c.head == D2T_HEAD_ATTN) {
    const int V = c.vocab, Hs = c.attn_hidden, Kd = c.attn_kernel_dim, taps = 2 * c.attn_kernel_size + 1;
    const std::string a = PRED + "attention_cell.attn.";
    if ((rc = upload_key(e, PRED + "embedding.weight", {V, D}))) return rc;
    if ((rc = upload_key(e, a + "loc_conv.weight", {Kd, 1, taps}))) return rc;
    if ((rc = upload_key(e, a + "loc_conv.bias", {Kd}))) return rc;
    if ((rc = upload_key(e, a + "loc_proj.weight", {Hs, Kd}))) return rc;
    if ((rc = upload_key(e, a + "loc_proj.bias", {Hs}))) return rc;
    if ((rc = upload_key(e, a + "query_proj.weight", {Hs, Hs}))) return rc;
    if ((rc = upload_key(e, a + "query_proj.bias", {Hs}))) return rc;
    if ((rc = upload_key(e, a + "key_proj.weight", {Hs, D}))) return rc;
    if ((rc = upload_key(e, a + "key_proj.bias", {Hs}))) return rc;
    if ((rc = upload_key(e, a + "score.weight", {1, Hs}))) return rc;
    if ((rc = upload_key(e, a + "score.bias", {1}))) return rc;
    const std::string r = PRED + "attention_cell.rnn.";
    // LSTMCell input = [context ; embedding] (attention1D.py:236-239): gates = W_ih x + b_ih + W_hh h + b_hh.
    // Packed once as one [4H, D+D+H] matrix over [context ; embedding ; h] with the two biases kept separate.
    {
      const HostTensor *wih, *whh, *bih, *bhh;
      if ((rc = need(e, r + "weight_ih", &wih, {4 * Hs, 2 * D}))) return rc;
      if ((rc = need(e, r + "weight_hh", &whh, {4 * Hs, Hs}))) return rc;
      if ((rc = need(e, r + "bias_ih", &bih, {4 * Hs}))) return rc;
      if ((rc = need(e, r + "bias_hh", &bhh, {4 * Hs}))) return rc;
      const int Kc = 2 * D + Hs;
      std::vector<float> cat((size_t)4 * Hs * Kc);
      for (int g = 0; g < 4 * Hs; ++g) {
        memcpy(&cat[(size_t)g * Kc], &wih->f[(size_t)g * 2 * D], 2 * D * sizeof(float));
        memcpy(&cat[(size_t)g * Kc + 2 * D], &whh->f[(size_t)g * Hs], Hs * sizeof(float));
      }
      float* p;
      if ((rc = upload(e, cat.data(), cat.size(), &p))) return rc;
      e->dev["lstm.w_cat"] = p;
      if ((rc = upload_key(e, r + "bias_ih"))) return rc;
      if ((rc = upload_key(e, r + "bias_hh"))) return rc;
    }
    if ((rc = upload_key(e, PRED + "attention_cell.generator.weight", {V, Hs}))) return rc;
    if ((rc = upload_key(e, PRED + "attention_cell.generator.bias", {V}))) return rc;
    for (const char* n : {"proj_init_h", "proj_init_c"}) {
      if ((rc = upload_key(e, PRED + n + ".weight", {Hs, D}))) return rc;
      if ((rc = upload_key(e, PRED + n + ".bias", {Hs}))) return rc;
    }
    if ((rc = finalize_attn_extras(e))) return rc;
  }
  // operand planes + TMA maps for the tensor-core contraction path: every conv (but the Cin=1 stem conv) and
  // every Linear, including the row blocks of the packed in_proj matrices that are used on their own
  if (c.precision != D2T_PREC_FP32) {
    auto prep = [&](const float* w, int N, int K) -> int {
      if (!w || e->tcw.count(w)) return 0;
      TcWeight tw;
      cudaError_t st = tc_prepare_weight(w, N, K, c.precision, &tw, &e->owned);
      if (st != cudaSuccess) return e->fail(D2T_ERR_CUDA, "tc_prepare_weight(N=%d,K=%d): %s", N, K, cudaGetErrorString(st));
      e->tcw[w] = tw;
      return 0;
    };
    for (auto& kv : e->conv) {
      if (kv.first == "conv0_1") continue;
      if ((rc = prep(kv.second.w, kv.second.cout, kv.second.kh * kv.second.kw * kv.second.cin))) return rc;
      if (c.precision == D2T_PREC_BF16X3 || c.precision == D2T_PREC_BF16) {
        Tc3Maps m3;
        cudaError_t st = tc3_prepare_maps(e->tcw[kv.second.w], &m3);
        if (st != cudaSuccess) return e->fail(D2T_ERR_CUDA, "tc3_prepare_maps(%s): %s", kv.first.c_str(), cudaGetErrorString(st));
        e->tc3[kv.second.w] = m3;
      }
    }
    for (int i = 0; i < c.depth; ++i) {
      const std::string p = SEQ + "blocks." + std::to_string(i) + ".";
      if ((rc = prep(e->dev[p + "attn.qkv.weight"], 3 * D, D))) return rc;
      if ((rc = prep(e->dev[p + "attn.proj.weight"], D, D))) return rc;
      if ((rc = prep(e->dev[p + "mlp.fc1.weight"], 4 * D, D))) return rc;
      if ((rc = prep(e->dev[p + "mlp.fc2.weight"], D, 4 * D))) return rc;
      if (c.precision == D2T_PREC_BF16X3 || c.precision == D2T_PREC_BF16) {   // option "vit_planes": same kernels as the stem
        for (const char* n : {"attn.qkv.weight", "attn.proj.weight", "mlp.fc1.weight", "mlp.fc2.weight"}) {
          const float* w = e->dev[p + n];
          Tc3Maps m3;
          cudaError_t st = tc3_prepare_maps(e->tcw[w], &m3);
          if (st != cudaSuccess) return e->fail(D2T_ERR_CUDA, "tc3_prepare_maps(%s%s): %s", p.c_str(), n, cudaGetErrorString(st));
          e->tc3[w] = m3;
        }
      }
    }
    if (c.head == D2T_HEAD_TFM) {
      for (int l = 0; l < c.dec_layers; ++l) {
        const std::string p = PRED + "model.layers." + std::to_string(l) + ".";
        if ((rc = prep(e->dev[p + "self_attn.in_proj_weight"], 3 * D, D))) return rc;
        if ((rc = prep(e->dev[p + "self_attn.out_proj.weight"], D, D))) return rc;
        if ((rc = prep(e->dev[p + "multihead_attn.in_proj_weight"], D, D))) return rc;                      // q rows
        if ((rc = prep(e->dev[p + "multihead_attn.in_proj_weight"] + (size_t)D * D, 2 * D, D))) return rc;  // k|v rows
        if ((rc = prep(e->dev[p + "multihead_attn.out_proj.weight"], D, D))) return rc;
        if ((rc = prep(e->dev[p + "linear1.weight"], c.dec_ff, D))) return rc;
        if ((rc = prep(e->dev[p + "linear2.weight"], D, c.dec_ff))) return rc;
      }
      if ((rc = prep(e->dev[PRED + "proj.weight"], c.vocab, D))) return rc;
      if (c.precision == D2T_PREC_BF16X3 || c.precision == D2T_PREC_BF16) {   // option "wide_decode": qkv, lin1 and the vocabulary projection
        std::vector<const float*> ws = {e->dev[PRED + "proj.weight"]};
        for (int l = 0; l < c.dec_layers; ++l) {
          ws.push_back(e->dev[PRED + "model.layers." + std::to_string(l) + ".linear1.weight"]);
          ws.push_back(e->dev[PRED + "model.layers." + std::to_string(l) + ".self_attn.in_proj_weight"]);
        }
        for (const float* w : ws) {
          Tc3Maps m3;
          cudaError_t st = tc3_prepare_maps(e->tcw[w], &m3);
          if (st != cudaSuccess) return e->fail(D2T_ERR_CUDA, "tc3_prepare_maps(decoder): %s", cudaGetErrorString(st));
          e->tc3[w] = m3;
        }
      }
    } else if (c.head == D2T_HEAD_ATTNV2 || c.head == D2T_HEAD_ATTN) {
      const int Hs = c.attn_hidden;
      const std::string a = PRED + "attention_cell.attn.";
      if ((rc = prep(e->dev[a + "key_proj.weight"], Hs, D))) return rc;
      if ((rc = prep(e->dev[a + "query_proj.weight"], Hs, Hs))) return rc;
      if ((rc = prep(e->dev["lstm.w_cat"], 4 * Hs, 2 * D + Hs))) return rc;
      if ((rc = prep(e->dev[PRED + "attention_cell.generator.weight"], c.vocab, Hs))) return rc;
      if ((rc = prep(e->dev[PRED + "proj_init_h.weight"], Hs, D))) return rc;
      if ((rc = prep(e->dev[PRED + "proj_init_c.weight"], Hs, D))) return rc;
    }
  }
  CUDA_TRY(e, cudaDeviceSynchronize());
  e->finalized = true;
  return D2T_OK;
}

int d2t_encoder_geometry(const d2t_engine* e, int H, int W, int* gh, int* gw, int* pad_w, int* pad_h, int* ntok) {
  if (!e) return D2T_ERR_INVALID;
  if (H < 32 || W < 32 || H % 32 || W % 32) return const_cast<d2t_engine*>(e)->fail(D2T_ERR_INVALID, "H and W must be multiples of 32 (got %dx%d)", H, W);
  const int fh = H / 16 - 1, fw = W / 4 + 1;  // quirk Q2
  const int g_h = (fh + 1) / 2, g_w = (fw + 1) / 2;
  if (gh) *gh = g_h;
  if (gw) *gw = g_w;
  if (pad_h) *pad_h = fh % 2;
  if (pad_w) *pad_w = fw % 2;
  if (ntok) *ntok = 1 + g_h * g_w;
  return D2T_OK;
}

int d2t_set_option(d2t_engine* e, const char* key, int value) {
  if (!e || !key) return D2T_ERR_INVALID;
  const std::string k = key;
  bool decode_affecting = true;
  if (k == "encoder_sms") {
    if (value < 8 || value > e->num_sms) return e->fail(D2T_ERR_INVALID, "encoder_sms must be in [8, %d]", e->num_sms);
    e->enc_sms = value;
    decode_affecting = false;
  } else if (k == "decode_groups") {
    if (value < 0 || value > D2T_MAX_GROUPS) return e->fail(D2T_ERR_INVALID, "decode_groups must be in [0, %d]", D2T_MAX_GROUPS);
    e->decode_groups = value;
  } else if (k == "pos_interpolate") {
    e->pos_interpolate = value != 0; decode_affecting = false;
  } else if (k == "pos_grid_h") {
    e->pos_grid_h = value; decode_affecting = false;
  } else if (k == "pos_grid_w") {
    e->pos_grid_w = value; decode_affecting = false;
  } else if (k == "time_conv") {
    e->time_conv = value != 0; decode_affecting = false;
  } else if (k == "time_decode") {
    e->time_decode = value != 0;
  } else if (k == "attn_kpi") {
    e->attn_kpi = value;
  } else if (k == "attn_split") {
    e->attn_split = value;
  } else if (k == "stack_mma") {
    e->stack_mma = value != 0;
  } else if (k == "steps_per_graph") {
    if (value < 1 || value > 16) return e->fail(D2T_ERR_INVALID, "steps_per_graph must be in [1, 16]");
    e->steps_per_graph = value;
  } else if (k == "split_k") {
    e->split_k = value;
  } else if (k == "pdl_max_rows") {
    e->pdl_max_rows = value;
  } else if (k == "pdl") {
    e->use_pdl = value != 0;
  } else if (k == "fuse_pick") {
    e->fuse_pick = value != 0;
  } else if (k == "kv_bf16") {
    e->kv_bf16 = value != 0;
  } else if (k == "lean_acts") {
    e->lean_acts = value != 0; decode_affecting = false;
  } else if (k == "wide_decode") {
    e->wide_decode = value != 0;
  } else if (k == "vit_planes") {
    e->vit_planes = value; decode_affecting = false;
  } else if (k == "fuse_pool") {
    e->fuse_pool = value != 0; decode_affecting = false;
  } else if (k == "pair") {
    e->use_pair = value; decode_affecting = false;
  } else if (k == "tma_a") {
    tc3_use_a_tma() = value != 0; decode_affecting = false;
  } else if (k == "tc3") {
    e->use_tc3 = value != 0; decode_affecting = false;
  } else if (k == "dbg_decode") {
    e->dbg_decode = value != 0;
  } else if (k == "dbg_timeline") {
    e->dbg_timeline = value != 0;
  } else {
    return e->fail(D2T_ERR_INVALID, "unknown option '%s'", key);
  }
  if (decode_affecting) {   // a captured step graph replays the OLD kernel selection
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    drop_graphs(e);
  }
  return D2T_OK;
}

int d2t_debug_conv_time(d2t_engine* e, double* total_ms, int64_t* launches, double* flops_per_launch) {
  if (!e || !total_ms || !launches || !flops_per_launch) return D2T_ERR_INVALID;
  CUDA_TRY(e, cudaSetDevice(e->device));
  CUDA_TRY(e, cudaDeviceSynchronize());
  double tot = 0.0;
  for (auto& ev : e->conv_events) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) tot += ms;
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  *total_ms = tot;
  *launches = (int64_t)e->conv_events.size();
  *flops_per_launch = e->conv_flops;
  e->conv_events.clear();
  return D2T_OK;
}

int d2t_debug_decode_time(d2t_engine* e, int kind, double* total_ms, int64_t* launches, double* total_bytes) {
  if (!e || !total_ms || !launches || !total_bytes) return D2T_ERR_INVALID;
  CUDA_TRY(e, cudaSetDevice(e->device));
  CUDA_TRY(e, cudaDeviceSynchronize());
  double ms_sum = 0.0, bytes = 0.0;
  int64_t n = 0;
  std::vector<d2t_engine::DecEvent> keep;
  for (auto& ev : e->dec_events) {
    if (ev.kind != kind) { keep.push_back(ev); continue; }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev.a, ev.b) == cudaSuccess) { ms_sum += ms; bytes += ev.bytes; ++n; }
    cudaEventDestroy(ev.a);
    cudaEventDestroy(ev.b);
  }
  e->dec_events.swap(keep);
  *total_ms = ms_sum; *launches = n; *total_bytes = bytes;
  return D2T_OK;
}

int d2t_debug_beam_runner_up(d2t_engine* e, float* runner_up_dev) {
  if (!e) return D2T_ERR_INVALID;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  drop_graphs(e);   // the pointer is baked into the captured beam step
  e->beam_runner_up = runner_up_dev;
  return D2T_OK;
}

int d2t_set_debug(d2t_engine* e, int keep_taps) {
  if (!e) return D2T_ERR_INVALID;
  e->keep_taps = keep_taps != 0;
  return D2T_OK;
}

int d2t_encode(d2t_engine* e, const float* img, int B, int H, int W, float* ctx, d2t_stream stream) {
  if (!e) return D2T_ERR_INVALID;
  if (!e->finalized) return e->fail(D2T_ERR_STATE, "d2t_encode before d2t_finalize_weights");
  if (!img || !ctx || B <= 0) return e->fail(D2T_ERR_INVALID, "bad encode arguments");
  int gh, gw, ntok;
  if (int rc = d2t_encoder_geometry(e, H, W, &gh, &gw, nullptr, nullptr, &ntok)) return rc;
  if (ntok > e->cfg.max_tokens)
    return e->fail(D2T_ERR_INVALID, "image %dx%d needs %d tokens but pos_embed has %d rows (max_dimension)", H, W, ntok, e->cfg.max_tokens);
  CUDA_TRY(e, cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  e->active_sms = e->enc_sms;
  e->enc_pool.release_all();
  e->taps.clear();
  const d2t_config& c = e->cfg;
  const int D = c.hidden;
  int rc;

  // ---- ResNet stem, NHWC (resnet.py:205-245) ----
  Fmap x, y;
  {
    const ConvW& c0 = e->conv["conv0_1"];
    // conv0_1 feeds conv0_2 only: operand planes, no fp32 copy (unless taps are kept)
    if ((rc = alloc_act(e, e->enc_pool, &x, B, H, W, c0.cout, stem_planes(e), !(stem_planes(e) && e->lean_acts) || e->keep_taps))) return rc;
    const long long total = (long long)B * H * W * (c0.cout / 4);
    const size_t smem = (size_t)11 * c0.cout * sizeof(float);
    if (c0.cout % 8 == 0 && W % 4 == 0 && 256 % (c0.cout / 8) == 0) {
      // four pixels x eight channels per thread; the grid stride (a multiple of 256) keeps a thread on its channel group
      const long long groups = (long long)B * H * (W / 4) * (c0.cout / 8);
      conv0_direct4x8_kernel<<<grid_for(groups, 256, e->active_sms), 256, 0, s>>>(img, c0.w, c0.scale, c0.shift, x.p, B, H, W, c0.cout, x.hi, x.lo);
    } else
    conv0_direct_kernel<<<grid_for(total, 256, e->active_sms), 256, smem, s>>>(img, c0.w, c0.scale, c0.shift, x.p, B, H, W, c0.cout, x.hi, x.lo);
    e->launches += 1;
    CUDA_TRY(e, cudaGetLastError());
    tap(e, "conv0_1", x);
  }
  bool pooled = false;
  // conv0_2 + max-pool 1 (resnet.py:214-217): fused in the epilogue on the plane path, else the map is read by the pool kernel
  if ((rc = conv_layer(e, "conv0_2", x, &y, 1, 1, 1, 1, nullptr, ACT_RELU, s, -1, -1, NEED_F32, true, &pooled))) return rc;
  free_act(e, e->enc_pool, x);
  if (pooled) {
    x = y; y = Fmap{};
  } else {
    tap(e, "conv0_2", y);
    if ((rc = pool_layer(e, y, &x, 2, 2, 0, 0, s))) return rc;
    free_act(e, e->enc_pool, y);
  }
  if ((rc = basic_block(e, "layer1.0", x, s))) return rc;
  tap(e, "layer1", x);
  if ((rc = conv_layer(e, "conv1", x, &y, 1, 1, 1, 1, nullptr, ACT_RELU, s, -1, -1, NEED_F32, true, &pooled))) return rc;   // + max-pool 2
  free_act(e, e->enc_pool, x);
  if (pooled) {
    x = y; y = Fmap{};
  } else {
    tap(e, "conv1", y);
    if ((rc = pool_layer(e, y, &x, 2, 2, 0, 0, s))) return rc;
    free_act(e, e->enc_pool, y);
  }
  for (int b = 0; b < 2; ++b) if ((rc = basic_block(e, "layer2." + std::to_string(b), x, s))) return rc;
  tap(e, "layer2", x);
  if ((rc = conv_layer(e, "conv2", x, &y, 1, 1, 1, 1, nullptr, ACT_RELU, s, -1, -1, NEED_F32))) return rc;
  free_act(e, e->enc_pool, x); tap(e, "conv2", y);
  if ((rc = pool_layer(e, y, &x, 2, 1, 0, 1, s))) return rc;  // maxpool3: k2 s(2,1) p(0,1), -inf padding
  free_act(e, e->enc_pool, y); tap(e, "pool3", x);
  for (int b = 0; b < 5; ++b) if ((rc = basic_block(e, "layer3." + std::to_string(b), x, s))) return rc;
  tap(e, "layer3", x);
  if ((rc = conv_layer(e, "conv3", x, &y, 1, 1, 1, 1, nullptr, ACT_RELU, s))) return rc;
  free_act(e, e->enc_pool, x); tap(e, "conv3", y);
  x = y; y = Fmap{};
  for (int b = 0; b < 3; ++b) if ((rc = basic_block(e, "layer4." + std::to_string(b), x, s))) return rc;
  tap(e, "layer4", x);
  if ((rc = conv_layer(e, "conv4_1", x, &y, 2, 1, 0, 1, nullptr, ACT_RELU, s))) return rc;
  free_act(e, e->enc_pool, x); tap(e, "conv4_1", y);
  if ((rc = conv_layer(e, "conv4_2", y, &x, 1, 1, 0, 0, nullptr, ACT_RELU, s))) return rc;
  free_act(e, e->enc_pool, y); tap(e, "conv4_2", x);
  if (x.H != H / 16 - 1 || x.W != W / 4 + 1) return e->fail(D2T_ERR_INVALID, "internal: stem geometry %dx%d", x.H, x.W);

  // ---- HybridEmbed: zero-pad right/bottom to even, Conv2d(k2,s2)+bias, flatten (patchembed.py:121-135) ----
  Fmap tok;
  if ((rc = conv_layer(e, "patch_embed.proj", x, &tok, 2, 2, 0, 0, nullptr, ACT_NONE, s, gh, gw))) return rc;
  free_act(e, e->enc_pool, x);
  {
    Fmap t = tok; t.B = B; t.H = 1; t.W = gh * gw; tap(e, "patch_embed", t, true);
  }
  const int N = gh * gw, T = N + 1, rows = B * T;
  Fmap xs, hs, qkv, att, x2, ff;
  // Plane path (option "vit_planes", tensor-core precisions): LayerNorm, the attention kernel and fc1's GELU epilogue write
  // bf16 hi/lo operand planes and no fp32 copy (nobody reads one), and the four Linears of a block run on the kernels of the
  // stem — both operands by TMA, CTA pair where the tile count fills the machine — instead of gathering fp32 through registers.
  bool vp = stem_planes(e) && e->vit_planes && D % 64 == 0;
  for (int i = 0; i < c.depth && vp; ++i)
    for (const char* n : {"attn.qkv.weight", "attn.proj.weight", "mlp.fc1.weight", "mlp.fc2.weight"}) {
      auto m3 = e->tc3.find(e->dev[SEQ + "blocks." + std::to_string(i) + "." + n]);
      if (m3 == e->tc3.end() || !m3->second.ready) vp = false;
    }
  if ((rc = alloc_act(e, e->enc_pool, &xs, B, 1, T, D))) return rc;
  if ((rc = alloc_act(e, e->enc_pool, &hs, B, 1, T, D, vp, !vp))) return rc;
  if ((rc = alloc_act(e, e->enc_pool, &qkv, B, 1, T, 3 * D))) return rc;
  if ((rc = alloc_act(e, e->enc_pool, &att, B, 1, T, D, vp, !vp))) return rc;
  if ((rc = alloc_act(e, e->enc_pool, &ff, B, 1, T, 4 * D, vp, !vp))) return rc;
  LinPlanes from_hs, from_att, to_ff, from_ff;
  if (vp) {
    from_hs.x_hi = hs.hi; from_hs.x_lo = hs.lo;
    from_att.x_hi = att.hi; from_att.x_lo = att.lo;
    to_ff = from_hs; to_ff.out_hi = ff.hi; to_ff.out_lo = ff.lo;
    from_ff.x_hi = ff.hi; from_ff.x_lo = ff.lo;
  }
  const LinPlanes* const p_hs = vp ? &from_hs : nullptr;
  const LinPlanes* const p_att = vp ? &from_att : nullptr;
  const LinPlanes* const p_toff = vp ? &to_ff : nullptr;
  const LinPlanes* const p_ff = vp ? &from_ff : nullptr;
  const float* pos = e->dev[SEQ + "pos_embed"];
  if (e->pos_interpolate && (gh != e->pos_grid_h || gw != e->pos_grid_w)) {
    if (e->pos_grid_h <= 0 || e->pos_grid_w <= 0 || 1 + e->pos_grid_h * e->pos_grid_w != c.max_tokens)
      return e->fail(D2T_ERR_STATE, "pos_interpolate needs pos_grid_h x pos_grid_w = the rows of pos_embed - 1");
    auto it = e->pos_tables.find({gh, gw});
    if (it == e->pos_tables.end()) {
      float* tbl = nullptr;
      CUDA_TRY(e, cudaMalloc(&tbl, (size_t)T * D * sizeof(float)));
      e->owned.push_back(tbl);
      // torch: scale_factor = (h0 + 0.1) / emb_height with h0 = grid rows; source index uses 1 / scale_factor
      const float sfh = (float)((gh + 0.1) / e->pos_grid_h), sfw = (float)((gw + 0.1) / e->pos_grid_w);
      pos_embed_bicubic_kernel<<<grid_for((long long)T * D, 256, e->num_sms), 256, 0, s>>>(
          pos, e->pos_grid_h, e->pos_grid_w, tbl, gh, gw, 1.0f / sfh, 1.0f / sfw, D);
      e->launches += 1;
      CUDA_TRY(e, cudaGetLastError());
      it = e->pos_tables.emplace(std::make_pair(gh, gw), tbl).first;
    }
    pos = it->second;
  }
  assemble_tokens_kernel<<<grid_for((long long)rows * D / 4, 256, e->active_sms), 256, 0, s>>>(
      tok.p, e->dev[SEQ + "cls_token"], pos, xs.p, B, N, D);
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  free_act(e, e->enc_pool, tok);
  const float scale = 1.0f / sqrtf((float)(D / c.heads));
  for (int i = 0; i < c.depth; ++i) {
    const std::string p = SEQ + "blocks." + std::to_string(i) + ".";
    if ((rc = alloc_act(e, e->enc_pool, &x2, B, 1, T, D))) return rc;
    if ((rc = layernorm(e, xs.p, e->dev[p + "norm1.weight"], e->dev[p + "norm1.bias"], hs.p, rows, D, 1e-6f, s, hs.hi, hs.lo))) return rc;
    if ((rc = linear(e, hs.p, p + "attn.qkv.weight", p + "attn.qkv.bias", qkv.p, rows, 3 * D, D, ACT_NONE, nullptr, s, 0, p_hs))) return rc;
    {
      dim3 grid((T + ENC_ATT_QT - 1) / ENC_ATT_QT, B * c.heads);
      encoder_attention_kernel<32><<<grid, 4 * ENC_ATT_QT, 0, s>>>(qkv.p, att.p, T, D, scale, att.hi, att.lo);
      e->launches += 1;
      CUDA_TRY(e, cudaGetLastError());
    }
    if ((rc = linear(e, att.p, p + "attn.proj.weight", p + "attn.proj.bias", x2.p, rows, D, D, ACT_NONE, xs.p, s, 0, p_att))) return rc;
    if ((rc = layernorm(e, x2.p, e->dev[p + "norm2.weight"], e->dev[p + "norm2.bias"], hs.p, rows, D, 1e-6f, s, hs.hi, hs.lo))) return rc;
    if ((rc = linear(e, hs.p, p + "mlp.fc1.weight", p + "mlp.fc1.bias", ff.p, rows, 4 * D, D, ACT_GELU, nullptr, s, 0, p_toff))) return rc;
    free_act(e, e->enc_pool, xs);
    if ((rc = alloc_act(e, e->enc_pool, &xs, B, 1, T, D))) return rc;
    if ((rc = linear(e, ff.p, p + "mlp.fc2.weight", p + "mlp.fc2.bias", xs.p, rows, D, 4 * D, ACT_NONE, x2.p, s, 0, p_ff))) return rc;
    free_act(e, e->enc_pool, x2);
    tap(e, "block" + std::to_string(i), xs, true);
  }
  if ((rc = layernorm(e, xs.p, e->dev[SEQ + "norm.weight"], e->dev[SEQ + "norm.bias"], ctx, rows, D, 1e-6f, s))) return rc;
  return D2T_OK;
}

int d2t_prep_measure(d2t_engine* e, const uint8_t* packed, const d2t_prep_image* imgs, int n, int32_t* stats, d2t_stream stream) {
  if (!e) return D2T_ERR_INVALID;
  if (!packed || !imgs || !stats || n <= 0) return e->fail(D2T_ERR_INVALID, "bad d2t_prep_measure arguments");
  CUDA_TRY(e, cudaSetDevice(e->device));
  prep_stats_kernel<<<n, 1024, 0, (cudaStream_t)stream>>>(packed, imgs, stats);
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  return D2T_OK;
}

int d2t_prep_render(d2t_engine* e, const uint8_t* packed, const d2t_prep_image* imgs, const d2t_prep_plan* plans, int n,
                    const int32_t* coefs, uint8_t* scratch, int any_resize, float sub, float mul, d2t_stream stream) {
  if (!e) return D2T_ERR_INVALID;
  if (!packed || !imgs || !plans || !scratch || n <= 0) return e->fail(D2T_ERR_INVALID, "bad d2t_prep_render arguments");
  if (any_resize && !coefs) return e->fail(D2T_ERR_INVALID, "d2t_prep_render: resize requested without coefficient tables");
  CUDA_TRY(e, cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  // grid.x sized for the largest image the YAML surface allows (448 x 960) at 4 pixels per thread; grid-stride beyond
  const dim3 grid(e->num_sms >= 64 ? 420 : 128, n);
  prep_crop_kernel<<<grid, 256, 0, s>>>(packed, imgs, plans, scratch);
  e->launches += 1;
  if (any_resize) {
    prep_resample_kernel<<<grid, 256, 0, s>>>(plans, coefs, scratch, 0);
    prep_resample_kernel<<<grid, 256, 0, s>>>(plans, coefs, scratch, 1);
    e->launches += 2;
  }
  prep_finish_kernel<<<grid, 256, 0, s>>>(plans, scratch, sub, mul);
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  return D2T_OK;
}

int d2t_debug_tap(d2t_engine* e, const char* name, float* out, int64_t* numel, int64_t* shape4, d2t_stream stream) {
  if (!e || !name) return D2T_ERR_INVALID;
  auto it = e->taps.find(name);
  if (it == e->taps.end()) return e->fail(D2T_ERR_INVALID, "no tap '%s' (enable d2t_set_debug before encode)", name);
  const Fmap& a = it->second.a;
  if (numel) *numel = (int64_t)a.numel();
  if (shape4) {
    if (it->second.tokens) { shape4[0] = a.B; shape4[1] = a.W; shape4[2] = a.C; shape4[3] = 0; }
    else { shape4[0] = a.B; shape4[1] = a.C; shape4[2] = a.H; shape4[3] = a.W; }
  }
  if (!out) return D2T_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (it->second.tokens) {
    CUDA_TRY(e, cudaMemcpyAsync(out, a.p, a.numel() * sizeof(float), cudaMemcpyDeviceToDevice, s));
  } else {
    nhwc_to_nchw_kernel<<<grid_for((long long)a.numel(), 256, e->num_sms), 256, 0, s>>>(a.p, out, a.B, a.H, a.W, a.C);
    CUDA_TRY(e, cudaGetLastError());
  }
  return D2T_OK;
}

int d2t_debug_gemm(d2t_engine* e, const float* a, const float* w, const float* scale, const float* shift, float* c,
                   int M, int N, int K, int act, int precision, d2t_stream stream) {
  if (!e) return D2T_ERR_INVALID;
  CUDA_TRY(e, cudaSetDevice(e->device));
  e->active_sms = e->num_sms;
  ConvGemm p = linear_params(a, w, shift, c, M, N, K);
  p.scale = scale; p.act = act;
  if (precision == D2T_PREC_FP32) return run_contraction(e, p, nullptr, precision, (cudaStream_t)stream);
  TcWeight tw;
  std::vector<void*> tmp;
  cudaError_t st = tc_prepare_weight(w, N, K, precision, &tw, &tmp);
  if (st != cudaSuccess) return e->fail(D2T_ERR_CUDA, "tc_prepare_weight: %s", cudaGetErrorString(st));
  cudaDeviceSynchronize();
  if (!tw.ready || !tc_supported(p)) {
    for (void* q : tmp) cudaFree(q);
    return e->fail(D2T_ERR_UNSUPPORTED, "tcgen05 path does not support M=%d N=%d K=%d", M, N, K);
  }
  int rc = run_contraction(e, p, &tw, precision, (cudaStream_t)stream);
  cudaStreamSynchronize((cudaStream_t)stream);
  for (void* q : tmp) cudaFree(q);
  return rc;
}

// Micro-benchmark hook: `iters` back-to-back launches of one contraction (weights prepared once), optionally
// interleaved with a small LayerNorm launch (the decode-step pattern); returns the average ms per iteration.
int d2t_debug_gemm_bench(d2t_engine* e, const float* a, const float* w, float* c, int M, int N, int K, int precision,
                         int iters, int interleave, float* ms_out, d2t_stream stream) {
  if (!e || !ms_out) return D2T_ERR_INVALID;
  CUDA_TRY(e, cudaSetDevice(e->device));
  cudaStream_t s = e->work;
  e->active_sms = e->num_sms;
  ConvGemm p = linear_params(a, w, nullptr, c, M, N, K);
  long long* dbg_dev = nullptr;
  cudaMalloc(&dbg_dev, 16 * sizeof(long long));
  cudaMemset(dbg_dev, 0, 16 * sizeof(long long));
  p.dbg = dbg_dev;
  if (const char* v = getenv("D2T_DBG_ACT")) p.act = atoi(v);
  TcWeight tw;
  std::vector<void*> tmp;
  if (precision != D2T_PREC_FP32) {
    cudaError_t st = tc_prepare_weight(w, N, K, precision, &tw, &tmp);
    if (st != cudaSuccess) return e->fail(D2T_ERR_CUDA, "tc_prepare_weight: %s", cudaGetErrorString(st));
    cudaDeviceSynchronize();
  }
  float *lnw = nullptr, *lnb = nullptr;
  cudaMalloc(&lnw, 256 * 4); cudaMalloc(&lnb, 256 * 4);
  cudaMemset(lnw, 0, 256 * 4); cudaMemset(lnb, 0, 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int rc = 0;
  for (int pass = 0; pass < 2 && !rc; ++pass) {
    cudaEventRecord(e0, s);
    for (int i = 0; i < iters && !rc; ++i) {
      rc = run_contraction(e, p, precision != D2T_PREC_FP32 ? &tw : nullptr, precision, s);
      if (interleave && !rc && N % 128 == 0 && N <= 1024) rc = layernorm(e, c, lnw, lnb, c, M, N > 256 ? 256 : N, 1e-5f, s);
    }
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  *ms_out = ms / iters;
  if (precision != D2T_PREC_FP32) {
    long long h[12];
    cudaMemcpy(h, dbg_dev, sizeof h, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tc dbg epi chunk0] ld start +%lld, ld done +%lld, staged +%lld, stored +%lld\n", h[8] - h[0],
            h[9] - h[0], h[10] - h[0], h[11] - h[0]);
    fprintf(stderr, "[tc dbg M=%d N=%d K=%d] prologue %lld ns, first full +%lld, last full +%lld, last commit +%lld, "
                    "epi start +%lld, epi done +%lld, exit +%lld\n", M, N, K, h[1] - h[0], h[2] - h[0], h[3] - h[0],
            h[4] - h[0], h[5] - h[0], h[6] - h[0], h[7] - h[0]);
  }
  cudaFree(dbg_dev);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(lnw); cudaFree(lnb);
  for (void* q : tmp) cudaFree(q);
  return rc;
}

}  // extern "C"

#include "decode_host.inl"
