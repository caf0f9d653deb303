// The resident-state engine behind the C ABI (include/conan_b200.h).
//
// Resident state lives in HBM as a slab of per-slot tensors (struct-of-arrays, `max_slots` entries):
//   * Emformer: K|V ring per layer [slot, ring_rows, 2D] + past_len[slot]
//   * Conan / vocoder: the H = (k-1)*dil history rows in front of every causal conv, [slot, H, C]
//               (fp32 for Conan, fp16 = tensor-core operand type or fp32 for the vocoder),
//               per-session style vector and aligner K/V cache
// A chunk step works on COMPACT buffers indexed by the position i of a stream in the ready list:
// every causal conv reads a context buffer [i, H + L, C]; one gather launch per sub-model copies the
// H history rows of slot_ids[i] in front of the L rows this chunk produces, every layer then runs
// as one launch over contiguous rows (so a tile's operand is one TMA box, whatever slots are ready),
// and one scatter launch writes the last H rows back to the slots.  Gather + scatter move exactly the
// bytes an in-place ring shift would.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <cmath>
#include <string>
#include <unordered_map>
#include <vector>

#include "kernels.cuh"

namespace conan {

static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};
void set_error(const std::string& msg) { g_last_error = msg; }
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

struct Ctx {                 // compact context buffer [i, H + L + R, C]; resident history [slot, H, C]
  void* p = nullptr;
  void* hist = nullptr;
  int H = 0, L = 0, R = 0, C = 0;
  int is_half = 0;           // 0 fp32, 1 fp16, 2 split fp16 pair (two planes: hi, lo)
  long long plane = 0;       // elements between the hi and the lo plane of a split buffer
  long long hist_plane = 0;
  int rows() const { return H + L + R; }
  long long slot_stride() const { return (long long)rows() * C; }
  size_t elem() const { return is_half ? 2 : 4; }
  RowView new_rows() const { return RowView{p, slot_stride(), C, H, is_half, plane}; }
  void* at_row(int r) const { return (char*)p + (size_t)r * C * elem(); }
};

struct WeightSlot { std::string name; size_t numel; int dtype; const void* ptr; };

}  // namespace conan

using namespace conan;

struct conan_engine {
  conan_config_t cfg;
  std::vector<WeightSlot> weights;
  std::unordered_map<std::string, int> windex;
  bool finalized = false;
  std::vector<void*> allocs;
  size_t state_bytes = 0;
  int S = 0, tp_max = 0;     // S: slots allocated = max_slots + kPadSlots
  int Su = 0;                // slots the caller may use (cfg.max_slots); the rest pad ready lists to graph bucket sizes
  int* padIds = nullptr;     // device: ids of the pad slots
  bool lin_tc = false;       // Emformer / Conan contractions on tcgen05 with split-fp16 operands
  bool ses_tc = false;       // session setup: the style encoder's ConvBlocks (95 % of the setup FLOPs) on tcgen05, split-fp16 operands
  int DP = 0, QP = 0, LP = 0;   // (padded) Emformer model dim, QKV width, logits width
  // ---- Emformer
  int ring_rows = 0;
  float *eX = nullptr, *eQKV = nullptr, *eR1 = nullptr, *eR2 = nullptr, *eLOG = nullptr;
  Ctx eXN, eATT, eFN, eHF;   // GEMM operands (fp32 or split fp16)
  std::vector<float> postTapsHost;                 // host copy of conv_post's taps + bias (kernel-parameter fast path)
  bool ffnFused = false; float* eFFP = nullptr;    // fused FFN: partial outputs [split][rows][DP]
  bool blockFused = false; float* bfP = nullptr; long long bfRows = 0;   // fused Conan blocks: partial outputs [split][rows][256]
  bool blockCluster = true;      // fused Conan blocks sum their hidden slices through distributed shared memory (cluster of 4 CTAs)
  std::vector<float*> eRing;
  // generic step (memory bank M > 0, summary query, partial segments): its own work buffers with seg + rc + 1 rows per stream
  int eM = 0;                                   // max_memory_size
  float *gX = nullptr, *gQKV = nullptr, *gR1 = nullptr, *gR2 = nullptr, *gXNf = nullptr, *gMKV = nullptr, *gMEM[2] = {nullptr, nullptr};
  Ctx gXN, gATT, gFN, gHF, gMB;
  std::vector<float*> eBank;                    // [layer][slot, M, D] resident
  int* ePast = nullptr;
  int* TOK = nullptr;
  // ---- Conan chunk path
  Ctx cC, cUV[5], cD[8][2], cP;
  Ctx cPOSTO;                                      // tensor-core path: post-conv rows as the split operand of the mel projection
  float* dMELP = nullptr; int MP = 0;              // mel rows padded to MP = 96 columns (the projection's padded N)
  bool melDirect = false;                          // the mel projection also writes the vocoder's input rows (vPRE) in its epilogue
  Ctx cX0, cATT, cO1, cHF, cPROS[2], cDECH;     // GEMM operands without history (fp32 or split fp16)
  float *dX0 = nullptr, *dQ = nullptr, *dT1 = nullptr, *dO1 = nullptr, *dT2 = nullptr,
        *dPROS[2] = {nullptr, nullptr}, *dPINP = nullptr, *dUVH = nullptr, *dDECX = nullptr, *dPOST = nullptr,
        *dMEL = nullptr, *dUVP = nullptr, *dMASK0 = nullptr, *dMASKB = nullptr;
  float *sSTYLE = nullptr, *sKV = nullptr, *sKPM = nullptr;
  int* sNKEYS = nullptr;
  // ---- vocoder
  Ctx vPRE, vUP[8], vXA[8], vC1[8][4][4], vC2[8][4][4], vPOST;
  float *vXS = nullptr, *vXR[2] = {nullptr, nullptr}, *vSUM = nullptr;
  __half* vSUMh = nullptr;       // running MRF sum as fp16 (residual-from-context mode)
  // fused residual blocks (resblock_fused.cu): per (scale, block) packed weights / biases and the resident history block
  bool vFused[8] = {false, false, false, false, false, false, false, false};
  __half* fW[8][4] = {}; float* fB[8][4] = {}; __half* fHist[8][4] = {}; int fHistRows[8][4] = {};
  __half* fHistOut[8][4] = {};                     // two-lane fused kernel: compact staging of the new history (scattered with the contexts)
  int vL[9], vC[9];     // rows / channels entering scale i (vL[0] = segment, vC[0] = initial channel)
  // ---- history gather/scatter tables
  HistDesc* histConan = nullptr; int nHistConan = 0;
  HistDesc* histVoc = nullptr; int nHistVoc = 0;
  float* sSTYLEW = nullptr;      // compact copy of the style vectors of the ready streams
  ZeroDesc* zeroEmf = nullptr; int nZeroEmf = 0;
  ZeroDesc* zeroConan = nullptr; int nZeroConan = 0;
  ZeroDesc* zeroVoc = nullptr; int nZeroVoc = 0;
  // ---- session scratch (compact index)
  int SB = 0;
  __half *qC31h = nullptr, *qHGh = nullptr, *qC3Gh = nullptr;   // split-fp16 operands of the style encoder's tensor-core path
  float* qMAp = nullptr;                                          // the reference-frame mask with the padded row stride of that path
  float *qMA = nullptr, *qMF = nullptr, *qXG = nullptr, *qC31 = nullptr, *qHG = nullptr, *qC3G = nullptr, *qPG = nullptr,
        *qXW = nullptr, *qCW = nullptr, *qAW = nullptr, *qACT = nullptr, *qRS = nullptr, *qSKIP = nullptr, *qGRP = nullptr,
        *qMP = nullptr, *qMPB = nullptr, *qXP = nullptr, *qC5 = nullptr, *qHP = nullptr, *qC3P = nullptr, *qPZ = nullptr,
        *qXE = nullptr, *qZC = nullptr, *qPE = nullptr, *qKVs = nullptr, *qMGB = nullptr;
  int* qVQ = nullptr;
  int* qSlots = nullptr;
  // ---- optional per-launch event timing of the conv engines (bench.py's roofline leg)
  bool profiling = false;
  struct ProfRec { cudaEvent_t a, b; int cat; double flops, bytes; };
  mutable std::vector<ProfRec> prof;
  // ---- host-call staging
  int* hIds = nullptr; float* hChunk = nullptr; float* hWav = nullptr; float* hMel = nullptr; int* hTok = nullptr; int* hIdsSmall = nullptr;
  // pipelined host stepping (conan_step_host_submit / _wait): second set of device outputs, a copy stream, per-ticket events
  float* hWav2 = nullptr; float* hMel2 = nullptr; int* hTok2 = nullptr;
  cudaStream_t copyStream = nullptr; cudaEvent_t evCompute[2] = {nullptr, nullptr}, evCopy[2] = {nullptr, nullptr};
  unsigned long long submitCount = 0; bool ticketPending[2] = {false, false};
  std::vector<uint8_t> idSeen;     // host-side duplicate check of a ready list
  // the three MRF branches of a vocoder scale (kernel sizes 3 / 7 / 11) are independent chains of six convs until their outputs are
  // summed: they run on three streams (fork after the upsampling conv, join through the running sum), so one branch's tail wave,
  // launch latency and prologue run under another branch's MMAs
  cudaStream_t branchStream[2] = {nullptr, nullptr};
  cudaEvent_t evFork = nullptr, evBranch[4] = {nullptr, nullptr, nullptr, nullptr};
  int vocStreams = 3;
  // CUDA graphs of whole chunk steps: a step's ~130 launches (and the fork / join above) are captured once per distinct
  // (ready count, buffer set) on an engine-owned stream and replayed with one cudaGraphLaunch on the caller's stream.  The ready
  // list itself is NOT baked in: kernels read the slot ids from the device buffer the graph was captured with (slot indirection).
  struct StepGraph { int n; const void* ids; const void* chunk; void* wav; void* mel; void* tok; cudaGraphExec_t exec; uint64_t launches; uint64_t last_use; };
  std::vector<StepGraph> stepGraphs;
  std::vector<StepGraph> stepSeen;       // keys met once (the first eager run doubles as warm-up of attributes / tensor maps)
  cudaStream_t graphStream = nullptr;
  int graphMode = 0;                     // 0 off, 1 on
  uint64_t graphClock = 0, graphReplays = 0;

  const WeightSlot* W(const std::string& name) const {
    auto it = windex.find(name);
    return it == windex.end() ? nullptr : &weights[it->second];
  }
  const float* F(const std::string& name) const { auto* w = W(name); return w ? (const float*)w->ptr : nullptr; }
  const void* P(const std::string& name) const { auto* w = W(name); return w ? w->ptr : nullptr; }
};

namespace {

void need(conan_engine* e, const std::string& name, size_t numel, int dtype = CONAN_DTYPE_F32) {
  e->windex[name] = (int)e->weights.size();
  e->weights.push_back(WeightSlot{name, numel, dtype, nullptr});
}

inline int pad32(int x) { return (x + 31) / 32 * 32; }

// A contraction that runs on the tensor cores when lin_tc is on: fp16 [Npad, 3*k*Kpad] = [W_hi | W_lo | W_hi]
// (pre-scaled by 2^10) + fp32 bias [Npad]; otherwise fp32 [N, k*K] + bias [N].
void need_linear(conan_engine* e, const std::string& name, int N, int K, int k = 1, int force_tc = -1) {
  if (force_tc < 0 ? e->lin_tc : force_tc != 0) {
    need(e, name + ".w", (size_t)pad32(N) * 3 * k * pad32(K), CONAN_DTYPE_F16);
    need(e, name + ".b", pad32(N));
  } else {
    need(e, name + ".w", (size_t)N * k * K);
    need(e, name + ".b", N);
  }
}

void declare_weights(conan_engine* e) {
  const conan_config_t& c = e->cfg;
  const int D = c.emformer_dim, F = c.emformer_ffn, H = c.hidden_size;
  for (int l = 0; l < c.emformer_layers; ++l) {
    std::string p = "emf." + std::to_string(l) + ".";
    need(e, p + "ln_in.g", D); need(e, p + "ln_in.b", D);
    need_linear(e, p + "qkv", 3 * D, D);
    need_linear(e, p + "out", D, D);
    if (e->lin_tc && c.lin_fuse_ffn) { need(e, p + "out.wt", (size_t)D * D); need(e, p + "out.bf", D); }   // fp32 out_proj^T for the fused attention epilogue
    need(e, p + "ffn_ln.g", D); need(e, p + "ffn_ln.b", D);
    need_linear(e, p + "ffn1", F, D);
    need_linear(e, p + "ffn2", D, F);
    need(e, p + "ln_out.g", D); need(e, p + "ln_out.b", D);
  }
  need_linear(e, "emf.proj", c.emformer_output_dim, D);

  need(e, "conan.content_embedding", (size_t)102 * H);
  need_linear(e, "conan.content_proj", H, H, c.content_kernel);
  for (int l = 0; l < 2; ++l) {
    std::string p = "conan.align." + std::to_string(l) + ".";
    need_linear(e, p + "q", H, H);
    need(e, p + "kv.w", (size_t)2 * H * H); need(e, p + "kv.b", 2 * H);          // session setup: fp32
    need_linear(e, p + "out", H, H);
    need(e, p + "norm1.g", H); need(e, p + "norm1.b", H);
    need_linear(e, p + "ffn1", 2048, H);
    need_linear(e, p + "ffn2", H, 2048);
    need(e, p + "norm2.g", H); need(e, p + "norm2.b", H);
  }
  for (int i = 0; i < 5; ++i) need_linear(e, "conan.uv." + std::to_string(i), 128, i == 0 ? H : 128, c.predictor_kernel);
  need(e, "conan.uv.ln.g", 128); need(e, "conan.uv.ln.b", 128);
  need(e, "conan.uv.lin.w", 256); need(e, "conan.uv.lin.b", 2);
  need(e, "conan.pitch_embed", (size_t)300 * H);
  for (int b = 0; b < c.dec_blocks; ++b)
    for (int s = 0; s < 2; ++s) {
      std::string p = "conan.dec." + std::to_string(b) + "." + std::to_string(s) + ".";
      need(e, p + "ln.g", H); need(e, p + "ln.b", H);
      need_linear(e, p + "conv", 2 * H, H, c.dec_kernel);
      need_linear(e, p + "pw", H, 2 * H);
    }
  need(e, "conan.dec.last_norm.g", H); need(e, "conan.dec.last_norm.b", H);
  need_linear(e, "conan.dec.post", H, H, c.dec_post_kernel);
  need_linear(e, "conan.mel_out", c.n_mels, H);
  // session-setup branch
  need(e, "conan.global_in.w", (size_t)H * c.n_mels); need(e, "conan.global_in.b", H);
  for (int b = 0; b < 5; ++b)
    for (int s = 0; s < 2; ++s) {
      std::string p = "conan.genc." + std::to_string(b) + "." + std::to_string(s) + ".";
      need(e, p + "ln.g", H); need(e, p + "ln.b", H);
      need_linear(e, p + "conv", 2 * H, H, 31, e->ses_tc);
      need_linear(e, p + "pw", H, 2 * H, 1, e->ses_tc);
      std::string q = "conan.penc." + std::to_string(b) + "." + std::to_string(s) + ".";
      need(e, q + "ln.g", 80); need(e, q + "ln.b", 80);
      need(e, q + "conv.w", (size_t)160 * 5 * 80); need(e, q + "conv.b", 160);
      need(e, q + "pw.w", (size_t)80 * 160); need(e, q + "pw.b", 80);
    }
  need(e, "conan.genc.last_norm.g", H); need(e, "conan.genc.last_norm.b", H);
  need_linear(e, "conan.genc.post", H, H, 3, e->ses_tc);
  need(e, "conan.penc.last_norm.g", 80); need(e, "conan.penc.last_norm.b", 80);
  need(e, "conan.penc.post.w", (size_t)H * 3 * 80); need(e, "conan.penc.post.b", H);
  for (int i = 0; i < 4; ++i) {
    std::string p = "conan.wn." + std::to_string(i) + ".";
    int co = i < 3 ? 160 : 80;
    need(e, p + "in.w", (size_t)160 * 3 * 80); need(e, p + "in.b", 160);
    need(e, p + "rs.w", (size_t)co * 80); need(e, p + "rs.b", co);
  }
  need(e, "conan.vq.embedding", (size_t)c.n_vq * H); need(e, "conan.vq.e2", c.n_vq);
  need(e, "conan.pos_table", (size_t)(e->tp_max + 1) * H);
  need(e, "conan.l1.w", (size_t)H * 2 * H); need(e, "conan.l1.b", H);
  // vocoder
  const int wdt = c.voc_precision ? CONAN_DTYPE_F16 : CONAN_DTYPE_F32;
  const size_t vsp = c.voc_precision == 2 ? 3 : 1;      // split fp16: [W_hi | W_lo | W_hi] of 2^10 W (fp32-grade on tensor cores)
  int ch = c.voc_initial_channel;
  need(e, "voc.pre.w", vsp * ch * 7 * (c.voc_use_tensor_cores ? pad32(c.n_mels) : c.n_mels), wdt); need(e, "voc.pre.b", ch);
  for (int i = 0; i < c.voc_n_ups; ++i) {
    int co = ch / 2;
    std::string p = "voc.up." + std::to_string(i) + ".";
    need(e, p + "w", vsp * co * c.voc_rates[i] * c.voc_up_kernels[i] * ch, wdt); need(e, p + "b", (size_t)co * c.voc_rates[i]);
    for (int r = 0; r < c.voc_n_res; ++r)
      for (int j = 0; j < c.voc_n_dil; ++j) {
        std::string q = "voc.res." + std::to_string(i) + "." + std::to_string(r) + ".";
        need(e, q + "c1." + std::to_string(j) + ".w", vsp * co * c.voc_res_kernels[r] * co, wdt);
        need(e, q + "c1." + std::to_string(j) + ".b", co);
        need(e, q + "c2." + std::to_string(j) + ".w", vsp * co * c.voc_res_kernels[r] * co, wdt);
        need(e, q + "c2." + std::to_string(j) + ".b", co);
      }
    ch = co;
  }
  need(e, "voc.post.w", (size_t)7 * ch); need(e, "voc.post.b", 1);
}

template <typename T>
int dalloc(conan_engine* e, T** out, size_t count) {
  void* p = nullptr;
  size_t bytes = count * sizeof(T);
  if (bytes == 0) bytes = 16;
  bytes = (bytes + 255) & ~(size_t)255;
  CONAN_CUDA_OK(cudaMalloc(&p, bytes));
  CONAN_CUDA_OK(cudaMemset(p, 0, bytes));
  e->allocs.push_back(p);
  e->state_bytes += bytes;
  *out = (T*)p;
  return 0;
}

// is_half: 0 fp32, 1 fp16, 2 split fp16 pair
int alloc_ctx(conan_engine* e, Ctx* c, int H, int L, int R, int C, int is_half) {
  c->H = H; c->L = L; c->R = R; c->C = C; c->is_half = is_half;
  const int planes = is_half == 2 ? 2 : 1;
  c->plane = (long long)e->S * c->rows() * C;            // compact work buffer (up to max_slots streams per step)
  c->hist_plane = (long long)e->S * H * C;
  size_t count = (size_t)planes * c->plane, hcount = (size_t)planes * c->hist_plane;
  if (is_half) {
    __half* p; if (dalloc(e, &p, count)) return 1; c->p = p;
    if (hcount) { __half* h; if (dalloc(e, &h, hcount)) return 1; c->hist = h; }
  } else {
    float* p; if (dalloc(e, &p, count)) return 1; c->p = p;
    if (hcount) { float* h; if (dalloc(e, &h, hcount)) return 1; c->hist = h; }
  }
  return 0;
}

constexpr int kFusedWeightCopies = 16;
inline int fused_weight_copies() {        // CONAN_FUSED_WCOPIES=1..16 (A/B of the L2 hot-spot relief)
  static int n = [] { const char* v = getenv("CONAN_FUSED_WCOPIES"); int x = v ? atoi(v) : 8; return x < 1 ? 1 : (x > kFusedWeightCopies ? kFusedWeightCopies : x); }();
  return n;
}
// Ready lists of the host entry points are padded to a multiple of kStepBucket with dedicated pad slots (distinct ids past
// max_slots, inputs = whatever the staging buffer holds, outputs dropped), so a serving loop whose ready count changes every
// step replays a handful of captured graphs instead of capturing one per distinct count.
constexpr int kStepBucket = 8, kPadSlots = kStepBucket - 1;
// Ready counts up to kGraphSmallN are padded to a bucket and captured the second time a bucket is met (at most 32 buckets: the
// graph cache holds them all).  Larger ready counts are not padded and are captured only once the very same count keeps coming
// back (lock-step serving); a large count that changes from step to step is launched eagerly -- at that size the launches are
// hidden behind the kernels anyway, and capturing a graph per distinct count would thrash the cache.
constexpr int kGraphSmallN = 256, kGraphBigSightings = 3;
constexpr float kSplitWeightScale = 1024.f;      // split weights are packed as 2^10 * W (keeps W_lo out of the fp16 subnormals)

// ---- conv parameter builders (all compact: stream i of the ready list, no slot indirection) -----
conan_conv_params_t conv_on_ctx(const conan_engine* e, const Ctx& in, int k, int dil, const void* w, const float* bias,
                                int cout, int n, bool causal = true, int row0_extra = 0, int L = -1) {
  conan_conv_params_t p;
  memset(&p, 0, sizeof(p));
  p.x = in.p; p.x_slot_stride = in.slot_stride(); p.x_row_stride = in.C; p.x_rows = in.rows(); p.x_is_half = in.is_half ? 1 : 0;
  p.row0 = (causal ? in.H - (k - 1) * dil : in.H - ((k - 1) * dil) / 2) + row0_extra;
  p.L = L < 0 ? in.L : L; p.cin = in.C; p.k = k; p.dil = dil; p.cout = cout; p.w = w; p.bias = bias;
  p.n_streams = n; p.slot_ids = nullptr; p.n_slots = e->S;
  p.scale = 1.f; p.out_scale = 1.f;
  if (in.is_half == 2) { p.x_split = 1; p.x_lo_slot_off = e->S; p.acc_scale = 1.f / kSplitWeightScale; }
  return p;
}
conan_conv_params_t conv_on_rows(const conan_engine* e, const float* x, int rows_per_slot, int row0, int L, int C,
                                 const void* w, const float* bias, int cout, int n) {
  conan_conv_params_t p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.x_slot_stride = (long long)rows_per_slot * C; p.x_row_stride = C; p.x_rows = rows_per_slot; p.x_is_half = 0;
  p.row0 = row0; p.L = L; p.cin = C; p.k = 1; p.dil = 1; p.cout = cout; p.w = w; p.bias = bias;
  p.n_streams = n; p.slot_ids = nullptr; p.n_slots = e->S;
  p.scale = 1.f; p.out_scale = 1.f;
  return p;
}
void out_rows(conan_conv_params_t& p, float* y, int L, int C) { p.y = y; p.y_slot_stride = (long long)L * C; p.y_row_stride = C; p.y_row0 = 0; }
void out_ctx(conan_conv_params_t& p, const Ctx& c) {     // fp32 context buffer as the primary output
  p.y = (float*)c.p; p.y_slot_stride = c.slot_stride(); p.y_row_stride = c.C; p.y_row0 = c.H;
}
void out2_ctx(conan_conv_params_t& p, const Ctx& c, int act2, float slope2) {
  p.y2 = c.p; p.y2_slot_stride = c.slot_stride(); p.y2_row_stride = c.C; p.y2_row0 = c.H; p.y2_is_half = c.is_half ? 1 : 0;
  p.act2 = act2; p.slope2 = slope2;
  if (c.is_half == 2) { p.y2_split = 1; p.y2_lo_off = c.plane; }
}
void res_rows(conan_conv_params_t& p, const float* r, int L, int C) { p.res = r; p.res_slot_stride = (long long)L * C; p.res_row_stride = C; }

#define TRY_RC(x) do { if ((x) != 0) return 1; } while (0)

int run_conv(const conan_engine* e, const conan_conv_params_t& p, cudaStream_t st, bool allow_tc = false) {
  const bool tc = allow_tc && conv_gemm_tc_eligible(p);
  if (p.x_split && !tc) { set_error("internal: split-fp16 operand on a shape the tcgen05 engine cannot run"); return 1; }
  if (!e->profiling) return tc ? launch_conv_gemm_tc(p, st) : launch_conv_gemm_ffma(p, st);
  conan_engine::ProfRec r;
  // categories: 0 FFMA, 1 tcgen05 ring kernel (vocoder), 2 tcgen05 window kernel, 3 tcgen05 ring kernel with split operands
  r.cat = !tc ? 0 : (p.x_split ? 3 : (conv_gemm_tc_uses_window(p) ? 2 : 1));
  r.flops = 2.0 * (double)p.n_streams * p.L * p.cout * p.k * p.cin;
  {
    // algorithmic HBM bytes of the launch: input rows once, weights once, residual / old output read, outputs written
    const double esz = p.x_is_half ? 2.0 : 4.0, rows_out = (double)p.n_streams * p.L;
    double b = (double)p.n_streams * (p.L + (p.k - 1) * p.dil) * p.cin * esz * (p.x_split ? 2 : 1);
    b += (double)p.cout * p.k * p.cin * esz * (p.x_split ? 3 : 1);
    if (p.res) b += rows_out * p.cout * (p.res_is_half ? 2 : 4) * (p.res_row_stride ? 1.0 : 1.0 / p.L);
    if (p.res2) b += rows_out * p.cout * (p.res2_is_half ? 2 : 4);
    if (p.y) b += rows_out * p.cout * (p.y_is_half ? 2 : 4) * (p.accumulate ? 2 : 1);
    if (p.y2) b += rows_out * p.cout * (p.y2_is_half ? 2.0 : 4.0) * (p.y2_split ? 2 : 1);
    r.bytes = b;
  }
  CONAN_CUDA_OK(cudaEventCreate(&r.a)); CONAN_CUDA_OK(cudaEventCreate(&r.b));
  CONAN_CUDA_OK(cudaEventRecord(r.a, st));
  int rc = tc ? launch_conv_gemm_tc(p, st) : launch_conv_gemm_ffma(p, st);
  CONAN_CUDA_OK(cudaEventRecord(r.b, st));
  e->prof.push_back(r);
  return rc;
}

// G independent convs of one shape: one grouped launch of the CTA-pair kernel where they qualify, else one launch each
// sum: the problems are terms of one output (launch_conv_gemm_tc_group's sum mode); returns -1 if the group does not qualify
int run_conv_group(const conan_engine* e, const conan_conv_params_t* ps, int G, cudaStream_t st, bool sum = false) {
  conan_engine::ProfRec r;
  if (e->profiling) {
    r.cat = 1; r.flops = 0; r.bytes = 0;
    for (int i = 0; i < G; ++i) {
      const conan_conv_params_t& p = ps[i];
      const double rows_out = (double)p.n_streams * p.L;
      r.flops += 2.0 * rows_out * p.cout * p.k * p.cin;
      r.bytes += (double)p.n_streams * (p.L + (p.k - 1) * p.dil) * p.cin * 2.0 + (double)p.cout * p.k * p.cin * 2.0 + rows_out * p.cout * 2.0 * 2.0;
    }
    CONAN_CUDA_OK(cudaEventCreate(&r.a)); CONAN_CUDA_OK(cudaEventCreate(&r.b));
    CONAN_CUDA_OK(cudaEventRecord(r.a, st));
  }
  const int rc = launch_conv_gemm_tc_group(ps, G, st, sum);
  if (rc < 0) {                                   // not a group for the pair kernel: launch them one by one (profiled individually)
    if (e->profiling) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    if (sum) return -1;
    for (int i = 0; i < G; ++i) TRY_RC(run_conv(e, ps[i], st, true));
    return 0;
  }
  if (e->profiling) { CONAN_CUDA_OK(cudaEventRecord(r.b, st)); e->prof.push_back(r); }
  return rc;
}

// profiling category 4: fused residual block (six convs per launch)
// CONAN_FUSED_LANES: bit 0 -> two-lane kernel at C = 64, bit 1 -> at C = 32 (default 3: both; 0 = one stream per CTA)
bool fused_two_lanes(int C, int k, const int* dil) {
  static const int lanes = [] { const char* v = getenv("CONAN_FUSED_LANES"); return v ? atoi(v) : 3; }();
  return ((C == 64 && (lanes & 1)) || (C == 32 && (lanes & 2))) && resblock_fused2_smem(C, k, dil) != 0;
}
int fused_launch(const ResblockFusedParams& f, cudaStream_t st) {
  return fused_two_lanes(f.C, f.k, f.dil) ? launch_resblock_fused2(f, st) : launch_resblock_fused(f, st);
}

int run_fused(const conan_engine* e, const ResblockFusedParams& f, cudaStream_t st) {
  if (!e->profiling) return fused_launch(f, st);
  conan_engine::ProfRec r;
  r.cat = 4;
  r.flops = 2.0 * 6.0 * (double)f.n_streams * f.L * f.C * f.k * f.C;
  {
    // algorithmic HBM bytes: input rows once, weights once, running sum in / out, next-layer rows, history in + out
    const double rows = (double)f.n_streams * f.L;
    double b = (double)f.n_streams * (f.L + (f.k - 1) * f.dil[0]) * f.C * 2.0 + 6.0 * f.C * f.k * f.C * 2.0;
    if (f.sum_in) b += rows * f.C * 2.0;
    if (f.sum_out) b += rows * f.C * 2.0;
    if (f.next) b += rows * f.C * 2.0;
    b += 2.0 * (double)f.n_streams * f.hist_slot_stride * 2.0;
    r.bytes = b;
  }
  CONAN_CUDA_OK(cudaEventCreate(&r.a)); CONAN_CUDA_OK(cudaEventCreate(&r.b));
  CONAN_CUDA_OK(cudaEventRecord(r.a, st));
  int rc = fused_launch(f, st);
  CONAN_CUDA_OK(cudaEventRecord(r.b, st));
  e->prof.push_back(r);
  return rc;
}

// profiling category 5: fused position-wise FFN
int run_ffn(const conan_engine* e, const FfnFusedParams& f, cudaStream_t st) {
  if (!e->profiling) return launch_ffn_fused(f, st);
  conan_engine::ProfRec r;
  r.cat = 5;
  r.flops = 2.0 * (double)f.M * f.hidden * (f.K + f.N);
  r.bytes = (double)f.M * f.K * 4.0 + (double)f.hidden * (f.K + f.N) * 6.0 + (double)f.FS * f.M * f.N * 4.0;
  CONAN_CUDA_OK(cudaEventCreate(&r.a)); CONAN_CUDA_OK(cudaEventCreate(&r.b));
  CONAN_CUDA_OK(cudaEventRecord(r.a, st));
  int rc = launch_ffn_fused(f, st);
  CONAN_CUDA_OK(cudaEventRecord(r.b, st));
  e->prof.push_back(r);
  return rc;
}

int ln_rows(const float* in, int in_rows, int in_ld, int in_row0, RowView out, const float* g, const float* b, int C, int L, int n,
            cudaStream_t st, const float* premask = nullptr, const float* postmask = nullptr, int mask_stride = 0,
            float* write_mask = nullptr, float* write_mask2 = nullptr, RowView out2 = RowView{}) {
  LnArgs a;
  a.in = RowView{(void*)in, (long long)in_rows * in_ld, in_ld, in_row0, 0, 0};
  a.out = out; a.out2 = out2; a.gamma = g; a.beta = b; a.eps = 1e-5f; a.C = C; a.L = L; a.n = n; a.slot_ids = nullptr;
  a.premask = premask; a.premask_slot_stride = mask_stride;
  a.postmask = postmask; a.postmask_slot_stride = mask_stride;
  a.write_mask = write_mask; a.write_mask_slot_stride = mask_stride; a.write_mask2 = write_mask2;
  return launch_layernorm(a, st);
}

#define TRY(x) do { if ((x) != 0) return 1; } while (0)

// profiling category 6: fused two-GEMM Conan block (decoder residual block body / aligner feed-forward)
int run_block(const conan_engine* e, const BlockFusedParams& f, cudaStream_t st) {
  if (!e->profiling) return launch_block_fused(f, st);
  conan_engine::ProfRec r;
  r.cat = 6;
  const double rows = (double)f.n_streams * f.L;
  r.flops = 2.0 * rows * f.hidden * ((double)f.k * f.C1 + f.N2);
  r.bytes = rows * f.C1 * 4.0 + (double)f.hidden * (f.k * f.C1 + f.N2) * 4.0 + (double)f.FS * rows * f.N2 * 4.0;
  CONAN_CUDA_OK(cudaEventCreate(&r.a)); CONAN_CUDA_OK(cudaEventCreate(&r.b));
  CONAN_CUDA_OK(cudaEventRecord(r.a, st));
  int rc = launch_block_fused(f, st);
  CONAN_CUDA_OK(cudaEventRecord(r.b, st));
  e->prof.push_back(r);
  return rc;
}

BlockFusedParams block_on_ctx(const conan_engine* e, const Ctx& in, int k, const void* w1, const float* b1, int hidden, float scale1, int act,
                              const void* w2, int n) {
  BlockFusedParams f;
  memset(&f, 0, sizeof(f));
  f.x = in.p; f.x_slot_stride = in.slot_stride(); f.x_rows = in.rows(); f.n_slots = e->S; f.lo_slot_off = e->S;
  f.C1 = in.C; f.k = k; f.row0 = in.H - (k - 1); f.L = in.L; f.n_streams = n;
  f.w1 = w1; f.b1 = b1; f.hidden = hidden; f.scale1 = scale1; f.act = act; f.w2 = w2; f.N2 = 256;
  f.partials = e->bfP; f.FS = block_fused_split(n, in.L, hidden, e->bfRows); f.acc_scale = 1.f / kSplitWeightScale;
  return f;
}

// ============================================================================ allocation
int allocate_state(conan_engine* e) {
  const conan_config_t& c = e->cfg;
  const int S = e->S, D = c.emformer_dim, H = c.hidden_size, seg = c.segment, rows = c.segment + c.right_context;
  // ---- Emformer (scratch compact, K|V ring + past_len resident).  GEMM operands are fp32 rows (FFMA mode) or
  // split-fp16 pairs with the model dim padded to a multiple of 32 (tensor-core mode); pad columns stay zero.
  const int lt = e->lin_tc ? 2 : 0;
  e->DP = e->lin_tc ? pad32(D) : D; e->QP = e->lin_tc ? pad32(3 * D) : 3 * D;
  e->LP = e->lin_tc ? pad32(c.emformer_output_dim) : c.emformer_output_dim;
  const int DP = e->DP;
  e->ring_rows = ((c.left_context + seg + seg - 1) / seg) * seg;      // >= lc + seg, multiple of seg
  TRY(dalloc(e, &e->eX, (size_t)S * rows * DP)); TRY(dalloc(e, &e->eQKV, (size_t)S * rows * e->QP));
  TRY(dalloc(e, &e->eR1, (size_t)S * rows * DP)); TRY(dalloc(e, &e->eR2, (size_t)S * rows * DP));
  TRY(dalloc(e, &e->eLOG, (size_t)S * seg * e->LP));
  TRY(alloc_ctx(e, &e->eXN, 0, rows, 0, DP, lt)); TRY(alloc_ctx(e, &e->eATT, 0, rows, 0, DP, lt));
  TRY(alloc_ctx(e, &e->eFN, 0, rows, 0, DP, lt)); TRY(alloc_ctx(e, &e->eHF, 0, rows, 0, c.emformer_ffn, lt));
  e->ffnFused = e->lin_tc && c.lin_fuse_ffn && ffn_fused_eligible(DP, c.emformer_ffn, DP);   // M = 0 fast path only; M > 0 steps take the generic path
  if (e->ffnFused) TRY(dalloc(e, &e->eFFP, (size_t)4 * S * rows * DP));
  e->eRing.resize(c.emformer_layers);
  for (int l = 0; l < c.emformer_layers; ++l) TRY(dalloc(e, &e->eRing[l], (size_t)S * e->ring_rows * 2 * D));
  {
    // generic-step buffers (allocated for every engine: a full-utterance forward with a partial last segment uses them at M = 0 too)
    const int er = rows + 1, M = e->eM;
    TRY(dalloc(e, &e->gX, (size_t)S * er * DP)); TRY(dalloc(e, &e->gQKV, (size_t)S * er * e->QP));
    TRY(dalloc(e, &e->gR1, (size_t)S * er * DP)); TRY(dalloc(e, &e->gR2, (size_t)S * er * DP));
    TRY(dalloc(e, &e->gXNf, (size_t)S * er * DP));
    TRY(alloc_ctx(e, &e->gXN, 0, er, 0, DP, lt)); TRY(alloc_ctx(e, &e->gATT, 0, er, 0, DP, lt));
    TRY(alloc_ctx(e, &e->gFN, 0, er, 0, DP, lt)); TRY(alloc_ctx(e, &e->gHF, 0, er, 0, c.emformer_ffn, lt));
    TRY(dalloc(e, &e->gMEM[0], (size_t)S * D)); TRY(dalloc(e, &e->gMEM[1], (size_t)S * D));
    if (M > 0) {
      TRY(alloc_ctx(e, &e->gMB, 0, M, 0, DP, lt)); TRY(dalloc(e, &e->gMKV, (size_t)S * M * e->QP));
      e->eBank.resize(c.emformer_layers);
      for (int l = 0; l < c.emformer_layers; ++l) TRY(dalloc(e, &e->eBank[l], (size_t)S * M * D));
    }
  }
  TRY(dalloc(e, &e->ePast, (size_t)S)); TRY(dalloc(e, &e->TOK, (size_t)S * seg));
  // ---- Conan chunk path
  TRY(alloc_ctx(e, &e->cC, c.content_kernel - 1, seg, 0, H, lt));
  for (int i = 0; i < 5; ++i) TRY(alloc_ctx(e, &e->cUV[i], c.predictor_kernel - 1, seg, 0, i == 0 ? H : 128, lt));
  for (int b = 0; b < c.dec_blocks; ++b)
    for (int s = 0; s < 2; ++s) TRY(alloc_ctx(e, &e->cD[b][s], c.dec_kernel - 1, seg, 0, H, lt));
  TRY(alloc_ctx(e, &e->cP, c.dec_post_kernel - 1, seg, 0, H, lt));
  TRY(alloc_ctx(e, &e->cX0, 0, seg, 0, H, lt)); TRY(alloc_ctx(e, &e->cATT, 0, seg, 0, H, lt));
  TRY(alloc_ctx(e, &e->cO1, 0, seg, 0, H, lt)); TRY(alloc_ctx(e, &e->cHF, 0, seg, 0, 2048, lt));
  TRY(alloc_ctx(e, &e->cPROS[0], 0, seg, 0, H, lt)); TRY(alloc_ctx(e, &e->cPROS[1], 0, seg, 0, H, lt));
  TRY(alloc_ctx(e, &e->cDECH, 0, seg, 0, 2 * H, lt));
  static const int fuse_blocks_env = [] { const char* v = getenv("CONAN_FUSE_BLOCKS"); return v ? atoi(v) : -1; }();
  e->blockFused = e->lin_tc && (fuse_blocks_env < 0 ? c.lin_fuse_blocks != 0 : fuse_blocks_env != 0) && H == 256 && block_fused_eligible(H, c.dec_kernel, 2 * H, H, seg) &&
                  block_fused_eligible(H, 1, 2048, H, seg);
  if (e->blockFused) { e->bfRows = (long long)4 * S * seg; TRY(dalloc(e, &e->bfP, (size_t)e->bfRows * H)); }
  { const char* v = getenv("CONAN_BLOCK_CLUSTER"); if (v) e->blockCluster = atoi(v) != 0; }
  TRY(dalloc(e, &e->dX0, (size_t)S * seg * H)); TRY(dalloc(e, &e->dQ, (size_t)S * seg * H));
  TRY(dalloc(e, &e->dT1, (size_t)S * seg * H)); TRY(dalloc(e, &e->dO1, (size_t)S * seg * H));
  TRY(dalloc(e, &e->dT2, (size_t)S * seg * H));
  TRY(dalloc(e, &e->dPROS[0], (size_t)S * seg * H)); TRY(dalloc(e, &e->dPROS[1], (size_t)S * seg * H));
  TRY(dalloc(e, &e->dPINP, (size_t)S * seg * H)); TRY(dalloc(e, &e->dUVH, (size_t)S * seg * 128));
  TRY(dalloc(e, &e->dDECX, (size_t)S * seg * H));
  TRY(dalloc(e, &e->dPOST, (size_t)S * seg * H)); TRY(dalloc(e, &e->dMEL, (size_t)S * seg * c.n_mels));
  e->MP = pad32(c.n_mels);
  if (e->lin_tc) { TRY(alloc_ctx(e, &e->cPOSTO, 0, seg, 0, H, lt)); TRY(dalloc(e, &e->dMELP, (size_t)S * seg * e->MP)); }
  TRY(dalloc(e, &e->dUVP, (size_t)S * seg * 4)); TRY(dalloc(e, &e->dMASK0, (size_t)S * seg));
  TRY(dalloc(e, &e->dMASKB, (size_t)S * seg));
  TRY(dalloc(e, &e->sSTYLE, (size_t)S * H)); TRY(dalloc(e, &e->sSTYLEW, (size_t)S * H));
  TRY(dalloc(e, &e->sKV, (size_t)S * 2 * e->tp_max * 2 * H));
  TRY(dalloc(e, &e->sKPM, (size_t)S * e->tp_max)); TRY(dalloc(e, &e->sNKEYS, (size_t)S));
  // ---- vocoder
  const int hf = c.voc_precision;       // context element type: 0 fp32, 1 fp16, 2 split fp16 pair
  e->vL[0] = seg; e->vC[0] = c.voc_initial_channel;
  for (int i = 0; i < c.voc_n_ups; ++i) { e->vL[i + 1] = e->vL[i] * c.voc_rates[i]; e->vC[i + 1] = e->vC[i] / 2; }
  // tensor-core mode: mel rows padded to a multiple of 32 channels (pad columns stay zero) so conv_pre is tcgen05-eligible
  TRY(alloc_ctx(e, &e->vPRE, 6, seg, 0, c.voc_use_tensor_cores ? pad32(c.n_mels) : c.n_mels, hf));
  size_t maxLC = 0;
  std::vector<ZeroDesc> fused_zero;
  std::vector<HistDesc> fused_hist;
  for (int i = 0; i < c.voc_n_ups; ++i) {
    TRY(alloc_ctx(e, &e->vUP[i], c.voc_up_kernels[i] - 1, e->vL[i], 0, e->vC[i], hf));
    int L = e->vL[i + 1], C = e->vC[i + 1];
    int hmax = 0;
    for (int r = 0; r < c.voc_n_res; ++r) hmax = std::max(hmax, (c.voc_res_kernels[r] - 1) * c.voc_res_dilations[0]);
    TRY(alloc_ctx(e, &e->vXA[i], hmax, L, 0, C, hf));
    e->vFused[i] = c.voc_fuse_resblocks && c.voc_use_tensor_cores && c.voc_residual_from_ctx && hf == 1 && c.voc_n_dil == 3;
    for (int r = 0; r < c.voc_n_res && e->vFused[i]; ++r)
      e->vFused[i] = resblock_fused_eligible(C, L, c.voc_res_kernels[r], c.voc_res_dilations) && c.voc_res_dilations[0] == 1;
    maxLC = std::max(maxLC, (size_t)L * C);
    if (e->vFused[i]) {
      // no per-conv context buffers at this scale: the block's intermediate activations live in shared memory; only
      // the newest halo rows of each conv window are resident per slot
      for (int r = 0; r < c.voc_n_res; ++r) {
        const int k = c.voc_res_kernels[r];
        e->fHistRows[i][r] = resblock_fused_hist_rows(k, c.voc_res_dilations);
        TRY(dalloc(e, &e->fHist[i][r], (size_t)S * e->fHistRows[i][r] * C));
        fused_zero.push_back(ZeroDesc{e->fHist[i][r], (long long)e->fHistRows[i][r] * C * 2, (long long)e->fHistRows[i][r] * C * 2});
        if (fused_two_lanes(C, k, c.voc_res_dilations)) {
          // the kernel leaves the new history in a compact staging block; the vocoder's history scatter moves it to the slot
          // (so lanes may cut a stream between tiles: nobody overwrites history another lane has yet to read)
          const int hb = e->fHistRows[i][r] * C * 2;
          TRY(dalloc(e, &e->fHistOut[i][r], (size_t)S * e->fHistRows[i][r] * C));
          fused_hist.push_back(HistDesc{e->fHistOut[i][r], (long long)hb, e->fHist[i][r], hb, 0, 2});
        }
        const size_t wn = (size_t)C * k * C;
        TRY(dalloc(e, &e->fW[i][r], (size_t)kFusedWeightCopies * 6 * wn)); TRY(dalloc(e, &e->fB[i][r], (size_t)6 * C));
        for (int j = 0; j < 3; ++j)
          for (int h = 0; h < 2; ++h) {
            std::string q = "voc.res." + std::to_string(i) + "." + std::to_string(r) + (h ? ".c2." : ".c1.") + std::to_string(j);
            for (int cp = 0; cp < kFusedWeightCopies; ++cp)
              CONAN_CUDA_OK(cudaMemcpy(e->fW[i][r] + ((size_t)cp * 6 + 2 * j + h) * wn, e->P(q + ".w"), wn * 2, cudaMemcpyDeviceToDevice));
            CONAN_CUDA_OK(cudaMemcpy(e->fB[i][r] + (size_t)(2 * j + h) * C, e->F(q + ".b"), (size_t)C * 4, cudaMemcpyDeviceToDevice));
          }
      }
      continue;
    }
    for (int r = 0; r < c.voc_n_res; ++r)
      for (int j = 0; j < c.voc_n_dil; ++j) {
        if (j > 0) TRY(alloc_ctx(e, &e->vC1[i][r][j], (c.voc_res_kernels[r] - 1) * c.voc_res_dilations[j], L, 0, C, hf));
        TRY(alloc_ctx(e, &e->vC2[i][r][j], c.voc_res_kernels[r] - 1, L, 0, C, hf));
      }
    maxLC = std::max(maxLC, (size_t)L * C);
  }
  TRY(alloc_ctx(e, &e->vPOST, 6, e->vL[c.voc_n_ups], 0, e->vC[c.voc_n_ups], hf));
  {
    const char* v = getenv("CONAN_MEL_DIRECT");
    e->melDirect = e->lin_tc && e->vPRE.is_half != 0 && e->vPRE.C == e->MP && c.voc_group <= 0 && !(v && atoi(v) == 0);
  }
  TRY(dalloc(e, &e->vXS, (size_t)S * maxLC)); TRY(dalloc(e, &e->vXR[0], (size_t)S * maxLC));
  TRY(dalloc(e, &e->vXR[1], (size_t)S * maxLC)); TRY(dalloc(e, &e->vSUM, (size_t)S * maxLC));
  TRY(dalloc(e, &e->vSUMh, (size_t)S * maxLC));

  // ---- history / zero tables
  auto build_tables = [&](std::vector<const Ctx*> ctxs, std::vector<HistDesc> extra, HistDesc** hist, int* nhist, ZeroDesc** zeros,
                          int* nzeros, std::vector<ZeroDesc> extra_zero) -> int {
    std::vector<HistDesc> r = extra; std::vector<ZeroDesc> z = extra_zero;
    for (const Ctx* cx : ctxs) {
      if (cx->H <= 0) continue;
      size_t rowb = (size_t)cx->C * cx->elem();
      for (int pl = 0; pl < (cx->is_half == 2 ? 2 : 1); ++pl) {          // a split buffer is two planes (hi, lo)
        HistDesc d{(char*)cx->p + (size_t)pl * cx->plane * cx->elem(), (long long)(cx->slot_stride() * cx->elem()),
                   (char*)cx->hist + (size_t)pl * cx->hist_plane * cx->elem(), (int)(cx->H * rowb), (int)(cx->L * rowb), 1};
        if (d.hist_bytes % 16 || d.new_bytes % 16 || d.work_stride_bytes % 16) { set_error("context sizes must be multiples of 16 bytes"); return 1; }
        r.push_back(d);
        z.push_back(ZeroDesc{d.hist, (long long)d.hist_bytes, (long long)d.hist_bytes});
      }
    }
    *nhist = (int)r.size(); *nzeros = (int)z.size();
    if (!r.empty()) {
      TRY(dalloc(e, hist, r.size()));
      CONAN_CUDA_OK(cudaMemcpy(*hist, r.data(), r.size() * sizeof(HistDesc), cudaMemcpyHostToDevice));
    }
    if (!z.empty()) {
      TRY(dalloc(e, zeros, z.size()));
      CONAN_CUDA_OK(cudaMemcpy(*zeros, z.data(), z.size() * sizeof(ZeroDesc), cudaMemcpyHostToDevice));
    }
    return 0;
  };
  {
    std::vector<ZeroDesc> ez;
    for (int l = 0; l < c.emformer_layers; ++l)
      ez.push_back(ZeroDesc{e->eRing[l], (long long)e->ring_rows * 2 * D * 4, (long long)e->ring_rows * 2 * D * 4});
    ez.push_back(ZeroDesc{e->ePast, 4, 4});
    for (size_t l = 0; l < e->eBank.size(); ++l) ez.push_back(ZeroDesc{e->eBank[l], (long long)e->eM * D * 4, (long long)e->eM * D * 4});
    HistDesc* dummy = nullptr; int nd = 0;
    TRY(build_tables({}, {}, &dummy, &nd, &e->zeroEmf, &e->nZeroEmf, ez));
  }
  {
    std::vector<const Ctx*> cs{&e->cC, &e->cP};
    for (int i = 0; i < 5; ++i) cs.push_back(&e->cUV[i]);
    for (int b = 0; b < c.dec_blocks; ++b) for (int s = 0; s < 2; ++s) cs.push_back(&e->cD[b][s]);
    // the session's style vector is read-only state: gathered into a compact copy, never scattered back
    std::vector<HistDesc> extra{HistDesc{e->sSTYLEW, (long long)H * 4, e->sSTYLE, H * 4, 0, 0}};
    TRY(build_tables(cs, extra, &e->histConan, &e->nHistConan, &e->zeroConan, &e->nZeroConan, {}));
  }
  {
    std::vector<const Ctx*> cs{&e->vPRE, &e->vPOST};
    for (int i = 0; i < c.voc_n_ups; ++i) {
      cs.push_back(&e->vUP[i]); cs.push_back(&e->vXA[i]);
      if (e->vFused[i]) continue;
      for (int r = 0; r < c.voc_n_res; ++r)
        for (int j = 0; j < c.voc_n_dil; ++j) { if (j > 0) cs.push_back(&e->vC1[i][r][j]); cs.push_back(&e->vC2[i][r][j]); }
    }
    TRY(build_tables(cs, fused_hist, &e->histVoc, &e->nHistVoc, &e->zeroVoc, &e->nZeroVoc, fused_zero));
  }
  // ---- session scratch
  e->SB = std::min(S, 64);
  // T: rows per session in the scratch buffers.  The tensor-core path pads every session to a multiple of 32 rows (tile geometry)
  const int SB = e->SB, T = (c.max_ref_frames + 31) / 32 * 32, Tp = e->tp_max;
  if (e->ses_tc) {
    TRY(dalloc(e, &e->qC31h, (size_t)2 * SB * (T + 30) * H)); TRY(dalloc(e, &e->qHGh, (size_t)2 * SB * T * 2 * H));
    TRY(dalloc(e, &e->qC3Gh, (size_t)2 * SB * (T + 2) * H)); TRY(dalloc(e, &e->qMAp, (size_t)SB * T));
  }
  TRY(dalloc(e, &e->qMA, (size_t)SB * T)); TRY(dalloc(e, &e->qMF, (size_t)SB * T)); TRY(dalloc(e, &e->qMGB, (size_t)SB * T));
  TRY(dalloc(e, &e->qXG, (size_t)SB * T * H)); TRY(dalloc(e, &e->qC31, (size_t)SB * (T + 30) * H));
  TRY(dalloc(e, &e->qHG, (size_t)SB * T * 2 * H)); TRY(dalloc(e, &e->qC3G, (size_t)SB * (T + 2) * H));
  TRY(dalloc(e, &e->qPG, (size_t)SB * T * H));
  TRY(dalloc(e, &e->qXW, (size_t)SB * T * 80)); TRY(dalloc(e, &e->qCW, (size_t)SB * (T + 2) * 80));
  TRY(dalloc(e, &e->qAW, (size_t)SB * T * 160)); TRY(dalloc(e, &e->qACT, (size_t)SB * T * 80));
  TRY(dalloc(e, &e->qRS, (size_t)SB * T * 160)); TRY(dalloc(e, &e->qSKIP, (size_t)SB * T * 80));
  TRY(dalloc(e, &e->qGRP, (size_t)SB * Tp * 80)); TRY(dalloc(e, &e->qMP, (size_t)SB * Tp)); TRY(dalloc(e, &e->qMPB, (size_t)SB * Tp));
  TRY(dalloc(e, &e->qXP, (size_t)SB * Tp * 80)); TRY(dalloc(e, &e->qC5, (size_t)SB * (Tp + 4) * 80));
  TRY(dalloc(e, &e->qHP, (size_t)SB * Tp * 160)); TRY(dalloc(e, &e->qC3P, (size_t)SB * (Tp + 2) * 80));
  TRY(dalloc(e, &e->qPZ, (size_t)SB * Tp * H)); TRY(dalloc(e, &e->qXE, (size_t)SB * Tp * c.n_vq));
  TRY(dalloc(e, &e->qZC, (size_t)SB * Tp * 2 * H)); TRY(dalloc(e, &e->qPE, (size_t)SB * Tp * H));
  TRY(dalloc(e, &e->qKVs, (size_t)SB * Tp * 2 * H)); TRY(dalloc(e, &e->qVQ, (size_t)SB * Tp));
  TRY(dalloc(e, &e->qSlots, (size_t)SB));
  // ---- host-call staging
  TRY(dalloc(e, &e->hIds, (size_t)S)); TRY(dalloc(e, &e->hIdsSmall, (size_t)S));
  {
    int pad[kPadSlots];
    for (int i = 0; i < kPadSlots; ++i) pad[i] = e->Su + i;
    TRY(dalloc(e, &e->padIds, (size_t)kPadSlots));
    CONAN_CUDA_OK(cudaMemcpy(e->padIds, pad, sizeof(pad), cudaMemcpyHostToDevice));
  }
  TRY(dalloc(e, &e->hChunk, (size_t)S * rows * D));
  TRY(dalloc(e, &e->hWav, (size_t)S * e->vL[c.voc_n_ups])); TRY(dalloc(e, &e->hMel, (size_t)S * seg * c.n_mels));
  TRY(dalloc(e, &e->hTok, (size_t)S * seg));
  TRY(dalloc(e, &e->hWav2, (size_t)S * e->vL[c.voc_n_ups])); TRY(dalloc(e, &e->hMel2, (size_t)S * seg * c.n_mels));
  TRY(dalloc(e, &e->hTok2, (size_t)S * seg));
  CONAN_CUDA_OK(cudaStreamCreateWithFlags(&e->copyStream, cudaStreamNonBlocking));
  { const char* v = getenv("CONAN_VOC_STREAMS"); if (v) e->vocStreams = std::max(1, std::min(3, atoi(v))); }
  for (int i = 0; i < 2; ++i) CONAN_CUDA_OK(cudaStreamCreateWithFlags(&e->branchStream[i], cudaStreamNonBlocking));
  CONAN_CUDA_OK(cudaStreamCreateWithFlags(&e->graphStream, cudaStreamNonBlocking));
  e->graphMode = e->cfg.step_graphs;
  { const char* v = getenv("CONAN_STEP_GRAPH"); if (v) e->graphMode = atoi(v) != 0; }
  CONAN_CUDA_OK(cudaEventCreateWithFlags(&e->evFork, cudaEventDisableTiming));
  for (int i = 0; i < 4; ++i) CONAN_CUDA_OK(cudaEventCreateWithFlags(&e->evBranch[i], cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) {
    CONAN_CUDA_OK(cudaEventCreateWithFlags(&e->evCompute[i], cudaEventDisableTiming));
    CONAN_CUDA_OK(cudaEventCreateWithFlags(&e->evCopy[i], cudaEventDisableTiming));
  }
  return 0;
}

int emformer_step_generic(conan_engine* e, int n, const int* ids, const float* src, long long src_stride, int utt_row0, int rc_row0,
                          int n_utt, float* enc_out, float* logits_out, int* tokens_out, int out_row0, int out_nrows, cudaStream_t st);

// ============================================================================ Emformer step
int emformer_step(conan_engine* e, int n, const int* ids, const float* chunk, float* enc_out, float* logits_out,
                  int* tokens_out, cudaStream_t st) {
  const conan_config_t& c = e->cfg;
  const int D = c.emformer_dim, DP = e->DP, QP = e->QP, LP = e->LP, seg = c.segment, rc = c.right_context, rows = seg + rc,
            F = c.emformer_ffn;
  const bool tc = e->lin_tc;
  if (e->eM > 0)        // memory bank: generic step; the chunk is [utt (seg) | look-ahead (rc)] rows per stream
    return emformer_step_generic(e, n, ids, chunk, (long long)rows * D, 0, seg, seg, enc_out, logits_out, tokens_out, 0, seg, st);
  TRY(launch_emformer_assemble(chunk, e->eX, DP, n, nullptr, seg, rc, D, st));
  for (int l = 0; l < c.emformer_layers; ++l) {
    std::string p = "emf." + std::to_string(l) + ".";
    // layers > 0: the input norm was applied by the previous layer's output-norm launch (chained LayerNorm)
    if (l == 0) TRY(ln_rows(e->eX, rows, DP, 0, e->eXN.new_rows(), e->F(p + "ln_in.g"), e->F(p + "ln_in.b"), D, rows, n, st));
    auto q = conv_on_ctx(e, e->eXN, 1, 1, e->P(p + "qkv.w"), e->F(p + "qkv.b"), QP, n);
    out_rows(q, e->eQKV, rows, QP);
    TRY(run_conv(e, q, st, tc));
    if (e->ffnFused) {
      // attention + out_proj (fp32) + residual + the FFN's LayerNorm in one kernel: writes eR1 (fp32) and eFN (split fp16)
      EmfAttnEpilogue ep;
      ep.wt = e->F(p + "out.wt"); ep.bias = e->F(p + "out.bf"); ep.x_res = e->eX; ep.ld = DP; ep.r1 = e->eR1;
      ep.ln_g = e->F(p + "ffn_ln.g"); ep.ln_b = e->F(p + "ffn_ln.b"); ep.eps = 1e-5f; ep.fn = e->eFN.new_rows();
      TRY(launch_emformer_attention(e->eQKV, e->eRing[l], e->ePast, e->eATT.new_rows(), n, ids, seg, rc, c.left_context, e->ring_rows, D,
                                    c.emformer_heads, QP, st, &ep));
    } else {
      TRY(launch_emformer_attention(e->eQKV, e->eRing[l], e->ePast, e->eATT.new_rows(), n, ids, seg, rc, c.left_context, e->ring_rows, D,
                                    c.emformer_heads, QP, st));
      auto o = conv_on_ctx(e, e->eATT, 1, 1, e->P(p + "out.w"), e->F(p + "out.b"), DP, n);
      out_rows(o, e->eR1, rows, DP); res_rows(o, e->eX, rows, DP);
      TRY(run_conv(e, o, st, tc));
      TRY(ln_rows(e->eR1, rows, DP, 0, e->eFN.new_rows(), e->F(p + "ffn_ln.g"), e->F(p + "ffn_ln.b"), D, rows, n, st));
    }
    // the last layer's output is also the operand of the projection GEMM
    const bool last = (l == c.emformer_layers - 1);
    if (e->ffnFused) {
      // one kernel for 80 -> 2048 -> 80: partial outputs per hidden slice; the LayerNorm sums them with b2 and the residual
      FfnFusedParams f;
      memset(&f, 0, sizeof(f));
      f.x = e->eFN.p; f.x_lo_off = e->eFN.plane; f.x_rows = e->S * rows; f.M = n * rows; f.K = DP; f.hidden = F; f.N = DP;
      f.w1 = e->P(p + "ffn1.w"); f.b1 = e->F(p + "ffn1.b"); f.w2 = e->P(p + "ffn2.w");
      f.partials = e->eFFP; f.FS = ffn_fused_split(f.M); f.acc_scale = 1.f / kSplitWeightScale;
      TRY(run_ffn(e, f, st));
      LnArgs a;
      a.in = RowView{(void*)e->eR1, (long long)rows * DP, DP, 0, 0, 0};
      a.out = view_f32(e->eX, (long long)rows * DP, DP); a.out2 = last ? e->eXN.new_rows() : RowView{};
      a.gamma = e->F(p + "ln_out.g"); a.beta = e->F(p + "ln_out.b"); a.eps = 1e-5f; a.C = D; a.L = rows; a.n = n; a.slot_ids = nullptr;
      a.premask = nullptr; a.premask_slot_stride = 0; a.postmask = nullptr; a.postmask_slot_stride = 0;
      a.write_mask = nullptr; a.write_mask_slot_stride = 0; a.write_mask2 = nullptr;
      a.part = e->eFFP; a.n_part = f.FS; a.part_stride = (long long)f.M * DP; a.part_ld = DP;
      a.part_bias = e->F(p + "ffn2.b"); a.part_res = e->eR1; a.part_res_ld = DP;
      if (!last) {
        std::string pn = "emf." + std::to_string(l + 1) + ".";
        a.gamma2 = e->F(pn + "ln_in.g"); a.beta2 = e->F(pn + "ln_in.b"); a.out3 = e->eXN.new_rows();
      }
      TRY(launch_layernorm(a, st));
    } else {
      auto f1 = conv_on_ctx(e, e->eFN, 1, 1, e->P(p + "ffn1.w"), e->F(p + "ffn1.b"), F, n);
      out2_ctx(f1, e->eHF, ACT_NONE, 0.f); f1.act = ACT_RELU;
      TRY(run_conv(e, f1, st, tc));
      auto f2 = conv_on_ctx(e, e->eHF, 1, 1, e->P(p + "ffn2.w"), e->F(p + "ffn2.b"), DP, n);
      out_rows(f2, e->eR2, rows, DP); res_rows(f2, e->eR1, rows, DP);
      TRY(run_conv(e, f2, st, tc));
      LnArgs a;
      a.in = RowView{(void*)e->eR2, (long long)rows * DP, DP, 0, 0, 0};
      a.out = view_f32(e->eX, (long long)rows * DP, DP); a.out2 = last ? e->eXN.new_rows() : RowView{};
      a.gamma = e->F(p + "ln_out.g"); a.beta = e->F(p + "ln_out.b"); a.eps = 1e-5f; a.C = D; a.L = rows; a.n = n; a.slot_ids = nullptr;
      a.premask = nullptr; a.premask_slot_stride = 0; a.postmask = nullptr; a.postmask_slot_stride = 0;
      a.write_mask = nullptr; a.write_mask_slot_stride = 0; a.write_mask2 = nullptr;
      if (!last) {
        std::string pn = "emf." + std::to_string(l + 1) + ".";
        a.gamma2 = e->F(pn + "ln_in.g"); a.beta2 = e->F(pn + "ln_in.b"); a.out3 = e->eXN.new_rows();
      }
      TRY(launch_layernorm(a, st));
    }
  }
  TRY(launch_advance_past_len(e->ePast, n, ids, seg, st));
  auto pj = conv_on_ctx(e, e->eXN, 1, 1, e->P("emf.proj.w"), e->F("emf.proj.b"), LP, n, true, rc, seg);   // utterance rows only
  out_rows(pj, e->eLOG, seg, LP);
  TRY(run_conv(e, pj, st, tc));
  TRY(launch_argmax_rows(e->eLOG, LP, e->TOK, tokens_out, n, seg, c.emformer_output_dim, st));
  if (enc_out) TRY(launch_copy_rows_out(e->eX, (long long)rows * DP, DP, rc, enc_out, n, nullptr, seg, D, st));
  if (logits_out)
    TRY(launch_copy_rows_out(e->eLOG, (long long)seg * LP, LP, 0, logits_out, n, nullptr, seg, c.emformer_output_dim, st));
  return 0;
}

// ============================================================================ Emformer step, generic form
// Memory bank (M > 0), summary query, partial segments (n_utt < seg): see emformer_mem.cu.  `src` holds, for stream i, the
// utterance rows at src[i*src_stride + (utt_row0 + t)*D] and the look-ahead rows at src[i*src_stride + (rc_row0 + q)*D].
// Outputs (any may be null) are written for the n_utt real rows at row offset out_row0 of buffers with out_nrows rows per stream.
int emformer_step_generic(conan_engine* e, int n, const int* ids, const float* src, long long src_stride, int utt_row0, int rc_row0,
                          int n_utt, float* enc_out, float* logits_out, int* tokens_out, int out_row0, int out_nrows, cudaStream_t st) {
  const conan_config_t& c = e->cfg;
  const int D = c.emformer_dim, DP = e->DP, QP = e->QP, LP = e->LP, seg = c.segment, rc = c.right_context, rows = seg + rc, er = rows + 1,
            F = c.emformer_ffn, M = e->eM;
  const bool tc = e->lin_tc;
  TRY(launch_emformer_assemble_generic(src, src_stride, utt_row0, rc_row0, e->gX, DP, M > 0 ? e->gMEM[0] : nullptr, n, seg, n_utt, rc, D, st));
  int cur = 0;
  for (int l = 0; l < c.emformer_layers; ++l) {
    std::string p = "emf." + std::to_string(l) + ".";
    // input norm over the rc + seg real rows (TA:427-434); an fp32 copy feeds the summary mean
    TRY(ln_rows(e->gX, er, DP, 0, e->gXN.new_rows(), e->F(p + "ln_in.g"), e->F(p + "ln_in.b"), D, rows, n, st, nullptr, nullptr, 0, nullptr,
                nullptr, view_f32(e->gXNf, (long long)er * DP, DP)));
    if (M > 0)
      TRY(launch_emformer_mem_prepare(e->gXNf, DP, e->gXN.new_rows(), e->eBank[l], e->gMB.new_rows(), n, ids, seg, n_utt, rc, D, M, st));
    auto q = conv_on_ctx(e, e->gXN, 1, 1, e->P(p + "qkv.w"), e->F(p + "qkv.b"), QP, n);       // Q | K | V of [rc | utt | summary]
    out_rows(q, e->gQKV, er, QP);
    TRY(run_conv(e, q, st, tc));
    if (M > 0) {
      auto mk = conv_on_ctx(e, e->gMB, 1, 1, e->P(p + "qkv.w"), e->F(p + "qkv.b"), QP, n);    // K | V of the bank rows (their Q columns are unused)
      out_rows(mk, e->gMKV, M, QP);
      TRY(run_conv(e, mk, st, tc));
    }
    TRY(launch_emformer_attention_mem(e->gQKV, M > 0 ? e->gMKV : nullptr, e->eRing[l], e->ePast, e->gATT.new_rows(), n, ids, seg, n_utt, rc,
                                      c.left_context, e->ring_rows, D, c.emformer_heads, QP, M, st));
    auto o = conv_on_ctx(e, e->gATT, 1, 1, e->P(p + "out.w"), e->F(p + "out.b"), DP, n);      // out_proj; + residual (the summary row's is 0)
    out_rows(o, e->gR1, er, DP); res_rows(o, e->gX, er, DP);
    TRY(run_conv(e, o, st, tc));
    if (M > 0) {
      TRY(launch_emformer_mem_update(e->gR1, DP, e->gMEM[cur], e->gMEM[cur ^ 1], e->eBank[l], n, ids, seg, rc, D, M, st));
      cur ^= 1;
    }
    TRY(ln_rows(e->gR1, er, DP, 0, e->gFN.new_rows(), e->F(p + "ffn_ln.g"), e->F(p + "ffn_ln.b"), D, rows, n, st));
    auto f1 = conv_on_ctx(e, e->gFN, 1, 1, e->P(p + "ffn1.w"), e->F(p + "ffn1.b"), F, n);
    out2_ctx(f1, e->gHF, ACT_NONE, 0.f); f1.act = ACT_RELU;
    TRY(run_conv(e, f1, st, tc));
    auto f2 = conv_on_ctx(e, e->gHF, 1, 1, e->P(p + "ffn2.w"), e->F(p + "ffn2.b"), DP, n);
    out_rows(f2, e->gR2, er, DP); res_rows(f2, e->gR1, er, DP);
    TRY(run_conv(e, f2, st, tc));
    // output norm over the real rows -> next layer's input (row `rows` of gX stays zero: the summary has no residual)
    const bool last = (l == c.emformer_layers - 1);
    TRY(ln_rows(e->gR2, er, DP, 0, view_f32(e->gX, (long long)er * DP, DP), e->F(p + "ln_out.g"), e->F(p + "ln_out.b"), D, rows, n, st,
                nullptr, nullptr, 0, nullptr, nullptr, last ? e->gXN.new_rows() : RowView{}));
  }
  TRY(launch_advance_past_len_by(e->ePast, n, ids, n_utt, st));
  auto pj = conv_on_ctx(e, e->gXN, 1, 1, e->P("emf.proj.w"), e->F("emf.proj.b"), LP, n, true, rc, seg);   // utterance rows
  out_rows(pj, e->eLOG, seg, LP);
  TRY(run_conv(e, pj, st, tc));
  // tokens of the real rows only, straight into the caller's layout
  if (tokens_out || n_utt == seg) {
    if (n_utt == seg && out_nrows == seg && out_row0 == 0) {
      TRY(launch_argmax_rows(e->eLOG, LP, e->TOK, tokens_out, n, seg, c.emformer_output_dim, st));
    } else {
      TRY(launch_argmax_rows(e->eLOG, LP, e->TOK, nullptr, n, seg, c.emformer_output_dim, st));
      if (tokens_out)
        CONAN_CUDA_OK(cudaMemcpy2DAsync(tokens_out + out_row0, (size_t)out_nrows * 4, e->TOK, (size_t)seg * 4, (size_t)n_utt * 4, n,
                                        cudaMemcpyDeviceToDevice, st));
    }
  }
  if (enc_out)
    TRY(launch_copy_rows_strided(e->gX, (long long)er * DP, DP, rc, enc_out, (long long)out_nrows * D, D, out_row0, n, n_utt, D, st));
  if (logits_out)
    TRY(launch_copy_rows_strided(e->eLOG, (long long)seg * LP, LP, 0, logits_out, (long long)out_nrows * c.emformer_output_dim,
                                 c.emformer_output_dim, out_row0, n, n_utt, c.emformer_output_dim, st));
  return 0;
}

// ============================================================================ Conan chunk step
int decoder_step(conan_engine* e, int n, const int* ids, const int* tokens_ext, float* mel_out, cudaStream_t st) {
  const conan_config_t& c = e->cfg;
  const int H = c.hidden_size, seg = c.segment;
  const bool tc = e->lin_tc;
  const int* tok = tokens_ext ? tokens_ext : e->TOK;            // compact [n, seg]
  TRY(launch_hist_gather(e->histConan, e->nHistConan, n, ids, st));
  TRY(launch_embedding_rows(tok, e->F("conan.content_embedding"), 102, e->cC.new_rows(), n, nullptr, seg, H, st));
  {
    auto p = conv_on_ctx(e, e->cC, c.content_kernel, 1, e->P("conan.content_proj.w"), e->F("conan.content_proj.b"), H, n);
    out_rows(p, e->dX0, seg, H); p.act = ACT_LRELU; p.slope = 0.01f;
    p.res = e->sSTYLEW; p.res_slot_stride = H; p.res_row_stride = 0;                 // + style_embed (Conan.py:162)
    out2_ctx(p, e->cX0, ACT_NONE, 0.f);
    TRY(run_conv(e, p, st, tc));
  }
  const float* cur = e->dX0;
  const Ctx* curc = &e->cX0;
  bool pinp_done = false;
  for (int l = 0; l < 2; ++l) {
    std::string a = "conan.align." + std::to_string(l) + ".";
    auto q = conv_on_ctx(e, *curc, 1, 1, e->P(a + "q.w"), e->F(a + "q.b"), H, n);
    out_rows(q, e->dQ, seg, H);
    TRY(run_conv(e, q, st, tc));
    TRY(launch_cross_attention(e->dQ, e->sKV, e->sKPM, e->sNKEYS, e->cATT.new_rows(), n, ids, seg, H, 2, l, 2, e->tp_max, st));
    auto o = conv_on_ctx(e, e->cATT, 1, 1, e->P(a + "out.w"), e->F(a + "out.b"), H, n);
    out_rows(o, e->dT1, seg, H); res_rows(o, cur, seg, H);
    TRY(run_conv(e, o, st, tc));
    TRY(ln_rows(e->dT1, seg, H, 0, e->cO1.new_rows(), e->F(a + "norm1.g"), e->F(a + "norm1.b"), H, seg, n, st, nullptr, nullptr, 0,
                nullptr, nullptr, view_f32(e->dO1, (long long)seg * H, H)));
    if (e->blockFused) {
      // feed-forward 256 -> 2048 ReLU -> 256 in one kernel; the LayerNorm sums its partials with b2 and the residual
      auto f = block_on_ctx(e, e->cO1, 1, e->P(a + "ffn1.w"), e->F(a + "ffn1.b"), 2048, 1.f, ACT_RELU, e->P(a + "ffn2.w"), n);
      const bool cl = e->blockCluster && f.FS == 4;
      if (cl) {
        // finished rows, summed in-cluster, and norm2 applied by the reducing warps: operand rows for the next GEMM + an fp32 copy
        f.out = e->dT2; f.res = e->dO1; f.b2 = e->F(a + "ffn2.b"); f.mask = nullptr; f.ld = H;
        f.ln_g = e->F(a + "norm2.g"); f.ln_b = e->F(a + "norm2.b"); f.ln_eps = 1e-5f;
        f.ln_out = e->cPROS[l].new_rows(); f.ln_out2 = e->dPROS[l]; f.ln_out2_ld = H;
        if (l == 1) {                              // + content rows (Conan.py:168): the pitch predictor's input, no separate add launch
          f.ln_add = e->dX0; f.ln_out = e->cUV[0].new_rows(); f.ln_out2 = e->dPINP;
          pinp_done = true;
        }
      }
      TRY(run_block(e, f, st));
      if (cl) {
        cur = e->dPROS[l]; curc = &e->cPROS[l];
        continue;
      }
      LnArgs ln;
      ln.in = RowView{(void*)e->dO1, (long long)seg * H, H, 0, 0, 0};
      ln.out = e->cPROS[l].new_rows(); ln.out2 = view_f32(e->dPROS[l], (long long)seg * H, H);
      ln.gamma = e->F(a + "norm2.g"); ln.beta = e->F(a + "norm2.b"); ln.eps = 1e-5f; ln.C = H; ln.L = seg; ln.n = n; ln.slot_ids = nullptr;
      ln.premask = nullptr; ln.premask_slot_stride = 0; ln.postmask = nullptr; ln.postmask_slot_stride = 0;
      ln.write_mask = nullptr; ln.write_mask_slot_stride = 0; ln.write_mask2 = nullptr;
      ln.part = e->bfP; ln.n_part = f.FS; ln.part_stride = (long long)n * seg * H; ln.part_ld = H;
      ln.part_bias = e->F(a + "ffn2.b"); ln.part_res = e->dO1; ln.part_res_ld = H;
      TRY(launch_layernorm(ln, st));
    } else {
    auto f1 = conv_on_ctx(e, e->cO1, 1, 1, e->P(a + "ffn1.w"), e->F(a + "ffn1.b"), 2048, n);
      out2_ctx(f1, e->cHF, ACT_NONE, 0.f); f1.act = ACT_RELU;
      TRY(run_conv(e, f1, st, tc));
      auto f2 = conv_on_ctx(e, e->cHF, 1, 1, e->P(a + "ffn2.w"), e->F(a + "ffn2.b"), H, n);
      out_rows(f2, e->dT2, seg, H); res_rows(f2, e->dO1, seg, H);
      TRY(run_conv(e, f2, st, tc));
      TRY(ln_rows(e->dT2, seg, H, 0, e->cPROS[l].new_rows(), e->F(a + "norm2.g"), e->F(a + "norm2.b"), H, seg, n, st, nullptr, nullptr, 0,
                  nullptr, nullptr, view_f32(e->dPROS[l], (long long)seg * H, H)));
    }
    cur = e->dPROS[l]; curc = &e->cPROS[l];
  }
  if (!pinp_done) TRY(launch_add_rows(e->dX0, cur, e->dPINP, e->cUV[0].new_rows(), n, nullptr, seg, H, st));  // Conan.py:168
  for (int i = 0; i < 5; ++i) {                                                                // uv_predictor convs
    std::string u = "conan.uv." + std::to_string(i) + ".";
    auto p = conv_on_ctx(e, e->cUV[i], c.predictor_kernel, 1, e->P(u + "w"), e->F(u + "b"), 128, n);
    p.act = ACT_RELU;
    if (i < 4) out2_ctx(p, e->cUV[i + 1], ACT_NONE, 0.f); else out_rows(p, e->dUVH, seg, 128);
    TRY(run_conv(e, p, st, tc));
  }
  TRY(launch_pitch(e->dUVH, e->F("conan.uv.ln.g"), e->F("conan.uv.ln.b"), e->F("conan.uv.lin.w"), e->F("conan.uv.lin.b"), tok,
                   c.silent_token, e->F("conan.pitch_embed"), e->dPINP, e->dDECX, e->dUVP, n, nullptr, seg, 128, H, st));
  if (e->blockFused) {
    // per residual block body: LayerNorm -> ONE kernel (causal conv k5 -> x 5^-1/2 -> GELU -> 1x1); the next LayerNorm assembles
    // x = (x + y + b2) * nonpadding from the kernel's partial outputs, writes it back as the residual stream and normalises it
    int fs_prev = 0; std::string d_prev;
    auto ln_assemble = [&](LnArgs& ln) {
      if (!fs_prev || e->blockCluster) return;       // cluster mode: the kernel already wrote the finished residual stream
      ln.part = e->bfP; ln.n_part = fs_prev; ln.part_stride = (long long)n * seg * H; ln.part_ld = H;
      ln.part_bias = e->F(d_prev + "pw.b"); ln.part_res = e->dDECX; ln.part_res_ld = H;
      ln.part_mask = e->dMASKB; ln.part_out = e->dDECX; ln.part_out_ld = H;
    };
    const bool fold = e->blockCluster;               // cluster mode: every LayerNorm but the first is applied by the previous block kernel
    for (int b = 0; b < c.dec_blocks; ++b)
      for (int s = 0; s < 2; ++s) {
        std::string d = "conan.dec." + std::to_string(b) + "." + std::to_string(s) + ".";
        if (!fold || (b == 0 && s == 0)) {
          LnArgs ln;
          ln.in = RowView{(void*)e->dDECX, (long long)seg * H, H, 0, 0, 0};
          ln.out = e->cD[b][s].new_rows(); ln.out2 = RowView{};
          ln.gamma = e->F(d + "ln.g"); ln.beta = e->F(d + "ln.b"); ln.eps = 1e-5f; ln.C = H; ln.L = seg; ln.n = n; ln.slot_ids = nullptr;
          ln.premask = nullptr; ln.premask_slot_stride = seg; ln.postmask = nullptr; ln.postmask_slot_stride = seg;
          ln.write_mask = s == 0 ? e->dMASKB : nullptr; ln.write_mask_slot_stride = seg;
          ln.write_mask2 = (b == 0 && s == 0) ? e->dMASK0 : nullptr;
          ln_assemble(ln);
          TRY(launch_layernorm(ln, st));
        }
        auto f = block_on_ctx(e, e->cD[b][s], c.dec_kernel, e->P(d + "conv.w"), e->F(d + "conv.b"), 2 * H, 1.0f / sqrtf((float)c.dec_kernel),
                              ACT_GELU, e->P(d + "pw.w"), n);
        if (e->blockCluster && f.FS == 4) { f.out = e->dDECX; f.res = e->dDECX; f.b2 = e->F(d + "pw.b"); f.mask = e->dMASKB; f.ld = H; }
        else if (e->blockCluster) { set_error("internal: fused decoder block without 4 hidden slices"); return 1; }
        if (fold) {
          // the LayerNorm that consumes this block's output: the next block's input norm (its mask is taken at s == 0), or last_norm
          const bool last = b == c.dec_blocks - 1 && s == 1;
          f.ln_eps = 1e-5f;
          if (!last) {
            const int nb = s == 0 ? b : b + 1, ns = s == 0 ? 1 : 0;
            std::string dn = "conan.dec." + std::to_string(nb) + "." + std::to_string(ns) + ".";
            f.ln_g = e->F(dn + "ln.g"); f.ln_b = e->F(dn + "ln.b"); f.ln_out = e->cD[nb][ns].new_rows();
            f.ln_wmask = ns == 0 ? e->dMASKB : nullptr;
          } else {
            f.ln_g = e->F("conan.dec.last_norm.g"); f.ln_b = e->F("conan.dec.last_norm.b"); f.ln_out = e->cP.new_rows();
            f.ln_pre = e->dMASK0; f.ln_post = e->dMASK0;
          }
        }
        TRY(run_block(e, f, st));
        fs_prev = f.FS; d_prev = d;
      }
    if (!fold) {
      LnArgs ln;
      ln.in = RowView{(void*)e->dDECX, (long long)seg * H, H, 0, 0, 0};
      ln.out = e->cP.new_rows(); ln.out2 = RowView{};
      ln.gamma = e->F("conan.dec.last_norm.g"); ln.beta = e->F("conan.dec.last_norm.b"); ln.eps = 1e-5f; ln.C = H; ln.L = seg; ln.n = n;
      ln.slot_ids = nullptr; ln.premask = e->dMASK0; ln.premask_slot_stride = seg; ln.postmask = e->dMASK0; ln.postmask_slot_stride = seg;
      ln.write_mask = nullptr; ln.write_mask_slot_stride = seg; ln.write_mask2 = nullptr;
      ln_assemble(ln);
      TRY(launch_layernorm(ln, st));
    }
  } else {
    for (int b = 0; b < c.dec_blocks; ++b)
      for (int s = 0; s < 2; ++s) {
        std::string d = "conan.dec." + std::to_string(b) + "." + std::to_string(s) + ".";
        TRY(ln_rows(e->dDECX, seg, H, 0, e->cD[b][s].new_rows(), e->F(d + "ln.g"), e->F(d + "ln.b"), H, seg, n, st, nullptr, nullptr, seg,
                    s == 0 ? e->dMASKB : nullptr, (b == 0 && s == 0) ? e->dMASK0 : nullptr));
        auto p = conv_on_ctx(e, e->cD[b][s], c.dec_kernel, 1, e->P(d + "conv.w"), e->F(d + "conv.b"), 2 * H, n);
        out2_ctx(p, e->cDECH, ACT_NONE, 0.f); p.scale = 1.0f / sqrtf((float)c.dec_kernel); p.act = ACT_GELU;
        TRY(run_conv(e, p, st, tc));
        auto w = conv_on_ctx(e, e->cDECH, 1, 1, e->P(d + "pw.w"), e->F(d + "pw.b"), H, n);
        out_rows(w, e->dDECX, seg, H); res_rows(w, e->dDECX, seg, H); w.rowmask = e->dMASKB; w.mask_slot_stride = seg;
        TRY(run_conv(e, w, st, tc));
      }
    TRY(ln_rows(e->dDECX, seg, H, 0, e->cP.new_rows(), e->F("conan.dec.last_norm.g"), e->F("conan.dec.last_norm.b"), H, seg, n, st,
                e->dMASK0, e->dMASK0, seg));
  }
  {
    auto p = conv_on_ctx(e, e->cP, c.dec_post_kernel, 1, e->P("conan.dec.post.w"), e->F("conan.dec.post.b"), H, n);
    p.rowmask = e->dMASK0; p.mask_slot_stride = seg;
    if (tc) {
      // mel projection on tensor cores too: N padded to 96 (zero weight rows); its epilogue also writes the vocoder's input rows
      out2_ctx(p, e->cPOSTO, ACT_NONE, 0.f);
      TRY(run_conv(e, p, st, tc));
      auto m = conv_on_ctx(e, e->cPOSTO, 1, 1, e->P("conan.mel_out.w"), e->F("conan.mel_out.b"), e->MP, n);
      out_rows(m, e->dMELP, seg, e->MP);
      if (e->melDirect) out2_ctx(m, e->vPRE, ACT_NONE, 0.f);
      TRY(run_conv(e, m, st, tc));
      if (!e->melDirect)
        CONAN_CUDA_OK(cudaMemcpy2DAsync(e->dMEL, (size_t)c.n_mels * 4, e->dMELP, (size_t)e->MP * 4, (size_t)c.n_mels * 4, (size_t)n * seg,
                                        cudaMemcpyDeviceToDevice, st));
    } else {
      out_rows(p, e->dPOST, seg, H);
      TRY(run_conv(e, p, st, tc));
      auto m = conv_on_rows(e, e->dPOST, seg, 0, seg, H, e->P("conan.mel_out.w"), e->F("conan.mel_out.b"), c.n_mels, n);
      out_rows(m, e->dMEL, seg, c.n_mels);
      TRY(run_conv(e, m, st));
    }
  }
  TRY(launch_hist_scatter(e->histConan, e->nHistConan, n, ids, st));
  if (mel_out) {
    if (tc && e->melDirect)
      CONAN_CUDA_OK(cudaMemcpy2DAsync(mel_out, (size_t)c.n_mels * 4, e->dMELP, (size_t)e->MP * 4, (size_t)c.n_mels * 4, (size_t)n * seg,
                                      cudaMemcpyDeviceToDevice, st));
    else
      CONAN_CUDA_OK(cudaMemcpyAsync(mel_out, e->dMEL, (size_t)n * seg * c.n_mels * 4, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

// ============================================================================ vocoder step
// mel: compact fp32 rows [n, seg, n_mels] of the streams ids[0..n)
int vocoder_pass(conan_engine* e, int n, const int* ids, const float* mel, float* wav_out, cudaStream_t st) {
  const conan_config_t& c = e->cfg;
  const float sl = 0.1f;
  // from_ctx: x_j is not kept as an fp32 stream; conv c2_j reads it back from lrelu(x_j), the rows conv c1_j consumed
  const bool from_ctx = c.voc_residual_from_ctx != 0;
  TRY(launch_hist_gather(e->histVoc, e->nHistVoc, n, ids, st));
  if (mel) TRY(launch_rows_to_view(mel, e->vPRE.new_rows(), n, nullptr, c.segment, c.n_mels, st));   // null: the mel projection wrote them
  {
    auto p = conv_on_ctx(e, e->vPRE, 7, 1, e->P("voc.pre.w"), e->F("voc.pre.b"), e->vC[0], n);
    out2_ctx(p, e->vUP[0], ACT_LRELU, sl);
    TRY(run_conv(e, p, st, e->cfg.voc_use_tensor_cores != 0));
  }
  for (int i = 0; i < c.voc_n_ups; ++i) {
    const int r_up = c.voc_rates[i], L = e->vL[i + 1], C = e->vC[i + 1];
    std::string u = "voc.up." + std::to_string(i) + ".";
    {
      // conv -> pixel shuffle folded into the weight row order: output row t holds r_up consecutive
      // output frames, i.e. [i, L_in, r*C] viewed as [i, L_in*r, C]  (hifigan_causal.py:186-188)
      auto p = conv_on_ctx(e, e->vUP[i], c.voc_up_kernels[i], 1, e->P(u + "w"), e->F(u + "b"), r_up * C, n);
      if (!from_ctx) { p.y = e->vXS; p.y_slot_stride = (long long)L * C; p.y_row_stride = r_up * C; p.y_row0 = 0; }
      p.y2 = e->vXA[i].at_row(e->vXA[i].H); p.y2_slot_stride = e->vXA[i].slot_stride(); p.y2_row_stride = r_up * C; p.y2_row0 = 0;
      p.y2_is_half = e->vXA[i].is_half ? 1 : 0; p.act2 = ACT_LRELU; p.slope2 = sl;
      if (e->vXA[i].is_half == 2) { p.y2_split = 1; p.y2_lo_off = e->vXA[i].plane; }
      TRY(run_conv(e, p, st, e->cfg.voc_use_tensor_cores != 0));
    }
    const bool last_scale = (i == c.voc_n_ups - 1);
    const Ctx& next = last_scale ? e->vPOST : e->vUP[i + 1];
    for (int r = 0; r < c.voc_n_res && e->vFused[i]; ++r) {
      // MRF average with the running sum in fp16: block 0 writes it, block 1 adds to it, the last block emits
      // lrelu(sum / n_res) into the next layer's context (hifigan_causal.py:324-329)
      ResblockFusedParams f;
      memset(&f, 0, sizeof(f));
      const Ctx& in = e->vXA[i];
      f.x = in.p; f.x_slot_stride = in.slot_stride(); f.x_rows = in.rows(); f.x_hist_rows = in.H; f.n_slots = e->S;
      f.C = C; f.L = L; f.k = c.voc_res_kernels[r]; f.n_streams = n; f.slot_ids = ids;
      for (int j = 0; j < 3; ++j) f.dil[j] = c.voc_res_dilations[j];
      f.w = e->fW[i][r]; f.bias = e->fB[i][r]; f.w_copies = fused_weight_copies();
      f.hist = e->fHist[i][r]; f.hist_slot_stride = (long long)e->fHistRows[i][r] * C; f.hist_out = e->fHistOut[i][r];
      f.sum_in = r > 0 ? e->vSUMh : nullptr;
      f.sum_out = r < c.voc_n_res - 1 ? e->vSUMh : nullptr;
      if (r == c.voc_n_res - 1) { f.next = next.at_row(next.H); f.next_slot_stride = next.slot_stride(); f.next_row0 = 0; }
      f.out_scale = 1.0f / (float)c.voc_n_res; f.slope = sl;
      TRY(run_fused(e, f, st));
    }
    // per-conv path.  The three MRF branches are independent chains until their outputs are summed, so the j-th conv of every branch
    // goes out as ONE grouped launch of the CTA-pair kernel (launch_conv_gemm_tc_group) where the layer qualifies; only the last conv of
    // each branch stays a launch of its own (the running sum orders them).  Without grouping the branches run on three streams.
    static const int group_env = [] { const char* v = getenv("CONAN_VOC_GROUP"); return v ? atoi(v) : 1; }();
    const bool fast16 = from_ctx && e->vXA[i].is_half == 1;
    const bool grouped = !e->vFused[i] && group_env && fast16 && c.voc_n_res > 1 && c.voc_n_res <= 3 && e->cfg.voc_use_tensor_cores;
    const bool multi = !e->vFused[i] && !grouped && e->vocStreams > 1 && c.voc_n_res > 1 && c.voc_n_res <= 4 && !e->profiling && fast16;
    auto make_p1 = [&](int r, int j) {
      const Ctx& in1 = (j == 0) ? e->vXA[i] : e->vC1[i][r][j];
      std::string q = "voc.res." + std::to_string(i) + "." + std::to_string(r) + ".";
      auto p1 = conv_on_ctx(e, in1, c.voc_res_kernels[r], c.voc_res_dilations[j], e->P(q + "c1." + std::to_string(j) + ".w"),
                            e->F(q + "c1." + std::to_string(j) + ".b"), C, n);
      out2_ctx(p1, e->vC2[i][r][j], ACT_LRELU, sl);
      return p1;
    };
    const float* xj_of[4] = {e->vXS, e->vXS, e->vXS, e->vXS};
    auto make_p2 = [&](int r, int j) {
      const Ctx& in1 = (j == 0) ? e->vXA[i] : e->vC1[i][r][j];
      std::string q = "voc.res." + std::to_string(i) + "." + std::to_string(r) + ".";
      auto p2 = conv_on_ctx(e, e->vC2[i][r][j], c.voc_res_kernels[r], 1, e->P(q + "c2." + std::to_string(j) + ".w"),
                            e->F(q + "c2." + std::to_string(j) + ".b"), C, n);
      if (from_ctx) {
        p2.res = (const float*)in1.at_row(in1.H); p2.res_slot_stride = in1.slot_stride(); p2.res_row_stride = C;
        p2.res_is_half = in1.is_half ? 1 : 0; p2.res_inv_slope = 1.0f / sl;
      } else {
        res_rows(p2, xj_of[r], L, C);
      }
      if (j + 1 < c.voc_n_dil) {
        if (!from_ctx) { float* xn = e->vXR[j & 1]; out_rows(p2, xn, L, C); xj_of[r] = xn; }
        out2_ctx(p2, e->vC1[i][r][j + 1], ACT_LRELU, sl);
      } else if (fast16) {
        // MRF average (hifigan_causal.py:324-329) with the running sum carried in fp16: resblock 0 writes it, 1 adds to it,
        // the last one only reads it and emits lrelu(sum / 3) into the next layer's context
        if (r > 0) { p2.res2 = e->vSUMh; p2.res2_slot_stride = (long long)L * C; p2.res2_row_stride = C; p2.res2_is_half = 1; }
        if (r < c.voc_n_res - 1) {
          p2.y = (float*)e->vSUMh; p2.y_slot_stride = (long long)L * C; p2.y_row_stride = C; p2.y_row0 = 0; p2.y_is_half = 1;
        } else {
          p2.out_scale = 1.0f / (float)c.voc_n_res;
          out2_ctx(p2, next, ACT_LRELU, sl);
        }
      } else {
        out_rows(p2, e->vSUM, L, C);
        p2.out_scale = 1.0f / (float)c.voc_n_res; p2.accumulate = r > 0;                  // MRF average (hifigan_causal.py:324-329)
        if (r == c.voc_n_res - 1) out2_ctx(p2, next, ACT_LRELU, sl);
      }
      return p2;
    };
    if (grouped) {
      conan_conv_params_t ps[4];
      for (int j = 0; j < c.voc_n_dil; ++j) {
        for (int r = 0; r < c.voc_n_res; ++r) ps[r] = make_p1(r, j);
        TRY(run_conv_group(e, ps, c.voc_n_res, st));
        for (int r = 0; r < c.voc_n_res; ++r) ps[r] = make_p2(r, j);
        if (j + 1 < c.voc_n_dil) { TRY(run_conv_group(e, ps, c.voc_n_res, st)); continue; }
        // last convs of the branches: x_out = sum_r (c2_r(.) + x_r) / n_res is ONE accumulator -- every tile runs the k-blocks of all
        // branches into it, the epilogue adds the biases and the branches' residual rows and emits lrelu(sum / n_res) into the next
        // layer's context (no running-sum tensor, one launch instead of n_res ordered ones)
        static const int sum_env = [] { const char* v = getenv("CONAN_VOC_SUM"); return v ? atoi(v) : 1; }();
        int rc_sum = -1;
        if (sum_env) {
          conan_conv_params_t qs[4];
          for (int r = 0; r < c.voc_n_res; ++r) {
            const Ctx& in1 = e->vC1[i][r][j];
            std::string q = "voc.res." + std::to_string(i) + "." + std::to_string(r) + ".";
            qs[r] = conv_on_ctx(e, e->vC2[i][r][j], c.voc_res_kernels[r], 1, e->P(q + "c2." + std::to_string(j) + ".w"),
                                e->F(q + "c2." + std::to_string(j) + ".b"), C, n);
            qs[r].res = (const float*)in1.at_row(in1.H); qs[r].res_slot_stride = in1.slot_stride(); qs[r].res_row_stride = C;
            qs[r].res_is_half = 1; qs[r].res_inv_slope = 1.0f / sl;
          }
          qs[c.voc_n_res - 1].out_scale = 1.0f / (float)c.voc_n_res;
          out2_ctx(qs[c.voc_n_res - 1], next, ACT_LRELU, sl);
          rc_sum = j > 0 ? run_conv_group(e, qs, c.voc_n_res, st, true) : -1;
          if (rc_sum > 0) return rc_sum;
        }
        if (rc_sum < 0)
          for (int r = 0; r < c.voc_n_res; ++r) TRY(run_conv(e, ps[r], st, true));          // the running sum orders these
      }
    } else if (!e->vFused[i]) {
      if (multi) {
        CONAN_CUDA_OK(cudaEventRecord(e->evFork, st));
        for (int q = 0; q < e->vocStreams - 1; ++q) CONAN_CUDA_OK(cudaStreamWaitEvent(e->branchStream[q], e->evFork, 0));
      }
      for (int r = 0; r < c.voc_n_res; ++r) {
        const int lane = multi ? r % e->vocStreams : 0;
        cudaStream_t bs = lane == 0 ? st : e->branchStream[lane - 1];
        for (int j = 0; j < c.voc_n_dil; ++j) {
          auto p1 = make_p1(r, j);
          TRY(run_conv(e, p1, bs, e->cfg.voc_use_tensor_cores != 0));
          auto p2 = make_p2(r, j);
          const bool last_conv = j + 1 == c.voc_n_dil;
          if (multi && last_conv && r > 0) CONAN_CUDA_OK(cudaStreamWaitEvent(bs, e->evBranch[r - 1], 0));   // the running sum of branch r - 1
          TRY(run_conv(e, p2, bs, e->cfg.voc_use_tensor_cores != 0));
          if (multi && last_conv) CONAN_CUDA_OK(cudaEventRecord(e->evBranch[r], bs));
        }
      }
      if (multi) CONAN_CUDA_OK(cudaStreamWaitEvent(st, e->evBranch[c.voc_n_res - 1], 0));             // join (the chain of sums implies every branch)
    }
  }
  const int Lw = e->vL[c.voc_n_ups];
  TRY(launch_conv_post_tanh(e->vPOST.p, e->vPOST.is_half ? 1 : 0, e->vPOST.slot_stride(), e->vPOST.C, e->vPOST.H - 6, Lw, e->vPOST.C, 7,
                            e->F("voc.post.w"), e->F("voc.post.b"), wav_out, n, nullptr, st,
                            e->postTapsHost.empty() ? nullptr : e->postTapsHost.data(), e->vPOST.is_half == 2 ? e->vPOST.plane : 0));
  TRY(launch_hist_scatter(e->histVoc, e->nHistVoc, n, ids, st));
  return 0;
}

int vocoder_step(conan_engine* e, int n, const int* ids, const float* mel_ext, float* wav_out, cudaStream_t st) {
  const conan_config_t& c = e->cfg;
  const bool direct = !mel_ext && e->melDirect;      // decoder_step of this step already wrote vPRE's new rows
  const float* mel = mel_ext ? mel_ext : e->dMEL;
  if (direct) return vocoder_pass(e, n, ids, nullptr, wav_out, st);
  // voc_group > 0: the compact buffers [0, G) are reused by consecutive groups of streams, so one group's
  // activations (G x ~5 MB) can stay L2-resident between producer and consumer layers
  const int G = c.voc_group > 0 ? c.voc_group : n;
  const int Lw = e->vL[c.voc_n_ups];
  for (int g = 0; g < n; g += G)
    TRY(vocoder_pass(e, std::min(G, n - g), ids + g, mel + (size_t)g * c.segment * c.n_mels, wav_out + (size_t)g * Lw, st));
  return 0;
}

// ============================================================================ session setup
// ConvBlocks.forward on compact [n, TS, C] rows of which the first T are real (modules/commons/conv.py:84-125 /
// prosody_util.py:299-336).  TS > T only on the tensor-core path (tile geometry): the convs then run over TS rows per session,
// the LayerNorms over T, so the context rows T.. stay zero (the reference's zero padding) and the rows T.. of X stay zero
// through the block masks.  tc: operands are split-fp16 pairs in ctxk_h / Hbuf_h / ctx3_h (capacity 2 * SB sessions).
int conv_blocks_noncausal(conan_engine* e, const std::string& pre, float* X, int C, int k, float* ctxk, float* Hbuf, float* ctx3,
                          float* OUT, int outC, const float* nonpad, float* maskb, int n, int T, int TS, cudaStream_t st,
                          bool tc = false, __half* ctxk_h = nullptr, __half* Hbuf_h = nullptr, __half* ctx3_h = nullptr) {
  const int pad = (k - 1) / 2, SB = e->SB;
  Ctx ck; ck.p = tc ? (void*)ctxk_h : (void*)ctxk; ck.H = pad; ck.L = TS; ck.R = pad; ck.C = C; ck.is_half = tc ? 2 : 0;
  Ctx c3; c3.p = tc ? (void*)ctx3_h : (void*)ctx3; c3.H = 1; c3.L = TS; c3.R = 1; c3.C = C; c3.is_half = tc ? 2 : 0;
  Ctx hb; hb.p = Hbuf_h; hb.H = 0; hb.L = TS; hb.R = 0; hb.C = 2 * C; hb.is_half = 2;
  ck.plane = (long long)SB * ck.rows() * C; c3.plane = (long long)SB * c3.rows() * C; hb.plane = (long long)SB * hb.rows() * 2 * C;
  auto fix = [&](conan_conv_params_t& p) {          // session scratch holds SB sessions (not max_slots); lo plane SB slots further
    p.n_slots = tc ? SB : n;
    if (p.x_split) p.x_lo_slot_off = SB;
  };
  CONAN_CUDA_OK(cudaMemsetAsync(ck.p, 0, (size_t)(tc ? 2 * SB : n) * ck.rows() * C * ck.elem(), st));
  CONAN_CUDA_OK(cudaMemsetAsync(c3.p, 0, (size_t)(tc ? 2 * SB : n) * c3.rows() * C * c3.elem(), st));
  if (TS != T) CONAN_CUDA_OK(cudaMemsetAsync(maskb, 0, (size_t)n * TS * 4, st));
  for (int b = 0; b < 5; ++b)
    for (int s = 0; s < 2; ++s) {
      std::string d = pre + "." + std::to_string(b) + "." + std::to_string(s) + ".";
      TRY(ln_rows(X, TS, C, 0, ck.new_rows(), e->F(d + "ln.g"), e->F(d + "ln.b"), C, T, n, st, nullptr, nullptr, TS,
                  s == 0 ? maskb : nullptr, nullptr));
      auto p = conv_on_ctx(e, ck, k, 1, e->P(d + "conv.w"), e->F(d + "conv.b"), 2 * C, n, false);
      fix(p); p.scale = 1.0f / sqrtf((float)k); p.act = ACT_GELU;
      if (tc) out2_ctx(p, hb, ACT_NONE, 0.f); else out_rows(p, Hbuf, TS, 2 * C);
      TRY(run_conv(e, p, st, tc));
      auto w = tc ? conv_on_ctx(e, hb, 1, 1, e->P(d + "pw.w"), e->F(d + "pw.b"), C, n)
                  : conv_on_rows(e, Hbuf, TS, 0, TS, 2 * C, e->P(d + "pw.w"), e->F(d + "pw.b"), C, n);
      fix(w); out_rows(w, X, TS, C); res_rows(w, X, TS, C); w.rowmask = maskb; w.mask_slot_stride = TS;
      TRY(run_conv(e, w, st, tc));
    }
  TRY(ln_rows(X, TS, C, 0, c3.new_rows(), e->F(pre + ".last_norm.g"), e->F(pre + ".last_norm.b"), C, T, n, st, nonpad, nonpad, TS));
  auto p = conv_on_ctx(e, c3, 3, 1, e->P(pre + ".post.w"), e->F(pre + ".post.b"), outC, n, false);
  fix(p); out_rows(p, OUT, TS, outC); p.rowmask = nonpad; p.mask_slot_stride = TS;
  TRY(run_conv(e, p, st, tc));
  return 0;
}

int session_open_batch(conan_engine* e, int n, const int* slots_host, const float* ref, int T, cudaStream_t st) {
  const conan_config_t& c = e->cfg;
  const int H = c.hidden_size, M = c.n_mels;
  const int Tp = (T - 1) / 4 + 1;
  CONAN_CUDA_OK(cudaMemcpyAsync(e->qSlots, slots_host, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  TRY(launch_row_masks(ref, e->qMA, e->qMF, n, T, M, st));
  // ---- global style encoder (Conan.py:200-219)
  {
    // tensor-core path: sessions padded to TS = multiple of 32 rows (rows T.. of X and of the masks are zero)
    const bool tc = e->ses_tc;
    const int TS = tc ? (T + 31) / 32 * 32 : T;
    const float* mask = e->qMA;
    if (TS != T) {
      CONAN_CUDA_OK(cudaMemsetAsync(e->qXG, 0, (size_t)n * TS * H * 4, st));
      CONAN_CUDA_OK(cudaMemsetAsync(e->qMAp, 0, (size_t)n * TS * 4, st));
      CONAN_CUDA_OK(cudaMemcpy2DAsync(e->qMAp, (size_t)TS * 4, e->qMA, (size_t)T * 4, (size_t)T * 4, n, cudaMemcpyDeviceToDevice, st));
      mask = e->qMAp;
    }
    auto p = conv_on_rows(e, ref, T, 0, T, M, e->P("conan.global_in.w"), e->F("conan.global_in.b"), H, n);
    p.n_slots = n; out_rows(p, e->qXG, TS, H); p.rowmask = e->qMA; p.mask_slot_stride = T;
    TRY(run_conv(e, p, st));
    TRY(conv_blocks_noncausal(e, "conan.genc", e->qXG, H, 31, e->qC31, e->qHG, e->qC3G, e->qPG, H, mask, e->qMGB, n, T, TS, st,
                              tc, e->qC31h, e->qHGh, e->qC3Gh));
    TRY(launch_masked_time_mean(e->qPG, mask, e->sSTYLE, e->qSlots, n, T, TS, H, st));
  }
  // ---- LocalStyleAdaptor: WN (wavenet.py:55-88)
  {
    Ctx cw; cw.p = e->qCW; cw.H = 1; cw.L = T; cw.R = 1; cw.C = M; cw.is_half = 0;
    CONAN_CUDA_OK(cudaMemsetAsync(e->qCW, 0, (size_t)n * cw.rows() * M * 4, st));
    CONAN_CUDA_OK(cudaMemsetAsync(e->qSKIP, 0, (size_t)n * T * M * 4, st));
    CONAN_CUDA_OK(cudaMemcpyAsync(e->qXW, ref, (size_t)n * T * M * 4, cudaMemcpyDeviceToDevice, st));
    CONAN_CUDA_OK(cudaMemcpy2DAsync((float*)cw.p + M, (size_t)cw.rows() * M * 4, ref, (size_t)T * M * 4, (size_t)T * M * 4, n,
                                    cudaMemcpyDeviceToDevice, st));
    for (int i = 0; i < 4; ++i) {
      std::string w = "conan.wn." + std::to_string(i) + ".";
      auto p = conv_on_ctx(e, cw, 3, 1, e->P(w + "in.w"), e->F(w + "in.b"), 2 * M, n, false);
      p.n_slots = n; out_rows(p, e->qAW, T, 2 * M);
      TRY(run_conv(e, p, st));
      TRY(launch_gated_tanh_sigmoid(e->qAW, e->qACT, (long long)n * T, M, st));
      int co = i < 3 ? 2 * M : M;
      auto r = conv_on_rows(e, e->qACT, T, 0, T, M, e->P(w + "rs.w"), e->F(w + "rs.b"), co, n);
      r.n_slots = n; out_rows(r, e->qRS, T, co);
      TRY(run_conv(e, r, st));
      TRY(launch_wn_update(e->qRS, e->qXW, cw.new_rows(), e->qSKIP, e->qMF, (long long)n * T, T, M, i == 3, st));
    }
    TRY(launch_group_mean4(e->qSKIP, e->qMF, e->qGRP, n, T, Tp, M, st));
  }
  // ---- prosody encoder ConvBlocks(80 -> H, k5) + VQ + positions + l1 + aligner K/V
  {
    TRY(launch_row_masks(e->qGRP, e->qMP, e->qMPB, n, Tp, M, st));      // qMPB is overwritten by the block masks below
    CONAN_CUDA_OK(cudaMemcpyAsync(e->qXP, e->qGRP, (size_t)n * Tp * M * 4, cudaMemcpyDeviceToDevice, st));
    TRY(conv_blocks_noncausal(e, "conan.penc", e->qXP, M, 5, e->qC5, e->qHP, e->qC3P, e->qPZ, H, e->qMP, e->qMPB, n, Tp, Tp, st));
    auto d = conv_on_rows(e, e->qPZ, Tp, 0, Tp, H, e->P("conan.vq.embedding"), nullptr, c.n_vq, n);
    d.n_slots = n; out_rows(d, e->qXE, Tp, c.n_vq);
    TRY(run_conv(e, d, st));
    TRY(launch_vq_quantize(e->qPZ, e->qXE, e->F("conan.vq.embedding"), e->F("conan.vq.e2"), e->F("conan.pos_table"), e->qZC, e->qVQ,
                           n, Tp, H, c.n_vq, st));
    auto l1 = conv_on_rows(e, e->qZC, Tp, 0, Tp, 2 * H, e->P("conan.l1.w"), e->F("conan.l1.b"), H, n);
    l1.n_slots = n; out_rows(l1, e->qPE, Tp, H);
    TRY(run_conv(e, l1, st));
    TRY(launch_kpm(e->qPE, e->sKPM, e->sNKEYS, e->qSlots, n, Tp, H, e->tp_max, st));
    for (int l = 0; l < 2; ++l) {
      std::string a = "conan.align." + std::to_string(l) + ".";
      auto kv = conv_on_rows(e, e->qPE, Tp, 0, Tp, H, e->P(a + "kv.w"), e->F(a + "kv.b"), 2 * H, n);
      kv.n_slots = n; out_rows(kv, e->qKVs, Tp, 2 * H);
      TRY(run_conv(e, kv, st));
      TRY(launch_scatter_kv(e->qKVs, e->sKV, e->qSlots, n, Tp, 2 * H, l, 2, e->tp_max, st));
    }
  }
  return 0;
}

// A ready list handed over in HOST memory is checked before it reaches the device: an out-of-range id would be an
// out-of-bounds read / write of resident state, a duplicate would make two CTAs race on one stream's K/V ring and history.
int check_ids_host(conan_engine* e, int n, const int32_t* ids, const char* who) {
  e->idSeen.assign((size_t)e->Su, 0);
  for (int i = 0; i < n; ++i) {
    const int s = ids[i];
    if (s < 0 || s >= e->Su) { set_error(std::string(who) + ": slot id " + std::to_string(s) + " out of range [0, " + std::to_string(e->Su) + ")"); return 1; }
    if (e->idSeen[s]) { set_error(std::string(who) + ": slot id " + std::to_string(s) + " appears twice in one step"); return 1; }
    e->idSeen[s] = 1;
  }
  return 0;
}

int check_ready(conan_engine* e) {
  if (!e) { set_error("null engine"); return 1; }
  if (!e->finalized) { set_error("engine not finalized"); return 1; }
  return 0;
}

}  // namespace

// ================================================================================ C ABI
extern "C" {

const char* conan_last_error(void) { return g_last_error.c_str(); }
int conan_abi_version(void) { return CONAN_B200_ABI_VERSION; }
size_t conan_sizeof_config(void) { return sizeof(conan_config_t); }
size_t conan_sizeof_conv_params(void) { return sizeof(conan_conv_params_t); }

int conan_engine_create(const conan_config_t* cfg, conan_engine_t** out) {
  if (!cfg || !out) { set_error("null argument"); return 1; }
  if (cfg->abi_version != CONAN_B200_ABI_VERSION) { set_error("ABI version mismatch"); return 1; }
  if (cfg->max_slots <= 0 || cfg->max_ref_frames <= 0) { set_error("max_slots and max_ref_frames must be positive"); return 1; }
  if (cfg->voc_n_ups > 8 || cfg->voc_n_res > 4 || cfg->voc_n_dil > 4 || cfg->dec_blocks > 8) { set_error("config above compiled limits"); return 1; }
  if (cfg->emformer_memory_size < 0 || cfg->emformer_memory_size > 8 ||
      cfg->emformer_memory_size + cfg->right_context + cfg->left_context + cfg->segment > 64) {
    set_error("emformer_memory_size must be in [0, 8] and memory + contexts + segment at most 64 keys");
    return 1;
  }
  if (cfg->voc_precision < 0 || cfg->voc_precision > 2) { set_error("voc_precision must be 0 (fp32), 1 (fp16 operands) or 2 (split fp16 operands)"); return 1; }
  if (cfg->voc_use_tensor_cores && !cfg->voc_precision) { set_error("voc_use_tensor_cores requires voc_precision 1 (fp16 operands) or 2 (split fp16 operands)"); return 1; }
  if (cfg->voc_precision == 2 && (!cfg->voc_use_tensor_cores || cfg->voc_residual_from_ctx || cfg->voc_fuse_resblocks)) {
    set_error("voc_precision 2 (split fp16) runs on the tensor cores with an fp32 residual stream: needs voc_use_tensor_cores = 1, "
              "voc_residual_from_ctx = 0, voc_fuse_resblocks = 0");
    return 1;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= cfg->device) {
    set_error("no CUDA device: the conan_b200 engine has no CPU fallback");
    return 1;
  }
  cudaDeviceProp prop;
  CONAN_CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) { set_error(std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + ", this library is built for sm_100a only"); return 1; }
  CONAN_CUDA_OK(cudaSetDevice(cfg->device));
  conan_engine* e = new conan_engine();
  e->cfg = *cfg;
  e->Su = cfg->max_slots;
  e->S = cfg->max_slots + kPadSlots;
  e->tp_max = (cfg->max_ref_frames - 1) / 4 + 1;
  e->lin_tc = cfg->lin_use_tensor_cores != 0;
  e->ses_tc = cfg->ses_use_tensor_cores != 0;
  e->eM = cfg->emformer_memory_size;
  declare_weights(e);
  *out = e;
  return 0;
}

void conan_engine_destroy(conan_engine_t* e) {
  if (!e) return;
  if (e->copyStream) { cudaStreamSynchronize(e->copyStream); cudaStreamDestroy(e->copyStream); }
  for (int i = 0; i < 2; ++i) if (e->branchStream[i]) { cudaStreamSynchronize(e->branchStream[i]); cudaStreamDestroy(e->branchStream[i]); }
  for (auto& g : e->stepGraphs) cudaGraphExecDestroy(g.exec);
  if (e->graphStream) cudaStreamDestroy(e->graphStream);
  if (e->evFork) cudaEventDestroy(e->evFork);
  for (int i = 0; i < 4; ++i) if (e->evBranch[i]) cudaEventDestroy(e->evBranch[i]);
  for (int i = 0; i < 2; ++i) { if (e->evCompute[i]) cudaEventDestroy(e->evCompute[i]); if (e->evCopy[i]) cudaEventDestroy(e->evCopy[i]); }
  for (void* p : e->allocs) cudaFree(p);
  delete e;
}

int conan_engine_num_weights(const conan_engine_t* e) { return e ? (int)e->weights.size() : 0; }

int conan_engine_weight_info(const conan_engine_t* e, int idx, const char** name, size_t* numel, int* dtype) {
  if (!e || idx < 0 || idx >= (int)e->weights.size()) { set_error("weight index out of range"); return 1; }
  if (name) *name = e->weights[idx].name.c_str();
  if (numel) *numel = e->weights[idx].numel;
  if (dtype) *dtype = e->weights[idx].dtype;
  return 0;
}

int conan_engine_bind_weight(conan_engine_t* e, const char* name, const void* data_dev, size_t numel, int dtype) {
  if (!e || !name || !data_dev) { set_error("null argument"); return 1; }
  auto it = e->windex.find(name);
  if (it == e->windex.end()) { set_error(std::string("unknown weight '") + name + "'"); return 1; }
  WeightSlot& w = e->weights[it->second];
  if (w.numel != numel || w.dtype != dtype) {
    set_error(std::string("weight '") + name + "': expected " + std::to_string(w.numel) + " elements of dtype " + std::to_string(w.dtype) +
              ", got " + std::to_string(numel) + " of dtype " + std::to_string(dtype));
    return 1;
  }
  if (((uintptr_t)data_dev) % 16 != 0) { set_error(std::string("weight '") + name + "' is not 16-byte aligned"); return 1; }
  w.ptr = data_dev;
  return 0;
}

int conan_engine_finalize(conan_engine_t* e) {
  if (!e) { set_error("null engine"); return 1; }
  if (e->finalized) return 0;
  for (auto& w : e->weights)
    if (!w.ptr) { set_error("weight '" + w.name + "' was never bound"); return 1; }
  CONAN_CUDA_OK(cudaSetDevice(e->cfg.device));
  if (allocate_state(e)) return 1;
  {
    const size_t nt = (size_t)7 * e->vC[e->cfg.voc_n_ups];
    e->postTapsHost.resize(nt + 1);
    CONAN_CUDA_OK(cudaMemcpy(e->postTapsHost.data(), e->F("voc.post.w"), nt * sizeof(float), cudaMemcpyDeviceToHost));
    CONAN_CUDA_OK(cudaMemcpy(e->postTapsHost.data() + nt, e->F("voc.post.b"), sizeof(float), cudaMemcpyDeviceToHost));
  }
  CONAN_CUDA_OK(cudaDeviceSynchronize());
  e->finalized = true;
  {
    // the pad slots get a real (dummy) session so that every kernel sees a well-formed stream in them: 8 reference frames of a
    // constant spectrum; their outputs are never returned
    const int T = std::min(8, e->cfg.max_ref_frames), M = e->cfg.n_mels;
    std::vector<float> ref((size_t)kPadSlots * T * M);
    for (size_t i = 0; i < ref.size(); ++i) ref[i] = -3.0f + 0.01f * (float)(i % 7);
    std::vector<int> pad(kPadSlots);
    for (int i = 0; i < kPadSlots; ++i) pad[i] = e->Su + i;
    float* dref = nullptr;
    CONAN_CUDA_OK(cudaMalloc(&dref, ref.size() * sizeof(float)));
    CONAN_CUDA_OK(cudaMemcpy(dref, ref.data(), ref.size() * sizeof(float), cudaMemcpyHostToDevice));
    int rc = 0;
    for (int g = 0; g < kPadSlots && !rc; g += e->SB) {
      const int nb = std::min(e->SB, kPadSlots - g);
      rc = session_open_batch(e, nb, pad.data() + g, dref + (size_t)g * T * M, T, nullptr);
    }
    cudaError_t ce = cudaDeviceSynchronize();
    cudaFree(dref);
    if (rc || ce != cudaSuccess) { if (!rc) set_error(std::string("pad-slot session setup: ") + cudaGetErrorString(ce)); e->finalized = false; return 1; }
  }
  return 0;
}

size_t conan_engine_state_bytes(const conan_engine_t* e) { return e ? e->state_bytes : 0; }
uint64_t conan_engine_launch_count(const conan_engine_t*) { return g_launches.load(); }
uint64_t conan_engine_graph_replays(const conan_engine_t* e) { return e ? e->graphReplays : 0; }

int conan_slots_reset(conan_engine_t* e, int n, const int32_t* slots_host, int parts, void* stream) {
  if (check_ready(e)) return 1;
  if (n <= 0) return 0;
  if (n > e->Su) { set_error("more slots than max_slots"); return 1; }
  for (int i = 0; i < n; ++i) if (slots_host[i] < 0 || slots_host[i] >= e->Su) { set_error("slot id out of range"); return 1; }
  cudaStream_t st = (cudaStream_t)stream;
  CONAN_CUDA_OK(cudaMemcpyAsync(e->hIdsSmall, slots_host, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  if (parts & 1) TRY(launch_zero_slots(e->zeroEmf, e->nZeroEmf, n, e->hIdsSmall, st));
  if (parts & 2) TRY(launch_zero_slots(e->zeroConan, e->nZeroConan, n, e->hIdsSmall, st));
  if (parts & 4) TRY(launch_zero_slots(e->zeroVoc, e->nZeroVoc, n, e->hIdsSmall, st));
  // no synchronisation: the id copy above is stream-ordered (a pageable source is staged before cudaMemcpyAsync returns; a
  // page-locked one must stay valid until the stream reaches this call), and so is the reuse of hIdsSmall by the next call
  return 0;
}

int conan_session_open(conan_engine_t* e, int n, const int32_t* slots_host, const float* ref_mel_dev, int ref_frames, void* stream) {
  if (check_ready(e)) return 1;
  if (n <= 0) return 0;
  if (ref_frames < 1 || ref_frames > e->cfg.max_ref_frames) { set_error("ref_frames outside [1, max_ref_frames]"); return 1; }
  for (int i = 0; i < n; ++i) if (slots_host[i] < 0 || slots_host[i] >= e->Su) { set_error("slot id out of range"); return 1; }
  cudaStream_t st = (cudaStream_t)stream;
  for (int g = 0; g < n; g += e->SB) {
    int nb = std::min(e->SB, n - g);
    // stream-ordered, no host synchronisation: the next batch's copy into qSlots and its use of the scratch buffers queue
    // behind this batch on `st` (a serving loop opens sessions between chunk steps without stalling the host)
    TRY(session_open_batch(e, nb, slots_host + g, ref_mel_dev + (size_t)g * ref_frames * e->cfg.n_mels, ref_frames, st));
  }
  return 0;
}

int conan_emformer_step(conan_engine_t* e, int n, const int32_t* slot_ids_dev, const float* chunk_dev, float* enc_out_dev,
                        float* logits_out_dev, int32_t* tokens_out_dev, void* stream) {
  if (check_ready(e)) return 1;
  if (n < 0 || n > e->Su || !slot_ids_dev || !chunk_dev) { set_error("bad arguments to conan_emformer_step"); return 1; }
  return emformer_step(e, n, slot_ids_dev, chunk_dev, enc_out_dev, logits_out_dev, tokens_out_dev, (cudaStream_t)stream);
}

int conan_emformer_forward(conan_engine_t* e, int n, const int32_t* slots_host, const float* input_dev, int frames, float* enc_out_dev,
                           float* logits_out_dev, int32_t* tokens_out_dev, void* stream) {
  if (check_ready(e)) return 1;
  if (n <= 0) return 0;
  const conan_config_t& c = e->cfg;
  const int seg = c.segment, rc = c.right_context, D = c.emformer_dim, T = frames - rc;
  if (n > e->Su || !slots_host || !input_dev || T < 1) { set_error("bad arguments to conan_emformer_forward (needs frames > right_context)"); return 1; }
  if (check_ids_host(e, n, slots_host, "conan_emformer_forward")) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  CONAN_CUDA_OK(cudaMemcpyAsync(e->hIds, slots_host, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  TRY(launch_zero_slots(e->zeroEmf, e->nZeroEmf, n, e->hIds, st));
  // segment i: utterance rows [i*seg, min((i+1)*seg, T)); its look-ahead block is the rc rows after it, the last segment's the
  // rc rows at the end of the input (_gen_right_context, TA:619-628)
  for (int t0 = 0; t0 < T; t0 += seg) {
    const int n_utt = std::min(seg, T - t0);
    const int rc_row0 = t0 + seg < T ? t0 + seg : T;
    if (e->eM == 0 && n_utt == seg) {
      // whole segment, no memory: the fused M = 0 step on a gathered [utt | rc] chunk
      for (int q = 0; q < 2; ++q)
        CONAN_CUDA_OK(cudaMemcpy2DAsync(e->hChunk + (size_t)(q ? seg : 0) * D, (size_t)(seg + rc) * D * 4,
                                        input_dev + (size_t)(q ? rc_row0 : t0) * D, (size_t)frames * D * 4, (size_t)(q ? rc : seg) * D * 4, n,
                                        cudaMemcpyDeviceToDevice, st));
      TRY(emformer_step(e, n, e->hIds, e->hChunk, nullptr, nullptr, nullptr, st));
      // rows of the step's work buffers (eX: [rc | utt] rows of DP floats; eLOG: seg rows of LP floats) -> the caller's layout
      const int rows = seg + rc, OD = c.emformer_output_dim;
      if (enc_out_dev)
        TRY(launch_copy_rows_strided(e->eX, (long long)rows * e->DP, e->DP, rc, enc_out_dev, (long long)T * D, D, t0, n, seg, D, st));
      if (logits_out_dev)
        TRY(launch_copy_rows_strided(e->eLOG, (long long)seg * e->LP, e->LP, 0, logits_out_dev, (long long)T * OD, OD, t0, n, seg, OD, st));
      if (tokens_out_dev)
        CONAN_CUDA_OK(cudaMemcpy2DAsync(tokens_out_dev + t0, (size_t)T * 4, e->TOK, (size_t)seg * 4, (size_t)seg * 4, n,
                                        cudaMemcpyDeviceToDevice, st));
    } else {
      TRY(emformer_step_generic(e, n, e->hIds, input_dev, (long long)frames * D, t0, rc_row0, n_utt, enc_out_dev, logits_out_dev,
                                tokens_out_dev, t0, T, st));
    }
  }
  return 0;
}

int conan_decoder_step(conan_engine_t* e, int n, const int32_t* slot_ids_dev, const int32_t* tokens_dev, float* mel_out_dev, void* stream) {
  if (check_ready(e)) return 1;
  if (n < 0 || n > e->Su || !slot_ids_dev || !tokens_dev) { set_error("bad arguments to conan_decoder_step"); return 1; }
  return decoder_step(e, n, slot_ids_dev, tokens_dev, mel_out_dev, (cudaStream_t)stream);
}

int conan_vocoder_step(conan_engine_t* e, int n, const int32_t* slot_ids_dev, const float* mel_dev, float* wav_out_dev, void* stream) {
  if (check_ready(e)) return 1;
  if (n < 0 || n > e->Su || !slot_ids_dev || !mel_dev || !wav_out_dev) { set_error("bad arguments to conan_vocoder_step"); return 1; }
  return vocoder_step(e, n, slot_ids_dev, mel_dev, wav_out_dev, (cudaStream_t)stream);
}

static int step_eager(conan_engine_t* e, int n, const int32_t* slot_ids_dev, const float* chunk_dev, float* wav_out_dev, float* mel_out_dev,
                      int32_t* tokens_out_dev, cudaStream_t st) {
  TRY(emformer_step(e, n, slot_ids_dev, chunk_dev, nullptr, nullptr, tokens_out_dev, st));
  TRY(decoder_step(e, n, slot_ids_dev, nullptr, mel_out_dev, st));
  TRY(vocoder_step(e, n, slot_ids_dev, nullptr, wav_out_dev, st));
  return 0;
}

static int step_graphed(conan_engine_t* e, int n, const int32_t* slot_ids_dev, const float* chunk_dev, float* wav_out_dev, float* mel_out_dev,
                        int32_t* tokens_out_dev, cudaStream_t st);

int conan_step(conan_engine_t* e, int n, const int32_t* slot_ids_dev, const float* chunk_dev, float* wav_out_dev, float* mel_out_dev,
               int32_t* tokens_out_dev, void* stream) {
  if (check_ready(e)) return 1;
  if (n < 0 || n > e->Su || !slot_ids_dev || !chunk_dev || !wav_out_dev) { set_error("bad arguments to conan_step"); return 1; }
  return step_graphed(e, n, slot_ids_dev, chunk_dev, wav_out_dev, mel_out_dev, tokens_out_dev, (cudaStream_t)stream);
}

// pads a ready list staged in hIds to the next multiple of kStepBucket with the pad slots (graph mode only); returns the padded count
static int pad_ready_list(conan_engine_t* e, int n, cudaStream_t st) {
  if (!e->graphMode || n == 0 || n > kGraphSmallN) return n;
  const int np = std::min((n + kStepBucket - 1) / kStepBucket * kStepBucket, e->S);
  if (np > n) cudaMemcpyAsync(e->hIds + n, e->padIds, (size_t)(np - n) * sizeof(int), cudaMemcpyDeviceToDevice, st);
  return np;
}

static int step_graphed(conan_engine_t* e, int n, const int32_t* slot_ids_dev, const float* chunk_dev, float* wav_out_dev, float* mel_out_dev,
                        int32_t* tokens_out_dev, cudaStream_t st) {
  if (!e->graphMode || e->profiling || n == 0) return step_eager(e, n, slot_ids_dev, chunk_dev, wav_out_dev, mel_out_dev, tokens_out_dev, st);
  auto same = [&](const conan_engine::StepGraph& g) {
    return g.n == n && g.ids == slot_ids_dev && g.chunk == chunk_dev && g.wav == wav_out_dev && g.mel == mel_out_dev && g.tok == tokens_out_dev;
  };
  ++e->graphClock;
  for (auto& g : e->stepGraphs)
    if (same(g)) {
      g.last_use = e->graphClock;
      CONAN_CUDA_OK(cudaGraphLaunch(g.exec, st));
      count_launch((int)g.launches);
      ++e->graphReplays;
      return 0;
    }
  uint64_t sightings = 0;
  for (auto& g : e->stepSeen) if (same(g)) sightings = ++g.launches;           // (launches doubles as the sighting counter here)
  if (sightings < (uint64_t)(n <= kGraphSmallN ? 2 : 1 + kGraphBigSightings)) {
    // not met often enough yet: run eagerly (the first run also sets kernel attributes and fills the tensor-map cache)
    if (!sightings) {
      if (e->stepSeen.size() >= 64) e->stepSeen.erase(e->stepSeen.begin());
      e->stepSeen.push_back(conan_engine::StepGraph{n, slot_ids_dev, chunk_dev, wav_out_dev, mel_out_dev, tokens_out_dev, nullptr, 1, 0});
    }
    return step_eager(e, n, slot_ids_dev, chunk_dev, wav_out_dev, mel_out_dev, tokens_out_dev, st);
  }
  // second time: capture on the engine's stream (the caller's may be the legacy default stream, which cannot capture), instantiate,
  // and replay on the caller's stream from now on
  cudaGraph_t graph = nullptr;
  const uint64_t l0 = g_launches.load();
  CONAN_CUDA_OK(cudaStreamBeginCapture(e->graphStream, cudaStreamCaptureModeThreadLocal));
  const int rc = step_eager(e, n, slot_ids_dev, chunk_dev, wav_out_dev, mel_out_dev, tokens_out_dev, e->graphStream);
  cudaError_t ce = cudaStreamEndCapture(e->graphStream, &graph);
  const uint64_t nl = g_launches.load() - l0;
  g_launches.fetch_sub(nl);                              // nothing has run yet
  if (rc != 0 || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    (void)cudaGetLastError();
    e->graphMode = 0;                                    // capture is not possible in this process: stay on eager launches
    return step_eager(e, n, slot_ids_dev, chunk_dev, wav_out_dev, mel_out_dev, tokens_out_dev, st);
  }
  cudaGraphExec_t exec = nullptr;
  ce = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) { (void)cudaGetLastError(); e->graphMode = 0; return step_eager(e, n, slot_ids_dev, chunk_dev, wav_out_dev, mel_out_dev, tokens_out_dev, st); }
  if (e->stepGraphs.size() >= 32) {                      // bounded: drop the least recently used
    size_t lru = 0;
    for (size_t i = 1; i < e->stepGraphs.size(); ++i) if (e->stepGraphs[i].last_use < e->stepGraphs[lru].last_use) lru = i;
    cudaGraphExecDestroy(e->stepGraphs[lru].exec);
    e->stepGraphs.erase(e->stepGraphs.begin() + lru);
  }
  e->stepGraphs.push_back(conan_engine::StepGraph{n, slot_ids_dev, chunk_dev, wav_out_dev, mel_out_dev, tokens_out_dev, exec, nl, e->graphClock});
  CONAN_CUDA_OK(cudaGraphLaunch(exec, st));
  count_launch((int)nl);
  ++e->graphReplays;
  return 0;
}

int conan_step_host(conan_engine_t* e, int n, const int32_t* slot_ids_host, const float* chunk_host, float* wav_out_host,
                    float* mel_out_host, int32_t* tokens_out_host, void* stream) {
  if (check_ready(e)) return 1;
  if (n < 0 || n > e->Su || !slot_ids_host || !chunk_host || !wav_out_host) { set_error("bad arguments to conan_step_host"); return 1; }
  if (e->ticketPending[0] || e->ticketPending[1]) {
    // the synchronous call shares output set 0 with the pipelined one: its result copy may still be in flight
    set_error("conan_step_host: pipelined steps are in flight; call conan_step_host_wait first");
    return 1;
  }
  if (check_ids_host(e, n, slot_ids_host, "conan_step_host")) return 1;
  const conan_config_t& c = e->cfg;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rows = c.segment + c.right_context, Lw = e->vL[c.voc_n_ups];
  CONAN_CUDA_OK(cudaMemcpyAsync(e->hIds, slot_ids_host, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  CONAN_CUDA_OK(cudaMemcpyAsync(e->hChunk, chunk_host, (size_t)n * rows * c.emformer_dim * 4, cudaMemcpyHostToDevice, st));
  const int np = pad_ready_list(e, n, st);
  TRY(step_graphed(e, np, e->hIds, e->hChunk, e->hWav, mel_out_host ? e->hMel : nullptr, tokens_out_host ? e->hTok : nullptr, st));
  CONAN_CUDA_OK(cudaMemcpyAsync(wav_out_host, e->hWav, (size_t)n * Lw * 4, cudaMemcpyDeviceToHost, st));
  if (mel_out_host) CONAN_CUDA_OK(cudaMemcpyAsync(mel_out_host, e->hMel, (size_t)n * c.segment * c.n_mels * 4, cudaMemcpyDeviceToHost, st));
  if (tokens_out_host) CONAN_CUDA_OK(cudaMemcpyAsync(tokens_out_host, e->hTok, (size_t)n * c.segment * 4, cudaMemcpyDeviceToHost, st));
  CONAN_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int conan_step_host_submit(conan_engine_t* e, int n, const int32_t* slot_ids_host, const float* chunk_host, float* wav_out_host,
                           float* mel_out_host, int32_t* tokens_out_host, void* stream, int* ticket) {
  if (check_ready(e)) return 1;
  if (n < 0 || n > e->Su || !slot_ids_host || !chunk_host || !wav_out_host || !ticket) { set_error("bad arguments to conan_step_host_submit"); return 1; }
  const conan_config_t& c = e->cfg;
  cudaStream_t st = (cudaStream_t)stream;
  const int b = (int)(e->submitCount & 1);
  if (e->ticketPending[b]) { set_error("conan_step_host_submit: two steps already in flight; call conan_step_host_wait first"); return 1; }
  if (check_ids_host(e, n, slot_ids_host, "conan_step_host_submit")) return 1;
  const size_t rows = c.segment + c.right_context, Lw = e->vL[c.voc_n_ups];
  float* dWav = b ? e->hWav2 : e->hWav; float* dMel = b ? e->hMel2 : e->hMel; int* dTok = b ? e->hTok2 : e->hTok;
  CONAN_CUDA_OK(cudaMemcpyAsync(e->hIds, slot_ids_host, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  CONAN_CUDA_OK(cudaMemcpyAsync(e->hChunk, chunk_host, (size_t)n * rows * c.emformer_dim * 4, cudaMemcpyHostToDevice, st));
  // output set b was last read by the copy of ticket b two submits ago: that copy has been waited for by the caller
  // (ticketPending[b] is clear), so the step may overwrite it
  const int np = pad_ready_list(e, n, st);
  TRY(step_graphed(e, np, e->hIds, e->hChunk, dWav, mel_out_host ? dMel : nullptr, tokens_out_host ? dTok : nullptr, st));
  CONAN_CUDA_OK(cudaEventRecord(e->evCompute[b], st));
  CONAN_CUDA_OK(cudaStreamWaitEvent(e->copyStream, e->evCompute[b], 0));
  CONAN_CUDA_OK(cudaMemcpyAsync(wav_out_host, dWav, (size_t)n * Lw * 4, cudaMemcpyDeviceToHost, e->copyStream));
  if (mel_out_host) CONAN_CUDA_OK(cudaMemcpyAsync(mel_out_host, dMel, (size_t)n * c.segment * c.n_mels * 4, cudaMemcpyDeviceToHost, e->copyStream));
  if (tokens_out_host) CONAN_CUDA_OK(cudaMemcpyAsync(tokens_out_host, dTok, (size_t)n * c.segment * 4, cudaMemcpyDeviceToHost, e->copyStream));
  CONAN_CUDA_OK(cudaEventRecord(e->evCopy[b], e->copyStream));
  e->ticketPending[b] = true;
  ++e->submitCount;
  *ticket = b;
  return 0;
}

int conan_step_host_wait(conan_engine_t* e, int ticket) {
  if (check_ready(e)) return 1;
  if (ticket < 0 || ticket > 1 || !e->ticketPending[ticket]) { set_error("conan_step_host_wait: no such step in flight"); return 1; }
  CONAN_CUDA_OK(cudaEventSynchronize(e->evCopy[ticket]));
  e->ticketPending[ticket] = false;
  return 0;
}

int conan_engine_set_profiling(conan_engine_t* e, int enabled) {
  if (!e) { set_error("null engine"); return 1; }
  for (auto& r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  e->prof.clear();
  e->profiling = enabled != 0;
  return 0;
}

int conan_engine_profile_read(conan_engine_t* e, int category, double* ms, uint64_t* launches, double* flops, double* bytes) {
  if (!e) { set_error("null engine"); return 1; }
  CONAN_CUDA_OK(cudaDeviceSynchronize());
  double t = 0, f = 0, b = 0; uint64_t n = 0;
  for (auto& r : e->prof) {
    if (r.cat != category) continue;
    float dt = 0.f;
    CONAN_CUDA_OK(cudaEventElapsedTime(&dt, r.a, r.b));
    t += dt; f += r.flops; b += r.bytes; ++n;
  }
  if (ms) *ms = t;
  if (launches) *launches = n;
  if (flops) *flops = f;
  if (bytes) *bytes = b;
  return 0;
}

int conan_debug_read(conan_engine_t* e, const char* name, int slot, float* out_dev, size_t capacity, size_t* numel, void* stream) {
  if (check_ready(e)) return 1;
  if (slot < 0 || slot >= e->Su) { set_error("slot id out of range"); return 1; }
  const conan_config_t& c = e->cfg;
  const void* src = nullptr; size_t cnt = 0;
  std::string nm(name ? name : "");
  if (nm == "style") { src = e->sSTYLE + (size_t)slot * c.hidden_size; cnt = c.hidden_size; }
  else if (nm == "kv_cache") { cnt = (size_t)2 * e->tp_max * 2 * c.hidden_size; src = e->sKV + (size_t)slot * cnt; }
  else if (nm == "kpm") { cnt = e->tp_max; src = e->sKPM + (size_t)slot * cnt; }
  else if (nm == "emformer_past_len") { cnt = 1; src = e->ePast + slot; }
  else if (nm == "uv_pred") { cnt = (size_t)c.segment * 4; src = e->dUVP + (size_t)slot * cnt; }   // `slot` = index in the last ready list
  else if (nm == "vq_index") { cnt = e->tp_max; src = e->qVQ; }   // compact index 0 of the last session batch
  else { set_error("unknown debug tensor '" + nm + "'"); return 1; }
  if (numel) *numel = cnt;
  if (out_dev) {
    if (capacity < cnt) { set_error("debug buffer too small"); return 1; }
    CONAN_CUDA_OK(cudaMemcpyAsync(out_dev, src, cnt * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  }
  return 0;
}

int conan_logmel(const float* wav_rows, int n_streams, int rows_per_stream, int hop, int taps, int row0, int n_frames,
                 const float* dft_w, int bins, const float* mel_basis_t, int n_mels, float eps, float vmin, float vmax,
                 float* spec_scratch, float* mel_out, void* stream) {
  if (!wav_rows || !dft_w || !mel_basis_t || !spec_scratch || !mel_out) { set_error("null argument"); return 1; }
  if (n_streams <= 0 || n_frames <= 0) return 0;
  if (row0 < 0 || row0 + n_frames + taps - 1 > rows_per_stream) { set_error("logmel: frames outside the padded signal"); return 1; }
  cudaStream_t st = (cudaStream_t)stream;
  // frame f = taps consecutive hop-sample rows starting at row f of the centre-padded signal: an implicit conv with the
  // windowed DFT basis [2*bins, taps*hop] as weights (fp32 FFMA engine: the spectrum feeds a log)
  conan_conv_params_t p;
  memset(&p, 0, sizeof(p));
  p.x = wav_rows; p.x_slot_stride = (long long)rows_per_stream * hop; p.x_row_stride = hop; p.x_rows = rows_per_stream;
  p.row0 = row0; p.L = n_frames; p.cin = hop; p.k = taps; p.dil = 1; p.cout = 2 * bins; p.w = dft_w;
  p.n_streams = n_streams; p.n_slots = n_streams; p.scale = 1.f; p.out_scale = 1.f;
  p.y = spec_scratch; p.y_slot_stride = (long long)n_frames * 2 * bins; p.y_row_stride = 2 * bins;
  if (launch_conv_gemm_ffma(p, st)) return 1;
  return launch_logmel(spec_scratch, 2 * bins, bins, mel_basis_t, n_mels, eps, vmin, vmax, mel_out, (long long)n_streams * n_frames, st);
}

int conan_conv_gemm(const conan_conv_params_t* p, int engine, void* stream) {
  if (!p) { set_error("null params"); return 1; }
  if (engine == 1) {
    if (!conv_gemm_tc_eligible(*p)) { set_error("conv_gemm: shape/dtype not eligible for the tcgen05 engine"); return 1; }
    return launch_conv_gemm_tc(*p, (cudaStream_t)stream);
  }
  return launch_conv_gemm_ffma(*p, (cudaStream_t)stream);
}

}  // extern "C"
