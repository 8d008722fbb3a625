// Host-side orchestration of the refinement loop: lookup -> motion encoder -> GRU -> heads,
// `iters` times, then the upsample.  Mirrors the data flow of
//   ptlflow/models/raft/raft.py:170-192 and ptlflow/models/raft/update.py:115-153
// but every stage is one launch of this library's kernels on the caller's stream (so the whole
// loop can be captured in a CUDA graph by the host), no torch.cat / permute copies exist, and
// the mask head + upsample run once (eval returns only the last prediction, raft.py:192).
#include <initializer_list>

#include <stdlib.h>

#include "common.cuh"

#define PFB_TRY(expr)        \
  do {                       \
    int rc__ = (expr);       \
    if (rc__ != PFB_OK) return rc__; \
  } while (0)

namespace pfb {

int launch_flow_from_coords(const float* coords, float* flow, int B, int H, int W, cudaStream_t s);

struct Workspace {
  // element strides (channels per pixel) and byte offsets inside the caller's workspace
  int planes, corr_stride;
  size_t off_corr, off_cor1, off_corflo, off_flo1, off_motion, off_z, off_rh, off_fh, off_mh, off_mask, off_flow;
  size_t off_taps;          // flow head conv2 per-tap products [P][32] fp32 (tensor-core path)
  size_t off_vbuf, off_vT;  // gma: to_v(motion) [P][128] and its per-sample transpose [B][128][n_pad]
  size_t off_flags;         // on-the-fly tensor-core lookup: one flag per query (queries recomputed by the SIMT pass)
  size_t off_ctx[4];        // iteration-invariant context terms of the GRU gates: zr1 [P][2hd], q1 [P][hd], zr2, q2 (tensor path)
  int n_pad;
  size_t total;
  int c_cor1, c_corflo, c_cor2, c_flo1, c_flo2, c_motion, c_fh;
};

static Workspace plan(const pfb_raft_cfg* c) {
  Workspace w{};
  const size_t P = (size_t)c->B * c->H * c->W;
  const size_t es = dtype_size(c->dtype);
  const int K = 2 * c->corr_radius + 1;
  w.planes = c->corr_levels * K * K;
  // 16-byte aligned rows for the tensor-core path (the TMA unit zero-fills a partial last 64-channel K chunk itself)
  w.corr_stride = (c->dtype == PFB_F32) ? w.planes : (int)align_up(w.planes, 8);
  if (c->variant == 0 || c->variant == 2) {
    w.c_cor1 = 256; w.c_cor2 = 192; w.c_flo1 = 128; w.c_flo2 = 64; w.c_fh = 256;
    // gma keeps [motion | motion_global] side by side so the GRU still sees three sources (update.py:150-151)
    w.c_motion = c->variant == 2 ? 256 : 128;
  } else {
    w.c_cor1 = 0; w.c_cor2 = 96; w.c_flo1 = 64; w.c_flo2 = 32; w.c_motion = 82; w.c_fh = 128;
  }
  w.c_corflo = w.c_cor2 + w.c_flo2;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  w.off_corr = take(P * w.corr_stride * es);
  w.off_cor1 = take(P * (size_t)w.c_cor1 * es);
  w.off_corflo = take(P * (size_t)w.c_corflo * es);
  w.off_flo1 = take(P * (size_t)w.c_flo1 * es);
  w.off_motion = take(P * (size_t)w.c_motion * es);
  w.off_z = take(P * (size_t)c->hidden_dim * es);
  w.off_rh = take(P * (size_t)c->hidden_dim * es);
  w.off_fh = take(P * (size_t)w.c_fh * es);
  w.off_mh = take(c->variant != 1 ? P * 256 * es : 0);
  w.off_mask = take(c->variant != 1 ? P * 576 * es : 0);
  w.off_flow = take(P * 2 * sizeof(float));
  w.off_taps = take(c->variant != 1 ? P * 32 * sizeof(float) : 0);
  w.n_pad = (int)align_up((size_t)c->H * c->W, 64);
  w.off_vbuf = take(c->variant == 2 ? P * 128 * es : 0);
  w.off_vT = take(c->variant == 2 ? (size_t)c->B * 128 * w.n_pad * es : 0);
  w.off_flags = take(c->alternate_corr ? P : 0);
  for (int i = 0; i < 4; ++i)
    w.off_ctx[i] = take((c->variant != 1 && c->dtype != PFB_F32) ? P * (size_t)((i & 1) ? c->hidden_dim : 2 * c->hidden_dim) * es : 0);
  w.total = off;
  return w;
}

static int check_cfg(const pfb_raft_cfg* c) {
  PFB_CHECK_ARG(c, "raft: null cfg");
  PFB_CHECK_ARG(c->variant >= 0 && c->variant <= 2, "raft: variant=%d", c->variant);
  PFB_CHECK_ARG(dtype_ok(c->dtype), "raft: bad dtype");
  PFB_CHECK_ARG(c->B > 0 && c->H > 0 && c->W > 0, "raft: bad grid %dx%dx%d", c->B, c->H, c->W);
  PFB_CHECK_ARG(c->corr_levels >= 1 && c->corr_levels <= PFB_MAX_LEVELS && c->corr_radius >= 0 && c->corr_radius <= 15,
                "raft: corr_levels=%d corr_radius=%d", c->corr_levels, c->corr_radius);
  PFB_CHECK_ARG((c->H >> (c->corr_levels - 1)) >= 1 && (c->W >> (c->corr_levels - 1)) >= 1,
                "raft: %dx%d grid too small for %d levels", c->H, c->W, c->corr_levels);
  PFB_CHECK_ARG(c->hidden_dim > 0 && c->context_dim > 0 && c->iters >= 0, "raft: bad dims");
  PFB_CHECK_ARG(c->volume_layout == 0 || (c->volume_layout == 1 && c->dtype != PFB_F32 && !c->alternate_corr && c->corr_levels <= 4),
                "raft: volume_layout=%d needs f16/bf16, a materialised pyramid and <= 4 levels", c->volume_layout);
  if (c->variant != 1) PFB_CHECK_ARG(c->hidden_dim == 128 && c->context_dim == 128, "raft/gma: the update block expects hidden=context=128");
  else PFB_CHECK_ARG(c->hidden_dim == 96 && c->context_dim == 64, "raft_small: SmallUpdateBlock expects hidden=96 context=64");
  return PFB_OK;
}

struct Ctx {
  const pfb_raft_cfg* c;
  const pfb_raft_weights* w;
  const pfb_raft_buffers* b;
  Workspace ws;
  char* base;
  cudaStream_t s;
  // flow branch of the motion encoder on a second stream (fork_flow_branch): null when not forked
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  void* at(size_t off) const { return base + off; }
};

static pfb_conv_src src_of(const void* ptr, int channels, int stride, int offset = 0, int is_f32 = 0) {
  pfb_conv_src s;
  s.ptr = ptr; s.channels = channels; s.stride = stride; s.offset = offset; s.is_f32 = is_f32;
  return s;
}

static int run_conv(const Ctx& x, int layer, std::initializer_list<pfb_conv_src> srcs, int epi, void* out,
                    int out_stride, int out_offset, float scale = 1.f, const void* addend = nullptr, int addend_stride = 0) {
  const pfb_layer& L = x.w->layers[layer];
  PFB_CHECK_ARG(L.weight, "raft: layer %d has no packed weight", layer);
  pfb_conv_params p{};
  int i = 0, cin = 0;
  for (const auto& s : srcs) { p.src[i++] = s; cin += s.channels; }
  p.nsrc = i;
  // the lookup buffer may be wider than the layer's real Cin (zero pad columns <-> zero weight rows)
  PFB_CHECK_ARG(cin == L.Cin, "raft: layer %d expects Cin=%d, sources provide %d", layer, L.Cin, cin);
  p.B = x.c->B; p.H = x.c->H; p.W = x.c->W;
  p.KH = L.KH; p.KW = L.KW; p.Cout = L.Cout; p.Cout_pad = L.Cout_pad;
  p.weight = L.weight; p.bias = L.bias;
  p.epilogue = epi; p.scale = scale;
  p.out = out; p.out_stride = out_stride; p.out_offset = out_offset;
  p.aux_h = x.b->net; p.aux_z = x.at(x.ws.off_z); p.hidden = x.c->hidden_dim;
  p.coords = x.b->coords; p.flow = reinterpret_cast<const float*>(x.at(x.ws.off_flow));
  p.dtype = x.c->dtype; p.impl = x.c->impl;
  p.weight_k = L.weight_k; p.Cin_pad = L.Cin_pad; p.Cout_pad_k = L.Cout_pad_k;
  p.addend = addend; p.addend_stride = addend_stride;
  return pfb_conv2d(&p, (pfb_stream)x.s);
}

// round-2 restructurings of the tensor-core path; PFB_GRU_CTX_SPLIT=0 / PFB_MERGE_C2F2=0 fall back to the plain layers
static bool tensor_path(const Ctx& x) { return x.c->variant != 1 && x.c->dtype != PFB_F32 && x.c->impl != 1; }
static bool ctx_split_active(const Ctx& x) {
  static const int env = getenv("PFB_GRU_CTX_SPLIT") ? atoi(getenv("PFB_GRU_CTX_SPLIT")) : 1;
  return env && tensor_path(x) && x.w->layers[PFB_L_GRUX_ZR1].weight_k && x.w->layers[PFB_L_CTX_ZR1].weight_k;
}
static bool merged_c2f2_active(const Ctx& x) {
  // measured (launch list r02b): 70.5 us merged vs 46.3 + 23.8 us separate -- the zero blocks cost what the better shape saves,
  // so the merged layer is opt-in
  static const int env = getenv("PFB_MERGE_C2F2") ? atoi(getenv("PFB_MERGE_C2F2")) : 0;
  return env && tensor_path(x) && x.w->layers[PFB_L_CONVC2F2].weight_k;
}
// The motion encoder has two independent branches: lookup -> convc1 -> convc2 (correlation) and convf1 -> convf2 (flow); both
// start from the previous iteration's coordinates and meet in `conv`.  cfg.fork_flow puts the flow branch on a second stream
// of this host thread (fork / join with events, which a CUDA-graph capture turns into parallel branches of the graph), so that
// the SIMT lookup can share the SMs with the two small tensor-core layers (+0.3 ... 2 % per step).
static bool fork_flow_active(const Ctx& x) { return x.c->fork_flow && tensor_path(x) && !merged_c2f2_active(x); }
static int fork_flow_setup(Ctx& x) {
  if (!fork_flow_active(x)) return PFB_OK;
  // one side stream + event pair per (host thread, device): forwards of different host threads run concurrently
  struct Lane {
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  };
  constexpr int kMaxDev = 64;
  thread_local Lane lanes[kMaxDev];
  int dev = 0;
  PFB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev) return PFB_OK;  // no fork on exotic device numbers: everything stays on the caller's stream
  Lane& l = lanes[dev];
  if (!l.side) {
    PFB_CUDA(cudaStreamCreateWithFlags(&l.side, cudaStreamNonBlocking));
    PFB_CUDA(cudaEventCreateWithFlags(&l.ev_fork, cudaEventDisableTiming));
    PFB_CUDA(cudaEventCreateWithFlags(&l.ev_join, cudaEventDisableTiming));
  }
  x.side = l.side; x.ev_fork = l.ev_fork; x.ev_join = l.ev_join;
  return PFB_OK;
}

// conv_inp(inp) + bias for the four GRU convolutions: once per forward (inp does not change over the iterations)
static int run_context_terms(const Ctx& x) {
  if (!ctx_split_active(x)) return PFB_OK;
  const int hd = x.c->hidden_dim, cd = x.c->context_dim;
  const int layers[4] = {PFB_L_CTX_ZR1, PFB_L_CTX_Q1, PFB_L_CTX_ZR2, PFB_L_CTX_Q2};
  for (int i = 0; i < 4; ++i) {
    const int n = (i & 1) ? hd : 2 * hd;
    PFB_TRY(run_conv(x, layers[i], {src_of(x.b->inp, cd, cd)}, PFB_EPI_LINEAR, x.at(x.ws.off_ctx[i]), n, 0));
  }
  return PFB_OK;
}

static int lookup(const Ctx& x) {
  const pfb_raft_cfg* c = x.c;
  if (c->alternate_corr) {
    static const int env_tc = getenv("PFB_ONTHEFLY_TC") ? atoi(getenv("PFB_ONTHEFLY_TC")) : 1;
    if (env_tc && c->impl != 1 && corr_onthefly_umma_supported(c->B, c->H, c->W, c->feat_dim, c->corr_levels, c->corr_radius, c->dtype, x.ws.corr_stride))
      return pfb_corr_lookup_onthefly_tc(x.b->fmap1, x.b->pyramid, x.b->coords, x.at(x.ws.off_corr), x.at(x.ws.off_flags), c->B, c->H, c->W,
                                         c->feat_dim, c->corr_levels, c->corr_radius, c->dtype, x.ws.corr_stride, (pfb_stream)x.s);
  }
  if (c->alternate_corr)
    return pfb_corr_lookup_onthefly(x.b->fmap1, x.b->pyramid, x.b->coords, x.at(x.ws.off_corr), c->B, c->H, c->W,
                                    c->feat_dim, c->corr_levels, c->corr_radius, c->dtype, c->dtype, 0,
                                    x.ws.corr_stride, (pfb_stream)x.s);
  if (c->volume_layout == 1)
    return pfb_corr_lookup_tiled(x.b->pyramid, x.b->coords, x.at(x.ws.off_corr), c->B, c->H, c->W, c->H, c->W, c->corr_levels,
                                 c->corr_radius, c->dtype, x.ws.corr_stride, (pfb_stream)x.s);
  return pfb_corr_lookup(x.b->pyramid, x.b->coords, x.at(x.ws.off_corr), c->B, c->H, c->W, c->corr_levels,
                         c->corr_radius, c->dtype, c->dtype, 0, x.ws.corr_stride, (pfb_stream)x.s);
}

// One BasicUpdateBlock / SmallUpdateBlock evaluation + coords update.  update.py:122-153
static int update_iter(const Ctx& x, const void* corr_ext, void* mask_out) {
  const pfb_raft_cfg* c = x.c;
  const Workspace& ws = x.ws;
  const int hd = c->hidden_dim, cd = c->context_dim;
  const void* corr = corr_ext ? corr_ext : x.at(ws.off_corr);
  const int corr_stride = corr_ext ? ws.planes : ws.corr_stride;
  void* corflo = x.at(ws.off_corflo);
  void* flo1 = x.at(ws.off_flo1);
  void* motion = x.at(ws.off_motion);
  void* rh = x.at(ws.off_rh);
  void* fh = x.at(ws.off_fh);
  float* flow = reinterpret_cast<float*>(x.at(ws.off_flow));

  // ---- motion encoder (update.py:76-112) ----
  const bool merged_c2f2 = merged_c2f2_active(x);
  Ctx xf = x;  // the flow branch's launches: the side stream when forked (the fork itself is in the caller, before the lookup)
  if (x.side) xf.s = x.side;
  if (c->variant != 1) {
    void* cor1 = x.at(ws.off_cor1);
    PFB_TRY(run_conv(x, PFB_L_CONVC1, {src_of(corr, ws.planes, corr_stride)}, PFB_EPI_RELU, cor1, ws.c_cor1, 0));
    if (!merged_c2f2) PFB_TRY(run_conv(x, PFB_L_CONVC2, {src_of(cor1, ws.c_cor1, ws.c_cor1)}, PFB_EPI_RELU, corflo, ws.c_corflo, 0));
  } else {
    PFB_TRY(run_conv(x, PFB_L_CONVC1, {src_of(corr, ws.planes, corr_stride)}, PFB_EPI_RELU, corflo, ws.c_corflo, 0));
  }
  {
    const pfb_layer& LF = x.w->layers[PFB_L_CONVF1];
    static const int env_fc = getenv("PFB_FLOW_CONV_UMMA") ? atoi(getenv("PFB_FLOW_CONV_UMMA")) : 1;
    if (env_fc && LF.weight_k && LF.KH == 7 && LF.KW == 7 && LF.Cin == 2 && LF.Cout == 128 && c->dtype != PFB_F32 && c->impl != 1)
      PFB_TRY(pfb_flow_conv7x7(flow, LF.weight_k, LF.bias, flo1, ws.c_flo1, 0, c->B, c->H, c->W, c->dtype, (pfb_stream)xf.s));
    else
      PFB_TRY(run_conv(xf, PFB_L_CONVF1, {src_of(flow, 2, 2, 0, 1)}, PFB_EPI_RELU, flo1, ws.c_flo1, 0));
  }
  if (merged_c2f2)  // convc2 | convf2 as one block-diagonal layer: [cor1 (256) | flo1 (128)] -> [cor (192) | flo (64)]
    PFB_TRY(run_conv(x, PFB_L_CONVC2F2, {src_of(x.at(ws.off_cor1), ws.c_cor1, ws.c_cor1), src_of(flo1, ws.c_flo1, ws.c_flo1)}, PFB_EPI_RELU,
                     corflo, ws.c_corflo, 0));
  else
    PFB_TRY(run_conv(xf, PFB_L_CONVF2, {src_of(flo1, ws.c_flo1, ws.c_flo1)}, PFB_EPI_RELU, corflo, ws.c_corflo, ws.c_cor2));
  if (x.side) {  // join: `conv` reads both branches
    PFB_CUDA(cudaEventRecord(x.ev_join, x.side));
    PFB_CUDA(cudaStreamWaitEvent(x.s, x.ev_join, 0));
  }
  PFB_TRY(run_conv(x, PFB_L_CONV, {src_of(corflo, ws.c_corflo, ws.c_corflo)}, PFB_EPI_RELU_APPEND_FLOW, motion, ws.c_motion, 0));

  // ---- gma: motion_global = motion + gamma * (attention @ to_v(motion))   gma_utils.py:101-113, gma/update.py:149 ----
  if (c->variant == 2) {
    PFB_CHECK_ARG(x.b->attention, "gma: null attention");
    const int N = c->H * c->W;
    const size_t es = dtype_size(c->dtype);
    char* vbuf = reinterpret_cast<char*>(x.at(ws.off_vbuf));
    char* vT = reinterpret_cast<char*>(x.at(ws.off_vT));
    PFB_TRY(run_conv(x, PFB_L_AGG_V, {src_of(motion, 128, ws.c_motion)}, PFB_EPI_LINEAR, vbuf, 128, 0));
    const bool tensor_path = c->dtype != PFB_F32 && c->impl != 1 && (N % 8) == 0;
    if (tensor_path) PFB_TRY(pfb_transpose_pm(vbuf, vT, c->B, N, 128, ws.n_pad, c->dtype, (pfb_stream)x.s));
    static const int env_batched = getenv("PFB_GMA_BATCHED") ? atoi(getenv("PFB_GMA_BATCHED")) : 1;
    if (tensor_path && env_batched) {
      // ONE launch for all samples: pixels = queries, input channels = the N attention columns, per-sample weights = that
      // sample's v^T (K-major [128][n_pad], rows b * 128 ...).  B x 55 row tiles fill the machine; one launch per sample left
      // 55 of 148 SMs busy.
      pfb_conv_params p{};
      p.src[0] = src_of(x.b->attention, N, N);
      p.nsrc = 1;
      p.B = c->B; p.H = c->H; p.W = c->W; p.KH = 1; p.KW = 1;
      p.Cout = 128; p.Cout_pad = 128;
      p.weight = vbuf;  // (SIMT layout, unused on this path)
      p.bias = nullptr;
      p.epilogue = PFB_EPI_AXPY; p.scale = x.b->agg_gamma;
      p.out = motion; p.out_stride = ws.c_motion; p.out_offset = 128;
      p.aux_h = motion; p.hidden = ws.c_motion;
      p.dtype = c->dtype; p.impl = 2;
      p.weight_k = vT; p.Cin_pad = ws.n_pad; p.Cout_pad_k = 128;
      p.w_rows_per_sample = 128;
      PFB_TRY(pfb_conv2d(&p, (pfb_stream)x.s));
    } else
    for (int b = 0; b < c->B; ++b) {
      // one "1x1 convolution" per sample: pixels = queries, input channels = the N attention columns,
      // weights = this sample's v (SIMT layout [N][128]) / v^T (K-major [128][n_pad])
      pfb_conv_params p{};
      p.src[0] = src_of(reinterpret_cast<const char*>(x.b->attention) + (size_t)b * N * N * es, N, N);
      p.nsrc = 1;
      p.B = 1; p.H = c->H; p.W = c->W; p.KH = 1; p.KW = 1;
      p.Cout = 128; p.Cout_pad = 128;
      p.weight = vbuf + (size_t)b * N * 128 * es;
      p.bias = nullptr;
      p.epilogue = PFB_EPI_AXPY; p.scale = x.b->agg_gamma;
      char* mrow = reinterpret_cast<char*>(motion) + (size_t)b * N * ws.c_motion * es;
      p.out = mrow; p.out_stride = ws.c_motion; p.out_offset = 128;
      p.aux_h = mrow; p.hidden = ws.c_motion;
      p.dtype = c->dtype; p.impl = c->impl;
      if (tensor_path) { p.weight_k = vT + (size_t)b * 128 * ws.n_pad * es; p.Cin_pad = ws.n_pad; p.Cout_pad_k = 128; }
      PFB_TRY(pfb_conv2d(&p, (pfb_stream)x.s));
    }
  }

  // ---- GRU (update.py:24-32 ConvGRU, :58-73 SepConvGRU); x = [inp, motion (, motion_global)] ----
  const int halves = (c->variant != 1) ? 2 : 1;
  const bool split = ctx_split_active(x);
  for (int h = 0; h < halves; ++h) {
    if (split) {
      // conv([h | inp | motion]) = conv_inp(inp) + bias (once per forward, run_context_terms) + conv_rest([h | motion]) (here)
      const int lzr = h == 0 ? PFB_L_GRUX_ZR1 : PFB_L_GRUX_ZR2, lq = h == 0 ? PFB_L_GRUX_Q1 : PFB_L_GRUX_Q2;
      PFB_TRY(run_conv(x, lzr, {src_of(x.b->net, hd, hd), src_of(motion, ws.c_motion, ws.c_motion)}, PFB_EPI_GRU_ZR, rh, hd, 0, 1.f,
                       x.at(ws.off_ctx[2 * h]), 2 * hd));
      PFB_TRY(run_conv(x, lq, {src_of(rh, hd, hd), src_of(motion, ws.c_motion, ws.c_motion)}, PFB_EPI_GRU_Q, x.b->net, hd, 0, 1.f,
                       x.at(ws.off_ctx[2 * h + 1]), hd));
      continue;
    }
    const int lzr = h == 0 ? PFB_L_GRU_ZR1 : PFB_L_GRU_ZR2, lq = h == 0 ? PFB_L_GRU_Q1 : PFB_L_GRU_Q2;
    PFB_TRY(run_conv(x, lzr, {src_of(x.b->net, hd, hd), src_of(x.b->inp, cd, cd), src_of(motion, ws.c_motion, ws.c_motion)},
                     PFB_EPI_GRU_ZR, rh, hd, 0));
    PFB_TRY(run_conv(x, lq, {src_of(rh, hd, hd), src_of(x.b->inp, cd, cd), src_of(motion, ws.c_motion, ws.c_motion)},
                     PFB_EPI_GRU_Q, x.b->net, hd, 0));
  }

  // ---- heads (update.py:6-14, :138-152) ----
  const bool fork_mask = x.side && mask_out && c->variant != 1;
  if (fork_mask) {  // last iteration: the mask head beside the flow head (both read the final hidden state)
    PFB_CUDA(cudaEventRecord(x.ev_fork, x.s));
    PFB_CUDA(cudaStreamWaitEvent(x.side, x.ev_fork, 0));
    void* mh = x.at(ws.off_mh);
    PFB_TRY(run_conv(xf, PFB_L_MASK1, {src_of(x.b->net, hd, hd)}, PFB_EPI_RELU, mh, 256, 0));
    PFB_TRY(run_conv(xf, PFB_L_MASK2, {src_of(mh, 256, 256)}, PFB_EPI_LINEAR, mask_out, 576, 0, 0.25f));
    PFB_CUDA(cudaEventRecord(x.ev_join, x.side));
  }
  PFB_TRY(run_conv(x, PFB_L_FLOW1, {src_of(x.b->net, hd, hd)}, PFB_EPI_RELU, fh, ws.c_fh, 0));
  const pfb_layer& LT = x.w->layers[PFB_L_FLOW2T];
  if (c->variant != 1 && c->dtype != PFB_F32 && c->impl != 1 && LT.weight_k) {
    // tensor-core form of the 3x3 -> 2 convolution: one 1x1 GEMM to the 18 (tap, output) products, then a 9-tap gather
    float* taps = reinterpret_cast<float*>(x.at(ws.off_taps));
    PFB_TRY(run_conv(x, PFB_L_FLOW2T, {src_of(fh, ws.c_fh, ws.c_fh)}, PFB_EPI_LINEAR_F32, taps, 32, 0));
    PFB_TRY(pfb_flow_tap_gather(taps, 32, x.w->layers[PFB_L_FLOW2].bias, x.b->coords, flow, c->B, c->H, c->W, (pfb_stream)x.s));
  } else {
    PFB_TRY(run_conv(x, PFB_L_FLOW2, {src_of(fh, ws.c_fh, ws.c_fh)}, PFB_EPI_FLOW, flow, 2, 0));
  }
  if (fork_mask) {
    PFB_CUDA(cudaStreamWaitEvent(x.s, x.ev_join, 0));  // the upsample reads the mask
  } else if (mask_out && c->variant != 1) {
    void* mh = x.at(ws.off_mh);
    PFB_TRY(run_conv(x, PFB_L_MASK1, {src_of(x.b->net, hd, hd)}, PFB_EPI_RELU, mh, 256, 0));
    PFB_TRY(run_conv(x, PFB_L_MASK2, {src_of(mh, 256, 256)}, PFB_EPI_LINEAR, mask_out, 576, 0, 0.25f));
  }
  return PFB_OK;
}

static int make_ctx(Ctx& x, const pfb_raft_cfg* cfg, const pfb_raft_weights* w, const pfb_raft_buffers* buf,
                    cudaStream_t s, bool need_pyramid) {
  PFB_TRY(check_cfg(cfg));
  PFB_CHECK_ARG(w && buf, "raft: null weights/buffers");
  PFB_CHECK_ARG(buf->net && buf->inp && buf->coords && buf->workspace, "raft: null state buffer");
  if (need_pyramid) {
    PFB_CHECK_ARG(buf->pyramid, "raft: null pyramid");
    PFB_CHECK_ARG(!cfg->alternate_corr || (buf->fmap1 && cfg->feat_dim > 0), "raft: alternate_corr needs fmap1 and feat_dim");
  }
  x.c = cfg; x.w = w; x.b = buf; x.s = s;
  x.ws = plan(cfg);
  PFB_CHECK_ARG(buf->workspace_bytes >= x.ws.total, "raft: workspace %zu bytes < required %zu", buf->workspace_bytes, x.ws.total);
  x.base = reinterpret_cast<char*>(buf->workspace);
  return PFB_OK;
}

}  // namespace pfb

using namespace pfb;

extern "C" PFB_API size_t pfb_raft_workspace_bytes(const pfb_raft_cfg* cfg) {
  if (check_cfg(cfg) != PFB_OK) return 0;
  return plan(cfg).total;
}

extern "C" PFB_API int pfb_raft_update_iter(const pfb_raft_cfg* cfg, const pfb_raft_weights* w, const pfb_raft_buffers* buf,
                                    const void* corr, void* mask_out, pfb_stream stream) {
  Ctx x;
  PFB_TRY(make_ctx(x, cfg, w, buf, as_stream(stream), corr == nullptr));
  PFB_TRY(launch_flow_from_coords(buf->coords, reinterpret_cast<float*>(x.at(x.ws.off_flow)), cfg->B, cfg->H, cfg->W, x.s));
  PFB_TRY(run_context_terms(x));
  if (!corr) PFB_TRY(lookup(x));
  return update_iter(x, corr, mask_out);
}

extern "C" PFB_API int pfb_raft_refine(const pfb_raft_cfg* cfg, const pfb_raft_weights* w, const pfb_raft_buffers* buf,
                               pfb_stream stream) {
  Ctx x;
  PFB_TRY(make_ctx(x, cfg, w, buf, as_stream(stream), true));
  PFB_CHECK_ARG(buf->flow_up, "raft_refine: null flow_up");
  PFB_CHECK_ARG(cfg->variant == 1 || cfg->iters >= 1, "raft_refine: the convex upsample needs at least one iteration (mask)");
  PFB_TRY(launch_flow_from_coords(buf->coords, reinterpret_cast<float*>(x.at(x.ws.off_flow)), cfg->B, cfg->H, cfg->W, x.s));
  void* mask = cfg->variant != 1 ? x.at(x.ws.off_mask) : nullptr;
  PFB_TRY(fork_flow_setup(x));
  if (x.side) {
    // the once-per-forward context terms ride the side stream too: they are first read by the GRU of iteration 0, after that
    // iteration's join (same stream as its flow branch, so the join covers them)
    PFB_CUDA(cudaEventRecord(x.ev_fork, x.s));
    PFB_CUDA(cudaStreamWaitEvent(x.side, x.ev_fork, 0));
    Ctx xs = x;
    xs.s = x.side;
    PFB_TRY(run_context_terms(xs));
  } else {
    PFB_TRY(run_context_terms(x));
  }
  for (int it = 0; it < cfg->iters; ++it) {
    if (x.side) {  // fork: the flow branch may start as soon as the previous iteration's coordinates are written
      PFB_CUDA(cudaEventRecord(x.ev_fork, x.s));
      PFB_CUDA(cudaStreamWaitEvent(x.side, x.ev_fork, 0));
    }
    PFB_TRY(lookup(x));
    PFB_TRY(update_iter(x, nullptr, it == cfg->iters - 1 ? mask : nullptr));
  }
  if (cfg->variant != 1)
    return pfb_convex_upsample(buf->coords, mask, buf->flow_up, buf->flow_small, cfg->B, cfg->H, cfg->W, cfg->out_h,
                               cfg->out_w, cfg->pad_top, cfg->pad_left, cfg->dtype, stream);
  return pfb_upflow8(buf->coords, buf->flow_up, buf->flow_small, cfg->B, cfg->H, cfg->W, cfg->out_h, cfg->out_w,
                     cfg->pad_top, cfg->pad_left, stream);
}
