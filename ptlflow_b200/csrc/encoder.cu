// Encoder-side kernels (first step): the convolutions of BasicEncoder /
// SmallEncoder still run in cuDNN, but everything around them is native and fused --
//   * pfb_preprocess_frames : (x - 0.5) * 2, BGR->RGB, replicate padding, NCHW -> pixel-major, one pass
//                             (raft.py:127-135, base_model.py:206-246)
//   * pfb_instance_norm_act : per-(sample, channel) statistics + normalise + ReLU (+ residual add + ReLU)
//                             (extractor.py:29-31,49-58 -- torch runs this as batch_norm_collect_statistics,
//                             batch_norm_transform_input, relu, add, relu: five passes)
//   * pfb_add_act           : relu(residual + relu(x)) for the batch-norm (folded) context encoder
#include <algorithm>

#include "common.cuh"

namespace pfb {

template <typename T>
__global__ void preprocess_frames_kernel(const T* __restrict__ img, T* __restrict__ out, int B, int H, int W, int Hp,
                                         int Wp, int pad_top, int pad_left) {
  // out: [2B][Hp][Wp][3], frame-major (all first frames, then all second frames)
  const size_t total = (size_t)2 * B * Hp * Wp;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    int x = (int)(idx % Wp);
    size_t t = idx / Wp;
    int y = (int)(t % Hp);
    int n = (int)(t / Hp);
    const int f = n / B, b = n - f * B;
    int sy = y - pad_top, sx = x - pad_left;
    sy = sy < 0 ? 0 : (sy >= H ? H - 1 : sy);
    sx = sx < 0 ? 0 : (sx >= W ? W - 1 : sx);
    const T* src = img + (((size_t)b * 2 + f) * 3) * H * W + (size_t)sy * W + sx;
    T* o = out + idx * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {  // output channel c (RGB) <- input channel 2-c (BGR)
      float v = to_f32(src[(size_t)(2 - c) * H * W]);
      o[c] = from_f32<T>((v + (-0.5f)) * 2.0f);
    }
  }
}

// ---- instance norm ---------------------------------------------------------------------------------
// stats: [B][C][2] doubles (sum, sum of squares), zeroed by the caller-side memset node.
template <typename T>
__global__ void __launch_bounds__(256) inorm_stats_kernel(const T* __restrict__ x, double* __restrict__ stats, int HW, int C,
                                                          int pix_per_block) {
  // thread -> (channel pair, pixel lane): consecutive threads read consecutive channels (coalesced)
  const int b = blockIdx.y;
  const int c2n = C / 2;                       // channel pairs
  const int lanes = blockDim.x / c2n;          // pixel lanes per block
  const int cp = threadIdx.x % c2n, pl = threadIdx.x / c2n;
  const bool active = pl < lanes;  // trailing threads (blockDim not a multiple of C/2) only take part in the barrier
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, HW);
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  const T* base = x + (size_t)b * HW * C + 2 * cp;
  for (int p = p0 + pl; active && p < p1; p += lanes) {
    const T* v = base + (size_t)p * C;
    float a = to_f32(v[0]), bb = to_f32(v[1]);
    s0 += a; q0 = fmaf(a, a, q0);
    s1 += bb; q1 = fmaf(bb, bb, q1);
  }
  extern __shared__ float red[];  // [4][blockDim.x]
  red[threadIdx.x] = s0; red[blockDim.x + threadIdx.x] = q0;
  red[2 * blockDim.x + threadIdx.x] = s1; red[3 * blockDim.x + threadIdx.x] = q1;
  __syncthreads();
  if (active && pl == 0) {
    double ds0 = 0, dq0 = 0, ds1 = 0, dq1 = 0;
    for (int l = 0; l < lanes; ++l) {
      const int t = l * c2n + cp;
      ds0 += red[t]; dq0 += red[blockDim.x + t]; ds1 += red[2 * blockDim.x + t]; dq1 += red[3 * blockDim.x + t];
    }
    double* st = stats + ((size_t)b * C + 2 * cp) * 2;
    atomicAdd(st + 0, ds0); atomicAdd(st + 1, dq0);
    atomicAdd(st + 2, ds1); atomicAdd(st + 3, dq1);
  }
}

// y = act(norm(x)) ; with residual: y = relu(residual + act(norm(x)))
template <typename T>
__global__ void inorm_apply_kernel(const T* __restrict__ x, const double* __restrict__ stats, const T* __restrict__ residual,
                                   T* __restrict__ y, int B, int HW, int C, float eps, int relu) {
  const size_t total = (size_t)B * HW * C / 2;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t e = idx * 2;
    const int c = (int)(e % C);
    const int b = (int)(e / ((size_t)HW * C));
    const double* st = stats + ((size_t)b * C + c) * 2;
    float o[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double mean = st[2 * k] / HW;
      double var = st[2 * k + 1] / HW - mean * mean;
      var = var < 0 ? 0 : var;
      const float rstd = rsqrtf((float)var + eps);
      float v = (to_f32(x[e + k]) - (float)mean) * rstd;
      if (relu) v = fmaxf(v, 0.f);
      if (residual) v = fmaxf(to_f32(residual[e + k]) + v, 0.f);
      o[k] = v;
    }
    y[e] = from_f32<T>(o[0]);
    y[e + 1] = from_f32<T>(o[1]);
  }
}

template <typename T>
__global__ void add_act_kernel(const T* __restrict__ x, const T* __restrict__ residual, T* __restrict__ y, size_t n, int relu_x) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = to_f32(x[i]);
    if (relu_x) v = fmaxf(v, 0.f);
    y[i] = from_f32<T>(fmaxf(to_f32(residual[i]) + v, 0.f));
  }
}

}  // namespace pfb

using namespace pfb;

extern "C" PFB_API int pfb_preprocess_frames(const void* images, void* out, int B, int H, int W, int Hp, int Wp, int pad_top,
                                             int pad_left, pfb_dtype dtype, pfb_stream stream) {
  PFB_CHECK_ARG(images && out, "preprocess_frames: null pointer");
  PFB_CHECK_ARG(dtype_ok(dtype) && B > 0 && H > 0 && W > 0 && Hp >= H && Wp >= W && pad_top >= 0 && pad_left >= 0 &&
                    pad_top + H <= Hp && pad_left + W <= Wp,
                "preprocess_frames: bad geometry %dx%d -> %dx%d (+%d,+%d)", H, W, Hp, Wp, pad_top, pad_left);
  cudaStream_t s = as_stream(stream);
  size_t total = (size_t)2 * B * Hp * Wp;
  unsigned blocks = (unsigned)std::min<size_t>(ceil_div_sz(total, 256), (size_t)sm_count() * 16);
  ProfScope prof(KC_MISC, s);
  PFB_DISPATCH_DTYPE(dtype, T, {
    preprocess_frames_kernel<T><<<blocks, 256, 0, s>>>((const T*)images, (T*)out, B, H, W, Hp, Wp, pad_top, pad_left);
  });
  PFB_LAUNCH_CHECK();
  return PFB_OK;
}

extern "C" PFB_API size_t pfb_instance_norm_workspace_bytes(int B, int C) { return (size_t)B * C * 2 * sizeof(double); }

extern "C" PFB_API int pfb_instance_norm_act(const void* x, void* y, const void* residual, void* workspace, int B, int H, int W,
                                             int C, float eps, int relu, pfb_dtype dtype, pfb_stream stream) {
  PFB_CHECK_ARG(x && y && workspace, "instance_norm_act: null pointer");
  PFB_CHECK_ARG(dtype_ok(dtype) && B > 0 && H > 0 && W > 0 && C > 0 && C % 2 == 0 && C <= 512, "instance_norm_act: bad shape (C=%d must be even, <= 512)", C);
  cudaStream_t s = as_stream(stream);
  const int HW = H * W;
  double* stats = reinterpret_cast<double*>(workspace);
  PFB_CUDA(cudaMemsetAsync(stats, 0, pfb_instance_norm_workspace_bytes(B, C), s));
  const int threads = 256;
  PFB_CHECK_ARG(threads % (C / 2) == 0 || (C / 2) <= threads, "instance_norm_act: unsupported C=%d", C);
  int slabs = ceil_div(4 * sm_count(), B);
  if (slabs > ceil_div(HW, 64)) slabs = ceil_div(HW, 64);
  if (slabs < 1) slabs = 1;
  const int ppb = ceil_div(HW, slabs);
  dim3 grid(ceil_div(HW, ppb), B);
  const size_t total2 = (size_t)B * HW * C / 2;
  unsigned blocks = (unsigned)std::min<size_t>(ceil_div_sz(total2, 256), (size_t)sm_count() * 16);
  ProfScope prof(KC_MISC, s);
  PFB_DISPATCH_DTYPE(dtype, T, {
    inorm_stats_kernel<T><<<grid, threads, 4 * threads * sizeof(float), s>>>((const T*)x, stats, HW, C, ppb);
    inorm_apply_kernel<T><<<blocks, 256, 0, s>>>((const T*)x, stats, (const T*)residual, (T*)y, B, HW, C, eps, relu);
  });
  PFB_LAUNCH_CHECK();
  return PFB_OK;
}

extern "C" PFB_API int pfb_add_act(const void* x, const void* residual, void* y, size_t n, int relu_x, pfb_dtype dtype,
                                   pfb_stream stream) {
  PFB_CHECK_ARG(x && residual && y && n > 0 && dtype_ok(dtype), "add_act: bad arguments");
  cudaStream_t s = as_stream(stream);
  unsigned blocks = (unsigned)std::min<size_t>(ceil_div_sz(n, 256), (size_t)sm_count() * 16);
  ProfScope prof(KC_MISC, s);
  PFB_DISPATCH_DTYPE(dtype, T, { add_act_kernel<T><<<blocks, 256, 0, s>>>((const T*)x, (const T*)residual, (T*)y, n, relu_x); });
  PFB_LAUNCH_CHECK();
  return PFB_OK;
}
