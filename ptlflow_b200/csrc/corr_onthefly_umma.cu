// a4 on the tensor cores: on-the-fly correlation + lookup without the 4D volume (f16 / bf16, radius 4, C % 64 == 0, C <= 256).
//
// Replaces, for those shapes, the reference's alt_cuda_corr kernel (ptlflow/utils/external/alt_cuda_corr/
// correlation_kernel.cu:18-119: one warp per 4x8 query tile, 32-channel chunks, scalar FMAs, global RMW per tap) and this
// library's own SIMT kernel (csrc/corr.cu: one warp per query, a 256-long dot product per window tap from L2).
//
// Work item = (tile of 8 x 16 neighbouring queries, pyramid level).  The windows of neighbouring queries overlap almost
// completely when the flow is locally smooth, so the item multiplies the tile's 128 query vectors with a REGION of the
// level's feature map that contains all their windows -- anchored at the tile's smallest window origin, 32 targets wide, in
// bands of 8 rows (band stride 7, so that every vertical tap pair lies inside one band):
//     D[128 queries][256 targets] = F1_tile[128][C] . F2_band[256][C]^T        (tcgen05, M = 128, N = 256, fp32 in TMEM)
// Both operands are TMA boxes of the pixel-major feature maps (out-of-map targets are zero-filled by the TMA unit: the
// zero padding of raft/utils.py:71-75 for free).  Each epilogue thread owns one query: it dumps its accumulator row to
// shared memory (scaled, storage type -- the same rounding the materialised volume has), then blends the window rows that
// fall into this band (x-major order of corr.py:43-47) and stages the level's 81 outputs for a coalesced store.
// Queries whose window does not fit the region (rough flow inside a tile) are flagged and recomputed by the SIMT kernel
// (exact same values, one warp per flagged query), so the result never depends on the smoothness of the flow.
#include <stdlib.h>

#include <algorithm>
#include <vector>
#include <stdio.h>

#include "umma.cuh"

namespace pfb {
using namespace sm100;

constexpr int kOtfTileBytes = 128 * 128;       // 128 rows x 64 channels
constexpr int kOtfBandRows = 256;              // targets per band: 8 rows x 32 columns
constexpr int kOtfMaxBands = 8;                // 7 * 8 + 1 = 57 region rows at most
constexpr int kOtfRW = 32;
constexpr int kOtfDumpPitch = 528;             // bytes per accumulator-dump row (256 targets x 2 bytes + 16)

struct OtfArgs {
  const float* coords;
  void* out;
  unsigned char* flags;  // [B*H*W]: set to 1 for queries the region could not serve (recomputed by the SIMT kernel)
  int B, H, W, kchunks, levels, out_stride;
  int lh[4], lw[4];
  float scale;
  int ab_fmt;
  int tiles_x, tiles_y, n_tiles;
  unsigned long long* trace;  // PFB_OTF_TRACE: [CTA][64] clock64 stamps of the CTA's second work item (phase timeline), else null
};

struct __align__(8) OtfBars {
  uint64_t a_full, b_full, acc_full;
  uint32_t tmem_base;
  int bx0, by0, nb, any;
  int red[4][3];
};

template <typename T>
__device__ __forceinline__ uint32_t otf_pack2(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t otf_pack2<__half>(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <>
__device__ __forceinline__ uint32_t otf_pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <typename T>
__device__ __forceinline__ float otf_bits_to_f32(unsigned short b);
template <>
__device__ __forceinline__ float otf_bits_to_f32<__half>(unsigned short b) { return __half2float(__ushort_as_half(b)); }
template <>
__device__ __forceinline__ float otf_bits_to_f32<__nv_bfloat16>(unsigned short b) { return __uint_as_float((unsigned)b << 16); }
template <typename T>
__device__ __forceinline__ unsigned short otf_f32_to_bits(float v);
template <>
__device__ __forceinline__ unsigned short otf_f32_to_bits<__half>(float v) { return __half_as_ushort(__float2half_rn(v)); }
template <>
__device__ __forceinline__ unsigned short otf_f32_to_bits<__nv_bfloat16>(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }

template <typename T, int R>
__global__ void __launch_bounds__(192, 1)
corr_onthefly_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                          const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2,
                          const __grid_constant__ CUtensorMap tmB3, const OtfArgs a) {
  constexpr int D = 2 * R + 2, K = 2 * R + 1, KK = K * K;
  constexpr int SP = KK + 1;  // staging row pitch in halfs (odd number of 32-bit words: conflict-free per-thread rows)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                       // kchunks x 16 KB
  uint8_t* sB = sA + a.kchunks * kOtfTileBytes;             // kchunks x 32 KB (>= 64 KB); after the MMAs: the accumulator dump [128][256]
  const int b_bytes = a.kchunks * 2 * kOtfTileBytes < 128 * kOtfDumpPitch ? 5 * kOtfTileBytes : a.kchunks * 2 * kOtfTileBytes;
  unsigned short* sOut = reinterpret_cast<unsigned short*>(sB + b_bytes);  // [128][SP] staged outputs of one level
  OtfBars* bars = reinterpret_cast<OtfBars*>(reinterpret_cast<uint8_t*>(sOut) + ((128 * SP * 2 + 15) & ~15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars->a_full, 1);
    mbar_init(&bars->b_full, 1);
    mbar_init(&bars->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<256>(&bars->tmem_base);
  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int q_local = threadIdx.x;  // epilogue threads: 0..127 = (row in tile) * 16 + (column in tile)
  uint32_t par_a = 0, par_b = 0, par_acc = 0;

  // work item = (query tile, level): 4x more, 4x shorter items than whole tiles -- 255 tiles on 148 SMs are 2 rounds with the
  // second one 72 % full, 1020 items are 6.9 rounds
  int item_no = 0;
#define OTF_TR(slot) do { if (a.trace && item_no == 1 && (slot) < 64) a.trace[blockIdx.x * 64 + (slot)] = clock64(); } while (0)
  for (int item = blockIdx.x; item < a.n_tiles * a.levels; item += gridDim.x, ++item_no) {
    const int tile = item / a.levels, l = item - tile * a.levels;
    const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % a.tiles_y, b = tile / (a.tiles_x * a.tiles_y);
    __syncthreads();  // the previous item's MMAs and gathers are done: sA / sB may be overwritten
    if (threadIdx.x == 0) OTF_TR(0);
    if (warp == 4 && lane == 0) {
      mbar_arrive_expect_tx(&bars->a_full, a.kchunks * kOtfTileBytes);
      for (int k = 0; k < a.kchunks; ++k) tma_load_4d(sA + k * kOtfTileBytes, &tmA, &bars->a_full, k * 64, tx * 16, ty * 8, b);
    }
    // ---- this thread's query ----
    const int qy = ty * 8 + (q_local >> 4), qx = tx * 16 + (q_local & 15);
    const bool q_in = warp < 4 && qy < a.H && qx < a.W;
    const size_t q = ((size_t)b * a.H + (q_in ? qy : 0)) * a.W + (q_in ? qx : 0);
    float cx = 0.f, cy = 0.f;
    if (q_in) {
      const float2 c = __ldg(reinterpret_cast<const float2*>(a.coords) + q);
      cx = c.x;
      cy = c.y;
    }
    bool a_waited = false;
    {
      const CUtensorMap* tmB = l == 0 ? &tmB0 : (l == 1 ? &tmB1 : (l == 2 ? &tmB2 : &tmB3));
      const int Hl = a.lh[l], Wl = a.lw[l];
      // ---- window geometry of this query at this level ----
      int x0 = 0, y0 = 0;
      float w00 = 0.f, w10 = 0.f, w01 = 0.f, w11 = 0.f;
      bool live = false;  // the window overlaps the map (otherwise all 81 outputs are zero)
      if (q_in) {
        const float sc = 1.0f / (float)(1 << l);
        const float x = cx * sc, y = cy * sc;
        const bool finite = (fabsf(x) < 1e7f) && (fabsf(y) < 1e7f);
        const float xf = finite ? floorf(x) : -1e6f, yf = finite ? floorf(y) : -1e6f;
        const float fx = finite ? x - xf : 0.f, fy = finite ? y - yf : 0.f;
        w00 = (1.f - fx) * (1.f - fy) * a.scale;
        w10 = fx * (1.f - fy) * a.scale;
        w01 = (1.f - fx) * fy * a.scale;
        w11 = fx * fy * a.scale;
        x0 = (int)xf - R;
        y0 = (int)yf - R;
        live = x0 + D - 1 >= 0 && x0 < Wl && y0 + D - 1 >= 0 && y0 < Hl;
      }
      // ---- region of the item: anchored at the smallest window origin of the live queries ----
      if (warp < 4) {
        int mx = live ? x0 : 0x7fffffff, my = live ? y0 : 0x7fffffff, My = live ? y0 : -0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          mx = min(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          my = min(my, __shfl_xor_sync(0xffffffffu, my, o));
          My = max(My, __shfl_xor_sync(0xffffffffu, My, o));
        }
        if (lane == 0) {
          bars->red[warp][0] = mx;
          bars->red[warp][1] = my;
          bars->red[warp][2] = My;
        }
        named_barrier_sync(1, 128);
        if (threadIdx.x == 0) {
          int bx = 0x7fffffff, by = 0x7fffffff, By = -0x7fffffff;
          for (int w = 0; w < 4; ++w) {
            bx = min(bx, bars->red[w][0]);
            by = min(by, bars->red[w][1]);
            By = max(By, bars->red[w][2]);
          }
          const int any = bx != 0x7fffffff;
          int nb = 0;
          if (any) {
            nb = (By + D - 1 - by + 6) / 7;  // bands of 8 rows at stride 7 that cover region rows 0 .. By + D - 1 - by
            nb = nb < 1 ? 1 : (nb > kOtfMaxBands ? kOtfMaxBands : nb);
          }
          bars->bx0 = bx;
          bars->by0 = by;
          bars->nb = nb;
          bars->any = any;
        }
      }
      __syncthreads();  // region known to everybody (and the staging rows of the previous level have been written out)
      const int bx0 = bars->bx0, by0 = bars->by0, nb = bars->nb;
      if (threadIdx.x == 0) { OTF_TR(1); if (a.trace && item_no == 1) a.trace[blockIdx.x * 64 + 2] = (unsigned long long)nb; }
      const int cxo = x0 - bx0, ryo = y0 - by0;  // column / row of the window's first tap inside the region
      // a window the region cannot hold: too far right of the anchor, or below the last band
      const bool outlier = live && (cxo + D > kOtfRW || ryo + D - 1 > 7 * nb);
      if (q_in && outlier) a.flags[q] = 1;
      const bool mine = live && !outlier;
      unsigned short* orow = sOut + q_local * SP;
      if (warp < 4 && !mine) {
#pragma unroll 9
        for (int c = 0; c < KK; ++c) orow[c] = 0;  // zero window (or a flagged query: its row is rewritten by the SIMT pass)
      }
      for (int kb = 0; kb < nb; ++kb) {
        if (kb > 0) __syncthreads();  // everybody is done with the previous band's dump (it aliases sB)
        if (warp == 4) {
          if (lane == 0) {
            OTF_TR(8 + kb * 6 + 0);
            mbar_arrive_expect_tx(&bars->b_full, a.kchunks * 2 * kOtfTileBytes);
            for (int k = 0; k < a.kchunks; ++k)
              tma_load_4d(sB + k * 2 * kOtfTileBytes, tmB, &bars->b_full, k * 64, bx0, by0 + 7 * kb, b);
          }
        } else if (warp == 5) {
          if (lane == 0) {
            if (!a_waited) mbar_wait(&bars->a_full, par_a);
            mbar_wait(&bars->b_full, par_b);
            OTF_TR(8 + kb * 6 + 1);
            tc_fence_after();
            const uint32_t idesc = make_idesc_f16(128, 256, a.ab_fmt);
            for (int k = 0; k < a.kchunks; ++k) {
              const uint64_t da = make_desc_k_sw128(smem_u32(sA + k * kOtfTileBytes));
              const uint64_t db = make_desc_k_sw128(smem_u32(sB + k * 2 * kOtfTileBytes));
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base, desc_advance(da, kk * 32), desc_advance(db, kk * 32), idesc, (k | kk) != 0);
            }
            umma_commit(&bars->acc_full);
          }
        } else {
          // ---- epilogue: accumulator row -> shared memory (scaled by the blend weights later; storage-type rounding) ----
          mbar_wait(&bars->acc_full, par_acc);
          if (threadIdx.x == 0) OTF_TR(8 + kb * 6 + 2);
          tc_fence_after();
          uint8_t* drow = sB + q_local * kOtfDumpPitch;  // 528-byte rows: the 32 lanes' 16-byte stores fall on 32 different bank groups
          const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 u;
              u.x = otf_pack2<T>(__uint_as_float(r[8 * g + 0]), __uint_as_float(r[8 * g + 1]));
              u.y = otf_pack2<T>(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3]));
              u.z = otf_pack2<T>(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
              u.w = otf_pack2<T>(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
              *reinterpret_cast<uint4*>(drow + (((c << 2) | g) << 4)) = u;
            }
          }
          tc_fence_before();
          if (threadIdx.x == 0) OTF_TR(8 + kb * 6 + 3);
          // ---- gather: the window rows j whose tap pair (region rows ryo + j, ryo + j + 1) lies in this band ----
          if (mine) {
            const unsigned short* dr = reinterpret_cast<const unsigned short*>(drow);
#pragma unroll 1
            for (int j = 0; j < K; ++j) {
              const int rr = ryo + j - 7 * kb;  // region row of the upper tap, relative to the band
              if (rr < 0 || rr > 6) continue;
              float up[D], dn[D];
#pragma unroll
              for (int i = 0; i < D; ++i) {
                const int e0 = rr * kOtfRW + cxo + i, e1 = e0 + kOtfRW;
                up[i] = otf_bits_to_f32<T>(dr[e0]);
                dn[i] = otf_bits_to_f32<T>(dr[e1]);
              }
#pragma unroll
              for (int i = 0; i < K; ++i)
                orow[i * K + j] = otf_f32_to_bits<T>(w00 * up[i] + w10 * up[i + 1] + w01 * dn[i] + w11 * dn[i + 1]);
            }
          }
          if (threadIdx.x == 0) OTF_TR(8 + kb * 6 + 4);
          fence_proxy_async();  // this thread's generic-proxy accesses of the dump precede the next band's TMA writes to the same bytes
        }
        a_waited = true;
        par_b ^= 1;
        par_acc ^= 1;
      }
      // ---- the level's 81 outputs of the 128 queries: coalesced 2-byte runs (81 consecutive channels per query) ----
      if (warp < 4) {
        named_barrier_sync(1, 128);
        unsigned short* outp = reinterpret_cast<unsigned short*>(a.out);
        for (int rq = warp; rq < 128; rq += 4) {
          const int yy = ty * 8 + (rq >> 4), xx = tx * 16 + (rq & 15);
          if (yy >= a.H || xx >= a.W) continue;
          unsigned short* dst = outp + (((size_t)b * a.H + yy) * a.W + xx) * a.out_stride + l * KK;
          const unsigned short* src = sOut + rq * SP;
          for (int c = lane; c < KK; c += 32) dst[c] = src[c];
        }
        named_barrier_sync(1, 128);  // sOut is free for the next level
        if (threadIdx.x == 0) OTF_TR(3);
      }
    }
    if (a_waited) par_a ^= 1;
    else if (warp == 5 && lane == 0) {  // no level had a live window: the A tile was loaded but never consumed
      mbar_wait(&bars->a_full, par_a);
      par_a ^= 1;
    } else par_a ^= 1;
    // pad columns of the pixel-major rows (out_stride > levels * 81): zero
    if (warp < 4 && q_in && l == 0) {
      unsigned short* dst = reinterpret_cast<unsigned short*>(a.out) + q * a.out_stride;
      for (int c = a.levels * KK; c < a.out_stride; ++c) dst[c] = 0;
    }
  }
#undef OTF_TR
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<256>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
int corr_onthefly_simt_flagged(const void* fmap1, void* const* pyr, const float* coords, void* out, const unsigned char* flags, int B, int H,
                               int W, int C, int levels, int radius, pfb_dtype dtype, int out_stride, cudaStream_t s);  // corr.cu

bool corr_onthefly_umma_supported(int B, int H, int W, int C, int levels, int radius, pfb_dtype dt, int out_stride) {
  if (dt != PFB_F16 && dt != PFB_BF16) return false;
  if (radius != 4 || levels < 1 || levels > 4) return false;
  if (C % 64 != 0 || C > 256) return false;
  if (out_stride < levels * 81) return false;
  return B > 0 && H > 0 && W > 0;
}

int corr_onthefly_umma(const void* fmap1, void* const* pyr, const float* coords, void* out, unsigned char* flags, int B, int H, int W, int C,
                       int levels, pfb_dtype dt, int out_stride, cudaStream_t s) {
  OtfArgs a{};
  a.coords = coords; a.out = out; a.flags = flags;
  a.B = B; a.H = H; a.W = W; a.kchunks = C / 64; a.levels = levels; a.out_stride = out_stride;
  a.scale = 1.0f / sqrtf((float)C);
  a.ab_fmt = dt == PFB_F16 ? 0 : 1;
  a.tiles_x = ceil_div(W, 16); a.tiles_y = ceil_div(H, 8); a.n_tiles = a.tiles_x * a.tiles_y * B;
  CUtensorMap tmA, tmB[4];
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {64, 16, 8, 1};
    int rc = make_tensor_map(&tmA, fmap1, dt, 4, dims, str, box);
    if (rc) return rc;
  }
  for (int l = 0; l < 4; ++l) {
    const int ll = l < levels ? l : 0;
    a.lh[l] = H >> ll; a.lw[l] = W >> ll;
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)a.lw[l], (uint64_t)a.lh[l], (uint64_t)B};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)a.lw[l] * C * 2, (uint64_t)a.lh[l] * a.lw[l] * C * 2};
    uint32_t box[4] = {64, (uint32_t)kOtfRW, 8, 1};
    int rc = make_tensor_map(&tmB[l], pyr[ll], dt, 4, dims, str, box);
    if (rc) return rc;
  }
  PFB_CUDA(cudaMemsetAsync(flags, 0, (size_t)B * H * W, s));
  const int b_bytes = a.kchunks * 2 * kOtfTileBytes < 128 * kOtfDumpPitch ? 5 * kOtfTileBytes : a.kchunks * 2 * kOtfTileBytes;
  const size_t smem = (size_t)a.kchunks * kOtfTileBytes + b_bytes + ((128 * 82 * 2 + 15) & ~15) + sizeof(OtfBars) + 1024;
  int grid = sm_count();
  if (grid > a.n_tiles) grid = a.n_tiles;
  // PFB_OTF_TRACE=<file>: per-CTA phase timeline of the second work item (clock64 at the role hand-overs), appended as JSON lines
  static const char* trace_path = getenv("PFB_OTF_TRACE");
  if (trace_path) {
    PFB_CUDA(cudaMalloc(&a.trace, (size_t)grid * 64 * sizeof(unsigned long long)));
    PFB_CUDA(cudaMemsetAsync(a.trace, 0, (size_t)grid * 64 * sizeof(unsigned long long), s));
  }
  {
    ProfScope prof(KC_ONTHEFLY, s);
    if (dt == PFB_F16) {
      PFB_CUDA(cudaFuncSetAttribute(corr_onthefly_umma_kernel<__half, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      corr_onthefly_umma_kernel<__half, 4><<<grid, 192, smem, s>>>(tmA, tmB[0], tmB[1], tmB[2], tmB[3], a);
    } else {
      PFB_CUDA(cudaFuncSetAttribute(corr_onthefly_umma_kernel<__nv_bfloat16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      corr_onthefly_umma_kernel<__nv_bfloat16, 4><<<grid, 192, smem, s>>>(tmA, tmB[0], tmB[1], tmB[2], tmB[3], a);
    }
    PFB_LAUNCH_CHECK();
  }
  if (trace_path) {
    std::vector<unsigned long long> host((size_t)grid * 64);
    PFB_CUDA(cudaStreamSynchronize(s));
    PFB_CUDA(cudaMemcpy(host.data(), a.trace, host.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    PFB_CUDA(cudaFree(a.trace));
    if (FILE* f = fopen(trace_path, "a")) {
      fprintf(f, "{\"grid\": %d, \"kchunks\": %d, \"H\": %d, \"W\": %d, \"stamps\": [", grid, a.kchunks, H, W);
      for (size_t i = 0; i < host.size(); ++i) fprintf(f, "%s%llu", i ? "," : "", host[i]);
      fprintf(f, "]}\n");
      fclose(f);
    }
  }
  // queries whose windows did not fit their tile's region: the SIMT kernel, one warp per flagged query
  return corr_onthefly_simt_flagged(fmap1, pyr, coords, out, flags, B, H, W, C, levels, 4, dt, out_stride, s);
}

}  // namespace pfb

extern "C" PFB_API size_t pfb_corr_lookup_onthefly_tc_workspace_bytes(int B, int H, int W) {
  return B > 0 && H > 0 && W > 0 ? (size_t)B * H * W : 0;
}

extern "C" PFB_API int pfb_corr_lookup_onthefly_tc(const void* fmap1, void* const* fmap2_pyramid, const float* coords, void* out,
                                                   void* workspace, int B, int H, int W, int C, int levels, int radius, pfb_dtype dtype,
                                                   int out_stride, pfb_stream stream) {
  using namespace pfb;
  PFB_CHECK_ARG(fmap1 && fmap2_pyramid && coords && out && workspace, "corr_lookup_onthefly_tc: null pointer");
  for (int l = 0; l < levels && l < 4; ++l) PFB_CHECK_ARG(fmap2_pyramid[l], "corr_lookup_onthefly_tc: fmap2 level %d is null", l);
  PFB_CHECK_ARG((H >> (levels - 1)) >= 1 && (W >> (levels - 1)) >= 1, "corr_lookup_onthefly_tc: grid too small for %d levels", levels);
  if (!corr_onthefly_umma_supported(B, H, W, C, levels, radius, dtype, out_stride)) {
    set_error("corr_lookup_onthefly_tc: needs f16/bf16, radius 4, 1-4 levels, C a multiple of 64 <= 256 (C=%d r=%d L=%d dtype=%d)", C, radius, levels, (int)dtype);
    return PFB_ERR_UNSUPPORTED;
  }
  return corr_onthefly_umma(fmap1, fmap2_pyramid, coords, out, reinterpret_cast<unsigned char*>(workspace), B, H, W, C, levels, dtype, out_stride,
                            as_stream(stream));
}
