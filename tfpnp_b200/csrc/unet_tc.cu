// UNet(2,1) denoiser (tfpnp/pnp/denoiser/models/unet.py:34-131, denoiser/base.py:23-32) on the
// 5th-generation tensor cores: every 3x3 convolution except the 2-channel input layer is an
// implicit GEMM   D[128 pixels, BN couts] += A[128 pixels, kc cin] * W[BN couts, kc cin]^T
// issued as tcgen05.mma (kind::f16, fp32 accumulators in TMEM), with
//   * activations NHWC fp16 in HBM; the A tile of tap (dy,dx) is ONE 4-D TMA box
//     {kc channels, TW, TH, TB} at (w0+dx, h0+dy): the conv's zero padding is TMA's
//     out-of-bounds fill, and the decoder's torch.cat([skip, up]) (unet.py:119) is a K-loop over
//     two tensor maps -- neither padding nor concat is ever materialised;
//   * weights re-laid-out once to [tap][Cout][Cin] fp16 (K-major B operand), 3-D TMA boxes;
//   * 128B/64B-swizzled K-major smem tiles feeding UMMA descriptors directly;
//   * a warp-specialised CTA: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue
//     (tcgen05.ld -> +bias -> LeakyReLU(0.2) -> fp16 -> NHWC store), mbarrier full/empty ring.
// Precision modes: FP16 (one product) and FP16X3 (activations and weights split into fp16
// hi+lo; hi*hi + lo*hi + hi*lo accumulate in the same fp32 TMEM tile: ~22-bit operands).
#include "common.cuh"
#include "sm100.cuh"
#include <cuda.h>
#include <vector>
#include <map>
#include <cstring>
#include <cstdlib>
#include <string>

namespace tfpnp {
namespace {

using namespace sm100;

using grad_elem::kLoScale;           // residual planes of the split-fp16 mode are stored x 2^11 (grad_elem.cuh)
using grad_elem::kLoInv;
constexpr int kTileM = 128;          // pixels per CTA tile (UMMA M)
constexpr int kConvThreads = 192;    // 6 warps

struct ConvParams {
  CUtensorMap a_map[2][2];  // [source][plane]  activations {C, W, H, B}
  CUtensorMap w_map[2];     // [plane]          weights     {Cin, Cout, 9}
  int nchunk0, nchunk1;     // channel chunks of source 0 / 1
  int kc;                   // channels per chunk (32 -> 64B rows, 64 -> 128B rows)
  int nprod;                // 1 (FP16) or 3 (FP16X3)
  int TW, TH, TB;           // tile geometry, TW*TH*TB == 128
  int box_rows;             // pixel rows one A box really carries (TB clamped to B)
  int tiles_w, tiles_h;
  int B, H, W, Cout;
  int dil;                  // tap spacing (1 for the UNet; 1..4 for the dilated IRCNN layers)
  float slope;              // activation max(v, slope*v): 0.2 = LeakyReLU(0.2), 0 = ReLU
  const float* bias;
  __half* out_hi;
  __half* out_lo;           // nullptr unless FP16X3
};


struct H8 { __half2 v[4]; };  // 8 channels = 16 bytes

__device__ __forceinline__ uint64_t pack_desc(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

// packed fp32 arithmetic (sm_100 FFMA2 / FMUL2): two lanes per instruction, scalar first operand broadcast
__device__ __forceinline__ float2 fmul2(float s, float2 a) {
  unsigned long long rs, ra, rd;
  asm("mov.b64 %0, {%1, %1};" : "=l"(rs) : "f"(s));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(rs), "l"(ra));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 ffma2(float s, float2 a, float2 c) {   // s * a + c
  unsigned long long rs, ra, rc, rd;
  asm("mov.b64 %0, {%1, %1};" : "=l"(rs) : "f"(s));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(rs), "l"(ra), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// Epilogue pieces for one 32-column slice of an accumulator row.
// act: + bias (smem broadcast), LeakyReLU(0.2) (unet.py:22) in fp32
__device__ __forceinline__ void epilogue_act32(const uint32_t (&r)[32], const float* __restrict__ sbias, float (&v)[32],
                                               const float slope = 0.2f) {
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 bb = *reinterpret_cast<const float4*>(sbias + 4 * j4);
    v[4 * j4 + 0] = __uint_as_float(r[4 * j4 + 0]) + bb.x;
    v[4 * j4 + 1] = __uint_as_float(r[4 * j4 + 1]) + bb.y;
    v[4 * j4 + 2] = __uint_as_float(r[4 * j4 + 2]) + bb.z;
    v[4 * j4 + 3] = __uint_as_float(r[4 * j4 + 3]) + bb.w;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], slope * v[j]);
}
// store: pack to fp16 (hi) and, in FP16X3 mode, the fp16 residual (lo); 64-byte NHWC stores per plane
__device__ __forceinline__ void epilogue_store_nhwc32(const float (&v)[32], __half* __restrict__ out_hi,
                                                      __half* __restrict__ out_lo, size_t off) {
  uint32_t hi[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    __half2 hh = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    hi[j] = *reinterpret_cast<uint32_t*>(&hh);
  }
  // 32-byte (sector-sized) stores: STG.256, two per 32 channels
#pragma unroll
  for (int q = 0; q < 2; ++q) st_global_256(out_hi + off + 16 * q, &hi[8 * q]);
  if (out_lo) {
    uint32_t lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float2 back = __half22float2(*reinterpret_cast<__half2*>(&hi[j]));
      __half2 ll = __floats2half2_rn((v[2 * j] - back.x) * kLoScale, (v[2 * j + 1] - back.y) * kLoScale);
      lo[j] = *reinterpret_cast<uint32_t*>(&ll);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) st_global_256(out_lo + off + 16 * q, &lo[8 * q]);
  }
}
__device__ __forceinline__ void epilogue_store32(const uint32_t (&r)[32], const float* __restrict__ sbias,
                                                 __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                                                 size_t off, bool store, float slope) {
  float v[32];
  epilogue_act32(r, sbias, v, slope);
  if (store) epilogue_store_nhwc32(v, out_hi, out_lo, off);
}

template <int BN>
struct ConvCfg {
  static constexpr int kStages = BN >= 128 ? 6 : 4;   // BN=128: 192 KB, 1 CTA/SM (these grids are < 148 CTAs anyway)
  static constexpr int kABytes = kTileM * 128;   // room for kc = 64
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;   // [main | corrections of the split-fp16 mode]
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + BN * 4 /*bias*/;
};

template <int BN>
__global__ void __launch_bounds__(kConvThreads, BN <= 128 ? 2 : 1)
conv3x3_tc(const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* empty = full + S;
  uint64_t* accum = empty + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
  float* sbias = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  const int mt = blockIdx.x;
  const int w0 = (mt % p.tiles_w) * p.TW;
  const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.TH;
  const int b0 = (mt / (p.tiles_w * p.tiles_h)) * p.TB;
  const int n0 = blockIdx.y * BN;
  const int nchunks = p.nchunk0 + p.nchunk1;
  const int kiters = 9 * nchunks * p.nprod;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.a_map[0][0]);
    prefetch_tensormap(&p.w_map[0]);
    if (p.nchunk1) prefetch_tensormap(&p.a_map[1][0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  if (threadIdx.x >= 64 && threadIdx.x - 64 < BN) sbias[threadIdx.x - 64] = p.bias[n0 + threadIdx.x - 64];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous layer's tail; its activations are needed from here on

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      const uint32_t stage_tx = (uint32_t)(p.box_rows + BN) * p.kc * 2;
      for (int it = 0; it < kiters; ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        const int prod = it % p.nprod;
        const int t = it / p.nprod;
        const int chunk = t % nchunks, tap = t / nchunks;
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        const int src = chunk < p.nchunk0 ? 0 : 1;
        const int cc = (src == 0 ? chunk : chunk - p.nchunk0) * p.kc;
        uint8_t* sA = smem + s * Cfg::kStageBytes;
        uint8_t* sB = sA + Cfg::kABytes;
        mbar_arrive_expect_tx(&full[s], stage_tx);
        tma_load_4d(sA, &p.a_map[src][prod == 1 ? 1 : 0], &full[s], cc, w0 + dx * p.dil, h0 + dy * p.dil, b0);
        tma_load_3d(sB, &p.w_map[prod == 2 ? 1 : 0], &full[s], chunk * p.kc, n0, tap);
      }
    }
  } else if (warp == 1) {
    {
      // ---------------- MMA issuer (whole warp converged; one elected lane issues) ----------------
      const uint32_t idesc = make_idesc_f16(kTileM, BN);
      const uint32_t row_bytes = p.kc * 2;
      const int ksteps = p.kc / 16;
      const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, row_bytes) >> 32);
      for (int it = 0; it < kiters; ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_lo = ((smem_u32(smem) + s * Cfg::kStageBytes) >> 4) | (1u << 16);
        const uint32_t b_lo = a_lo + (Cfg::kABytes >> 4);
        // split-fp16: product 0 (a_hi w_hi) -> columns [0, BN); products 1, 2 (a_lo w_hi, a_hi w_lo; residual planes are
        // stored x 2^11) -> columns [BN, 2 BN), scaled back in the epilogue
        const int prod = it % p.nprod, t = it / p.nprod;
        const uint32_t dcol = tmem_base + (prod ? BN : 0);
        if (elect_one()) {
          for (int kk = 0; kk < ksteps; ++kk) {
            umma_f16(dcol, pack_desc(a_lo + kk * 2, desc_hi), pack_desc(b_lo + kk * 2, desc_hi), idesc,
                     ((t | kk) != 0 || prod == 2) ? 1u : 0u);
          }
          umma_commit(&empty[s]);   // frees the smem slot once these MMAs have read it
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(accum);          // accumulator complete
      __syncwarp();
    }
  } else {
    // ---------------- epilogue: TMEM -> bias -> LeakyReLU -> fp16 NHWC ----------------
    mbar_wait(accum, 0);
    tc_fence_after();
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int m = q * 32 + lane;            // tile row = pixel
    const int tw = m % p.TW, th = (m / p.TW) % p.TH, tb = m / (p.TW * p.TH);
    const int b = b0 + tb, h = h0 + th, w = w0 + tw;
    const bool valid = b < p.B && h < p.H && w < p.W;
    const size_t pix = ((size_t)b * p.H + h) * p.W + w;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, r);
      if (p.nprod == 3) {
        uint32_t rc[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + BN + c0, rc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaf(__uint_as_float(rc[j]), kLoInv, __uint_as_float(r[j])));
      } else {
        tmem_ld_wait();
      }
      epilogue_store32(r, sbias + c0, p.out_hi, p.out_lo, pix * p.Cout + n0 + c0, valid, p.slope);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ==========================================================================================
// v2: persistent halo-tile convolution (used wherever the image tiles by 16 x 16 pixels)
// ==========================================================================================
// One CTA loops over output SUPER-TILES of 16x16 pixels = two UMMA M-tiles (left / right 16x8
// halves).  Per channel chunk it loads ONE halo box {KC, 18, 18, 1} at (w0-1, h0-1) -- 324 pixel
// rows of KC*2 bytes, TMA-swizzled -- and all nine taps of both halves read it through shifted
// UMMA descriptors: tap (ky,kx) of half h starts at pixel row ky*18 + kx + 8h and its sixteen
// 8-row groups are one halo row (18 pixels) apart (SBO).  Activation traffic from L2 drops from
// 9x to 1.27x of the tile, and every weight slab fetched feeds two M-tiles.  Weights are either
// RESIDENT in smem for the whole kernel (small layers: 9*Cin*BN*2 bytes) or streamed per
// (chunk, tap) through a deep second ring.  Accumulators are double-buffered in TMEM so the
// epilogue of super-tile i overlaps the MMAs of i+1; TMA runs ahead across tile boundaries.
// The single MMA-issuing thread only does 32-bit adds on precomputed descriptor words: the
// descriptor high words are loop constants (an issue loop with per-MMA descriptor construction
// measured ~300 clk/MMA, 10x the MMA itself).
constexpr int kHaloW = 18, kHaloH = 18, kHaloRows = kHaloW * kHaloH;

struct Conv2Params {
  CUtensorMap a_map[2][2];  // [source][plane]  halo boxes {KC, 18, 18, 1}
  CUtensorMap w_map[2];     // [plane]          {KC, BN, 1} over {Cin, Cout, 9}
  int nchunk0, nchunk1, nprod;
  int tiles_w, tiles_h, num_m_tiles, num_n_tiles;
  int B, H, W, Cout;
  int a_stage_bytes, num_a_stages;   // A ring
  int b_stage_bytes, num_b_stages;   // B ring (streamed) -- or resident slab size, 0 stages
  // fused nn.Upsample(x2, bilinear, align_corners=True) of source 1 (unet.py:99): a_map[1] then addresses the
  // LOW-RES tensor [B, H/2, W/2, C1] with an (unswizzled) 11x11 box and four transform warps interpolate the
  // 18x18 halo tile straight into the swizzled A stage; the up-sampled tensor is never materialised.
  int up_fused;
  float up_sy, up_sx;                // (Hin-1)/(H-1), (Win-1)/(W-1)
  int stg_bytes;                     // staging slot size (121 rows, 1024-aligned)
  int cluster;                       // CTAs per cluster sharing each streamed weight slab via TMA multicast (1, 2, 4)
  int dbg;                           // knock-out switches for bottleneck hunting (TFPNP_DBG): 1 = no stores,
                                     // 2 = no MMA issue, 4 = no activation TMA, 8 = no weight TMA
  unsigned long long* trace;         // optional [8][1024] globaltimer samples of CTA 0 (TFPNP_TRACE_FILE)
  const float* bias;
  __half* out_hi;
  __half* out_lo;
  // fused nn.MaxPool2d(2) of the output (unet.py:83): pooled NHWC tensor [B,H/2,W/2,Cout] (nullptr = off)
  __half* pool_hi;
  __half* pool_lo;
  // fused outconv 1x1 (Cout=32 -> 1) + residual + clamp (unet.py:124-131,65-66; denoiser/base.py:32):
  // x_out[pix] = clamp(d_in[pix] + b + sum_c w[c] * act[c], 0, 1); the 32-channel tensor is not stored
  const float* outc_w;               // [33] = w[32], b  (nullptr = off)
  const float* d_in;
  float* x_out;
};

constexpr int kMaxStages = 16;

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TRACE(row, idx)                                                                         \
  do {                                                                                          \
    if (p.trace && blockIdx.x == 0 && (idx) < 1024) p.trace[(row) * 1024 + (idx)] = gtimer();  \
  } while (0)

constexpr int kUpBox = 11;                     // low-res window edge feeding an 18-pixel halo edge
constexpr int kXformThreads = 288;                      // 9 transform warps
// Epilogue warps: 4 (one per TMEM lane quadrant) or 8 (two per quadrant, one per M-tile half).  The epilogue of a
// 256-pixel x BN tile is a long dependent instruction stream (~5 instructions per value); with one warp per quadrant
// it is as long as the tile's MMAs for BN >= 64 (measured: 64->64 @64x64 ran at the epilogue's pace, and a one-tile
// CTA exposes all of it).  The 64-channel-chunk variants run one CTA per SM and have the registers for 8 warps.
// Streamed weights: one ring stage holds TPS taps of one (n-tile, chunk): the three taps of a kernel row.  Measured on
// B200: per-tap stages for BN = 128 (8 x 16 KB instead of 2 x 48 KB) are SLOWER (60.0k -> 55.7k image-iters/s): the
// extra full/empty handshakes on the single MMA-issuing thread cost more than the deeper ring hides.
template <int BN> constexpr int conv2_taps_per_stage() { return 3; }
inline int conv2_taps_per_stage_rt(int) { return 3; }
template <int KC, bool FUSE> constexpr int conv2_nepi() { return (KC == 64 && !FUSE) ? 8 : 4; }
template <int KC, bool FUSE> constexpr int conv2_threads() { return 64 + 32 * conv2_nepi<KC, FUSE>() + (FUSE ? kXformThreads : 0); }

// SMALL: 8x8 images (the bottleneck level of a 128x128 input).  One UMMA M-tile = TWO images: the A box is
// {KC, 10 (x), 2 (image), 10 (y)} over the tensor viewed as {C, W, B, H}, i.e. smem row = y'*20 + img*10 + x', so the
// sixteen 8-pixel row groups (y, img) of tap (ky,kx) start at row ky*20 + kx and are uniformly 10 rows apart (SBO) --
// the same shifted-descriptor scheme as the 18x18 halo tile, one M-tile (one accumulator) per work item.
// XF (experiments, TFPNP_XFORM2=1|2; FUSE only).  1: the transform warps compute every output row of their row group
// independently from its four source vectors (same arithmetic, bit-identical results) instead of walking the rows with a
// dependent load -> interpolate chain: more instructions, but they can all be in flight.
// 2: the same with all three lerps in packed fp16 (HFMA2 / HMUL2, no conversions; three fp16 roundings instead of one).
template <int BN, int KC, bool RESIDENT, bool FUSE, bool SMALL = false, int XF = 0>
__global__ void __launch_bounds__(conv2_threads<KC, FUSE>(), (FUSE || KC == 64) ? 1 : 2)
conv3x3_tc2(const __grid_constant__ Conv2Params p) {
  constexpr int NEPI = conv2_nepi<KC, FUSE>();
  static_assert(!SMALL || (!FUSE && !RESIDENT && KC == 64 && NEPI == 8), "SMALL is built for the streamed 64-channel-chunk variant");
  constexpr int TPS = conv2_taps_per_stage<BN>();        // taps per weight-ring stage
  constexpr int TAP_ROWS = SMALL ? 20 : kHaloW;          // smem rows between kernel rows ky
  constexpr int SBO_ROWS = SMALL ? 10 : kHaloW;          // smem rows between the M-tile's 8-pixel row groups
  constexpr int A_ROWS = SMALL ? 200 : kHaloRows;        // rows one A stage receives
  static_assert(NEPI == 4 || BN >= 64, "two epilogue warps per quadrant split the tile by M-tile half in 64-column blocks");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t ROW = KC * 2;                 // bytes per pixel row of a chunk
  constexpr int KSTEPS = KC / 16;
  constexpr uint32_t SLAB = BN * ROW;              // one (tap, chunk) weight slab
  const int nchunks = p.nchunk0 + p.nchunk1;
  const int SA = p.num_a_stages, SB = p.num_b_stages;
  uint8_t* sA = smem;
  uint8_t* sW = smem + SA * p.a_stage_bytes;   // resident weights or the B ring
  const int w_region = RESIDENT ? 9 * nchunks * (int)SLAB : SB * TPS * (int)SLAB;
  uint8_t* sStg = sW + w_region;                                  // [2][stg_bytes] low-res windows (FUSE only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + (FUSE ? 2 * p.stg_bytes : 0));
  uint64_t* full_a = bars;
  uint64_t* empty_a = bars + kMaxStages;
  uint64_t* full_b = bars + 2 * kMaxStages;
  uint64_t* empty_b = bars + 3 * kMaxStages;
  uint64_t* w_full = bars + 4 * kMaxStages;
  uint64_t* tmem_full = w_full + 1;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;  // [2]
  uint64_t* stg_full = tmem_empty + 2;   // [2]
  uint64_t* stg_empty = stg_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stg_empty + 2);
  float* sbias = reinterpret_cast<float*>(bars + 4 * kMaxStages + 12);   // [Cout] <= 512 floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (p.dbg & 64) { pdl_wait(); return; }   // knock-out: launch + dependency wait only
  if (threadIdx.x == 0) TRACE(0, 1000);
  constexpr uint32_t kTmemCols = 4 * BN;
  // work items: (group of `cs` consecutive super-tiles, n-tile); CTA `crank` of a cluster takes
  // super-tile g*cs + crank, all CTAs of the cluster walk the same (n-tile, chunk, tap) sequence
  const int cs = RESIDENT ? 1 : p.cluster;
  const int crank = cs > 1 ? (int)cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << cs) - 1);
  const int total_items = ((p.num_m_tiles + cs - 1) / cs) * p.num_n_tiles;
  const int item0 = blockIdx.x / cs, item_step = gridDim.x / cs;           // 2 buffers x 2 M-tiles x BN (128..512, power of 2)

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.a_map[0][0]);
    prefetch_tensormap(&p.w_map[0]);
    if (p.nchunk1) prefetch_tensormap(&p.a_map[1][0]);
  }
  if (warp == 1) {
    if (lane < kMaxStages) {                                   // one ring slot per lane
      mbar_init(&full_a[lane], 1); mbar_init(&empty_a[lane], 1);
      mbar_init(&full_b[lane], 1); mbar_init(&empty_b[lane], cs);   // a weight slot is free when ALL cluster CTAs released it
    } else if (lane < kMaxStages + 2) {
      const int i = lane - kMaxStages;
      mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], NEPI);
      mbar_init(&stg_full[i], 1); mbar_init(&stg_empty[i], 1);
    } else if (lane == kMaxStages + 2) {
      mbar_init(w_full, 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) sbias[i] = p.bias[i];
  float* soutc = sbias + 512;                                    // [33] fused outconv weights + bias
  if (p.outc_w && threadIdx.x < 33) soutc[threadIdx.x] = p.outc_w[threadIdx.x];
  if (threadIdx.x == 64) TRACE(0, 1001);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (cs > 1) cluster_sync_all();   // peers' barriers are initialised before any remote arrive / multicast
  const uint32_t tmem_base = *tmem_slot;
  if (RESIDENT && warp == 0 && lane == 0) {     // weights are constants: fetch them before the dependency wait
    mbar_arrive_expect_tx(w_full, (uint32_t)(9 * nchunks) * SLAB);
    for (int tap = 0; tap < 9; ++tap)
      for (int c = 0; c < nchunks; ++c)
        tma_load_3d(sW + (tap * nchunks + c) * SLAB, &p.w_map[0], w_full, c * KC, 0, tap);
  }
  pdl_wait();   // prologue + weight prefetch overlapped the previous layer's tail
  if (threadIdx.x == 0) TRACE(0, 1002);
  const bool skip_roles = (p.dbg & 128) != 0;   // knock-out: prologue + teardown only

  if (skip_roles) {
    if (RESIDENT && warp == 1) mbar_wait(w_full, 0);   // do not exit with the weight TMA in flight
  } else if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      uint32_t ia = 0, ib = 0, iu = 0;
      const int rows_mc = BN / cs;   // weight-slab rows this CTA fetches (and multicasts)
      for (int t = item0; t < total_items; t += item_step) {
        const int nt = t % p.num_n_tiles;
        int m = (t / p.num_n_tiles) * cs + crank;
        if (m >= p.num_m_tiles) m = p.num_m_tiles - 1;      // padding CTA: recompute the last tile, stores masked
        const int w0 = (m % p.tiles_w) * 16, h0 = ((m / p.tiles_w) % p.tiles_h) * 16;
        const int b = m / (p.tiles_w * p.tiles_h);
        for (int c = 0; c < nchunks; ++c) {
          const int src = c < p.nchunk0 ? 0 : 1;
          const int cc = (src == 0 ? c : c - p.nchunk0) * KC;
          for (int prod = 0; prod < p.nprod; ++prod) {
            const int s = ia % SA;
            if (!(FUSE && src == 1)) mbar_wait(&empty_a[s], ((ia / SA) & 1) ^ 1);
            TRACE(0, ia);
            if (FUSE && src == 1) {
              // hand the low-res window to the transform warps as soon as a staging slot is free (the load runs
              // ahead of the A ring); they wait for A stage s themselves and fill it
              const int st = iu & 1;
              mbar_wait(&stg_empty[st], ((iu >> 1) & 1) ^ 1);
              mbar_arrive_expect_tx(&stg_full[st], (uint32_t)(kUpBox * kUpBox) * ROW);
              const int ys = (int)(p.up_sy * (float)(h0 > 0 ? h0 - 1 : 0));
              const int xs = (int)(p.up_sx * (float)(w0 > 0 ? w0 - 1 : 0));
              tma_load_4d(sStg + st * p.stg_bytes, &p.a_map[1][0], &stg_full[st], cc, xs, ys, b);
              ++iu;
            } else if (p.dbg & 4) mbar_arrive(&full_a[s]);
            else {
              mbar_arrive_expect_tx(&full_a[s], (uint32_t)A_ROWS * ROW);
              if (SMALL) tma_load_4d(sA + s * p.a_stage_bytes, &p.a_map[src][prod == 1 ? 1 : 0], &full_a[s], cc, -1, 2 * m, -1);
              else tma_load_4d(sA + s * p.a_stage_bytes, &p.a_map[src][prod == 1 ? 1 : 0], &full_a[s], cc, w0 - 1, h0 - 1, b);
            }
            ++ia;
            if (!RESIDENT) {
              const CUtensorMap* wm = &p.w_map[prod == 2 ? 1 : 0];
#pragma unroll 1
              for (int tg = 0; tg < 9 / TPS; ++tg) {     // one ring stage = TPS taps
                const int sb = ib % SB;
                mbar_wait(&empty_b[sb], ((ib / SB) & 1) ^ 1);
                if (p.dbg & 8) mbar_arrive(&full_b[sb]);
                else {
                  mbar_arrive_expect_tx(&full_b[sb], TPS * SLAB);
#pragma unroll
                  for (int tt = 0; tt < TPS; ++tt) {
                    uint8_t* dst = sW + sb * (TPS * SLAB) + tt * SLAB;
                    if (cs == 1) tma_load_3d(dst, wm, &full_b[sb], c * KC, nt * BN, tg * TPS + tt);
                    else tma_load_3d_mc(dst + crank * rows_mc * ROW, wm, &full_b[sb], cmask, c * KC,
                                        nt * BN + crank * rows_mc, tg * TPS + tt);
                  }
                }
                ++ib;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ---------------- MMA issuer (whole warp converged; one elected lane issues) ----------------
      const uint32_t idesc = make_idesc_f16(kTileM, BN);
      // descriptor words: hi = {SBO, version, swizzle} is constant per operand; lo = addr>>4 | LBO
      const uint32_t a_hi = (uint32_t)(make_smem_desc_ex(0, ROW, SBO_ROWS * ROW, 0) >> 32);
      const uint32_t b_hi = (uint32_t)(make_smem_desc(0, ROW) >> 32);
      const uint32_t lo_flags = 1u << 16;
      const uint32_t sA_lo = (smem_u32(sA) >> 4) | lo_flags;
      const uint32_t sW_lo = (smem_u32(sW) >> 4) | lo_flags;
      const uint32_t a_stage16 = (uint32_t)p.a_stage_bytes >> 4;
      if (RESIDENT) { mbar_wait(w_full, 0); tc_fence_after(); }
      uint32_t ia = 0, ib = 0, it = 0;
      uint32_t sa = 0, pha = 0, sb = 0, phb = 0;     // ring cursors (stage, phase)
      for (int t = item0; t < total_items; t += item_step, ++it) {
        const uint32_t buf = it & 1;
        if (lane == 0) TRACE(1, it);
        mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        if (lane == 0) TRACE(2, it);
        const uint32_t d0 = tmem_base + buf * (2 * BN);   // left half; right half at +BN
        uint32_t accumulate = 0;
        uint32_t acc_corr = 0;     // SMALL split-fp16: products 1, 2 accumulate in columns [BN, 2 BN) (residual planes are x 2^11)
        for (int c = 0; c < nchunks; ++c) {
          for (int prod = 0; prod < p.nprod; ++prod) {
            mbar_wait(&full_a[sa], pha);
            tc_fence_after();
            if (lane == 0) TRACE(3, ia);
            ++ia;
            const uint32_t a_lo = sA_lo + sa * a_stage16;
            const bool last_item = (c == nchunks - 1) && (prod == p.nprod - 1);
            if (RESIDENT) {
              if (elect_one()) {
                if (!(p.dbg & 2))
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                  const uint32_t a_tap = a_lo + (((tap / 3) * TAP_ROWS + tap % 3) * ROW >> 4);
                  const uint32_t b_lo = sW_lo + (uint32_t)(tap * nchunks + c) * (SLAB >> 4);
#pragma unroll
                  for (int kk = 0; kk < KSTEPS; ++kk) {
                    const uint64_t bd = pack_desc(b_lo + kk * 2, b_hi);
                    umma_f16(d0, pack_desc(a_tap + kk * 2, a_hi), bd, idesc, accumulate);
                    umma_f16(d0 + BN, pack_desc(a_tap + (8 * ROW >> 4) + kk * 2, a_hi), bd, idesc, accumulate);
                    accumulate = 1;
                  }
                }
                umma_commit(&empty_a[sa]);                 // smem slot free once these MMAs have read it
                if (last_item) umma_commit(&tmem_full[buf]);   // accumulator complete
              }
              accumulate = 1;
              __syncwarp();
            } else {
#pragma unroll 1
              for (int tg = 0; tg < 9 / TPS; ++tg) {
                mbar_wait(&full_b[sb], phb);
                tc_fence_after();
                const uint32_t b_stage = sW_lo + sb * (TPS * SLAB >> 4);
                // first tap of the stage: kernel row (tg*TPS)/3, column (tg*TPS)%3
                const uint32_t a_row = a_lo + ((((tg * TPS) / 3) * TAP_ROWS + (tg * TPS) % 3) * ROW >> 4);
                const bool corr = SMALL && prod > 0;
                if (elect_one()) {
                  if (!(p.dbg & 2)) {
                    uint32_t acc = corr ? acc_corr : accumulate;
                    const uint32_t dd = corr ? d0 + BN : d0;
#pragma unroll
                    for (int tt = 0; tt < TPS; ++tt) {
                      const uint32_t a_tap = a_row + (tt * ROW >> 4);
                      const uint32_t b_lo = b_stage + tt * (SLAB >> 4);
#pragma unroll
                      for (int kk = 0; kk < KSTEPS; ++kk) {
                        const uint64_t bd = pack_desc(b_lo + kk * 2, b_hi);
                        umma_f16(dd, pack_desc(a_tap + kk * 2, a_hi), bd, idesc, acc);
                        if (!SMALL) umma_f16(d0 + BN, pack_desc(a_tap + (8 * ROW >> 4) + kk * 2, a_hi), bd, idesc, acc);
                        acc = 1;
                      }
                    }
                  }
                  if (cs == 1) umma_commit(&empty_b[sb]);
                  else umma_commit_mc(&empty_b[sb], cmask);
                }
                if (corr) acc_corr = 1; else accumulate = 1;
                __syncwarp();
                if (++sb == (uint32_t)SB) { sb = 0; phb ^= 1; }
              }
            }
            if (!RESIDENT) {
              if (elect_one()) {
                umma_commit(&empty_a[sa]);
                if (last_item) umma_commit(&tmem_full[buf]);
              }
              __syncwarp();
            }
            if (++sa == (uint32_t)SA) { sa = 0; pha ^= 1; }
          }
        }
        if (lane == 0) TRACE(4, it);
      }
      (void)ia; (void)ib;
    }
  } else if (FUSE && warp >= 2 + NEPI) {
    // ---------------- transform warps: bilinear x2 (align_corners) of the low-res window into the A stage ------
    // Thread = (halo column k, 16-byte channel group j, row group g).  All index arithmetic of the interpolation is
    // done once per TILE (column set-up: source offsets + lx; row set-up: source row + ly for each of the thread's
    // RPG rows, kept in registers); per chunk the thread walks its rows top-down keeping the horizontally
    // interpolated source rows in registers (consecutive output rows share them) and uses packed fp32 FMAs
    // (FFMA2 / FMUL2), so an output row of 8 channels costs ~40 instructions instead of ~90.
    const int tid = threadIdx.x - (64 + 32 * NEPI);        // 0..287
    constexpr int CH16 = KC / 8;                           // 16-byte channel groups per pixel row
    constexpr int SLOTS = kHaloW * CH16;                   // 144 (KC=64) / 72 (KC=32)
    constexpr int GROUPS = kXformThreads / SLOTS;          // 2 / 4 row groups (SLOTS * GROUPS == 288)
    constexpr int RPG = (kHaloH + GROUPS - 1) / GROUPS;    // rows per group: 9 / 5
    constexpr int iROW = (int)ROW;
    const int g = tid / SLOTS, slot = tid % SLOTS;
    const int k = slot / CH16, j = slot % CH16;
    const int Hin = p.H >> 1, Win = p.W >> 1;
    const int r_begin = g * RPG;
    uint32_t ia = 0, iu = 0;
    for (int t = item0; t < total_items; t += item_step) {
      int m = (t / p.num_n_tiles) * cs + crank;
      if (m >= p.num_m_tiles) m = p.num_m_tiles - 1;
      const int w0 = (m % p.tiles_w) * 16, h0 = ((m / p.tiles_w) % p.tiles_h) * 16;
      const int ys = (int)(p.up_sy * (float)(h0 > 0 ? h0 - 1 : 0));
      const int xs = (int)(p.up_sx * (float)(w0 > 0 ? w0 - 1 : 0));
      // column set-up (fixed for the tile)
      const int xo = w0 - 1 + k;
      const bool x_ok = xo >= 0 && xo < p.W;
      const float fx = p.up_sx * (float)xo;
      const int x0 = (int)fx;
      const int x1 = x0 + (x0 < Win - 1 ? 1 : 0);
      const float lx = fx - (float)x0, wx = 1.f - lx;
      const int ox0 = (x0 - xs) * iROW + j * 16, ox1 = (x1 - xs) * iROW + j * 16;
      // row set-up (fixed for the tile): yrow = (staging row of y0) * 2 + (y1 != y0), or -1 for conv zero padding
      int yrow[RPG];
      float lyv[RPG];
#pragma unroll
      for (int rr = 0; rr < RPG; ++rr) {
        const int r = r_begin + rr;
        const int yo = h0 - 1 + r;
        const float fy = p.up_sy * (float)yo;
        const int y0 = (int)fy;
        lyv[rr] = fy - (float)y0;
        yrow[rr] = (x_ok && r < kHaloH && yo >= 0 && yo < p.H) ? (((y0 - ys) << 1) | (y0 < Hin - 1 ? 1 : 0)) : -1;
      }
      for (int c = 0; c < nchunks; ++c) {
        for (int prod = 0; prod < p.nprod; ++prod, ++ia) {
          if (c < p.nchunk0) continue;                      // source 0 comes by TMA
          const int s = ia % SA, st = iu & 1;
          mbar_wait(&stg_full[st], (iu >> 1) & 1);
          mbar_wait(&empty_a[s], ((ia / SA) & 1) ^ 1);      // the staging load ran ahead of the A ring: claim the stage here
          if (tid == 0) TRACE(5, 2 * iu);
          const uint8_t* stg = sStg + st * p.stg_bytes;
          uint8_t* dstA = sA + s * p.a_stage_bytes;
          float2 ha[4], hb[4];                              // horizontally interpolated source rows ya, yb
          int ya = -1, yb = -1;
          auto hrow = [&](int y, float2 (&hr)[4]) {
            const uint8_t* rp = stg + y * (kUpBox * iROW);
            const H8 a = *reinterpret_cast<const H8*>(rp + ox0), bq = *reinterpret_cast<const H8*>(rp + ox1);
#pragma unroll
            for (int e = 0; e < 4; ++e) hr[e] = ffma2(lx, __half22float2(bq.v[e]), fmul2(wx, __half22float2(a.v[e])));
          };
          if constexpr (XF != 0) {
            // every row on its own: 4 loads + 3 lerps per row, no state carried from row to row; batches of <= 5 rows keep
            // the live registers under the kernel's 128-register cap
            constexpr int RB = 5;
#pragma unroll
            for (int r0 = 0; r0 < RPG; r0 += RB) {
              uint4 outs[RB];
#pragma unroll
              for (int q = 0; q < RB; ++q) {
                const int rr = r0 + q;
                outs[q] = make_uint4(0, 0, 0, 0);
                if (rr < RPG && yrow[rr < RPG ? rr : 0] >= 0) {
                  const int yr = yrow[rr < RPG ? rr : 0];
                  const int y0 = yr >> 1, y1 = y0 + (yr & 1);
                  const float ly = lyv[rr < RPG ? rr : 0], wy = 1.f - ly;
                  H8 o;
                  if constexpr (XF == 2) {
                    const __half2 lx2 = __float2half2_rn(lx), wx2 = __float2half2_rn(wx);
                    const __half2 ly2 = __float2half2_rn(ly), wy2 = __float2half2_rn(wy);
                    const uint8_t* rp0 = stg + y0 * (kUpBox * iROW);
                    const uint8_t* rp1 = stg + y1 * (kUpBox * iROW);
                    const H8 a0 = *reinterpret_cast<const H8*>(rp0 + ox0), b0 = *reinterpret_cast<const H8*>(rp0 + ox1);
                    const H8 a1 = *reinterpret_cast<const H8*>(rp1 + ox0), b1 = *reinterpret_cast<const H8*>(rp1 + ox1);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const __half2 top = __hfma2(lx2, b0.v[e], __hmul2(wx2, a0.v[e]));
                      const __half2 bot = __hfma2(lx2, b1.v[e], __hmul2(wx2, a1.v[e]));
                      o.v[e] = __hfma2(ly2, bot, __hmul2(wy2, top));
                    }
                  } else {
                    float2 top[4], bot[4];
                    hrow(y0, top);
                    hrow(y1, bot);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const float2 v = ffma2(ly, bot[e], fmul2(wy, top[e]));
                      o.v[e] = __floats2half2_rn(v.x, v.y);
                    }
                  }
                  outs[q] = *reinterpret_cast<uint4*>(&o);
                }
              }
#pragma unroll
              for (int q = 0; q < RB; ++q) {
                const int rr = r0 + q, r = r_begin + rr;
                if (rr < RPG && r < kHaloH) {
                  const int pr = r * kHaloW + k;
                  const int swz = (ROW == 128) ? (pr & 7) : ((pr >> 1) & 3);
                  *reinterpret_cast<uint4*>(dstA + pr * iROW + ((j ^ swz) << 4)) = outs[q];
                }
              }
            }
          } else {
#pragma unroll
          for (int rr = 0; rr < RPG; ++rr) {
            const int r = r_begin + rr;
            uint4 outv = make_uint4(0, 0, 0, 0);            // conv zero padding outside the image
            if (yrow[rr] >= 0) {
              const int y0 = yrow[rr] >> 1, y1 = y0 + (yrow[rr] & 1);
              if (y0 != ya) {
                if (y0 == yb) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) ha[e] = hb[e];
                } else hrow(y0, ha);
                ya = y0;
              }
              if (y1 != yb) {
                if (y1 == ya) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) hb[e] = ha[e];
                } else hrow(y1, hb);
                yb = y1;
              }
              const float ly = lyv[rr], wy = 1.f - ly;
              H8 o;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 v = ffma2(ly, hb[e], fmul2(wy, ha[e]));
                o.v[e] = __floats2half2_rn(v.x, v.y);
              }
              outv = *reinterpret_cast<uint4*>(&o);
            }
            if (r < kHaloH) {
              // TMA-compatible swizzle of the A stage: 16-byte chunk index ^= row bits
              const int pr = r * kHaloW + k;
              const int swz = (ROW == 128) ? (pr & 7) : ((pr >> 1) & 3);
              *reinterpret_cast<uint4*>(dstA + pr * iROW + ((j ^ swz) << 4)) = outv;
            }
          }
          }   // XF == 0
          fence_proxy_async();                               // generic-proxy smem writes -> visible to the MMA (async proxy)
          asm volatile("bar.sync 1, 288;" ::: "memory");     // the nine transform warps
          if (tid == 0) { mbar_arrive(&full_a[s]); mbar_arrive(&stg_empty[st]); TRACE(5, 2 * iu + 1); }
          ++iu;
        }
      }
    }
  } else if (warp >= 2 && warp < 2 + NEPI) {
    // ---------------- epilogue ----------------
    const int q = warp & 3;                                // TMEM lane quadrant this warp may access
    // with 8 epilogue warps, warp pair member e handles M-tile half e (columns [e*BN, (e+1)*BN) of the tile's 2*BN)
    // (SMALL has one M-tile per item: the two warps of a quadrant split its BN columns instead)
    const int cb_begin = SMALL ? ((warp - 2) >> 2) * (BN / 2) : (NEPI == 8 ? ((warp - 2) >> 2) * BN : 0);
    const int cb_end = SMALL ? cb_begin + BN / 2 : (NEPI == 8 ? cb_begin + BN : 2 * BN);
    const int ml = q * 32 + lane;
    const int tw = ml & 7, th = ml >> 3;
    uint32_t it = 0;
    for (int t = item0; t < total_items; t += item_step, ++it) {
      const int nt = t % p.num_n_tiles;
      int m = (t / p.num_n_tiles) * cs + crank;
      const bool real_tile = m < p.num_m_tiles;
      if (!real_tile) m = p.num_m_tiles - 1;
      // SMALL: M row = (y*2 + img)*8 + x of images 2m, 2m+1
      const int w = SMALL ? tw : (m % p.tiles_w) * 16 + tw;
      const int h = SMALL ? (th >> 1) : ((m / p.tiles_w) % p.tiles_h) * 16 + th;
      const int b = SMALL ? 2 * m + (th & 1) : m / (p.tiles_w * p.tiles_h);
      const int n0 = nt * BN;
      const size_t pix = ((size_t)b * p.H + h) * p.W + w;
      const uint32_t buf = it & 1;
      mbar_wait(&tmem_full[buf], (it >> 1) & 1);
      tc_fence_after();
      if (warp == 2 && lane == 0) TRACE(6, it);
      // the two M-tile halves sit side by side in TMEM: walk their 2*BN columns in blocks of 64
      // (one tcgen05.ld round trip per two 32-column slices)
      if (SMALL && p.nprod == 3) {
        // split-fp16 at the 8x8 level: main product in columns [0, BN), the two correction products (x 2^11) in [BN, 2 BN)
#pragma unroll 1
        for (int c0 = cb_begin; c0 < cb_end; c0 += 32) {
          uint32_t r[32], rc[32];
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (2 * BN) + c0;
          tmem_ld_32x32(ta, r);
          tmem_ld_32x32(ta + BN, rc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaf(__uint_as_float(rc[j]), kLoInv, __uint_as_float(r[j])));
          float v[32];
          epilogue_act32(r, sbias + n0 + c0, v);
          if (real_tile && b < p.B) epilogue_store_nhwc32(v, p.out_hi, p.out_lo, pix * p.Cout + n0 + c0);
        }
      } else
#pragma unroll 1
      for (int cb = cb_begin; cb < cb_end; cb += 64) {
        if (p.dbg & 32) break;                  // knock-out: no TMEM reads, no epilogue math
        uint32_t r64[64];
        tmem_ld_32x64(tmem_base + ((uint32_t)(q * 32) << 16) + buf * (2 * BN) + cb, r64);
        tmem_ld_wait();
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const int col = cb + 32 * sl;
          const int half = col / BN, c0 = col % BN;
          uint32_t (&r)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r64[32 * sl]);
          float v[32];
          epilogue_act32(r, sbias + n0 + c0, v);
          const bool st_ok = real_tile && !(p.dbg & 1) && (!SMALL || b < p.B);
          if (p.outc_w) {                       // last layer: 1x1 conv + residual + clamp, fp32 out (BN == 32)
            float acc = soutc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc = fmaf(soutc[j], v[j], acc);
            if (st_ok) p.x_out[pix + half * 8] = fminf(fmaxf(p.d_in[pix + half * 8] + acc, 0.f), 1.f);
            continue;
          }
          if (st_ok) epilogue_store_nhwc32(v, p.out_hi, p.out_lo, (pix + half * 8) * p.Cout + n0 + c0);
          if (p.pool_hi) {                      // 2x2 max over (tw^1, th^1) = lanes ^1 and ^8
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
              v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 8));
            }
            if (st_ok && !(lane & 9)) {
              const size_t ppix = ((size_t)b * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + ((w + half * 8) >> 1);
              epilogue_store_nhwc32(v, p.pool_hi, p.pool_lo, ppix * p.Cout + n0 + c0);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      if (warp == 2 && lane == 0) TRACE(7, it);
    }
  }
  if (lane == 0) TRACE(0, 1010 + warp);
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // no CTA exits while a peer may still arrive on its barriers
  if (threadIdx.x == 0) TRACE(0, 1003);
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
  if (threadIdx.x == 64) TRACE(0, 1004);
}

// ==========================================================================================
// v2-pair: the streamed-weight kernel on CTA PAIRS (tcgen05.mma.cta_group::2) for the 128/256/512-channel levels
// ==========================================================================================
// A work item = (16x16 super-tile, 128-cout n-tile) on a cluster of two CTAs.  CTA r owns the 16-row x 8-column half r of
// the tile = one 128-row M-tile; the MMA runs M = 256 over the pair.  Per CTA: the half-halo A box {64, 10, 18, 1} (180 rows,
// 23 KB instead of 41.5 KB; tap (ky,kx) starts at row ky*10 + kx, 8-row groups 10 rows apart) and HALF of each weight slab
// (rows [64 r, 64 r + 64) of the n-tile: 8 KB per tap instead of 16 KB): a whole 64-channel chunk (half-halo + nine half
// slabs = 95 KB) is ONE ring stage behind ONE full / empty barrier pair, two stages deep, and every SM ingests half the
// weight bytes.  (Per-kernel-row stages were 1.5x slower than the single-CTA kernel: with M = 256 per MMA a 3-tap stage
// is 0.39 us of tensor work while each hand-shake crosses SMs twice.)  Only the leader (rank 0) issues MMAs; its
// full_* barriers collect the TMA bytes of BOTH CTAs (the peer's loads name the leader's barrier), completions are multicast
// to both CTAs (empty_* and tmem_full), and the peer's epilogue releases the accumulator on the leader's tmem_empty.
// (Bring-up: tools/ubench/mma_pair.cu.)
struct ConvPairParams {
  CUtensorMap a_map[2][2];  // [source][plane] half-halo boxes {64, 10, 18, 1}
  CUtensorMap w_map[2];     // [plane] {64, 64 rows, 1 tap} over {Cin, Cout, 9}
  int nchunk0, nchunk1;
  int tiles_w, tiles_h, num_m_tiles, num_n_tiles;
  int B, H, W, Cout;
  int num_a_stages, num_b_stages;
  const float* bias;
  __half* out_hi;
  __half* out_lo;           // fp16 residual plane (split-fp16 mode) or nullptr
  __half* pool_hi;          // fused nn.MaxPool2d(2) output or nullptr
  __half* pool_lo;
  int single_buf;           // experiment (TFPNP_PAIR_SINGLE=1): one accumulator buffer, 256 TMEM columns in the X3 instantiation
  // 8x8 images (the 512-channel level of a 128x128 input): one M = 256 tile = FOUR images, two per CTA.  The A box is
  // {KC, 10 (x), 2 (image), 10 (y)} over the tensor viewed as {C, W, B, H} (as in conv3x3_tc2<..., SMALL>): smem row = y'*20 + img*10 + x',
  // tap (ky,kx) starts at row ky*20 + kx, the sixteen 8-pixel row groups (y, img) are 10 rows apart.
  int small;
  int a_rows;               // rows one A box carries: 180 (half-halo) or 200 (small)
  int a_plane;              // bytes reserved per A plane (a_rows x row bytes, padded to 1 KB)
  int tap_rows;             // smem rows between kernel rows ky: 10 or 20
};

constexpr int kPairThreads = 64 + 32 * 8;
constexpr int kPairARows = 180, kPairABytes = 23552;            // 180 x 128 B, padded to a multiple of 1024
constexpr int kPairSlab = 64 * 128;                             // one tap's half slab
constexpr int kPairTPS = 9;                                     // taps per weight-ring stage (3: a kernel row, 9: a whole chunk)
constexpr int kPairBStage = kPairTPS * kPairSlab;
// split-fp16 (X3): separate rings -- A stages [hi half-halo | lo half-halo], weight stages of one kernel row
// [3 taps x (W_hi half slab | W_lo half slab)]: 36 MMAs (3 taps x 4 k-steps x 3 products) = 1.2 us per weight stage, the same
// tensor time per hand-shake as the fp16 kernel's whole-chunk stage
constexpr int kPairX3AStage = 2 * kPairABytes;
constexpr int kPairX3BStage = 3 * 2 * kPairSlab;

template <bool X3>
__global__ void __launch_bounds__(kPairThreads, 1)
conv3x3_pair(const __grid_constant__ ConvPairParams p) {
  constexpr int BN = 128, KC = 64, KSTEPS = 4;
  constexpr uint32_t ROW = 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.num_a_stages;                 // fp16: ring of chunk stages [A half-halo | 9 half slabs]; X3: the A ring
  const int SB = p.num_b_stages;                // X3: the weight ring
  const int nchunks = p.nchunk0 + p.nchunk1;
  const int kStage = X3 ? kPairX3AStage : p.a_plane + kPairBStage;
  uint8_t* sWr = smem + S * kStage;             // X3 weight ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * kStage + (X3 ? SB * kPairX3BStage : 0));
  uint64_t* full = bars;                        // [8]   (the leader's are the ones waited on)
  uint64_t* empty = bars + 8;                   // [8]
  uint64_t* full_b = bars + 16;                 // [8]   X3 weight ring
  uint64_t* empty_b = bars + 24;                // [8]
  uint64_t* tmem_full = bars + 48;              // [2]
  uint64_t* tmem_empty = bars + 50;             // [2]  (leader's collects both CTAs' epilogues)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 52);
  float* sbias = reinterpret_cast<float*>(bars + 54);           // [Cout] <= 512

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.a_map[0][0]);
    prefetch_tensormap(&p.w_map[0]);
    if (p.nchunk1) prefetch_tensormap(&p.a_map[1][0]);
    if (X3) {
      prefetch_tensormap(&p.a_map[0][1]);
      prefetch_tensormap(&p.w_map[1]);
      if (p.nchunk1) prefetch_tensormap(&p.a_map[1][1]);
    }
  }
  if (warp == 1) {
    if (lane < 8) { mbar_init(&full[lane], 2); mbar_init(&empty[lane], 1); mbar_init(&full_b[lane], 2); mbar_init(&empty_b[lane], 1); }
    if (lane < 2) { mbar_init(&tmem_full[lane], 1); mbar_init(&tmem_empty[lane], 16); }
    fence_barrier_init();
  }
  const uint32_t kTmem = (X3 && !p.single_buf) ? 512 : 256;     // X3: two buffers of [main | corrections (x 2^11)]
  constexpr uint32_t kBufCols = X3 ? 2 * BN : BN;
  const bool single = X3 && p.single_buf;
  if (warp == 2) tmem_alloc_2sm(tmem_slot, kTmem);
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) sbias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // peer's barriers initialised, TMEM allocated in both CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int total_items = p.num_m_tiles * p.num_n_tiles;
  const int item0 = blockIdx.x >> 1, item_step = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer (both CTAs): own half-halo + own half of the chunk's nine weight slabs ----------------
      uint32_t ia = 0, ib = 0;
      for (int t = item0; t < total_items; t += item_step) {
        const int nt = t % p.num_n_tiles, m = t / p.num_n_tiles;
        const int w0 = (m % p.tiles_w) * 16, h0 = ((m / p.tiles_w) % p.tiles_h) * 16;
        const int b = m / (p.tiles_w * p.tiles_h);
        for (int c = 0; c < nchunks; ++c, ++ia) {
          const int src = c < p.nchunk0 ? 0 : 1;
          const int cc = (src == 0 ? c : c - p.nchunk0) * KC;
          const int s = ia % S;
          mbar_wait(&empty[s], ((ia / S) & 1) ^ 1);
          const uint32_t fb = mapa_u32(&full[s], 0);            // the LEADER's barrier
          uint8_t* st = smem + s * kStage;
          if constexpr (!X3) {
            if (rank == 0) mbar_arrive_expect_tx(&full[s], 2u * ((uint32_t)p.a_rows * ROW + kPairBStage));
            else mbar_arrive_cluster(fb);
            if (p.small) tma_load_4d_2sm(st, &p.a_map[src][0], fb, cc, -1, 4 * m + 2 * (int)rank, -1);
            else tma_load_4d_2sm(st, &p.a_map[src][0], fb, cc, w0 - 1 + 8 * (int)rank, h0 - 1, b);
#pragma unroll
            for (int tt = 0; tt < 9; ++tt)
              tma_load_3d_2sm(st + p.a_plane + tt * kPairSlab, &p.w_map[0], fb, c * KC, nt * BN + 64 * (int)rank, tt);
          } else {
            if (rank == 0) mbar_arrive_expect_tx(&full[s], 2u * 2u * (kPairARows * ROW));
            else mbar_arrive_cluster(fb);
            tma_load_4d_2sm(st, &p.a_map[src][0], fb, cc, w0 - 1 + 8 * (int)rank, h0 - 1, b);
            tma_load_4d_2sm(st + kPairABytes, &p.a_map[src][1], fb, cc, w0 - 1 + 8 * (int)rank, h0 - 1, b);
#pragma unroll 1
            for (int tg = 0; tg < 3; ++tg, ++ib) {
              const int sb = ib % SB;
              mbar_wait(&empty_b[sb], ((ib / SB) & 1) ^ 1);
              const uint32_t fbb = mapa_u32(&full_b[sb], 0);
              if (rank == 0) mbar_arrive_expect_tx(&full_b[sb], 2u * kPairX3BStage);
              else mbar_arrive_cluster(fbb);
              uint8_t* wst = sWr + sb * kPairX3BStage;
#pragma unroll
              for (int tt = 0; tt < 3; ++tt) {
                tma_load_3d_2sm(wst + (2 * tt) * kPairSlab, &p.w_map[0], fbb, c * KC, nt * BN + 64 * (int)rank, tg * 3 + tt);
                tma_load_3d_2sm(wst + (2 * tt + 1) * kPairSlab, &p.w_map[1], fbb, c * KC, nt * BN + 64 * (int)rank, tg * 3 + tt);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ---------------- MMA issuer (leader only) ----------------
      const uint32_t idesc = make_idesc_f16(256, BN);
      const uint32_t a_hi = (uint32_t)(make_smem_desc_ex(0, ROW, 10 * ROW, 0) >> 32);
      const uint32_t b_hi = (uint32_t)(make_smem_desc(0, ROW) >> 32);
      const uint32_t lo_flags = 1u << 16;
      const uint32_t s_lo = (smem_u32(smem) >> 4) | lo_flags;
      const uint32_t w_lo = (smem_u32(sWr) >> 4) | lo_flags;
      uint32_t ia = 0, ib = 0, it = 0;
      for (int t = item0; t < total_items; t += item_step, ++it) {
        const uint32_t buf = single ? 0 : (it & 1);
        mbar_wait(&tmem_empty[buf], ((single ? it : (it >> 1)) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + buf * kBufCols;
        uint32_t accumulate = 0;
        for (int c = 0; c < nchunks; ++c, ++ia) {
          const int sa = ia % S;
          mbar_wait(&full[sa], (ia / S) & 1);
          tc_fence_after();
          const uint32_t a_lo = s_lo + sa * (kStage >> 4);
          if constexpr (!X3) {
            // one barrier wait and one release per chunk
            const uint32_t b_stage = a_lo + ((uint32_t)p.a_plane >> 4);
            const uint32_t tapr = (uint32_t)p.tap_rows;
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const uint32_t a_tap = a_lo + (((tap / 3) * tapr + tap % 3) * ROW >> 4);
                const uint32_t b_lo = b_stage + tap * (kPairSlab >> 4);
#pragma unroll
                for (int kk = 0; kk < KSTEPS; ++kk) {
                  umma2_f16(d0, pack_desc(a_tap + kk * 2, a_hi), pack_desc(b_lo + kk * 2, b_hi), idesc, accumulate);
                  accumulate = 1;
                }
              }
              umma2_commit_mc(&empty[sa], 3);                     // frees the chunk stage in BOTH CTAs
              if (c == nchunks - 1) umma2_commit_mc(&tmem_full[buf], 3);
            }
            accumulate = 1;
            __syncwarp();
          } else {
#pragma unroll 1
            for (int tg = 0; tg < 3; ++tg, ++ib) {
              const int sb = ib % SB;
              mbar_wait(&full_b[sb], (ib / SB) & 1);
              tc_fence_after();
              const uint32_t b_stage = w_lo + sb * (kPairX3BStage >> 4);
              const uint32_t a_row = a_lo + ((tg * 10) * ROW >> 4);
              if (elect_one()) {
#pragma unroll
                for (int tt = 0; tt < 3; ++tt) {
                  const uint32_t a_tap = a_row + (tt * ROW >> 4);
                  const uint32_t bh = b_stage + (2 * tt) * (kPairSlab >> 4), bl = bh + (kPairSlab >> 4);
#pragma unroll
                  for (int kk = 0; kk < KSTEPS; ++kk) {
                    const uint64_t ad = pack_desc(a_tap + kk * 2, a_hi);
                    umma2_f16(d0, ad, pack_desc(bh + kk * 2, b_hi), idesc, accumulate);                               // a_hi w_hi
                    umma2_f16(d0 + BN, ad, pack_desc(bl + kk * 2, b_hi), idesc, accumulate);                          // a_hi w_lo
                    umma2_f16(d0 + BN, pack_desc(a_tap + (kPairABytes >> 4) + kk * 2, a_hi), pack_desc(bh + kk * 2, b_hi), idesc, 1);   // a_lo w_hi
                    accumulate = 1;
                  }
                }
                umma2_commit_mc(&empty_b[sb], 3);
                if (tg == 2) {
                  umma2_commit_mc(&empty[sa], 3);
                  if (c == nchunks - 1) umma2_commit_mc(&tmem_full[buf], 3);
                }
              }
              accumulate = 1;
              __syncwarp();
            }
          }
        }
      }
    }
  } else {
    // ---------------- epilogue (both CTAs): own 128 rows; the two warps of a quadrant split the 128 columns ----------------
    const int q = warp & 3;
    const int e = (warp - 2) >> 2;
    const int ml = q * 32 + lane;
    const int tw = ml & 7, th = ml >> 3;
    const uint32_t te = mapa_u32(&tmem_empty[0], 0);            // the leader's tmem_empty[0]; [1] is 8 bytes further
    uint32_t it = 0;
    for (int t = item0; t < total_items; t += item_step, ++it) {
      const int nt = t % p.num_n_tiles, m = t / p.num_n_tiles;
      // small: M row = (y*2 + img)*8 + x of this CTA's images 4m + 2 rank, + 1
      const int w = p.small ? tw : (m % p.tiles_w) * 16 + 8 * (int)rank + tw;
      const int h = p.small ? (th >> 1) : ((m / p.tiles_w) % p.tiles_h) * 16 + th;
      const int b = p.small ? 4 * m + 2 * (int)rank + (th & 1) : m / (p.tiles_w * p.tiles_h);
      const bool b_ok = b < p.B;
      const int n0 = nt * BN;
      const size_t pix = ((size_t)b * p.H + h) * p.W + w;
      const uint32_t buf = single ? 0 : (it & 1);
      mbar_wait(&tmem_full[buf], (single ? it : (it >> 1)) & 1);
      tc_fence_after();
      uint32_t r64[64];
      if constexpr (!X3) {
        tmem_ld_32x64(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + e * 64, r64);
        tmem_ld_wait();
      }
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        const int c0 = e * 64 + 32 * sl;
        uint32_t (&r)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r64[32 * sl]);
        if constexpr (X3) {                    // main + corrections / 2^11
          uint32_t (&rc)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r64[32 * (sl ^ 1)]);
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + buf * kBufCols + c0;
          tmem_ld_32x32(ta, r);
          tmem_ld_32x32(ta + BN, rc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaf(__uint_as_float(rc[j]), kLoInv, __uint_as_float(r[j])));
        }
        float v[32];
        epilogue_act32(r, sbias + n0 + c0, v);
        if (b_ok) epilogue_store_nhwc32(v, p.out_hi, X3 ? p.out_lo : nullptr, pix * p.Cout + n0 + c0);
        if (p.pool_hi) {                      // 2x2 max over (tw^1, th^1) = lanes ^1 and ^8
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
            v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 8));
          }
          if (!(lane & 9)) {
            const size_t ppix = ((size_t)b * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
            epilogue_store_nhwc32(v, p.pool_hi, X3 ? p.pool_lo : nullptr, ppix * p.Cout + n0 + c0);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(te + buf * 8);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // nobody leaves (or frees TMEM) while the peer may still signal it
  if (warp == 2) tmem_dealloc_2sm(tmem_base, kTmem);
}

#include "conv_x3.cuh"
#include "conv_ws.cuh"

// ---- CUDA-core layers around the tensor-core convs ---------------------------------------

// inc.conv-0: 3x3 conv over cat[d, sigma*ones] (2 ch, denoiser/base.py:29-30) -> 32 ch, fp32 FFMA
// (K = 18: 0.2 % of the FLOPs), bias + LeakyReLU, NHWC fp16 out.  The 576 weights + 32 biases travel
// as kernel parameters, so every FFMA takes its weight straight from the constant bank.
// Tried in round 2 and rejected by measurement (34 us today at 48 x 128^2): folding the constant sigma channel into a
// per-channel term (10 instead of 18 FMAs per output, two threads per pixel: 37 us -- the kernel is not FMA-bound), and
// register-resident weights with a thread walking a 32-row strip (91 us: too few threads in flight); deferring the
// programmatic launch trigger to the end of the CTA (no change).
struct FirstLayerW { float w[32 * 18]; float b[32]; };

__global__ void __launch_bounds__(128)
conv_first_kernel(const float* __restrict__ d, const float* __restrict__ sigma, int64_t sstride,
                  const __grid_constant__ FirstLayerW wb, __half* __restrict__ out_hi,
                  __half* __restrict__ out_lo, int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.z, y = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const float sg = sigma[b * sstride];
  float in[18];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
    bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
    in[k] = ok ? d[((size_t)b * H + yy) * W + xx] : 0.f;
    in[9 + k] = ok ? sg : 0.f;
  }
  const size_t o = (((size_t)b * H + y) * W + x) * 32;
#pragma unroll
  for (int c0 = 0; c0 < 32; c0 += 8) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      float a0 = wb.b[c0 + j], a1 = wb.b[c0 + j + 1];
#pragma unroll
      for (int k = 0; k < 18; ++k) {
        a0 = fmaf(wb.w[(c0 + j) * 18 + k], in[k], a0);
        a1 = fmaf(wb.w[(c0 + j + 1) * 18 + k], in[k], a1);
      }
      a0 = fmaxf(a0, 0.2f * a0);
      a1 = fmaxf(a1, 0.2f * a1);
      __half2 hh = __floats2half2_rn(a0, a1);
      hi[j / 2] = *reinterpret_cast<uint32_t*>(&hh);
      float2 back = __half22float2(hh);
      __half2 ll = __floats2half2_rn((a0 - back.x) * kLoScale, (a1 - back.y) * kLoScale);
      lo[j / 2] = *reinterpret_cast<uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(out_hi + o + c0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (out_lo) *reinterpret_cast<uint4*>(out_lo + o + c0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}


__device__ __forceinline__ void load8(const __half* hi, const __half* lo, size_t off, float (&f)[8]) {
  H8 a = *reinterpret_cast<const H8*>(hi + off);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __half22float2(a.v[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  if (lo) {
    H8 l = *reinterpret_cast<const H8*>(lo + off);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __half22float2(l.v[i]);
      f[2 * i] = fmaf(t.x, kLoInv, f[2 * i]); f[2 * i + 1] = fmaf(t.y, kLoInv, f[2 * i + 1]);
    }
  }
}
__device__ __forceinline__ void store8(__half* hi, __half* lo, size_t off, const float (&f)[8]) {
  H8 a, l;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a.v[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    float2 back = __half22float2(a.v[i]);
    l.v[i] = __floats2half2_rn((f[2 * i] - back.x) * kLoScale, (f[2 * i + 1] - back.y) * kLoScale);
  }
  *reinterpret_cast<H8*>(hi + off) = a;
  if (lo) *reinterpret_cast<H8*>(lo + off) = l;
}

// nn.MaxPool2d(2) (unet.py:83), NHWC, 8 channels per thread; grid (x*C8 blocks, Ho, B)
__global__ void __launch_bounds__(256)
maxpool2_nhwc(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, __half* __restrict__ out_hi,
              __half* __restrict__ out_lo, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Wo * C8) return;
  const int c8 = idx % C8, x = idx / C8, y = blockIdx.y;
  const size_t b = blockIdx.z;
  float m[8], t[8];
  const size_t base = ((b * H + 2 * y) * W + 2 * x) * C + c8 * 8;
  load8(in_hi, in_lo, base, m);
  load8(in_hi, in_lo, base + C, t);
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], t[k]);
  load8(in_hi, in_lo, base + (size_t)W * C, t);
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], t[k]);
  load8(in_hi, in_lo, base + (size_t)W * C + C, t);
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], t[k]);
  store8(out_hi, out_lo, ((b * Ho + y) * Wo + x) * C + c8 * 8, m);
}

// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) (unet.py:99), NHWC, for the decoder heads whose
// up-sampling is not fused into the convolution.  One CTA per (output row, image): the two source rows of that output row
// (both planes in split-fp16) are staged in shared memory with coalesced 16-byte loads, then every thread interpolates
// (pixel, 8-channel group) items out of shared memory.  (The first version read its four corners straight from global
// memory: 4x the output size in L2 traffic -- 34.6 us for the 256-channel 16^2 -> 32^2 tensor of the split-fp16 mode,
// where this one moves 2x the INPUT.)
template <bool X3>
__global__ void __launch_bounds__(256)
upsample2_nhwc(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, __half* __restrict__ out_hi,
               __half* __restrict__ out_lo, int H, int W, int C) {
  extern __shared__ uint4 up_smem[];              // [plane][row 0/1][W * C / 8] 16-byte groups
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = 2 * H, Wo = 2 * W, C8 = C >> 3;
  const int y = blockIdx.x, bimg = blockIdx.y;
  const float sy = (float)(H - 1) / (float)(Ho - 1), sx = (float)(W - 1) / (float)(Wo - 1);
  const float fy = sy * y;
  const int y0 = (int)fy;
  const int y1 = y0 + (y0 < H - 1 ? 1 : 0);
  const float ly = fy - y0;
  const int row16 = W * C8;                       // 16-byte groups per source row
  const size_t img_in = (size_t)bimg * H * W * C;
  {
    const uint4* s0 = reinterpret_cast<const uint4*>(in_hi + img_in + (size_t)y0 * W * C);
    const uint4* s1 = reinterpret_cast<const uint4*>(in_hi + img_in + (size_t)y1 * W * C);
    for (int i = threadIdx.x; i < row16; i += blockDim.x) { up_smem[i] = s0[i]; up_smem[row16 + i] = s1[i]; }
    if (X3) {
      const uint4* l0 = reinterpret_cast<const uint4*>(in_lo + img_in + (size_t)y0 * W * C);
      const uint4* l1 = reinterpret_cast<const uint4*>(in_lo + img_in + (size_t)y1 * W * C);
      for (int i = threadIdx.x; i < row16; i += blockDim.x) { up_smem[2 * row16 + i] = l0[i]; up_smem[3 * row16 + i] = l1[i]; }
    }
  }
  __syncthreads();
  const size_t out_row = ((size_t)bimg * Ho + y) * Wo * C;
  for (int idx = threadIdx.x; idx < Wo * C8; idx += blockDim.x) {
    const int c8 = idx % C8, x = idx / C8;
    const float fx = sx * x;
    const int x0 = (int)fx;
    const int x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float lx = fx - x0;
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    float r[8];
    {
      const H8 a = *reinterpret_cast<const H8*>(&up_smem[x0 * C8 + c8]), b = *reinterpret_cast<const H8*>(&up_smem[x1 * C8 + c8]);
      const H8 c = *reinterpret_cast<const H8*>(&up_smem[row16 + x0 * C8 + c8]), d = *reinterpret_cast<const H8*>(&up_smem[row16 + x1 * C8 + c8]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 fa = __half22float2(a.v[i]), fb = __half22float2(b.v[i]);
        const float2 fc = __half22float2(c.v[i]), fd = __half22float2(d.v[i]);
        r[2 * i] = w00 * fa.x + w01 * fb.x + w10 * fc.x + w11 * fd.x;
        r[2 * i + 1] = w00 * fa.y + w01 * fb.y + w10 * fc.y + w11 * fd.y;
      }
    }
    if (X3) {
      const uint4* lo = up_smem + 2 * row16;
      const H8 a = *reinterpret_cast<const H8*>(&lo[x0 * C8 + c8]), b = *reinterpret_cast<const H8*>(&lo[x1 * C8 + c8]);
      const H8 c = *reinterpret_cast<const H8*>(&lo[row16 + x0 * C8 + c8]), d = *reinterpret_cast<const H8*>(&lo[row16 + x1 * C8 + c8]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 fa = __half22float2(a.v[i]), fb = __half22float2(b.v[i]);
        const float2 fc = __half22float2(c.v[i]), fd = __half22float2(d.v[i]);
        r[2 * i] += (w00 * fa.x + w01 * fb.x + w10 * fc.x + w11 * fd.x) * kLoInv;
        r[2 * i + 1] += (w00 * fa.y + w01 * fb.y + w10 * fc.y + w11 * fd.y) * kLoInv;
      }
    }
    H8 hi;
#pragma unroll
    for (int i = 0; i < 4; ++i) hi.v[i] = __floats2half2_rn(r[2 * i], r[2 * i + 1]);
    *reinterpret_cast<H8*>(out_hi + out_row + (size_t)idx * 8) = hi;
    if (X3) {
      H8 lo;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 back = __half22float2(hi.v[i]);
        lo.v[i] = __floats2half2_rn((r[2 * i] - back.x) * kLoScale, (r[2 * i + 1] - back.y) * kLoScale);
      }
      *reinterpret_cast<H8*>(out_lo + out_row + (size_t)idx * 8) = lo;
    }
  }
}

// outconv 1x1 32->1 (unet.py:124-131) + residual (unet.py:65-66) + clamp (denoiser/base.py:32)
__global__ void outc_nhwc(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo,
                          const float* __restrict__ w, const float* __restrict__ bias,
                          const float* __restrict__ d, float* __restrict__ out, size_t npix) {
  pdl_launch_dependents();
  pdl_wait();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  float acc = bias[0];
#pragma unroll
  for (int c8 = 0; c8 < 4; ++c8) {
    float f[8];
    load8(in_hi, in_lo, i * 32 + c8 * 8, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = fmaf(__ldg(w + c8 * 8 + k), f[k], acc);
  }
  float r = d[i] + acc;
  out[i] = fminf(fmaxf(r, 0.f), 1.f);
}

// ---- host side ---------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
    set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int encode_map(CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
               const cuuint32_t* box, int inner_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return TFPNP_ERR_CUDA;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = inner_bytes < 0 ? CU_TENSOR_MAP_SWIZZLE_NONE
                          : inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu, box %u %u %u)", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], box[0],
              box[1], box[2]);
    return TFPNP_ERR_CUDA;
  }
  return 0;
}

int set_conv_attrs() {
  static unsigned long long done_mask = 0;   // one bit per device
  int dev = 0;
  cudaGetDevice(&dev);
  const bool done = (done_mask >> (dev & 63)) & 1ull;
  if (done) return 0;
  TFPNP_CUDA_OK(cudaFuncSetAttribute(conv3x3_tc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<32>::kSmemBytes));
  TFPNP_CUDA_OK(cudaFuncSetAttribute(conv3x3_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<64>::kSmemBytes));
  TFPNP_CUDA_OK(cudaFuncSetAttribute(conv3x3_tc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<128>::kSmemBytes));
  TFPNP_CUDA_OK(cudaFuncSetAttribute(upsample2_nhwc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  TFPNP_CUDA_OK(cudaFuncSetAttribute(upsample2_nhwc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  done_mask |= 1ull << (dev & 63);
  return 0;
}

struct Conv2Plan {
  Conv2Params p;
  int BN = 0;
  int kc = 0;
  bool resident = false;
  bool small = false;      // 8x8 images, two per M-tile (conv3x3_tc2<..., SMALL>)
  bool x3n = false;        // split-fp16 with N-concatenated products (conv3x3_x3, conv_x3.cuh)
  int smem_bytes = 0;
  int grid = 0;
};

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
// programmatic dependent launch between the kernels of one denoiser call (TFPNP_PDL=0 disables)
thread_local bool g_no_pdl = false;   // set while tfpnp_denoiser_layer_profile times launches one by one
bool use_pdl() {
  static int v = env_int("TFPNP_PDL", 1);
  return v != 0 && !g_no_pdl;
}

template <int BN, int KC, bool RES, bool FUSE, bool SMALL = false, int XF = 0>
int launch_conv2_t(const Conv2Plan& c, cudaStream_t st) {
  static unsigned long long attr_set = 0;   // one bit per device (a function attribute is per device)
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_set >> (dev & 63) & 1ull)) {
    TFPNP_CUDA_OK(cudaFuncSetAttribute(conv3x3_tc2<BN, KC, RES, FUSE, SMALL, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       227 * 1024));
    attr_set |= 1ull << (dev & 63);
  }
  TFPNP_CUDA_OK(launch_ex(conv3x3_tc2<BN, KC, RES, FUSE, SMALL, XF>, dim3(c.grid), dim3(conv2_threads<KC, FUSE>()),
                          c.smem_bytes, st, use_pdl(), c.p.cluster, c.p));
  TFPNP_COUNT_LAUNCH();
  return 0;
}

template <int BN, bool RES, bool FUSE>
int launch_conv_x3_t(const Conv2Plan& c, cudaStream_t st) {
  static unsigned long long attr_set = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_set >> (dev & 63) & 1ull)) {
    TFPNP_CUDA_OK(cudaFuncSetAttribute(conv3x3_x3<BN, RES, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set |= 1ull << (dev & 63);
  }
  TFPNP_CUDA_OK(launch_ex(conv3x3_x3<BN, RES, FUSE>, dim3(c.grid), dim3(conv_x3_threads<FUSE>()), c.smem_bytes, st, use_pdl(),
                          c.p.cluster, c.p));
  TFPNP_COUNT_LAUNCH();
  return 0;
}

int launch_conv2(const Conv2Plan& c, cudaStream_t st) {
  if (c.x3n && c.p.up_fused) {      // the decoder heads 96 -> 32 (resident weights) and 192 -> 64 (streamed)
    if (c.BN == 32 && c.resident) return launch_conv_x3_t<32, true, true>(c, st);
    if (c.BN == 32 && !c.resident) return launch_conv_x3_t<32, false, true>(c, st);
    if (c.BN == 64 && !c.resident) return launch_conv_x3_t<64, false, true>(c, st);
    set_error("conv_x3: no fused-upsample variant for BN %d (resident %d)", c.BN, (int)c.resident);
    return TFPNP_ERR_INVALID;
  }
  if (c.x3n) {
    if (c.BN == 32) return c.resident ? launch_conv_x3_t<32, true, false>(c, st) : launch_conv_x3_t<32, false, false>(c, st);
    if (c.BN == 64) return c.resident ? launch_conv_x3_t<64, true, false>(c, st) : launch_conv_x3_t<64, false, false>(c, st);
    set_error("conv_x3: unsupported BN %d", c.BN);
    return TFPNP_ERR_INVALID;
  }
  const int key = c.BN * 1000 + c.kc * 10 + (c.resident ? 1 : 0);
  if (c.small) {
    if (key == 128640 && !c.p.up_fused) return launch_conv2_t<128, 64, false, false, true>(c, st);
    set_error("conv2: no 8x8 variant for BN %d KC %d (resident %d)", c.BN, c.kc, (int)c.resident);
    return TFPNP_ERR_INVALID;
  }
  const int xf = c.p.up_fused ? env_int("TFPNP_XFORM2", 2) : 0;   // 2 (default since round 2: +1.2 %, packed-fp16 lerps); 1: row-independent fp32; 0: row walk
  if (xf == 1) {   // experiment: row-independent transform warps (bit-identical results)
    switch (key) {
      case 32321: return launch_conv2_t<32, 32, true, true, false, 1>(c, st);
      case 64640: return launch_conv2_t<64, 64, false, true, false, 1>(c, st);
      case 128640: return launch_conv2_t<128, 64, false, true, false, 1>(c, st);
    }
  } else if (xf == 2) {   // experiment: the same with packed-fp16 lerps
    switch (key) {
      case 32321: return launch_conv2_t<32, 32, true, true, false, 2>(c, st);
      case 64640: return launch_conv2_t<64, 64, false, true, false, 2>(c, st);
      case 128640: return launch_conv2_t<128, 64, false, true, false, 2>(c, st);
    }
  }
  if (c.p.up_fused) {
    switch (key) {      // the four decoder conv-0 layers of UNet(2,1): 96->32, 192->64, 384->128, 768->256
      case 32321: return launch_conv2_t<32, 32, true, true>(c, st);
      case 64640: return launch_conv2_t<64, 64, false, true>(c, st);
      case 128640: return launch_conv2_t<128, 64, false, true>(c, st);
    }
    set_error("conv2: no fused-upsample variant for BN %d KC %d (resident %d)", c.BN, c.kc, (int)c.resident);
    return TFPNP_ERR_INVALID;
  }
  switch (key) {
    case 32321: return launch_conv2_t<32, 32, true, false>(c, st);
    case 64321: return launch_conv2_t<64, 32, true, false>(c, st);
    case 64641: return launch_conv2_t<64, 64, true, false>(c, st);
    case 32320: return launch_conv2_t<32, 32, false, false>(c, st);
    case 64320: return launch_conv2_t<64, 32, false, false>(c, st);
    case 64640: return launch_conv2_t<64, 64, false, false>(c, st);
    case 128640: return launch_conv2_t<128, 64, false, false>(c, st);
  }
  set_error("conv2: unsupported BN %d KC %d (resident %d)", c.BN, c.kc, (int)c.resident);
  return TFPNP_ERR_INVALID;
}

bool conv2_eligible(int H, int W) { return env_int("TFPNP_CONV_V2", 1) != 0 && W % 16 == 0 && H % 16 == 0; }
// 8x8 images with >= 2 images, 64-channel chunks and Cout a multiple of 128 (the 512-channel level of a 128x128 input)
bool conv2_small_eligible(int H, int W, int B, int C0, int C1, int Cout) {
  return env_int("TFPNP_CONV_V2", 1) != 0 && env_int("TFPNP_CONV_SMALL", 1) != 0 && H == 8 && W == 8 && B >= 2 &&
         C0 % 64 == 0 && C1 % 64 == 0 && Cout % 128 == 0;
}

int encode_map(CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
               const cuuint32_t* box, int inner_bytes);

// Fill everything of a Conv2Plan except the tensor maps.
int plan_conv2_geometry(Conv2Plan& c, int C0, int C1, int Cout, int B, int H, int W, bool x3, bool fuse_up = false) {
  Conv2Params& p = c.p;
  c.small = (H == 8 && W == 8);
  const int Cin = C0 + C1;
  // split-fp16 on the 32/64-channel levels: the N-concatenated kernel (32-channel chunks, both planes in one stage)
  c.x3n = x3 && !c.small && (Cout == 32 || Cout == 64);
  if (x3 && !c.x3n && !c.small) {
    // (the K-loop-over-products form of conv3x3_tc2 has no room for the separate correction accumulator the scaled residual
    // planes need, except at the 8x8 level; everything else of a split-fp16 UNet runs on conv3x3_x3 / conv3x3_pair<true>)
    set_error("split-fp16: no halo-tile kernel for %d+%d -> %d at %dx%d", C0, C1, Cout, H, W);
    return TFPNP_ERR_UNSUPPORTED;
  }
  const int kc = (C0 % 64 == 0 && C1 % 64 == 0 && !c.x3n) ? 64 : 32;
  c.kc = kc;
  c.BN = Cout >= 128 ? 128 : Cout;
  if (kc == 32 && c.BN == 128) c.BN = 64;                   // (not a UNet(2,1) shape; keeps the variant table small)
  p.nchunk0 = C0 / kc; p.nchunk1 = C1 / kc; p.nprod = x3 ? 3 : 1;
  p.tiles_w = c.small ? 1 : W / 16; p.tiles_h = c.small ? 1 : H / 16;
  p.num_m_tiles = c.small ? (B + 1) / 2 : p.tiles_w * p.tiles_h * B;
  p.num_n_tiles = Cout / c.BN;
  p.B = B; p.H = H; p.W = W; p.Cout = Cout;
  p.dbg = env_int("TFPNP_DBG", 0);
  const int row_bytes = kc * 2;
  p.a_stage_bytes = ((c.small ? 200 : kHaloRows) * row_bytes + 1023) & ~1023;
  p.b_stage_bytes = c.BN * row_bytes;                       // multiple of 1024 for all (BN, kc) used
  if (c.x3n) {
    // stage = [hi plane | lo plane]; weights [W_hi | W_lo] per (tap, chunk); one CTA per SM
    p.nprod = 1;                                            // the three products are issued per tap, not as K passes
    p.a_stage_bytes *= 2;
    const int slab2 = 2 * p.b_stage_bytes;
    const int w_all = 9 * (Cin / kc) * slab2;
    const int misc3 = 1024 + 1024 + 2048 + 256 + (fuse_up ? 2 * kX3StgSlot : 0);   // align + barriers + bias + outc (+ staging)
    const int budget = 227 * 1024 - misc3;
    c.resident = w_all + 2 * p.a_stage_bytes <= budget && env_int("TFPNP_CONV_RESIDENT", 1) != 0 &&
                 Cin / kc <= env_int("TFPNP_X3N_RES_MAXCHUNKS", 3);
    if (c.resident) {
      p.num_b_stages = 0;
      p.num_a_stages = (budget - w_all) / p.a_stage_bytes;
      if (p.num_a_stages > 4) p.num_a_stages = 4;      // (the fused variant tracks at most 4 slots)
      c.smem_bytes = p.num_a_stages * p.a_stage_bytes + w_all + misc3;
    } else {
      p.num_a_stages = 3;
      int sb = (budget - p.num_a_stages * p.a_stage_bytes) / (3 * slab2);
      if (sb < 2) { p.num_a_stages = 2; sb = (budget - p.num_a_stages * p.a_stage_bytes) / (3 * slab2); }
      p.num_b_stages = sb > kMaxStages ? kMaxStages : sb;
      c.smem_bytes = p.num_a_stages * p.a_stage_bytes + p.num_b_stages * 3 * slab2 + misc3;
    }
    int dev3 = 0, sms3 = 148;
    cudaGetDevice(&dev3);
    cudaDeviceGetAttribute(&sms3, cudaDevAttrMultiProcessorCount, dev3);
    int cs3 = 1;
    if (!c.resident) {
      cs3 = env_int("TFPNP_CONV_CLUSTER", 2);
      while (cs3 > 1 && (p.num_m_tiles < cs3 || (c.BN / cs3) * row_bytes % 1024 != 0)) cs3 /= 2;
    }
    p.cluster = cs3;
    p.up_fused = fuse_up ? 1 : 0;
    p.up_sy = (float)(H / 2 - 1) / (float)(H - 1);
    p.up_sx = (float)(W / 2 - 1) / (float)(W - 1);
    const int items3 = (p.num_m_tiles + cs3 - 1) / cs3;
    const int maxc3 = sms3 / cs3;
    c.grid = (items3 < maxc3 ? items3 : maxc3) * cs3;
    return 0;
  }
  const int w_bytes = 9 * (Cin / kc) * p.b_stage_bytes;
  p.up_fused = fuse_up ? 1 : 0;
  p.stg_bytes = (kUpBox * kUpBox * row_bytes + 1023) & ~1023;
  p.up_sy = (float)(H / 2 - 1) / (float)(H - 1);
  p.up_sx = (float)(W / 2 - 1) / (float)(W - 1);
  // alignment slack + barriers + bias[<=512] + outc[33] (+ two low-res staging slots when the upsample is fused)
  const int misc = 1024 + 1024 + 2048 + 256 + (fuse_up ? 2 * p.stg_bytes : 0);
  c.resident = !x3 && !c.small && p.num_n_tiles == 1 && c.BN <= 64 && w_bytes <= 100 * 1024 &&
               env_int("TFPNP_CONV_RESIDENT", 1) != 0;
  if (c.resident) {
    p.num_b_stages = 0;
    // 3 A stages if that lets two CTAs share an SM, else as many (<= 4) as fit one CTA
    if (3 * p.a_stage_bytes + w_bytes + misc <= 112 * 1024) p.num_a_stages = 3;
    else {
      p.num_a_stages = 4;
      while (p.num_a_stages > 2 && p.num_a_stages * p.a_stage_bytes + w_bytes + misc > 224 * 1024) --p.num_a_stages;
    }
    c.smem_bytes = p.num_a_stages * p.a_stage_bytes + w_bytes + misc;
  } else {
    p.num_a_stages = c.small ? 3 : 2;
    const int budget = 224 * 1024 - misc - p.num_a_stages * p.a_stage_bytes;
    const int tps = conv2_taps_per_stage_rt(c.BN);          // taps per ring stage (1 for BN = 128, else a kernel row)
    int sb = budget / (tps * p.b_stage_bytes);
    p.num_b_stages = sb > kMaxStages ? kMaxStages : sb;
    c.smem_bytes = p.num_a_stages * p.a_stage_bytes + p.num_b_stages * tps * p.b_stage_bytes + misc;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int occ = (c.smem_bytes <= 112 * 1024 && c.BN <= 64 && !fuse_up && kc == 32) ? 2 : 1;   // = the kernel's launch bounds
  // streamed weights: clusters of 2 (or 4) CTAs fetch each slab once and multicast it
  int cs = 1;
  if (!c.resident) {
    cs = env_int("TFPNP_CONV_CLUSTER", 2);
    while (cs > 1 && (p.num_m_tiles < cs || (c.BN / cs) * row_bytes % 1024 != 0)) cs /= 2;
  }
  p.cluster = cs;
  const int items = ((p.num_m_tiles + cs - 1) / cs) * p.num_n_tiles;
  const int max_clusters = (sms * occ) / cs;
  c.grid = (items < max_clusters ? items : max_clusters) * cs;
  return 0;
}

// 8x8 level: tensor viewed as {C, W, B, H}; box {kc, 10, 2, 10} -> smem row = y'*20 + img*10 + x'
int encode_small_map(CUtensorMap* m, const __half* base, int C, int B, int H, int W, int kc) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)B, (cuuint64_t)H};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kc, 10, 2, 10};
  return encode_map(m, const_cast<__half*>(base), 4, dims, strides, box, kc * 2);
}

int encode_halo_map(CUtensorMap* m, const __half* base, int C, int B, int H, int W, int kc) {
  if (H == 8 && W == 8) return encode_small_map(m, base, C, B, H, W, kc);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)kHaloW, (cuuint32_t)kHaloH, 1};
  return encode_map(m, const_cast<__half*>(base), 4, dims, strides, box, kc * 2);
}

struct ConvPairPlan {
  ConvPairParams p;
  bool x3 = false;
  bool x3_kc32 = false;     // conv3x3_pair_x3: 32-channel whole-chunk stages (default for split-fp16; TFPNP_PAIR_X3_KC64=1: first layout)
  int grid = 0, smem_bytes = 0;
};

// CTA-pair kernel: streamed weights, 64-channel chunks, Cout a multiple of 128, 16x16 tiling (TFPNP_CONV_PAIR=0
// falls back to the single-CTA kernel).  Measured in fp16: 14.6 / 16.6 / 39.7 us against 20.4 / 20.4 / 47.5 us (128->128 @32x32,
// 256->256 @16x16, 768->256 @16x16, 48 images); 60.3k -> 63.8k image-iters/s end to end.  The split-fp16 mode always uses it.
bool conv_pair_eligible(int C0, int C1, int Cout, int H, int W, bool x3, bool fuse_up, int B = 4) {
  if (!(x3 || env_int("TFPNP_CONV_PAIR", 1) != 0) || fuse_up || C0 % 64 != 0 || C1 % 64 != 0 || Cout % 128 != 0) return false;
  if (H == 8 && W == 8) return B >= 2 && env_int("TFPNP_PAIR_SMALL", 1) != 0 && (!x3 || env_int("TFPNP_PAIR_X3_KC64", 0) == 0);
  return H % 16 == 0 && W % 16 == 0;
}

// x*_lo / w_lo / out_lo: the fp16 residual planes (all non-null selects the split-fp16 instantiation)
int plan_conv_pair(ConvPairPlan& c, const __half* x0, const __half* x0_lo, int C0, const __half* x1, const __half* x1_lo, int C1,
                   const __half* w_taps, const __half* w_lo, const float* bias, __half* out, __half* out_lo, int B, int H, int W,
                   int Cout) {
  ConvPairParams& p = c.p;
  memset(&p, 0, sizeof(p));
  c.x3 = w_lo != nullptr;
  c.x3_kc32 = c.x3 && env_int("TFPNP_PAIR_X3_KC64", 0) == 0;
  const int kc = c.x3_kc32 ? 32 : 64;
  p.single_buf = env_int("TFPNP_PAIR_SINGLE", 0);
  p.nchunk0 = C0 / kc; p.nchunk1 = C1 / kc;
  p.small = (H == 8 && W == 8) ? 1 : 0;
  p.a_rows = p.small ? 200 : kPairARows;
  p.a_plane = (p.a_rows * kc * 2 + 1023) & ~1023;
  p.tap_rows = p.small ? 20 : 10;
  p.tiles_w = p.small ? 1 : W / 16; p.tiles_h = p.small ? 1 : H / 16;
  p.num_m_tiles = p.small ? (B + 3) / 4 : p.tiles_w * p.tiles_h * B;
  p.num_n_tiles = Cout / 128;
  p.B = B; p.H = H; p.W = W; p.Cout = Cout;
  p.bias = bias; p.out_hi = out; p.out_lo = c.x3 ? out_lo : nullptr;
  const int misc = 1024 /*align*/ + 512 /*barriers*/ + 2048 /*bias*/ + 256;
  if (!c.x3) {
    p.num_a_stages = 2;                            // chunk stages: [A half-halo 23 KB | nine half slabs 72 KB]
    p.num_b_stages = 0;
    c.smem_bytes = p.num_a_stages * (p.a_plane + kPairBStage) + misc;
  } else if (c.x3_kc32) {
    p.num_a_stages = 2;                            // chunk stages: [A_hi | A_lo half-halos 24 KB | nine slab pairs 72 KB]
    p.num_b_stages = 0;
    c.smem_bytes = p.num_a_stages * (2 * p.a_plane + 9 * 2 * kPX3Slab) + misc;
  } else {
    p.num_a_stages = 2;                            // [hi half-halo | lo half-halo] 46 KB
    p.num_b_stages = 2;                            // one kernel row of [W_hi | W_lo] half slabs, 48 KB
    c.smem_bytes = p.num_a_stages * kPairX3AStage + p.num_b_stages * kPairX3BStage + misc;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int items = p.num_m_tiles * p.num_n_tiles;
  const int pairs = items < sms / 2 ? items : sms / 2;
  c.grid = 2 * pairs;
  const __half* srcs[2][2] = {{x0, x0_lo}, {x1, x1_lo}};
  const int cs[2] = {C0, C1};
  for (int s = 0; s < 2; ++s) {
    if (!srcs[s][0]) { p.a_map[s][0] = p.a_map[0][0]; p.a_map[s][1] = p.a_map[0][1]; continue; }
    cuuint64_t dims[4] = {(cuuint64_t)cs[s], (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)cs[s] * 2, (cuuint64_t)W * cs[s] * 2, (cuuint64_t)H * W * cs[s] * 2};
    cuuint32_t box[4] = {(cuuint32_t)kc, 10, 18, 1};
    if (p.small) {          // {C, W, B, H} view, box {kc, 10, 2, 10}
      TFPNP_TRY(encode_small_map(&p.a_map[s][0], srcs[s][0], cs[s], B, H, W, kc));
      if (c.x3) TFPNP_TRY(encode_small_map(&p.a_map[s][1], srcs[s][1], cs[s], B, H, W, kc));
      else p.a_map[s][1] = p.a_map[s][0];
      continue;
    }
    TFPNP_TRY(encode_map(&p.a_map[s][0], const_cast<__half*>(srcs[s][0]), 4, dims, strides, box, 2 * kc));
    if (c.x3) TFPNP_TRY(encode_map(&p.a_map[s][1], const_cast<__half*>(srcs[s][1]), 4, dims, strides, box, 2 * kc));
    else p.a_map[s][1] = p.a_map[s][0];
  }
  const int Cin = C0 + C1;
  cuuint64_t wd[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, 9};
  cuuint64_t ws[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cin * Cout * 2};
  cuuint32_t wb[3] = {(cuuint32_t)kc, 64, 1};
  TFPNP_TRY(encode_map(&p.w_map[0], const_cast<__half*>(w_taps), 3, wd, ws, wb, 2 * kc));
  if (c.x3) TFPNP_TRY(encode_map(&p.w_map[1], const_cast<__half*>(w_lo), 3, wd, ws, wb, 2 * kc));
  else p.w_map[1] = p.w_map[0];
  return 0;
}

int launch_conv_pair(const ConvPairPlan& c, cudaStream_t st) {
  static unsigned long long attr_set = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_set >> (dev & 63) & 1ull)) {
    TFPNP_CUDA_OK(cudaFuncSetAttribute(conv3x3_pair<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    TFPNP_CUDA_OK(cudaFuncSetAttribute(conv3x3_pair<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    TFPNP_CUDA_OK(cudaFuncSetAttribute(conv3x3_pair_x3, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set |= 1ull << (dev & 63);
  }
  if (c.x3_kc32) TFPNP_CUDA_OK(launch_ex(conv3x3_pair_x3, dim3(c.grid), dim3(kPairThreads), c.smem_bytes, st, use_pdl(), 2, c.p));
  else if (c.x3) TFPNP_CUDA_OK(launch_ex(conv3x3_pair<true>, dim3(c.grid), dim3(kPairThreads), c.smem_bytes, st, use_pdl(), 2, c.p));
  else TFPNP_CUDA_OK(launch_ex(conv3x3_pair<false>, dim3(c.grid), dim3(kPairThreads), c.smem_bytes, st, use_pdl(), 2, c.p));
  TFPNP_COUNT_LAUNCH();
  return 0;
}

struct ConvWsPlan {
  ConvWsParams p;
  bool x3 = false;
  int grid = 0, smem_bytes = 0;
};

// Weight-stationary pair kernel (conv_ws.cuh): single-source layers on 16x16 tiles whose 64-cout weight slice fits next to a
// 3-deep A ring.  In the UNet: 64->128, 128->128 @ H/4 and 128->256, 256->256 @ H/8 in fp16; 64->128, 128->128 in split-fp16.
bool conv_ws_eligible(int C0, int C1, int Cout, int H, int W, bool x3) {
  // measured (round 2, graph-timed, 48 images): 16.7 / 16.8 us against 14.6 / 16.7 us of the streamed-weight pair kernel for
  // 128->128 @32^2 / 256->256 @16^2 -- with NO weight streaming at all the layer time does not move, i.e. the deep levels are
  // bound by the per-kernel fixed costs (prologue, first-tile latency, last-tile epilogue, drain: ~7 us of a ~15 us layer), not
  // by L2 -> SM weight traffic.  Kept as an experiment (TFPNP_CONV_WS=1), off by default.
  if (env_int("TFPNP_CONV_WS", 0) == 0 || C1 != 0 || Cout % 128 != 0 || H % 16 != 0 || W % 16 != 0) return false;
  const int kc = x3 ? 32 : 64;
  if (C0 % kc != 0) return false;
  const int wpair = (x3 ? 2 : 1) * 32 * kc * 2;
  const int astage = x3 ? 2 * kPX3APlane : kPairABytes;
  return 9 * (C0 / kc) * wpair + 3 * astage + 2048 <= 227 * 1024;
}

int plan_conv_ws(ConvWsPlan& c, const __half* x, const __half* x_lo, int C0, const __half* w_taps, const __half* w_lo,
                 const float* bias, __half* out, __half* out_lo, int B, int H, int W, int Cout) {
  ConvWsParams& p = c.p;
  memset(&p, 0, sizeof(p));
  c.x3 = w_lo != nullptr;
  const int kc = c.x3 ? 32 : 64;
  p.nchunks = C0 / kc;
  p.tiles_w = W / 16; p.tiles_h = H / 16;
  p.num_m_tiles = p.tiles_w * p.tiles_h * B;
  p.n_slices = Cout / 64;
  p.B = B; p.H = H; p.W = W; p.Cout = Cout;
  p.bias = bias; p.out_hi = out; p.out_lo = c.x3 ? out_lo : nullptr;
  const int wpair = (c.x3 ? 2 : 1) * 32 * kc * 2;
  const int astage = c.x3 ? 2 * kPX3APlane : kPairABytes;
  const int w_all = 9 * p.nchunks * wpair;
  int S = (227 * 1024 - 2048 - w_all) / astage;
  p.num_a_stages = S > 6 ? 6 : S;
  c.smem_bytes = p.num_a_stages * astage + w_all + 2048;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_slice = (sms / 2) / p.n_slices;                      // pairs per cout slice
  if (per_slice > p.num_m_tiles) per_slice = p.num_m_tiles;
  if (per_slice < 1) { set_error("conv_ws: %d cout slices do not fit %d SM pairs", p.n_slices, sms / 2); return TFPNP_ERR_INVALID; }
  c.grid = 2 * per_slice * p.n_slices;
  cuuint64_t dims[4] = {(cuuint64_t)C0, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C0 * 2, (cuuint64_t)W * C0 * 2, (cuuint64_t)H * W * C0 * 2};
  cuuint32_t box[4] = {(cuuint32_t)kc, 10, 18, 1};
  TFPNP_TRY(encode_map(&p.a_map[0], const_cast<__half*>(x), 4, dims, strides, box, 2 * kc));
  if (c.x3) TFPNP_TRY(encode_map(&p.a_map[1], const_cast<__half*>(x_lo), 4, dims, strides, box, 2 * kc));
  else p.a_map[1] = p.a_map[0];
  cuuint64_t wd[3] = {(cuuint64_t)C0, (cuuint64_t)Cout, 9};
  cuuint64_t ws[2] = {(cuuint64_t)C0 * 2, (cuuint64_t)C0 * Cout * 2};
  cuuint32_t wb[3] = {(cuuint32_t)kc, 32, 1};
  TFPNP_TRY(encode_map(&p.w_map[0], const_cast<__half*>(w_taps), 3, wd, ws, wb, 2 * kc));
  if (c.x3) TFPNP_TRY(encode_map(&p.w_map[1], const_cast<__half*>(w_lo), 3, wd, ws, wb, 2 * kc));
  else p.w_map[1] = p.w_map[0];
  return 0;
}

int launch_conv_ws(const ConvWsPlan& c, cudaStream_t st) {
  static unsigned long long attr_set = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_set >> (dev & 63) & 1ull)) {
    TFPNP_CUDA_OK(cudaFuncSetAttribute(conv3x3_pair_ws<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    TFPNP_CUDA_OK(cudaFuncSetAttribute(conv3x3_pair_ws<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set |= 1ull << (dev & 63);
  }
  if (c.x3) TFPNP_CUDA_OK(launch_ex(conv3x3_pair_ws<true>, dim3(c.grid), dim3(kPairThreads), c.smem_bytes, st, use_pdl(), 2, c.p));
  else TFPNP_CUDA_OK(launch_ex(conv3x3_pair_ws<false>, dim3(c.grid), dim3(kPairThreads), c.smem_bytes, st, use_pdl(), 2, c.p));
  TFPNP_COUNT_LAUNCH();
  return 0;
}

int launch_conv_params(const ConvParams& p, int BN, cudaStream_t st) {
  dim3 grid(p.tiles_w * p.tiles_h * cdiv(p.B, p.TB), p.Cout / BN);
  switch (BN) {
    case 32: TFPNP_CUDA_OK(launch_ex(conv3x3_tc<32>, grid, dim3(kConvThreads), ConvCfg<32>::kSmemBytes, st, use_pdl(), 1, p)); break;
    case 64: TFPNP_CUDA_OK(launch_ex(conv3x3_tc<64>, grid, dim3(kConvThreads), ConvCfg<64>::kSmemBytes, st, use_pdl(), 1, p)); break;
    case 128: TFPNP_CUDA_OK(launch_ex(conv3x3_tc<128>, grid, dim3(kConvThreads), ConvCfg<128>::kSmemBytes, st, use_pdl(), 1, p)); break;
    default: set_error("unsupported BN %d", BN); return TFPNP_ERR_INVALID;
  }
  TFPNP_COUNT_LAUNCH();
  return 0;
}

void tile_geom(int H, int W, int& TW, int& TH, int& TB) {
  TW = W < 8 ? W : 8;
  TH = H < kTileM / TW ? H : kTileM / TW;
  TB = kTileM / (TW * TH);
}

struct Act {            // NHWC fp16 activation tensor (hi plane, optional lo plane)
  __half* hi = nullptr;
  __half* lo = nullptr;
  int C = 0, H = 0, W = 0;
};

struct UNetTc : Denoiser {
  bool x3 = false;
  bool shares_weights = false;   // clone_shared(): weight buffers belong to the parent engine
  // weights
  DevBuf w_hi, w_lo, biases, w_out;
  FirstLayerW first_w;          // inc.conv-0 weights [32][2*9] + bias, passed by value
  size_t w_off[kNumUnetConv3];   // element offset of layer l in w_hi / w_lo ([9][Cout][Cin])
  size_t b_off[kNumUnetConv3];
  // plan for one (B,H,W)
  int pB = 0, pH = 0, pW = 0;
  DevBuf act;                    // all activation planes
  Act S0, S1, S2, skip[5];
  std::vector<ConvParams> convs; // layers 1..26 -> convs[l]
  std::vector<int> conv_bn;
  std::vector<Conv2Plan> convs2; // v2 (halo-tile) plans; grid == 0 -> layer uses v1
  std::vector<ConvPairPlan> convsp; // CTA-pair plans (default where eligible); grid == 0 -> layer uses v2 / v1
  std::vector<ConvWsPlan> convsw;   // weight-stationary pair plans (first choice where eligible)

  int init(const float* host) {
    const ConvSpec* sp = unet_conv_specs();
    // first layer + biases + outc stay fp32
    size_t off = 0, woff = 0, boff = 0, total_b = 0, total_w = 0;
    for (int l = 0; l < kNumUnetConv3; ++l) { total_b += sp[l].cout; if (l) total_w += (size_t)9 * sp[l].cout * sp[l].cin; }
    std::vector<float> hb(total_b);
    std::vector<__half> hhi(total_w), hlo(total_w);
    for (int l = 0; l < kNumUnetConv3; ++l) {
      const int ci = sp[l].cin, co = sp[l].cout;
      const float* w = host + off;
      off += (size_t)co * ci * 9;
      const float* b = host + off;
      off += co;
      b_off[l] = boff;
      for (int i = 0; i < co; ++i) hb[boff + i] = b[i];
      boff += co;
      if (l == 0) {
        for (int i = 0; i < 576; ++i) first_w.w[i] = w[i];   // [32][2][9]
        for (int i = 0; i < 32; ++i) first_w.b[i] = b[i];

        w_off[l] = 0;
        continue;
      }
      w_off[l] = woff;
      // [Cout][Cin][3][3] fp32 -> [tap][Cout][Cin] fp16 (+ residual plane)
      for (int t = 0; t < 9; ++t)
        for (int o = 0; o < co; ++o)
          for (int c = 0; c < ci; ++c) {
            float v = w[((size_t)o * ci + c) * 9 + t];
            __half h = __float2half_rn(v);
            size_t idx = woff + ((size_t)t * co + o) * ci + c;
            hhi[idx] = h;
            hlo[idx] = __float2half_rn((v - __half2float(h)) * kLoScale);
          }
      woff += (size_t)9 * co * ci;
    }
    std::vector<float> hout(33);
    for (int i = 0; i < 33; ++i) hout[i] = host[off + i];
    off += 33;
    if (off != kUnetParamCount) { set_error("unet param table mismatch"); return TFPNP_ERR_INVALID; }
    TFPNP_TRY(biases.alloc(total_b * sizeof(float)));
    TFPNP_TRY(w_out.alloc(33 * sizeof(float)));
    TFPNP_TRY(w_hi.alloc(total_w * sizeof(__half)));
    TFPNP_CUDA_OK(cudaMemcpy(biases.p, hb.data(), total_b * sizeof(float), cudaMemcpyHostToDevice));
    TFPNP_CUDA_OK(cudaMemcpy(w_out.p, hout.data(), 33 * sizeof(float), cudaMemcpyHostToDevice));
    TFPNP_CUDA_OK(cudaMemcpy(w_hi.p, hhi.data(), total_w * sizeof(__half), cudaMemcpyHostToDevice));
    if (x3) {
      TFPNP_TRY(w_lo.alloc(total_w * sizeof(__half)));
      TFPNP_CUDA_OK(cudaMemcpy(w_lo.p, hlo.data(), total_w * sizeof(__half), cudaMemcpyHostToDevice));
    }
    return set_conv_attrs();
  }

  int make_act_maps(CUtensorMap (&maps)[2], const Act& a, int B, int kc, int TW, int TH, int TB) {
    cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)a.C * 2, (cuuint64_t)a.W * a.C * 2, (cuuint64_t)a.H * a.W * a.C * 2};
    cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)(TB < B ? TB : B)};
    TFPNP_TRY(encode_map(&maps[0], a.hi, 4, dims, strides, box, kc * 2));
    if (x3) TFPNP_TRY(encode_map(&maps[1], a.lo, 4, dims, strides, box, kc * 2));
    else maps[1] = maps[0];
    return 0;
  }

  // per-launch timing (tfpnp_denoiser_layer_profile): an event after every launch of forward()
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<std::string> prof_names;
  size_t prof_n = 0;
  void mark(cudaStream_t st, const char* name) {
    if (!prof_on) return;
    if (prof_n == prof_ev.size()) { cudaEvent_t e; cudaEventCreate(&e); prof_ev.push_back(e); prof_names.push_back(name); }
    cudaEventRecord(prof_ev[prof_n++], st);
  }
  std::string conv_label(int l) const {
    const ConvSpec& sp = unet_conv_specs()[l];
    const char* kind = "v1";
    int h = 0;
    if (convsw[l].grid > 0) { kind = convsw[l].x3 ? "pair-ws-x3" : "pair-ws"; h = convsw[l].p.H; }
    else if (convsp[l].grid > 0) { kind = convsp[l].x3_kc32 ? "pair-x3/32" : convsp[l].x3 ? "pair-x3/64" : "pair"; h = convsp[l].p.H; }
    else if (convs2[l].grid > 0) {
      const Conv2Plan& c = convs2[l];
      h = c.p.H;
      kind = c.x3n ? (c.resident ? "x3n-res" : "x3n") : c.small ? "small" : c.p.up_fused ? "fuse-up" : c.resident ? "tc2-res" : "tc2";
    } else h = convs[l].H;
    char buf[48];
    snprintf(buf, sizeof(buf), "l%02d %d->%d @%d %s", l, sp.cin, sp.cout, h, kind);
    return buf;
  }

  int layer_profile(const float* x, const float* sigma, float* out, int B, int H, int W, int reps, float* ms_out, int cap,
                    int* n_out, char* names, cudaStream_t st) override {
    g_no_pdl = true;
    prof_on = true;
    std::vector<double> acc;
    int rc = 0;
    for (int r = 0; r < reps + 1 && rc == 0; ++r) {        // first repetition is a warm-up
      // an un-timed call first, so that the timed launches are enqueued behind a busy GPU (no host launch gaps in the intervals)
      prof_on = false;
      rc = forward(x, sigma, 1, out, B, H, W, st);
      prof_on = true;
      if (rc != 0) break;
      prof_n = 0;
      mark(st, "start");
      rc = forward(x, sigma, 1, out, B, H, W, st);
      if (rc != 0) break;
      if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("layer_profile: sync failed"); rc = TFPNP_ERR_CUDA; break; }
      if (acc.empty()) acc.assign(prof_n - 1, 0.0);
      if (r == 0) continue;
      for (size_t i = 1; i < prof_n; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, prof_ev[i - 1], prof_ev[i]);
        acc[i - 1] += ms;
      }
    }
    prof_on = false;
    g_no_pdl = false;
    if (rc != 0) return rc;
    const int n = (int)acc.size() < cap ? (int)acc.size() : cap;
    for (int i = 0; i < n; ++i) {
      ms_out[i] = (float)(acc[i] / reps);
      snprintf(names + 48 * i, 48, "%s", prof_names[i + 1].c_str());
    }
    *n_out = n;
    return 0;
  }

  bool fused_pool[kNumUnetConv3] = {};   // layer l also wrote its 2x2-max-pooled output (into S2)
  bool fused_outc = false;               // layer 26 produced x directly

  bool fused_up[kNumUnetConv3] = {};     // decoder conv-0 layer l interpolates its second source itself

  // `low`: when non-null, source 1 is the bilinear x2 up-sampling of this low-resolution tensor, fused into the layer
  int plan_conv_v2(int l, const Act& s0, const Act* s1, const Act& dst, int B, const Act* low = nullptr) {
    const ConvSpec& sp = unet_conv_specs()[l];
    Conv2Plan& c = convs2[l];
    memset(&c.p, 0, sizeof(c.p));
    const int c1 = s1 ? s1->C : 0;
    fused_up[l] = low != nullptr;
    TFPNP_TRY(plan_conv2_geometry(c, s0.C, c1, sp.cout, B, dst.H, dst.W, x3, low != nullptr));
    Conv2Params& p = c.p;
    p.bias = biases.as<float>() + b_off[l];
    p.out_hi = dst.hi;
    p.out_lo = x3 ? dst.lo : nullptr;
    const bool fuse = env_int("TFPNP_CONV_FUSE", 1) != 0;
    fused_pool[l] = fuse && !c.small && (l == 2 || l == 5 || l == 8 || l == 11);   // conv-2 of inc / down1..3 feeds a MaxPool2d
    if (fused_pool[l]) { p.pool_hi = S2.hi; p.pool_lo = x3 ? S2.lo : nullptr; }   // S2 is idle in the encoder
    if (l == 26) {
      fused_outc = fuse;
      if (fuse) p.outc_w = w_out.as<float>();      // d_in / x_out are per-call pointers, set in forward()
    }
    const Act* srcs[2] = {&s0, s1};
    for (int s = 0; s < 2; ++s) {
      if (!srcs[s]) { p.a_map[s][0] = p.a_map[0][0]; p.a_map[s][1] = p.a_map[0][1]; continue; }
      if (s == 1 && low) {       // unswizzled 11x11 window of the low-resolution tensor
        cuuint64_t dims[4] = {(cuuint64_t)low->C, (cuuint64_t)low->W, (cuuint64_t)low->H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)low->C * 2, (cuuint64_t)low->W * low->C * 2,
                                 (cuuint64_t)low->H * low->W * low->C * 2};
        cuuint32_t box[4] = {(cuuint32_t)c.kc, (cuuint32_t)kUpBox, (cuuint32_t)kUpBox, 1};
        TFPNP_TRY(encode_map(&p.a_map[1][0], low->hi, 4, dims, strides, box, -1));
        if (x3) TFPNP_TRY(encode_map(&p.a_map[1][1], low->lo, 4, dims, strides, box, -1));
        else p.a_map[1][1] = p.a_map[1][0];
        continue;
      }
      TFPNP_TRY(encode_halo_map(&p.a_map[s][0], srcs[s]->hi, srcs[s]->C, B, dst.H, dst.W, c.kc));
      if (x3) TFPNP_TRY(encode_halo_map(&p.a_map[s][1], srcs[s]->lo, srcs[s]->C, B, dst.H, dst.W, c.kc));
      else p.a_map[s][1] = p.a_map[s][0];
    }
    cuuint64_t wd[3] = {(cuuint64_t)sp.cin, (cuuint64_t)sp.cout, 9};
    cuuint64_t ws[2] = {(cuuint64_t)sp.cin * 2, (cuuint64_t)sp.cin * sp.cout * 2};
    cuuint32_t wb[3] = {(cuuint32_t)c.kc, (cuuint32_t)(c.BN / p.cluster), 1};   // each cluster CTA fetches BN/cluster rows
    TFPNP_TRY(encode_map(&p.w_map[0], w_hi.as<__half>() + w_off[l], 3, wd, ws, wb, c.kc * 2));
    if (x3) TFPNP_TRY(encode_map(&p.w_map[1], w_lo.as<__half>() + w_off[l], 3, wd, ws, wb, c.kc * 2));
    else p.w_map[1] = p.w_map[0];
    return 0;
  }

  // build ConvParams for layer l reading (src0 [, src1]) and writing dst
  int plan_conv(int l, const Act& s0, const Act* s1, const Act& dst, int B, const Act* low = nullptr) {
    convsp[l].grid = 0;
    convsw[l].grid = 0;
    if (!low && conv_ws_eligible(s0.C, s1 ? s1->C : 0, unet_conv_specs()[l].cout, dst.H, dst.W, x3)) {
      ConvWsPlan& c = convsw[l];
      TFPNP_TRY(plan_conv_ws(c, s0.hi, x3 ? s0.lo : nullptr, s0.C, w_hi.as<__half>() + w_off[l],
                             x3 ? w_lo.as<__half>() + w_off[l] : nullptr, biases.as<float>() + b_off[l], dst.hi,
                             x3 ? dst.lo : nullptr, B, dst.H, dst.W, unet_conv_specs()[l].cout));
      convs2[l].grid = 0;
      fused_up[l] = false;
      fused_pool[l] = env_int("TFPNP_CONV_FUSE", 1) != 0 && (l == 2 || l == 5 || l == 8 || l == 11);
      if (fused_pool[l]) { c.p.pool_hi = S2.hi; c.p.pool_lo = x3 ? S2.lo : nullptr; }
      return 0;
    }
    if (conv_pair_eligible(s0.C, s1 ? s1->C : 0, unet_conv_specs()[l].cout, dst.H, dst.W, x3, low != nullptr, B)) {
      ConvPairPlan& c = convsp[l];
      TFPNP_TRY(plan_conv_pair(c, s0.hi, x3 ? s0.lo : nullptr, s0.C, s1 ? s1->hi : nullptr, (s1 && x3) ? s1->lo : nullptr,
                               s1 ? s1->C : 0, w_hi.as<__half>() + w_off[l], x3 ? w_lo.as<__half>() + w_off[l] : nullptr,
                               biases.as<float>() + b_off[l], dst.hi, x3 ? dst.lo : nullptr, B, dst.H, dst.W,
                               unet_conv_specs()[l].cout));
      convs2[l].grid = 0;
      fused_up[l] = false;
      // (8x8 tiles interleave two images per row group: the epilogue's lane-shuffle pooling does not apply there)
      fused_pool[l] = env_int("TFPNP_CONV_FUSE", 1) != 0 && !c.p.small && (l == 2 || l == 5 || l == 8 || l == 11);
      if (fused_pool[l]) { c.p.pool_hi = S2.hi; c.p.pool_lo = x3 ? S2.lo : nullptr; }
      return 0;
    }
    if (conv2_eligible(dst.H, dst.W) ||
        (!low && conv2_small_eligible(dst.H, dst.W, B, s0.C, s1 ? s1->C : 0, unet_conv_specs()[l].cout)))
      return plan_conv_v2(l, s0, s1, dst, B, low);
    convs2[l].grid = 0;
    fused_pool[l] = false;
    fused_up[l] = false;
    if (l == 26) fused_outc = false;
    const ConvSpec& sp = unet_conv_specs()[l];
    ConvParams& p = convs[l];
    memset(&p, 0, sizeof(p));
    const int c1 = s1 ? s1->C : 0;
    TFPNP_CHECK(s0.C + c1 == sp.cin && dst.C == sp.cout, "plan_conv %d: channel mismatch", l);
    const int kc = (s0.C % 64 == 0 && c1 % 64 == 0) ? 64 : 32;
    const int BN = sp.cout >= 128 ? 128 : sp.cout;
    conv_bn[l] = BN;
    p.kc = kc;
    p.nchunk0 = s0.C / kc;
    p.nchunk1 = c1 / kc;
    p.nprod = x3 ? 3 : 1;
    tile_geom(dst.H, dst.W, p.TW, p.TH, p.TB);
    p.box_rows = p.TW * p.TH * (p.TB < B ? p.TB : B);
    p.tiles_w = dst.W / p.TW;
    p.tiles_h = dst.H / p.TH;
    p.B = B; p.H = dst.H; p.W = dst.W; p.Cout = sp.cout;
    p.dil = 1; p.slope = 0.2f;
    p.bias = biases.as<float>() + b_off[l];
    p.out_hi = dst.hi;
    p.out_lo = x3 ? dst.lo : nullptr;
    TFPNP_TRY(make_act_maps(p.a_map[0], s0, B, kc, p.TW, p.TH, p.TB));
    if (s1) TFPNP_TRY(make_act_maps(p.a_map[1], *s1, B, kc, p.TW, p.TH, p.TB));
    else { p.a_map[1][0] = p.a_map[0][0]; p.a_map[1][1] = p.a_map[0][1]; }
    cuuint64_t wd[3] = {(cuuint64_t)sp.cin, (cuuint64_t)sp.cout, 9};
    cuuint64_t ws[2] = {(cuuint64_t)sp.cin * 2, (cuuint64_t)sp.cin * sp.cout * 2};
    cuuint32_t wb[3] = {(cuuint32_t)kc, (cuuint32_t)BN, 1};
    TFPNP_TRY(encode_map(&p.w_map[0], w_hi.as<__half>() + w_off[l], 3, wd, ws, wb, kc * 2));
    if (x3) TFPNP_TRY(encode_map(&p.w_map[1], w_lo.as<__half>() + w_off[l], 3, wd, ws, wb, kc * 2));
    else p.w_map[1] = p.w_map[0];
    return 0;
  }

  int prepare(int B, int H, int W) override {
    TFPNP_CHECK(H % 16 == 0 && W % 16 == 0 && H >= 16 && W >= 16, "UNet needs H,W multiples of 16, got %dx%d", H, W);
    if (B == pB && H == pH && W == pW) return 0;
    const size_t HW = (size_t)H * W;
    const int planes = x3 ? 2 : 1;
    // elements per image: S0 64 | S1 32 | S2 32 | x1 32 | x2 16 | x3 8 | x4 4 | x5 2  (units of HW)
    const size_t per_plane = (size_t)(64 + 32 + 32 + 32 + 16 + 8 + 4 + 2) * HW * B;
    const void* before = act.p;
    TFPNP_TRY(act.alloc(per_plane * planes * sizeof(__half)));
    if (act.p != before) ++generation;
    __half* base = act.as<__half>();
    size_t cur = 0;
    auto carve = [&](Act& a, size_t units) {
      a.hi = base + cur;
      a.lo = x3 ? base + per_plane + cur : nullptr;
      cur += units * HW * B;
    };
    carve(S0, 64); carve(S1, 32); carve(S2, 32);
    const int ch[5] = {32, 64, 128, 256, 512};
    const size_t units[5] = {32, 16, 8, 4, 2};
    for (int i = 0; i < 5; ++i) { carve(skip[i], units[i]); skip[i].C = ch[i]; skip[i].H = H >> i; skip[i].W = W >> i; }
    convs.assign(kNumUnetConv3, ConvParams{});
    conv_bn.assign(kNumUnetConv3, 0);
    convs2.assign(kNumUnetConv3, Conv2Plan{});
    convsp.assign(kNumUnetConv3, ConvPairPlan{});
    convsw.assign(kNumUnetConv3, ConvWsPlan{});
    auto view = [](const Act& buf, int C, int h, int w) { Act a = buf; a.C = C; a.H = h; a.W = w; return a; };
    // encoder level 0: first -> S0 ; conv1: S0 -> S1 ; conv2: S1 -> x1
    TFPNP_TRY(plan_conv(1, view(S0, 32, H, W), nullptr, view(S1, 32, H, W), B));
    TFPNP_TRY(plan_conv(2, view(S1, 32, H, W), nullptr, skip[0], B));
    for (int lv = 1; lv <= 4; ++lv) {
      int h = H >> lv, w = W >> lv, l0 = 3 * lv;
      // pooled input: written by the fused epilogue of the previous block (-> S2) or by maxpool2_nhwc (-> S0)
      const Act& pooled = fused_pool[l0 - 1] ? S2 : S0;
      TFPNP_TRY(plan_conv(l0, view(pooled, ch[lv - 1], h, w), nullptr, view(S1, ch[lv], h, w), B));
      TFPNP_TRY(plan_conv(l0 + 1, view(S1, ch[lv], h, w), nullptr, view(S0, ch[lv], h, w), B));
      TFPNP_TRY(plan_conv(l0 + 2, view(S0, ch[lv], h, w), nullptr, skip[lv], B));
    }
    for (int k = 0; k < 4; ++k) {
      int lv = 3 - k, h = H >> lv, w = W >> lv, l0 = 15 + 3 * k;
      Act up = view(S0, ch[lv + 1], h, w);
      // the block input: x5 for the first block, else the previous block's output in S2
      Act low = view(k == 0 ? skip[4] : S2, ch[lv + 1], h / 2, w / 2);
      // fused up-sampling pays where the up-sampled tensor is large; TFPNP_FUSE_UP_MIN = smallest output height fused
      // (split-fp16: the two heads that run on conv3x3_x3, 96 -> 32 and 192 -> 64)
      const bool fuse_up = conv2_eligible(h, w) && env_int("TFPNP_CONV_FUSE_UP", 1) != 0 &&
                           (x3 ? (ch[lv] <= env_int("TFPNP_X3_FUSE_MAXCH", 64) && ch[lv] >= env_int("TFPNP_X3_FUSE_MINCH", 32) && env_int("TFPNP_X3_FUSE_UP", 1) != 0)
                               : h >= env_int("TFPNP_FUSE_UP_MIN", 64));   // measured (fp16, round 2, row-staged up-sampling kernel):
                                                                              // 384->128 @32^2 un-fused on the pair kernel 14.7 + 40.8 us
                                                                              // against 68.8 us fused; 16x16 outputs: un-fused since round 1
      TFPNP_TRY(plan_conv(l0, skip[lv], &up, view(S1, ch[lv], h, w), B, fuse_up ? &low : nullptr));
      TFPNP_TRY(plan_conv(l0 + 1, view(S1, ch[lv], h, w), nullptr, view(S0, ch[lv], h, w), B));
      TFPNP_TRY(plan_conv(l0 + 2, view(S0, ch[lv], h, w), nullptr, view(S2, ch[lv], h, w), B));
    }
    pB = B; pH = H; pW = W;
    return 0;
  }

  int launch_conv(int l, cudaStream_t st) {
    const int rc = launch_conv_impl(l, st);
    if (prof_on && rc == 0) mark(st, conv_label(l).c_str());
    return rc;
  }
  int launch_conv_impl(int l, cudaStream_t st) {
    if (convsw[l].grid > 0) return launch_conv_ws(convsw[l], st);
    if (convsp[l].grid > 0) return launch_conv_pair(convsp[l], st);
    if (convs2[l].grid > 0) {
      // debugging aid: TFPNP_TRACE_LAYER=l + TFPNP_TRACE_FILE dump CTA 0's role timeline of layer l (eager calls only)
      static const int trace_layer = env_int("TFPNP_TRACE_LAYER", -1);
      const char* tf = trace_layer == l ? getenv("TFPNP_TRACE_FILE") : nullptr;
      if (tf) {
        unsigned long long* dtrace = nullptr;
        TFPNP_CUDA_OK(cudaMalloc(&dtrace, 8 * 1024 * 8));
        TFPNP_CUDA_OK(cudaMemset(dtrace, 0, 8 * 1024 * 8));
        Conv2Plan c = convs2[l];
        c.p.trace = dtrace;
        TFPNP_TRY(launch_conv2(c, st));
        std::vector<unsigned long long> h(8 * 1024);
        TFPNP_CUDA_OK(cudaStreamSynchronize(st));
        TFPNP_CUDA_OK(cudaMemcpy(h.data(), dtrace, h.size() * 8, cudaMemcpyDeviceToHost));
        cudaFree(dtrace);
        FILE* f = fopen(tf, "wb");
        if (f) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
        return 0;
      }
      return launch_conv2(convs2[l], st);
    }
    return launch_conv_params(convs[l], conv_bn[l], st);
  }

  int forward(const float* x, const float* sigma, int64_t sstride, float* out, int B, int H, int W,
              cudaStream_t st) override {
    TFPNP_CHECK(B == pB && H == pH && W == pW, "prepare(%d,%d,%d) not called (plan is %d,%d,%d)", B, H, W, pB, pH, pW);
    const int ch[5] = {32, 64, 128, 256, 512};
    const int T = 256;
    // head of the chain: a normal (fully serialised) launch; every later kernel of this call may overlap its
    // prologue with its predecessor's tail (PDL)
    TFPNP_CUDA_OK(launch_ex(conv_first_kernel, dim3(cdiv(W, 128), H, B), dim3(128), 0, st, use_pdl(), 1, x, sigma, sstride,
                            first_w, S0.hi, x3 ? S0.lo : nullptr, H, W));
    TFPNP_COUNT_LAUNCH();
    mark(st, "l00 2->32 first (CUDA cores)");
    TFPNP_TRY(launch_conv(1, st));
    TFPNP_TRY(launch_conv(2, st));
    for (int lv = 1; lv <= 4; ++lv) {
      int h = H >> lv, w = W >> lv;
      if (!fused_pool[3 * lv - 1]) {
        TFPNP_CUDA_OK(launch_ex(maxpool2_nhwc, dim3(cdiv(w * (ch[lv - 1] / 8), T), h, B), dim3(T), 0, st, use_pdl(), 1,
                                skip[lv - 1].hi, skip[lv - 1].lo, S0.hi, S0.lo, 2 * h, 2 * w, ch[lv - 1]));
        TFPNP_COUNT_LAUNCH();
        mark(st, "maxpool");
      }
      for (int k = 0; k < 3; ++k) TFPNP_TRY(launch_conv(3 * lv + k, st));
    }
    for (int k = 0; k < 4; ++k) {
      int lv = 3 - k, h = H >> lv, w = W >> lv;
      const Act& src = k == 0 ? skip[4] : S2;
      if (!fused_up[15 + 3 * k]) {
        const int up_smem = (x3 ? 4 : 2) * (w / 2) * ch[lv + 1] * 2;     // [plane][2 rows][W_in * C] fp16
        TFPNP_CHECK(up_smem <= 200 * 1024, "upsample: a source row pair of %d bytes does not fit shared memory", up_smem);
        TFPNP_CUDA_OK(launch_ex(x3 ? upsample2_nhwc<true> : upsample2_nhwc<false>, dim3(h, B), dim3(T), up_smem, st, use_pdl(), 1,
                                src.hi, src.lo, S0.hi, S0.lo, h / 2, w / 2, ch[lv + 1]));
      }
      if (!fused_up[15 + 3 * k]) { TFPNP_COUNT_LAUNCH(); char nb[48]; snprintf(nb, sizeof(nb), "upsample %d ch -> @%d", ch[lv + 1], h); mark(st, nb); }
      for (int j = 0; j < 3; ++j) {
        const int l = 15 + 3 * k + j;
        if (l == 26 && fused_outc) { convs2[l].p.d_in = x; convs2[l].p.x_out = out; }
        TFPNP_TRY(launch_conv(l, st));
      }
    }
    if (!fused_outc) {
      size_t npix = (size_t)B * H * W;
      TFPNP_CUDA_OK(launch_ex(outc_nhwc, dim3((unsigned)((npix + T - 1) / T)), dim3(T), 0, st, use_pdl(), 1, S2.hi, S2.lo,
                              w_out.as<float>(), w_out.as<float>() + 32, x, out, npix));
      TFPNP_COUNT_LAUNCH();
      mark(st, "outconv");
    }
    TFPNP_CUDA_OK(cudaGetLastError());
    return 0;
  }

  Denoiser* clone_shared() override {
    UNetTc* c = new UNetTc();
    c->precision = precision; c->x3 = x3; c->shares_weights = true;
    c->w_hi = w_hi; c->w_lo = w_lo; c->biases = biases; c->w_out = w_out;   // non-owning aliases
    c->first_w = first_w;
    memcpy(c->w_off, w_off, sizeof(w_off));
    memcpy(c->b_off, b_off, sizeof(b_off));
    return c;
  }

  ~UNetTc() override {
    if (shares_weights) { w_hi.p = w_lo.p = biases.p = w_out.p = nullptr; }
    w_hi.release(); w_lo.release(); biases.release(); w_out.release(); act.release();
    for (cudaEvent_t e : prof_ev) cudaEventDestroy(e);
  }
};

// One 3x3 conv (pad 1) + bias + LeakyReLU(0.2) on NHWC fp16 tensors, FP16 mode: the kernel-level
// entry point behind tfpnp_conv3x3_nhwc (per-kernel parity tests).
int conv3x3_nhwc_standalone(const __half* x0, int C0, const __half* x1, int C1, const __half* w_taps,
                            const float* bias, __half* out, int B, int H, int W, int Cout, cudaStream_t st) {
  TFPNP_CHECK(C0 > 0 && C0 % 32 == 0 && C1 % 32 == 0, "channel counts must be multiples of 32");
  TFPNP_CHECK(Cout == 32 || Cout == 64 || Cout % 128 == 0, "Cout must be 32, 64 or a multiple of 128");
  TFPNP_TRY(set_conv_attrs());
  if (conv_ws_eligible(C0, C1, Cout, H, W, false)) {
    ConvWsPlan cw;
    TFPNP_TRY(plan_conv_ws(cw, x0, nullptr, C0, w_taps, nullptr, bias, out, nullptr, B, H, W, Cout));
    TFPNP_TRY(launch_conv_ws(cw, st));
    TFPNP_CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (conv_pair_eligible(C0, C1, Cout, H, W, false, false, B)) {
    ConvPairPlan cp;
    TFPNP_TRY(plan_conv_pair(cp, x0, nullptr, C0, x1, nullptr, C1, w_taps, nullptr, bias, out, nullptr, B, H, W, Cout));
    TFPNP_TRY(launch_conv_pair(cp, st));
    TFPNP_CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (conv2_eligible(H, W) || conv2_small_eligible(H, W, B, C0, C1, Cout)) {
    Conv2Plan c;
    memset(&c.p, 0, sizeof(c.p));
    TFPNP_TRY(plan_conv2_geometry(c, C0, C1, Cout, B, H, W, false));
    Conv2Params& q = c.p;
    q.bias = bias; q.out_hi = out; q.out_lo = nullptr;
    TFPNP_TRY(encode_halo_map(&q.a_map[0][0], x0, C0, B, H, W, c.kc));
    q.a_map[0][1] = q.a_map[0][0];
    if (x1) TFPNP_TRY(encode_halo_map(&q.a_map[1][0], x1, C1, B, H, W, c.kc));
    else q.a_map[1][0] = q.a_map[0][0];
    q.a_map[1][1] = q.a_map[1][0];
    cuuint64_t wd2[3] = {(cuuint64_t)(C0 + C1), (cuuint64_t)Cout, 9};
    cuuint64_t ws2[2] = {(cuuint64_t)(C0 + C1) * 2, (cuuint64_t)(C0 + C1) * Cout * 2};
    cuuint32_t wb2[3] = {(cuuint32_t)c.kc, (cuuint32_t)(c.BN / q.cluster), 1};
    TFPNP_TRY(encode_map(&q.w_map[0], const_cast<__half*>(w_taps), 3, wd2, ws2, wb2, c.kc * 2));
    q.w_map[1] = q.w_map[0];
    const char* tf = getenv("TFPNP_TRACE_FILE");
    unsigned long long* dtrace = nullptr;
    if (tf) {
      TFPNP_CUDA_OK(cudaMalloc(&dtrace, 8 * 1024 * 8));
      TFPNP_CUDA_OK(cudaMemset(dtrace, 0, 8 * 1024 * 8));
      q.trace = dtrace;
    }
    TFPNP_TRY(launch_conv2(c, st));
    TFPNP_CUDA_OK(cudaGetLastError());
    if (tf) {
      std::vector<unsigned long long> h(8 * 1024);
      TFPNP_CUDA_OK(cudaStreamSynchronize(st));
      TFPNP_CUDA_OK(cudaMemcpy(h.data(), dtrace, h.size() * 8, cudaMemcpyDeviceToHost));
      cudaFree(dtrace);
      FILE* f = fopen(tf, "wb");
      if (f) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
    }
    return 0;
  }
  ConvParams p;
  memset(&p, 0, sizeof(p));
  const int Cin = C0 + C1;
  const int kc = (C0 % 64 == 0 && C1 % 64 == 0) ? 64 : 32;
  const int BN = Cout >= 128 ? 128 : Cout;
  p.kc = kc; p.nchunk0 = C0 / kc; p.nchunk1 = C1 / kc; p.nprod = 1;
  tile_geom(H, W, p.TW, p.TH, p.TB);
  TFPNP_CHECK(W % p.TW == 0 && H % p.TH == 0, "H, W must tile by %dx%d", p.TH, p.TW);
  p.box_rows = p.TW * p.TH * (p.TB < B ? p.TB : B);
  p.tiles_w = W / p.TW; p.tiles_h = H / p.TH;
  p.B = B; p.H = H; p.W = W; p.Cout = Cout;
  p.dil = 1; p.slope = 0.2f;
  p.bias = bias; p.out_hi = out; p.out_lo = nullptr;
  const __half* srcs[2] = {x0, x1};
  const int cs[2] = {C0, C1};
  for (int s = 0; s < 2; ++s) {
    if (!srcs[s]) { p.a_map[s][0] = p.a_map[0][0]; p.a_map[s][1] = p.a_map[0][0]; continue; }
    cuuint64_t dims[4] = {(cuuint64_t)cs[s], (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)cs[s] * 2, (cuuint64_t)W * cs[s] * 2, (cuuint64_t)H * W * cs[s] * 2};
    cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)(p.TB < B ? p.TB : B)};
    TFPNP_TRY(encode_map(&p.a_map[s][0], const_cast<__half*>(srcs[s]), 4, dims, strides, box, kc * 2));
    p.a_map[s][1] = p.a_map[s][0];
  }
  cuuint64_t wd[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, 9};
  cuuint64_t ws[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cin * Cout * 2};
  cuuint32_t wb[3] = {(cuuint32_t)kc, (cuuint32_t)BN, 1};
  TFPNP_TRY(encode_map(&p.w_map[0], const_cast<__half*>(w_taps), 3, wd, ws, wb, kc * 2));
  p.w_map[1] = p.w_map[0];
  TFPNP_TRY(launch_conv_params(p, BN, st));
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

// ---- one-tile-per-CTA tensor-core conv as a reusable, pre-planned layer (used by ircnn.cu) -------------
struct ConvV1Layer { ConvParams p; int BN; };

int conv_v1_plan(ConvV1Layer** out, const __half* x_hi, const __half* x_lo, int Cin, const __half* w_hi,
                 const __half* w_lo, const float* bias, __half* out_hi, __half* out_lo, int B, int H, int W,
                 int Cout, int dil, float slope) {
  TFPNP_CHECK(Cin % 32 == 0 && (Cout == 32 || Cout == 64 || Cout % 128 == 0), "conv_v1_plan: unsupported channels %d -> %d", Cin, Cout);
  TFPNP_TRY(set_conv_attrs());
  const bool x3 = x_lo != nullptr;
  ConvV1Layer* L = new ConvV1Layer();
  ConvParams& p = L->p;
  memset(&p, 0, sizeof(p));
  const int kc = Cin % 64 == 0 ? 64 : 32;
  L->BN = Cout >= 128 ? 128 : Cout;
  p.kc = kc; p.nchunk0 = Cin / kc; p.nchunk1 = 0; p.nprod = x3 ? 3 : 1;
  tile_geom(H, W, p.TW, p.TH, p.TB);
  if (W % p.TW != 0 || H % p.TH != 0) {
    set_error("conv_v1_plan: H, W must tile by %dx%d", p.TH, p.TW);
    delete L;
    return TFPNP_ERR_INVALID;
  }
  p.box_rows = p.TW * p.TH * (p.TB < B ? p.TB : B);
  p.tiles_w = W / p.TW; p.tiles_h = H / p.TH;
  p.B = B; p.H = H; p.W = W; p.Cout = Cout;
  p.dil = dil; p.slope = slope;
  p.bias = bias; p.out_hi = out_hi; p.out_lo = x3 ? out_lo : nullptr;
  cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)(p.TB < B ? p.TB : B)};
  int rc = encode_map(&p.a_map[0][0], const_cast<__half*>(x_hi), 4, dims, strides, box, kc * 2);
  if (rc == 0 && x3) rc = encode_map(&p.a_map[0][1], const_cast<__half*>(x_lo), 4, dims, strides, box, kc * 2);
  if (!x3) p.a_map[0][1] = p.a_map[0][0];
  p.a_map[1][0] = p.a_map[0][0]; p.a_map[1][1] = p.a_map[0][1];
  cuuint64_t wd[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, 9};
  cuuint64_t ws[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cin * Cout * 2};
  cuuint32_t wb[3] = {(cuuint32_t)kc, (cuuint32_t)L->BN, 1};
  if (rc == 0) rc = encode_map(&p.w_map[0], const_cast<__half*>(w_hi), 3, wd, ws, wb, kc * 2);
  if (rc == 0 && x3) rc = encode_map(&p.w_map[1], const_cast<__half*>(w_lo), 3, wd, ws, wb, kc * 2);
  if (!x3) p.w_map[1] = p.w_map[0];
  if (rc != 0) { delete L; return rc; }
  *out = L;
  return 0;
}
int conv_v1_launch(const ConvV1Layer* L, cudaStream_t st) {
  TFPNP_TRY(launch_conv_params(L->p, L->BN, st));
  return 0;
}
void conv_v1_free(ConvV1Layer* L) { delete L; }

int conv3x3_nhwc(const void* x0, int C0, const void* x1, int C1, const void* w_taps, const float* bias,
                 void* out, int B, int H, int W, int Cout, cudaStream_t st) {
  return conv3x3_nhwc_standalone(static_cast<const __half*>(x0), C0, static_cast<const __half*>(x1), C1,
                                 static_cast<const __half*>(w_taps), bias, static_cast<__half*>(out), B, H, W,
                                 Cout, st);
}

Denoiser* make_unet_tc(const float* weights_host, int precision) {
  UNetTc* u = new UNetTc();
  u->precision = precision;
  u->x3 = precision == TFPNP_PREC_FP16X3;
  if (u->init(weights_host) != 0) { delete u; return nullptr; }
  return u;
}

}  // namespace tfpnp
