// Weight-stationary CTA-pair convolution for the C -> C layers of the 128- and 256-channel levels (included by unet_tc.cu).
//
// The streamed-weight pair kernels fetch 72 KB of weight slabs per 64-channel chunk per CTA for 1.2 us of MMAs: 12 TB/s
// chip-wide out of L2 in fp16 -- the deep levels ran at ~40 % of their MMA bound because of it (layer profile, round 2).
// Here a CTA pair OWNS a 64-cout slice of the layer for its whole lifetime: CTA r keeps rows [64 s + 32 r, +32) of every
// (tap, chunk) slab resident in shared memory (72 KB at 128 channels, 144 KB at 256 in fp16; 144 KB at 128 channels in
// split-fp16), fetched once before the dependency wait, and the pair walks the 16x16 super-tiles of the layer with only the
// half-halo A tile (23 KB per chunk per CTA) streaming through a 3-4 deep ring.  tcgen05.mma.cta_group::2 with M = 256,
// N = 64 (43 cycles per MMA instead of the ideal 32: 75 % of the tensor rate, against 40 % with streamed weights), and the
// work unit shrinks to (tile, 64 couts), which also balances the 16x16 level (192 units on 72 pairs instead of 96 on 74).
// Split-fp16 (X3): 32-channel chunks, stage = [A_hi | A_lo]; main product in TMEM columns [0, 64), the two correction
// products (residual planes are x 2^11) in [64, 128).
struct ConvWsParams {
  CUtensorMap a_map[2];     // [plane] half-halo boxes {KC, 10, 18, 1}
  CUtensorMap w_map[2];     // [plane] {KC, 32 rows, 1 tap} over {Cin, Cout, 9}
  int nchunks;
  int tiles_w, tiles_h, num_m_tiles, n_slices;
  int B, H, W, Cout;
  int num_a_stages;
  const float* bias;
  __half* out_hi;
  __half* out_lo;
  __half* pool_hi;          // fused nn.MaxPool2d(2) output or nullptr
  __half* pool_lo;
};

template <bool X3>
__global__ void __launch_bounds__(kPairThreads, 1)
conv3x3_pair_ws(const __grid_constant__ ConvWsParams p) {
  constexpr int KC = X3 ? 32 : 64, KSTEPS = KC / 16;
  constexpr uint32_t ROW = KC * 2;
  constexpr int BNS = 64;                                       // couts per pair
  constexpr uint32_t APLANE = X3 ? kPX3APlane : kPairABytes;    // one half-halo plane (180 rows, padded to 1 KB)
  constexpr uint32_t ASTAGE = X3 ? 2 * APLANE : APLANE;
  constexpr uint32_t WSLAB = 32 * ROW;                          // this CTA's 32 rows of one (tap, chunk) slab, one plane
  constexpr uint32_t WPAIR = X3 ? 2 * WSLAB : WSLAB;            // [W_hi | W_lo]
  constexpr uint32_t kTmem = X3 ? 256 : 128;                    // two buffers of [main (| corrections)]
  constexpr uint32_t kBufCols = X3 ? 2 * BNS : BNS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.num_a_stages;
  const int nchunks = p.nchunks;
  uint8_t* sA = smem;
  uint8_t* sW = smem + S * ASTAGE;                              // [tap][chunk] slab pairs
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + 9 * nchunks * WPAIR);
  uint64_t* full = bars;                        // [8]   (the leader's are the ones waited on)
  uint64_t* empty = bars + 8;                   // [8]
  uint64_t* w_full = bars + 16;                 // leader: weights of BOTH CTAs have landed
  uint64_t* tmem_full = bars + 18;              // [2]
  uint64_t* tmem_empty = bars + 20;             // [2]  (leader's collects both CTAs' epilogues)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
  float* sbias = reinterpret_cast<float*>(bars + 24);           // [64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  pdl_launch_dependents();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int slice = pair % p.n_slices;                          // npairs is a multiple of n_slices
  const int m0 = pair / p.n_slices, m_step = npairs / p.n_slices;
  const int n0 = slice * BNS;
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.a_map[0]);
    prefetch_tensormap(&p.w_map[0]);
    if (X3) { prefetch_tensormap(&p.a_map[1]); prefetch_tensormap(&p.w_map[1]); }
  }
  if (warp == 1) {
    if (lane < 8) { mbar_init(&full[lane], 2); mbar_init(&empty[lane], 1); }
    if (lane < 2) { mbar_init(&tmem_full[lane], 1); mbar_init(&tmem_empty[lane], 16); }
    if (lane == 2) mbar_init(w_full, 2);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, kTmem);
  if (threadIdx.x < BNS) sbias[threadIdx.x] = p.bias[n0 + threadIdx.x];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // peer's barriers initialised, TMEM allocated in both CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0 && lane == 0) {
    // this CTA's rows of every slab: constants, fetched before the dependency wait; both CTAs signal the leader's barrier
    const uint32_t wb = mapa_u32(w_full, 0);
    if (rank == 0) mbar_arrive_expect_tx(w_full, 2u * 9u * (uint32_t)nchunks * WPAIR);
    else mbar_arrive_cluster(wb);
    for (int tap = 0; tap < 9; ++tap)
      for (int c = 0; c < nchunks; ++c) {
        uint8_t* dst = sW + (tap * nchunks + c) * WPAIR;
        tma_load_3d_2sm(dst, &p.w_map[0], wb, c * KC, n0 + 32 * (int)rank, tap);
        if (X3) tma_load_3d_2sm(dst + WSLAB, &p.w_map[1], wb, c * KC, n0 + 32 * (int)rank, tap);
      }
  }
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer (both CTAs): own half-halo tile(s) of every chunk ----------------
      uint32_t ia = 0;
      for (int m = m0; m < p.num_m_tiles; m += m_step) {
        const int w0 = (m % p.tiles_w) * 16, h0 = ((m / p.tiles_w) % p.tiles_h) * 16;
        const int b = m / (p.tiles_w * p.tiles_h);
        for (int c = 0; c < nchunks; ++c, ++ia) {
          const int s = ia % S;
          mbar_wait(&empty[s], ((ia / S) & 1) ^ 1);
          const uint32_t fb = mapa_u32(&full[s], 0);            // the LEADER's barrier
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 2u * (X3 ? 2u : 1u) * (kPairARows * ROW));
          else mbar_arrive_cluster(fb);
          uint8_t* st = sA + s * ASTAGE;
          tma_load_4d_2sm(st, &p.a_map[0], fb, c * KC, w0 - 1 + 8 * (int)rank, h0 - 1, b);
          if (X3) tma_load_4d_2sm(st + APLANE, &p.a_map[1], fb, c * KC, w0 - 1 + 8 * (int)rank, h0 - 1, b);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ---------------- MMA issuer (leader only) ----------------
      const uint32_t idesc = make_idesc_f16(256, BNS);
      const uint32_t a_hi = (uint32_t)(make_smem_desc_ex(0, ROW, 10 * ROW, 0) >> 32);
      const uint32_t b_hi = (uint32_t)(make_smem_desc(0, ROW) >> 32);
      const uint32_t lo_flags = 1u << 16;
      const uint32_t sA_lo = (smem_u32(sA) >> 4) | lo_flags;
      const uint32_t sW_lo = (smem_u32(sW) >> 4) | lo_flags;
      mbar_wait(w_full, 0);
      tc_fence_after();
      uint32_t ia = 0, it = 0;
      for (int m = m0; m < p.num_m_tiles; m += m_step, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + buf * kBufCols;
        uint32_t accumulate = 0;
        for (int c = 0; c < nchunks; ++c, ++ia) {
          const int sa = ia % S;
          mbar_wait(&full[sa], (ia / S) & 1);
          tc_fence_after();
          const uint32_t a_lo = sA_lo + sa * (ASTAGE >> 4);
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t a_tap = a_lo + (((tap / 3) * 10 + tap % 3) * ROW >> 4);
              const uint32_t bh = sW_lo + (uint32_t)(tap * nchunks + c) * (WPAIR >> 4);
#pragma unroll
              for (int kk = 0; kk < KSTEPS; ++kk) {
                const uint64_t ad = pack_desc(a_tap + kk * 2, a_hi);
                const uint64_t bhd = pack_desc(bh + kk * 2, b_hi);
                umma2_f16(d0, ad, bhd, idesc, accumulate);
                if (X3) {
                  umma2_f16(d0 + BNS, ad, pack_desc(bh + (WSLAB >> 4) + kk * 2, b_hi), idesc, accumulate);          // a_hi w_lo
                  umma2_f16(d0 + BNS, pack_desc(a_tap + (APLANE >> 4) + kk * 2, a_hi), bhd, idesc, 1);              // a_lo w_hi
                }
                accumulate = 1;
              }
            }
            umma2_commit_mc(&empty[sa], 3);                     // frees the A stage in BOTH CTAs
            if (c == nchunks - 1) umma2_commit_mc(&tmem_full[buf], 3);
          }
          accumulate = 1;
          __syncwarp();
        }
      }
    }
  } else {
    // ---------------- epilogue (both CTAs): own 128 rows; the two warps of a quadrant split the 64 columns ----------------
    const int q = warp & 3;
    const int e = (warp - 2) >> 2;
    const int ml = q * 32 + lane;
    const int tw = ml & 7, th = ml >> 3;
    const uint32_t te = mapa_u32(&tmem_empty[0], 0);            // the leader's tmem_empty[0]; [1] is 8 bytes further
    const int c0 = 32 * e;
    uint32_t it = 0;
    for (int m = m0; m < p.num_m_tiles; m += m_step, ++it) {
      const int w = (m % p.tiles_w) * 16 + 8 * (int)rank + tw, h = ((m / p.tiles_w) % p.tiles_h) * 16 + th;
      const int b = m / (p.tiles_w * p.tiles_h);
      const size_t pix = ((size_t)b * p.H + h) * p.W + w;
      const uint32_t buf = it & 1;
      mbar_wait(&tmem_full[buf], (it >> 1) & 1);
      tc_fence_after();
      uint32_t r[32];
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + buf * kBufCols + c0;
      tmem_ld_32x32(ta, r);
      if (X3) {
        uint32_t rc[32];
        tmem_ld_32x32(ta + BNS, rc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaf(__uint_as_float(rc[j]), kLoInv, __uint_as_float(r[j])));
      } else {
        tmem_ld_wait();
      }
      // the accumulators are in registers: hand the TMEM buffer back before the stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(te + buf * 8);
      float v[32];
      epilogue_act32(r, sbias + c0, v);
      epilogue_store_nhwc32(v, p.out_hi, X3 ? p.out_lo : nullptr, pix * p.Cout + n0 + c0);
      if (p.pool_hi) {                        // 2x2 max over (tw^1, th^1) = lanes ^1 and ^8
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
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // nobody leaves (or frees TMEM) while the peer may still signal it
  if (warp == 2) tmem_dealloc_2sm(tmem_base, kTmem);
}
