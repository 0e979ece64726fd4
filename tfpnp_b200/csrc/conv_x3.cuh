// Split-fp16 ("fp16x3") halo-tile convolution for the 32- and 64-channel levels of the UNet: the mode that meets the
// 1e-4 contract for any weights.  Included by unet_tc.cu (same translation unit, anonymous namespace).
//
// Operands: a = a_hi + a_lo, w = w_hi + w_lo (fp16 + fp16 residual, ~22 bits each); the product keeps the three leading
// terms  a_hi w_hi + a_hi w_lo + a_lo w_hi  in fp32 TMEM accumulators.  What is different from running the fp16 kernel three
// times (the round-1 path: 147 us for a 32->32 layer at 48 x 128^2 against 37 us in fp16):
//   * ONE halo tile per plane per chunk -- a_hi and a_lo are each fetched once (the K-loop-over-products form fetched
//     a_hi twice and every weight slab up to twice);
//   * N-concatenation: [W_hi | W_lo] of a (tap, chunk) sit back to back in shared memory, so a_hi w_hi and a_hi w_lo are ONE
//     tcgen05.mma with N = 2 BN writing D[:, 0:BN] and D[:, BN:2BN]; a second MMA with N = BN adds a_lo w_hi into D[:, BN:2BN].
//     (The residual planes a_lo, w_lo are stored x 2^11 so that they stay in fp16's normal range -- grad_elem.cuh kLoScale -- so
//     both correction products belong in the second column block, which the epilogue scales back.)
//     Below N = 128 an SS-mode MMA is bound by the shared-memory operand read, 32 + N/4 cycles per 128 x N x 16 (measured,
//     profiles/r01_ubench_mma_chip.txt), so the three products cost 88 cycles at BN = 32 (fp16: 40) and 115 at BN = 64
//     (fp16: 48) instead of three times the fp16 figure;
//   * the epilogue computes main + corrections / 2^11 (the small terms are accumulated apart from the big one, which also helps
//     the rounding), applies bias + LeakyReLU and writes the fp16 hi plane and the scaled fp16 residual plane.
// 32-channel chunks (64-byte rows) so that two planes x 3-4 stages of halo tiles plus the resident / streamed weights fit;
// one CTA per SM, eight epilogue warps (two per TMEM lane quadrant, one per M-tile half), accumulators double-buffered.

template <int BN, bool RESIDENT>
__global__ void __launch_bounds__(64 + 32 * 8, 1)
conv3x3_x3(const __grid_constant__ Conv2Params p) {
  static_assert(BN == 32 || BN == 64, "N-concatenated split-fp16 products: BN = 32 or 64");
  constexpr int KC = 32, KSTEPS = 2, TPS = 3;
  constexpr uint32_t ROW = KC * 2;                  // 64-byte pixel rows (SWIZZLE_64B)
  constexpr uint32_t SLAB = BN * ROW;               // one plane of one (tap, chunk) weight slab
  constexpr uint32_t SLAB2 = 2 * SLAB;              // [W_hi | W_lo]
  constexpr int NEPI = 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nchunks = p.nchunk0 + p.nchunk1;
  const int SA = p.num_a_stages, SB = p.num_b_stages;
  const uint32_t plane = (uint32_t)p.a_stage_bytes >> 1;     // a stage = [hi plane | lo plane]
  uint8_t* sA = smem;
  uint8_t* sW = smem + SA * p.a_stage_bytes;
  const int w_region = RESIDENT ? 9 * nchunks * (int)SLAB2 : SB * TPS * (int)SLAB2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + w_region);
  uint64_t* full_a = bars;
  uint64_t* empty_a = bars + kMaxStages;
  uint64_t* full_b = bars + 2 * kMaxStages;
  uint64_t* empty_b = bars + 3 * kMaxStages;
  uint64_t* w_full = bars + 4 * kMaxStages;
  uint64_t* tmem_full = w_full + 1;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* sbias = reinterpret_cast<float*>(bars + 4 * kMaxStages + 12);   // [Cout]
  float* soutc = sbias + 512;                                            // [33] fused outconv weights + bias

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  constexpr uint32_t kTmemCols = 8 * BN;            // 2 buffers x 2 M-tile halves x (BN main + BN correction) columns
  const int cs = RESIDENT ? 1 : p.cluster;
  const int crank = cs > 1 ? (int)cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << cs) - 1);
  const int total_items = (p.num_m_tiles + cs - 1) / cs;
  const int item0 = blockIdx.x / cs, item_step = gridDim.x / cs;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.a_map[0][0]);
    prefetch_tensormap(&p.a_map[0][1]);
    prefetch_tensormap(&p.w_map[0]);
    prefetch_tensormap(&p.w_map[1]);
    if (p.nchunk1) { prefetch_tensormap(&p.a_map[1][0]); prefetch_tensormap(&p.a_map[1][1]); }
  }
  if (warp == 1) {
    if (lane < kMaxStages) {
      mbar_init(&full_a[lane], 1); mbar_init(&empty_a[lane], 1);
      mbar_init(&full_b[lane], 1); mbar_init(&empty_b[lane], cs);
    } else if (lane < kMaxStages + 2) {
      const int i = lane - kMaxStages;
      mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], NEPI);
    } else if (lane == kMaxStages + 2) {
      mbar_init(w_full, 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) sbias[i] = p.bias[i];
  if (p.outc_w && threadIdx.x < 33) soutc[threadIdx.x] = p.outc_w[threadIdx.x];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (cs > 1) cluster_sync_all();
  const uint32_t tmem_base = *tmem_slot;
  if (RESIDENT && warp == 0 && lane == 0) {     // weights are constants: fetch both planes before the dependency wait
    mbar_arrive_expect_tx(w_full, (uint32_t)(9 * nchunks) * SLAB2);
    for (int tap = 0; tap < 9; ++tap)
      for (int c = 0; c < nchunks; ++c) {
        uint8_t* dst = sW + (tap * nchunks + c) * SLAB2;
        tma_load_3d(dst, &p.w_map[0], w_full, c * KC, 0, tap);
        tma_load_3d(dst + SLAB, &p.w_map[1], w_full, c * KC, 0, tap);
      }
  }
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      uint32_t ia = 0, ib = 0;
      const int rows_mc = BN / cs;
      for (int t = item0; t < total_items; t += item_step) {
        int m = t * cs + crank;
        if (m >= p.num_m_tiles) m = p.num_m_tiles - 1;      // padding CTA of a cluster: recompute the last tile, stores masked
        const int w0 = (m % p.tiles_w) * 16, h0 = ((m / p.tiles_w) % p.tiles_h) * 16;
        const int b = m / (p.tiles_w * p.tiles_h);
        for (int c = 0; c < nchunks; ++c, ++ia) {
          const int src = c < p.nchunk0 ? 0 : 1;
          const int cc = (src == 0 ? c : c - p.nchunk0) * KC;
          const int s = ia % SA;
          mbar_wait(&empty_a[s], ((ia / SA) & 1) ^ 1);
          mbar_arrive_expect_tx(&full_a[s], 2u * (uint32_t)kHaloRows * ROW);
          tma_load_4d(sA + s * p.a_stage_bytes, &p.a_map[src][0], &full_a[s], cc, w0 - 1, h0 - 1, b);
          tma_load_4d(sA + s * p.a_stage_bytes + plane, &p.a_map[src][1], &full_a[s], cc, w0 - 1, h0 - 1, b);
          if (!RESIDENT) {
#pragma unroll 1
            for (int tg = 0; tg < 9 / TPS; ++tg, ++ib) {
              const int sb = ib % SB;
              mbar_wait(&empty_b[sb], ((ib / SB) & 1) ^ 1);
              mbar_arrive_expect_tx(&full_b[sb], TPS * SLAB2);
#pragma unroll
              for (int tt = 0; tt < TPS; ++tt) {
                uint8_t* dst = sW + sb * (TPS * SLAB2) + tt * SLAB2;
                if (cs == 1) {
                  tma_load_3d(dst, &p.w_map[0], &full_b[sb], c * KC, 0, tg * TPS + tt);
                  tma_load_3d(dst + SLAB, &p.w_map[1], &full_b[sb], c * KC, 0, tg * TPS + tt);
                } else {
                  tma_load_3d_mc(dst + crank * rows_mc * ROW, &p.w_map[0], &full_b[sb], cmask, c * KC, crank * rows_mc,
                                 tg * TPS + tt);
                  tma_load_3d_mc(dst + SLAB + crank * rows_mc * ROW, &p.w_map[1], &full_b[sb], cmask, c * KC,
                                 crank * rows_mc, tg * TPS + tt);
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (whole warp converged; one elected lane issues) ----------------
    const uint32_t idesc2 = make_idesc_f16(kTileM, 2 * BN);   // a_hi x [w_hi | w_lo]
    const uint32_t idesc1 = make_idesc_f16(kTileM, BN);       // a_lo x w_hi
    const uint32_t a_hi = (uint32_t)(make_smem_desc_ex(0, ROW, kHaloW * ROW, 0) >> 32);
    const uint32_t b_hi = (uint32_t)(make_smem_desc(0, ROW) >> 32);
    const uint32_t lo_flags = 1u << 16;
    const uint32_t sA_lo = (smem_u32(sA) >> 4) | lo_flags;
    const uint32_t sW_lo = (smem_u32(sW) >> 4) | lo_flags;
    const uint32_t a_stage16 = (uint32_t)p.a_stage_bytes >> 4;
    const uint32_t plane16 = plane >> 4;
    if (RESIDENT) { mbar_wait(w_full, 0); tc_fence_after(); }
    uint32_t it = 0;
    uint32_t sa = 0, pha = 0, sb = 0, phb = 0;
    for (int t = item0; t < total_items; t += item_step, ++it) {
      const uint32_t buf = it & 1;
      mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + buf * (4 * BN);   // left half [main | corr]; right half at + 2 BN
      uint32_t accumulate = 0;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&full_a[sa], pha);
        tc_fence_after();
        const uint32_t ah = sA_lo + sa * a_stage16;     // hi plane; lo plane at + plane16
        const bool last = c == nchunks - 1;
        if (RESIDENT) {
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t a_tap = ah + (((tap / 3) * kHaloW + tap % 3) * ROW >> 4);
              const uint32_t b_lo = sW_lo + (uint32_t)(tap * nchunks + c) * (SLAB2 >> 4);
#pragma unroll
              for (int kk = 0; kk < KSTEPS; ++kk) {
                const uint64_t bd = pack_desc(b_lo + kk * 2, b_hi);
                const uint64_t bdl = bd;       // N = BN reads the first BN rows of the slab pair = W_hi
                umma_f16(d0, pack_desc(a_tap + kk * 2, a_hi), bd, idesc2, accumulate);
                umma_f16(d0 + 2 * BN, pack_desc(a_tap + (8 * ROW >> 4) + kk * 2, a_hi), bd, idesc2, accumulate);
                umma_f16(d0 + BN, pack_desc(a_tap + plane16 + kk * 2, a_hi), bdl, idesc1, 1);
                umma_f16(d0 + 3 * BN, pack_desc(a_tap + plane16 + (8 * ROW >> 4) + kk * 2, a_hi), bdl, idesc1, 1);
                accumulate = 1;
              }
            }
            umma_commit(&empty_a[sa]);
            if (last) umma_commit(&tmem_full[buf]);
          }
          accumulate = 1;
          __syncwarp();
        } else {
#pragma unroll 1
          for (int tg = 0; tg < 9 / TPS; ++tg) {
            mbar_wait(&full_b[sb], phb);
            tc_fence_after();
            const uint32_t b_stage = sW_lo + sb * (TPS * SLAB2 >> 4);
            const uint32_t a_row = ah + ((tg * kHaloW) * ROW >> 4);     // TPS = 3: stage tg = kernel row tg
            if (elect_one()) {
#pragma unroll
              for (int tt = 0; tt < TPS; ++tt) {
                const uint32_t a_tap = a_row + (tt * ROW >> 4);
                const uint32_t b_lo = b_stage + tt * (SLAB2 >> 4);
#pragma unroll
                for (int kk = 0; kk < KSTEPS; ++kk) {
                  const uint64_t bd = pack_desc(b_lo + kk * 2, b_hi);
                  const uint64_t bdl = bd;     // N = BN reads the first BN rows of the slab pair = W_hi
                  umma_f16(d0, pack_desc(a_tap + kk * 2, a_hi), bd, idesc2, accumulate);
                  umma_f16(d0 + 2 * BN, pack_desc(a_tap + (8 * ROW >> 4) + kk * 2, a_hi), bd, idesc2, accumulate);
                  umma_f16(d0 + BN, pack_desc(a_tap + plane16 + kk * 2, a_hi), bdl, idesc1, 1);
                  umma_f16(d0 + 3 * BN, pack_desc(a_tap + plane16 + (8 * ROW >> 4) + kk * 2, a_hi), bdl, idesc1, 1);
                  accumulate = 1;
                }
              }
              if (cs == 1) umma_commit(&empty_b[sb]);
              else umma_commit_mc(&empty_b[sb], cmask);
            }
            accumulate = 1;
            __syncwarp();
            if (++sb == (uint32_t)SB) { sb = 0; phb ^= 1; }
          }
          if (elect_one()) {
            umma_commit(&empty_a[sa]);
            if (last) umma_commit(&tmem_full[buf]);
          }
          __syncwarp();
        }
        if (++sa == (uint32_t)SA) { sa = 0; pha ^= 1; }
      }
    }
  } else {
    // ---------------- epilogue: warp pair member e takes M-tile half e ----------------
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int ml = q * 32 + lane;
    const int tw = ml & 7, th = ml >> 3;
    uint32_t it = 0;
    for (int t = item0; t < total_items; t += item_step, ++it) {
      int m = t * cs + crank;
      const bool real_tile = m < p.num_m_tiles;
      if (!real_tile) m = p.num_m_tiles - 1;
      const int w = (m % p.tiles_w) * 16 + tw + half * 8;
      const int h = ((m / p.tiles_w) % p.tiles_h) * 16 + th;
      const int b = m / (p.tiles_w * p.tiles_h);
      const size_t pix = ((size_t)b * p.H + h) * p.W + w;
      const uint32_t buf = it & 1;
      mbar_wait(&tmem_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (4 * BN) + half * (2 * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32], rc[32];
        tmem_ld_32x32(tbase + c0, r);
        tmem_ld_32x32(tbase + BN + c0, rc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaf(__uint_as_float(rc[j]), kLoInv, __uint_as_float(r[j])));
        float v[32];
        epilogue_act32(r, sbias + c0, v);
        if (p.outc_w) {                       // last layer: 1x1 conv + residual + clamp, fp32 out (BN == 32)
          float acc = soutc[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) acc = fmaf(soutc[j], v[j], acc);
          if (real_tile) p.x_out[pix] = fminf(fmaxf(p.d_in[pix] + acc, 0.f), 1.f);
          continue;
        }
        if (real_tile) epilogue_store_nhwc32(v, p.out_hi, p.out_lo, pix * p.Cout + c0);
        if (p.pool_hi) {                      // fused nn.MaxPool2d(2): max over (tw^1, th^1) = lanes ^1 and ^8
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
            v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 8));
          }
          if (real_tile && !(lane & 9)) {
            const size_t ppix = ((size_t)b * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
            epilogue_store_nhwc32(v, p.pool_hi, p.pool_lo, ppix * p.Cout + c0);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}
