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

//
// FUSE: the decoder heads (96 -> 32 @ full resolution, 192 -> 64 @ half resolution): source 1 is the bilinear x2 up-sampling
// (align_corners=True, unet.py:99) of a low-resolution tensor that is never materialised (in split-fp16 the up-sampled tensor
// is 200 MB per plane pair at 48 x 128^2 x 64 channels, written and read back once per call: ~250 us of a 1.8 ms call).
// The producer stages the 11 x 11 low-resolution window of both planes (unswizzled TMA box); nine transform warps interpolate
// in fp32 from hi + lo / 2^11, split the result again and write both planes of the 18 x 18 halo tile into the swizzled A
// stage.  Work item of a transform thread = (coarse row interval i, fine column k, 8-channel group j): it loads the two
// source rows of the interval at the column's two source pixels ONCE (8 x LDS.128 for both planes), interpolates
// horizontally, then emits the 1-3 fine rows that fall into the interval -- no state carried from row to row, all loads of
// an item in flight together.  Four epilogue warps (each takes both M-tile halves) keep the CTA at 480 threads.
constexpr int kX3StgPlane = 61 * 128;                  // 121 rows x 64 B = 7744, padded to the TMA's 128-byte alignment
constexpr int kX3StgSlot = 2 * kX3StgPlane;            // [hi | lo]
constexpr int kX3XformThreads = 288;

template <bool FUSE> constexpr int conv_x3_nepi() { return FUSE ? 4 : 8; }
template <bool FUSE> constexpr int conv_x3_threads() { return 64 + 32 * conv_x3_nepi<FUSE>() + (FUSE ? kX3XformThreads : 0); }

template <int BN, bool RESIDENT, bool FUSE>
__global__ void __launch_bounds__(conv_x3_threads<FUSE>(), 1)
conv3x3_x3(const __grid_constant__ Conv2Params p) {
  static_assert(BN == 32 || BN == 64, "N-concatenated split-fp16 products: BN = 32 or 64");
  constexpr int KC = 32, KSTEPS = 2, TPS = 3;
  constexpr uint32_t ROW = KC * 2;                  // 64-byte pixel rows (SWIZZLE_64B)
  constexpr uint32_t SLAB = BN * ROW;               // one plane of one (tap, chunk) weight slab
  constexpr uint32_t SLAB2 = 2 * SLAB;              // [W_hi | W_lo]
  constexpr int NEPI = conv_x3_nepi<FUSE>();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nchunks = p.nchunk0 + p.nchunk1;
  const int SA = p.num_a_stages, SB = p.num_b_stages;
  const uint32_t plane = (uint32_t)p.a_stage_bytes >> 1;     // a stage = [hi plane | lo plane]
  uint8_t* sA = smem;
  uint8_t* sW = smem + SA * p.a_stage_bytes;
  const int w_region = RESIDENT ? 9 * nchunks * (int)SLAB2 : SB * TPS * (int)SLAB2;
  uint8_t* sStg = sW + w_region;                                  // [2][kX3StgSlot] low-resolution windows (FUSE only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + (FUSE ? 2 * kX3StgSlot : 0));
  uint64_t* full_a = bars;
  uint64_t* empty_a = bars + kMaxStages;
  uint64_t* full_b = bars + 2 * kMaxStages;
  uint64_t* empty_b = bars + 3 * kMaxStages;
  uint64_t* w_full = bars + 4 * kMaxStages;
  uint64_t* tmem_full = w_full + 1;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;  // [2]
  uint64_t* stg_full = tmem_empty + 2;   // [2]
  uint64_t* stg_empty = stg_full + 2;    // [2]
  // FUSE: a stage of the A ring is claimed by the TMA producer (source-0 chunks) or by the transform warps (fused chunks).
  // A claimant that skipped the other party's uses of a slot could not tell mbarrier phases apart by parity (it may be two
  // phases ahead: measured as a data race with resident weights, where nothing else throttles the producer), so every slot
  // has TWO release barriers: the MMA warp signals empty_a[s] when the NEXT use of slot s is a TMA chunk and empty_x[s] when
  // it is a fused chunk; each party then sees exactly one phase per claim of its own.
  uint64_t* empty_x = stg_empty + 2;     // [kMaxStages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty_x + kMaxStages);
  float* sbias = reinterpret_cast<float*>(bars + 5 * kMaxStages + 12);   // [Cout]
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
      mbar_init(&empty_x[lane], 1);
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
      uint32_t ia = 0, ib = 0, iu = 0;
      uint32_t tcnt[kMaxStages];               // FUSE: releases of each slot this thread has consumed (phase counter)
#pragma unroll
      for (int i = 0; i < kMaxStages; ++i) tcnt[i] = 0;
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
          if (FUSE && src == 1) {
            // hand the low-resolution window (both planes) to the transform warps as soon as a staging slot is free; they
            // claim A stage s themselves and fill it
            const int st = iu & 1;
            mbar_wait(&stg_empty[st], ((iu >> 1) & 1) ^ 1);
            mbar_arrive_expect_tx(&stg_full[st], 2u * (uint32_t)(kUpBox * kUpBox) * ROW);
            const int ys = (int)(p.up_sy * (float)(h0 > 0 ? h0 - 1 : 0));
            const int xs = (int)(p.up_sx * (float)(w0 > 0 ? w0 - 1 : 0));
            tma_load_4d(sStg + st * kX3StgSlot, &p.a_map[1][0], &stg_full[st], cc, xs, ys, b);
            tma_load_4d(sStg + st * kX3StgSlot + kX3StgPlane, &p.a_map[1][1], &stg_full[st], cc, xs, ys, b);
            ++iu;
          } else {
            if (FUSE) {
              if (ia >= (uint32_t)SA) { mbar_wait(&empty_a[s], tcnt[s] & 1); ++tcnt[s]; }   // first use of a slot: free
            } else {
              mbar_wait(&empty_a[s], ((ia / SA) & 1) ^ 1);
            }
            mbar_arrive_expect_tx(&full_a[s], 2u * (uint32_t)kHaloRows * ROW);
            tma_load_4d(sA + s * p.a_stage_bytes, &p.a_map[src][0], &full_a[s], cc, w0 - 1, h0 - 1, b);
            tma_load_4d(sA + s * p.a_stage_bytes + plane, &p.a_map[src][1], &full_a[s], cc, w0 - 1, h0 - 1, b);
          }
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
        // who claims this slot next: chunk (c + SA) of the tile cycle
        uint64_t* rel_a = (FUSE && ((c + SA) % nchunks) >= p.nchunk0) ? &empty_x[sa] : &empty_a[sa];
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
            umma_commit(rel_a);
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
            umma_commit(rel_a);
            if (last) umma_commit(&tmem_full[buf]);
          }
          __syncwarp();
        }
        if (++sa == (uint32_t)SA) { sa = 0; pha ^= 1; }
      }
    }
  } else if (FUSE && warp >= 2 + NEPI) {
    // ---------------- transform warps: bilinear x2 (align_corners=True) of the staged low-resolution window -> A stage ----
    const int tid = threadIdx.x - (64 + 32 * NEPI);          // 0..287
    const int i0 = tid / 72, slot = tid % 72;                // intervals i0, i0 + 4, i0 + 8 of fine column k, channel group j
    const int k = slot >> 2, j = slot & 3;
    const int Hin = p.H >> 1, Win = p.W >> 1;
    constexpr int iROW = (int)ROW;
    uint32_t ia = 0, iu = 0;
    uint32_t xc0 = 0, xc1 = 0, xc2 = 0, xc3 = 0;             // releases of slots 0..3 consumed by the transform warps (SA <= 4)
    auto store_px = [&](uint8_t* dstA, int r, uint4 vhi, uint4 vlo) {
      const int pr = r * kHaloW + k;
      const int off = pr * iROW + ((j ^ ((pr >> 1) & 3)) << 4);      // TMA-compatible SWIZZLE_64B of the A stage
      *reinterpret_cast<uint4*>(dstA + off) = vhi;
      *reinterpret_cast<uint4*>(dstA + plane + off) = vlo;
    };
    for (int t = item0; t < total_items; t += item_step) {
      int m = t * cs + crank;
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
      const int ox0 = (x0 - xs) * 128 + j * 32, ox1 = (x1 - xs) * 128 + j * 32;      // fp32 staging rows of 128 bytes
      // row set-up (fixed for the tile): first fine row and number of fine rows of this thread's three coarse intervals
      int rf[3], cn[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int i = i0 + 4 * q, Y = ys + i;
        rf[q] = 0; cn[q] = 0;
        if (x_ok && i < kUpBox - 1 && Y < Hin) {
          int r = 2 * i - 1;
          if (h0 - 1 + r < 0) r = 1 - h0;                    // first fine row inside the image
          while (r < kHaloH && (int)(p.up_sy * (float)(h0 - 1 + r)) < Y) ++r;
          int n = 0;
          while (r + n < kHaloH && h0 - 1 + r + n < p.H && (int)(p.up_sy * (float)(h0 - 1 + r + n)) == Y) ++n;
          rf[q] = r; cn[q] = n;
        }
      }
      for (int c = 0; c < nchunks; ++c, ++ia) {
        if (c < p.nchunk0) continue;                         // source 0 comes by TMA
        const int s = ia % SA, st = iu & 1;
        mbar_wait(&stg_full[st], (iu >> 1) & 1);
        uint8_t* stg = sStg + st * kX3StgSlot;
        // pre-pass: [hi plane | lo plane] fp16 -> fp32 pixel rows IN PLACE (121 x 32 channels x 4 bytes = exactly the slot), so
        // that every staged value is combined and converted once instead of once per fine pixel that reads it
        {
          uint4 ph[2], pl[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int gi = tid + kX3XformThreads * q;
            if (gi < 4 * kUpBox * kUpBox) {
              ph[q] = *reinterpret_cast<const uint4*>(stg + gi * 16);
              pl[q] = *reinterpret_cast<const uint4*>(stg + kX3StgPlane + gi * 16);
            }
          }
          asm volatile("bar.sync 1, 288;" ::: "memory");     // every fp16 value is in registers
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int gi = tid + kX3XformThreads * q;
            if (gi < 4 * kUpBox * kUpBox) {
              const __half2* hh = reinterpret_cast<const __half2*>(&ph[q]);
              const __half2* ll = reinterpret_cast<const __half2*>(&pl[q]);
              float2 f[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) f[e] = ffma2(kLoInv, __half22float2(ll[e]), __half22float2(hh[e]));
              float4* dst = reinterpret_cast<float4*>(stg + gi * 32);
              dst[0] = make_float4(f[0].x, f[0].y, f[1].x, f[1].y);
              dst[1] = make_float4(f[2].x, f[2].y, f[3].x, f[3].y);
            }
          }
          asm volatile("bar.sync 1, 288;" ::: "memory");
        }
        if (ia >= (uint32_t)SA) {                            // the staging load ran ahead of the A ring: claim the stage here
          uint32_t& xc = s == 0 ? xc0 : s == 1 ? xc1 : s == 2 ? xc2 : xc3;
          mbar_wait(&empty_x[s], xc & 1);
          ++xc;
        }
        uint8_t* dstA = sA + s * p.a_stage_bytes;
        const uint4 zero = make_uint4(0, 0, 0, 0);
        if (!x_ok) {                                         // conv zero padding left / right of the image
          for (int r = i0; r < kHaloH; r += 4) store_px(dstA, r, zero, zero);
        } else {
          if (i0 == 0 && h0 == 0) store_px(dstA, 0, zero, zero);                       // ... above
          if (i0 == 3 && h0 + 16 == p.H) store_px(dstA, kHaloH - 1, zero, zero);       // ... below
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            if (cn[q] == 0) continue;
            const int i = i0 + 4 * q, Y = ys + i;
            // the interval's two source rows at the column's two source pixels
            const uint8_t* rp0 = stg + i * (kUpBox * 128);
            const uint8_t* rp1 = rp0 + (Y < Hin - 1 ? kUpBox * 128 : 0);
            const float4 a0a = *reinterpret_cast<const float4*>(rp0 + ox0), a0b = *reinterpret_cast<const float4*>(rp0 + ox0 + 16);
            const float4 b0a = *reinterpret_cast<const float4*>(rp0 + ox1), b0b = *reinterpret_cast<const float4*>(rp0 + ox1 + 16);
            const float4 a1a = *reinterpret_cast<const float4*>(rp1 + ox0), a1b = *reinterpret_cast<const float4*>(rp1 + ox0 + 16);
            const float4 b1a = *reinterpret_cast<const float4*>(rp1 + ox1), b1b = *reinterpret_cast<const float4*>(rp1 + ox1 + 16);
            float2 top[4], bot[4];
            top[0] = ffma2(lx, make_float2(b0a.x, b0a.y), fmul2(wx, make_float2(a0a.x, a0a.y)));
            top[1] = ffma2(lx, make_float2(b0a.z, b0a.w), fmul2(wx, make_float2(a0a.z, a0a.w)));
            top[2] = ffma2(lx, make_float2(b0b.x, b0b.y), fmul2(wx, make_float2(a0b.x, a0b.y)));
            top[3] = ffma2(lx, make_float2(b0b.z, b0b.w), fmul2(wx, make_float2(a0b.z, a0b.w)));
            bot[0] = ffma2(lx, make_float2(b1a.x, b1a.y), fmul2(wx, make_float2(a1a.x, a1a.y)));
            bot[1] = ffma2(lx, make_float2(b1a.z, b1a.w), fmul2(wx, make_float2(a1a.z, a1a.w)));
            bot[2] = ffma2(lx, make_float2(b1b.x, b1b.y), fmul2(wx, make_float2(a1b.x, a1b.y)));
            bot[3] = ffma2(lx, make_float2(b1b.z, b1b.w), fmul2(wx, make_float2(a1b.z, a1b.w)));
#pragma unroll 1
            for (int n = 0; n < cn[q]; ++n) {
              const int r = rf[q] + n;
              const float fy = p.up_sy * (float)(h0 - 1 + r);
              const float ly = fy - (float)Y, wy = 1.f - ly;
              H8 ohi, olo;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 v = ffma2(ly, bot[e], fmul2(wy, top[e]));
                ohi.v[e] = __floats2half2_rn(v.x, v.y);
                const float2 back = __half22float2(ohi.v[e]);
                olo.v[e] = __floats2half2_rn((v.x - back.x) * kLoScale, (v.y - back.y) * kLoScale);
              }
              store_px(dstA, r, *reinterpret_cast<uint4*>(&ohi), *reinterpret_cast<uint4*>(&olo));
            }
          }
        }
        fence_proxy_async();                                 // generic-proxy smem writes -> visible to the MMA (async proxy)
        asm volatile("bar.sync 1, 288;" ::: "memory");       // the nine transform warps
        if (tid == 0) { mbar_arrive(&full_a[s]); mbar_arrive(&stg_empty[st]); }
        ++iu;
      }
    }
  } else if (warp >= 2 && warp < 2 + NEPI) {
    // ---------------- epilogue: 8 warps -> warp pair member e takes M-tile half e; 4 warps (FUSE) -> both halves each ----
    const int q = warp & 3;
    const int ml = q * 32 + lane;
    const int tw = ml & 7, th = ml >> 3;
    constexpr int NHALF = NEPI == 8 ? 1 : 2;
    uint32_t it = 0;
    for (int t = item0; t < total_items; t += item_step, ++it) {
      int m = t * cs + crank;
      const bool real_tile = m < p.num_m_tiles;
      if (!real_tile) m = p.num_m_tiles - 1;
      const int b = m / (p.tiles_w * p.tiles_h);
      const int h = ((m / p.tiles_w) % p.tiles_h) * 16 + th;
      const uint32_t buf = it & 1;
      mbar_wait(&tmem_full[buf], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int hh = 0; hh < NHALF; ++hh) {
      const int half = NEPI == 8 ? (warp - 2) >> 2 : hh;
      const int w = (m % p.tiles_w) * 16 + tw + half * 8;
      const size_t pix = ((size_t)b * p.H + h) * p.W + w;
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

// ==========================================================================================
// Split-fp16 on CTA pairs (tcgen05.mma.cta_group::2) for the 128/256/512-channel levels, whole-chunk stages
// ==========================================================================================
// Same work decomposition as conv3x3_pair (16x16 super-tile x 128 couts per pair, CTA r owns the 16x8 half r and rows
// [64 r, 64 r + 64) of every weight slab), but with 32-channel chunks so that ONE ring stage holds everything a chunk needs in
// both planes: [A_hi half-halo | A_lo half-halo | 9 taps x (W_hi half slab | W_lo half slab)] = 96 KB, two stages, ONE
// full / empty barrier pair per chunk and 54 MMAs (9 taps x 2 k-steps x 3 products = 1.8 us of tensor work) per hand-shake.
// (The first x3 instantiation of conv3x3_pair used 64-channel chunks with a separate A ring and kernel-row weight stages --
// the layout that was 1.8x slower than whole-chunk stages in the fp16 bring-up, DESIGN.md 9.1 -- 43 us for 128->128 @32^2.)
// Accumulators: [main | corrections (residual planes are x 2^11)] x 2 buffers = 512 TMEM columns.
constexpr int kPX3APlane = 12288;                       // 180 rows x 64 B = 11520, padded to a multiple of 1024
constexpr int kPX3Slab = 64 * 64;                       // one tap's half slab (64 of the 128 couts), one plane
constexpr int kPX3Stage = 2 * kPX3APlane + 9 * 2 * kPX3Slab;   // 98304

__global__ void __launch_bounds__(kPairThreads, 1)
conv3x3_pair_x3(const __grid_constant__ ConvPairParams p) {
  constexpr int BN = 128, KC = 32, KSTEPS = 2;
  constexpr uint32_t ROW = 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.num_a_stages;
  const int nchunks = p.nchunk0 + p.nchunk1;
  const uint32_t aplane = (uint32_t)p.a_plane;                  // kPX3APlane, or 13 KB for the 8x8 level (200 rows)
  const uint32_t stage = 2 * aplane + 9 * 2 * kPX3Slab;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * stage);
  uint64_t* full = bars;                        // [8]   (the leader's are the ones waited on)
  uint64_t* empty = bars + 8;                   // [8]
  uint64_t* tmem_full = bars + 48;              // [2]
  uint64_t* tmem_empty = bars + 50;             // [2]  (leader's collects both CTAs' epilogues)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 52);
  float* sbias = reinterpret_cast<float*>(bars + 54);           // [Cout] <= 512

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.a_map[0][0]); prefetch_tensormap(&p.a_map[0][1]);
    prefetch_tensormap(&p.w_map[0]); prefetch_tensormap(&p.w_map[1]);
    if (p.nchunk1) { prefetch_tensormap(&p.a_map[1][0]); prefetch_tensormap(&p.a_map[1][1]); }
  }
  if (warp == 1) {
    if (lane < 8) { mbar_init(&full[lane], 2); mbar_init(&empty[lane], 1); }
    if (lane < 2) { mbar_init(&tmem_full[lane], 1); mbar_init(&tmem_empty[lane], 16); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, 512);
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
      // ---------------- TMA producer (both CTAs): own half-halo planes + own half of the chunk's nine slab pairs ----------------
      uint32_t ia = 0;
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
          uint8_t* st = smem + s * stage;
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 2u * (2u * (uint32_t)p.a_rows * ROW + 9u * 2u * kPX3Slab));
          else mbar_arrive_cluster(fb);
          if (p.small) {                                       // four 8x8 images per pair, two per CTA ({C, W, B, H} view)
            tma_load_4d_2sm(st, &p.a_map[src][0], fb, cc, -1, 4 * m + 2 * (int)rank, -1);
            tma_load_4d_2sm(st + aplane, &p.a_map[src][1], fb, cc, -1, 4 * m + 2 * (int)rank, -1);
          } else {
            tma_load_4d_2sm(st, &p.a_map[src][0], fb, cc, w0 - 1 + 8 * (int)rank, h0 - 1, b);
            tma_load_4d_2sm(st + aplane, &p.a_map[src][1], fb, cc, w0 - 1 + 8 * (int)rank, h0 - 1, b);
          }
          uint8_t* wst = st + 2 * aplane;
#pragma unroll
          for (int tt = 0; tt < 9; ++tt) {
            tma_load_3d_2sm(wst + (2 * tt) * kPX3Slab, &p.w_map[0], fb, c * KC, nt * BN + 64 * (int)rank, tt);
            tma_load_3d_2sm(wst + (2 * tt + 1) * kPX3Slab, &p.w_map[1], fb, c * KC, nt * BN + 64 * (int)rank, tt);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ---------------- MMA issuer (leader only): one barrier wait and one release per chunk ----------------
      const uint32_t idesc = make_idesc_f16(256, BN);
      const uint32_t a_hi = (uint32_t)(make_smem_desc_ex(0, ROW, 10 * ROW, 0) >> 32);
      const uint32_t b_hi = (uint32_t)(make_smem_desc(0, ROW) >> 32);
      const uint32_t s_lo = (smem_u32(smem) >> 4) | (1u << 16);
      uint32_t ia = 0, it = 0;
      for (int t = item0; t < total_items; t += item_step, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + buf * (2 * BN);         // [main | corrections]
        uint32_t accumulate = 0;
        for (int c = 0; c < nchunks; ++c, ++ia) {
          const int sa = ia % S;
          mbar_wait(&full[sa], (ia / S) & 1);
          tc_fence_after();
          const uint32_t a_lo = s_lo + sa * (stage >> 4);
          const uint32_t b_stage = a_lo + (2 * aplane >> 4);
          const uint32_t tapr = (uint32_t)p.tap_rows;
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t a_tap = a_lo + (((tap / 3) * tapr + tap % 3) * ROW >> 4);
              const uint32_t bh = b_stage + (2 * tap) * (kPX3Slab >> 4), bl = bh + (kPX3Slab >> 4);
#pragma unroll
              for (int kk = 0; kk < KSTEPS; ++kk) {
                const uint64_t ad = pack_desc(a_tap + kk * 2, a_hi);
                const uint64_t bhd = pack_desc(bh + kk * 2, b_hi);
                umma2_f16(d0, ad, bhd, idesc, accumulate);                                                        // a_hi w_hi
                umma2_f16(d0 + BN, ad, pack_desc(bl + kk * 2, b_hi), idesc, accumulate);                          // a_hi w_lo
                umma2_f16(d0 + BN, pack_desc(a_tap + (aplane >> 4) + kk * 2, a_hi), bhd, idesc, 1);               // a_lo w_hi
                accumulate = 1;
              }
            }
            umma2_commit_mc(&empty[sa], 3);                     // frees the chunk stage in BOTH CTAs
            if (c == nchunks - 1) umma2_commit_mc(&tmem_full[buf], 3);
          }
          accumulate = 1;
          __syncwarp();
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
      const uint32_t buf = it & 1;
      mbar_wait(&tmem_full[buf], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int sl = 0; sl < 2; ++sl) {
        const int c0 = e * 64 + 32 * sl;
        uint32_t r[32], rc[32];
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (2 * BN) + c0;
        tmem_ld_32x32(ta, r);
        tmem_ld_32x32(ta + BN, rc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaf(__uint_as_float(rc[j]), kLoInv, __uint_as_float(r[j])));
        float v[32];
        epilogue_act32(r, sbias + n0 + c0, v);
        if (b_ok) epilogue_store_nhwc32(v, p.out_hi, p.out_lo, pix * p.Cout + n0 + c0);
        if (p.pool_hi) {                      // 2x2 max over (tw^1, th^1) = lanes ^1 and ^8
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
            v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 8));
          }
          if (!(lane & 9)) {
            const size_t ppix = ((size_t)b * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
            epilogue_store_nhwc32(v, p.pool_hi, p.pool_lo, ppix * p.Cout + n0 + c0);
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
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}
