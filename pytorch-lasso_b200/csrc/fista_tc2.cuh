// fista_tc2_kernel -- two tiles in flight per SM ("slots"), included by fista_tc.cu.
//
// The single-tile kernel (fista_tc_kernel) keeps an fp32 master copy of y in TMEM (256 of 512
// columns) from phase A to phase C, which leaves no room to overlap two tiles: the HBM-bound
// phase A and the MMA-bound GEMM2 / phase C run back to back.  Here a CTA runs TWO independent
// slots, each with its own producer warp, MMA warp, 8 compute warps, two shared-memory stages and
// 224 TMEM columns, working on different tiles; while one slot streams codes for phase A the
// other one is usually in GEMM2 / phase C.  There is no y master: phase C re-reads the tile's
// z_cur / z_prev chunks (L2 hits: phase A loads them with an evict-last hint) and recomputes
// y with the same instruction sequence, i.e. the same bits.  z_next is written in place over
// the z_prev box of the stage and stored by the slot's producer thread (TMA store) just before
// it recycles the stage, so the compute warps never wait for a store.
//
// Per slot: TMEM [0,96) piece stages (2 x 48) / r pieces (3 x 32) | [96,160) R_big / G buffer 0
// | [160,224) R_small / G buffer 1.  Shared memory: W pieces 96 KB (shared) + 2 x 2 x 32 KB stages.

constexpr int kThreads2 = 640;   // warps 0,1: producers; 2,3: MMA issuers; 4..19: compute (2 x 8)
constexpr uint32_t kSlotCols = 224;
constexpr uint32_t kColStage2 = 0;
constexpr uint32_t kColAcc0_2 = 96;
constexpr uint32_t kColAcc1_2 = 160;

struct SlotBarriers {
  uint64_t full[2], cons[2], aready[2], sfree[2], rfull, rready, gfull[2], gfree[2];
};

template <int kDSteps>
__global__ void __launch_bounds__(kThreads2, 1)
fista_tc2_kernel(const __grid_constant__ CUtensorMap tm_za, const __grid_constant__ CUtensorMap tm_zb,
                 TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_w, bar_stagger;
  __shared__ SlotBarriers sb[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ double red[16];

  if (p.ctl.tol_abs >= 0.0 && p.ctl.iter >= 2 && p.ctl.hist[p.ctl.iter - 2] <= p.ctl.tol_abs) return;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int tr_n = 0;
  const int64_t ntiles = (p.n + p.tile_rows - 1) / p.tile_rows;
  const int nc = (p.k + kChunk - 1) / kChunk;   // 32-atom chunks (phase A) = sub-chunks (phase C)
  const int nq = (p.k + kQ - 1) / kQ;           // 64-atom GEMM2 chunks
  constexpr int dsteps = kDSteps;   // ceil(d / 16)
  const CUtensorMap* tm_cur = p.cur_is_a ? &tm_za : &tm_zb;
  const CUtensorMap* tm_prev = p.cur_is_a ? &tm_zb : &tm_za;

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_stagger, 1);
    for (int s = 0; s < 2; ++s) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&sb[s].full[i], 1);
        mbar_init(&sb[s].cons[i], 256);
        mbar_init(&sb[s].aready[i], 256);
        mbar_init(&sb[s].sfree[i], 1);
        mbar_init(&sb[s].gfull[i], 1);
        mbar_init(&sb[s].gfree[i], 256);
      }
      mbar_init(&sb[s].rfull, 1);
      mbar_init(&sb[s].rready, 256);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_s, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase0 = tmem_base_s;

  // this slot's tiles: blockIdx.x + gridDim.x * (2 i + slot)
  const int slot = warp < 2 ? warp : (warp < 4 ? warp - 2 : (warp - 4) >> 3);
  SlotBarriers& B = sb[slot];
  const uint32_t tbase = tbase0 + slot * kSlotCols;
  uint8_t* stage_base = smem + kSmemStage + (uint32_t)slot * 2u * kStageBytes;
  const int64_t tile0 = (int64_t)blockIdx.x + (int64_t)gridDim.x * slot;
  const int64_t tile_step = 2 * (int64_t)gridDim.x;

  if (warp < 2) {
    // ===================== TMA producer of this slot =====================
    if (warp == 0 && elect_one()) {
      prefetch_tmap(&tm_za);
      prefetch_tmap(&tm_zb);
      mbar_expect_tx(&bar_w, kWBytes);
      for (uint32_t off = 0; off < kWBytes; off += 16384)
        bulk_load(smem + kSmemW + off, p.w_image + off, 16384, &bar_w);
    }
    __syncwarp();
    // Stagger: slot 1 starts streaming when slot 0 has finished phase A of its first tile, so
    // that one slot is in the HBM-bound phase while the other is in the MMA-bound phases.
    if (slot == 1 && tile0 < ntiles) TC_WAIT(&bar_stagger, 0);
    uint32_t seq = 0;
    int pend_col[2] = {-1, -1}, pend_row[2] = {0, 0};   // z_next box waiting in a stage, if any
    const uint32_t tx_bytes = 2u * (uint32_t)p.tile_rows * 128u;
    for (int64_t tile = tile0; tile < ntiles; tile += tile_step) {
      const int row0 = (int)(tile * p.tile_rows);
      for (int step = 0; step < 2 * nc; ++step, ++seq) {
        const bool is_c = step >= nc;          // phase-C re-read of the same chunks
        const int c = is_c ? step - nc : step;
        const uint32_t s = seq & 1, ph = (seq >> 1) & 1;
        TC_WAIT(&B.cons[s], ph ^ 1);
        TRACE(1);
        if (elect_one()) {
          uint8_t* dst = stage_base + s * kStageBytes;
          if (pend_col[s] >= 0) {
            // the stage holds a finished z_next box (written in place over z_prev): store it,
            // and wait until the TMA engine has read it before the stage is overwritten
            tma_store_2d_hint(tm_prev, dst + kBoxBytes, pend_col[s], pend_row[s], kEvictNormal);
            tma_store_commit();
            tma_store_wait_read<0>();
          }
          mbar_expect_tx(&B.full[s], tx_bytes);
          const uint64_t pol = is_c ? kEvictFirst : kEvictLast;
          tma_load_2d_hint(dst, tm_cur, c * kChunk, row0, &B.full[s], pol);
          tma_load_2d_hint(dst + kBoxBytes, tm_prev, c * kChunk, row0, &B.full[s], pol);
        }
        __syncwarp();
        pend_col[s] = is_c ? c * kChunk : -1;
        pend_row[s] = row0;
      }
    }
    // drain: the last two stages may still hold z_next boxes
    for (int e = 0; e < 2; ++e, ++seq) {
      const uint32_t s = seq & 1, ph = (seq >> 1) & 1;
      if (pend_col[s] < 0) continue;
      TC_WAIT(&B.cons[s], ph ^ 1);
      if (elect_one()) {
        tma_store_2d_hint(tm_prev, stage_base + s * kStageBytes + kBoxBytes, pend_col[s], pend_row[s],
                          kEvictNormal);
        tma_store_commit();
      }
      __syncwarp();
    }
    tma_store_wait_all<0>();
  } else if (warp < 4) {
    // ===================== MMA issuer of this slot =====================
    const uint32_t idesc1 = make_idesc(kFmtBF16, 128, kDP, 0, 0);   // B K-major  (GEMM1)
    const uint32_t idesc2 = make_idesc(kFmtBF16, 128, kQ, 0, 1);    // B MN-major (GEMM2)
    const uint32_t w_addr = smem_u32(smem + kSmemW);
    const uint64_t desc1 = make_smem_desc_sw128(w_addr, 0, 1024);
    const uint64_t desc2 = make_smem_desc_sw128(w_addr, kSlabBytes, 1024);
    const uint32_t d1_lo = (uint32_t)desc1, d1_hi = (uint32_t)(desc1 >> 32);
    const uint32_t d2_lo = (uint32_t)desc2, d2_hi = (uint32_t)(desc2 >> 32);
    constexpr uint32_t kPiece16 = kPieceBytes >> 4;
    auto make64 = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
    TC_WAIT(&bar_w, 0);
    uint32_t a_cnt0 = 0, a_cnt1 = 0, g_cnt0 = 0, g_cnt1 = 0, ti = 0;
    for (int64_t tile = tile0; tile < ntiles; tile += tile_step, ++ti) {
      // accumulators alias the G buffers of the previous tile: wait until both were drained
      TC_WAIT(&B.gfree[0], (g_cnt0 & 1) ^ 1);
      TC_WAIT(&B.gfree[1], (g_cnt1 & 1) ^ 1);
      TRACE(10);
      tc_fence_after();
      for (int c = 0; c < nc; ++c) {
        const int b = c & 1;
        if (b == 0) {
          TC_WAIT(&B.aready[0], a_cnt0 & 1);
          ++a_cnt0;
        } else {
          TC_WAIT(&B.aready[1], a_cnt1 & 1);
          ++a_cnt1;
        }
        TRACE(11);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t t_stage = tbase + kColStage2 + b * 48;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t koff = (uint32_t)(c >> 1) * (kSlabBytes >> 4) + (uint32_t)((c & 1) * 4 + ks * 2);
            const uint32_t acc_on = (c > 0 || ks > 0) ? 1u : 0u;
            const uint64_t q1 = make64(d1_lo + koff, d1_hi);
            const uint64_t q2 = make64(d1_lo + koff + kPiece16, d1_hi);
            const uint64_t q3 = make64(d1_lo + koff + 2 * kPiece16, d1_hi);
            const uint32_t p1 = t_stage + ks * 8, p2 = p1 + 16, p3 = p1 + 32;
            mma_ts<false>(tbase + kColAcc1_2, p1, q3, idesc1, acc_on);
            mma_ts<false>(tbase + kColAcc1_2, p2, q2, idesc1, 1);
            mma_ts<false>(tbase + kColAcc1_2, p3, q1, idesc1, 1);
            mma_ts<false>(tbase + kColAcc1_2, p1, q2, idesc1, 1);
            mma_ts<false>(tbase + kColAcc1_2, p2, q1, idesc1, 1);
            mma_ts<false>(tbase + kColAcc0_2, p1, q1, idesc1, acc_on);
          }
          if (c == nc - 1) mma_commit(&B.rfull);
          else mma_commit(&B.sfree[b]);
          if (slot == 0 && ti == 0 && c == nc - 1) mbar_arrive(&bar_stagger);
        }
        __syncwarp();
        TRACE(12);
      }
      TC_WAIT(&B.rready, ti & 1);
      TRACE(14);
      tc_fence_after();
      for (int q = 0; q < nq; ++q) {
        const int b = q & 1;
        if (b == 0) {
          TC_WAIT(&B.gfree[0], (g_cnt0 & 1) ^ 1);
          ++g_cnt0;
        } else {
          TC_WAIT(&B.gfree[1], (g_cnt1 & 1) ^ 1);
          ++g_cnt1;
        }
        TRACE(15);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t t_acc = tbase + (b ? kColAcc1_2 : kColAcc0_2);
          const uint32_t t_r = tbase + kColStage2;
          const uint32_t qoff = (uint32_t)q * (kSlabBytes >> 4);
          uint32_t acc_on = 0;
#pragma unroll
          for (int t = 0; t < 6; ++t) {
            constexpr int pa[6] = {2, 1, 0, 1, 0, 0}, pb[6] = {0, 1, 2, 0, 1, 0};
#pragma unroll
            for (int ks = 0; ks < dsteps; ++ks) {
              const uint64_t bd = make64(d2_lo + qoff + pb[t] * kPiece16 + ks * 128, d2_hi);
              mma_ts<false>(t_acc, t_r + pa[t] * 32 + ks * 8, bd, idesc2, acc_on);
              acc_on = 1;
            }
          }
          mma_commit(&B.gfull[b]);
        }
        __syncwarp();
        TRACE(16);
      }
    }
  } else {
    // ===================== compute warps of this slot =====================
    const int quad = warp & 3;                      // TMEM lane quadrant
    const int half = ((warp - 4) >> 2) & 1;         // which 16 of a chunk's 32 atoms / 32 of 64 features
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const float2 lr2 = make_float2(p.lr, p.lr);
    const float2 beta2 = make_float2(p.beta, p.beta);
    double dsum = 0.0;
    uint32_t seq = 0, g_cnt0 = 0, g_cnt1 = 0, ti = 0, sf_base0 = 0, sf_base1 = 0;
    const uint32_t sf_tile0 = (uint32_t)((nc + 1) / 2) - (((nc - 1) & 1) == 0 ? 1u : 0u);
    const uint32_t sf_tile1 = (uint32_t)(nc / 2) - (((nc - 1) & 1) == 1 ? 1u : 0u);
    for (int64_t tile = tile0; tile < ntiles;
         tile += tile_step, ++ti, sf_base0 += sf_tile0, sf_base1 += sf_tile1) {
      const int64_t grow = tile * p.tile_rows + row;
      const bool row_ok = row < p.tile_rows && grow < p.n;
      if (row_ok && half * 32 < p.d)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + grow * p.d + half * 32));
      // ---------------- phase A ----------------
      for (int c = 0; c < nc; ++c, ++seq) {
        const uint32_t s = seq & 1, ph = (seq >> 1) & 1;
        TC_WAIT(&B.full[s], ph);
        TRACE(20);
        const uint8_t* zc_s = stage_base + s * kStageBytes;
        const uint8_t* zp_s = zc_s + kBoxBytes;
        uint32_t w1[8], w2[8], w3[8];
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = sw128_offset(row, (half * 4 + j) * 16);
          const float4 zc = *reinterpret_cast<const float4*>(zc_s + off);
          float2 ya = make_float2(zc.x, zc.y), yc = make_float2(zc.z, zc.w);
          if (p.use_prev) {
            const float4 zp = *reinterpret_cast<const float4*>(zp_s + off);
            const float2 da = sub2(ya, make_float2(zp.x, zp.y));
            const float2 dc = sub2(yc, make_float2(zp.z, zp.w));
            part += (fabsf(da.x) + fabsf(da.y)) + (fabsf(dc.x) + fabsf(dc.y));
            ya = __fadd2_rn(ya, __fmul2_rn(beta2, da));
            yc = __fadd2_rn(yc, __fmul2_rn(beta2, dc));
          }
          split3_pair(ya, w1[2 * j], w2[2 * j], w3[2 * j]);
          split3_pair(yc, w1[2 * j + 1], w2[2 * j + 1], w3[2 * j + 1]);
        }
        mbar_arrive(&B.cons[s]);
        if (row_ok) dsum += (double)part;
        TRACE(21);
        const int b = c & 1;
        if (c >= 2)
          TC_WAIT(&B.sfree[b], ((b ? sf_base1 : sf_base0) + (uint32_t)(c >> 1) - 1u) & 1u);
        TRACE(22);
        tc_fence_after();
        const uint32_t t_stage = tbase + lane_base + kColStage2 + b * 48 + half * 8;
        tmem_st8(t_stage, w1);
        tmem_st8(t_stage + 16, w2);
        tmem_st8(t_stage + 32, w3);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&B.aready[b]);
        TRACE(23);
      }
      // ---------------- phase B: r = R - x, pieces of r (2 x 16 features per thread) ----------------
      {
        float4 xv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = half * 32 + 4 * j;
          xv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && col < p.d)
            xv[j] = __ldg(reinterpret_cast<const float4*>(p.x + grow * p.d + col));
        }
        TC_WAIT(&B.rfull, ti & 1);
        TRACE(30);
        tc_fence_after();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t rb[16], rs[16], w1[8], w2[8], w3[8];
          tmem_ld16(tbase + lane_base + kColAcc0_2 + half * 32 + hh * 16, rb);
          tmem_ld16(tbase + lane_base + kColAcc1_2 + half * 32 + hh * 16, rs);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 xq = xv[hh * 4 + j];
            const float2 ra = sub2(__fadd2_rn(make_float2(__uint_as_float(rb[4 * j + 0]), __uint_as_float(rb[4 * j + 1])),
                                              make_float2(__uint_as_float(rs[4 * j + 0]), __uint_as_float(rs[4 * j + 1]))),
                                   make_float2(xq.x, xq.y));
            const float2 rc = sub2(__fadd2_rn(make_float2(__uint_as_float(rb[4 * j + 2]), __uint_as_float(rb[4 * j + 3])),
                                              make_float2(__uint_as_float(rs[4 * j + 2]), __uint_as_float(rs[4 * j + 3]))),
                                   make_float2(xq.z, xq.w));
            split3_pair(ra, w1[2 * j], w2[2 * j], w3[2 * j]);
            split3_pair(rc, w1[2 * j + 1], w2[2 * j + 1], w3[2 * j + 1]);
          }
          // r pieces: piece p at columns [32 p, 32 p + 32), 2 features per column
          const uint32_t t_r = tbase + lane_base + kColStage2 + half * 16 + hh * 8;
          tmem_st8(t_r, w1);
          tmem_st8(t_r + 32, w2);
          tmem_st8(t_r + 64, w3);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&B.rready);
        TRACE(31);
      }
      // ---------------- phase C: re-read the codes, recompute y, fused update in place ----------------
      for (int c = 0; c < nc; ++c, ++seq) {
        const int q = c >> 1, gb = q & 1;
        if ((c & 1) == 0) {
          if (gb == 0) {
            TC_WAIT(&B.gfull[0], g_cnt0 & 1);
            ++g_cnt0;
          } else {
            TC_WAIT(&B.gfull[1], g_cnt1 & 1);
            ++g_cnt1;
          }
          TRACE(40);
          tc_fence_after();
        }
        uint32_t g[16];
        tmem_ld16(tbase + lane_base + (gb ? kColAcc1_2 : kColAcc0_2) + (c & 1) * 32 + half * 16, g);
        const uint32_t s = seq & 1, ph = (seq >> 1) & 1;
        TC_WAIT(&B.full[s], ph);
        uint8_t* zc_s = stage_base + s * kStageBytes;
        uint8_t* zp_s = zc_s + kBoxBytes;
        tmem_wait_ld();
        if ((c & 1) == 1 || c == nc - 1) {
          // both halves of this 64-atom accumulator have been read by this thread
          tc_fence_before();
          mbar_arrive(&B.gfree[gb]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = sw128_offset(row, (half * 4 + j) * 16);
          const float4 zc = *reinterpret_cast<const float4*>(zc_s + off);
          float2 ya = make_float2(zc.x, zc.y), yc = make_float2(zc.z, zc.w);
          if (p.use_prev) {
            const float4 zp = *reinterpret_cast<const float4*>(zp_s + off);
            ya = __fadd2_rn(ya, __fmul2_rn(beta2, sub2(ya, make_float2(zp.x, zp.y))));
            yc = __fadd2_rn(yc, __fmul2_rn(beta2, sub2(yc, make_float2(zp.z, zp.w))));
          }
          const float2 oa = ista_update_pair(
              ya, make_float2(__uint_as_float(g[4 * j + 0]), __uint_as_float(g[4 * j + 1])), lr2, p.lam);
          const float2 oc = ista_update_pair(
              yc, make_float2(__uint_as_float(g[4 * j + 2]), __uint_as_float(g[4 * j + 3])), lr2, p.lam);
          *reinterpret_cast<float4*>(zp_s + off) = make_float4(oa.x, oa.y, oc.x, oc.y);
        }
        fence_proxy_async_smem();   // in-place z_next box -> visible to the TMA store
        mbar_arrive(&B.cons[s]);
        TRACE(41);
      }
    }
    dsum = warp_sum(dsum);
    if (lane == 0) red[warp - 4] = dsum;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (warp == 4 && lane == 0 && p.use_prev && p.ctl.hist != nullptr && p.ctl.iter >= 1) {
      double s = 0.0;
      for (int i = 0; i < 16; ++i) s += red[i];
      atomicAdd(&p.ctl.hist[p.ctl.iter - 1], s);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tbase0, kTmemCols);
}
