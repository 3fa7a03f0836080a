// nope_b200 -- 2-CTA (cta_group::2) variant of the implicit-GEMM convolution.
//
// Same math, parameters and epilogue as conv_tc_kernel (conv_tc.cuh); the difference is
// the tile: a CTA PAIR (thread-block cluster of 2, one TPC) owns a 256-pixel x BN tile.
// Each CTA TMA-loads its own 128 pixel rows of A and only HALF of the weight tile
// (BN/2 rows); one tcgen05.mma.cta_group::2 issued by the leader CTA multiplies both halves
// of A against the full weight tile (each SM reads the peer's weight half over the pair
// link), accumulating 128 x BN fp32 in each CTA's own TMEM.  Per K-step a CTA pulls
// 16 KB + BN*64 B from L2 instead of 16 KB + BN*128 B: the L2->SM traffic that bounds the
// 1-CTA kernel (77 FLOP/B at BN=192) drops by 30 % (110 FLOP/B), and the smaller stage
// buys a 6-deep ring.
//
// Tile widths: 192 (the default UNet: every width is a multiple of 192), 256 (the LDM variant:
// multiples of 256; a 128-wide tile reads 16 KB of A + 8 KB of B per 64 tensor-core clocks,
// exactly the 128 B/clk shared-memory limit, a 256-wide one 32 KB per 128), 128 / 64 (template
// encoder, GEGLU).  Tiles of <= 128 columns double-buffer the output staging.  EPI selects the
// epilogue at compile time: 0 plain (+ GroupNorm partial sums), 1 extras (ReLU, residual add,
// (hi, lo) split, fp32 store: template encoder, LDM out conv), 2 GEGLU (conv_tc.cuh).
//
// Protocol (per CTA unless noted; barriers live at identical smem offsets in both CTAs):
//   full[s]   leader only, count 2: leader's arrive.expect_tx(bytes of BOTH CTAs) + the
//             peer producer's remote arrive; both CTAs' TMA loads complete_tx on it
//             (cta_group::2 loads with the peer bit of the barrier address cleared)
//   empty[s]  count 1: tcgen05.commit multicast from the leader's MMA thread to both CTAs
//   tfull[a]  count 1: same multicast commit after the last K-step of a tile
//   tempty[a] leader only, count 8: the 4 epilogue warps of each CTA (peer: remote arrive)
#pragma once
#include "conv_tc.cuh"

namespace nope {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // address of the same offset in the pair's CTA 0
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar,
                                                int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask),
      "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar,
                                                int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask),
      "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once all prior MMAs of this thread retire) on the barrier at this offset in every
// CTA of `mask`
__device__ __forceinline__ void umma_commit_2cta_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
      " [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}

template <int BN, int STAGES>
struct Conv2Smem {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = (BN / 2) * kBK * 2;      // this CTA's half of the weight tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutBytes = (BN / 64) * kBM * 128;
  // narrow tiles double-buffer the output staging: the TMA store of tile i drains while the
  // epilogue of tile i+1 fills the other buffer (the short-K 1x1 layers are epilogue bound)
  static constexpr int kOutBufs = BN <= 128 ? 2 : 1;
  static constexpr int kBarOffset = STAGES * kStageBytes + kOutBufs * kOutBytes;
  static constexpr int kBiasOffset = kBarOffset + 256;
  static constexpr int kTotal = kBiasOffset + BN * 4 + 1024;
};

// EPI: 0 = plain epilogue (the sweep), 1 = extras (ReLU / residual / hi-lo / fp32), 2 = GEGLU
template <int BN, int STAGES, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv_tc2_kernel(const __grid_constant__ ConvParams p) {
  using S = Conv2Smem<BN, STAGES>;
  constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static_assert(BN % 64 == 0 && BN <= 256, "BN must be a multiple of 64");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* out_stage = smem + STAGES * S::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(smem + S::kBiasOffset);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m_pairs = (p.m_tiles + 1) >> 1;
  const int num_tiles = m_pairs * p.n_tiles;          // pair tiles
  const int tile0 = cluster_id_x(), tile_step = num_clusters_x();

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.n_amaps; ++i) prefetch_tmap(&p.amap[i]);
    prefetch_tmap(&p.bmap_half);
    for (int i = 0; i < p.n_par; ++i) prefetch_tmap(&p.omap[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 2 * kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2cta<kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();     // barriers of BOTH CTAs initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int m_pair = tile / p.n_tiles;
      const int n_tile = tile - m_pair * p.n_tiles;
      const int m_tile = 2 * m_pair + (int)rank;       // may be one past the end: TMA zero-fills
      const int par = n_tile / p.n_tiles_par;
      const int py = par >> 1, px = par & 1;
      int b0, y0;
      conv_tile_coords(p, m_tile, b0, y0);
      int kcol = 0;
      for (int s = 0; s < p.nseg; ++s) {
        const ConvSeg sg = p.seg[s];
        const CUtensorMap* am = &p.amap[sg.map];
        for (int ch = 0; ch < sg.nchunks; ++ch) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem + stage * S::kStageBytes;
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
            tma_load_4d_2sm(sa, am, &full_bar[stage], ch * kBK, sg.dx + px, y0 + sg.dy + py, b0);
            tma_load_2d_2sm(sa + S::kABytes, &p.bmap_half, &full_bar[stage], kcol,
                            n_tile * BN + (int)rank * (BN / 2));
            if (!leader) mbar_arrive_remote(&full_bar[stage], 0);
          }
          kcol += kBK;
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader CTA only) =====================
    constexpr uint32_t idesc = make_idesc_f16(2 * kBM, BN, false);
    const uint32_t smem_base = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int ks = 0; ks < p.ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = (smem_base + stage * S::kStageBytes) >> 4;
          const uint64_t adesc = kDescHi | a_lo;
          const uint64_t bdesc = kDescHi | (a_lo + (S::kABytes >> 4));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_f16_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (ks | k) != 0 ? 1u : 0u);
          umma_commit_2cta_mc(&empty_bar[stage], 3);
          if (ks == p.ksteps - 1) umma_commit_2cta_mc(&tfull_bar[acc], 3);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows, 8 warps) =====================
    const int e = warp - 4;
    const int etid = threadIdx.x - 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    int obuf = 0;
    constexpr bool kPrefetchRes = EPI == 1 && BN <= 128;   // 32 registers; wider tiles load in place
    ResPrefetch<kPrefetchRes ? BN : 64> pre;
    if constexpr (kPrefetchRes) {
      if (tile0 < num_tiles) {     // operands of this CTA's first tile (residual convs have n_par == 1)
        const int mp0 = tile0 / p.n_tiles;
        pre.load(p, 2 * mp0 + (int)rank, (tile0 - mp0 * p.n_tiles) * BN, e, lane);
      }
    }
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int m_pair = tile / p.n_tiles;
      const int n_tile = tile - m_pair * p.n_tiles;
      const int m_tile = 2 * m_pair + (int)rank;
      const int par = n_tile / p.n_tiles_par;
      const int n_chan0 = (n_tile - par * p.n_tiles_par) * BN;
      int b0, y0;
      conv_tile_coords(p, m_tile, b0, y0);
      if (etid < BN) s_bias[etid] = p.bias ? __ldg(p.bias + n_chan0 + etid) : 0.f;
      if constexpr (kPrefetchRes) {
        const int nt = tile + tile_step;
        const int nmp = nt / p.n_tiles;
        pre.next_m_tile = nt < num_tiles ? 2 * nmp + (int)rank : -1;
        pre.next_n_chan0 = (nt - nmp * p.n_tiles) * BN;
      }
      uint8_t* ost = out_stage + obuf * S::kOutBytes;
      if (etid == 0) {       // the store that last used this staging buffer has read it
        if constexpr (S::kOutBufs == 2) tma_store_wait_read1();
        else tma_store_wait_read0();
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if constexpr (EPI == 2)
        conv_epilogue_geglu(ost, s_bias, tmem_base + acc * BN, e, lane);
      else if constexpr (kPrefetchRes)
        conv_epilogue_tile<BN, true>(p, ost, s_bias, tmem_base + acc * BN, m_tile, n_chan0, e, lane, &pre);
      else if constexpr (EPI == 1)
        conv_epilogue_tile<BN, true>(p, ost, s_bias, tmem_base + acc * BN, m_tile, n_chan0, e, lane);
      else
        conv_epilogue_tile<BN, false>(p, ost, s_bias, tmem_base + acc * BN, m_tile, n_chan0, e, lane);
      // this CTA's accumulator half is drained: tell the leader's MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tempty_bar[acc]);
        else mbar_arrive_remote(&tempty_bar[acc], 0);
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (etid == 0) {
        if constexpr (EPI == 2) {
          tma_store_4d(&p.omap[0], ost, n_chan0 / 2, 0, y0, b0);
        } else {
#pragma unroll 1
          for (int cc = 0; cc < BN / 64; ++cc)
            tma_store_4d(&p.omap[par], ost + cc * (kBM * 128), n_chan0 + cc * 64, 0, y0, b0);
        }
        tma_store_commit();
      }
      obuf ^= S::kOutBufs - 1;
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (etid == 0) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();      // neither CTA may exit (or free TMEM) while its peer still uses it
  if (warp == 2) tmem_dealloc_2cta<kTmemCols>(tmem_base);
}

template <int BN, int STAGES, int EPI>
inline int launch_conv_tc2_t(const ConvParams& p, int num_sms, cudaStream_t stream) {
  using S = Conv2Smem<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    NOPE_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<BN, STAGES, EPI>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr_set = true;
  }
  const int pair_tiles = ((p.m_tiles + 1) / 2) * p.n_tiles;
  const int max_clusters = num_sms / 2;
  const int clusters = pair_tiles < max_clusters ? pair_tiles : max_clusters;
  conv_tc2_kernel<BN, STAGES, EPI><<<2 * clusters, kConvThreads, S::kTotal, stream>>>(p);
  NOPE_CUDA(cudaGetLastError());
  return 0;
}

inline int launch_conv_tc2(const ConvParams& p, int bn, int num_sms, cudaStream_t stream) {
  const bool ex = conv_needs_extras(p);
  if (p.geglu) {
    if (bn != 128 || ex || p.stats || p.n_par != 1) return fail("launch_conv_tc2: GEGLU epilogue needs BN = 128, no extras");
    return launch_conv_tc2_t<128, 6, 2>(p, num_sms, stream);
  }
  switch (bn) {
    // 256-wide tiles (LDM variant: every width is a multiple of 256): per K-step a CTA reads
    // 16 KB of A + 16 KB of B for 128 tensor-core clocks, against 16 + 8 KB for 64 clocks at
    // BN = 128, which sits exactly on the 128 B/clk shared-memory read limit
    case 256: return ex ? launch_conv_tc2_t<256, 5, 1>(p, num_sms, stream)
                        : launch_conv_tc2_t<256, 5, 0>(p, num_sms, stream);
    case 192: return ex ? launch_conv_tc2_t<192, 6, 1>(p, num_sms, stream)
                        : launch_conv_tc2_t<192, 6, 0>(p, num_sms, stream);
    case 128: return ex ? launch_conv_tc2_t<128, 6, 1>(p, num_sms, stream)
                        : launch_conv_tc2_t<128, 6, 0>(p, num_sms, stream);
    case 64: return ex ? launch_conv_tc2_t<64, 8, 1>(p, num_sms, stream)
                       : launch_conv_tc2_t<64, 8, 0>(p, num_sms, stream);
  }
  return fail("launch_conv_tc2: unsupported BN");
}

}  // namespace nope
