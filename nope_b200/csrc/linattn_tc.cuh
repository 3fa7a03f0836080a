// nope_b200 -- LinearAttention core on tcgen05 (model_utils.py:403-417), heads = 4, dim_head = 32:
//   q = softmax_d(q) * scale ; k = softmax_n(k) ; ctx[d,e] = sum_n k[d,n] v[e,n] ; out[e,n] = sum_d ctx[d,e] q[d,n]
// qkv: [n_img, n_tok, 384] fp16 (q | k | v, each (head, 32)); out: [n_img, n_tok, 128] fp16; n_tok % 128 == 0.
//
// The SIMT kernel (kernels.cuh) runs both contractions on CUDA cores: 16.8 MFLOP per image is ~0.2 ms of
// fp32 FMA per 642-image launch before any memory time.  Here they are two tensor-core GEMMs per image with
// all four heads stacked (the off-diagonal head blocks are wasted work, 4x of a negligible amount):
//   ctx [128 d x 128 e] += ek^T [128 d x 128 tok] * v [128 tok x 128 e]          per 128-token tile
//   out [128 tok x 128 e] = qs [128 tok x 128 d] * ctxm [128 d x 128 e]          ctxm = block-diagonal ctx / ksum
// ek = exp(k - max_n k) and qs = softmax_d(q) * scale are written back IN PLACE over the TMA-loaded tiles
// by the SIMT warps (fp16), so the token-major tiles serve directly as operands: token-major [tok][channel]
// is the canonical MN-major SWIZZLE_128B layout for ek^T / v (M resp. N = channel is contiguous, K = token
// runs over rows) and the canonical K-major layout for qs (K = channel).
//
// One persistent CTA per SM; per image: pass 1 streams k (column max), pass 2 streams (k, v) (second read of
// k comes from L2), pass 3 builds ctxm, pass 4 streams q and stores out.  Warp roles: 0 TMA producer,
// 1 tcgen05.mma issuer, 2 TMEM allocator, 4..11 SIMT transforms + epilogues.
#pragma once
#include "conv_tc.cuh"

namespace nope {

constexpr int kLaSlots = 4;                       // ring of 32 KB slots: one [128 tok][128 ch] fp16 tile each
constexpr int kLaSlotBytes = 2 * kBM * 128;       // two 64-channel boxes
constexpr int kLaThreads = 384;
constexpr int kLaSimt = 256;

struct LinAttnParams {
  CUtensorMap qkv;        // [n_img][n_tok][384], box {64, 128, 1}
  CUtensorMap out;        // [n_img][n_tok][128], box {64, 128, 1}
  int n_img, n_tok;
  int bf16;               // qkv / out (and the ek / qs / ctxm operands written here) are bf16
};

struct LinAttnSmem {
  static constexpr int kRing = kLaSlots * kLaSlotBytes;          // 128 KB
  static constexpr int kCtxOff = kRing;                          // ctxm operand [128 e][128 d] fp16, 2 boxes
  static constexpr int kOutOff = kCtxOff + kLaSlotBytes;         // output staging, 2 boxes
  static constexpr int kBarOff = kOutOff + kLaSlotBytes;
  static constexpr int kVecOff = kBarOff + 256;                  // kmax[128], ksum[128] fp32, scratch [8][128]
  static constexpr int kTotal = kVecOff + (2 * 128 + 8 * 128) * 4 + 1024;
};

// MN-major SWIZZLE_128B operand descriptor: 64 MN-elements (128 B) contiguous, 8 K-rows per 1024-byte atom
// (SBO), the next 64 MN-elements 16 KB further (LBO = the second 64-channel box).
constexpr uint64_t kDescMN = (static_cast<uint64_t>(16384 >> 4) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
                             (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
// instruction descriptor, kind::f16, fp32 accumulate, M = N = 128; bit 15 / 16: A / B MN-major
constexpr uint32_t kIdescMN = make_idesc_f16(128, 128, false) | (1u << 15) | (1u << 16);
constexpr uint32_t kIdescK = make_idesc_f16(128, 128, false);

__global__ void __launch_bounds__(kLaThreads, 1) linattn_tc_kernel(const __grid_constant__ LinAttnParams p) {
  using S = LinAttnSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* s_ctx = smem + S::kCtxOff;
  uint8_t* s_out = smem + S::kOutOff;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarOff);   // [slots] TMA landed
  uint64_t* empty = full + kLaSlots;                                 // [slots] slot may be refilled
  uint64_t* ready = empty + kLaSlots;                                // [slots] transformed in place (SIMT -> MMA)
  uint64_t* ctx_full = ready + kLaSlots;                             // ctx accumulated (MMA -> SIMT)
  uint64_t* ctxm_ready = ctx_full + 1;                               // ctxm operand written (SIMT -> MMA)
  uint64_t* d_full = ctxm_ready + 1;                                 // [2] out tile accumulated
  uint64_t* d_empty = d_full + 2;                                    // [2] out tile drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);
  float* s_kmax = reinterpret_cast<float*>(smem + S::kVecOff);
  float* s_ksum = s_kmax + 128;
  float* s_scr = s_ksum + 128;                                       // [8 warps][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.n_tok / kBM;                                       // 128-token tiles per image

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.qkv);
    prefetch_tmap(&p.out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kLaSlots; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&ready[s], 1);
    }
    mbar_init(ctx_full, 1);
    mbar_init(ctxm_ready, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&d_full[a], 1);
      mbar_init(&d_empty[a], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();       // the prologue above overlapped the previous kernel's tail; its output is read from here on
  const uint32_t t_ctx = tmem_base;              // columns [0, 128)
  const uint32_t t_d0 = tmem_base + 128;         // out-tile accumulators: columns [128, 256), [256, 384)

  // Ring protocol: items are consumed in production order; item i lives in slot i % kLaSlots and EVERY
  // item completes exactly one phase of full[s] (TMA), ready[s] (SIMT warps: operand usable by the MMA) and
  // empty[s] (MMA commit, or the SIMT warps for pass-1 items), so all three flip with parity
  // (i / kLaSlots) & 1.  Per image the item sequence is
  //   T x k (pass 1) | T x (k, v) (pass 2) | T x q (pass 4)
  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t item = 0;
    auto load = [&](int ch0, int tok0, int img) {
      const int s = item % kLaSlots;
      mbar_wait(&empty[s], ((item / kLaSlots) & 1) ^ 1);
      if (elect_one()) {
        uint8_t* dst = smem + s * kLaSlotBytes;
        mbar_expect_tx(&full[s], kLaSlotBytes);
        tma_load_3d(dst, &p.qkv, &full[s], ch0, tok0, img);
        tma_load_3d(dst + kBM * 128, &p.qkv, &full[s], ch0 + 64, tok0, img);
      }
      __syncwarp();
      ++item;
    };
    for (int img = blockIdx.x; img < p.n_img; img += gridDim.x) {
      for (int t = 0; t < T; ++t) load(128, t * kBM, img);                 // k
      for (int t = 0; t < T; ++t) {
        load(128, t * kBM, img);                                           // k again (L2)
        load(256, t * kBM, img);                                           // v
      }
      for (int t = 0; t < T; ++t) load(0, t * kBM, img);                   // q
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t fmt = p.bf16 ? ((1u << 7) | (1u << 10)) : 0u;       // A / B format bits of the instruction descriptor
    uint32_t item = 0, n_img_done = 0, dcount = 0;
    for (int img = blockIdx.x; img < p.n_img; img += gridDim.x, ++n_img_done) {
      // parity waits only work for a thread that observes EVERY phase of a barrier: this warp walks the
      // pass-1 items too, and it (not the SIMT warps) hands their slots back, so the producer can never
      // run a phase ahead of it
      for (int t = 0; t < T; ++t, ++item) {
        mbar_wait(&full[item % kLaSlots], (item / kLaSlots) & 1);
        mbar_wait(&ready[item % kLaSlots], (item / kLaSlots) & 1);       // SIMT warps are done with the tile
        if (elect_one()) mbar_arrive(&empty[item % kLaSlots]);
        __syncwarp();
      }
      // ---- pass 2: ctx += ek^T v
      for (int t = 0; t < T; ++t) {
        const int sk = item % kLaSlots, sv = (item + 1) % kLaSlots;
        mbar_wait(&full[sk], (item / kLaSlots) & 1);
        mbar_wait(&ready[sk], (item / kLaSlots) & 1);                      // ek written in place
        mbar_wait(&full[sv], ((item + 1) / kLaSlots) & 1);                 // v landed
        mbar_wait(&ready[sv], ((item + 1) / kLaSlots) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = kDescMN | ((smem_base + sk * kLaSlotBytes) >> 4);
          const uint64_t bdesc = kDescMN | ((smem_base + sv * kLaSlotBytes) >> 4);
#pragma unroll
          for (int k = 0; k < kBM / 16; ++k)        // 16 tokens = two 8-row atoms = 2048 B
            umma_f16(t_ctx, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), kIdescMN | fmt, (t | k) != 0 ? 1u : 0u);
          umma_commit(&empty[sk]);
          umma_commit(&empty[sv]);
          if (t == T - 1) umma_commit(ctx_full);
        }
        __syncwarp();
        item += 2;
      }
      // ---- pass 4: out tile = qs ctxm
      mbar_wait(ctxm_ready, n_img_done & 1);
      for (int t = 0; t < T; ++t, ++dcount) {
        const int sq = item % kLaSlots, db = dcount & 1;
        mbar_wait(&full[sq], (item / kLaSlots) & 1);
        mbar_wait(&ready[sq], (item / kLaSlots) & 1);                      // qs written in place
        mbar_wait(&d_empty[db], ((dcount >> 1) & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = (smem_base + sq * kLaSlotBytes) >> 4;
          const uint32_t b_lo = (smem_base + S::kCtxOff) >> 4;
#pragma unroll
          for (int k = 0; k < 8; ++k) {             // K = 128 channels d: 4 steps of 16 per 64-channel box
            const uint32_t off = (k >> 2) * ((kBM * 128) >> 4) + (k & 3) * 2;
            umma_f16(t_d0 + db * 128, kDescHi | (a_lo + off), kDescHi | (b_lo + off), kIdescK | fmt, k != 0 ? 1u : 0u);
          }
          umma_commit(&empty[sq]);
          umma_commit(&d_full[db]);
        }
        __syncwarp();
        item += 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== SIMT transforms + epilogues (8 warps) =====================
    const int tid = threadIdx.x - 128, w8 = warp - 4;
    const int cx = tid & 15;                       // 16-byte chunk column: channels [8 cx, 8 cx + 8)
    const int r0 = tid >> 4;                       // token rows r0 + 16 i
    const int box = cx >> 3, cin = cx & 7;
    const int q4 = w8 & 3, ch = w8 >> 2;           // TMEM lane quarter / 64-column half of this warp
    const float scale = 0.17677669529663687f;      // 32^-0.5
    const bool bf = p.bf16 != 0;
    uint32_t item = 0, n_img_done = 0, dcount = 0;
    for (int img = blockIdx.x; img < p.n_img; img += gridDim.x, ++n_img_done) {
      // ---- pass 1: per-channel max of k over the image's tokens
      float mx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx[i] = -INFINITY;
      for (int t = 0; t < T; ++t, ++item) {
        const int s = item % kLaSlots;
        mbar_wait(&full[s], (item / kLaSlots) & 1);
        const uint8_t* tile = smem + s * kLaSlotBytes + box * (kBM * 128);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 16 * i;
          const uint4 v = *reinterpret_cast<const uint4*>(tile + r * 128 + ((cin ^ (r & 7)) << 4));
          const uint32_t* h2 = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) {
            const float2 f = unpack2(h2[k2], bf);
            mx[2 * k2] = fmaxf(mx[2 * k2], f.x);
            mx[2 * k2 + 1] = fmaxf(mx[2 * k2 + 1], f.y);
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) mbar_arrive(&ready[s]);        // the MMA warp returns the slot
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 16));
      if (lane < 16) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s_scr[w8 * 128 + cx * 8 + i] = mx[i];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid < 128) {
        float m = s_scr[tid];
        for (int w = 1; w < 8; ++w) m = fmaxf(m, s_scr[w * 128 + tid]);
        s_kmax[tid] = m;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float km[8], ks[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { km[i] = s_kmax[cx * 8 + i]; ks[i] = 0.f; }
      // ---- pass 2: ek = exp(k - max) in place; column sums of the stored values
      for (int t = 0; t < T; ++t, item += 2) {
        const int s = item % kLaSlots;
        mbar_wait(&full[s], (item / kLaSlots) & 1);
        uint8_t* tile = smem + s * kLaSlotBytes + box * (kBM * 128);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 16 * i;
          uint4* pv = reinterpret_cast<uint4*>(tile + r * 128 + ((cin ^ (r & 7)) << 4));
          uint4 v = *pv;
          uint32_t* h2 = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) {
            const float2 f = unpack2(h2[k2], bf);
            const uint32_t e2 = pack2(__expf(f.x - km[2 * k2]), __expf(f.y - km[2 * k2 + 1]), bf);
            const float2 b = unpack2(e2, bf);
            ks[2 * k2] += b.x;
            ks[2 * k2 + 1] += b.y;
            h2[k2] = e2;
          }
          *pv = v;
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) {
          mbar_arrive(&ready[s]);
          const int sv = (item + 1) % kLaSlots;               // v needs no transform: usable as it lands
          mbar_wait(&full[sv], ((item + 1) / kLaSlots) & 1);
          mbar_arrive(&ready[sv]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) ks[i] += __shfl_xor_sync(0xffffffffu, ks[i], 16);
      if (lane < 16) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s_scr[w8 * 128 + cx * 8 + i] = ks[i];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid < 128) {
        float a = 0.f;
        for (int w = 0; w < 8; ++w) a += s_scr[w * 128 + tid];
        s_ksum[tid] = a;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // ---- pass 3: ctxm[e][d] = ctx[d][e] / ksum[d] inside a head, 0 across heads; fp16, K-major (K = d)
      mbar_wait(ctx_full, n_img_done & 1);
      tc_fence_after();
      {
        const int d = q4 * 32 + lane;                           // TMEM lane = ctx row d; head of d = q4
        const float inv = 1.0f / s_ksum[d];
        uint32_t v[32];
        // previous image's out-tile MMAs have retired before ctx_full of this image can complete, so the
        // ctxm operand buffer is free to overwrite
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int e0 = ch * 64 + half * 32;                   // 32 columns e0..e0+31 of ctx row d
          tmem_ld_32x32(t_ctx + (static_cast<uint32_t>(q4 * 32) << 16) + e0, v);
          tmem_ld_wait();
          const bool same_head = (e0 >> 5) == q4;
          uint8_t* cbox = s_ctx + (d >> 6) * (kBM * 128);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int e = e0 + j;                               // row of the operand
            const float val = same_head ? __uint_as_float(v[j]) * inv : 0.f;
            st16(reinterpret_cast<__half*>(cbox + e * 128 + ((((d & 63) >> 3) ^ (e & 7)) << 4) + (d & 7) * 2), val, bf);
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid == 0) mbar_arrive(ctxm_ready);
      // ---- pass 4: qs = softmax_d(q) * scale in place; epilogue of the previous tile while the MMA runs
      auto drain = [&](uint32_t dc, int tok0) {
        const int db = dc & 1;
        mbar_wait(&d_full[db], (dc >> 1) & 1);
        tc_fence_after();
        if (tid == 0) tma_store_wait_read0();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int row = q4 * 32 + lane;
        uint32_t v[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tmem_ld_32x32(t_d0 + db * 128 + (static_cast<uint32_t>(q4 * 32) << 16) + ch * 64 + half * 32, v);
          tmem_ld_wait();
          uint8_t* srow = s_out + ch * (kBM * 128) + row * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 w;
            w.x = pack2(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]), bf);
            w.y = pack2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]), bf);
            w.z = pack2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]), bf);
            w.w = pack2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]), bf);
            *reinterpret_cast<uint4*>(srow + (((half * 4 + j) ^ (row & 7)) << 4)) = w;
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) {
          mbar_arrive(&d_empty[db]);
          tma_store_3d(&p.out, s_out, 0, tok0, img);
          tma_store_3d(&p.out, s_out + kBM * 128, 64, tok0, img);
          tma_store_commit();
        }
      };
      for (int t = 0; t < T; ++t, ++item) {
        const int s = item % kLaSlots;
        mbar_wait(&full[s], (item / kLaSlots) & 1);
        {
          // thread -> (token row, 64-channel box = head pair): two softmaxes over 32 channels each
          const int r = tid & 127, bx = tid >> 7;
          uint8_t* rowp = smem + s * kLaSlotBytes + bx * (kBM * 128) + r * 128;
#pragma unroll
          for (int hd = 0; hd < 2; ++hd) {
            uint4 v[4];
            float f[32];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              v[c4] = *reinterpret_cast<const uint4*>(rowp + (((hd * 4 + c4) ^ (r & 7)) << 4));
              const uint32_t* h2 = reinterpret_cast<const uint32_t*>(&v[c4]);
#pragma unroll
              for (int k2 = 0; k2 < 4; ++k2) {
                const float2 t2 = unpack2(h2[k2], bf);
                f[c4 * 8 + 2 * k2] = t2.x;
                f[c4 * 8 + 2 * k2 + 1] = t2.y;
              }
            }
            float m = f[0];
#pragma unroll
            for (int i = 1; i < 32; ++i) m = fmaxf(m, f[i]);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) { f[i] = __expf(f[i] - m); sum += f[i]; }
            const float qs = scale / sum;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              uint4 w;
              w.x = pack2(f[c4 * 8 + 0] * qs, f[c4 * 8 + 1] * qs, bf);
              w.y = pack2(f[c4 * 8 + 2] * qs, f[c4 * 8 + 3] * qs, bf);
              w.z = pack2(f[c4 * 8 + 4] * qs, f[c4 * 8 + 5] * qs, bf);
              w.w = pack2(f[c4 * 8 + 6] * qs, f[c4 * 8 + 7] * qs, bf);
              *reinterpret_cast<uint4*>(rowp + (((hd * 4 + c4) ^ (r & 7)) << 4)) = w;
            }
          }
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) mbar_arrive(&ready[s]);
        if (t > 0) { drain(dcount, (t - 1) * kBM); ++dcount; }
      }
      drain(dcount, (T - 1) * kBM);
      ++dcount;
    }
    if (tid == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// fp16 3-D tensor map [n_img][n_tok][C], box {64, 128, 1}, 128-byte swizzle
inline int make_token_map(CUtensorMap* out, const void* base, int n_img, int n_tok, int C) {
  uint64_t dims[3] = {(uint64_t)C, (uint64_t)n_tok, (uint64_t)n_img};
  uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)n_tok * C * 2};
  uint32_t box[3] = {64, (uint32_t)kBM, 1};
  return make_tmap_f16(out, base, 3, dims, str, box);
}

inline int launch_linattn_tc(const __half* qkv, __half* out, int n_img, int n_tok, int num_sms, cudaStream_t st,
                             bool bf16 = false) {
  if (n_tok % kBM != 0) return fail("linattn_tc: n_tok must be a multiple of 128");
  static bool attr_set[kMaxDevices];
  int dev = 0;
  NOPE_CUDA(cudaGetDevice(&dev));
  NOPE_CHECK(dev >= 0 && dev < kMaxDevices, "device index out of range");
  if (!attr_set[dev]) {
    NOPE_CUDA(cudaFuncSetAttribute(linattn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LinAttnSmem::kTotal));
    attr_set[dev] = true;
  }
  LinAttnParams p;
  if (make_token_map(&p.qkv, qkv, n_img, n_tok, 384) || make_token_map(&p.out, out, n_img, n_tok, 128)) return -1;
  p.n_img = n_img;
  p.n_tok = n_tok;
  p.bf16 = bf16 ? 1 : 0;
  const int grid = n_img < num_sms ? n_img : num_sms;
  NOPE_CUDA(launch_pdl(linattn_tc_kernel, dim3(grid), dim3(kLaThreads), LinAttnSmem::kTotal, st, p));
  return 0;
}

}  // namespace nope
