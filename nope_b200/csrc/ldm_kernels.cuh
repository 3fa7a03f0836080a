// nope_b200 -- kernels of the LDM-variant pose-conditioned UNet (UNetModelPose; reference:
// src/model/u_net/ldm/adapt_openaimodel.py:127-158, ldm/openaimodel.py:180-286 ResBlock,
// ldm/attention.py:149-277 CrossAttention / BasicTransformerBlock / SpatialTransformer).
//
// The GEMM-shaped work (3x3 / strided / 1x1 convolutions, q|k|v, to_out, GEGLU projections)
// runs on the tcgen05 implicit-GEMM kernel of conv_tc2.cuh.  This file holds what surrounds it:
//   * GroupNorm(32) statistics + apply over one or two (concatenated) NHWC sources,
//   * LayerNorm over channels (+ the cross-attention term, see ldm_ln_kernel),
//   * multi-head self-attention: QK^T and PV on tcgen05 (S and O accumulators in TMEM, softmax
//     in registers, P staged in shared memory as a swizzled K-major operand),
//   * the stand-alone GEGLU kernel (A/B switch; the default path fuses GEGLU into the GEMM
//     epilogue, conv_tc.cuh), the l2 score of the (tensor-core) output convolution.
// Activations are NHWC fp16 (tokens = pixels, so "b (h w) c" is the same memory); statistics,
// softmax, LayerNorm and scores are fp32.
#pragma once
#include "common.cuh"
#include "conv_tc.cuh"
#include "kernels.cuh"

namespace nope {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float2 t = __half22float2(hv[q]);
    f[2 * q] = t.x;
    f[2 * q + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_half2(f[0], f[1]), pack_half2(f[2], f[3]), pack_half2(f[4], f[5]),
                    pack_half2(f[6], f[7]));
}

// ----------------------------------------------------------------------------
// GroupNorm statistics of cat(x0, x1) along channels, in the conv-epilogue format
// (conv_tc.cuh: stats[(img * parts + part) * noct + octet] = (sum, sum of squares) over a
// 32-pixel segment x 8 channels; parts = max(1, hw/32)).  grid (parts, n_img), 256 threads.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ldm_stats_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1,
                 float2* __restrict__ stats, int hw) {
  __shared__ float2 s_red[256];
  const int part = blockIdx.x, img = blockIdx.y, parts = gridDim.x;
  const int noct = (C0 + C1) / 8;          // <= 256
  const int npx = hw < 32 ? hw : 32;
  const int rows = 256 / noct;             // >= 1
  const int t = threadIdx.x;
  const int o = t % noct, r = t / noct;
  if (r < rows) {
    const int c = o * 8;
    const __half* base = c < C0 ? x0 + c : x1 + (c - C0);
    const int Cs = c < C0 ? C0 : C1;
    base += ((size_t)img * hw + (size_t)part * 32) * Cs;
    float s = 0.f, ss = 0.f;
    for (int p = r; p < npx; p += rows) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(base + (size_t)p * Cs), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s += f[i];
        ss = fmaf(f[i], f[i], ss);
      }
    }
    s_red[t] = make_float2(s, ss);
  }
  __syncthreads();
  if (t < noct) {
    float s = 0.f, ss = 0.f;
    for (int rr = 0; rr < rows; ++rr) {
      s += s_red[rr * noct + t].x;
      ss += s_red[rr * noct + t].y;
    }
    stats[((size_t)img * parts + part) * noct + t] = make_float2(s, ss);
  }
}

// ----------------------------------------------------------------------------
// y = [SiLU](GroupNorm32(cat(x0, x1)))   (normalization() = GroupNorm32(32, C), eps 1e-5,
// ldm/util.py:187-204; SpatialTransformer.norm = GroupNorm(32, C, eps 1e-6),
// ldm/attention.py:73-76).  Statistics in the conv-epilogue format.  grid (nslab, n_img).
// ----------------------------------------------------------------------------
constexpr int kLdmMaxC = 2048;
struct LdmGnArgs {
  const __half* x0;
  const __half* x1;
  __half* y;
  const float2* stats;
  const float* gamma;
  const float* beta;
  int C0, C1, st_parts, hw, pps;   // pps: pixels per CTA
  float eps;
};

template <bool SILU>
__global__ void __launch_bounds__(256) ldm_gn_apply_kernel(const LdmGnArgs a) {
  __shared__ float s_scale[kLdmMaxC], s_shift[kLdmMaxC];
  __shared__ float2 s_red[256];
  __shared__ float2 s_grp[32];
  const int C = a.C0 + a.C1, noct = C / 8, opg = noct / 32;
  const int slab = blockIdx.x, img = blockIdx.y, t = threadIdx.x;
  {
    // fixed-order reduction of this image's partials: 8 threads per group
    const int g = t >> 3, li = t & 7;
    const int E = a.st_parts * opg;
    float s = 0.f, ss = 0.f;
    for (int e = li; e < E; e += 8) {
      const int part = e / opg, oo = e - part * opg;
      const float2 v = a.stats[((size_t)img * a.st_parts + part) * noct + g * opg + oo];
      s += v.x;
      ss += v.y;
    }
    s_red[t] = make_float2(s, ss);
  }
  __syncthreads();
  if (t < 32) {
    float s = 0.f, ss = 0.f;
    for (int i = 0; i < 8; ++i) {
      s += s_red[t * 8 + i].x;
      ss += s_red[t * 8 + i].y;
    }
    const float cnt = (float)a.hw * (float)(C / 32);
    const float mean = s / cnt;
    const float var = fmaxf(ss / cnt - mean * mean, 0.f);
    s_grp[t] = make_float2(mean, rsqrtf(var + a.eps));
  }
  __syncthreads();
  for (int c = t; c < C; c += 256) {
    const float2 mr = s_grp[c / (C / 32)];
    const float sc = mr.y * a.gamma[c];
    s_scale[c] = sc;
    s_shift[c] = a.beta[c] - mr.x * sc;
  }
  __syncthreads();
  const int total = a.pps * noct;
  const size_t pix0 = (size_t)img * a.hw + (size_t)slab * a.pps;
  for (int i0 = t; i0 < total; i0 += 256 * 4) {
    uint4 v[4];
    int pp[4], cc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 256;
      if (i < total) {
        const int p = i / noct, c = (i - p * noct) * 8;
        pp[u] = p;
        cc[u] = c;
        v[u] = c < a.C0 ? ld_stream16(a.x0 + (pix0 + p) * a.C0 + c)
                        : ld_stream16(a.x1 + (pix0 + p) * a.C1 + (c - a.C0));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 256;
      if (i >= total) break;
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float y = fmaf(f[k], s_scale[cc[u] + k], s_shift[cc[u] + k]);
        if (SILU) y = silu_f(y);
        f[k] = y;
      }
      *reinterpret_cast<uint4*>(a.y + (pix0 + pp[u]) * C + cc[u]) = pack8(f);
    }
  }
}

// ----------------------------------------------------------------------------
// LayerNorm over channels, one warp per token (nn.LayerNorm(dim), eps 1e-5,
// ldm/attention.py:218-220).  C = 256 * NV.
//
// With `cb` the kernel first adds the cross-attention term and writes the updated residual
// stream back:  x <- x + cb[img, :].  BasicTransformerBlock.attn2 attends to a context of ONE
// token (context = pose_mlp(pose).unsqueeze(1), adapt_openaimodel.py:147): the softmax over a
// single key is exactly 1 for every query, so attn2(norm2(x), ctx) = to_out(to_v(ctx)) does not
// depend on x -- a per-hypothesis channel vector, precomputed from the pose (ldm_cross_kernel).
// ----------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
ldm_ln_kernel(const __half* __restrict__ x, __half* __restrict__ xout, const float* __restrict__ cb,
              int cb_stride, const float* __restrict__ gamma, const float* __restrict__ beta,
              __half* __restrict__ y, long long n_tok, int tok_per_img, const int* __restrict__ src_img) {
  constexpr int C = 256 * NV;
  constexpr int TPW = 4 / NV;      // tokens in flight per warp: 1024 channels = 4 x 16-byte loads per lane
  const long long tok0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * TPW;
  const int lane = threadIdx.x & 31;
  if (tok0 >= n_tok) return;
  float f[TPW][NV][8];
  long long img[TPW];
  // src_img: x holds one image per REFERENCE (pose-independent prefix); hypothesis i reads image
  // src_img[i] of it.  Outputs are always per hypothesis.
#pragma unroll
  for (int t = 0; t < TPW; ++t) {
    const long long tok = tok0 + t < n_tok ? tok0 + t : n_tok - 1;    // clamp: tail tokens recompute the last one
    img[t] = tok / tok_per_img;
    const long long src_tok = src_img ? (long long)src_img[img[t]] * tok_per_img + (tok - img[t] * tok_per_img) : tok;
    const __half* xp = x + src_tok * C;
#pragma unroll
    for (int j = 0; j < NV; ++j)
      unpack8(*reinterpret_cast<const uint4*>(xp + (j * 32 + lane) * 8), f[t][j]);
  }
#pragma unroll
  for (int t = 0; t < TPW; ++t) {
    const long long tok = tok0 + t;
    if (tok >= n_tok) break;
    if (cb) {
      const float* cp = cb + (size_t)img[t] * cb_stride;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 c0 = *reinterpret_cast<const float4*>(cp + (j * 32 + lane) * 8);
        const float4 c1 = *reinterpret_cast<const float4*>(cp + (j * 32 + lane) * 8 + 4);
        f[t][j][0] += c0.x; f[t][j][1] += c0.y; f[t][j][2] += c0.z; f[t][j][3] += c0.w;
        f[t][j][4] += c1.x; f[t][j][5] += c1.y; f[t][j][6] += c1.z; f[t][j][7] += c1.w;
        const uint4 w = pack8(f[t][j]);
        *reinterpret_cast<uint4*>(xout + tok * C + (j * 32 + lane) * 8) = w;
        unpack8(w, f[t][j]);     // normalise what is stored (the residual the next layer adds)
      }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) s += f[t][j][i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = f[t][j][i] - mean;
        q = fmaf(d, d, q);
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / C) + 1e-5f);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 8;
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + c);
      const float4 g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + c);
      const float4 b1 = *reinterpret_cast<const float4*>(beta + c + 4);
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o8[i] = fmaf((f[t][j][i] - mean) * rstd, gm[i], bt[i]);
      *reinterpret_cast<uint4*>(y + tok * C + c) = pack8(o8);
    }
  }
}

// cross-attention terms of all transformer blocks: cb[h, :] = Wc pose[h] + bc with
// Wc = to_out.W to_v.W pose_mlp.W, bc = to_out.W to_v.W pose_mlp.b + to_out.b folded on the host
// in double (ldm/attention.py:170-195 with one context token).
__global__ void ldm_cross_kernel(const float* __restrict__ poses, const float* __restrict__ Wc,
                                 const float* __restrict__ bc, float* __restrict__ cb, int n_hyp,
                                 int rot_dim, int width) {
  const int h = blockIdx.x;
  __shared__ float sp[8];
  if (threadIdx.x < 8) sp[threadIdx.x] = threadIdx.x < rot_dim ? poses[(size_t)h * rot_dim + threadIdx.x] : 0.f;
  __syncthreads();
  for (int j = threadIdx.x; j < width; j += blockDim.x) {
    float a = bc[j];
    for (int i = 0; i < rot_dim; ++i) a = fmaf(Wc[(size_t)j * rot_dim + i], sp[i], a);
    cb[(size_t)h * width + j] = a;
  }
}

// ----------------------------------------------------------------------------
// GEGLU (ldm/attention.py:44-51): hg [tok][2*inner] = (x | gate) -> y [tok][inner] = x * gelu(gate),
// exact (erf) GELU as F.gelu's default.
// ----------------------------------------------------------------------------
__global__ void ldm_geglu_kernel(const __half* __restrict__ hg, __half* __restrict__ y,
                                 long long n_tok, int inner) {
  const int octs = inner / 8;
  const long long total = n_tok * octs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long tok = i / octs;
    const int o = (int)(i - tok * octs);
    const __half* row = hg + tok * 2 * inner;
    float a[8], g[8];
    unpack8(ld_stream16(row + o * 8), a);
    unpack8(ld_stream16(row + inner + o * 8), g);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] *= gelu_erf_fast(g[k]);
    *reinterpret_cast<uint4*>(y + tok * inner + o * 8) = pack8(a);
  }
}

// ----------------------------------------------------------------------------
// Self-attention operand staging.  Q and K are read by TMA straight out of qkv [img][n][3C]
// (q | k | v, each (head, 32)): a 64-channel box covers the heads (2i, 2i+1) as one 128-byte
// swizzle row, and head h uses the K-steps of its own 32 channels.  Only V needs a copy: the
// P V product wants it K-major, i.e. transposed:
//   Vt [img*H + h][32][n]
// grid (n / 64, n_img), 256 threads.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ldm_attn_prep_kernel(const __half* __restrict__ qkv, __half* __restrict__ Vt, int n, int C) {
  __shared__ __align__(16) __half sV[4][64][40];
  const int H = C / 32;
  const int tb = blockIdx.x, img = blockIdx.y, t = threadIdx.x;
  const size_t row0 = (size_t)img * n + (size_t)tb * 64;
  for (int h0 = 0; h0 < H; h0 += 4) {       // four heads per pass: 256-byte runs of every token row
    const int nh = min(4, H - h0);
    for (int i = t; i < 64 * 16; i += 256) {
      const int j = i & 15, tok = i >> 4;     // 16 x 16-byte chunks = 4 heads x 32 channels
      if ((j >> 2) < nh)
        *reinterpret_cast<uint4*>(&sV[j >> 2][tok][(j & 3) * 8]) =
            *reinterpret_cast<const uint4*>(qkv + (row0 + tok) * 3 * C + 2 * C + h0 * 32 + j * 8);
    }
    __syncthreads();
    for (int i = t; i < 4 * 32 * 8; i += 256) {
      const int tc = i & 7, d = (i >> 3) & 31, hh = i >> 8;
      if (hh < nh) {
        __align__(16) __half tmp[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) tmp[jj] = sV[hh][tc * 8 + jj][d];
        *reinterpret_cast<uint4*>(Vt + (((size_t)img * H + h0 + hh) * 32 + d) * n + (size_t)tb * 64 + tc * 8) =
            *reinterpret_cast<const uint4*>(tmp);
      }
    }
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------
// Multi-head self-attention core (CrossAttention.forward with context = x,
// ldm/attention.py:170-195): out = softmax(q k^T * d^-1/2) v per (image, head), d = 32.
//
// One CTA (128 threads) per (image*head, block of 128 queries); thread r owns query row r.
//   S = Q K^T : tcgen05.mma M=128 (queries) x N=128 (keys) x K=32, accumulator in TMEM cols 0..127
//   softmax   : each thread reads its row of S with tcgen05.ld (no shuffles), online max / sum,
//               P -> fp16 -> shared memory as the K-major SWIZZLE_128B A operand
//   O_blk = P V : tcgen05.mma M=128 x N=32 (head channels) x K=128 (keys), TMEM cols 128..159,
//               folded into the running O in registers: O = O * alpha + O_blk
// K / V^T blocks are double-buffered TMA loads.  The MMAs are software-pipelined against the
// softmax (see the loop) and two CTAs share an SM.
// ----------------------------------------------------------------------------
struct AttnParams {
  CUtensorMap qkmap, vmap;        // 3-D: qkv {3C, n, img} (box 64 ch x 128 tok) / Vt {n, 32, img*H}
  __half* out;                    // [img][n][C], head h at channels 32h..32h+31
  int n, H, C;
  float scale_log2e;              // d^-1/2 * log2(e)
};
constexpr int kAttnQBytes = 128 * 128, kAttnKBytes = 128 * 128, kAttnVBytes = 2 * 32 * 128,
              kAttnPBytes = 2 * 128 * 128;
constexpr int kAttnSmem = kAttnQBytes + 2 * kAttnKBytes + 2 * kAttnVBytes + kAttnPBytes + 256 + 1024;

// 2^x on the SFU (MUFU.EX2, ~2 ulp); exp2f() adds a denormal-range fix-up the softmax never needs.
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(128, 2) ldm_attn_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t attn_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(attn_smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kAttnQBytes;
  uint8_t* sV = sK + 2 * kAttnKBytes;
  uint8_t* sP = sV + 2 * kAttnVBytes;
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(sP + kAttnPBytes);
  uint64_t* bar_kv = bar_q + 1;   // [2]
  uint64_t* bar_s = bar_q + 3;
  uint64_t* bar_o = bar_q + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 8);

  // query blocks of one (image, head) are adjacent CTAs: they run together and share K / V^T in L2
  // (with the head-major order of the first version every query block re-read them from DRAM)
  const int nblk = (p.n + 127) / 128;
  const int bh = blockIdx.x / nblk, qb = blockIdx.x - bh * nblk;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int img = bh / p.H, head = bh - img * p.H;
  const int qc0 = (head >> 1) * 64, kc0 = p.C + qc0;      // channel of the head pair's 64-wide box
  const uint32_t kstep0 = (head & 1) * 4;                   // descriptor offset (>>4) of this head's 32 channels

  if (tid == 0) {
    prefetch_tmap(&p.qkmap);
    prefetch_tmap(&p.vmap);
    for (int i = 0; i < 5; ++i) mbar_init(bar_q + i, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem + 128;
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;

  constexpr uint32_t idesc_s = make_idesc_f16(128, 128, false);
  constexpr uint32_t idesc_o = make_idesc_f16(128, 32, false);

  auto issue_kv = [&](int jb) {      // TMA of K / V^T block jb into buffer jb & 1
    const int buf = jb & 1;
    const int valid = min(128, p.n - jb * 128);
    const int nat = (valid + 63) / 64;
    mbar_expect_tx(&bar_kv[buf], kAttnKBytes + nat * 4096);
    tma_load_3d(sK + buf * kAttnKBytes, &p.qkmap, &bar_kv[buf], kc0, jb * 128, img);
    for (int a = 0; a < nat; ++a)
      tma_load_3d(sV + buf * kAttnVBytes + a * 4096, &p.vmap, &bar_kv[buf], jb * 128 + a * 64, 0, bh);
  };
  auto issue_qk = [&](int jb) {      // S = Q K_jb^T (K = 32 real head channels: two 16-wide steps)
    const int buf = jb & 1;
    mbar_wait(&bar_kv[buf], (jb >> 1) & 1);
    tc_fence_after();
    const uint64_t adesc = (kDescHi | (smem_u32(sQ) >> 4)) + kstep0;
    const uint64_t bdesc = (kDescHi | (smem_u32(sK + buf * kAttnKBytes) >> 4)) + kstep0;
    umma_f16(tS, adesc, bdesc, idesc_s, 0u);
    umma_f16(tS, adesc + 2, bdesc + 2, idesc_s, 1u);
    umma_commit(bar_s);
  };
  if (tid == 0) {
    mbar_expect_tx(bar_q, kAttnQBytes);
    tma_load_3d(sQ, &p.qkmap, bar_q, qc0, qb * 128, img);
    issue_kv(0);
    if (nblk > 1) issue_kv(1);
    mbar_wait(bar_q, 0);
    issue_qk(0);
  }
  __syncwarp();

  // Software pipeline: while the threads run the softmax of block j, the tensor core finishes
  // P V of block j-1 and (issued right after it) Q K^T of block j+1; O_blk(j-1) is folded into
  // the running O in the middle of iteration j, when its MMA has long retired.
  float O[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) O[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
  const float c = p.scale_log2e;
  uint8_t* prow = sP + tid * 128;

  for (int j = 0; j < nblk; ++j) {
    const int b = j & 1;
    const int valid = min(128, p.n - j * 128);
    const int nat = (valid + 63) / 64;
    const bool full = valid == 128;

    mbar_wait(bar_s, j & 1);
    tc_fence_after();
    uint32_t s[128];
    tmem_ld_32x32(tS + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
    tmem_ld_32x32(tS + lane_off + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
    tmem_ld_32x32(tS + lane_off + 64, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
    tmem_ld_32x32(tS + lane_off + 96, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
    tmem_ld_wait();
    tc_fence_before();      // S is in registers: Q K^T of the next block may overwrite it after the barrier

    float bm;
    if (full) {
      float m0 = __uint_as_float(s[0]), m1 = __uint_as_float(s[1]), m2 = __uint_as_float(s[2]),
            m3 = __uint_as_float(s[3]);
#pragma unroll
      for (int i = 4; i < 128; i += 4) {
        m0 = fmaxf(m0, __uint_as_float(s[i]));
        m1 = fmaxf(m1, __uint_as_float(s[i + 1]));
        m2 = fmaxf(m2, __uint_as_float(s[i + 2]));
        m3 = fmaxf(m3, __uint_as_float(s[i + 3]));
      }
      bm = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    } else {
      bm = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; ++i)
        if (i < valid) bm = fmaxf(bm, __uint_as_float(s[i]));
    }
    const float m_new = fmaxf(m_run, bm);
    const float alpha = fast_exp2((m_run - m_new) * c);     // 0 on the first block (m_run = -inf)
    const float mc = m_new * c;
    float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
#pragma unroll
    for (int i = 0; i < 128; i += 4) {
      float e0 = fast_exp2(fmaf(__uint_as_float(s[i]), c, -mc));
      float e1 = fast_exp2(fmaf(__uint_as_float(s[i + 1]), c, -mc));
      float e2 = fast_exp2(fmaf(__uint_as_float(s[i + 2]), c, -mc));
      float e3 = fast_exp2(fmaf(__uint_as_float(s[i + 3]), c, -mc));
      if (!full) {
        if (i >= valid) e0 = 0.f;
        if (i + 1 >= valid) e1 = 0.f;
        if (i + 2 >= valid) e2 = 0.f;
        if (i + 3 >= valid) e3 = 0.f;
      }
      // row sum in fp32 of the unrounded probabilities (the fp16 rounding of P is unbiased)
      rs0 += e0; rs1 += e1; rs2 += e2; rs3 += e3;
      s[i] = __float_as_uint(e0);
      s[i + 1] = __float_as_uint(e1);
      s[i + 2] = __float_as_uint(e2);
      s[i + 3] = __float_as_uint(e3);
    }
    l_run = fmaf(l_run, alpha, (rs0 + rs1) + (rs2 + rs3));
    m_run = m_new;

    if (j > 0) {
      // fold in O_blk(j-1); its completion also frees sP and K/V buffer (j-1)&1 = (j+1)&1
      mbar_wait(bar_o, (j - 1) & 1);
      tc_fence_after();
      if (tid == 0 && j + 1 < nblk) issue_kv(j + 1);
      __syncwarp();
      uint32_t v[32];
      tmem_ld_32x32(tO + lane_off, v);
      tmem_ld_wait();
      tc_fence_before();
#pragma unroll
      for (int i = 0; i < 32; ++i) O[i] = fmaf(O[i], alpha_prev, __uint_as_float(v[i]));
    }
    alpha_prev = alpha;

    // P -> fp16 -> swizzled K-major rows of sP (atoms of 64 keys)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      if (ch * 32 < nat * 64) {
        uint8_t* arow = prow + (ch >> 1) * (128 * 128);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int i0 = ch * 32 + g * 8;
          const uint4 w = make_uint4(
              pack_half2(__uint_as_float(s[i0]), __uint_as_float(s[i0 + 1])),
              pack_half2(__uint_as_float(s[i0 + 2]), __uint_as_float(s[i0 + 3])),
              pack_half2(__uint_as_float(s[i0 + 4]), __uint_as_float(s[i0 + 5])),
              pack_half2(__uint_as_float(s[i0 + 6]), __uint_as_float(s[i0 + 7])));
          const int chunk = (ch & 1) * 4 + g;
          *reinterpret_cast<uint4*>(arow + ((chunk ^ (tid & 7)) * 16)) = w;
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      for (int a = 0; a < nat; ++a) {
        const uint64_t adesc = kDescHi | (smem_u32(sP + a * (128 * 128)) >> 4);
        const uint64_t bdesc = kDescHi | (smem_u32(sV + b * kAttnVBytes + a * 4096) >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tO, adesc + 2 * k, bdesc + 2 * k, idesc_o, (a | k) != 0 ? 1u : 0u);
      }
      umma_commit(bar_o);
      if (j + 1 < nblk) issue_qk(j + 1);
    }
    __syncwarp();
  }

  {
    mbar_wait(bar_o, (nblk - 1) & 1);
    tc_fence_after();
    uint32_t v[32];
    tmem_ld_32x32(tO + lane_off, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) O[i] = fmaf(O[i], alpha_prev, __uint_as_float(v[i]));
  }
  const int q_row = qb * 128 + tid;
  if (q_row < p.n) {
    const float inv = 1.f / l_run;
    __half* dst = p.out + ((size_t)img * p.n + q_row) * p.C + head * 32;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float o8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o8[i] = O[g * 8 + i] * inv;
      *reinterpret_cast<uint4*>(dst + g * 8) = pack8(o8);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

// CUDA-core twin of ldm_attn_tc_kernel on the same staged operands (debug / bring-up:
// nope_ldm_set_impl(.., attn_impl = 1)); same grid, thread r owns query row r.
__global__ void __launch_bounds__(128)
ldm_attn_simt_kernel(const __half* __restrict__ qkv, const __half* __restrict__ Vt,
                     __half* __restrict__ out, int n, int H, int C, float scale_log2e) {
  __shared__ float sK[64][33];
  __shared__ float sVt[32][65];
  const int nqb = (n + 127) / 128;
  const int bh = blockIdx.x / nqb, qb = blockIdx.x - bh * nqb, tid = threadIdx.x;
  const int q_row = qb * 128 + tid;
  float q[32], O[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { q[i] = 0.f; O[i] = 0.f; }
  if (q_row < n) {
    const __half* qp = qkv + ((size_t)(bh / H) * n + q_row) * 3 * C + (bh % H) * 32;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(qp + g * 8), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) q[g * 8 + i] = f[i];
    }
  }
  float m_run = -INFINITY, l_run = 0.f;
  for (int k0 = 0; k0 < n; k0 += 64) {
    __syncthreads();
    for (int i = tid; i < 64 * 32; i += 128) {
      const int kk = i >> 5, d = i & 31;
      sK[kk][d] = (k0 + kk < n)
                      ? __half2float(qkv[((size_t)(bh / H) * n + k0 + kk) * 3 * C + C + (bh % H) * 32 + d])
                      : 0.f;
    }
    for (int i = tid; i < 32 * 64; i += 128) {
      const int d = i >> 6, kk = i & 63;
      sVt[d][kk] = (k0 + kk < n) ? __half2float(Vt[((size_t)bh * 32 + d) * n + k0 + kk]) : 0.f;
    }
    __syncthreads();
    const int valid = min(64, n - k0);
    float s[64];
    float bm = -INFINITY;
#pragma unroll 4
    for (int kk = 0; kk < 64; ++kk) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) a = fmaf(q[d], sK[kk][d], a);
      s[kk] = a;
      if (kk < valid) bm = fmaxf(bm, a);
    }
    const float m_new = fmaxf(m_run, bm);
    const float alpha = exp2f((m_run - m_new) * scale_log2e);
    float rs = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) O[d] *= alpha;
#pragma unroll 4
    for (int kk = 0; kk < 64; ++kk) {
      const float pe = kk < valid ? exp2f((s[kk] - m_new) * scale_log2e) : 0.f;
      rs += pe;
#pragma unroll
      for (int d = 0; d < 32; ++d) O[d] = fmaf(pe, sVt[d][kk], O[d]);
    }
    l_run = fmaf(l_run, alpha, rs);
    m_run = m_new;
  }
  if (q_row < n) {
    const float inv = 1.f / l_run;
    __half* dst = out + ((size_t)(bh / H) * n + q_row) * C + (bh % H) * 32;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float o8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o8[i] = O[g * 8 + i] * inv;
      *reinterpret_cast<uint4*>(dst + g * 8) = pack8(o8);
    }
  }
}

// 3-D fp16 tensor map with 128-byte swizzle (attention operands).
inline int make_tmap3_f16(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                          uint32_t b0, uint32_t b1) {
  uint64_t dims[3] = {d0, d1, d2};
  uint64_t str[2] = {d0 * 2, d0 * d1 * 2};
  uint32_t box[3] = {b0, b1, 1};
  return make_tmap_f16(m, base, 3, dims, str, box);
}

// ----------------------------------------------------------------------------
// UNetModel.out[2] (conv3x3, C -> Cl <= 8 latent channels, openaimodel.py:722-726) runs on the
// tensor-core kernel with its output channels padded to 64 and an fp32 epilogue store
// (ConvParams.out_f32); this kernel picks the Cl real channels out of o [n_hyp*hw][64] fp32,
// writes the embeddings (NCHW fp32) and the reference's "l2" score partials
// (model.py:260-262), same outputs as final_conv_score_kernel.  grid (hw / 128, n_hyp).
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(kFinalThreads)
ldm_score_kernel(const float* __restrict__ o, float* __restrict__ emb, const float* __restrict__ query,
                 const int* __restrict__ ref_of, float* __restrict__ partial, int hw, int Cl) {
  __shared__ float s_part[kFinalThreads / 32];
  const int slab = blockIdx.x, h = blockIdx.y, nslab = gridDim.x;
  const int p = slab * kFinalThreads + threadIdx.x;
  float dist = 0.f;
  if (p < hw) {
    const float* op = o + ((size_t)h * hw + p) * 64;
    float acc[kMaxLatent];
    const float4 a0 = *reinterpret_cast<const float4*>(op);
    const float4 a1 = *reinterpret_cast<const float4*>(op + 4);
    acc[0] = a0.x; acc[1] = a0.y; acc[2] = a0.z; acc[3] = a0.w;
    acc[4] = a1.x; acc[5] = a1.y; acc[6] = a1.z; acc[7] = a1.w;
    if (emb) {
#pragma unroll
      for (int cc = 0; cc < kMaxLatent; ++cc)
        if (cc < Cl) emb[((size_t)h * Cl + cc) * hw + p] = acc[cc];
    }
    if (query) {
      const float* qp = query + (size_t)ref_of[h] * Cl * hw + p;
      float s4 = 0.f;
#pragma unroll
      for (int cc = 0; cc < kMaxLatent; ++cc)
        if (cc < Cl) {
          const float d = qp[(size_t)cc * hw] - acc[cc];
          const float d2 = d * d;
          s4 = fmaf(d2, d2, s4);
        }
      dist = sqrtf(s4);
    }
  }
  if (partial) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dist += __shfl_xor_sync(0xffffffffu, dist, off);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = dist;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < kFinalThreads / 32; ++i) t += s_part[i];
      partial[(size_t)h * nslab + slab] = t;
    }
  }
}

}  // namespace nope
