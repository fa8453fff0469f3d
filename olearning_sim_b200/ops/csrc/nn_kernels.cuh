// nn_kernels.cuh — the bandwidth-bound kernels around the batched tcgen05 GEMMs of the ResNet-18 / BERT-mini local
// step: GroupNorm and LayerNorm forward / backward, embedding gather + scatter-SGD, classifier heads with fused
// softmax cross-entropy, bias column sums and the fused SGD / FedProx step on the "small" parameters (norm scales,
// biases, heads) whose gradients are accumulated in a per-client fp32 side buffer.
//
// Conventions: activations are bf16, row-major [client][sample][...][channels]; parameters are read as fp32 from the
// packed master rows master[c * P + offset]; `active[c] == 0` skips a client (offline / straggler this step);
// `gsmall[c * gs_stride + i]` accumulates the gradient of master[c * P + small_off + i].
// Reference parity: this is (part of) the user training script the reference runs once per simulated phone
// (ols_core/taskMgr/utils/utils_run_task.py:481-514).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "bgemm_tc.cuh"

namespace nnk {
using bg::pack8;
using bg::unpack8;
typedef __nv_bfloat16 bf16;

constexpr float kGnEps = 1e-5f;
constexpr float kLnEps = 1e-5f;

__device__ __forceinline__ void load_f8(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ---------------------------------------------------------------------------------------------------------
// stem im2col: uint8 [N][32][32][3] -> bf16 [N][32][32][32], channel k = (kh*3+kw)*3 + c holds the normalised pixel
// x[h+kh-1][w+kw-1][c] (zero outside, k >= 27 zero).  Done once per data upload: the 3x3x3 stem then runs as a 1-tap
// K=32 GEMM instead of wasting 61/64 of every 64-channel chunk.
// ---------------------------------------------------------------------------------------------------------
__global__ void stem_im2col_kernel(const uint8_t* __restrict__ src, bf16* __restrict__ dst, size_t npix) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const int w = i & 31, h = (i >> 5) & 31;
  const uint8_t* img = src + (i - (size_t)h * 32 - w) * 3;
  float v[32];
#pragma unroll
  for (int k = 27; k < 32; ++k) v[k] = 0.f;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int hh = h + kh - 1, ww = w + kw - 1;
      const bool in = hh >= 0 && hh < 32 && ww >= 0 && ww < 32;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        v[(kh * 3 + kw) * 3 + c] = in ? (static_cast<float>(img[(hh * 32 + ww) * 3 + c]) * (1.f / 255.f) - 0.5f) * 2.f : 0.f;
    }
  uint4* o = reinterpret_cast<uint4*>(dst + i * 32);
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = pack8(v + 8 * j);
}

// ---------------------------------------------------------------------------------------------------------
// GroupNorm forward:  out = relu?( gn(raw; stats, gamma, beta) [+ gn(raw2; stats2, gamma2, beta2)] [+ res] )
// `stats` = (sum, sum of squares) per (image, group), accumulated by the E_GNSTAT epilogue of the convolution.
// ---------------------------------------------------------------------------------------------------------
struct GnApplyParams {
  int C, imgs, HW, Ch, groups; const uint8_t* active;
  const bf16* raw; const float* stats; int gamma_off, beta_off;
  const bf16* raw2; const float* stats2; int gamma2_off, beta2_off;
  const bf16* res;
  const float* master; long long P;
  int relu; bf16* out;
};

__global__ void gn_apply_kernel(const GnApplyParams p) {
  const int c = blockIdx.z, b = blockIdx.y;
  if (p.active && !p.active[c]) return;
  __shared__ float s_mu[2][8], s_rs[2][8];
  const int cpg = p.Ch / p.groups;
  const float inv_n = 1.f / (float)(p.HW * cpg);
  if (threadIdx.x < p.groups * 2) {
    const int which = threadIdx.x / p.groups, g = threadIdx.x % p.groups;
    const float* st = which ? p.stats2 : p.stats;
    if (st) {
      const float s = st[((size_t)(c * p.imgs + b) * p.groups + g) * 2], q = st[((size_t)(c * p.imgs + b) * p.groups + g) * 2 + 1];
      const float mu = s * inv_n;
      const float var = fmaxf(q * inv_n - mu * mu, 0.f);
      s_mu[which][g] = mu; s_rs[which][g] = rsqrtf(var + kGnEps);
    }
  }
  __syncthreads();
  const int vpi = p.HW * p.Ch / 8, vpc = p.Ch / 8;
  const size_t base = (size_t)(c * p.imgs + b) * p.HW * p.Ch;
  const float* mrow = p.master + (size_t)c * p.P;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < vpi; v += gridDim.x * blockDim.x) {
    const int ch = (v % vpc) * 8, g = ch / cpg;
    float x[8], ga[8], be[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(p.raw + base + (size_t)v * 8), x);
    load_f8(mrow + p.gamma_off + ch, ga); load_f8(mrow + p.beta_off + ch, be);
    const float mu = s_mu[0][g], rs = s_rs[0][g];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = (x[i] - mu) * rs * ga[i] + be[i];
    if (p.raw2) {
      unpack8(*reinterpret_cast<const uint4*>(p.raw2 + base + (size_t)v * 8), x);
      load_f8(mrow + p.gamma2_off + ch, ga); load_f8(mrow + p.beta2_off + ch, be);
      const float mu2 = s_mu[1][g], rs2 = s_rs[1][g];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += (x[i] - mu2) * rs2 * ga[i] + be[i];
    }
    if (p.res) {
      unpack8(*reinterpret_cast<const uint4*>(p.res + base + (size_t)v * 8), x);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += x[i];
    }
    if (p.relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaxf(o[i], 0.f);
    }
    *reinterpret_cast<uint4*>(p.out + base + (size_t)v * 8) = pack8(o);
  }
}

// ---------------------------------------------------------------------------------------------------------
// GroupNorm backward, pass 1 (one CTA per image): with g = dy * [mask > 0], xhat = (raw - mu) * rstd
//   dgamma[ch] += sum g * xhat,  dbeta[ch] += sum g                         (atomics into gsmall, summed over images)
//   sb[(c, b, group)] = (mean(g * gamma), mean(g * gamma * xhat))             (consumed by pass 2)
// ---------------------------------------------------------------------------------------------------------
struct GnBwdParams {
  int C, imgs, HW, Ch, groups; const uint8_t* active;
  const bf16* dy; const bf16* mask; const bf16* raw; const float* stats;
  const float* master; long long P; int gamma_off;
  float* gsmall; long long gs_stride; int g_gamma_off, g_beta_off;   // offsets inside the small-gradient row
  float* sb;                                                          // [C*imgs][groups][2]
  bf16* dx; bf16* dres;                                               // pass 2 outputs (dres optional: g itself)
};

__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const GnBwdParams p) {
  const int c = blockIdx.y, b = blockIdx.x;
  if (p.active && !p.active[c]) return;
  __shared__ float s_mu[8], s_rs[8], s_S[8][2];
  __shared__ float s_dg[512], s_db[512];
  const int cpg = p.Ch / p.groups;
  const float inv_n = 1.f / (float)(p.HW * cpg);
  for (int i = threadIdx.x; i < p.Ch; i += blockDim.x) { s_dg[i] = 0.f; s_db[i] = 0.f; }
  if (threadIdx.x < p.groups) {
    const float* st = p.stats + ((size_t)(c * p.imgs + b) * p.groups + threadIdx.x) * 2;
    const float mu = st[0] * inv_n;
    s_mu[threadIdx.x] = mu; s_rs[threadIdx.x] = rsqrtf(fmaxf(st[1] * inv_n - mu * mu, 0.f) + kGnEps);
    s_S[threadIdx.x][0] = 0.f; s_S[threadIdx.x][1] = 0.f;
  }
  __syncthreads();
  const int vpc = p.Ch / 8;                      // 8, 16, 32 or 64 vector columns
  const int cv = threadIdx.x % vpc, pl = threadIdx.x / vpc, npl = blockDim.x / vpc;
  const int ch = cv * 8, g = ch / cpg;
  const size_t base = (size_t)(c * p.imgs + b) * p.HW * p.Ch;
  float ga[8], dg[8], db[8], S1 = 0.f, S2 = 0.f;
  load_f8(p.master + (size_t)c * p.P + p.gamma_off + ch, ga);
#pragma unroll
  for (int i = 0; i < 8; ++i) dg[i] = db[i] = 0.f;
  const float mu = s_mu[g], rs = s_rs[g];
  for (int pix = pl; pix < p.HW; pix += npl) {
    const size_t o = base + (size_t)pix * p.Ch + ch;
    float d[8], x[8], m[8];
    unpack8(*reinterpret_cast<const uint4*>(p.dy + o), d);
    unpack8(*reinterpret_cast<const uint4*>(p.raw + o), x);
    if (p.mask) {
      unpack8(*reinterpret_cast<const uint4*>(p.mask + o), m);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = m[i] > 0.f ? d[i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (x[i] - mu) * rs;
      dg[i] += d[i] * xh; db[i] += d[i];
      const float gg = d[i] * ga[i];
      S1 += gg; S2 += gg * xh;
    }
  }
  atomicAdd(&s_S[g][0], S1); atomicAdd(&s_S[g][1], S2);
#pragma unroll
  for (int i = 0; i < 8; ++i) { atomicAdd(&s_dg[ch + i], dg[i]); atomicAdd(&s_db[ch + i], db[i]); }
  __syncthreads();
  float* gr = p.gsmall + (size_t)c * p.gs_stride;
  for (int i = threadIdx.x; i < p.Ch; i += blockDim.x) {
    atomicAdd(gr + p.g_gamma_off + i, s_dg[i]);
    atomicAdd(gr + p.g_beta_off + i, s_db[i]);
  }
  if (threadIdx.x < p.groups) {
    float* o = p.sb + ((size_t)(c * p.imgs + b) * p.groups + threadIdx.x) * 2;
    o[0] = s_S[threadIdx.x][0] * inv_n; o[1] = s_S[threadIdx.x][1] * inv_n;
  }
}

// Both passes in one kernel, one CTA per image: the masked gradient is parked in shared memory (bf16, HW x Ch) during the
// reduction and the second pass re-reads only `raw` — from L2, the image was streamed through it microseconds earlier and
// at most one or two images per SM are in flight.  HBM traffic 3 reads + 2 writes instead of 6 + 2.
__global__ void __launch_bounds__(512) gn_bwd_fused_kernel(const GnBwdParams p) {
  extern __shared__ uint4 s_d[];                 // [HW * Ch / 8] masked gradient
  const int c = blockIdx.y, b = blockIdx.x;
  if (p.active && !p.active[c]) return;
  __shared__ float s_mu[8], s_rs[8], s_S[8][2];
  __shared__ float s_dg[512], s_db[512];
  const int cpg = p.Ch / p.groups;
  const float inv_n = 1.f / (float)(p.HW * cpg);
  for (int i = threadIdx.x; i < p.Ch; i += blockDim.x) { s_dg[i] = 0.f; s_db[i] = 0.f; }
  if (threadIdx.x < p.groups) {
    const float* st = p.stats + ((size_t)(c * p.imgs + b) * p.groups + threadIdx.x) * 2;
    const float mu = st[0] * inv_n;
    s_mu[threadIdx.x] = mu; s_rs[threadIdx.x] = rsqrtf(fmaxf(st[1] * inv_n - mu * mu, 0.f) + kGnEps);
    s_S[threadIdx.x][0] = 0.f; s_S[threadIdx.x][1] = 0.f;
  }
  __syncthreads();
  const int vpc = p.Ch / 8;
  const int cv = threadIdx.x % vpc, pl = threadIdx.x / vpc, npl = blockDim.x / vpc;
  const int ch = cv * 8, g = ch / cpg;
  const size_t base = (size_t)(c * p.imgs + b) * p.HW * p.Ch;
  float ga[8], dg[8], db[8], S1 = 0.f, S2 = 0.f;
  load_f8(p.master + (size_t)c * p.P + p.gamma_off + ch, ga);
#pragma unroll
  for (int i = 0; i < 8; ++i) dg[i] = db[i] = 0.f;
  const float mu = s_mu[g], rs = s_rs[g];
  for (int pix = pl; pix < p.HW; pix += 2 * npl) {
    // two pixels per iteration: twice the bytes in flight per thread
    const int pix2 = pix + npl;
    const bool has2 = pix2 < p.HW;
    const size_t o = base + (size_t)pix * p.Ch + ch, o2 = base + (size_t)pix2 * p.Ch + ch;
    uint4 qd = *reinterpret_cast<const uint4*>(p.dy + o), qx = *reinterpret_cast<const uint4*>(p.raw + o), qm = make_uint4(0, 0, 0, 0);
    uint4 qd2 = make_uint4(0, 0, 0, 0), qx2 = qd2, qm2 = qd2;
    if (p.mask) qm = *reinterpret_cast<const uint4*>(p.mask + o);
    if (has2) {
      qd2 = *reinterpret_cast<const uint4*>(p.dy + o2); qx2 = *reinterpret_cast<const uint4*>(p.raw + o2);
      if (p.mask) qm2 = *reinterpret_cast<const uint4*>(p.mask + o2);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !has2) break;
      float d[8], x[8], m[8];
      unpack8(h ? qd2 : qd, d); unpack8(h ? qx2 : qx, x);
      if (p.mask) {
        unpack8(h ? qm2 : qm, m);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = m[i] > 0.f ? d[i] : 0.f;
      }
      s_d[(size_t)(h ? pix2 : pix) * vpc + cv] = pack8(d);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = (x[i] - mu) * rs;
        dg[i] += d[i] * xh; db[i] += d[i];
        const float gg = d[i] * ga[i];
        S1 += gg; S2 += gg * xh;
      }
    }
  }
  atomicAdd(&s_S[g][0], S1); atomicAdd(&s_S[g][1], S2);
#pragma unroll
  for (int i = 0; i < 8; ++i) { atomicAdd(&s_dg[ch + i], dg[i]); atomicAdd(&s_db[ch + i], db[i]); }
  __syncthreads();
  float* gr = p.gsmall + (size_t)c * p.gs_stride;
  for (int i = threadIdx.x; i < p.Ch; i += blockDim.x) {
    atomicAdd(gr + p.g_gamma_off + i, s_dg[i]);
    atomicAdd(gr + p.g_beta_off + i, s_db[i]);
  }
  // ---- pass 2 (same thread -> same (pixel, channel vector) pairs as above: each thread re-reads what it wrote)
  const float m1 = s_S[g][0] * inv_n, m2 = s_S[g][1] * inv_n;
  for (int pix = pl; pix < p.HW; pix += npl) {
    const size_t o = base + (size_t)pix * p.Ch + ch;
    float d[8], x[8], r[8];
    const uint4 qd = s_d[(size_t)pix * vpc + cv];
    unpack8(qd, d);
    unpack8(*reinterpret_cast<const uint4*>(p.raw + o), x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (x[i] - mu) * rs;
      r[i] = rs * (d[i] * ga[i] - m1 - xh * m2);
    }
    *reinterpret_cast<uint4*>(p.dx + o) = pack8(r);
    if (p.dres) *reinterpret_cast<uint4*>(p.dres + o) = qd;
  }
}

// pass 2:  dx = rstd * (g * gamma - m1 - xhat * m2),  optionally dres = g
__global__ void gn_bwd_apply_kernel(const GnBwdParams p) {
  const int c = blockIdx.z, b = blockIdx.y;
  if (p.active && !p.active[c]) return;
  __shared__ float s_mu[8], s_rs[8], s_m1[8], s_m2[8];
  const int cpg = p.Ch / p.groups;
  const float inv_n = 1.f / (float)(p.HW * cpg);
  if (threadIdx.x < p.groups) {
    const size_t k = ((size_t)(c * p.imgs + b) * p.groups + threadIdx.x) * 2;
    const float mu = p.stats[k] * inv_n;
    s_mu[threadIdx.x] = mu; s_rs[threadIdx.x] = rsqrtf(fmaxf(p.stats[k + 1] * inv_n - mu * mu, 0.f) + kGnEps);
    s_m1[threadIdx.x] = p.sb[k]; s_m2[threadIdx.x] = p.sb[k + 1];
  }
  __syncthreads();
  const int vpi = p.HW * p.Ch / 8, vpc = p.Ch / 8;
  const size_t base = (size_t)(c * p.imgs + b) * p.HW * p.Ch;
  const float* mrow = p.master + (size_t)c * p.P;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < vpi; v += gridDim.x * blockDim.x) {
    const int ch = (v % vpc) * 8, g = ch / cpg;
    const size_t o = base + (size_t)v * 8;
    float d[8], x[8], m[8], ga[8], r[8];
    unpack8(*reinterpret_cast<const uint4*>(p.dy + o), d);
    unpack8(*reinterpret_cast<const uint4*>(p.raw + o), x);
    if (p.mask) {
      unpack8(*reinterpret_cast<const uint4*>(p.mask + o), m);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = m[i] > 0.f ? d[i] : 0.f;
    }
    load_f8(mrow + p.gamma_off + ch, ga);
    const float mu = s_mu[g], rs = s_rs[g], m1 = s_m1[g], m2 = s_m2[g];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (x[i] - mu) * rs;
      r[i] = rs * (d[i] * ga[i] - m1 - xh * m2);
    }
    *reinterpret_cast<uint4*>(p.dx + o) = pack8(r);
    if (p.dres) *reinterpret_cast<uint4*>(p.dres + o) = pack8(d);
  }
}

// ---------------------------------------------------------------------------------------------------------
// ResNet head: global average pool + FC + softmax cross-entropy + FC backward + SGD on the FC parameters + gradient of
// the pooled feature map, one CTA per client.  h [C][imgs][HW][Ch] -> dh (same shape; the ReLU gate of the last block
// is applied by that block's GroupNorm backward).
// ---------------------------------------------------------------------------------------------------------
struct PoolHeadParams {
  int C, imgs, HW, Ch, classes; const uint8_t* active;
  const bf16* h; bf16* dh;
  const int32_t* labels; int label_stride, label_off;
  float* master; long long P; int w_off, b_off; const float* global_w; float lr, mu;
  float* loss_sum; int32_t* nonfinite;
  float* logits_out;      // optional [C][imgs][classes] (tests)
  int apply_update;
};

__global__ void __launch_bounds__(256) pool_fc_xent_kernel(const PoolHeadParams p) {
  const int c = blockIdx.x;
  if (p.active && !p.active[c]) return;
  extern __shared__ float sm[];
  float* s_w = sm;                               // [classes][Ch]
  float* s_dw = s_w + p.classes * p.Ch;          // [classes][Ch]
  float* s_pool = s_dw + p.classes * p.Ch;       // [Ch]
  float* s_log = s_pool + p.Ch;                  // [32]: logits / dlogits
  float* mrow = p.master + (size_t)c * p.P;
  for (int i = threadIdx.x; i < p.classes * p.Ch; i += blockDim.x) { s_w[i] = mrow[p.w_off + i]; s_dw[i] = 0.f; }
  __shared__ float s_dbias[16];
  __shared__ float s_loss;
  if (threadIdx.x < 16) s_dbias[threadIdx.x] = 0.f;
  if (threadIdx.x == 0) s_loss = 0.f;
  __syncthreads();
  const float inv_hw = 1.f / (float)p.HW, inv_b = 1.f / (float)p.imgs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int b = 0; b < p.imgs; ++b) {
    const bf16* hb = p.h + ((size_t)(c * p.imgs + b) * p.HW) * p.Ch;
    for (int ch = threadIdx.x; ch < p.Ch; ch += blockDim.x) {
      float a = 0.f;
      for (int px = 0; px < p.HW; ++px) a += __bfloat162float(hb[(size_t)px * p.Ch + ch]);
      s_pool[ch] = a * inv_hw;
    }
    __syncthreads();
    for (int k = warp; k < p.classes; k += nwarp) {
      float a = 0.f;
      for (int ch = lane; ch < p.Ch; ch += 32) a += s_pool[ch] * s_w[k * p.Ch + ch];
      for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) s_log[k] = a + mrow[p.b_off + k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float mx = -3e38f;
      for (int k = 0; k < p.classes; ++k) mx = fmaxf(mx, s_log[k]);
      float sum = 0.f;
      for (int k = 0; k < p.classes; ++k) sum += __expf(s_log[k] - mx);
      const int y = p.labels[(size_t)c * p.label_stride + p.label_off + b];
      s_loss += -(s_log[y] - mx - __logf(sum));
      for (int k = 0; k < p.classes; ++k) {
        if (p.logits_out) p.logits_out[((size_t)c * p.imgs + b) * p.classes + k] = s_log[k];
        const float d = (__expf(s_log[k] - mx) / sum - (k == y ? 1.f : 0.f)) * inv_b;
        s_log[k] = d;
        s_dbias[k] += d;
      }
    }
    __syncthreads();
    bf16* db = p.dh + ((size_t)(c * p.imgs + b) * p.HW) * p.Ch;
    for (int ch = threadIdx.x; ch < p.Ch; ch += blockDim.x) {
      float dp = 0.f;
      const float pl = s_pool[ch];
      for (int k = 0; k < p.classes; ++k) { dp += s_log[k] * s_w[k * p.Ch + ch]; s_dw[k * p.Ch + ch] += s_log[k] * pl; }
      const bf16 v = __float2bfloat16_rn(dp * inv_hw);
      for (int px = 0; px < p.HW; ++px) db[(size_t)px * p.Ch + ch] = v;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float mean = s_loss * inv_b;
    p.loss_sum[c] += mean;
    if (!(fabsf(mean) < 3.0e38f)) p.nonfinite[c] = 1;
  }
  if (p.apply_update) {
    for (int i = threadIdx.x; i < p.classes * p.Ch; i += blockDim.x) {
      float g = s_dw[i];
      const float w = s_w[i];
      if (p.global_w) g += p.mu * (w - p.global_w[p.w_off + i]);
      mrow[p.w_off + i] = w - p.lr * g;
    }
    if (threadIdx.x < p.classes) {
      float g = s_dbias[threadIdx.x];
      const float w = mrow[p.b_off + threadIdx.x];
      if (p.global_w) g += p.mu * (w - p.global_w[p.b_off + threadIdx.x]);
      mrow[p.b_off + threadIdx.x] = w - p.lr * g;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm over H = 256 features: one warp per token, 8 features per lane.
//   forward : (pre | embedding gather)  ->  xhat (bf16), y = xhat * gamma + beta (bf16), rstd (fp32)
//   backward: dx = rstd * (dy*gamma - mean(dy*gamma) - xhat * mean(dy*gamma*xhat));  dgamma, dbeta and the column sums
//             of dx (= bias gradient of the linear layer that produced `pre`) accumulate in registers over the CTA's
//             tokens and reach gsmall with one atomic per feature per CTA.
// ---------------------------------------------------------------------------------------------------------
constexpr int kH = 256;
struct LnParams {
  int C, T; const uint8_t* active;                  // T tokens per client
  const bf16* pre;                                  // [C][T][256] or null (embedding mode)
  const int32_t* ids; int ids_stride, ids_off, S;   // embedding mode: token ids [C][shard*S], this step's batch at ids_off
  int tok_off, pos_off;                             // embedding tables inside the master row
  const float* master; long long P; int gamma_off, beta_off;
  bf16* xhat; bf16* y; float* rstd;
  // backward
  const bf16* dy; bf16* dx;
  float* gsmall; long long gs_stride; int g_gamma_off, g_beta_off, g_bias_off;   // g_bias_off < 0: no bias gradient
};

__global__ void __launch_bounds__(256) ln_fwd_kernel(const LnParams p) {
  const int c = blockIdx.y;
  if (p.active && !p.active[c]) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* mrow = p.master + (size_t)c * p.P;
  float ga[8], be[8];
  load_f8(mrow + p.gamma_off + lane * 8, ga); load_f8(mrow + p.beta_off + lane * 8, be);
  for (int t = blockIdx.x * 8 + warp; t < p.T; t += gridDim.x * 8) {
    float x[8];
    const size_t o = ((size_t)c * p.T + t) * kH + lane * 8;
    if (p.pre) {
      unpack8(*reinterpret_cast<const uint4*>(p.pre + o), x);
    } else {
      const int id = p.ids[(size_t)c * p.ids_stride + p.ids_off + t];
      float e[8], q[8];
      load_f8(mrow + p.tok_off + (size_t)id * kH + lane * 8, e);
      load_f8(mrow + p.pos_off + (size_t)(t % p.S) * kH + lane * 8, q);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = e[i] + q[i];
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    for (int k = 16; k; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    const float mu = s * (1.f / kH);
    float q2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] -= mu; q2 += x[i] * x[i]; }
    for (int k = 16; k; k >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, k);
    const float rs = rsqrtf(q2 * (1.f / kH) + kLnEps);
    float yh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] *= rs; yh[i] = x[i] * ga[i] + be[i]; }
    *reinterpret_cast<uint4*>(p.xhat + o) = pack8(x);
    *reinterpret_cast<uint4*>(p.y + o) = pack8(yh);
    if (lane == 0) p.rstd[(size_t)c * p.T + t] = rs;
  }
}

__global__ void __launch_bounds__(256) ln_bwd_kernel(const LnParams p) {
  const int c = blockIdx.y;
  if (p.active && !p.active[c]) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ga[8], dg[8], db[8], dbias[8];
  load_f8(p.master + (size_t)c * p.P + p.gamma_off + lane * 8, ga);
#pragma unroll
  for (int i = 0; i < 8; ++i) dg[i] = db[i] = dbias[i] = 0.f;
  for (int t = blockIdx.x * 8 + warp; t < p.T; t += gridDim.x * 8) {
    const size_t o = ((size_t)c * p.T + t) * kH + lane * 8;
    float d[8], xh[8];
    unpack8(*reinterpret_cast<const uint4*>(p.dy + o), d);
    unpack8(*reinterpret_cast<const uint4*>(p.xhat + o), xh);
    const float rs = p.rstd[(size_t)c * p.T + t];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dg[i] += d[i] * xh[i]; db[i] += d[i];
      d[i] *= ga[i];
      s1 += d[i]; s2 += d[i] * xh[i];
    }
    for (int k = 16; k; k >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, k); s2 += __shfl_xor_sync(0xffffffffu, s2, k); }
    s1 *= (1.f / kH); s2 *= (1.f / kH);
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { r[i] = rs * (d[i] - s1 - xh[i] * s2); dbias[i] += r[i]; }
    *reinterpret_cast<uint4*>(p.dx + o) = pack8(r);
  }
  __shared__ float s_acc[3][kH];
  for (int i = threadIdx.x; i < 3 * kH; i += blockDim.x) (&s_acc[0][0])[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    atomicAdd(&s_acc[0][lane * 8 + i], dg[i]); atomicAdd(&s_acc[1][lane * 8 + i], db[i]); atomicAdd(&s_acc[2][lane * 8 + i], dbias[i]);
  }
  __syncthreads();
  float* gr = p.gsmall + (size_t)c * p.gs_stride;
  for (int i = threadIdx.x; i < kH; i += blockDim.x) {
    atomicAdd(gr + p.g_gamma_off + i, s_acc[0][i]);
    atomicAdd(gr + p.g_beta_off + i, s_acc[1][i]);
    if (p.g_bias_off >= 0) atomicAdd(gr + p.g_bias_off + i, s_acc[2][i]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Embedding backward: dpre [C][T][256] (gradient at the embedding sum).  Token rows: sparse SGD straight on the fp32
// master (atomicAdd of -lr * g; untouched rows keep their value, which is exactly SGD without momentum); position rows:
// dense gradient accumulated in gsmall.
// ---------------------------------------------------------------------------------------------------------
struct EmbBwdParams {
  int C, T, S; const uint8_t* active;
  const bf16* dpre; const int32_t* ids; int ids_stride, ids_off;
  float* master; long long P; int tok_off; float lr;
  float* gsmall; long long gs_stride; int g_pos_off;
};
__global__ void __launch_bounds__(256) emb_bwd_kernel(const EmbBwdParams p) {
  const int c = blockIdx.y;
  if (p.active && !p.active[c]) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* mrow = p.master + (size_t)c * p.P;
  float* gr = p.gsmall + (size_t)c * p.gs_stride;
  for (int t = blockIdx.x * 8 + warp; t < p.T; t += gridDim.x * 8) {
    float d[8];
    unpack8(*reinterpret_cast<const uint4*>(p.dpre + ((size_t)c * p.T + t) * kH + lane * 8), d);
    const int id = p.ids[(size_t)c * p.ids_stride + p.ids_off + t];
    float* tr = mrow + p.tok_off + (size_t)id * kH + lane * 8;
    float* pr = gr + p.g_pos_off + (size_t)(t % p.S) * kH + lane * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) { atomicAdd(tr + i, -p.lr * d[i]); atomicAdd(pr + i, d[i]); }
  }
}

// ---------------------------------------------------------------------------------------------------------
// BERT classifier head on token 0 of every sample: logits, softmax-xent, head SGD and the gradient of the last hidden
// state (zero everywhere except token 0).  One CTA per client.
// ---------------------------------------------------------------------------------------------------------
struct ClsHeadParams {
  int C, imgs, S, classes; const uint8_t* active;
  const bf16* y; bf16* dy;                    // [C][imgs*S][256]
  const int32_t* labels; int label_stride, label_off;
  float* master; long long P; int w_off, b_off; const float* global_w; float lr, mu;
  float* loss_sum; int32_t* nonfinite; float* logits_out; int apply_update;
};
__global__ void __launch_bounds__(256) cls_head_kernel(const ClsHeadParams p) {
  const int c = blockIdx.x;
  if (p.active && !p.active[c]) return;
  __shared__ float s_w[16 * kH], s_dw[16 * kH], s_dl[32][16], s_db[16];
  __shared__ float s_loss;
  float* mrow = p.master + (size_t)c * p.P;
  for (int i = threadIdx.x; i < p.classes * kH; i += blockDim.x) { s_w[i] = mrow[p.w_off + i]; s_dw[i] = 0.f; }
  if (threadIdx.x < 16) s_db[threadIdx.x] = 0.f;
  if (threadIdx.x == 0) s_loss = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_b = 1.f / (float)p.imgs;
  for (int b = warp; b < p.imgs; b += 8) {       // a warp per sample
    float x[8];
    unpack8(*reinterpret_cast<const uint4*>(p.y + ((size_t)c * p.imgs * p.S + (size_t)b * p.S) * kH + lane * 8), x);
    float lg[16];
    float mx = -3e38f;
    for (int k = 0; k < p.classes; ++k) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) a += x[i] * s_w[k * kH + lane * 8 + i];
      for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      lg[k] = a + mrow[p.b_off + k];
      mx = fmaxf(mx, lg[k]);
    }
    float sum = 0.f;
    for (int k = 0; k < p.classes; ++k) sum += __expf(lg[k] - mx);
    const int yl = p.labels[(size_t)c * p.label_stride + p.label_off + b];
    if (lane == 0) {
      float l = 0.f;
      for (int k = 0; k < p.classes; ++k) {
        if (k == yl) l = -(lg[k] - mx - __logf(sum));
        if (p.logits_out) p.logits_out[((size_t)c * p.imgs + b) * p.classes + k] = lg[k];
      }
      atomicAdd(&s_loss, l);
    }
    float dxv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dxv[i] = 0.f;
    for (int k = 0; k < p.classes; ++k) {
      const float d = (__expf(lg[k] - mx) / sum - (k == yl ? 1.f : 0.f)) * inv_b;
      if (lane == 0) { s_dl[b][k] = d; atomicAdd(&s_db[k], d); }
#pragma unroll
      for (int i = 0; i < 8; ++i) { dxv[i] += d * s_w[k * kH + lane * 8 + i]; atomicAdd(&s_dw[k * kH + lane * 8 + i], d * x[i]); }
    }
    *reinterpret_cast<uint4*>(p.dy + ((size_t)c * p.imgs * p.S + (size_t)b * p.S) * kH + lane * 8) = pack8(dxv);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float mean = s_loss * inv_b;
    p.loss_sum[c] += mean;
    if (!(fabsf(mean) < 3.0e38f)) p.nonfinite[c] = 1;
  }
  if (p.apply_update) {
    for (int i = threadIdx.x; i < p.classes * kH; i += blockDim.x) {
      float g = s_dw[i];
      const float w = s_w[i];
      if (p.global_w) g += p.mu * (w - p.global_w[p.w_off + i]);
      mrow[p.w_off + i] = w - p.lr * g;
    }
    if (threadIdx.x < p.classes) {
      float g = s_db[threadIdx.x];
      const float w = mrow[p.b_off + threadIdx.x];
      if (p.global_w) g += p.mu * (w - p.global_w[p.b_off + threadIdx.x]);
      mrow[p.b_off + threadIdx.x] = w - p.lr * g;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Column sums of a bf16 matrix per client (bias gradients): gsmall[c][g_off + n] += sum_t x[c][t][n].
// blockDim = (N/8, rows_per_block); grid = (token chunks, C).
// ---------------------------------------------------------------------------------------------------------
struct ColsumParams {
  int C, T, N, chunk; const uint8_t* active; const bf16* x; float* gsmall; long long gs_stride; int g_off;
};
__global__ void colsum_kernel(const ColsumParams p) {
  const int c = blockIdx.y;
  if (p.active && !p.active[c]) return;
  extern __shared__ float s_cs[];               // [blockDim.y][N]
  const int t0 = blockIdx.x * p.chunk, t1 = min(t0 + p.chunk, p.T);
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0.f;
  for (int t = t0 + threadIdx.y; t < t1; t += blockDim.y) {
    float x[8];
    unpack8(*reinterpret_cast<const uint4*>(p.x + ((size_t)c * p.T + t) * p.N + threadIdx.x * 8), x);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] += x[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s_cs[threadIdx.y * p.N + threadIdx.x * 8 + i] = a[i];
  __syncthreads();
  if (threadIdx.y == 0) {
    float* gr = p.gsmall + (size_t)c * p.gs_stride + p.g_off + threadIdx.x * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = 0.f;
      for (int y = 0; y < (int)blockDim.y; ++y) s += s_cs[y * p.N + threadIdx.x * 8 + i];
      atomicAdd(gr + i, s);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// SGD / FedProx step on a slice of the packed rows:  w[c][off+i] -= lr * (g[c][i] + mu * (w - w_global)),  g optional
// (g == nullptr: proximal term only — the dense part of FedProx for sparsely-updated embedding tables).
// ---------------------------------------------------------------------------------------------------------
__global__ void slice_sgd_kernel(float* __restrict__ master, long long P, int off, int n, const float* __restrict__ g,
                                 long long gs_stride, const float* __restrict__ global_w, const uint8_t* __restrict__ active,
                                 float lr, float mu) {
  const int c = blockIdx.y;
  if (active && !active[c]) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* w = master + (size_t)c * P + off + i;
  float gg = g ? g[(size_t)c * gs_stride + i] : 0.f;
  const float wv = *w;
  if (global_w) gg += mu * (wv - global_w[off + i]);
  *w = wv - lr * gg;
}

}  // namespace nnk
