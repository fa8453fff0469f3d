// fl_kernels.cuh — the non-GEMM device code of a federated round:
//   ingest_u8          uint8 NHWC client samples  -> normalised bf16, horizontally unrolled conv1 operand
//   bcast_scatter      global model (optionally pulled from the owning peer GPU over NVLink) -> packed
//                      per-client replicas: fp32 master rows + bf16 tensor-core operand copies
//   head_fc_xent       fused classifier head: FC forward, softmax-xent, FC backward, SGD on FC + conv2 bias,
//                      un-pool / ReLU-backward into the conv2 pre-activation gradient
//   fedavg_partial     masked, sample-count-weighted partial sums of client deltas  (on-GPU level of the reduce)
//   fedavg_finish      cross-GPU pull of the partials (P2P loads), normalise, server SGD/Adam, write new shard
//   deviceflow_mask    counter-based (Philox4x32-10) per-round participation / drop sampling
//   peer_barrier       device-side barrier over flags in peer-mapped memory
//
// Reference parity: model broadcast = per-actor model-file download (ols_core/taskMgr/utils/
// utils_run_task.py:327-397); update reduce = DeviceFlow sorter/dispatcher message path
// (ols_core/deviceflow/non_grpc/sorter.py:37-91, dispatcher.py:84-242); sampling/drops =
// ols_core/deviceflow/non_grpc/strategy.py:275-311 and dispatcher.py:103-109.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc05.cuh"

namespace flk {

// ---------------------------------------------------------------------------------------------
// CNN-2 packed parameter layout (floats per client).  Shared with python (ops/layout.py).
// ---------------------------------------------------------------------------------------------
struct Cnn2Layout {
  static constexpr int W1 = 0;                 // [32][5][16]   cout, kh, (kw*3+c ; 15 used, 1 pad)
  static constexpr int B1 = 2560;              // [32]
  static constexpr int W2 = 2592;              // [64][25][32]  cout, kh*5+kw, cin
  static constexpr int B2 = 53792;             // [64]
  static constexpr int WFC = 53856;            // [10][4096]    class, (ph*8+pw)*64+ch
  static constexpr int BFC = 94816;            // [10] (+6 pad)
  static constexpr int P = 94832;
  static constexpr int W1_N = 2560, W2_N = 51200, WFC_N = 40960;
};

constexpr int kMaxPeers = 8;
struct PeerPtrs {
  void* p[kMaxPeers];
};

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }

// uint8 [N][32][32][3]  ->  bf16 [N][32][32][16] with channel = kw*3+c holding x[h][w+kw-2][c] (zero outside).
// The 5 horizontal taps of conv1 are unrolled into the channel dim once at upload time, so conv1 runs as a
// 5-tap (vertical) implicit GEMM with K=16 per tap instead of wasting 13/16 of a K=16 MMA on 3 channels.
// preswizzle 2: additionally split every image row into its two column-parity planes (conv1_window_kernel).
// preswizzle 1: store the pixel's two 16-byte halves swapped when (w & 4).  That is exactly the 32-byte swizzle the
// tensor core applies to a tile of 32-byte pixel rows (address bit 4 ^= bit 7), done once in HBM, so TMA can fetch a
// whole 16-pixel image row as ONE 512-byte unswizzled box row instead of sixteen 32-byte rows (TMA cost is per row).
__global__ void ingest_u8_kernel(const uint8_t* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t npix,
                                 int preswizzle) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const int w = i & 31;
  const uint8_t* row = src + (i - w) * 3;
  __align__(16) __nv_bfloat16 v[16];
#pragma unroll
  for (int kw = 0; kw < 5; ++kw) {
    const int ww = w + kw - 2;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float f = 0.f;
      if (ww >= 0 && ww < 32) f = (static_cast<float>(row[ww * 3 + c]) * (1.f / 255.f) - 0.5f) * 2.f;
      v[kw * 3 + c] = __float2bfloat16_rn(f);
    }
  }
  v[15] = __float2bfloat16_rn(0.f);
  if (preswizzle == 2) {
    // column-parity planes: [row][w & 1][w >> 1][16]; inside a plane the 32-byte swizzle follows the plane index j
    const int j = w >> 1;
    uint4* o = reinterpret_cast<uint4*>(dst + ((i - w) + (w & 1) * 16 + j) * 16);
    const int sw = (j & 4) ? 1 : 0;
    o[sw] = reinterpret_cast<uint4*>(v)[0];
    o[sw ^ 1] = reinterpret_cast<uint4*>(v)[1];
    return;
  }
  uint4* o = reinterpret_cast<uint4*>(dst + i * 16);
  const int sw = (preswizzle && (w & 4)) ? 1 : 0;
  o[sw] = reinterpret_cast<uint4*>(v)[0];
  o[sw ^ 1] = reinterpret_cast<uint4*>(v)[1];
}

// ---------------------------------------------------------------------------------------------
// bcast_scatter: one pass that (a) pulls each parameter shard from its owner GPU (peer-mapped pointer; the
// owner of element i is i / shard_len), (b) refreshes the local copy of the global model, and (c) fans the
// values out into every local client replica row (fp32 master) plus the bf16 operand copies in the layouts
// the tensor-core kernels consume.  With one GPU the "pull" is a local read.
// ---------------------------------------------------------------------------------------------
struct BcastParams {
  PeerPtrs global_peers;        // float* [P] on each rank (symmetric)
  int nranks, rank;
  int shard_len;                // elements per owner shard (P padded / nranks)
  float* global_local;          // [P] local copy (FedProx anchor, checkpoint source)
  float* master;                // [C][P]
  int num_clients;
  const uint8_t* participate;   // [C] or nullptr — only refresh clients sampled for this round
  __nv_bfloat16* w1_fwd;        // [C][2560]
  __nv_bfloat16* w2_fwd;        // [C][51200]
  __nv_bfloat16* w2_dgrad;      // [C][25*32][64]
  int P;
};

__global__ void bcast_scatter_cnn2_kernel(const BcastParams p) {
  using L = Cnn2Layout;
  // each thread owns 4 consecutive parameters; blockIdx.y strides over clients
  const int i4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= p.P) return;
  const int owner = min(i4 / p.shard_len, p.nranks - 1);
  const float4 v = *reinterpret_cast<const float4*>(static_cast<const float*>(p.global_peers.p[owner]) + i4);
  if (blockIdx.y == 0) *reinterpret_cast<float4*>(p.global_local + i4) = v;
  const float vv[4] = {v.x, v.y, v.z, v.w};
  for (int c = blockIdx.y; c < p.num_clients; c += gridDim.y) {
    if (p.participate && !p.participate[c]) continue;
    *reinterpret_cast<float4*>(p.master + (size_t)c * p.P + i4) = v;
    if (i4 < L::B1) {
      uint2 o;
      __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
      o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(p.w1_fwd + (size_t)c * L::W1_N + i4) = o;
    } else if (i4 >= L::W2 && i4 < L::B2) {
      const int j = i4 - L::W2;
      uint2 o;
      __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
      o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(p.w2_fwd + (size_t)c * L::W2_N + j) = o;
      if (p.w2_dgrad) {   // optional transposed copy [tap][cin][cout] (the data-gradient kernel reads the forward copy MN-major)
        const int cout = j / 800, rem = j % 800;   // rem = tap*32 + cin (4 consecutive cin)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          p.w2_dgrad[(size_t)c * L::W2_N + (size_t)(rem + e) * 64 + cout] = __float2bfloat16_rn(vv[e]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused classifier head.  One CTA per client, 256 threads.  BATCH=32 samples, K=4096 features, 10 classes.
// ---------------------------------------------------------------------------------------------
struct HeadParams {
  int num_clients;
  const uint8_t* active;
  const __nv_bfloat16* p2;      // [C*32][8][8][64]  pooled conv2 activations
  const uint8_t* code2;         // same shape
  const int32_t* labels;        // [C][label_stride]
  int label_stride, label_off;  // this step's batch = labels[c][label_off .. +32)
  float* master;                // [C][P]
  const float* global_w;        // [P] FedProx anchor or nullptr
  int P;
  float lr, mu;
  __nv_bfloat16* dy2;           // [C*32][16][16][64]  grad wrt conv2 pre-pool output
  float* loss_sum;              // [C]  += mean loss of this step
  int32_t* nonfinite;           // [C]  set to 1 when the loss is not finite (poisoned client)
  float* logits_out;            // optional [C][32][10] (tests)
  int apply_update;             // 0 => gradient-check mode (no parameter update)
  float* dwfc_out;              // optional [C][10*4096+16] raw FC grads (tests)
  float* db2_out;               // optional [C][64]
};

constexpr int kHeadThreads = 256;
// backward-phase staging ring: a slice = 8 samples x 512 features of the pooled activation (bf16) and its argmax code,
// fetched by 1-D bulk copies four slices ahead of their use
constexpr int kHeadSlots = 4, kHeadSliceX = 8 * 512 * 2, kHeadSliceC = 8 * 512, kHeadSlice = kHeadSliceX + kHeadSliceC;
constexpr int kHeadSmem = (10 * 4096 + 32 * 10 + 32 * 10 + (kHeadThreads / 32) * 64 + 16) * 4 + kHeadSlots * kHeadSlice + 64;

__global__ void __launch_bounds__(kHeadThreads, 1) head_fc_xent_kernel(const HeadParams p) {
  using L = Cnn2Layout;
  extern __shared__ float hs[];
  float* Ws = hs;                    // [10][4096]
  float* logit = Ws + 10 * 4096;     // [32][10]
  float* dlog = logit + 320;         // [32][10]
  float* db2 = dlog + 320;           // [8 warps][64]: per-warp conv2-bias gradient partials (summed in warp order)
  float* misc = db2 + (kHeadThreads / 32) * 64;   // [16]
  uint8_t* ring = reinterpret_cast<uint8_t*>(misc + 16);
  uint64_t* rfull = reinterpret_cast<uint64_t*>(ring + kHeadSlots * kHeadSlice);
  uint64_t* rempty = rfull + kHeadSlots;      // slot consumed: one arrival per warp
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kHeadSlots; ++i) { tc05::mbar_init(&rfull[i], 1); tc05::mbar_init(&rempty[i], kHeadThreads / 32); }
    tc05::fence_mbar_init();
  }
  __syncthreads();
  uint32_t nslice = 0;               // slices consumed so far by this CTA (slot = nslice % 4, parity = nslice / 4 & 1)

  for (int c = blockIdx.x; c < p.num_clients; c += gridDim.x) {
    if (p.active && !p.active[c]) continue;
    float* wm = p.master + (size_t)c * p.P;
    // slice n of a client = (feature block it = n / 4, sample batch b = n % 4); issued by one thread
    // issued by the 16 low lanes of warp 0 together: lane l copies sample l/2's activation (even l) or code (odd l) row
    auto issue_slice = [&](int n, uint32_t seq) {
      const int it = n >> 2, s0 = (n & 3) * 8, slot = seq % kHeadSlots;
      uint8_t* dst = ring + slot * kHeadSlice;
      if (lane == 0) tc05::mbar_arrive_expect_tx(&rfull[slot], kHeadSlice);
      __syncwarp();
      if (lane < 16) {
        const int u = lane >> 1;
        if (lane & 1)
          tc05::bulk_copy_g2s(dst + kHeadSliceX + u * 512, p.code2 + ((size_t)c * 32 + s0 + u) * 4096 + it * 512, 512,
                              &rfull[slot]);
        else
          tc05::bulk_copy_g2s(dst + u * 1024, p.p2 + ((size_t)c * 32 + s0 + u) * 4096 + it * 512, 1024, &rfull[slot]);
      }
    };
    // the first four slices of the backward phase travel while the weights load and the logits are computed
    if (warp == 0) {
      tc05::fence_proxy_async_smem();
      for (int n = 0; n < kHeadSlots; ++n) issue_slice(n, nslice + n);
    }
    // 40 float4 per thread, 8 loads in flight per thread before the first shared-memory store
#pragma unroll 1
    for (int i0 = tid * 4; i0 < L::WFC_N; i0 += kHeadThreads * 4 * 8) {
      float4 t8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t8[u] = __ldg(reinterpret_cast<const float4*>(wm + L::WFC + i0 + u * kHeadThreads * 4));
#pragma unroll
      for (int u = 0; u < 8; ++u) *reinterpret_cast<float4*>(Ws + i0 + u * kHeadThreads * 4) = t8[u];
    }
    __syncthreads();

    // ---- logits: a warp owns 4 samples; per k-chunk the 10 weight rows are read from smem once and reused for
    // the 4 samples, and the activations of the next chunk are already in flight (register double buffer)
    const __nv_bfloat16* x = p.p2 + (size_t)c * 32 * 4096;
    {
      float acc[4][10];
#pragma unroll
      for (int si = 0; si < 4; ++si)
#pragma unroll
        for (int j = 0; j < 10; ++j) acc[si][j] = 0.f;
      const __nv_bfloat16* xw = x + (size_t)(warp * 4) * 4096 + lane * 8;
      uint4 nxt[4];
#pragma unroll
      for (int si = 0; si < 4; ++si) nxt[si] = __ldg(reinterpret_cast<const uint4*>(xw + (size_t)si * 4096));
#pragma unroll 1
      for (int k = lane * 8; k < 4096; k += 256) {
        uint4 cur[4];
#pragma unroll
        for (int si = 0; si < 4; ++si) cur[si] = nxt[si];
        if (k + 256 < 4096) {
#pragma unroll
          for (int si = 0; si < 4; ++si) nxt[si] = __ldg(reinterpret_cast<const uint4*>(xw + (size_t)si * 4096 + (k - lane * 8) + 256));
        }
        float xf[4][8];
#pragma unroll
        for (int si = 0; si < 4; ++si) {
          const uint32_t xw4[4] = {cur[si].x, cur[si].y, cur[si].z, cur[si].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) { xf[si][2 * e] = __uint_as_float(xw4[e] << 16); xf[si][2 * e + 1] = __uint_as_float(xw4[e] & 0xFFFF0000u); }
        }
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          const float4 w0 = *reinterpret_cast<const float4*>(Ws + j * 4096 + k);
          const float4 w1 = *reinterpret_cast<const float4*>(Ws + j * 4096 + k + 4);
#pragma unroll
          for (int si = 0; si < 4; ++si)
            acc[si][j] += xf[si][0] * w0.x + xf[si][1] * w0.y + xf[si][2] * w0.z + xf[si][3] * w0.w +
                          xf[si][4] * w1.x + xf[si][5] * w1.y + xf[si][6] * w1.z + xf[si][7] * w1.w;
        }
      }
#pragma unroll
      for (int si = 0; si < 4; ++si)
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          float v = acc[si][j];
          v += __shfl_xor_sync(0xffffffffu, v, 16); v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 4);  v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          if (lane == 0) logit[(warp * 4 + si) * 10 + j] = v + wm[L::BFC + j];
        }
    }
    __syncthreads();

    // ---- softmax cross-entropy, dlogits = (p - onehot)/B
    if (tid < 32) {
      const int s = tid;
      const int y = p.labels[(size_t)c * p.label_stride + p.label_off + s];
      float mx = -1e30f;
#pragma unroll
      for (int j = 0; j < 10; ++j) mx = fmaxf(mx, logit[s * 10 + j]);
      float e[10], sum = 0.f;
#pragma unroll
      for (int j = 0; j < 10; ++j) { e[j] = __expf(logit[s * 10 + j] - mx); sum += e[j]; }
      const float inv = 1.f / sum;
      float loss = 0.f;
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        const float pj = e[j] * inv;
        if (j == y) loss = -(logit[s * 10 + j] - mx - __logf(sum));
        dlog[s * 10 + j] = (pj - (j == y ? 1.f : 0.f)) * (1.f / 32.f);
      }
      if (p.logits_out) {
#pragma unroll
        for (int j = 0; j < 10; ++j) p.logits_out[((size_t)c * 32 + s) * 10 + j] = logit[s * 10 + j];
      }
      float ls = loss;
      ls += __shfl_xor_sync(0xffffffffu, ls, 16); ls += __shfl_xor_sync(0xffffffffu, ls, 8);
      ls += __shfl_xor_sync(0xffffffffu, ls, 4);  ls += __shfl_xor_sync(0xffffffffu, ls, 2);
      ls += __shfl_xor_sync(0xffffffffu, ls, 1);
      if (tid == 0) {
        const float m = ls * (1.f / 32.f);
        p.loss_sum[c] += m;
        if (!(fabsf(m) < 3.0e38f)) p.nonfinite[c] = 1;   // catches NaN and Inf
      }
    }
    __syncthreads();

    // ---- backward: dW_fc, dP2 -> un-pool -> dY2_pre, db2, db_fc ; SGD on W_fc
    const uint8_t* code = p.code2 + (size_t)c * 32 * 4096;
    __nv_bfloat16* dy = p.dy2 + (size_t)c * 32 * 16 * 16 * 64;
    const float* wg = p.global_w;
    // a thread's two features map to the same channel pair in every iteration (512 features = 8 pixels x 64 channels per
    // iteration), so the conv2-bias gradient accumulates in registers and is combined in a fixed order: no atomics,
    // bit-reproducible
    float dbs0 = 0.f, dbs1 = 0.f;
    for (int it = 0; it < 8; ++it) {
      const int k = (it * kHeadThreads + tid) * 2;       // two consecutive features (same pooled pixel)
      float w0[10], w1[10], g0[10], g1[10];
#pragma unroll
      for (int j = 0; j < 10; ++j) { w0[j] = Ws[j * 4096 + k]; w1[j] = Ws[j * 4096 + k + 1]; g0[j] = 0.f; g1[j] = 0.f; }
      const int pix = k >> 6, ch = k & 63, ph = pix >> 3, pw = pix & 7;
      float dbl0 = 0.f, dbl1 = 0.f;
      // 8 samples per batch, staged in shared memory by bulk copies issued four slices (~2 us of work) ahead
#pragma unroll 1
      for (int s0 = 0; s0 < 32; s0 += 8) {
        const uint32_t seq = nslice++;
        const int slot = seq % kHeadSlots;
        tc05::mbar_wait(&rfull[slot], (seq / kHeadSlots) & 1);
        const uint8_t* sl = ring + slot * kHeadSlice;
        uint32_t xr8[8];
        uint16_t cr8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          xr8[u] = *reinterpret_cast<const uint32_t*>(sl + u * 1024 + tid * 4);
          cr8[u] = *reinterpret_cast<const uint16_t*>(sl + kHeadSliceX + u * 512 + tid * 2);
        }
        // refill the slot with the slice four ahead once every warp has read it.  Only warp 0 (the issuer) waits for
        // that; the other warps signal "consumed" on the slot's mbarrier and run on (a full __syncthreads here cost
        // 16 % of all warp samples).  The 8 values read above are in registers, so the slot may be overwritten.
        __syncwarp();
        if (lane == 0) tc05::mbar_arrive(&rempty[slot]);
        if (warp == 0 && it * 4 + (s0 >> 3) + kHeadSlots < 32) {
          tc05::mbar_wait(&rempty[slot], (seq / kHeadSlots) & 1);
          tc05::fence_proxy_async_smem();
          issue_slice(it * 4 + (s0 >> 3) + kHeadSlots, seq + kHeadSlots);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int s = s0 + u;
          const uint32_t xr = xr8[u];
          const float x0 = __uint_as_float(xr << 16), x1 = __uint_as_float(xr & 0xFFFF0000u);
          const uint16_t cr = cr8[u];
          float d0 = 0.f, d1 = 0.f;
#pragma unroll
          for (int j = 0; j < 10; ++j) {
            const float dl = dlog[s * 10 + j];
            d0 += dl * w0[j]; d1 += dl * w1[j];
            g0[j] += dl * x0; g1[j] += dl * x1;
          }
          d0 = x0 > 0.f ? d0 : 0.f; d1 = x1 > 0.f ? d1 : 0.f;
          dbl0 += d0; dbl1 += d1;
          const int c0 = cr & 0xFF, c1 = cr >> 8;
#pragma unroll
          for (int pos = 0; pos < 4; ++pos) {
            const size_t o = (((size_t)s * 16 + (2 * ph + (pos >> 1))) * 16 + (2 * pw + (pos & 1))) * 64 + ch;
            __nv_bfloat162 ov = __floats2bfloat162_rn(c0 == pos ? d0 : 0.f, c1 == pos ? d1 : 0.f);
            *reinterpret_cast<__nv_bfloat162*>(dy + o) = ov;
          }
        }
      }
      dbs0 += dbl0; dbs1 += dbl1;
      if (p.dwfc_out) {
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          p.dwfc_out[(size_t)c * (L::WFC_N + 16) + j * 4096 + k] = g0[j];
          p.dwfc_out[(size_t)c * (L::WFC_N + 16) + j * 4096 + k + 1] = g1[j];
        }
      }
      if (p.apply_update) {
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          float a = w0[j], b = w1[j];
          float ga = g0[j], gb = g1[j];
          if (wg) { ga += p.mu * (a - wg[L::WFC + j * 4096 + k]); gb += p.mu * (b - wg[L::WFC + j * 4096 + k + 1]); }
          a -= p.lr * ga; b -= p.lr * gb;
          *reinterpret_cast<float2*>(wm + L::WFC + j * 4096 + k) = make_float2(a, b);
        }
      }
    }
    db2[warp * 64 + ((tid * 2) & 63)] = dbs0;
    db2[warp * 64 + ((tid * 2) & 63) + 1] = dbs1;
    __syncthreads();
    if (tid < 10) {
      float g = 0.f;
      for (int s = 0; s < 32; ++s) g += dlog[s * 10 + tid];
      if (p.dwfc_out) p.dwfc_out[(size_t)c * (L::WFC_N + 16) + L::WFC_N + tid] = g;
      if (p.apply_update) {
        float b = wm[L::BFC + tid];
        if (wg) g += p.mu * (b - wg[L::BFC + tid]);
        wm[L::BFC + tid] = b - p.lr * g;
      }
    }
    if (tid >= 64 && tid < 128) {
      const int ch = tid - 64;
      float g = 0.f;
#pragma unroll
      for (int w = 0; w < kHeadThreads / 32; ++w) g += db2[w * 64 + ch];
      if (p.db2_out) p.db2_out[(size_t)c * 64 + ch] = g;
      if (p.apply_update) {
        float b = wm[L::B2 + ch];
        if (wg) g += p.mu * (b - wg[L::B2 + ch]);
        wm[L::B2 + ch] = b - p.lr * g;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// FedAvg reduce, level 1 (on-GPU): partial[g][i] = sum_{c in group g} a_c * (master[c][i] - global[i]),
// a_c = participate_c * arrived_c * finite_c * n_c.  Streaming, HBM-bound; float4 per thread.
// Also produces wsum[g] = sum a_c (thread 0 of block x==0) and per-class success / failed counters.
// ---------------------------------------------------------------------------------------------
struct PartialParams {
  const float* master;      // [C][P]
  const float* global_w;    // [P]
  const float* weight;      // [C] effective aggregation weight a_c (0 => masked)
  const int32_t* nonfinite; // [C] or nullptr (1 => client poisoned; treated as weight 0)
  int num_clients, P, groups;
  float* partial;           // [groups][P]
  float* wsum;              // [groups]
};

__global__ void fedavg_partial_kernel(const PartialParams p) {
  const int i4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int g = blockIdx.y;
  const int per = (p.num_clients + p.groups - 1) / p.groups;
  const int c0 = g * per, c1 = min(c0 + per, p.num_clients);
  if (i4 < p.P) {
    const float4 gw = *reinterpret_cast<const float4*>(p.global_w + i4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float asum = 0.f;
#pragma unroll 4
    for (int c = c0; c < c1; ++c) {
      float a = __ldg(&p.weight[c]);
      if (p.nonfinite && __ldg(&p.nonfinite[c])) a = 0.f;
      if (a == 0.f) continue;
      const float4 w = __ldcs(reinterpret_cast<const float4*>(p.master + (size_t)c * p.P + i4));
      acc.x += a * (w.x - gw.x); acc.y += a * (w.y - gw.y); acc.z += a * (w.z - gw.z); acc.w += a * (w.w - gw.w);
      asum += a;
    }
    *reinterpret_cast<float4*>(p.partial + (size_t)g * p.P + i4) = acc;
    if (i4 == 0) p.wsum[g] = asum;
  }
}

// ---------------------------------------------------------------------------------------------
// Device-side barrier across ranks: flags live in peer-mapped memory (one uint32 per (dst, src) pair).
// Every rank writes `epoch` into slot [rank] of every peer's flag array, then waits until all of its own
// slots reached `epoch`.  The spin is bounded: when a peer does not show up within `spin_limit` polls the waiting
// thread records 1 + (peer index) in `*lost` and gives up.  The kernel then returns without touching the model, the
// CUDA context stays healthy and the host raises PeerLostError (parallel/comm.py) — a dead rank costs the survivors
// one round, not their process.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_barrier(const PeerPtrs& flags, int nranks, int rank, uint32_t epoch,
                                             uint32_t spin_limit, uint32_t* lost) {
  // caller: a single CTA-wide call; thread r (< nranks) signals peer r and waits on local slot r
  __syncthreads();
  if ((int)threadIdx.x < nranks) {
    __threadfence_system();
    st_release_sys(static_cast<uint32_t*>(flags.p[threadIdx.x]) + rank, epoch);
    const uint32_t* mine = static_cast<const uint32_t*>(flags.p[rank]) + threadIdx.x;
    uint32_t spins = 0;
    while (ld_acquire_sys(mine) < epoch) {
      if (++spins > spin_limit) { atomicMax(lost, 1u + threadIdx.x); break; }
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// FedAvg reduce, level 2 + server optimiser, fused with the cross-GPU transfer: the owner of a parameter
// shard pulls every peer's partial sums for that shard straight out of peer HBM (P2P loads over NVLink,
// no NCCL, no staging copy), normalises by the global weight sum, applies server SGD(momentum)/Adam and
// writes the new global shard in place.  Grid barrier at entry (partials ready) and exit (shard published).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxClasses = 24;
constexpr int kRoundRec = 8 + 2 * kMaxClasses;      // floats per rank record; round_out[kRoundRec] carries the peer-lost code
struct FinishParams {
  PeerPtrs partial_peers;   // float* [groups][P] per rank
  PeerPtrs wsum_peers;      // float* [groups] per rank
  PeerPtrs flags;           // uint32* [kMaxPeers] per rank
  int nranks, rank, groups, P, shard_len;
  uint32_t epoch;
  float* global_w;          // [P] on this rank (symmetric; peers pull from it in bcast_scatter)
  float* opt_m;             // [shard_len] momentum / Adam m   (owner shard only)
  float* opt_v;             // [shard_len] Adam v
  int opt;                  // 0 = plain averaging (server lr), 1 = SGD momentum, 2 = Adam
  float server_lr, beta1, beta2, eps;
  int step;                 // Adam step (1-based)
  float* stats;             // [4]: total weight, |delta|^2 of owned shard, participants(unused), spare
  uint32_t* arrive_ctr;     // [4]: [0],[1] scratch of the intra-grid barriers (zeroed per launch); [2] sticky peer-lost code
  uint32_t spin_limit;      // polls of a peer flag before the barrier gives up
  PeerPtrs round_peers;     // float* [round_len] per rank: this round's local accounting (round_stats_kernel) or null
  float* round_out;         // [round_len] sum over ranks of round_peers (every rank gets the job-wide numbers)
  int round_len;            // 8 (+ 2 per device class)
  // NVSwitch multicast path (both null: peer-pointer path).  partial_mc: multicast address of the [groups][P] partial
  // sums, read with multimem.ld_reduce (the switch adds the ranks' copies); global_mc: multicast address of the
  // replicated global model, written with multimem.st (every rank's copy receives the owner's new shard).
  const float* partial_mc;
  float* global_mc;
};

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, const float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void grid_gate(uint32_t* ctr, uint32_t nblocks) {
  // lightweight one-shot "all CTAs arrived" (all CTAs are co-resident: grid <= #SMs)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    uint32_t spins = 0;
    while (atomicAdd(ctr, 0u) < nblocks) { if (++spins > (1u << 28)) __trap(); }
    __threadfence();
  }
  __syncthreads();
}

__global__ void fedavg_finish_kernel(const FinishParams p) {
  // entry barrier: every rank's partials are complete & visible
  if (p.nranks > 1) {
    if (blockIdx.x == 0) peer_barrier(p.flags, p.nranks, p.rank, p.epoch * 2 + 1, p.spin_limit, p.arrive_ctr + 2);
    grid_gate(p.arrive_ctr + 0, gridDim.x);
    const uint32_t lost = *reinterpret_cast<volatile uint32_t*>(p.arrive_ctr + 2);
    if (lost != 0u) {       // a peer is gone (now or in an earlier round): leave the model untouched
      if (blockIdx.x == 0 && threadIdx.x == 0 && p.round_out) p.round_out[kRoundRec] = static_cast<float>(lost);
      return;
    }
  }
  float wtot = 0.f;
  for (int r = 0; r < p.nranks; ++r) {
    const float* ws = static_cast<const float*>(p.wsum_peers.p[r]);
    for (int g = 0; g < p.groups; ++g) wtot += ws[g];
  }
  const float inv = wtot > 0.f ? 1.f / wtot : 0.f;
  const int s0 = p.rank * p.shard_len;
  const int s1 = min(s0 + p.shard_len, p.P);
  float dn = 0.f;
  for (int i4 = s0 + (blockIdx.x * blockDim.x + threadIdx.x) * 4; i4 < s1; i4 += gridDim.x * blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.partial_mc) {
      // in-switch reduction: one load per group returns the sum over all ranks
      for (int g = 0; g < p.groups; ++g) {
        const float4 v = multimem_ld_reduce_add(p.partial_mc + (size_t)g * p.P + i4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    } else {
      for (int r = 0; r < p.nranks; ++r) {
        const float* pp = static_cast<const float*>(p.partial_peers.p[r]);
        for (int g = 0; g < p.groups; ++g) {
          const float4 v = *reinterpret_cast<const float4*>(pp + (size_t)g * p.P + i4);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
    }
    float d[4] = {acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv};   // averaged client delta (pseudo-gradient = -d)
    float4 w = *reinterpret_cast<const float4*>(p.global_w + i4);
    float wv[4] = {w.x, w.y, w.z, w.w};
    const int li = i4 - s0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float upd;
      if (p.opt == 0) {
        upd = p.server_lr * d[e];
      } else if (p.opt == 1) {
        const float m = p.beta1 * p.opt_m[li + e] + d[e];
        p.opt_m[li + e] = m;
        upd = p.server_lr * m;
      } else {
        const float m = p.beta1 * p.opt_m[li + e] + (1.f - p.beta1) * d[e];
        const float v = p.beta2 * p.opt_v[li + e] + (1.f - p.beta2) * d[e] * d[e];
        p.opt_m[li + e] = m; p.opt_v[li + e] = v;
        const float mh = m / (1.f - __powf(p.beta1, (float)p.step));
        const float vh = v / (1.f - __powf(p.beta2, (float)p.step));
        upd = p.server_lr * mh / (sqrtf(vh) + p.eps);
      }
      if (wtot <= 0.f) upd = 0.f;     // nobody arrived this round: keep the model
      wv[e] += upd;
      dn += upd * upd;
    }
    if (p.global_mc) multimem_st(p.global_mc + i4, make_float4(wv[0], wv[1], wv[2], wv[3]));   // all replicas at once
    else *reinterpret_cast<float4*>(p.global_w + i4) = make_float4(wv[0], wv[1], wv[2], wv[3]);
  }
  if (p.global_mc) __threadfence_system();
  // stats
  dn += __shfl_xor_sync(0xffffffffu, dn, 16); dn += __shfl_xor_sync(0xffffffffu, dn, 8);
  dn += __shfl_xor_sync(0xffffffffu, dn, 4);  dn += __shfl_xor_sync(0xffffffffu, dn, 2);
  dn += __shfl_xor_sync(0xffffffffu, dn, 1);
  if ((threadIdx.x & 31) == 0 && dn != 0.f) atomicAdd(&p.stats[1], dn);
  if (blockIdx.x == 0 && threadIdx.x == 0) p.stats[0] = wtot;
  // job-wide round accounting: sum the 8-float records of all ranks (they were complete before the entry barrier)
  if (blockIdx.x == 0 && (int)threadIdx.x < p.round_len && p.round_out) {
    float a = 0.f;
    for (int r = 0; r < p.nranks; ++r) a += static_cast<const float*>(p.round_peers.p[r])[threadIdx.x];
    p.round_out[threadIdx.x] = a;
  }
  // exit barrier: new shard visible to every peer before anyone starts the next broadcast pull
  if (p.nranks > 1) {
    grid_gate(p.arrive_ctr + 1, gridDim.x);
    if (blockIdx.x == 0) {
      peer_barrier(p.flags, p.nranks, p.rank, p.epoch * 2 + 2, p.spin_limit, p.arrive_ctr + 2);
      if (threadIdx.x == 0 && p.round_out)
        p.round_out[kRoundRec] = static_cast<float>(*reinterpret_cast<volatile uint32_t*>(p.arrive_ctr + 2));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Per-round result accounting (reference: analyze_results, ols_core/taskMgr/run_task.py:149-210, sums success /
// failed per device class on the host).  One CTA folds the per-client loss / non-finite flags / masks into an
// 8-float record: [masked loss sum / local_steps, finite participants, counters[0..3], non-finite clients,
// client-steps executed].  fedavg_finish_kernel then adds the records of all ranks.
// ---------------------------------------------------------------------------------------------
// With `class_id` (one small integer per client = its (dataset, device class) index) the record is followed by
// 2 floats per class: [failed_k, participants_k] — the per-class success / failed numbers the task status is computed
// from (ols_core/taskMgr/run_task.py:164-197, task_manager.py:739-754).
__global__ void round_stats_kernel(const float* __restrict__ loss_sum, const int* __restrict__ nonfinite,
                                   const uint8_t* __restrict__ participate, const uint8_t* __restrict__ step_active,
                                   const int* __restrict__ counters, int C, int S, float inv_steps,
                                   float* __restrict__ out, const int* __restrict__ class_id, int nclass) {
  float loss = 0.f, ok = 0.f, bad = 0.f, steps = 0.f;
  __shared__ float s_cls[2 * kMaxClasses];
  for (int i = threadIdx.x; i < 2 * kMaxClasses; i += blockDim.x) s_cls[i] = 0.f;
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const bool nf = nonfinite[c] != 0;
    const bool part = participate[c] != 0;
    if (part && !nf) { loss += loss_sum[c]; ok += 1.f; }
    if (nf) bad += 1.f;
    if (class_id) {
      const int k = class_id[c];
      if (k >= 0 && k < nclass && k < kMaxClasses) {
        if (nf) atomicAdd(&s_cls[2 * k], 1.f);
        if (part) atomicAdd(&s_cls[2 * k + 1], 1.f);
      }
    }
  }
  for (int i = threadIdx.x; i < C * S; i += blockDim.x) steps += step_active[i] ? 1.f : 0.f;
  __shared__ float red[4][32];
  float v[4] = {loss, ok, bad, steps};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float x = v[k];
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float x = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) x += red[threadIdx.x][w];
    v[0] = x;
    const int slot = threadIdx.x == 0 ? 0 : threadIdx.x == 1 ? 1 : threadIdx.x == 2 ? 6 : 7;
    out[slot] = threadIdx.x == 0 ? x * inv_steps : x;
  }
  if (threadIdx.x >= 4 && threadIdx.x < 8) out[threadIdx.x - 2] = (float)counters[threadIdx.x - 4];
  if (threadIdx.x < 2 * kMaxClasses) out[8 + threadIdx.x] = s_cls[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// DeviceFlow sampler: Philox4x32-10 keyed by (seed), counter (round, client, stream, 0).
// Bit-identical to olearning_sim_b200.deviceflow.philox (CPU) so a trace replays the same on either side.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct MaskParams {
  int num_clients;
  int client_base;                 // global id of local client 0 (so every rank draws the same stream per client)
  uint32_t seed_lo, seed_hi, round;
  const float* p_online;           // [C] probability that the client is online / sampled this round (or nullptr => 1)
  float drop_probability;          // DeviceFlow drop_simulation.drop_probability: update lost in flight
  const int32_t* arrive_tick;      // [C] or nullptr: dispatch slot of the client's update from the flow schedule (-1 = never)
  int32_t max_tick;                // updates with tick > max_tick are late -> discarded (clean_remain_message)
  const int32_t* n_samples;        // [C]
  const int32_t* local_steps;      // [C] heterogeneous local-step budget
  int max_steps;
  uint8_t* participate;            // [C] 1 => trains this round
  uint8_t* step_active;            // [max_steps][C]  1 => client runs local step s
  float* agg_weight;               // [C] n_c if the update arrives, else 0
  int32_t* counters;               // [4]: participants, dropped, late, arrived
};

__global__ void deviceflow_mask_kernel(const MaskParams p) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.num_clients) return;
  uint32_t r[4];
  philox4x32_10(p.round, (uint32_t)(p.client_base + c), 0u, 0u, p.seed_lo, p.seed_hi, r);
  const float u_on = (r[0] >> 8) * (1.f / 16777216.f);
  const float u_drop = (r[1] >> 8) * (1.f / 16777216.f);
  const float pon = p.p_online ? p.p_online[c] : 1.f;
  const bool part = u_on < pon;
  // reference semantics: forwarded iff random() > p  (p==1 drops everything, p==0 keeps everything)
  const bool dropped = part && (p.drop_probability >= 1.f || (p.drop_probability > 0.f && !(u_drop > p.drop_probability)));
  bool late = false;
  if (part && !dropped && p.arrive_tick) { const int t = p.arrive_tick[c]; late = (t < 0) || (t > p.max_tick); }
  const bool arrived = part && !dropped && !late;
  p.participate[c] = part ? 1 : 0;
  const int steps = p.local_steps ? min(p.local_steps[c], p.max_steps) : p.max_steps;
  for (int s = 0; s < p.max_steps; ++s) p.step_active[(size_t)s * p.num_clients + c] = (part && s < steps) ? 1 : 0;
  p.agg_weight[c] = arrived ? (float)p.n_samples[c] : 0.f;
  if (part) atomicAdd(&p.counters[0], 1);
  if (dropped) atomicAdd(&p.counters[1], 1);
  if (late) atomicAdd(&p.counters[2], 1);
  if (arrived) atomicAdd(&p.counters[3], 1);
}

// ---------------------------------------------------------------------------------------------------------
// Flow-dispatch schedule on the device (DeviceFlow flow mode, ols_core/deviceflow/non_grpc/dispatcher.py:174-252):
// message k of a flow's shelf is the k-th FINISHING participant of the round; slot s forwards amount[s] messages, the
// indices in its drop list are consumed but not forwarded, and whatever is left on the shelf after the last slot is
// discarded.  The host used to rank all clients with numpy every round (philox + argsort + H2D copy); these four
// kernels do it on the GPU from the per-task plan (cumulative amounts, uploaded once):
//   flow_keys     Philox stream of every GLOBAL client -> participation, 24-bit finishing key, histogram of the top 12 bits
//   flow_scan     exclusive scan of the 4096 buckets (one CTA)
//   flow_scatter  participants grouped by bucket
//   flow_ticks    rank inside the bucket (ties by client id = the stable argsort of the host twin) -> message index ->
//                 slot via binary search in the cumulative amounts -> arrive_tick of this rank's clients
// Every rank computes the same global ranking (counter-based RNG, no communication).
// ---------------------------------------------------------------------------------------------------------
constexpr int kFlowBuckets = 4096;
__global__ void flow_keys_kernel(int N, uint32_t seed_lo, uint32_t seed_hi, uint32_t round, const float* __restrict__ p_online,
                                 float scale, uint32_t* __restrict__ keys, int* __restrict__ hist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  uint32_t r[4];
  philox4x32_10(round, (uint32_t)i, 0u, 0u, seed_lo, seed_hi, r);
  const float u_on = (r[0] >> 8) * (1.f / 16777216.f);
  const bool part = u_on < p_online[i] * scale;
  const uint32_t key = r[2] >> 8;
  keys[i] = part ? key : 0xFFFFFFFFu;
  if (part) atomicAdd(&hist[key >> 12], 1);
}
__global__ void flow_scan_kernel(const int* __restrict__ hist, int* __restrict__ start, int* __restrict__ cursor) {
  // 4096 buckets, 1024 threads x 4; start[kFlowBuckets] = number of participants
  __shared__ int s_part[1024];
  const int t = threadIdx.x;
  int v[4], sum = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { v[k] = hist[t * 4 + k]; sum += v[k]; }
  s_part[t] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int add = t >= o ? s_part[t - o] : 0;
    __syncthreads();
    s_part[t] += add;
    __syncthreads();
  }
  int base = s_part[t] - sum;
#pragma unroll
  for (int k = 0; k < 4; ++k) { start[t * 4 + k] = base; cursor[t * 4 + k] = 0; base += v[k]; }
  if (t == 1023) start[kFlowBuckets] = s_part[1023];
}
__global__ void flow_scatter_kernel(int N, const uint32_t* __restrict__ keys, const int* __restrict__ start,
                                    int* __restrict__ cursor, int* __restrict__ items) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint32_t k = keys[i];
  if (k == 0xFFFFFFFFu) return;
  const int b = k >> 12;
  items[start[b] + atomicAdd(&cursor[b], 1)] = i;
}
__global__ void flow_ticks_kernel(int c0, int C, const uint32_t* __restrict__ keys, const int* __restrict__ start,
                                  const int* __restrict__ items, const int* __restrict__ cum, int nslots,
                                  const uint8_t* __restrict__ dropmap, int32_t* __restrict__ tick) {
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= C) return;
  const int i = c0 + li;
  const uint32_t k = keys[i];
  if (k == 0xFFFFFFFFu) { tick[li] = -1; return; }
  const int b = k >> 12, s0 = start[b], s1 = start[b + 1];
  int rank = s0;
  for (int j = s0; j < s1; ++j) {
    const int o = items[j];
    const uint32_t ko = keys[o];
    rank += (ko < k || (ko == k && o < i)) ? 1 : 0;
  }
  const int total = nslots > 0 ? cum[nslots - 1] : 0;
  if (rank >= total) { tick[li] = -1; return; }          // beyond the schedule: left on the shelf, discarded
  int lo = 0, hi = nslots - 1;                            // first slot with cum[slot] > rank
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (cum[mid] > rank) hi = mid; else lo = mid + 1; }
  tick[li] = (dropmap && dropmap[rank]) ? -1 : lo;
}

// Rate-curve integration (DeviceFlow specific_interval, ols_core/deviceflow/non_grpc/strategy.py:339-349): area of the
// positive trapezoids of every 1-second bin.  `y` holds f sampled at the 101 nodes of each bin (nodes are NOT shared
// between bins, matching the reference's per-bin linspace).  Double precision with explicitly rounded operations in the
// python loop's order (no FMA contraction), so the result is bit-identical to Strategy.interval_areas on the host and a
// schedule derived from it is the same on CPU and GPU.
__global__ void rate_curve_area_kernel(const double* __restrict__ y, double* __restrict__ area, int bins) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bins) return;
  const double* yy = y + (size_t)b * 101;
  double a = 0.0;
  const double w = 1.0 / 100.0;
  for (int j = 1; j <= 100; ++j) {
    const double s = __dmul_rn(__dmul_rn(0.5, __dadd_rn(yy[j], yy[j - 1])), w);
    if (s > 0.0) a = __dadd_rn(a, s);
  }
  area[b] = a;
}

// Generic per-client SGD / FedProx step on packed rows (used by the non-fused model families):
//   w_c <- w_c - lr * (g_c + mu * (w_c - w_g)),  skipped where step_active[c] == 0
__global__ void client_sgd_kernel(float* __restrict__ w, const float* __restrict__ g, const float* __restrict__ wg,
                                  const uint8_t* __restrict__ active, int num_clients, int P, float lr, float mu,
                                  float momentum, float* __restrict__ mom) {
  const int i4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int c = blockIdx.y;
  if (i4 >= P || c >= num_clients) return;
  if (active && !active[c]) return;
  const size_t o = (size_t)c * P + i4;
  float4 wv = *reinterpret_cast<float4*>(w + o);
  const float4 gv = *reinterpret_cast<const float4*>(g + o);
  float gg[4] = {gv.x, gv.y, gv.z, gv.w};
  float ww[4] = {wv.x, wv.y, wv.z, wv.w};
  if (wg) {
    const float4 a = *reinterpret_cast<const float4*>(wg + i4);
    gg[0] += mu * (ww[0] - a.x); gg[1] += mu * (ww[1] - a.y); gg[2] += mu * (ww[2] - a.z); gg[3] += mu * (ww[3] - a.w);
  }
  if (mom) {
    float4 mv = *reinterpret_cast<float4*>(mom + o);
    float mm[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { mm[e] = momentum * mm[e] + gg[e]; gg[e] = mm[e]; }
    *reinterpret_cast<float4*>(mom + o) = make_float4(mm[0], mm[1], mm[2], mm[3]);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) ww[e] -= lr * gg[e];
  *reinterpret_cast<float4*>(w + o) = make_float4(ww[0], ww[1], ww[2], ww[3]);
}

}  // namespace flk
