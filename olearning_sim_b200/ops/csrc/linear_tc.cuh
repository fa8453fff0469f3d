// linear_tc.cuh — generic per-client ("grouped over clients") GEMM on tcgen05:   D_c[M,N] = A_c[M,K] * B_c[N,K]^T
//
// One persistent kernel covers the three GEMMs of a fully-connected layer by choosing operand majors and epilogue:
//
//   forward   Y^T[f, s]  = W[f, :] . X[s, :]        A = W   (K-major)   B = X   (K-major)   EPI_BIAS_ACT_T
//   dgrad     dX^T[i, s] = sum_o W[o, i] dY[s, o]    A = W   (MN-major)  B = dY  (K-major)   EPI_MASK_T  (+ bias SGD)
//   wgrad     dW[o, i]   = sum_s dY[s, o] X[s, i]    A = dY  (MN-major)  B = X   (MN-major)  EPI_SGD
//
// The batch (32 samples per client) is far below the 128-row MMA tile, so the output *features* are put on the M axis
// ("swap" formulation) and the batch on N; results are stored transposed so activations stay [sample][feature].
// Operand tiles are TMA boxes of 64 elements (128 B, SWIZZLE_128B) x rows; out-of-bounds rows / K tails are zero-filled
// by TMA, so no shape has to be padded in memory (only the innermost stride must be a multiple of 16 B).
// Same warp-specialised structure as conv_tc.cuh (TMA warp, MMA warp, 4 epilogue warps, TMEM double buffering).
#pragma once
#include "tc05.cuh"

namespace lintc {
using namespace tc05;

constexpr int kThreads = 192;
constexpr int kEpiWarps = 4;
enum { EPI_BIAS_ACT_T = 0, EPI_MASK_T = 1, EPI_SGD = 2 };

struct GemmBars {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

struct GemmParams {
  int num_clients, M, N, K;            // per-client problem (K = reduction length)
  int m_tiles, n_tiles, k_chunks;
  const uint8_t* active;
  float lr, mu;
  // EPI_BIAS_ACT_T / EPI_MASK_T : transposed bf16 output  out_t[c*out_stride + n*out_ld + m]
  __nv_bfloat16* out_t; long long out_stride; int out_ld;
  const float* bias; long long bias_stride; int relu;
  const __nv_bfloat16* mask_t;         // forward activation with the same indexing as out_t (ReLU gate)
  float* b_master; long long b_stride; const float* b_global;   // EPI_MASK_T: bias of the layer below, SGD on row sums
  float* db_out;                       // optional raw bias gradient [C][M] (tests); then no bias update
  // EPI_SGD : W[c*w_stride + m*w_ld + n]
  float* w_master; long long w_stride; int w_ld; const float* w_global;
  __nv_bfloat16* w_bf16; long long wb_stride;
  float* grad_out; long long g_stride;  // optional raw gradient (tests); then no update
};

template <bool A_MN, bool B_MN, int N_TILE, int KC, int NSTAGE>
struct GemmCfg {
  static constexpr int A_BYTES = A_MN ? 2 * KC * 128 : 128 * 128;
  static constexpr int B_BYTES = B_MN ? (N_TILE / 64) * KC * 128 : N_TILE * 128;
  static constexpr int STAGE_BYTES = ((A_BYTES + B_BYTES + 1023) / 1024) * 1024;
  static constexpr int TMEM_COLS = (2 * N_TILE <= 32) ? 32 : (2 * N_TILE <= 64 ? 64 : (2 * N_TILE <= 128 ? 128 : (2 * N_TILE <= 256 ? 256 : 512)));
  static constexpr int SMEM_BYTES = 1024 + NSTAGE * STAGE_BYTES + (int)sizeof(GemmBars);
  static_assert(A_MN || KC == 64, "K-major tiles use 64-element (128 B) rows");
  static_assert(B_MN || KC == 64, "K-major tiles use 64-element (128 B) rows");
  static_assert(!B_MN || N_TILE % 64 == 0, "MN-major B is loaded in 64-column boxes");
  static_assert(N_TILE % 32 == 0 && N_TILE <= 256, "epilogue works on 32-column chunks");
};

template <bool A_MN, bool B_MN, int N_TILE, int KC, int EPI, int NSTAGE>
__global__ void __launch_bounds__(kThreads, 1)
client_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<A_MN, B_MN, N_TILE, KC, NSTAGE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GemmBars* bars = reinterpret_cast<GemmBars*>(smem + NSTAGE * Cfg::STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(&bars->tmem_base, Cfg::TMEM_COLS); tmem_relinquish(); }
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const int tiles_per_client = p.m_tiles * p.n_tiles;
  const int num_work = p.num_clients * tiles_per_client;

  if (warp == 4) {
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / tiles_per_client, r = wi % tiles_per_client;
      if (p.active && !p.active[c]) continue;
      const int m0 = (r / p.n_tiles) * 128, n0 = (r % p.n_tiles) * N_TILE;
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        mbar_wait(&bars->empty[stage], phase ^ 1);
        if (leader) {
          uint8_t* a_dst = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* b_dst = a_dst + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&bars->full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
          const int k0 = kc * KC;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_3d(a_dst + j * KC * 128, &tmA, &bars->full[stage], m0 + 64 * j, k0, c);
          } else {
            tma_load_3d(a_dst, &tmA, &bars->full[stage], k0, m0, c);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < N_TILE / 64; ++j) tma_load_3d(b_dst + j * KC * 128, &tmB, &bars->full[stage], n0 + 64 * j, k0, c);
          } else {
            tma_load_3d(b_dst, &tmB, &bars->full[stage], k0, n0, c);
          }
        }
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    const bool leader = elect_one();
    constexpr uint32_t IDESC = make_idesc_bf16(128, N_TILE, A_MN ? 1 : 0, B_MN ? 1 : 0);
    const uint32_t sbase = smem_u32(smem);
    const uint64_t a_desc0 = A_MN ? make_smem_desc(sbase, KC * 128, 1024, LT_SW128) : make_smem_desc(sbase, 16, 1024, LT_SW128);
    const uint64_t b_desc0 = B_MN ? make_smem_desc(sbase + Cfg::A_BYTES, KC * 128, 1024, LT_SW128)
                                  : make_smem_desc(sbase + Cfg::A_BYTES, 16, 1024, LT_SW128);
    constexpr uint32_t A_STEP = A_MN ? 2048 : 32, B_STEP = B_MN ? 2048 : 32;
    int stage = 0; uint32_t phase = 0; int buf = 0; uint32_t tphase0 = 0, tphase1 = 0;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / tiles_per_client;
      if (p.active && !p.active[c]) continue;
      mbar_wait(&bars->tempty[buf], (buf ? tphase1 : tphase0) ^ 1);
      fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * N_TILE;
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        mbar_wait(&bars->full[stage], phase);
        fence_after_sync();
        if (leader) {
          const uint64_t a_s = desc_advance(a_desc0, stage * Cfg::STAGE_BYTES);
          const uint64_t b_s = desc_advance(b_desc0, stage * Cfg::STAGE_BYTES);
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks)
            umma_bf16(d_tmem, desc_advance(a_s, ks * A_STEP), desc_advance(b_s, ks * B_STEP), IDESC, (kc | ks) ? 1u : 0u);
          umma_commit(&bars->empty[stage]);
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(&bars->tfull[buf]);
      __syncwarp();
      if (buf) tphase1 ^= 1; else tphase0 ^= 1;
      buf ^= 1;
    }
  } else {
    int buf = 0; uint32_t tphase0 = 0, tphase1 = 0;
    const int q = warp;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / tiles_per_client, r = wi % tiles_per_client;
      if (p.active && !p.active[c]) continue;
      const int m0 = (r / p.n_tiles) * 128, n0 = (r % p.n_tiles) * N_TILE;
      const int m = m0 + 32 * q + lane;
      const bool mvalid = m < p.M;
      mbar_wait(&bars->tfull[buf], buf ? tphase1 : tphase0);
      fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + buf * N_TILE;
      float rowsum = 0.f;
#pragma unroll 1
      for (int j = 0; j < N_TILE / 32; ++j) {
        const int nb = n0 + j * 32;
        if (nb >= p.N) break;                         // warp-uniform
        uint32_t v[32];
        tmem_ld32(taddr + j * 32, v);
        tmem_ld_wait();
        if (EPI == EPI_BIAS_ACT_T) {
          const float b = (mvalid && p.bias) ? __ldg(&p.bias[(size_t)c * p.bias_stride + m]) : 0.f;
          if (mvalid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (nb + i < p.N) {
                float f = __uint_as_float(v[i]) + b;
                if (p.relu) f = fmaxf(f, 0.f);
                p.out_t[(size_t)c * p.out_stride + (size_t)(nb + i) * p.out_ld + m] = __float2bfloat16_rn(f);
              }
            }
          }
        } else if (EPI == EPI_MASK_T) {
          if (mvalid) {
            __nv_bfloat16 mk[32];
#pragma unroll
            for (int i = 0; i < 32; ++i)
              mk[i] = (nb + i < p.N) ? p.mask_t[(size_t)c * p.out_stride + (size_t)(nb + i) * p.out_ld + m] : __float2bfloat16_rn(0.f);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (nb + i < p.N) {
                const float g = __bfloat162float(mk[i]) > 0.f ? __uint_as_float(v[i]) : 0.f;
                rowsum += g;
                p.out_t[(size_t)c * p.out_stride + (size_t)(nb + i) * p.out_ld + m] = __float2bfloat16_rn(g);
              }
            }
          }
        } else {  // EPI_SGD
          if (mvalid) {
            if (p.grad_out) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (nb + i < p.N) p.grad_out[(size_t)c * p.g_stride + (size_t)m * p.w_ld + nb + i] = __uint_as_float(v[i]);
            } else {
              float* wrow = p.w_master + (size_t)c * p.w_stride + (size_t)m * p.w_ld + nb;
              const float* grow = p.w_global ? p.w_global + (size_t)m * p.w_ld + nb : nullptr;
              __nv_bfloat16* brow = p.w_bf16 + (size_t)c * p.wb_stride + (size_t)m * p.w_ld + nb;
              float wv[32], av[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) wv[i] = (nb + i < p.N) ? wrow[i] : 0.f;       // all loads before any store
              if (grow) {
#pragma unroll
                for (int i = 0; i < 32; ++i) av[i] = (nb + i < p.N) ? __ldg(&grow[i]) : 0.f;
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (nb + i < p.N) {
                  float g = __uint_as_float(v[i]);
                  if (grow) g += p.mu * (wv[i] - av[i]);
                  const float w = wv[i] - p.lr * g;
                  wrow[i] = w;
                  brow[i] = __float2bfloat16_rn(w);
                }
              }
            }
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tempty[buf]);
      if (EPI == EPI_MASK_T && mvalid && (r % p.n_tiles) == 0) {
        // bias of the layer whose pre-activation gradient this is: db[m] = sum over the batch (n_tiles == 1)
        if (p.db_out) {
          p.db_out[(size_t)c * p.M + m] = rowsum;
        } else if (p.b_master) {
          float* bp = p.b_master + (size_t)c * p.b_stride + m;
          float g = rowsum;
          const float b = *bp;
          if (p.b_global) g += p.mu * (b - p.b_global[m]);
          *bp = b - p.lr * g;
        }
      }
      if (buf) tphase1 ^= 1; else tphase0 ^= 1;
      buf ^= 1;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------
// softmax cross-entropy over per-client logits (transposed-store layout [C][B][LD]) + last-layer bias SGD
// ---------------------------------------------------------------------------------------------------------
struct XentParams {
  int num_clients, batch, classes, ld;     // logits[c][s][ld] bf16 (ld >= classes, multiple of 8)
  const uint8_t* active;
  const __nv_bfloat16* logits;
  const int32_t* labels; int label_stride, label_off;
  __nv_bfloat16* dlogits;                  // same layout, (softmax - onehot)/batch, padding columns zero
  float* loss_sum; int32_t* nonfinite;
  float* b_master; long long b_stride; const float* b_global; float lr, mu;   // bias of the last layer (may be null)
  float* db_out;                           // optional raw bias gradient [C][classes]
};

__global__ void client_xent_kernel(const XentParams p) {
  // one warp per client, lane = sample (batch <= 32)
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= p.num_clients) return;
  if (p.active && !p.active[c]) return;
  const bool sv = lane < p.batch;
  float lg[16], dl[16];
  float loss = 0.f;
  const size_t base = ((size_t)c * p.batch + lane) * p.ld;
  if (sv) {
    float mx = -1e30f;
    for (int j = 0; j < p.classes; ++j) { lg[j] = __bfloat162float(p.logits[base + j]); mx = fmaxf(mx, lg[j]); }
    float sum = 0.f;
    for (int j = 0; j < p.classes; ++j) { dl[j] = __expf(lg[j] - mx); sum += dl[j]; }
    const int y = p.labels[(size_t)c * p.label_stride + p.label_off + lane];
    const float inv = 1.f / sum, ib = 1.f / (float)p.batch;
    for (int j = 0; j < p.classes; ++j) {
      if (j == y) loss = -(lg[j] - mx - __logf(sum));
      dl[j] = (dl[j] * inv - (j == y ? 1.f : 0.f)) * ib;
      p.dlogits[base + j] = __float2bfloat16_rn(dl[j]);
    }
    for (int j = p.classes; j < p.ld; ++j) p.dlogits[base + j] = __float2bfloat16_rn(0.f);
  } else {
    for (int j = 0; j < p.classes; ++j) dl[j] = 0.f;
  }
  float ls = loss;
  for (int o = 16; o; o >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, o);
  if (lane == 0) {
    const float mean = ls / (float)p.batch;
    p.loss_sum[c] += mean;
    if (!(fabsf(mean) < 3.0e38f)) p.nonfinite[c] = 1;
  }
  for (int j = 0; j < p.classes; ++j) {
    float g = dl[j];
    for (int o = 16; o; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
    if (lane == 0) {
      if (p.db_out) p.db_out[(size_t)c * p.classes + j] = g;
      else if (p.b_master) {
        float* bp = p.b_master + (size_t)c * p.b_stride + j;
        const float b = *bp;
        if (p.b_global) g += p.mu * (b - p.b_global[j]);
        *bp = b - p.lr * g;
      }
    }
  }
}

// generic packed-replica broadcast for model families without special operand layouts:
//   master[c][i] = global[i] (pulled from the shard owner), bf16 copy alongside
struct BcastGenericParams {
  void* global_peers[8]; int nranks, shard_len, P;
  float* global_local; float* master; __nv_bfloat16* master_bf16; int num_clients; const uint8_t* participate;
  int P_bf16; long long bf16_stride;     // the bf16 operand copy covers elements [0, P_bf16) of every row
};
__global__ void bcast_scatter_generic_kernel(const BcastGenericParams p) {
  const int i4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= p.P) return;
  const int owner = min(i4 / p.shard_len, p.nranks - 1);
  const float4 v = *reinterpret_cast<const float4*>(static_cast<const float*>(p.global_peers[owner]) + i4);
  if (blockIdx.y == 0) *reinterpret_cast<float4*>(p.global_local + i4) = v;
  uint2 o;
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
  for (int c = blockIdx.y; c < p.num_clients; c += gridDim.y) {
    if (p.participate && !p.participate[c]) continue;
    *reinterpret_cast<float4*>(p.master + (size_t)c * p.P + i4) = v;
    if (p.master_bf16 && i4 < p.P_bf16) *reinterpret_cast<uint2*>(p.master_bf16 + (size_t)c * p.bf16_stride + i4) = o;
  }
}

}  // namespace lintc
