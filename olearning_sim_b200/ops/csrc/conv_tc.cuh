// conv_tc.cuh — per-client ("grouped over clients") 5x5 convolution family on tcgen05 tensor cores.
//
// Every simulated client owns its own weights, so a federated local step is C independent small
// convolutions.  Each CTA is persistent over (client, tile-range) work items and keeps that client's
// weights resident in shared memory while the activation tiles stream through a TMA ring:
//
//   * implicit GEMM: M = 128 output pixels (an 8x16 patch of one image), N = output channels,
//     K = taps x input channels.  The A operand is never materialised: one 4-D tiled TMA box of
//     12x16 pixels (8 rows + 4 halo rows) per horizontal tap lands in smem with hardware swizzle and
//     zero-filled borders ("same" padding); the five vertical taps are pure start-address shifts of
//     the UMMA shared-memory descriptor (one image row = 16 pixel-rows = a whole number of swizzle atoms).
//   * accumulators live in TMEM (double buffered) and are drained by four epilogue warps with
//     tcgen05.ld; the epilogues are fused: bias+ReLU+2x2-maxpool(+argmax code) for the forward,
//     un-pool + ReLU-backward + bias-grad (+SGD on the bias) for the data gradient.
//   * weight gradient: both operands are MN-major (pixels are the reduction dim), all taps accumulate
//     in TMEM simultaneously (7 x [128 x 64] fp32 tiles = 448 columns for 5x5x32->64) while the pixels
//     stream through once; the epilogue applies the SGD / FedProx update to the fp32 master weights
//     and refreshes the bf16 operand copies, so gradients never touch HBM.
//
// Reference parity: this replaces the per-client `os.system("python3 <operator>")` local training
// step of the reference (ols_core/taskMgr/utils/utils_run_task.py:481-514).
#pragma once
#include "tc05.cuh"

namespace convtc {
using namespace tc05;

constexpr int kEpiWarps = 4;
constexpr int kThreads = 192;  // warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer
constexpr int kTileH = 8, kTileW = 16, kHalo = 2, kBoxH = kTileH + 2 * kHalo;  // 12
constexpr int kNKH = 5;

struct Barriers {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t wfull;
  uint64_t wempty;
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

// ================================================================================================
// Forward / data-gradient kernel
// ================================================================================================
enum { EPI_POOL = 0, EPI_UNPOOL = 1 };

struct FwdParams {
  int num_clients;
  int nsplit;            // work items per client
  int a_n_stride;        // samples per client in the A tensor
  int a_n_off;           // first sample of this step's batch inside the client's block
  const uint8_t* active; // [C] or nullptr: 0 => client skipped this step (straggler / not sampled)
  // EPI_POOL
  const float* bias;     // [C][bias_stride] fp32
  int bias_stride;
  __nv_bfloat16* out;    // [C*BATCH][H/2][W/2][BN]
  uint8_t* code;         // same shape, argmax position 0..3
  // EPI_UNPOOL
  const __nv_bfloat16* pooled;  // [C*BATCH][H][W][BN]   forward activations of the layer below (post-pool)
  const uint8_t* pcode;         // same shape
  __nv_bfloat16* dpre;          // [C*BATCH][2H][2W][BN]  gradient w.r.t. the pre-pool conv output
  float* bias_master;           // [C][bias_stride] fp32, updated in place:  b -= lr*(db + mu*(b-bg))
  const float* bias_global;     // [bias_stride] fp32 (FedProx anchor) or nullptr
  float lr, mu;
  float* dbias_out;             // optional [C][BN]: raw bias grad (tests); when set the bias is NOT updated
};

template <int ROWB, int NKW, int BN, int IMG_H, int IMG_W, int BATCH, bool FLIP, int EPI, int NSTAGE>
struct FwdCfg {
  static constexpr int KT = ROWB / 2;
  static constexpr int NTAPS = NKW * kNKH;
  static constexpr int W_TAP_BYTES = BN * ROWB;
  static constexpr int W_BYTES = NTAPS * W_TAP_BYTES;
  static constexpr int STAGE_BYTES = kBoxH * kTileW * ROWB;
  static constexpr int TILES_X = IMG_W / kTileW, TILES_Y = IMG_H / kTileH;
  static constexpr int TILES_PER_CLIENT = BATCH * TILES_X * TILES_Y;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : 256));
  static constexpr int SMEM_BYTES = 1024 + W_BYTES + NSTAGE * STAGE_BYTES + (int)sizeof(Barriers);
  static_assert(W_TAP_BYTES % 1024 == 0 || ROWB < 128, "weight tap chunk must keep swizzle alignment");
  static_assert(BN % 32 == 0, "epilogue works on 32-column chunks");
};

template <int ROWB, int NKW, int BN, int IMG_H, int IMG_W, int BATCH, bool FLIP, int EPI, int NSTAGE>
__global__ void __launch_bounds__(kThreads, 1)
conv_taps_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const FwdParams p) {
  using Cfg = FwdCfg<ROWB, NKW, BN, IMG_H, IMG_W, BATCH, FLIP, EPI, NSTAGE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_sm = smem;
  uint8_t* a_sm = smem + Cfg::W_BYTES;
  Barriers* bars = reinterpret_cast<Barriers*>(a_sm + NSTAGE * Cfg::STAGE_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    mbar_init(&bars->wfull, 1);
    mbar_init(&bars->wempty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(&bars->tmem_base, Cfg::TMEM_COLS); tmem_relinquish(); }
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  const int num_work = p.num_clients * p.nsplit;
  const int tiles_per_work = (Cfg::TILES_PER_CLIENT + p.nsplit - 1) / p.nsplit;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0; int cur_client = -1; uint32_t wphase = 0; bool first_w = true;
      for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
        const int c = wi / p.nsplit, sp = wi % p.nsplit;
        if (p.active && !p.active[c]) continue;
        if (c != cur_client) {
          if (!first_w) { mbar_wait(&bars->wempty, wphase); wphase ^= 1; }
          first_w = false; cur_client = c;
          mbar_arrive_expect_tx(&bars->wfull, Cfg::W_BYTES);
          for (int t = 0; t < Cfg::NTAPS; ++t)
            // forward copy  [C][BN][NTAPS*KT]  -> box (KT, BN, 1) at (t*KT, 0, c)
            // dgrad copy    [C][NTAPS*BN][K]   -> box (K,  BN, 1) at (0, t*BN, c)
            tma_load_3d(w_sm + t * Cfg::W_TAP_BYTES, &tmW, &bars->wfull, FLIP ? 0 : t * Cfg::KT,
                        FLIP ? t * BN : 0, c);
        }
        const int t0 = sp * tiles_per_work;
        const int t1 = min(t0 + tiles_per_work, Cfg::TILES_PER_CLIENT);
        for (int t = t0; t < t1; ++t) {
          const int s = t / (Cfg::TILES_X * Cfg::TILES_Y);
          const int r = t % (Cfg::TILES_X * Cfg::TILES_Y);
          const int h0 = (r / Cfg::TILES_X) * kTileH, w0 = (r % Cfg::TILES_X) * kTileW;
          const int n = c * p.a_n_stride + p.a_n_off + s;
#pragma unroll 1
          for (int kw = 0; kw < NKW; ++kw) {
            mbar_wait(&bars->empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&bars->full[stage], Cfg::STAGE_BYTES);
            const int dw = (NKW == 1) ? 0 : (FLIP ? (kHalo - kw) : (kw - kHalo));
            tma_load_4d(a_sm + stage * Cfg::STAGE_BYTES, &tmA, &bars->full[stage], 0, w0 + dw, h0 - kHalo, n);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t LT = layout_type_for_row_bytes(ROWB);
      constexpr uint32_t IDESC = make_idesc_bf16(128, BN, 0, 0);
      int stage = 0; uint32_t phase = 0; int cur_client = -1; uint32_t wphase = 0;
      int buf = 0; uint32_t tphase[2] = {0, 0};
      for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
        const int c = wi / p.nsplit, sp = wi % p.nsplit;
        if (p.active && !p.active[c]) continue;
        // is this the last work item this CTA will run for client c?
        bool last_of_client = true;
        {
          const int nwi = wi + gridDim.x;
          if (nwi < num_work && (nwi / p.nsplit) == c) last_of_client = false;
        }
        if (c != cur_client) { mbar_wait(&bars->wfull, wphase); wphase ^= 1; cur_client = c; }
        const int t0 = sp * tiles_per_work;
        const int t1 = min(t0 + tiles_per_work, Cfg::TILES_PER_CLIENT);
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&bars->tempty[buf], tphase[buf] ^ 1);
          fence_after_sync();
          const uint32_t d_tmem = tmem_base + buf * BN;
          uint32_t acc = 0;
#pragma unroll 1
          for (int kw = 0; kw < NKW; ++kw) {
            mbar_wait(&bars->full[stage], phase);
            fence_after_sync();
            const uint32_t a_base = smem_u32(a_sm + stage * Cfg::STAGE_BYTES);
#pragma unroll
            for (int kh = 0; kh < kNKH; ++kh) {
              const int row_shift = FLIP ? (2 * kHalo - kh) : kh;  // image rows above the tile start
              const int tap = kh * NKW + kw;
              const uint64_t a_desc = make_smem_desc(a_base + row_shift * kTileW * ROWB, 16, 8 * ROWB, LT);
              const uint64_t b_desc = make_smem_desc(smem_u32(w_sm + tap * Cfg::W_TAP_BYTES), 16, 8 * ROWB, LT);
#pragma unroll
              for (int ks = 0; ks < ROWB / 32; ++ks) {
                umma_bf16(d_tmem, desc_advance(a_desc, ks * 32), desc_advance(b_desc, ks * 32), IDESC, acc);
                acc = 1;
              }
            }
            umma_commit(&bars->empty[stage]);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
          umma_commit(&bars->tfull[buf]);
          tphase[buf] ^= 1;
          buf ^= 1;
        }
        if (last_of_client) umma_commit(&bars->wempty);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (TMEM lanes 32*warp ..)
    int buf = 0; uint32_t tphase[2] = {0, 0};
    const int q = warp;  // 0..3
    const int hl = 2 * q + (lane >> 4), wl = lane & 15;
    float dbacc[(EPI == EPI_UNPOOL) ? BN : 1];
    if (EPI == EPI_UNPOOL) {
#pragma unroll
      for (int i = 0; i < BN; ++i) dbacc[i] = 0.f;
    }
    int cur_client = -1;
    bool had_sp0 = false;
    __shared__ float s_db[(EPI == EPI_UNPOOL) ? BN : 1];

    auto flush_bias = [&](int c) {
      if (EPI != EPI_UNPOOL) return;
      // reduce the per-thread partial bias gradients of the 128 epilogue threads, then update the bias
      asm volatile("bar.sync 1, 128;");
      if (threadIdx.x < BN) s_db[threadIdx.x] = 0.f;
      asm volatile("bar.sync 1, 128;");
#pragma unroll
      for (int i = 0; i < BN; ++i) {
        float v = dbacc[i];
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        if (lane == 0) atomicAdd(&s_db[i], v);
        dbacc[i] = 0.f;
      }
      asm volatile("bar.sync 1, 128;");
      if (threadIdx.x < BN) {
        const float g = s_db[threadIdx.x];
        if (p.dbias_out) {
          atomicAdd(&p.dbias_out[(size_t)c * BN + threadIdx.x], g);
        } else {
          // several CTAs may share a client (nsplit>1): the bias update is linear in g, so apply the
          // partial gradient atomically; the FedProx pull is applied by the split-0 CTA only.
          float* bp = &p.bias_master[(size_t)c * p.bias_stride + threadIdx.x];
          float upd = -p.lr * g;
          if (had_sp0 && p.bias_global) upd -= p.lr * p.mu * (*bp - p.bias_global[threadIdx.x]);
          atomicAdd(bp, upd);
        }
      }
      asm volatile("bar.sync 1, 128;");
    };

    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / p.nsplit, sp = wi % p.nsplit;
      if (p.active && !p.active[c]) continue;
      if (c != cur_client) {
        if (cur_client >= 0) flush_bias(cur_client);
        cur_client = c;
        had_sp0 = false;
      }
      if (sp == 0) had_sp0 = true;
      const int t0 = sp * tiles_per_work;
      const int t1 = min(t0 + tiles_per_work, Cfg::TILES_PER_CLIENT);
      for (int t = t0; t < t1; ++t) {
        const int s = t / (Cfg::TILES_X * Cfg::TILES_Y);
        const int r = t % (Cfg::TILES_X * Cfg::TILES_Y);
        const int h0 = (r / Cfg::TILES_X) * kTileH, w0 = (r % Cfg::TILES_X) * kTileW;
        const size_t n_out = (size_t)c * BATCH + s;
        mbar_wait(&bars->tfull[buf], tphase[buf]);
        fence_after_sync();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + buf * BN;

        if (EPI == EPI_POOL) {
          const int pos = ((lane >> 4) << 1) | (lane & 1);      // position inside the 2x2 window
          const int sub = pos;                                   // which 8-channel slice this lane stores
          const int ph = (h0 >> 1) + q, pw = (w0 >> 1) + (wl >> 1);
          const size_t obase = ((n_out * (IMG_H / 2) + ph) * (IMG_W / 2) + pw) * BN;
          const float* bias = p.bias + (size_t)c * p.bias_stride;
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) {
            uint32_t v[32];
            tmem_ld32(taddr + j * 32, v);
            tmem_ld_wait();
            uint32_t m[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float f = __uint_as_float(v[i]) + __ldg(&bias[j * 32 + i]);
              f = fmaxf(f, 0.f);
              // non-negative floats order like unsigned ints: stash (3-pos) in the two LSBs so one max
              // yields both the pooled value and its argmax (ties -> smallest position)
              uint32_t u = (__float_as_uint(f) & ~3u) | static_cast<uint32_t>(3 - pos);
              u = max(u, __shfl_xor_sync(0xffffffffu, u, 1));
              u = max(u, __shfl_xor_sync(0xffffffffu, u, 16));
              m[i] = u;
            }
            uint32_t o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
              o[k] = sub == 0 ? m[k] : (sub == 1 ? m[8 + k] : (sub == 2 ? m[16 + k] : m[24 + k]));
            uint4 pv;
            pv.x = pack_bf16x2(__uint_as_float(o[0] & ~3u), __uint_as_float(o[1] & ~3u));
            pv.y = pack_bf16x2(__uint_as_float(o[2] & ~3u), __uint_as_float(o[3] & ~3u));
            pv.z = pack_bf16x2(__uint_as_float(o[4] & ~3u), __uint_as_float(o[5] & ~3u));
            pv.w = pack_bf16x2(__uint_as_float(o[6] & ~3u), __uint_as_float(o[7] & ~3u));
            uint2 cv;
            cv.x = (3 - (o[0] & 3)) | ((3 - (o[1] & 3)) << 8) | ((3 - (o[2] & 3)) << 16) | ((3 - (o[3] & 3)) << 24);
            cv.y = (3 - (o[4] & 3)) | ((3 - (o[5] & 3)) << 8) | ((3 - (o[6] & 3)) << 16) | ((3 - (o[7] & 3)) << 24);
            const size_t off = obase + j * 32 + sub * 8;
            *reinterpret_cast<uint4*>(p.out + off) = pv;
            *reinterpret_cast<uint2*>(p.code + off) = cv;
          }
        } else {
          // EPI_UNPOOL: this lane owns pooled pixel (h0+hl, wl) of sample n_out; BN == 32 channels.
          const int h = h0 + hl, w = w0 + wl;
          const size_t pbase = ((n_out * IMG_H + h) * IMG_W + w) * BN;
          uint32_t v[32];
          tmem_ld32(taddr, v);
          // forward value (ReLU mask) and argmax code of the layer below
          uint4 pq[BN / 8];
          uint4 cq[BN / 16];
#pragma unroll
          for (int k = 0; k < BN / 8; ++k) pq[k] = *reinterpret_cast<const uint4*>(p.pooled + pbase + k * 8);
#pragma unroll
          for (int k = 0; k < BN / 16; ++k) cq[k] = *reinterpret_cast<const uint4*>(p.pcode + pbase + k * 16);
          tmem_ld_wait();
          const uint32_t* pw32 = reinterpret_cast<const uint32_t*>(pq);
          const uint8_t* cb = reinterpret_cast<const uint8_t*>(cq);
          float g[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const uint32_t pr = pw32[i >> 1];
            const float pv = (i & 1) ? bf16hi(pr) : bf16lo(pr);
            g[i] = pv > 0.f ? __uint_as_float(v[i]) : 0.f;
            dbacc[i] += g[i];
          }
#pragma unroll
          for (int pos = 0; pos < 4; ++pos) {
            const size_t obase =
                ((n_out * (2 * IMG_H) + (2 * h + (pos >> 1))) * (2 * IMG_W) + (2 * w + (pos & 1))) * BN;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              uint4 o;
              uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i0 = k * 8 + e * 2;
                const float a = (cb[i0] == pos) ? g[i0] : 0.f;
                const float b = (cb[i0 + 1] == pos) ? g[i0 + 1] : 0.f;
                ow[e] = pack_bf16x2(a, b);
              }
              *reinterpret_cast<uint4*>(p.dpre + obase + k * 8) = o;
            }
          }
        }
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tempty[buf]);
        tphase[buf] ^= 1;
        buf ^= 1;
      }
    }
    if (cur_client >= 0) flush_bias(cur_client);
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ================================================================================================
// Weight-gradient kernel (+ fused SGD / FedProx update of master and operand copies)
// ================================================================================================
struct WgradParams {
  int num_clients;
  int x_n_stride, x_n_off;       // sample addressing of the X tensor (layer input)
  const uint8_t* active;
  float* w_master;               // [C][p_stride] fp32; this layer's weights start at w_off
  const float* w_global;         // [p_stride] fp32 FedProx anchor or nullptr
  size_t p_stride;               // floats per client in the packed parameter vector
  size_t w_off;                  // offset of this layer's weight block inside the packed vector
  __nv_bfloat16* w_fwd;          // [C][BN][NTAPS*KT] bf16 operand copy (forward layout)
  __nv_bfloat16* w_dgrad;        // [C][NTAPS][KT][BN] bf16 operand copy (data-gradient layout) or nullptr
  float lr, mu;
  float* grad_out;               // optional [C][BN*NTAPS*KT] raw gradient (tests); when set no update is applied
};

template <int XROWB, int NKW, int YROWB, int NSS>
struct WgradCfg {
  static constexpr int KT = XROWB / 2;          // input channels per tap chunk
  static constexpr int BN = YROWB / 2;          // output channels
  static constexpr int NTAPS = NKW * kNKH;
  static constexpr int MTILES = (NKW == 5) ? 7 : 1;
  static constexpr int XSTAGE = kBoxH * kTileW * XROWB;
  static constexpr int YSTAGE = kTileH * kTileW * YROWB;
  static constexpr int SS_BYTES = NKW * XSTAGE + YSTAGE;
  static constexpr int TMEM_NEED = MTILES * BN;
  static constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : (TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512)));
  static constexpr int SMEM_BYTES = 1024 + NSS * SS_BYTES + (int)sizeof(Barriers);
  static_assert(SS_BYTES % 1024 == 0, "superstage must keep 1024B alignment");
};

template <int XROWB, int NKW, int YROWB, int IMG_H, int IMG_W, int BATCH, int NSS>
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                  const WgradParams p) {
  using Cfg = WgradCfg<XROWB, NKW, YROWB, NSS>;
  constexpr int KT = Cfg::KT, BN = Cfg::BN;
  constexpr int TILES_X = IMG_W / kTileW, TILES_Y = IMG_H / kTileH;
  constexpr int CHUNKS = BATCH * TILES_X * TILES_Y;   // pixel chunks (K dimension, 128 pixels each)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Barriers* bars = reinterpret_cast<Barriers*>(smem + NSS * Cfg::SS_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSS; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    mbar_init(&bars->tfull[0], 1);
    mbar_init(&bars->tempty[0], kEpiWarps);
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(&bars->tmem_base, Cfg::TMEM_COLS); tmem_relinquish(); }
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmY); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 4) {
    if (lane == 0) {
      int ss = 0; uint32_t phase = 0;
      for (int c = blockIdx.x; c < p.num_clients; c += gridDim.x) {
        if (p.active && !p.active[c]) continue;
        for (int t = 0; t < CHUNKS; ++t) {
          const int s = t / (TILES_X * TILES_Y);
          const int r = t % (TILES_X * TILES_Y);
          const int h0 = (r / TILES_X) * kTileH, w0 = (r % TILES_X) * kTileW;
          mbar_wait(&bars->empty[ss], phase ^ 1);
          mbar_arrive_expect_tx(&bars->full[ss], Cfg::SS_BYTES);
          uint8_t* base = smem + ss * Cfg::SS_BYTES;
          const int nx = c * p.x_n_stride + p.x_n_off + s;
#pragma unroll 1
          for (int kw = 0; kw < NKW; ++kw) {
            const int dw = (NKW == 1) ? 0 : (kw - kHalo);
            tma_load_4d(base + kw * Cfg::XSTAGE, &tmX, &bars->full[ss], 0, w0 + dw, h0 - kHalo, nx);
          }
          tma_load_4d(base + NKW * Cfg::XSTAGE, &tmY, &bars->full[ss], 0, w0, h0, c * BATCH + s);
          if (++ss == NSS) { ss = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t LTX = layout_type_for_row_bytes(XROWB);
      constexpr uint32_t LTY = layout_type_for_row_bytes(YROWB);
      constexpr uint32_t IDESC = make_idesc_bf16(128, BN, 1, 1);
      int ss = 0; uint32_t phase = 0; uint32_t tphase = 0;
      for (int c = blockIdx.x; c < p.num_clients; c += gridDim.x) {
        if (p.active && !p.active[c]) continue;
        mbar_wait(&bars->tempty[0], tphase ^ 1);
        fence_after_sync();
        for (int t = 0; t < CHUNKS; ++t) {
          mbar_wait(&bars->full[ss], phase);
          fence_after_sync();
          const uint32_t base = smem_u32(smem + ss * Cfg::SS_BYTES);
          const uint64_t b_desc0 = make_smem_desc(base + NKW * Cfg::XSTAGE, 16, 8 * YROWB, LTY);
#pragma unroll
          for (int mt = 0; mt < Cfg::MTILES; ++mt) {
            uint32_t a_start, lbo;
            if (mt < NKW) { a_start = base + mt * Cfg::XSTAGE; lbo = kTileW * XROWB; }
            else if (mt == NKW) { a_start = base + 4 * kTileW * XROWB; lbo = Cfg::XSTAGE; }
            else { a_start = base + 4 * Cfg::XSTAGE + 4 * kTileW * XROWB; lbo = kTileW * XROWB; }
            const uint64_t a_desc0 = make_smem_desc(a_start, lbo, 8 * XROWB, LTX);
#pragma unroll
            for (int ks = 0; ks < (kTileH * kTileW) / 16; ++ks) {
              umma_bf16(tmem_base + mt * BN, desc_advance(a_desc0, ks * 16 * XROWB),
                        desc_advance(b_desc0, ks * 16 * YROWB), IDESC, (t > 0 || ks > 0) ? 1u : 0u);
            }
          }
          umma_commit(&bars->empty[ss]);
          if (++ss == NSS) { ss = 0; phase ^= 1; }
        }
        umma_commit(&bars->tfull[0]);
        tphase ^= 1;
      }
    }
  } else {
    uint32_t tphase = 0;
    const int q = warp;
    const int L = 32 * q + lane;            // accumulator row
    const int ci = L / KT, cin = L % KT;    // chunk index inside the M tile, input channel
    for (int c = blockIdx.x; c < p.num_clients; c += gridDim.x) {
      if (p.active && !p.active[c]) continue;
      mbar_wait(&bars->tfull[0], tphase);
      fence_after_sync();
      float* wm = p.w_master + (size_t)c * p.p_stride + p.w_off;
      const float* wg = p.w_global ? p.w_global + p.w_off : nullptr;
      __nv_bfloat16* wf = p.w_fwd + (size_t)c * BN * Cfg::NTAPS * KT;
      __nv_bfloat16* wd = p.w_dgrad ? p.w_dgrad + (size_t)c * BN * Cfg::NTAPS * KT : nullptr;
      float* go = p.grad_out ? p.grad_out + (size_t)c * BN * Cfg::NTAPS * KT : nullptr;
#pragma unroll 1
      for (int mt = 0; mt < Cfg::MTILES; ++mt) {
        int kh, kw; bool valid;
        if (mt < NKW) { kh = ci; kw = mt; valid = kh < kNKH; }
        else if (mt == NKW) { kh = 4; kw = ci; valid = kw < 4; }
        else { kh = 4; kw = 4; valid = ci == 0; }
        const int tap = kh * NKW + kw;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + mt * BN;
#pragma unroll 1
        for (int j = 0; j < BN / 32; ++j) {
          uint32_t v[32];
          tmem_ld32(taddr + j * 32, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int cout = j * 32 + i;
              const size_t idx = (size_t)cout * (Cfg::NTAPS * KT) + tap * KT + cin;
              const float gacc = __uint_as_float(v[i]);
              if (go) {
                go[idx] = gacc;
              } else {
                float w = wm[idx];
                float g = gacc;
                if (wg) g += p.mu * (w - wg[idx]);
                w -= p.lr * g;
                wm[idx] = w;
                const __nv_bfloat16 wb = __float2bfloat16_rn(w);
                wf[idx] = wb;
                if (wd) wd[((size_t)tap * KT + cin) * BN + cout] = wb;
              }
            }
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tempty[0]);
      tphase ^= 1;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace convtc
