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
//   * warp roles: warps 0-3 epilogue (TMEM lanes 32*w..), warp 4 TMA producer, warp 5 MMA issuer.  The
//     producer and issuer warps run *converged* and predicate only the TMA / MMA / commit instructions on an
//     elected lane, so descriptor arithmetic stays in the uniform datapath (one add per operand per MMA).
//   * accumulators live in TMEM (double buffered) and are drained with tcgen05.ld; epilogues are fused:
//     bias+ReLU+2x2-maxpool(+argmax code) for the forward (a shuffle butterfly leaves every lane with exactly
//     the 8 channels it stores), ReLU-backward gate + bias-grad (+SGD on the bias) for the data gradient.
//   * weight gradient: both operands are MN-major (pixels are the reduction dim), all taps accumulate
//     in TMEM simultaneously (7 x [128 x 64] fp32 tiles = 448 columns for 5x5x32->64) while the pixels
//     stream through once; the epilogue applies the SGD / FedProx update to the fp32 master weights
//     and refreshes the bf16 operand copies, so weight gradients never touch HBM.  For the first layer the
//     un-pooled output gradient is not materialised either: the four otherwise idle epilogue warps synthesise
//     the swizzled dY operand tile in shared memory from the pooled gradient + argmax codes.
//
// Reference parity: this replaces the per-client `os.system("python3 <operator>")` local training
// step of the reference (ols_core/taskMgr/utils/utils_run_task.py:481-514).
#pragma once
#include "tc05.cuh"

namespace convtc {
using namespace tc05;

constexpr int kEpiWarps = 4;
constexpr int kThreads = 192;  // warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer
// folded conv2 kernels: two epilogue warp sets (warps 0-7), warp 8 TMA producer, warp 9 MMA issuer
constexpr int kFoldTmaWarp = 8, kFoldMmaWarp = 9, kFoldThreads = 320;
constexpr int kTileH = 8, kTileW = 16, kHalo = 2, kBoxH = kTileH + 2 * kHalo;  // 12
constexpr int kNKH = 5;

struct Barriers {
  uint64_t full[32];
  uint64_t empty[32];
  uint64_t wfull;
  uint64_t wempty;
  uint64_t tfull[4];
  uint64_t tempty[4];
  uint32_t tmem_base;
  uint32_t pad;
};

// ================================================================================================
// Forward / data-gradient kernel
// ================================================================================================
enum { EPI_POOL = 0, EPI_GATE = 1 };

// 2x2 max-pool + ReLU of 32 channels held by the four lanes l, l^1, l^16, l^17 of a warp (one window = 2 rows x 2
// columns of the 8x16 tile).  x = accumulator + bias.  Keys are the raw fp32 bits with (3 - window position) in the two
// LSBs, compared as SIGNED integers: exact among non-negative values, and every negative value sorts below every
// positive one — which is all ReLU needs (an all-negative window yields 0 whoever wins, and its gradient is gated
// off downstream).  One integer max therefore carries value and argmax (ties -> smallest position) at one ALU op
// per element; the sign-fixing key transform this replaces made the conv1 forward kernel ALU-bound (80% pipe).
// Lane (b0 = lane&1 column parity, b4 = row parity; the row partner is lane ^ ROW_XOR: 16 for 8x16-pixel tiles, 8 for
// 16x8 tiles) ends with channels [8*(2*b0+b4), +8): pv = bf16 values, cv = argmax codes.
template <int ROW_XOR = 16>
__device__ __forceinline__ void pool2x2_relu(const float (&x)[32], int b0, int b4, uint4& pv, uint2& cv) {
  const int tag = 3 - ((b4 << 1) | b0);
  int v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = (__float_as_int(x[i]) & ~3) | tag;
  int m16[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {                       // partner lane^1: b0=0 keeps channels [0,16), b0=1 [16,32)
    const int give = b0 ? v[i] : v[16 + i];
    const int keep = b0 ? v[16 + i] : v[i];
    m16[i] = max(keep, __shfl_xor_sync(0xffffffffu, give, 1));
  }
  int m8[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {                        // partner lane^16: b4=0 keeps the low 8, b4=1 the high 8
    const int give = b4 ? m16[i] : m16[8 + i];
    const int keep = b4 ? m16[8 + i] : m16[i];
    m8[i] = max(keep, __shfl_xor_sync(0xffffffffu, give, ROW_XOR));
  }
  float f[8];
  uint32_t cd[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    cd[i] = 3u - (static_cast<uint32_t>(m8[i]) & 3u);
    f[i] = fmaxf(__int_as_float(m8[i] & ~3), 0.f);
  }
  pv.x = pack_bf16x2(f[0], f[1]); pv.y = pack_bf16x2(f[2], f[3]);
  pv.z = pack_bf16x2(f[4], f[5]); pv.w = pack_bf16x2(f[6], f[7]);
  cv.x = cd[0] | (cd[1] << 8) | (cd[2] << 16) | (cd[3] << 24);
  cv.y = cd[4] | (cd[5] << 8) | (cd[6] << 16) | (cd[7] << 24);
}

// Sum of the five folded accumulator chunks of a pixel, each taken from the lane that holds the horizontally shifted
// pixel:  acc[i] = init[i] + sum_kw D_kw[lane + dir*(kw-2)][i]   (dir = +1 forward, -1 data gradient).
// The TMEM load of chunk n+1 is in flight while chunk n goes through its 32 shuffle+FMA pairs, so only the first
// tcgen05.ld latency is exposed instead of all five.
template <int DIR>
__device__ __forceinline__ void fold_kw_sum(uint32_t taddr, int lane, int wl, const float* init, float (&acc)[32]) {
  constexpr int ORDER[5] = {2, 0, 1, 3, 4};            // centre tap first: it needs no shuffle
  uint32_t va[32], vb[32];
  tmem_ld32(taddr + ORDER[0] * 32, va);
  tmem_ld_wait();
  tmem_ld32(taddr + ORDER[1] * 32, vb);
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = __uint_as_float(va[i]) + (init ? init[i] : 0.f);
#pragma unroll
  for (int n = 1; n < 5; ++n) {
    uint32_t (&cur)[32] = (n & 1) ? vb : va;
    uint32_t (&nxt)[32] = (n & 1) ? va : vb;
    tmem_ld_wait();
    if (n + 1 < 5) tmem_ld32(taddr + ORDER[n + 1] * 32, nxt);
    const int src_w = wl + DIR * (ORDER[n] - 2);
    const float inb = ((src_w >= 0) && (src_w < kTileW)) ? 1.f : 0.f;        // outside the image = zero padding
    const int src_lane = (lane & 16) | (src_w & 15);
#pragma unroll
    for (int i = 0; i < 32; ++i)
      acc[i] = fmaf(__uint_as_float(__shfl_sync(0xffffffffu, cur[i], src_lane)), inb, acc[i]);
  }
}

struct FwdParams {
  int num_clients;
  int nsplit;            // work items per client
  int a_n_stride;        // samples per client in the A tensor
  int a_n_off;           // first sample of this step's batch inside the client's block
  const uint8_t* active; // [C] or nullptr: 0 => client skipped this step (straggler / not sampled)
  int a_rowmap;          // A tensor map is the no-swizzle "row" map (see make_x5_row_map): x coordinate = w0 / 16
  // EPI_POOL
  const float* bias;     // [C][bias_stride] fp32
  int bias_stride;
  __nv_bfloat16* out;    // [C*BATCH][H/2][W/2][BN]
  uint8_t* code;         // same shape, argmax position 0..3
  // EPI_GATE (data gradient w.r.t. a pooled+ReLU activation)
  const __nv_bfloat16* pooled;  // [C*BATCH][H][W][BN]  forward activation (ReLU mask)
  __nv_bfloat16* dpool;         // [C*BATCH][H][W][BN]  gated gradient w.r.t. the pooled activation
  float* bias_master;           // [C][bias_stride] fp32, updated in place:  b -= lr*(db + mu*(b-bg))
  const float* bias_global;     // [bias_stride] fp32 (FedProx anchor) or nullptr
  float lr, mu;
  float* dbias_out;             // optional [C][BN]: raw bias grad (tests); when set the bias is NOT updated
};

// NMMA: MMA-issuing warps.  With few, small MMAs per tile (conv1: five N=32 MMAs) the serial per-tile work of ONE
// issuing warp (two barrier waits, fences, descriptor setup, two commits, ~700 cycles) is the bottleneck of the whole
// pipeline; NMMA = 2 gives even and odd tiles their own issuer.
template <int ROWB, int NKW, int BN, int IMG_H, int IMG_W, int BATCH, bool FLIP, int EPI, int NSTAGE, int NEPG = 1,
          int NMMA = 1>
struct FwdCfg {
  static constexpr int NBUF = 2 * NEPG;                  // TMEM accumulator buffers (two per epilogue group)
  static constexpr int THREADS = (4 * NEPG + 1 + NMMA) * 32;
  static constexpr int KT = ROWB / 2;
  static constexpr int NTAPS = NKW * kNKH;
  static constexpr int W_TAP_BYTES = BN * ROWB;
  static constexpr int W_BYTES = NTAPS * W_TAP_BYTES;
  static constexpr int STAGE_BYTES = kBoxH * kTileW * ROWB;
  static constexpr int TILES_X = IMG_W / kTileW, TILES_Y = IMG_H / kTileH;
  static constexpr int TILES_PER_CLIENT = BATCH * TILES_X * TILES_Y;
  static constexpr int TMEM_COLS = (NBUF * BN <= 32) ? 32 : (NBUF * BN <= 64 ? 64 : (NBUF * BN <= 128 ? 128 : 256));
  static constexpr int SMEM_BYTES = 1024 + W_BYTES + NSTAGE * STAGE_BYTES + (int)sizeof(Barriers);
  static_assert(EPI == EPI_POOL || NEPG == 1, "the gate epilogue reduces over one epilogue group");
  static_assert(W_TAP_BYTES % 1024 == 0 || ROWB < 128, "weight tap chunk must keep swizzle alignment");
  static_assert(BN % 32 == 0, "epilogue works on 32-column chunks");
};

template <int ROWB, int NKW, int BN, int IMG_H, int IMG_W, int BATCH, bool FLIP, int EPI, int NSTAGE, int NEPG = 1,
          int NMMA = 1>
__global__ void __launch_bounds__((4 * NEPG + 1 + NMMA) * 32, 1)
conv_taps_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const FwdParams p) {
  using Cfg = FwdCfg<ROWB, NKW, BN, IMG_H, IMG_W, BATCH, FLIP, EPI, NSTAGE, NEPG, NMMA>;
  constexpr int kProdWarp = 4 * NEPG, kMmaWarp = 4 * NEPG + 1, NBUF = Cfg::NBUF;
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
    mbar_init(&bars->wempty, NMMA);
    for (int i = 0; i < NBUF; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) { tmem_alloc(&bars->tmem_base, Cfg::TMEM_COLS); tmem_relinquish(); }
  if (warp == kProdWarp && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  const int num_work = p.num_clients * p.nsplit;
  const int tiles_per_work = (Cfg::TILES_PER_CLIENT + p.nsplit - 1) / p.nsplit;

  if (warp == kProdWarp) {
    // ------------------------------------------------------------------ TMA producer (warp-converged)
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0; int cur_client = -1; uint32_t wphase = 0; bool first_w = true;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / p.nsplit, sp = wi % p.nsplit;
      if (p.active && !p.active[c]) continue;
      if (c != cur_client) {
        if (!first_w) { mbar_wait(&bars->wempty, wphase); wphase ^= 1; }
        first_w = false; cur_client = c;
        if (leader) {
          mbar_arrive_expect_tx(&bars->wfull, Cfg::W_BYTES);
#pragma unroll 1
          for (int t = 0; t < Cfg::NTAPS; ++t)
            // forward copy  [C][BN][NTAPS*KT]  -> box (KT, BN, 1) at (t*KT, 0, c)
            // dgrad copy    [C][NTAPS*BN][K]   -> box (K,  BN, 1) at (0, t*BN, c)
            tma_load_3d(w_sm + t * Cfg::W_TAP_BYTES, &tmW, &bars->wfull, FLIP ? 0 : t * Cfg::KT,
                        FLIP ? t * BN : 0, c);
        }
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
          if (leader) {
            mbar_arrive_expect_tx(&bars->full[stage], Cfg::STAGE_BYTES);
            const int dw = (NKW == 1) ? 0 : (FLIP ? (kHalo - kw) : (kw - kHalo));
            tma_load_4d(a_sm + stage * Cfg::STAGE_BYTES, &tmA, &bars->full[stage], 0,
                        p.a_rowmap ? (w0 / kTileW) : (w0 + dw), h0 - kHalo, n);
          }
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= kMmaWarp && warp < kMmaWarp + NMMA) {
    // ------------------------------------------------------------------ MMA issuer(s) (warp-converged); issuer m
    // owns the tiles with seq % NMMA == m and walks the stage / buffer rings in lock step with the others
    const int mpar = warp - kMmaWarp;
    const bool leader = elect_one();
    constexpr uint32_t LT = layout_type_for_row_bytes(ROWB);
    constexpr uint32_t IDESC = make_idesc_bf16(128, BN, 0, 0);
    const uint64_t a_desc_base = make_smem_desc(smem_u32(a_sm), 16, 8 * ROWB, LT);
    const uint64_t b_desc_base = make_smem_desc(smem_u32(w_sm), 16, 8 * ROWB, LT);
    int stage = 0; uint32_t phase = 0; int cur_client = -1; uint32_t wphase = 0;
    uint32_t seq = 0;                                    // tiles issued by this CTA: buffer = seq % NBUF
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / p.nsplit, sp = wi % p.nsplit;
      if (p.active && !p.active[c]) continue;
      bool last_of_client = true;
      {
        const int nwi = wi + gridDim.x;
        if (nwi < num_work && (nwi / p.nsplit) == c) last_of_client = false;
      }
      if (c != cur_client) { mbar_wait(&bars->wfull, wphase); wphase ^= 1; cur_client = c; }
      const int t0 = sp * tiles_per_work;
      const int t1 = min(t0 + tiles_per_work, Cfg::TILES_PER_CLIENT);
      for (int t = t0; t < t1; ++t) {
        if (NMMA > 1 && (int)(seq % NMMA) != mpar) {            // another issuer's tile: just advance the rings
#pragma unroll 1
          for (int kw = 0; kw < NKW; ++kw) { if (++stage == NSTAGE) { stage = 0; phase ^= 1; } }
          ++seq;
          continue;
        }
        const int buf = seq % NBUF;
        mbar_wait(&bars->tempty[buf], ((seq / NBUF) & 1) ^ 1);
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + buf * BN;
        uint32_t acc = 0;
#pragma unroll 1
        for (int kw = 0; kw < NKW; ++kw) {
          mbar_wait(&bars->full[stage], phase);
          fence_after_sync();
          if (leader) {
            const uint64_t a_stage = desc_advance(a_desc_base, stage * Cfg::STAGE_BYTES);
            const uint64_t b_kw = desc_advance(b_desc_base, kw * Cfg::W_TAP_BYTES);
#pragma unroll
            for (int kh = 0; kh < kNKH; ++kh) {
              constexpr int dummy = 0; (void)dummy;
              const int row_shift = FLIP ? (2 * kHalo - kh) : kh;          // image rows above the tile start
              const uint64_t a_desc = desc_advance(a_stage, row_shift * kTileW * ROWB);
              const uint64_t b_desc = desc_advance(b_kw, kh * NKW * Cfg::W_TAP_BYTES);
#pragma unroll
              for (int ks = 0; ks < ROWB / 32; ++ks) {
                umma_bf16(d_tmem, desc_advance(a_desc, ks * 32), desc_advance(b_desc, ks * 32), IDESC, acc);
                acc = 1;
              }
            }
            umma_commit(&bars->empty[stage]);
          }
          __syncwarp();
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        if (leader) umma_commit(&bars->tfull[buf]);
        __syncwarp();
        ++seq;
      }
      if (last_of_client) {
        if (leader) umma_commit(&bars->wempty);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (TMEM lanes 32*warp ..)
    uint32_t seq = 0;
    const int q = warp & 3;            // TMEM lane quarter
    const int grp = warp >> 2;         // epilogue group: handles tiles with seq % NEPG == grp
    const int hl = 2 * q + (lane >> 4), wl = lane & 15;
    constexpr int NDB = (EPI == EPI_GATE) ? BN : 1;
    float dbacc[NDB];
#pragma unroll
    for (int i = 0; i < NDB; ++i) dbacc[i] = 0.f;
    int cur_client = -1;
    bool had_sp0 = false;
    __shared__ float s_db[NDB];
    // this lane's output slice after the pooling butterfly: channels [j*32 + sl*8, +8)
    const int b0 = lane & 1, b4 = (lane >> 4) & 1;
    const int sl = b0 * 2 + b4;
    float bias_r[(EPI == EPI_POOL) ? BN : 1];

    auto flush_bias = [&](int c) {
      if (EPI != EPI_GATE) return;
      asm volatile("bar.sync 1, 128;");
      if (threadIdx.x < NDB) s_db[threadIdx.x] = 0.f;
      asm volatile("bar.sync 1, 128;");
#pragma unroll
      for (int i = 0; i < NDB; ++i) {
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
      if (threadIdx.x < NDB) {
        const float g = s_db[threadIdx.x];
        if (p.dbias_out) {
          atomicAdd(&p.dbias_out[(size_t)c * BN + threadIdx.x], g);
        } else {
          // several CTAs may share a client (nsplit>1): the update is linear in g, so apply the partial
          // gradient atomically; the FedProx pull is applied once, by the CTA that owns split 0.
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
        if (EPI == EPI_POOL) {
          const float* bias = p.bias + (size_t)c * p.bias_stride;
#pragma unroll
          for (int k = 0; k < BN; ++k) bias_r[k] = __ldg(&bias[k]);
        }
      }
      if (sp == 0) had_sp0 = true;
      const int t0 = sp * tiles_per_work;
      const int t1 = min(t0 + tiles_per_work, Cfg::TILES_PER_CLIENT);
      for (int t = t0; t < t1; ++t) {
        const int s = t / (Cfg::TILES_X * Cfg::TILES_Y);
        const int r = t % (Cfg::TILES_X * Cfg::TILES_Y);
        const int h0 = (r / Cfg::TILES_X) * kTileH, w0 = (r % Cfg::TILES_X) * kTileW;
        const size_t n_out = (size_t)c * BATCH + s;
        const uint32_t myseq = seq++;
        if ((int)(myseq % NEPG) != grp) continue;
        const int buf = myseq % NBUF;
        mbar_wait(&bars->tfull[buf], (myseq / NBUF) & 1);
        fence_after_sync();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + buf * BN;

        if (EPI == EPI_POOL) {
          // relu(max_i(acc_i) + b) == relu(max_i(acc_i + b)): the bias goes in first (FMA pipe) so the pooling keys
          // can be compared as signed integers (pool2x2_relu)
          const int ph = (h0 >> 1) + q, pw = (w0 >> 1) + (wl >> 1);
          const size_t obase = ((n_out * (IMG_H / 2) + ph) * (IMG_W / 2) + pw) * BN;
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) {
            uint32_t v[32];
            tmem_ld32(taddr + j * 32, v);
            tmem_ld_wait();
            if (j == BN / 32 - 1) {
              // accumulator is in registers: hand the TMEM buffer back before the pooling arithmetic and the stores
              // (released at the end of the tile, the MMA warp waited ~1/4 of its time for a free buffer)
              fence_before_sync();
              __syncwarp();
              if (lane == 0) mbar_arrive(&bars->tempty[buf]);
            }
            float x[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(v[i]) + bias_r[j * 32 + i];
            uint4 pv;
            uint2 cv;
            pool2x2_relu(x, b0, b4, pv, cv);
            const size_t off = obase + j * 32 + sl * 8;
            *reinterpret_cast<uint4*>(p.out + off) = pv;
            *reinterpret_cast<uint2*>(p.code + off) = cv;
          }
        } else {
          // EPI_GATE: this lane owns pixel (h0+hl, w0+wl) of sample n_out; BN == 32 channels.
          static_assert(EPI != EPI_GATE || BN == 32, "gate epilogue is written for 32 channels");
          const int h = h0 + hl, w = w0 + wl;
          const size_t pbase = ((n_out * IMG_H + h) * IMG_W + w) * BN;
          uint32_t v[32];
          tmem_ld32(taddr, v);
          uint4 pq[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) pq[k] = *reinterpret_cast<const uint4*>(p.pooled + pbase + k * 8);
          tmem_ld_wait();
          const uint32_t* pw32 = reinterpret_cast<const uint32_t*>(pq);
          uint4 o[4];
          uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t pr = pw32[i];
            const float g0 = bf16lo(pr) > 0.f ? __uint_as_float(v[2 * i]) : 0.f;
            const float g1 = bf16hi(pr) > 0.f ? __uint_as_float(v[2 * i + 1]) : 0.f;
            dbacc[2 * i] += g0; dbacc[2 * i + 1] += g1;
            ow[i] = pack_bf16x2(g0, g1);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(p.dpool + pbase + k * 8) = o[k];
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->tempty[buf]);
        }
      }
    }
    if (cur_client >= 0) flush_bias(cur_client);
  }

  fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ================================================================================================
// conv1 forward, pooling done by the accumulator layout ("window" form)
// ================================================================================================
// conv_taps_kernel computes one accumulator row per INPUT pixel and pools across lanes: 150 ALU-pipe instructions
// per lane (tag / select / shuffle / max), which made it ALU-bound (74-80 % of the pipe at 11-14 % tensor).  Here an
// accumulator row is a POOLED pixel and its four window positions are four 32-column accumulator groups, so pooling
// is a max over four registers of the same thread: no shuffles, no selects.
//
// That needs the A operand of window position (dy, dx) to be "every second pixel of every second row".  The input is
// therefore stored as two column-parity planes per image row ([n][h][parity][16 pixels][16 ch], written by the ingest
// kernel, already 32-byte-swizzled): a plane row holds the pixels 2j+dx contiguously, tiles are 16 pooled rows x 8
// pooled columns (one 8-pixel group per pooled row -> uniform group stride of two box rows), and (dy, dx, kh) are all
// start-address shifts of one K-major view of ONE 36-row x 2-plane box per image (36 TMA rows of 512 B).
template <int BATCH, int NSTAGE>
struct Conv1WinCfg {
  static constexpr int COUT = 32, KT = 16, ROWB = 32;
  static constexpr int IMG = 32, POOLED = 16;
  static constexpr int BOX_ROWS = IMG + 2 * kHalo;            // whole image + halo rows
  static constexpr int ROW_PITCH = 2 * 16 * ROWB;             // two parity planes of 16 pixels
  static constexpr int STAGE_BYTES = BOX_ROWS * ROW_PITCH;    // 36864
  static constexpr int W_KH_BYTES = COUT * ROWB;              // [32 cout][16 k] per vertical tap
  static constexpr int W_BYTES = kNKH * W_KH_BYTES;
  static constexpr int NBUF = 4;                              // accumulator buffers: 2 tiles x 2 images in flight
  static constexpr int BUF_COLS = 4 * COUT;                   // four window positions
  static constexpr int TMEM_COLS = 512;
  static constexpr int THREADS = (8 + 1 + 2) * 32;            // 8 epilogue warps, TMA warp, two MMA issuers
  static constexpr int SMEM_BYTES = 1024 + 1024 + W_BYTES + NSTAGE * STAGE_BYTES + (int)sizeof(Barriers);
  static_assert(STAGE_BYTES % 1024 == 0, "stage keeps the swizzle phase");
};

template <int BATCH, int NSTAGE>
__global__ void __launch_bounds__(Conv1WinCfg<BATCH, NSTAGE>::THREADS, 1)
conv1_window_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const FwdParams p) {
  using Cfg = Conv1WinCfg<BATCH, NSTAGE>;
  constexpr int kTma = 8, kMma0 = 9;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_sm = smem;                                         // 5 KB of weights, stage area starts 1 KB-aligned
  uint8_t* a_sm = smem + ((Cfg::W_BYTES + 1023) / 1024) * 1024;
  Barriers* bars = reinterpret_cast<Barriers*>(a_sm + NSTAGE * Cfg::STAGE_BYTES);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 2); }
    mbar_init(&bars->wfull, 1);
    mbar_init(&bars->wempty, 2);
    for (int i = 0; i < Cfg::NBUF; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == kMma0) { tmem_alloc(&bars->tmem_base, Cfg::TMEM_COLS); tmem_relinquish(); }
  if (warp == kTma && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const int num_work = p.num_clients * p.nsplit;
  const int imgs_per_work = BATCH / p.nsplit;

  if (warp == kTma) {
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0; int cur_client = -1; uint32_t wphase = 0; bool first_w = true;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / p.nsplit, sp = wi % p.nsplit;
      if (p.active && !p.active[c]) continue;
      if (c != cur_client) {
        if (!first_w) { mbar_wait(&bars->wempty, wphase); wphase ^= 1; }
        first_w = false; cur_client = c;
        if (leader) {
          mbar_arrive_expect_tx(&bars->wfull, Cfg::W_BYTES);
#pragma unroll 1
          for (int kh = 0; kh < kNKH; ++kh)
            tma_load_3d(w_sm + kh * Cfg::W_KH_BYTES, &tmW, &bars->wfull, kh * Cfg::KT, 0, c);
        }
      }
      for (int i = 0; i < imgs_per_work; ++i) {
        const int n = c * p.a_n_stride + p.a_n_off + sp * imgs_per_work + i;
        mbar_wait(&bars->empty[stage], phase ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&bars->full[stage], Cfg::STAGE_BYTES);
          tma_load_4d(a_sm + stage * Cfg::STAGE_BYTES, &tmA, &bars->full[stage], 0, 0, -kHalo, n);
        }
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= kMma0) {
    // issuer g owns tile g (left / right half) of every image
    const int g = warp - kMma0;
    const bool leader = elect_one();
    constexpr uint32_t IDESC = make_idesc_bf16(128, Cfg::COUT, 0, 0);
    // A: sixteen 8-pixel groups (one per pooled row) two box rows apart; B: [32 cout][16 k] per vertical tap
    const uint64_t a_desc_base = make_smem_desc(smem_u32(a_sm) + g * 8 * Cfg::ROWB, 16, 2 * Cfg::ROW_PITCH, LT_SW32);
    const uint64_t b_desc_base = make_smem_desc(smem_u32(w_sm), 16, 8 * Cfg::ROWB, LT_SW32);
    int stage = 0; uint32_t phase = 0; int cur_client = -1; uint32_t wphase = 0;
    uint32_t iseq = 0;                                  // images issued: buffer = (iseq % 2) * 2 + g
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / p.nsplit;
      if (p.active && !p.active[c]) continue;
      bool last_of_client = true;
      {
        const int nwi = wi + gridDim.x;
        if (nwi < num_work && (nwi / p.nsplit) == c) last_of_client = false;
      }
      if (c != cur_client) { mbar_wait(&bars->wfull, wphase); wphase ^= 1; cur_client = c; }
      for (int i = 0; i < imgs_per_work; ++i, ++iseq) {
        const int buf = (iseq & 1) * 2 + g;
        mbar_wait(&bars->tempty[buf], ((iseq >> 1) & 1) ^ 1);
        fence_after_sync();
        mbar_wait(&bars->full[stage], phase);
        fence_after_sync();
        if (leader) {
          const uint32_t d0 = tmem_base + buf * Cfg::BUF_COLS;
          const uint64_t a_stage = desc_advance(a_desc_base, stage * Cfg::STAGE_BYTES);
#pragma unroll
          for (int pos = 0; pos < 4; ++pos) {
#pragma unroll
            for (int kh = 0; kh < kNKH; ++kh) {
              // window position (dy, dx) = (pos >> 1, pos & 1): box row 2*ph + dy + kh, parity plane dx
              const uint32_t aoff = ((pos >> 1) + kh) * Cfg::ROW_PITCH + (pos & 1) * (Cfg::ROW_PITCH / 2);
              umma_bf16(d0 + pos * Cfg::COUT, desc_advance(a_stage, aoff),
                        desc_advance(b_desc_base, kh * Cfg::W_KH_BYTES), IDESC, kh ? 1u : 0u);
            }
          }
          umma_commit(&bars->empty[stage]);
          umma_commit(&bars->tfull[buf]);
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
      if (last_of_client) {
        if (leader) umma_commit(&bars->wempty);
        __syncwarp();
      }
    }
  } else {
    // epilogue warp set g (warps 4g..4g+3): tile g of every image; lane = pooled pixel (ph = 4q + lane/8, pw = 8g + lane%8)
    const int q = warp & 3, g = warp >> 2;
    const int ph = 4 * q + (lane >> 3), pw = 8 * g + (lane & 7);
    float bias_r[Cfg::COUT];
    int cur_client = -1;
    uint32_t iseq = 0;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / p.nsplit, sp = wi % p.nsplit;
      if (p.active && !p.active[c]) continue;
      if (c != cur_client) {
        cur_client = c;
        const float* bias = p.bias + (size_t)c * p.bias_stride;
#pragma unroll
        for (int k = 0; k < Cfg::COUT; ++k) bias_r[k] = __ldg(&bias[k]);
      }
      for (int i = 0; i < imgs_per_work; ++i, ++iseq) {
        const int buf = (iseq & 1) * 2 + g;
        const size_t n_out = (size_t)c * BATCH + sp * imgs_per_work + i;
        const size_t obase = ((n_out * Cfg::POOLED + ph) * Cfg::POOLED + pw) * Cfg::COUT;
        mbar_wait(&bars->tfull[buf], (iseq >> 1) & 1);
        fence_after_sync();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + buf * Cfg::BUF_COLS;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[4][16];
#pragma unroll
          for (int pos = 0; pos < 4; ++pos) tmem_ld16(taddr + pos * Cfg::COUT + half * 16, v[pos]);
          tmem_ld_wait();
          if (half == 1) {                                   // accumulator fully in registers: release the buffer
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tempty[buf]);
          }
          uint32_t ob[8], oc[4];
#pragma unroll
          for (int i2 = 0; i2 < 16; i2 += 2) {
            float f[2];
            uint32_t cd[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float b = bias_r[half * 16 + i2 + e];
              // raw fp32 bits with (3 - position) in the two LSBs, signed max: see pool2x2_relu
              int m = (__float_as_int(__uint_as_float(v[0][i2 + e]) + b) & ~3) | 3;
              m = max(m, (__float_as_int(__uint_as_float(v[1][i2 + e]) + b) & ~3) | 2);
              m = max(m, (__float_as_int(__uint_as_float(v[2][i2 + e]) + b) & ~3) | 1);
              m = max(m, (__float_as_int(__uint_as_float(v[3][i2 + e]) + b) & ~3));
              cd[e] = 3u - (static_cast<uint32_t>(m) & 3u);
              f[e] = fmaxf(__int_as_float(m & ~3), 0.f);
            }
            ob[i2 >> 1] = pack_bf16x2(f[0], f[1]);
            const int sh = (i2 & 3) * 8;
            if ((i2 & 3) == 0) oc[i2 >> 2] = 0;
            oc[i2 >> 2] |= (cd[0] << sh) | (cd[1] << (sh + 8));
          }
          *reinterpret_cast<uint4*>(p.out + obase + half * 16) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
          *reinterpret_cast<uint4*>(p.out + obase + half * 16 + 8) = make_uint4(ob[4], ob[5], ob[6], ob[7]);
          *reinterpret_cast<uint4*>(p.code + obase + half * 16) = make_uint4(oc[0], oc[1], oc[2], oc[3]);
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == kMma0) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ================================================================================================
// Data-gradient kernel, horizontal taps folded into N
// ================================================================================================
// dX[h,w,ci] = sum_{kh,kw} dY[h-(kh-2), w-(kw-2), co] * W[kh,kw][co,ci].   A shift of the *input* column is the same
// as a shift of the *output* column, so for every vertical tap kh ONE MMA with N = 5*CIN computes the five partial
// products D_kw[h,w'] = sum_co dY[h-(kh-2), w', co] W[kh,kw][co,:] from the same un-shifted A tile, and the epilogue
// sums D_kw taken from the lane of pixel w+2-kw (the 8x16 tile spans the full image width, so out-of-tile = zero
// padding).  Versus one MMA per (kh,kw): 5x fewer TMA boxes and MMAs, 2.8x less shared-memory operand traffic —
// the N=32 variant is capped at 40% tensor utilisation by the 128 B/clk smem read port.
template <int CIN, int COUT, int IMG_H, int BATCH, int NSTAGE>
struct DgradFoldCfg {
  static constexpr int ROWB = COUT * 2;                 // bytes per A pixel row (K = COUT)
  static constexpr int NF = 5 * CIN;                    // folded N
  static constexpr int W_BOX_BYTES = COUT * CIN * 2;    // one (kh,kw) tap as loaded from the forward copy: [COUT rows][CIN]
  static constexpr int W_KH_BYTES = 5 * W_BOX_BYTES;   // one vertical tap = 5 MN-major chunks (kw) of the B operand
  static constexpr int W_BYTES = kNKH * W_KH_BYTES;
  static constexpr int STAGE_BYTES = kBoxH * kTileW * ROWB;
  static constexpr int TILES_PER_CLIENT = BATCH * (IMG_H / kTileH);
  static constexpr int TMEM_BUF_STRIDE = 256;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = 1024 + W_BYTES + NSTAGE * STAGE_BYTES + (int)sizeof(Barriers);
  static_assert(ROWB == 128 && CIN == 32, "written for 64 -> 32 channels (SW128 rows, one 32-column chunk per tap)");
  static_assert(NF % 16 == 0 && NF <= 256, "UMMA N constraint");
};

template <int CIN, int COUT, int IMG_H, int BATCH, int NSTAGE>
__global__ void __launch_bounds__(kFoldThreads, 1)
conv_dgrad_fold_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                       const FwdParams p) {
  using Cfg = DgradFoldCfg<CIN, COUT, IMG_H, BATCH, NSTAGE>;
  constexpr int ROWB = Cfg::ROWB;
  constexpr int IMG_W = kTileW;
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
  if (warp == kFoldMmaWarp) { tmem_alloc(&bars->tmem_base, Cfg::TMEM_COLS); tmem_relinquish(); }
  if (warp == kFoldTmaWarp && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const int num_work = p.num_clients * p.nsplit;
  const int tiles_per_work = (Cfg::TILES_PER_CLIENT + p.nsplit - 1) / p.nsplit;

  if (warp == kFoldTmaWarp) {
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0; int cur_client = -1; uint32_t wphase = 0; bool first_w = true;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / p.nsplit, sp = wi % p.nsplit;
      if (p.active && !p.active[c]) continue;
      if (c != cur_client) {
        if (!first_w) { mbar_wait(&bars->wempty, wphase); wphase ^= 1; }
        first_w = false; cur_client = c;
        if (leader) {
          mbar_arrive_expect_tx(&bars->wfull, Cfg::W_BYTES);
          // B operand straight from the *forward* weight copy [C][COUT][25*CIN]: box (CIN, COUT, 1) at (tap*CIN, 0, c)
          // lands as [COUT k-rows][CIN n-elements] = one MN-major SW64 chunk per tap -> no transposed weight copy
#pragma unroll 1
          for (int t = 0; t < 5 * kNKH; ++t)
            tma_load_3d(w_sm + t * Cfg::W_BOX_BYTES, &tmW, &bars->wfull, t * CIN, 0, c);
        }
      }
      const int t0 = sp * tiles_per_work;
      const int t1 = min(t0 + tiles_per_work, Cfg::TILES_PER_CLIENT);
      for (int t = t0; t < t1; ++t) {
        const int s = t / (IMG_H / kTileH);
        const int h0 = (t % (IMG_H / kTileH)) * kTileH;
        mbar_wait(&bars->empty[stage], phase ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&bars->full[stage], Cfg::STAGE_BYTES);
          tma_load_4d(a_sm + stage * Cfg::STAGE_BYTES, &tmA, &bars->full[stage], 0, 0, h0 - kHalo,
                      c * p.a_n_stride + p.a_n_off + s);
        }
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kFoldMmaWarp) {
    const bool leader = elect_one();
    constexpr uint32_t IDESC = make_idesc_bf16(128, Cfg::NF, 0, 1);           // B is MN-major
    const uint64_t a_desc_base = make_smem_desc(smem_u32(a_sm), 16, 8 * ROWB, LT_SW128);
    // B: 5 chunks (kw) of [COUT k-rows][CIN] with 64-byte rows: LBO = chunk stride, SBO = 8 k-rows
    const uint64_t b_desc_base = make_smem_desc(smem_u32(w_sm), Cfg::W_BOX_BYTES, 8 * CIN * 2, LT_SW64);
    int stage = 0; uint32_t phase = 0; int cur_client = -1; uint32_t wphase = 0;
    int buf = 0; uint32_t tphase0 = 0, tphase1 = 0;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / p.nsplit, sp = wi % p.nsplit;
      if (p.active && !p.active[c]) continue;
      bool last_of_client = true;
      {
        const int nwi = wi + gridDim.x;
        if (nwi < num_work && (nwi / p.nsplit) == c) last_of_client = false;
      }
      if (c != cur_client) { mbar_wait(&bars->wfull, wphase); wphase ^= 1; cur_client = c; }
      const int t0 = sp * tiles_per_work;
      const int t1 = min(t0 + tiles_per_work, Cfg::TILES_PER_CLIENT);
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&bars->tempty[buf], (buf ? tphase1 : tphase0) ^ 1);
        fence_after_sync();
        mbar_wait(&bars->full[stage], phase);
        fence_after_sync();
        if (leader) {
          const uint32_t d_tmem = tmem_base + buf * Cfg::TMEM_BUF_STRIDE;
          const uint64_t a_stage = desc_advance(a_desc_base, stage * Cfg::STAGE_BYTES);
#pragma unroll
          for (int kh = 0; kh < kNKH; ++kh) {
            const uint64_t a_desc = desc_advance(a_stage, (2 * kHalo - kh) * kTileW * ROWB);
            const uint64_t b_desc = desc_advance(b_desc_base, kh * Cfg::W_KH_BYTES);
#pragma unroll
            for (int ks = 0; ks < ROWB / 32; ++ks)
              umma_bf16(d_tmem, desc_advance(a_desc, ks * 32), desc_advance(b_desc, ks * 16 * CIN * 2), IDESC,
                        (kh | ks) ? 1u : 0u);
          }
          umma_commit(&bars->empty[stage]);
          umma_commit(&bars->tfull[buf]);
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        if (buf) tphase1 ^= 1; else tphase0 ^= 1;
        buf ^= 1;
      }
      if (last_of_client) {
        if (leader) umma_commit(&bars->wempty);
        __syncwarp();
      }
    }
  } else {
    // two epilogue groups: group g drains TMEM buffer g (= every other tile), so one group's TMEM-load / shuffle
    // latency is covered by the other group's arithmetic on the same scheduler
    const int grp = warp >> 2;
    int buf = 0; uint32_t tph = 0;
    const int q = warp & 3;
    const int tid_g = threadIdx.x & 127;
    const int hl = 2 * q + (lane >> 4), wl = lane & 15;
    float dbacc[CIN];
#pragma unroll
    for (int i = 0; i < CIN; ++i) dbacc[i] = 0.f;
    int cur_client = -1;
    bool had_sp0 = false;
    __shared__ float s_db_all[2][CIN];
    float* s_db = s_db_all[grp];
    auto gsync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp)); };
    auto flush_bias = [&](int c) {
      gsync();
      if (tid_g < CIN) s_db[tid_g] = 0.f;
      gsync();
#pragma unroll
      for (int i = 0; i < CIN; ++i) {
        float v = dbacc[i];
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        if (lane == 0) atomicAdd(&s_db[i], v);
        dbacc[i] = 0.f;
      }
      gsync();
      if (tid_g < CIN) {
        const float g = s_db[tid_g];
        if (p.dbias_out) {
          atomicAdd(&p.dbias_out[(size_t)c * CIN + tid_g], g);
        } else {
          float* bp = &p.bias_master[(size_t)c * p.bias_stride + tid_g];
          float upd = -p.lr * g;
          if (grp == 0 && had_sp0 && p.bias_global) upd -= p.lr * p.mu * (*bp - p.bias_global[tid_g]);
          atomicAdd(bp, upd);
        }
      }
      gsync();
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
      for (int t = t0; t < t1; ++t, buf ^= 1) {
        if (buf != grp) continue;
        const int s = t / (IMG_H / kTileH);
        const int h0 = (t % (IMG_H / kTileH)) * kTileH;
        const size_t n_out = (size_t)c * BATCH + s;
        const size_t pbase = ((n_out * IMG_H + (h0 + hl)) * IMG_W + wl) * CIN;
        // forward activation of this pixel (ReLU gate) — issued before the TMEM wait to overlap its latency
        uint4 pq[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) pq[k] = __ldg(reinterpret_cast<const uint4*>(p.pooled + pbase + k * 8));
        mbar_wait(&bars->tfull[buf], tph);
        fence_after_sync();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + buf * Cfg::TMEM_BUF_STRIDE;
        float acc[CIN];
        fold_kw_sum<-1>(taddr, lane, wl, nullptr, acc);       // D_kw is taken from the pixel 2-kw columns to the right
        // TMEM buffer drained: release it before the (long-latency) global stores
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tempty[buf]);
        const uint32_t* pw32 = reinterpret_cast<const uint32_t*>(pq);
        uint4 o[4];
        uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
        for (int i = 0; i < CIN / 2; ++i) {
          const uint32_t pr = pw32[i];
          const float g0 = bf16lo(pr) > 0.f ? acc[2 * i] : 0.f;
          const float g1 = bf16hi(pr) > 0.f ? acc[2 * i + 1] : 0.f;
          dbacc[2 * i] += g0; dbacc[2 * i + 1] += g1;
          ow[i] = pack_bf16x2(g0, g1);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(p.dpool + pbase + k * 8) = o[k];
        tph ^= 1;
      }
    }
    if (cur_client >= 0) flush_bias(cur_client);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == kFoldMmaWarp) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ================================================================================================
// Data-gradient kernel, CTA-pair version (cta_group::2)
// ================================================================================================
// Same math and epilogue as conv_dgrad_fold_kernel; two CTAs on the SMs of one TPC work on two tiles of the same
// client at a time as ONE M=256 MMA.  The folded weight operand (N = 160) is split between the two CTAs (each holds
// 80 columns = five 16-channel MN-major chunks per vertical tap), so per tile the tensor core fetches 4 KB of A and
// 2.5 KB of B from its SM's shared memory instead of 4 + 5 KB — the folded kernels are bound by exactly that pipe.
//   warps 0-7  epilogue (two sets, alternate steps)      warp 8  TMA producer (own tile + own half of the weights)
//   warp 9     leader: MMA issuer for the pair           peer: relay that tells the leader "my operands have landed"
struct PairBarriers {
  uint64_t full[8];      // A stage landed in THIS CTA (tx count)
  uint64_t pfull[8];     // leader only: the peer's A stage landed (remote arrive by the peer's relay warp)
  uint64_t empty[8];     // stage consumed (multicast commit)
  uint64_t wfull, pwfull, wempty;
  uint64_t tfull[3];     // accumulator ready (multicast commit)
  uint64_t tempty[3];    // leader only: accumulator drained by both CTAs (4 local + 4 remote warps)
  uint32_t tmem_base;
  uint32_t pad;
};

template <int CIN, int COUT, int IMG_H, int BATCH, int NSTAGE>
struct DgradPairCfg {
  static constexpr int ROWB = COUT * 2;                 // bytes per A pixel row (K = COUT)
  static constexpr int NF = 5 * CIN;                    // folded N of the pair's MMA
  static constexpr int CHUNK_BYTES = COUT * 16 * 2;     // one MN-major SW32 chunk: [COUT k-rows][16 input channels]
  static constexpr int CHUNKS_PER_CTA = NF / 16 / 2;    // 5 chunks of 16 columns per vertical tap and CTA
  static constexpr int W_KH_BYTES = CHUNKS_PER_CTA * CHUNK_BYTES;
  static constexpr int W_BYTES = kNKH * W_KH_BYTES;
  static constexpr int STAGE_BYTES = kBoxH * kTileW * ROWB;
  static constexpr int TILES_PER_CLIENT = BATCH * (IMG_H / kTileH);
  // three accumulator buffers: draining one tile (five TMEM loads, 128 shuffles through the shared-memory pipe) takes
  // about two MMA steps, so with two buffers the tensor core idled a third of the time waiting for a free one
  static constexpr int NBUF = 3;
  static constexpr int TMEM_BUF_STRIDE = 160;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = 1024 + W_BYTES + NSTAGE * STAGE_BYTES + (int)sizeof(PairBarriers);
  static_assert(ROWB == 128 && CIN == 32 && NSTAGE <= 8, "written for 64 -> 32 channels");
};

template <int CIN, int COUT, int IMG_H, int BATCH, int NSTAGE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFoldThreads, 1)
conv_dgrad_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                       const FwdParams p) {
  using Cfg = DgradPairCfg<CIN, COUT, IMG_H, BATCH, NSTAGE>;
  constexpr int ROWB = Cfg::ROWB;
  constexpr int IMG_W = kTileW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_sm = smem;
  uint8_t* a_sm = smem + Cfg::W_BYTES;
  PairBarriers* bars = reinterpret_cast<PairBarriers*>(a_sm + NSTAGE * Cfg::STAGE_BYTES);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->pfull[i], 1); mbar_init(&bars->empty[i], 1); }
    mbar_init(&bars->wfull, 1); mbar_init(&bars->pwfull, 1); mbar_init(&bars->wempty, 1);
    for (int i = 0; i < Cfg::NBUF; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], 2 * kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == kFoldMmaWarp) { tmem_alloc2(&bars->tmem_base, Cfg::TMEM_COLS); tmem_relinquish2(); }
  if (warp == kFoldTmaWarp && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW); }
  fence_before_sync();
  cluster_sync();                       // barriers of both CTAs initialised before any remote arrive / multicast commit
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const int num_work = p.num_clients * p.nsplit;
  const int tiles_per_work = Cfg::TILES_PER_CLIENT / p.nsplit;     // even by construction (host picks nsplit)
  const int steps_per_work = tiles_per_work / 2;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == kFoldTmaWarp) {
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0; int cur_client = -1; uint32_t wphase = 0; bool first_w = true;
    for (int wi = pair; wi < num_work; wi += npairs) {
      const int c = wi / p.nsplit, sp = wi % p.nsplit;
      if (p.active && !p.active[c]) continue;
      if (c != cur_client) {
        if (!first_w) { mbar_wait(&bars->wempty, wphase); wphase ^= 1; }
        first_w = false; cur_client = c;
        if (leader) {
          mbar_arrive_expect_tx(&bars->wfull, Cfg::W_BYTES);
          // this CTA's half of the folded weight operand: global 16-column chunk g = 5*rank + j -> (kw = g/2, half)
#pragma unroll 1
          for (int kh = 0; kh < kNKH; ++kh)
#pragma unroll 1
            for (int j = 0; j < Cfg::CHUNKS_PER_CTA; ++j) {
              const int g = Cfg::CHUNKS_PER_CTA * (int)rank + j;
              tma_load_3d(w_sm + kh * Cfg::W_KH_BYTES + j * Cfg::CHUNK_BYTES, &tmW, &bars->wfull,
                          (kh * 5 + (g >> 1)) * CIN + (g & 1) * 16, 0, c);
            }
        }
      }
      const int t0 = sp * tiles_per_work;
      for (int k = 0; k < steps_per_work; ++k) {
        const int t = t0 + 2 * k + (int)rank;
        const int s = t / (IMG_H / kTileH);
        const int h0 = (t % (IMG_H / kTileH)) * kTileH;
        mbar_wait(&bars->empty[stage], phase ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&bars->full[stage], Cfg::STAGE_BYTES);
          tma_load_4d(a_sm + stage * Cfg::STAGE_BYTES, &tmA, &bars->full[stage], 0, 0, h0 - kHalo,
                      c * p.a_n_stride + p.a_n_off + s);
        }
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kFoldMmaWarp && !is_leader) {
    // relay: forward "landed" events of this CTA's operands to the leader's pfull / pwfull barriers
    int stage = 0; uint32_t phase = 0; int cur_client = -1; uint32_t wphase = 0;
    for (int wi = pair; wi < num_work; wi += npairs) {
      const int c = wi / p.nsplit;
      if (p.active && !p.active[c]) continue;
      if (c != cur_client) {
        mbar_wait(&bars->wfull, wphase); wphase ^= 1; cur_client = c;
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->pwfull), 0));
        __syncwarp();
      }
      for (int k = 0; k < steps_per_work; ++k) {
        mbar_wait(&bars->full[stage], phase);
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->pfull[stage]), 0));
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kFoldMmaWarp) {
    const bool leader = elect_one();
    constexpr uint32_t IDESC = make_idesc_bf16(256, Cfg::NF, 0, 1);           // pair MMA, B is MN-major
    const uint64_t a_desc_base = make_smem_desc(smem_u32(a_sm), 16, 8 * ROWB, LT_SW128);
    // B: chunks of [COUT k-rows][16 channels] with 32-byte rows: LBO = chunk stride, SBO = 8 k-rows
    const uint64_t b_desc_base = make_smem_desc(smem_u32(w_sm), Cfg::CHUNK_BYTES, 8 * 32, LT_SW32);
    int stage = 0; uint32_t phase = 0; int cur_client = -1; uint32_t wphase = 0;
    uint32_t seq = 0;                                   // steps issued: buffer = seq % NBUF
    for (int wi = pair; wi < num_work; wi += npairs) {
      const int c = wi / p.nsplit;
      if (p.active && !p.active[c]) continue;
      bool last_of_client = true;
      {
        const int nwi = wi + npairs;
        if (nwi < num_work && (nwi / p.nsplit) == c) last_of_client = false;
      }
      if (c != cur_client) {
        mbar_wait(&bars->wfull, wphase); mbar_wait(&bars->pwfull, wphase); wphase ^= 1; cur_client = c;
      }
      for (int k = 0; k < steps_per_work; ++k, ++seq) {
        const int buf = seq % Cfg::NBUF;
        mbar_wait(&bars->tempty[buf], ((seq / Cfg::NBUF) & 1) ^ 1);
        fence_after_sync();
        mbar_wait(&bars->full[stage], phase);
        mbar_wait(&bars->pfull[stage], phase);
        fence_after_sync();
        if (leader) {
          const uint32_t d_tmem = tmem_base + buf * Cfg::TMEM_BUF_STRIDE;
          const uint64_t a_stage = desc_advance(a_desc_base, stage * Cfg::STAGE_BYTES);
#pragma unroll
          for (int kh = 0; kh < kNKH; ++kh) {
            const uint64_t a_desc = desc_advance(a_stage, (2 * kHalo - kh) * kTileW * ROWB);
            const uint64_t b_desc = desc_advance(b_desc_base, kh * Cfg::W_KH_BYTES);
#pragma unroll
            for (int ks = 0; ks < ROWB / 32; ++ks)
              umma_bf16_2cta(d_tmem, desc_advance(a_desc, ks * 32), desc_advance(b_desc, ks * 16 * 32), IDESC,
                             (kh | ks) ? 1u : 0u);
          }
          umma_commit_2cta(&bars->empty[stage], 3);
          umma_commit_2cta(&bars->tfull[buf], 3);
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
      if (last_of_client) {
        if (leader) umma_commit_2cta(&bars->wempty, 3);
        __syncwarp();
      }
    }
  } else {
    const int grp = warp >> 2;
    uint32_t seq = 0;
    const int q = warp & 3;
    const int tid_g = threadIdx.x & 127;
    const int hl = 2 * q + (lane >> 4), wl = lane & 15;
    const uint32_t tempty_leader0 = mapa_u32(smem_u32(&bars->tempty[0]), 0);
    float dbacc[CIN];
#pragma unroll
    for (int i = 0; i < CIN; ++i) dbacc[i] = 0.f;
    int cur_client = -1;
    bool had_sp0 = false;
    __shared__ float s_db_all[2][CIN];
    float* s_db = s_db_all[grp];
    auto gsync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp)); };
    auto flush_bias = [&](int c) {
      gsync();
      if (tid_g < CIN) s_db[tid_g] = 0.f;
      gsync();
#pragma unroll
      for (int i = 0; i < CIN; ++i) {
        float v = dbacc[i];
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        if (lane == 0) atomicAdd(&s_db[i], v);
        dbacc[i] = 0.f;
      }
      gsync();
      if (tid_g < CIN) {
        const float g = s_db[tid_g];
        if (p.dbias_out) {
          atomicAdd(&p.dbias_out[(size_t)c * CIN + tid_g], g);
        } else {
          float* bp = &p.bias_master[(size_t)c * p.bias_stride + tid_g];
          float upd = -p.lr * g;
          // the FedProx pull is applied once per client: leader CTA, warp set 0, the work item that owns split 0
          if (is_leader && grp == 0 && had_sp0 && p.bias_global) upd -= p.lr * p.mu * (*bp - p.bias_global[tid_g]);
          atomicAdd(bp, upd);
        }
      }
      gsync();
    };
    for (int wi = pair; wi < num_work; wi += npairs) {
      const int c = wi / p.nsplit, sp = wi % p.nsplit;
      if (p.active && !p.active[c]) continue;
      if (c != cur_client) {
        if (cur_client >= 0) flush_bias(cur_client);
        cur_client = c;
        had_sp0 = false;
      }
      if (sp == 0) had_sp0 = true;
      const int t0 = sp * tiles_per_work;
      for (int k = 0; k < steps_per_work; ++k, ++seq) {
        if ((int)(seq & 1) != grp) continue;
        const int buf = seq % Cfg::NBUF;
        const int t = t0 + 2 * k + (int)rank;
        const int s = t / (IMG_H / kTileH);
        const int h0 = (t % (IMG_H / kTileH)) * kTileH;
        const size_t n_out = (size_t)c * BATCH + s;
        const size_t pbase = ((n_out * IMG_H + (h0 + hl)) * IMG_W + wl) * CIN;
        uint4 pq[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) pq[kk] = __ldg(reinterpret_cast<const uint4*>(p.pooled + pbase + kk * 8));
        mbar_wait(&bars->tfull[buf], (seq / Cfg::NBUF) & 1);
        fence_after_sync();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + buf * Cfg::TMEM_BUF_STRIDE;
        float acc[CIN];
        fold_kw_sum<-1>(taddr, lane, wl, nullptr, acc);       // D_kw is taken from the pixel 2-kw columns to the right
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(tempty_leader0 + buf * 8);   // the leader's barrier counts both CTAs' warps
        const uint32_t* pw32 = reinterpret_cast<const uint32_t*>(pq);
        uint4 o[4];
        uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
        for (int i = 0; i < CIN / 2; ++i) {
          const uint32_t pr = pw32[i];
          const float g0 = bf16lo(pr) > 0.f ? acc[2 * i] : 0.f;
          const float g1 = bf16hi(pr) > 0.f ? acc[2 * i + 1] : 0.f;
          dbacc[2 * i] += g0; dbacc[2 * i + 1] += g1;
          ow[i] = pack_bf16x2(g0, g1);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) *reinterpret_cast<uint4*>(p.dpool + pbase + kk * 8) = o[kk];
      }
    }
    if (cur_client >= 0) flush_bias(cur_client);
  }
  fence_before_sync();
  cluster_sync();                       // no CTA may exit (or free TMEM) while its peer can still signal it
  if (warp == kFoldMmaWarp) tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
}

// ================================================================================================
// Forward kernel, horizontal taps folded into N (same column-shift identity as the folded data gradient)
// ================================================================================================
// out[h,w,co] = sum_{kh,kw} X[h+kh-2, w+kw-2, ci] W[co][kh,kw][ci]  =  sum_kw D_kw[h, w+kw-2, co],
// D_kw[h,w',co] = sum_kh X[h+kh-2, w', ci] W[co][kh,kw][ci].  The B rows are ordered (co-chunk j, kw, co%32) so one
// accumulator group of N = 5*32 = 160 columns holds everything a 32-channel output chunk needs; the two groups
// (j = 0,1) ping-pong inside a tile: the epilogue pools group 0 while the tensor core fills group 1.
template <int CIN, int COUT, int IMG_H, int BATCH, int NSTAGE>
struct FwdFoldCfg {
  static constexpr int ROWB = CIN * 2;
  static constexpr int NG = COUT / 32;                  // accumulator groups (32 output channels each)
  static constexpr int NF = 5 * 32;                     // folded N per group
  static constexpr int W_BOX_BYTES = 32 * ROWB;         // one (kh, j, kw) block: 32 rows x CIN
  static constexpr int W_KH_BYTES = NG * 5 * W_BOX_BYTES;
  static constexpr int W_BYTES = kNKH * W_KH_BYTES;
  static constexpr int STAGE_BYTES = kBoxH * kTileW * ROWB;
  static constexpr int TILES_PER_CLIENT = BATCH * (IMG_H / kTileH);
  static constexpr int TMEM_BUF_STRIDE = 256;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = 1024 + W_BYTES + NSTAGE * STAGE_BYTES + (int)sizeof(Barriers);
  static_assert(ROWB == 64 && NG == 2, "written for 32 -> 64 channels (SW64 rows, two accumulator groups)");
};

template <int CIN, int COUT, int IMG_H, int BATCH, int NSTAGE>
__global__ void __launch_bounds__(kFoldThreads, 1)
conv_fwd_fold_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                     const FwdParams p) {
  using Cfg = FwdFoldCfg<CIN, COUT, IMG_H, BATCH, NSTAGE>;
  constexpr int ROWB = Cfg::ROWB;
  constexpr int IMG_W = kTileW;
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
  if (warp == kFoldMmaWarp) { tmem_alloc(&bars->tmem_base, Cfg::TMEM_COLS); tmem_relinquish(); }
  if (warp == kFoldTmaWarp && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const int num_work = p.num_clients * p.nsplit;
  const int tiles_per_work = (Cfg::TILES_PER_CLIENT + p.nsplit - 1) / p.nsplit;

  if (warp == kFoldTmaWarp) {
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0; int cur_client = -1; uint32_t wphase = 0; bool first_w = true;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / p.nsplit, sp = wi % p.nsplit;
      if (p.active && !p.active[c]) continue;
      if (c != cur_client) {
        if (!first_w) { mbar_wait(&bars->wempty, wphase); wphase ^= 1; }
        first_w = false; cur_client = c;
        if (leader) {
          mbar_arrive_expect_tx(&bars->wfull, Cfg::W_BYTES);
          // forward weight copy [C][COUT][25*CIN]: box (CIN, 32 rows, 1) at ((kh*5+kw)*CIN, j*32, c)
#pragma unroll 1
          for (int b = 0; b < kNKH * Cfg::NG * 5; ++b) {
            const int kh = b / (Cfg::NG * 5), r = b % (Cfg::NG * 5), j = r / 5, kw = r % 5;
            tma_load_3d(w_sm + b * Cfg::W_BOX_BYTES, &tmW, &bars->wfull, (kh * 5 + kw) * CIN, j * 32, c);
          }
        }
      }
      const int t0 = sp * tiles_per_work;
      const int t1 = min(t0 + tiles_per_work, Cfg::TILES_PER_CLIENT);
      for (int t = t0; t < t1; ++t) {
        const int s = t / (IMG_H / kTileH);
        const int h0 = (t % (IMG_H / kTileH)) * kTileH;
        mbar_wait(&bars->empty[stage], phase ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&bars->full[stage], Cfg::STAGE_BYTES);
          tma_load_4d(a_sm + stage * Cfg::STAGE_BYTES, &tmA, &bars->full[stage], 0, 0, h0 - kHalo,
                      c * p.a_n_stride + p.a_n_off + s);
        }
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kFoldMmaWarp) {
    const bool leader = elect_one();
    constexpr uint32_t IDESC = make_idesc_bf16(128, Cfg::NF, 0, 0);
    const uint64_t a_desc_base = make_smem_desc(smem_u32(a_sm), 16, 8 * ROWB, LT_SW64);
    const uint64_t b_desc_base = make_smem_desc(smem_u32(w_sm), 16, 8 * ROWB, LT_SW64);
    int stage = 0; uint32_t phase = 0; int cur_client = -1; uint32_t wphase = 0;
    uint32_t tphase = 0;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / p.nsplit, sp = wi % p.nsplit;
      if (p.active && !p.active[c]) continue;
      bool last_of_client = true;
      {
        const int nwi = wi + gridDim.x;
        if (nwi < num_work && (nwi / p.nsplit) == c) last_of_client = false;
      }
      if (c != cur_client) { mbar_wait(&bars->wfull, wphase); wphase ^= 1; cur_client = c; }
      const int t0 = sp * tiles_per_work;
      const int t1 = min(t0 + tiles_per_work, Cfg::TILES_PER_CLIENT);
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&bars->full[stage], phase);
        fence_after_sync();
        const uint64_t a_stage = desc_advance(a_desc_base, stage * Cfg::STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < Cfg::NG; ++j) {
          mbar_wait(&bars->tempty[j], tphase ^ 1);
          fence_after_sync();
          if (leader) {
            const uint32_t d_tmem = tmem_base + j * Cfg::TMEM_BUF_STRIDE;
#pragma unroll
            for (int kh = 0; kh < kNKH; ++kh) {
              const uint64_t a_desc = desc_advance(a_stage, kh * kTileW * ROWB);
              const uint64_t b_desc = desc_advance(b_desc_base, kh * Cfg::W_KH_BYTES + j * 5 * Cfg::W_BOX_BYTES);
#pragma unroll
              for (int ks = 0; ks < ROWB / 32; ++ks)
                umma_bf16(d_tmem, desc_advance(a_desc, ks * 32), desc_advance(b_desc, ks * 32), IDESC,
                          (kh | ks) ? 1u : 0u);
            }
            if (j == Cfg::NG - 1) umma_commit(&bars->empty[stage]);
            umma_commit(&bars->tfull[j]);
          }
          __syncwarp();
        }
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        tphase ^= 1;
      }
      if (last_of_client) {
        if (leader) umma_commit(&bars->wempty);
        __syncwarp();
      }
    }
  } else {
    // epilogue warp set j (warps 4j .. 4j+3) owns accumulator group j: two warps per scheduler, one per group
    uint32_t tphase = 0;
    const int q = warp & 3;
    const int jg = warp >> 2;
    const int wl = lane & 15;
    const int b0 = lane & 1, b4 = (lane >> 4) & 1;
    const int sl = b0 * 2 + b4;
    float bias_r[32];
    int cur_client = -1;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / p.nsplit, sp = wi % p.nsplit;
      if (p.active && !p.active[c]) continue;
      if (c != cur_client) {
        cur_client = c;
        const float* bias = p.bias + (size_t)c * p.bias_stride;
#pragma unroll
        for (int k = 0; k < 32; ++k) bias_r[k] = __ldg(&bias[jg * 32 + k]);
      }
      const int t0 = sp * tiles_per_work;
      const int t1 = min(t0 + tiles_per_work, Cfg::TILES_PER_CLIENT);
      for (int t = t0; t < t1; ++t) {
        const int s = t / (IMG_H / kTileH);
        const int h0 = (t % (IMG_H / kTileH)) * kTileH;
        const size_t n_out = (size_t)c * BATCH + s;
        const int ph = (h0 >> 1) + q, pw = wl >> 1;
        const size_t obase = ((n_out * (IMG_H / 2) + ph) * (IMG_W / 2) + pw) * COUT;
        {
          const int j = jg;
          mbar_wait(&bars->tfull[j], tphase);
          fence_after_sync();
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + j * Cfg::TMEM_BUF_STRIDE;
          float acc[32];
          fold_kw_sum<+1>(taddr, lane, wl, bias_r, acc);
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->tempty[j]);          // group drained: the tensor core may refill it
          uint4 pv;
          uint2 cv;
          pool2x2_relu(acc, b0, b4, pv, cv);
          const size_t off = obase + j * 32 + sl * 8;
          *reinterpret_cast<uint4*>(p.out + off) = pv;
          *reinterpret_cast<uint2*>(p.code + off) = cv;
        }
        tphase ^= 1;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == kFoldMmaWarp) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ================================================================================================
// conv2 forward, shifted-operand form
// ================================================================================================
// The folded kernel above reads 320 accumulator columns per pixel row back from TMEM (64 B/clk/SM: 2560 cycles per
// tile) and pays 128 shuffles per lane for the column shift.  Here every tap is its own MMA on a *shifted view* of one
// halo box instead: tiles are 16 rows x 8 pixels, so the 128 M rows form sixteen 8-pixel groups, one per image row,
// whose stride is the box pitch — uniform for ANY box width.  That makes both tap directions pure start-address
// shifts of the K-major A descriptor (kw: one 64-byte pixel, kh: one 12-pixel box row), the accumulator is 64 columns
// per tile, and the epilogue is just bias + pool.  The price, 50 N=64 MMAs whose weight operand would be re-fetched per
// tile, is paid once per TWO tiles with weight-stationary MMAs (B collector fill / lastuse).
template <int CIN, int COUT, int IMG_H, int IMG_W, int BATCH, int NSTAGE>
struct FwdShiftCfg {
  static constexpr int TH = 16, TW = 8;                 // tile: 16 rows x 8 pixels
  static constexpr int BOX_H = TH + 2 * kHalo, BOX_W = TW + 2 * kHalo;
  static constexpr int ROWB = CIN * 2;
  static constexpr int PITCH = BOX_W * ROWB;            // bytes per box row = stride between the 8-pixel M groups
  static constexpr int STAGE_BYTES = BOX_H * PITCH;     // 15360: a multiple of the 512-byte SW64 pattern
  static constexpr int W_TAP_BYTES = COUT * ROWB;       // [COUT rows][CIN] K-major
  static constexpr int W_BYTES = 25 * W_TAP_BYTES;
  static constexpr int TILES_PER_CLIENT = BATCH * (IMG_H / TH) * (IMG_W / TW);
  static constexpr int NT = 2;                          // tiles per weight-stationary batch
  static constexpr int NBUF = 2 * NT;                   // accumulator buffers (two batches in flight)
  static constexpr int TMEM_COLS = 256;
  static constexpr int SMEM_BYTES = 1024 + W_BYTES + NSTAGE * STAGE_BYTES + (int)sizeof(Barriers);
  static_assert(ROWB == 64 && COUT == 64 && IMG_H == TH && STAGE_BYTES % 512 == 0 && NSTAGE % NT == 0, "conv2 geometry");
};

template <int CIN, int COUT, int IMG_H, int IMG_W, int BATCH, int NSTAGE>
__global__ void __launch_bounds__(kFoldThreads, 1)
conv_fwd_shift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                      const FwdParams p) {
  using Cfg = FwdShiftCfg<CIN, COUT, IMG_H, IMG_W, BATCH, NSTAGE>;
  constexpr int ROWB = Cfg::ROWB, NT = Cfg::NT;
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
    for (int i = 0; i < Cfg::NBUF; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == kFoldMmaWarp) { tmem_alloc(&bars->tmem_base, Cfg::TMEM_COLS); tmem_relinquish(); }
  if (warp == kFoldTmaWarp && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const int num_work = p.num_clients * p.nsplit;
  const int tiles_per_work = Cfg::TILES_PER_CLIENT / p.nsplit;          // a multiple of NT (host picks nsplit)
  constexpr int TPS = IMG_W / Cfg::TW;                                    // tiles per sample

  if (warp == kFoldTmaWarp) {
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0; int cur_client = -1; uint32_t wphase = 0; bool first_w = true;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / p.nsplit, sp = wi % p.nsplit;
      if (p.active && !p.active[c]) continue;
      if (c != cur_client) {
        if (!first_w) { mbar_wait(&bars->wempty, wphase); wphase ^= 1; }
        first_w = false; cur_client = c;
        if (leader) {
          mbar_arrive_expect_tx(&bars->wfull, Cfg::W_BYTES);
          // forward weight copy [C][COUT][25*CIN]: one K-major [COUT][CIN] block per tap
#pragma unroll 1
          for (int tap = 0; tap < 25; ++tap)
            tma_load_3d(w_sm + tap * Cfg::W_TAP_BYTES, &tmW, &bars->wfull, tap * CIN, 0, c);
        }
      }
      const int t0 = sp * tiles_per_work;
      for (int t = t0; t < t0 + tiles_per_work; ++t) {
        const int s = t / TPS, w0 = (t % TPS) * Cfg::TW;
        mbar_wait(&bars->empty[stage], phase ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&bars->full[stage], Cfg::STAGE_BYTES);
          tma_load_4d(a_sm + stage * Cfg::STAGE_BYTES, &tmA, &bars->full[stage], 0, w0 - kHalo, -kHalo,
                      c * p.a_n_stride + p.a_n_off + s);
        }
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kFoldMmaWarp) {
    const bool leader = elect_one();
    constexpr uint32_t IDESC = make_idesc_bf16(128, COUT, 0, 0);
    // A: sixteen 8-pixel groups (one per image row of the tile) at the box pitch; B: [COUT][CIN] per tap
    const uint64_t a_desc_base = make_smem_desc(smem_u32(a_sm), 16, Cfg::PITCH, LT_SW64);
    const uint64_t b_desc_base = make_smem_desc(smem_u32(w_sm), 16, 8 * ROWB, LT_SW64);
    int stage = 0; uint32_t phase = 0; int cur_client = -1; uint32_t wphase = 0;
    uint32_t bseq = 0;                                   // batches issued: buffers (bseq % 2) * NT + {0, 1}
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / p.nsplit;
      if (p.active && !p.active[c]) continue;
      bool last_of_client = true;
      {
        const int nwi = wi + gridDim.x;
        if (nwi < num_work && (nwi / p.nsplit) == c) last_of_client = false;
      }
      if (c != cur_client) { mbar_wait(&bars->wfull, wphase); wphase ^= 1; cur_client = c; }
      for (int b = 0; b < tiles_per_work / NT; ++b, ++bseq) {
        const int buf0 = (bseq & 1) * NT;
        const uint32_t tpar = ((bseq >> 1) & 1) ^ 1;
        mbar_wait(&bars->tempty[buf0], tpar);
        mbar_wait(&bars->tempty[buf0 + 1], tpar);
        fence_after_sync();
        const int stage1 = stage + 1;                     // NSTAGE is a multiple of NT: a batch never wraps the ring
        mbar_wait(&bars->full[stage], phase);
        mbar_wait(&bars->full[stage1], phase);
        fence_after_sync();
        if (leader) {
          const uint32_t d0 = tmem_base + buf0 * COUT, d1 = d0 + COUT;
          const uint64_t a0 = desc_advance(a_desc_base, stage * Cfg::STAGE_BYTES);
          const uint64_t a1 = desc_advance(a_desc_base, stage1 * Cfg::STAGE_BYTES);
#pragma unroll
          for (int tap = 0; tap < 25; ++tap) {
            const uint32_t aoff = (tap / 5) * Cfg::PITCH + (tap % 5) * ROWB;    // (kh, kw) as a byte shift of the view
            const uint64_t bt = desc_advance(b_desc_base, tap * Cfg::W_TAP_BYTES);
#pragma unroll
            for (int ks = 0; ks < ROWB / 32; ++ks) {
              const uint32_t acc = (tap | ks) ? 1u : 0u;
              const uint64_t bd = desc_advance(bt, ks * 32);
              umma_bf16_ws<WS_FILL>(d0, desc_advance(a0, aoff + ks * 32), bd, IDESC, acc);
              umma_bf16_ws<WS_LAST>(d1, desc_advance(a1, aoff + ks * 32), bd, IDESC, acc);
            }
          }
          umma_commit(&bars->empty[stage]);
          umma_commit(&bars->empty[stage1]);
          umma_commit(&bars->tfull[buf0]);
          umma_commit(&bars->tfull[buf0 + 1]);
        }
        __syncwarp();
        stage += NT;
        if (stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
      if (last_of_client) {
        if (leader) umma_commit(&bars->wempty);
        __syncwarp();
      }
    }
  } else {
    // epilogue warp set g (warps 4g .. 4g+3) takes tile g of every batch: lane = (row 4q + lane/8, pixel lane%8)
    const int q = warp & 3, g = warp >> 2;
    const int b0 = lane & 1, b3 = (lane >> 3) & 1;
    const int sl = b0 * 2 + b3;
    float bias_r[COUT];
    int cur_client = -1;
    uint32_t bseq = 0;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int c = wi / p.nsplit, sp = wi % p.nsplit;
      if (p.active && !p.active[c]) continue;
      if (c != cur_client) {
        cur_client = c;
        const float* bias = p.bias + (size_t)c * p.bias_stride;
#pragma unroll
        for (int k = 0; k < COUT; ++k) bias_r[k] = __ldg(&bias[k]);
      }
      const int t0 = sp * tiles_per_work;
      for (int b = 0; b < tiles_per_work / NT; ++b, ++bseq) {
        const int t = t0 + b * NT + g;
        const int s = t / TPS, w0 = (t % TPS) * Cfg::TW;
        const int buf = (bseq & 1) * NT + g;
        const size_t n_out = (size_t)c * BATCH + s;
        const int ph = 2 * q + (lane >> 4), pw = (w0 >> 1) + ((lane & 7) >> 1);
        const size_t obase = ((n_out * (IMG_H / 2) + ph) * (IMG_W / 2) + pw) * COUT;
        mbar_wait(&bars->tfull[buf], (bseq >> 1) & 1);
        fence_after_sync();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + buf * COUT;
        uint32_t va[32], vb[32];
        tmem_ld32(taddr, va);
        tmem_ld32(taddr + 32, vb);
        tmem_ld_wait();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tempty[buf]);          // accumulator is in registers
        float x[32];
        uint4 pv;
        uint2 cv;
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(va[i]) + bias_r[i];
        pool2x2_relu<8>(x, b0, b3, pv, cv);
        *reinterpret_cast<uint4*>(p.out + obase + sl * 8) = pv;
        *reinterpret_cast<uint2*>(p.code + obase + sl * 8) = cv;
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(vb[i]) + bias_r[32 + i];
        pool2x2_relu<8>(x, b0, b3, pv, cv);
        *reinterpret_cast<uint4*>(p.out + obase + 32 + sl * 8) = pv;
        *reinterpret_cast<uint2*>(p.code + obase + 32 + sl * 8) = cv;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == kFoldMmaWarp) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
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
  // SYNTH_Y: the dY operand is synthesised from the pooled gradient instead of being loaded by TMA
  int x_rowmap;                  // X tensor map is the no-swizzle row map: x coordinate = w0 / 16
  int x_planes;                  // X rows are stored as two column-parity planes: tile column index = parity plane
  int mma_ws;                    // 1: weight-stationary MMAs (dY slice latched in the B collector across the M tiles)
  int desc_mode;                 // wide-box variant: 1 = encode the swizzle base offset of unaligned operand starts
  const __nv_bfloat16* dpool;    // [C*BATCH][H/2][W/2][BN]
  const uint8_t* pcode;          // same shape, argmax position 0..3
};

// XBW = width (pixels) of the X halo box.  kTileW: one box per horizontal tap (the tap is the TMA x-coordinate).
// kTileW + 2*kHalo: ONE box with both halos per chunk; the horizontal tap becomes a 1-pixel (= one smem row) shift of
// the MN-major A operand's start address, which cuts the L2->smem traffic of the layer input 4x (12x20 pixels once
// instead of 5 boxes of 12x16).
template <int XROWB, int NKW, int YROWB, int NSS, int XBW = kTileW>
struct WgradCfg {
  static constexpr int KT = XROWB / 2;          // input channels per tap chunk
  static constexpr int BN = YROWB / 2;          // output channels
  static constexpr int NTAPS = NKW * kNKH;
  static constexpr int MTILES = (NKW == 5) ? 7 : 1;
  static constexpr bool WIDE = XBW != kTileW;
  static constexpr int NXB = WIDE ? 1 : NKW;    // X boxes per stage
  static constexpr int XPITCH = XBW * XROWB;    // bytes per box row
  static constexpr int XBOX_BYTES = kBoxH * XPITCH;
  static constexpr int XSTAGE = (XBOX_BYTES + 1023) / 1024 * 1024;
  static constexpr int YSTAGE = kTileH * kTileW * YROWB;
  static constexpr int SS_BYTES = NXB * XSTAGE + YSTAGE;
  static constexpr int TMEM_NEED = MTILES * BN;
  static constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : (TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512)));
  static constexpr int SMEM_BYTES = 1024 + NSS * SS_BYTES + (int)sizeof(Barriers);
  static_assert(SS_BYTES % 1024 == 0, "superstage must keep 1024B alignment");
};

template <int XROWB, int NKW, int YROWB, int IMG_H, int IMG_W, int BATCH, int NSS, bool SYNTH_Y, int NEPG = 1,
          int XBW = kTileW>
__global__ void __launch_bounds__((4 * NEPG + 2) * 32, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                  const WgradParams p) {
  using Cfg = WgradCfg<XROWB, NKW, YROWB, NSS, XBW>;
  static_assert(!Cfg::WIDE || (NKW == 5 && IMG_W == kTileW && !SYNTH_Y), "wide box: conv2 geometry only");
  constexpr int KT = Cfg::KT, BN = Cfg::BN;
  constexpr int TILES_X = IMG_W / kTileW, TILES_Y = IMG_H / kTileH;
  constexpr int CHUNKS = BATCH * TILES_X * TILES_Y;   // pixel chunks (K dimension, 128 pixels each)
  static_assert(!SYNTH_Y || YROWB == 64, "operand synthesis is written for 32 output channels (SW64 rows)");
  constexpr int kProdWarp = 4 * NEPG, kMmaWarp = 4 * NEPG + 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Barriers* bars = reinterpret_cast<Barriers*>(smem + NSS * Cfg::SS_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSS; ++i) { mbar_init(&bars->full[i], SYNTH_Y ? 1 + kEpiWarps : 1); mbar_init(&bars->empty[i], 1); }
    mbar_init(&bars->tfull[0], 1);
    // SYNTH_Y: only group 0 drains TMEM (one M tile).  Otherwise the groups drain alternate M tiles: the fused
    // update is latency-bound (batches of 32 dependent-free loads), so twice the warps halve the tensor-idle time.
    mbar_init(&bars->tempty[0], (SYNTH_Y || Cfg::MTILES == 1) ? kEpiWarps : kEpiWarps * NEPG);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) { tmem_alloc(&bars->tmem_base, Cfg::TMEM_COLS); tmem_relinquish(); }
  if (warp == kProdWarp && lane == 0) { tma_prefetch_desc(&tmX); if (!SYNTH_Y) tma_prefetch_desc(&tmY); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == kProdWarp) {
    const bool leader = elect_one();
    int ss = 0; uint32_t phase = 0;
    for (int c = blockIdx.x; c < p.num_clients; c += gridDim.x) {
      if (p.active && !p.active[c]) continue;
      for (int t = 0; t < CHUNKS; ++t) {
        const int s = t / (TILES_X * TILES_Y);
        const int r = t % (TILES_X * TILES_Y);
        const int h0 = (r / TILES_X) * kTileH, w0 = (r % TILES_X) * kTileW;
        mbar_wait(&bars->empty[ss], phase ^ 1);
        if (leader && !p.grad_out && t == (CHUNKS > 16 ? CHUNKS - 12 : 0)) {
          // warm L2 with this client's fp32 master block shortly before the fused-SGD epilogue reads it (issued
          // late on purpose: the activation stream would evict an early prefetch from L2 again)
          const char* wmb = reinterpret_cast<const char*>(p.w_master + (size_t)c * p.p_stride + p.w_off);
          constexpr uint32_t WB = BN * Cfg::NTAPS * KT * 4;
#pragma unroll 1
          for (uint32_t o = 0; o < WB; o += 16384) prefetch_l2_bulk(wmb + o, min(16384u, WB - o));
        }
        if (leader) {
          mbar_arrive_expect_tx(&bars->full[ss], Cfg::NXB * Cfg::XBOX_BYTES + (SYNTH_Y ? 0 : Cfg::YSTAGE));
          uint8_t* base = smem + ss * Cfg::SS_BYTES;
          const int nx = c * p.x_n_stride + p.x_n_off + s;
          if (Cfg::WIDE) {
            tma_load_4d(base, &tmX, &bars->full[ss], 0, w0 - kHalo, h0 - kHalo, nx);
          } else {
#pragma unroll 1
            for (int kw = 0; kw < NKW; ++kw) {
              const int dw = (NKW == 1) ? 0 : (kw - kHalo);
              tma_load_4d(base + kw * Cfg::XSTAGE, &tmX, &bars->full[ss], 0, p.x_rowmap ? (w0 / kTileW) : (w0 + dw),
                          h0 - kHalo, nx);
            }
          }
          if (!SYNTH_Y) tma_load_4d(base + Cfg::NXB * Cfg::XSTAGE, &tmY, &bars->full[ss], 0, w0, h0, c * BATCH + s);
        }
        if (++ss == NSS) { ss = 0; phase ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    const bool leader = elect_one();
    constexpr uint32_t LTX = layout_type_for_row_bytes(XROWB);
    constexpr uint32_t LTY = layout_type_for_row_bytes(YROWB);
    constexpr uint32_t IDESC = make_idesc_bf16(128, BN, 1, 1);
    const uint32_t sbase = smem_u32(smem);
    // A-operand descriptor templates (stage 0): M tile = 4 taps stacked along MN via the leading-dim offset
    constexpr uint32_t XP = Cfg::XPITCH;
    // per-tap strides: next vertical tap = one box row, next horizontal tap = next box (narrow) / next pixel (wide)
    constexpr uint32_t KW_STEP = Cfg::WIDE ? XROWB : Cfg::XSTAGE;
    const uint64_t a_kw = make_smem_desc(sbase, XP, 8 * XROWB, LTX);                                   // kh 0..3 of one kw
    const uint64_t a_r = make_smem_desc(sbase + 4 * XP, KW_STEP, 8 * XROWB, LTX);                      // kh 4 of kw 0..3
    const uint64_t a_l = make_smem_desc(sbase + 4 * KW_STEP + 4 * XP, XP, 8 * XROWB, LTX);
    const uint64_t b_y = make_smem_desc(sbase + Cfg::NXB * Cfg::XSTAGE, 16, 8 * YROWB, LTY);
    int ss = 0; uint32_t phase = 0; uint32_t tphase = 0;
    for (int c = blockIdx.x; c < p.num_clients; c += gridDim.x) {
      if (p.active && !p.active[c]) continue;
      mbar_wait(&bars->tempty[0], tphase ^ 1);
      fence_after_sync();
      for (int t = 0; t < CHUNKS; ++t) {
        mbar_wait(&bars->full[ss], phase);
        fence_after_sync();
        if (leader) {
          const uint32_t soff = ss * Cfg::SS_BYTES;
          const uint64_t b0 = desc_advance(b_y, soff);
          const uint32_t accum = t > 0 ? 1u : 0u;
          auto a_tile = [&](int mt) -> uint64_t {
            return (mt < NKW) ? desc_advance(a_kw, soff + mt * KW_STEP)
                              : (mt == NKW ? desc_advance(a_r, soff) : desc_advance(a_l, soff));
          };
          if (Cfg::MTILES > 1 && p.mma_ws) {
            // k-step outermost: the 7 M tiles of a k-step share the dY slice, latched once in the B collector
#pragma unroll
            for (int ks = 0; ks < (kTileH * kTileW) / 16; ++ks) {
              const uint64_t bd = desc_advance(b0, ks * 16 * YROWB);
              const uint32_t acc = ks > 0 ? 1u : accum;
#pragma unroll
              for (int mt = 0; mt < Cfg::MTILES; ++mt) {
                const uint64_t ad = desc_advance(a_tile(mt), ks * XP);
                if (mt == 0) umma_bf16_ws<WS_FILL>(tmem_base + mt * BN, ad, bd, IDESC, acc);
                else if (mt == Cfg::MTILES - 1) umma_bf16_ws<WS_LAST>(tmem_base + mt * BN, ad, bd, IDESC, acc);
                else umma_bf16_ws<WS_USE>(tmem_base + mt * BN, ad, bd, IDESC, acc);
              }
            }
          } else {
#pragma unroll
            for (int mt = 0; mt < Cfg::MTILES; ++mt) {
              const uint64_t a0 = a_tile(mt);
#pragma unroll
              for (int ks = 0; ks < (kTileH * kTileW) / 16; ++ks) {
                uint64_t ad = desc_advance(a0, ks * XP);            // k-step = one tile row = one box row
                if (Cfg::WIDE && p.desc_mode) ad = desc_set_base_offset(ad);
                umma_bf16(tmem_base + mt * BN, ad, desc_advance(b0, ks * 16 * YROWB), IDESC, ks > 0 ? 1u : accum);
              }
            }
          }
          umma_commit(&bars->empty[ss]);
        }
        __syncwarp();
        if (++ss == NSS) { ss = 0; phase ^= 1; }
      }
      if (leader) umma_commit(&bars->tfull[0]);
      __syncwarp();
      tphase ^= 1;
    }
  } else {
    uint32_t tphase = 0;
    const int q = warp & 3;
    const int grp = warp >> 2;              // synthesis group: owns chunks with (running chunk index) % NEPG == grp
    const int L = 32 * q + lane;            // accumulator row / (SYNTH_Y) pixel of the chunk
    uint32_t cseq = 0;                      // running chunk index of this CTA: stage = cseq % NSS
    const int ci = L / KT, cin = L % KT;    // chunk index inside the M tile, input channel
    for (int c = blockIdx.x; c < p.num_clients; c += gridDim.x) {
      if (p.active && !p.active[c]) continue;
      if (SYNTH_Y) {
        // ---- operand synthesis: dY[pix][ch] = (code == pos) ? dpool : 0, written as a SW64 MN-major tile
        const int rr = L >> 4, cc = L & 15;                      // pixel (row, col) inside the 8x16 patch
        // plane layout: the 16 pixels of a patch row are the columns 2*cc + plane of the image row
        auto posy_of = [&](int t) -> uint32_t {
          const int plane = (t % (TILES_X * TILES_Y)) % TILES_X;
          return static_cast<uint32_t>(((rr & 1) << 1) | (p.x_planes ? plane : (cc & 1)));
        };
        auto src_off = [&](int t) -> size_t {
          const int s = t / (TILES_X * TILES_Y);
          const int r = t % (TILES_X * TILES_Y);
          const int h0 = (r / TILES_X) * kTileH, w0 = (r % TILES_X) * kTileW;
          const size_t n = (size_t)c * BATCH + s;
          const int pcol = p.x_planes ? cc : ((w0 + cc) >> 1);
          return ((n * (IMG_H / 2) + ((h0 + rr) >> 1)) * (IMG_W / 2) + pcol) * BN;
        };
        // register ring, kDepth chunks of global loads in flight per thread (hides ~kDepth x the per-chunk work)
        constexpr int kDepth = 4;
        uint4 gq[kDepth][4], cq[kDepth][2];
        // this group's chunks of the client: t = first + NEPG*i
        const int first = (int)((grp + NEPG - (cseq % NEPG)) % NEPG);
        const int mine = (CHUNKS - first + NEPG - 1) / NEPG;
        auto issue = [&](int i, int slot) {
          if (i < mine) {
            const size_t pb = src_off(first + NEPG * i);
#pragma unroll
            for (int k = 0; k < 4; ++k) gq[slot][k] = __ldg(reinterpret_cast<const uint4*>(p.dpool + pb + k * 8));
#pragma unroll
            for (int k = 0; k < 2; ++k) cq[slot][k] = __ldg(reinterpret_cast<const uint4*>(p.pcode + pb + k * 16));
          }
        };
#pragma unroll
        for (int d = 0; d < kDepth - 1; ++d) issue(d, d);
#pragma unroll 1
        for (int ib = 0; ib < mine; ib += kDepth) {
#pragma unroll
          for (int u = 0; u < kDepth; ++u) {
            const int i = ib + u;
            if (i >= mine) break;
            issue(i + kDepth - 1, (u + kDepth - 1) % kDepth);
            const uint32_t* gw = reinterpret_cast<const uint32_t*>(gq[u]);
            const uint32_t* cw = reinterpret_cast<const uint32_t*>(cq[u]);
            const uint32_t posy = posy_of(first + NEPG * i);
            uint4 o[4];
            uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
            for (int e = 0; e < 16; ++e) {                         // channel pair (2e, 2e+1)
              const uint32_t c2 = (cw[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu;
              uint32_t val = gw[e];
              if ((c2 & 0xFFu) != (uint32_t)posy) val &= 0xFFFF0000u;
              if ((c2 >> 8) != (uint32_t)posy) val &= 0x0000FFFFu;
              ow[e] = val;
            }
            const uint32_t gs = cseq + first + NEPG * i;           // global chunk sequence number
            const int ss = gs % NSS;
            mbar_wait(&bars->empty[ss], ((gs / NSS) & 1) ^ 1);
            uint8_t* ybase = smem + ss * Cfg::SS_BYTES + Cfg::NXB * Cfg::XSTAGE + L * YROWB;
            const int sw = (L >> 1) & 3;                           // SW64: 16B-chunk index ^= bits[7:8] of the byte address
#pragma unroll
            for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(ybase + ((k ^ sw) << 4)) = o[k];
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->full[ss]);
          }
        }
        cseq += CHUNKS;
      }
      constexpr bool kSplitDrain = !SYNTH_Y && Cfg::MTILES > 1;
      if (!kSplitDrain && grp != 0) continue;                       // group 0 drains TMEM / applies the update
      mbar_wait(&bars->tfull[0], tphase);
      fence_after_sync();
      float* wm = p.w_master + (size_t)c * p.p_stride + p.w_off;
      const float* wg = p.w_global ? p.w_global + p.w_off : nullptr;
      __nv_bfloat16* wf = p.w_fwd + (size_t)c * BN * Cfg::NTAPS * KT;
      __nv_bfloat16* wd = p.w_dgrad ? p.w_dgrad + (size_t)c * BN * Cfg::NTAPS * KT : nullptr;
      float* go = p.grad_out ? p.grad_out + (size_t)c * BN * Cfg::NTAPS * KT : nullptr;
#pragma unroll 1
      for (int mt = kSplitDrain ? grp : 0; mt < Cfg::MTILES; mt += kSplitDrain ? NEPG : 1) {
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
            uint32_t wdp[16];
            const size_t idx0 = (size_t)(j * 32) * (Cfg::NTAPS * KT) + tap * KT + cin;
            constexpr size_t ISTR = Cfg::NTAPS * KT;          // stride between consecutive output channels
            // Issue all 32 master-weight loads before the first store: stores through `wm` would otherwise order
            // every later load behind them (no alias information) and serialise 32 DRAM/L2 round trips.
            float wv[32], av[32];
            if (!go) {
#pragma unroll
              for (int i = 0; i < 32; ++i) wv[i] = wm[idx0 + i * ISTR];
              if (wg) {
#pragma unroll
                for (int i = 0; i < 32; ++i) av[i] = __ldg(&wg[idx0 + i * ISTR]);
              }
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const size_t idx = idx0 + i * ISTR;
              const float gacc = __uint_as_float(v[i]);
              float w = 0.f;
              if (go) {
                go[idx] = gacc;
              } else {
                w = wv[i];
                float g = gacc;
                if (wg) g += p.mu * (w - av[i]);
                w -= p.lr * g;
                wm[idx] = w;
                wf[idx] = __float2bfloat16_rn(w);
              }
              // pack pairs for the (contiguous-in-cout) data-gradient copy
              if (i & 1) wdp[i >> 1] |= static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(w))) << 16;
              else wdp[i >> 1] = static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(w)));
            }
            if (wd && !go) {
              uint4* dst = reinterpret_cast<uint4*>(wd + ((size_t)tap * KT + cin) * BN + j * 32);
#pragma unroll
              for (int k = 0; k < 4; ++k) dst[k] = make_uint4(wdp[4 * k], wdp[4 * k + 1], wdp[4 * k + 2], wdp[4 * k + 3]);
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
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace convtc
