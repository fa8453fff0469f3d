// bgemm_tc.cuh — batched ("grouped over clients") tcgen05 GEMM / implicit-GEMM convolution family.
//
//     D_b[M, N] = sum_k A_b[M, k] * B_b[N, k]          b = client, or (client, sample, head) for attention
//
// One persistent warp-specialised kernel (TMA warp, MMA warp, two groups of 4 epilogue warps — one per TMEM accumulator
// buffer —, NSTAGE-deep shared-memory ring) is instantiated over
//   * the A loader:  A_K     K-major rows straight from a row-major tensor                      (activations, dY)
//                    A_MN    MN-major (the tensor is read "transposed" by the tensor core)      (dY^T, P^T, dS^T)
//                    A_CONV  K-major 128-pixel window of an NHWC activation tensor, one TMA box per (tap, 64-channel
//                            chunk); padding = TMA out-of-bounds zero fill, stride 2 = TMA elementStrides
//                    A_HALO  3x3 / stride 1 / 64 -> 64 channels (ResNet stage 1): ONE 18 x 10-pixel halo box per
//                            16 x 8-pixel tile; the 128 M rows are sixteen 8-pixel groups at the box pitch, so every tap
//                            is a start-address shift of the same K-major view, and the client's nine 64 x 64 weight
//                            blocks stay resident in shared memory while a CTA walks a contiguous range of tiles
//                            (6x less L2 -> SM operand traffic than one box per tap: that feed bounds the N = 64 tiles)
//                    A_CONVX MN-major 64-pixel window of the conv input, two (tap, 64-channel chunk) blocks per M tile: the
//                            weight gradient with rows = (tap, cin) and columns = cout (B_MN over dY), so that the fused
//                            SGD epilogue E_SGD_T walks the fp32 master with the 32 lanes of a warp on contiguous bytes
//                    A_HALOW the weight gradient of the same layer: the halo box is the MN-major A operand (M = 2 taps x
//                            64 input channels, the second tap = the descriptor's leading-dimension offset), B = the
//                            16 x 8-pixel dY tile (MN-major), K = pixels; a work item is (client, set of M tiles) and
//                            accumulates over ALL tiles of the client in TMEM (E_SGD_T: rows = (tap, cin), columns = cout)
//   * the B loader:  B_K     K-major (torch Linear weight [out][in], conv weight [cout][tap][cin])
//                    B_MN    MN-major (W read transposed for data gradients, V / K / Q / X for attention / wgrad)
//                    B_WT    MN-major conv weight for the data gradient: box (64 cin, 1 tap, 64 cout) of the forward copy
//                    B_CONVX MN-major 64-pixel window of the conv input for the weight gradient; every 64-column block
//                            of the N tile is its own (tap, cin-chunk) so one MMA accumulates several taps at once
//   * the epilogue:  E_STORE    bf16 row-major store (+ per-column bias, + addend tensor, ReLU / GELU / GELU-backward)
//                    E_GNSTAT   store + GroupNorm statistics (sum, sum of squares per image and group, fp32 atomics)
//                    E_SOFTMAX  row softmax of the 128-wide score tile (a thread owns a whole accumulator row)
//                    E_DSOFT    softmax backward  dS = P o (dP - rowsum(dP o P)) * scale
//                    E_SGD      weight gradient -> SGD / FedProx step on the fp32 master + bf16 operand copy (rows = cout)
//                    E_SGD_T    the same with rows = (tap, cin), columns = cout: lanes of a warp on contiguous master bytes
// This is the per-client training step of the ResNet-18 and BERT-mini families (BASELINE configs 3 and 5): the role
// of the reference's per-phone subprocess (ols_core/taskMgr/utils/utils_run_task.py:481-514).
#pragma once
#include "tc05.cuh"

namespace bg {
using namespace tc05;

constexpr int kThreads = 320;          // warps 0-7 epilogue (group g = warp / 4 owns accumulator buffer g), 8 TMA, 9 MMA
constexpr int kEpiWarps = 4;          // warps per epilogue group = TMEM lane quadrants
constexpr int kTmaWarp = 8, kMmaWarp = 9;
constexpr int KC = 64;                 // K elements per pipeline stage (one 128-byte swizzle row)
enum { A_K = 0, A_MN = 1, A_CONV = 2, A_HALO = 3, A_HALOW = 4, A_CONVX = 5 };
enum { B_K = 0, B_MN = 1, B_CONVX = 2, B_WT = 3 };
enum { E_STORE = 0, E_GNSTAT = 1, E_SOFTMAX = 2, E_DSOFT = 3, E_SGD = 4, E_SGD_T = 5 };
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_GELU_BWD = 3, ACT_RELU_BWD = 4 };   // *_BWD gate / scale by aux_in

struct Bars {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint64_t wfull, wempty;                // A_HALO: resident weight blocks of the current client landed / may be overwritten
  uint32_t tmem_base;
  uint32_t pad;
};

// Plain-old-data; mirrored field by field by ops/bgemm.py (ctypes).  Keep the order in sync.
struct Params {
  int batches, M, N, m_tiles, n_tiles, k_chunks;
  const uint8_t* active; int act_div;          // work of batch bi is skipped unless active[bi / act_div]
  int a_bdiv, b_bdiv;                          // > 0: 4-D operand map, outer coordinates (bi % bdiv, bi / bdiv)
  int pf_ahead;                                // L2 prefetch distance in pipeline chunks (0 = off)
  int direct_store;                            // 1: every thread stores its own row pieces (no staging transpose)
  // ---- convolution geometry -------------------------------------------------------------------------
  int tw_log, th_log;                          // an M tile = 2^tw_log x 2^th_log pixels x (128 >> (tw_log+th_log)) images
  int tiles_x, tiles_y;                        // tiles per image
  int imgs;                                    // images per batch entry (samples per client)
  int src_stride;                              // window step in the source tensor (= elementStrides of its map)
  int src_cimgs, src_img0;                     // source image index = bi * src_cimgs + src_img0 + image (data shards hold
                                               // more samples per client than one batch)
  int cchunks, ntaps, ktap;                    // 64-channel chunks per tap; taps; K elements per tap of the weight operand
  int tap_dy[9], tap_dx[9], tap_w[9];          // window offset (source pixels) and weight tap index of every tap
  int kb_h, kb_n, kb_cpi;                      // B_CONVX: a 64-pixel K chunk = full rows x kb_h x kb_n images; chunks / image
  int nblocks, cin;                            // B_CONVX: valid 64-column blocks (taps * cchunks); source channels
  // ---- epilogue -------------------------------------------------------------------------------------
  __nv_bfloat16* out; long long o_bs1, o_bs0; int o_bdiv; int ld;
  int out_conv, os, py, px, OW, OH;            // out_conv: rows are pixels; output pixel = tile pixel * os + (py, px)
  const float* bias; long long bias_bs; int bias_div;
  const __nv_bfloat16* addend;
  int act; const __nv_bfloat16* aux_in; __nv_bfloat16* aux_out;
  float scale;
  float* gn_stats; int gn_groups, gn_cpg;      // [batches * imgs][groups][2]; channels per group
  float* w_master; long long w_bs; int w_ld; const float* w_global; __nv_bfloat16* w_bf16; long long wb_bs;
  float* grad_out; long long g_bs; float lr, mu;
};

template <int AM, int BM, int N_TILE, int NSTAGE>
struct Cfg {
  static constexpr bool HALO = (AM == A_HALO);
  static constexpr bool HALOW = (AM == A_HALOW);
  static constexpr int HALO_W = 10, HALO_H = 18;                  // box of a 16 x 8-pixel tile with a 1-pixel apron
  static constexpr int HALO_PITCH = HALO_W * 128;                 // bytes per box row = stride between the 8-pixel M groups
  static constexpr int HALO_BYTES = HALO_H * HALO_PITCH;          // 23040 (TMA transaction size)
  static constexpr int W_TAP_BYTES = 64 * 128;                    // one [64][64] bf16 weight block
  static constexpr int W_BYTES = HALO ? 9 * W_TAP_BYTES : 0;      // resident weights of one client
  static constexpr int A_BYTES = 128 * 128;                       // 128 rows x 64 bf16 (either major)
  static constexpr int B_BYTES = (BM == B_K) ? N_TILE * 128 : (N_TILE / 64) * KC * 128;
  static constexpr int HALO_STAGE = (HALO_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGE_BYTES = HALO ? HALO_STAGE : (HALOW ? HALO_STAGE + 128 * 128 : A_BYTES + ((B_BYTES + 1023) / 1024) * 1024);
  static constexpr int TMEM_COLS = (2 * N_TILE <= 32) ? 32 : (2 * N_TILE <= 64 ? 64 : (2 * N_TILE <= 128 ? 128 : (2 * N_TILE <= 256 ? 256 : 512)));
  static constexpr int SCRATCH_BYTES = 8 * 3072;                 // per epilogue warp: 32 rows x 64 B staging block + 1 KB bias row
  static constexpr int SMEM_BYTES = 1024 + NSTAGE * STAGE_BYTES + W_BYTES + SCRATCH_BYTES + (int)sizeof(Bars);
  static_assert(!HALO || N_TILE == 64, "halo form: 64 -> 64 channels");
  static_assert(!HALOW || N_TILE == 192, "halo weight gradient: up to three 64-column accumulators per buffer");
  static_assert(N_TILE % 32 == 0 && N_TILE <= 256, "epilogue works on 32-column chunks");
  static_assert(BM == B_K || N_TILE % 64 == 0, "MN-major B is loaded in 64-column boxes");
  static_assert(NSTAGE <= 8, "barrier arrays hold 8 stages");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// Single-instruction approximations (MUFU): the epilogues are issue-bound, the IEEE-rounded forms cost 5-10x more.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// GELU and its derivative in the tanh form, 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))), on MUFU.TANH: 6 / 11 instructions
// per element instead of ~25 for an erf evaluation — the two GELU GEMMs were 46 % of a BERT layer and ALU-bound in their
// epilogues.  The tanh form differs from torch's erf default by < 3e-4 absolute (plus < 5e-4 from tanh.approx), below the
// bf16 resolution of the stored activations (2^-9 relative); the engine-level test against the erf model bounds the effect.
__device__ __forceinline__ float gelu_f(float x) {
  const float x2 = x * x;
  const float t = tanh_approx(x * fmaf(x2, 0.0356774081f, 0.7978845608f));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float x2 = x * x;
  const float t = tanh_approx(x * fmaf(x2, 0.0356774081f, 0.7978845608f));
  const float a = 0.5f * x * fmaf(-t, t, 1.f);                                  // 0.5 x sech^2(u)
  return fmaf(a, fmaf(x2, 0.1070322243f, 0.7978845608f), fmaf(t, 0.5f, 0.5f));  // + 0.5 (1 + tanh u)
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
// 32 consecutive bf16 (64 bytes, 16-byte aligned) <-> registers
__device__ __forceinline__ void load32(const __nv_bfloat16* p, float* f) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const uint4 u = q[i]; unpack8(u, f + 8 * i); }
}
__device__ __forceinline__ void store32(__nv_bfloat16* p, const float* f) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = pack8(f + 8 * i);
}

// ---------------------------------------------------------------------------------------------------------
// Warp-private staging transposes.  A thread of the epilogue owns one accumulator ROW, so a direct store makes every
// warp-level instruction write 32 scattered 16-byte pieces (half sectors): measured at ~6x the per-byte cost of reads,
// it made the output-heavy GEMMs (K = 256) run at a third of their HBM roofline.  Rows are therefore exchanged through a
// 4 KB swizzled block per warp so that each global instruction touches four full 128-byte lines (8 lanes per row).
// Slot (row, c) holds 16-byte piece c of a 128-byte row segment at sc[row * 8 + (c ^ (row & 7))]: conflict-free both for
// the row-owner access (lane = row) and for the transposed access (8 lanes per row).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
template <int NQ>   // 16-byte pieces per row segment: 8 (128 B: fp32 chunk) or 4 (64 B: bf16 chunk)
__device__ __forceinline__ uint32_t stage_slot(int row, int c) {
  return (NQ == 8 ? row * 8 + (c ^ (row & 7)) : row * 4 + (c ^ ((row >> 1) & 3))) * 16;
}
template <int NQ>
__device__ __forceinline__ void stage_rows_out(uint4* sc_ptr, int lane, const uint4* pk, char* dst_base, size_t row_off, bool rvalid) {
  // pk[0..NQ): this thread's row segment; dst = dst_base + row_off(row) for every row of the warp
  const uint32_t sc = smem_u32(sc_ptr);              // explicit shared-space accesses (a generic pointer compiles to LD.E / ST.E)
  __syncwarp();
#pragma unroll
  for (int c = 0; c < NQ; ++c) sts128(sc + stage_slot<NQ>(lane, c), pk[c]);
  __syncwarp();
  constexpr int RPI = 32 / NQ;                       // rows per instruction
  const int c = lane % NQ;
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int row = RPI * i + lane / NQ;
    const uint32_t lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(row_off), row);
    const uint32_t hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(row_off >> 32), row);
    const int ok = __shfl_sync(0xffffffffu, rvalid ? 1 : 0, row);
    if (ok) *reinterpret_cast<uint4*>(dst_base + ((static_cast<size_t>(hi) << 32) | lo) + c * 16) = lds128(sc + stage_slot<NQ>(row, c));
  }
}
// 32 fp32 values of this thread's row -> 4 packed bf16 pieces
__device__ __forceinline__ void pack32(const float* f, uint4* pk) {
#pragma unroll
  for (int i = 0; i < 4; ++i) pk[i] = pack8(f + 8 * i);
}

struct TileCoord {
  int bi, mt, nt, m0, n0, img0, ty0, tx0;
};

template <int AM, int N_TILE>
__device__ __forceinline__ bool decode_work(const Params& p, int wi, TileCoord& t) {
  const int tpc = p.m_tiles * p.n_tiles;
  t.bi = wi / tpc;
  const int r = wi - t.bi * tpc;
  t.mt = r / p.n_tiles;
  t.nt = r - t.mt * p.n_tiles;
  t.m0 = t.mt * 128;
  t.n0 = t.nt * N_TILE;
  if (p.active && !p.active[t.bi / p.act_div]) return false;
  if (AM == A_CONV || AM == A_HALO || p.out_conv) {
    const int tpi = p.tiles_x * p.tiles_y;
    const int tn = 128 >> (p.tw_log + p.th_log);
    const int ti = t.mt / tpi, rem = t.mt - ti * tpi;
    t.img0 = ti * tn;
    t.ty0 = (rem / p.tiles_x) << p.th_log;
    t.tx0 = (rem % p.tiles_x) << p.tw_log;
  } else {
    t.img0 = t.ty0 = t.tx0 = 0;
  }
  return true;
}

// One pipeline chunk of operands for tile `t`: TMA loads into (a_dst, b_dst) signalling `bar`, or — PF — the same boxes as
// L2 prefetches.  The producer runs the prefetches `kPrefetchAhead` chunks in front of the loads: with at most ~190 KB of
// loads in flight per SM (the shared-memory ring) a kernel whose operands come from HBM is bound by bytes-in-flight /
// latency; prefetched boxes are L2 hits when their load is issued.
template <int AM, int BM, int N_TILE, bool PF>
__device__ __forceinline__ void issue_chunk(const Params& p, const CUtensorMap* tmA, const CUtensorMap* tmB, const TileCoord& t,
                                            int kc, uint8_t* a_dst, uint8_t* b_dst, uint64_t* bar) {
  auto op3 = [&](uint8_t* dst, const CUtensorMap* tm, int c0, int c1, int c2) {
    if (PF) tma_prefetch_3d(tm, c0, c1, c2); else tma_load_3d(dst, tm, bar, c0, c1, c2);
  };
  auto op4 = [&](uint8_t* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
    if (PF) tma_prefetch_4d(tm, c0, c1, c2, c3); else tma_load_4d(dst, tm, bar, c0, c1, c2, c3);
  };
  int tap = 0, ch0 = 0;
  if (AM == A_CONV) { tap = kc / p.cchunks; ch0 = (kc - tap * p.cchunks) * 64; }
  // ---- A
  if (AM == A_K) {
    if (p.a_bdiv) op4(a_dst, tmA, kc * KC, t.m0, t.bi % p.a_bdiv, t.bi / p.a_bdiv);
    else op3(a_dst, tmA, kc * KC, t.m0, t.bi);
  } else if (AM == A_CONVX) {
    const int ci = kc / p.kb_cpi;
    const int img = ci * p.kb_n, y0 = (kc - ci * p.kb_cpi) * p.kb_h;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int blk = t.mt * 2 + j;
      int tp = blk / p.cchunks;
      int c0 = (blk - tp * p.cchunks) * 64;
      if (blk >= p.nblocks) { tp = 0; c0 = p.cin + 64; }           // padding block: out of bounds -> zeros
      op4(a_dst + j * KC * 128, tmA, c0, p.tap_dx[tp], y0 * p.src_stride + p.tap_dy[tp], t.bi * p.src_cimgs + p.src_img0 + img);
    }
  } else if (AM == A_MN) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (p.a_bdiv) op4(a_dst + j * KC * 128, tmA, t.m0 + 64 * j, kc * KC, t.bi % p.a_bdiv, t.bi / p.a_bdiv);
      else op3(a_dst + j * KC * 128, tmA, t.m0 + 64 * j, kc * KC, t.bi);
    }
  } else {
    op4(a_dst, tmA, ch0, t.tx0 * p.src_stride + p.tap_dx[tap], t.ty0 * p.src_stride + p.tap_dy[tap],
        t.bi * p.src_cimgs + p.src_img0 + t.img0);
  }
  // ---- B
  if (BM == B_K) {
    const int kb = (AM == A_CONV) ? p.tap_w[tap] * p.ktap + ch0 : kc * KC;
    if (p.b_bdiv) op4(b_dst, tmB, kb, t.n0, t.bi % p.b_bdiv, t.bi / p.b_bdiv);
    else op3(b_dst, tmB, kb, t.n0, t.bi);
  } else if (BM == B_MN) {
#pragma unroll
    for (int j = 0; j < N_TILE / 64; ++j) {
      if (p.b_bdiv) op4(b_dst + j * KC * 128, tmB, t.n0 + 64 * j, kc * KC, t.bi % p.b_bdiv, t.bi / p.b_bdiv);
      else op3(b_dst + j * KC * 128, tmB, t.n0 + 64 * j, kc * KC, t.bi);
    }
  } else if (BM == B_WT) {
#pragma unroll
    for (int j = 0; j < N_TILE / 64; ++j) op4(b_dst + j * KC * 128, tmB, t.n0 + 64 * j, p.tap_w[tap], ch0, t.bi);
  } else {  // B_CONVX
    const int ci = kc / p.kb_cpi;
    const int img = ci * p.kb_n, y0 = (kc - ci * p.kb_cpi) * p.kb_h;
#pragma unroll
    for (int j = 0; j < N_TILE / 64; ++j) {
      const int blk = t.nt * (N_TILE / 64) + j;
      int tp = blk / p.cchunks;
      int c0 = (blk - tp * p.cchunks) * 64;
      if (blk >= p.nblocks) { tp = 0; c0 = p.cin + 64; }           // padding block: out of bounds -> zeros
      op4(b_dst + j * KC * 128, tmB, c0, p.tap_dx[tp], y0 * p.src_stride + p.tap_dy[tp], t.bi * p.src_cimgs + p.src_img0 + img);
    }
  }
}

template <int AM, int BM, int N_TILE, int EPI, int NSTAGE>
__global__ void __launch_bounds__(kThreads, 1)   // 10 warps are allocated as 12: 168 registers per thread
bgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ Params p) {
  using C = Cfg<AM, BM, N_TILE, NSTAGE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_sm = smem + NSTAGE * C::STAGE_BYTES;                  // A_HALO: nine resident weight blocks
  uint4* scratch_all = reinterpret_cast<uint4*>(smem + NSTAGE * C::STAGE_BYTES + C::W_BYTES);
  Bars* bars = reinterpret_cast<Bars*>(smem + NSTAGE * C::STAGE_BYTES + C::W_BYTES + C::SCRATCH_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], kEpiWarps); }
    mbar_init(&bars->wfull, 1); mbar_init(&bars->wempty, 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) { tmem_alloc(&bars->tmem_base, C::TMEM_COLS); tmem_relinquish(); }
  if (warp == kTmaWarp && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const int num_work = p.batches * p.m_tiles * p.n_tiles;
  // work items of this CTA: strided over the grid, or (A_HALO) one contiguous range so that a CTA changes client — and
  // reloads the resident weights — at most twice
  const int wi_begin = C::HALO ? (int)((long long)num_work * blockIdx.x / gridDim.x) : (int)blockIdx.x;
  const int wi_end = C::HALO ? (int)((long long)num_work * (blockIdx.x + 1) / gridDim.x) : num_work;
  const int wi_step = C::HALO ? 1 : (int)gridDim.x;

  if (warp == kTmaWarp && C::HALOW) {
    // ================================================================ TMA producer, halo weight gradient
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0;
    const int tpi = p.tiles_x * p.tiles_y, ktiles = p.imgs * tpi;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int bi = wi >> 1;
      if (p.active && !p.active[bi / p.act_div]) continue;
      for (int kt = 0; kt < ktiles; ++kt) {
        const int img = kt / tpi, rem = kt - img * tpi;
        const int ty0 = (rem / p.tiles_x) << p.th_log, tx0 = (rem % p.tiles_x) << p.tw_log;
        mbar_wait(&bars->empty[stage], phase ^ 1);
        if (leader) {
          uint8_t* dst = smem + stage * C::STAGE_BYTES;
          mbar_arrive_expect_tx(&bars->full[stage], C::HALO_BYTES + 128 * 128);
          tma_load_4d(dst, &tmA, &bars->full[stage], 0, tx0 - 1, ty0 - 1, bi * p.src_cimgs + p.src_img0 + img);   // x halo
          tma_load_4d(dst + C::HALO_STAGE, &tmB, &bars->full[stage], 0, tx0, ty0, bi * p.imgs + img);             // dY tile
        }
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kMmaWarp && C::HALOW) {
    // ================================================================ MMA issuer, halo weight gradient
    // M tile mt = weight taps (2mt, 2mt+1) x 64 input channels: tap (kh, kw) is the halo view shifted by kh rows and kw
    // pixels; the second tap of the pair is reached through the leading-dimension offset of the MN-major descriptor
    const bool leader = elect_one();
    constexpr uint32_t IDESC = make_idesc_bf16(128, 64, 1, 1);
    const uint32_t sbase = smem_u32(smem);
    uint64_t a_desc[5];
#pragma unroll
    for (int mt = 0; mt < 5; ++mt) {
      const int t0 = 2 * mt, t1 = (2 * mt + 1 < 9) ? 2 * mt + 1 : 2 * mt;
      const uint32_t o0 = ((t0 / 3) * C::HALO_W + (t0 % 3)) * 128, o1 = ((t1 / 3) * C::HALO_W + (t1 % 3)) * 128;
      a_desc[mt] = make_smem_desc(sbase + o0, (o1 > o0) ? (o1 - o0) : 128, C::HALO_PITCH, LT_SW128);
    }
    const uint64_t b_desc0 = make_smem_desc(sbase + C::HALO_STAGE, KC * 128, 1024, LT_SW128);
    const int ktiles = p.imgs * p.tiles_x * p.tiles_y;
    int stage = 0; uint32_t phase = 0; int buf = 0; uint32_t tphase0 = 0, tphase1 = 0;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int bi = wi >> 1, set = wi & 1;
      if (p.active && !p.active[bi / p.act_div]) continue;
      mbar_wait(&bars->tempty[buf], (buf ? tphase1 : tphase0) ^ 1);
      fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * N_TILE;
      for (int kt = 0; kt < ktiles; ++kt) {
        mbar_wait(&bars->full[stage], phase);
        fence_after_sync();
        if (leader) {
          const uint32_t soff = stage * C::STAGE_BYTES;
          const uint64_t b_s = desc_advance(b_desc0, soff);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (set == 1 && j == 2) break;                       // set 0: M tiles 0,1,2; set 1: M tiles 3,4
            const uint64_t a_s = desc_advance(set ? a_desc[3 + (j & 1)] : a_desc[j], soff);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)                       // K16 = two image rows of the tile (8 pixels each)
              umma_bf16(d_tmem + j * 64, desc_advance(a_s, ks * 2 * C::HALO_PITCH), desc_advance(b_s, ks * 2048), IDESC,
                        (kt | ks) ? 1u : 0u);
          }
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
  } else if (warp < 2 * kEpiWarps && C::HALOW) {
    // ================================================================ epilogue, halo weight gradient: accumulator row =
    // (tap slot, cin), column = cout -> W[cout][tap][cin]: for a fixed column the 32 lanes of a warp touch 128 contiguous
    // bytes of the fp32 master, so the update runs on direct coalesced accesses
    const int group = warp >> 2, q = warp & 3;
    const int row = 32 * q + lane;
    uint32_t tphase = 0;
    int item_n = 0;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int bi = wi >> 1, set = wi & 1;
      if (p.active && !p.active[bi / p.act_div]) continue;
      if ((item_n++ & 1) != group) continue;
      const int buf = group;
      mbar_wait(&bars->tfull[buf], tphase);
      fence_after_sync();
      const int nmt = set ? 2 : 3;
      for (int j = 0; j < nmt; ++j) {
        const int tap = 2 * (set * 3 + j) + (row >> 6), cin = row & 63;
        const bool valid = tap < 9;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + buf * N_TILE + j * 64;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
          if (!valid) continue;
          const size_t idx0 = (size_t)c0 * p.w_ld + tap * 64 + cin;       // + i * w_ld per output channel
          if (p.grad_out) {
            float* gp = p.grad_out + (size_t)bi * p.g_bs + idx0;
#pragma unroll
            for (int i = 0; i < 32; ++i) gp[(size_t)i * p.w_ld] = __uint_as_float(v[i]);
          } else {
            float* wp = p.w_master + (size_t)bi * p.w_bs + idx0;
            __nv_bfloat16* bp = p.w_bf16 + (size_t)bi * p.wb_bs + idx0;
            float w[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) w[i] = wp[(size_t)i * p.w_ld];    // all loads before any store
            if (p.w_global) {
              const float* ap = p.w_global + idx0;
#pragma unroll
              for (int i = 0; i < 32; ++i) w[i] -= p.lr * (__uint_as_float(v[i]) + p.mu * (w[i] - __ldg(ap + (size_t)i * p.w_ld)));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) w[i] -= p.lr * __uint_as_float(v[i]);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) { wp[(size_t)i * p.w_ld] = w[i]; bp[(size_t)i * p.w_ld] = __float2bfloat16_rn(w[i]); }
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tempty[buf]);
      tphase ^= 1;
    }
  } else if (warp == kTmaWarp && C::HALO) {
    // ================================================================ TMA producer, halo form
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0, wphase = 0; int cur_bi = -1;
    for (int wi = wi_begin; wi < wi_end; ++wi) {
      TileCoord t;
      if (!decode_work<AM, N_TILE>(p, wi, t)) continue;
      if (t.bi != cur_bi) {
        if (cur_bi >= 0) { mbar_wait(&bars->wempty, wphase); wphase ^= 1; }
        cur_bi = t.bi;
        if (leader) {
          mbar_arrive_expect_tx(&bars->wfull, C::W_BYTES);
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            // the [cout][cin] block of weight tap tap_w: K-major B of the forward conv, MN-major B of the data gradient
            if (BM == B_K) tma_load_3d(w_sm + tap * C::W_TAP_BYTES, &tmB, &bars->wfull, p.tap_w[tap] * p.ktap, 0, t.bi);
            else tma_load_4d(w_sm + tap * C::W_TAP_BYTES, &tmB, &bars->wfull, 0, p.tap_w[tap], 0, t.bi);
          }
        }
      }
      mbar_wait(&bars->empty[stage], phase ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&bars->full[stage], C::HALO_BYTES);
        tma_load_4d(smem + stage * C::STAGE_BYTES, &tmA, &bars->full[stage], 0, t.tx0 - 1, t.ty0 - 1,
                    t.bi * p.src_cimgs + p.src_img0 + t.img0);
      }
      if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
    }
  } else if (warp == kMmaWarp && C::HALO) {
    // ================================================================ MMA issuer, halo form: 9 taps x 4 k-steps per tile
    const bool leader = elect_one();
    constexpr uint32_t IDESC = make_idesc_bf16(128, N_TILE, 0, BM == B_K ? 0 : 1);
    const uint32_t sbase = smem_u32(smem);
    const uint64_t a_desc0 = make_smem_desc(sbase, 16, C::HALO_PITCH, LT_SW128);      // 8-pixel groups at the box pitch
    const uint64_t b_desc0 = (BM != B_K) ? make_smem_desc(smem_u32(w_sm), KC * 128, 1024, LT_SW128)
                                         : make_smem_desc(smem_u32(w_sm), 16, 1024, LT_SW128);
    constexpr uint32_t B_STEP = (BM != B_K) ? 2048 : 32;
    uint32_t aoff[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) aoff[tap] = ((p.tap_dy[tap] + 1) * C::HALO_W + (p.tap_dx[tap] + 1)) * 128;
    const int tpc = p.m_tiles * p.n_tiles;
    int stage = 0; uint32_t phase = 0, wphase = 0; int buf = 0; uint32_t tphase0 = 0, tphase1 = 0; int cur_bi = -1;
    for (int wi = wi_begin; wi < wi_end; ++wi) {
      TileCoord t;
      if (!decode_work<AM, N_TILE>(p, wi, t)) continue;
      if (t.bi != cur_bi) { mbar_wait(&bars->wfull, wphase); wphase ^= 1; cur_bi = t.bi; }
      mbar_wait(&bars->tempty[buf], (buf ? tphase1 : tphase0) ^ 1);
      mbar_wait(&bars->full[stage], phase);
      fence_after_sync();
      if (leader) {
        const uint32_t d_tmem = tmem_base + buf * N_TILE;
        const uint64_t a_s = desc_advance(a_desc0, stage * C::STAGE_BYTES);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint64_t b_t = desc_advance(b_desc0, tap * C::W_TAP_BYTES);
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks)
            umma_bf16(d_tmem, desc_advance(a_s, aoff[tap] + ks * 32), desc_advance(b_t, ks * B_STEP), IDESC, (tap | ks) ? 1u : 0u);
        }
        umma_commit(&bars->empty[stage]);
        umma_commit(&bars->tfull[buf]);
        if (wi + 1 >= wi_end || (wi + 1) / tpc != t.bi) umma_commit(&bars->wempty);   // last tile of this client here
      }
      __syncwarp();
      if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      if (buf) tphase1 ^= 1; else tphase0 ^= 1;
      buf ^= 1;
    }
  } else if (warp == kTmaWarp) {
    // ================================================================ TMA producer
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0;
    // prefetch cursor: (work item, chunk) kPrefetchAhead chunks ahead of the load cursor in this CTA's own sequence
    const int pf_ahead = p.pf_ahead;
    int pf_wi = blockIdx.x, pf_kc = 0;
    TileCoord pt;
    bool pf_ok = pf_ahead > 0 && pf_wi < num_work && decode_work<AM, N_TILE>(p, pf_wi, pt);
    auto pf_step = [&]() {                      // advance the prefetch cursor by one chunk (skipping inactive work items)
      if (pf_wi >= num_work) return;
      if (pf_ok && leader) issue_chunk<AM, BM, N_TILE, true>(p, &tmA, &tmB, pt, pf_kc, nullptr, nullptr, nullptr);
      if (!pf_ok || ++pf_kc == p.k_chunks) {
        pf_kc = 0;
        do { pf_wi += gridDim.x; } while (pf_wi < num_work && !(pf_ok = decode_work<AM, N_TILE>(p, pf_wi, pt)));
      }
    };
    for (int i = 0; i < pf_ahead; ++i) pf_step();
    for (int wi = wi_begin; wi < wi_end; wi += wi_step) {
      TileCoord t;
      if (!decode_work<AM, N_TILE>(p, wi, t)) continue;
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        mbar_wait(&bars->empty[stage], phase ^ 1);
        if (leader) {
          uint8_t* a_dst = smem + stage * C::STAGE_BYTES;
          uint8_t* b_dst = a_dst + C::A_BYTES;
          uint64_t* bar = &bars->full[stage];
          mbar_arrive_expect_tx(bar, C::A_BYTES + C::B_BYTES);
          issue_chunk<AM, BM, N_TILE, false>(p, &tmA, &tmB, t, kc, a_dst, b_dst, bar);
        }
        if (pf_ahead > 0) pf_step();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================================================ MMA issuer
    const bool leader = elect_one();
    constexpr bool A_IS_MN = (AM == A_MN || AM == A_CONVX);
    constexpr uint32_t IDESC = make_idesc_bf16(128, N_TILE, A_IS_MN ? 1 : 0, BM == B_K ? 0 : 1);
    const uint32_t sbase = smem_u32(smem);
    const uint64_t a_desc0 = A_IS_MN ? make_smem_desc(sbase, KC * 128, 1024, LT_SW128) : make_smem_desc(sbase, 16, 1024, LT_SW128);
    const uint64_t b_desc0 = (BM != B_K) ? make_smem_desc(sbase + C::A_BYTES, KC * 128, 1024, LT_SW128)
                                         : make_smem_desc(sbase + C::A_BYTES, 16, 1024, LT_SW128);
    constexpr uint32_t A_STEP = A_IS_MN ? 2048 : 32, B_STEP = (BM != B_K) ? 2048 : 32;
    int stage = 0; uint32_t phase = 0; int buf = 0; uint32_t tphase0 = 0, tphase1 = 0;
    for (int wi = wi_begin; wi < wi_end; wi += wi_step) {
      TileCoord t;
      if (!decode_work<AM, N_TILE>(p, wi, t)) continue;
      mbar_wait(&bars->tempty[buf], (buf ? tphase1 : tphase0) ^ 1);
      fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * N_TILE;
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        mbar_wait(&bars->full[stage], phase);
        fence_after_sync();
        if (leader) {
          const uint64_t a_s = desc_advance(a_desc0, stage * C::STAGE_BYTES);
          const uint64_t b_s = desc_advance(b_desc0, stage * C::STAGE_BYTES);
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
    // ================================================================ epilogue: two groups of 4 warps; group g drains accumulator
    // buffer g, i.e. every other tile — the epilogues of consecutive tiles overlap each other as well as the MMAs (the
    // forward GEMMs with K = 256 were epilogue-bound at 9 % tensor-pipe utilisation with a single group)
    const int group = warp >> 2, q = warp & 3;
    const int buf = group;
    uint4* scratch = scratch_all + warp * 192;                       // [0,128) staging block, [128,192) bias row of the tile
    float* sbias = reinterpret_cast<float*>(scratch + 128);
    uint32_t tphase = 0;
    int tile_n = 0;
    for (int wi = wi_begin; wi < wi_end; wi += wi_step) {
      TileCoord t;
      if (!decode_work<AM, N_TILE>(p, wi, t)) continue;
      if ((tile_n++ & 1) != group) continue;
      const int row = 32 * q + lane;
      const int m = t.m0 + row;
      const bool rvalid = m < p.M;
      // ---- output row base
      size_t obase;
      int gimg = 0;
      if (p.out_conv) {
        const int tx = row & ((1 << p.tw_log) - 1), ty = (row >> p.tw_log) & ((1 << p.th_log) - 1);
        const int tn = row >> (p.tw_log + p.th_log);
        const int y = (t.ty0 + ty) * p.os + p.py, x = (t.tx0 + tx) * p.os + p.px;
        gimg = t.bi * p.imgs + t.img0 + tn;
        obase = ((size_t)gimg * p.OH * p.OW + (size_t)y * p.OW + x) * p.ld;
      } else {
        obase = (size_t)(t.bi / p.o_bdiv) * p.o_bs1 + (size_t)(t.bi % p.o_bdiv) * p.o_bs0 + (size_t)m * p.ld;
      }
      if (EPI == E_SGD && rvalid && !p.grad_out) {
        // the fp32 master rows this thread is going to update: pull them into L2 while the tile's MMAs run
        const char* wr = reinterpret_cast<const char*>(p.w_master + (size_t)t.bi * p.w_bs + (size_t)m * p.w_ld + t.n0);
        const int nbytes = min(N_TILE, p.N - t.n0) * 4;
        for (int o = 0; o < nbytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(wr + o));
      }
      if (EPI == E_STORE && p.bias) {
        // the tile's bias row -> shared memory while the MMAs run (one cold L2 access per tile instead of one per chunk)
        const float* bp = p.bias + (size_t)(t.bi / p.bias_div) * p.bias_bs + t.n0;
        __syncwarp();
        for (int i = lane * 4; i < N_TILE; i += 128)
          if (t.n0 + i < p.N) *reinterpret_cast<float4*>(sbias + i) = __ldg(reinterpret_cast<const float4*>(bp + i));
        __syncwarp();
      }
      mbar_wait(&bars->tfull[buf], tphase);
      fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + buf * N_TILE;

      if (EPI == E_STORE || EPI == E_GNSTAT) {
        float gs = 0.f, gq = 0.f;
        int gcur = -1;
        const int rows_per_img = 1 << (p.tw_log + p.th_log);
        auto flush = [&](int g) {
          // reduce over the rows of one image held by this warp (whole warp, or 16-lane halves for 4x4 images)
          if (rows_per_img >= 32) {
#pragma unroll
            for (int o = 16; o; o >>= 1) { gs += __shfl_xor_sync(0xffffffffu, gs, o); gq += __shfl_xor_sync(0xffffffffu, gq, o); }
            if (lane == 0) {
              float* st = p.gn_stats + ((size_t)gimg * p.gn_groups + g) * 2;
              atomicAdd(st, gs); atomicAdd(st + 1, gq);
            }
          } else {
#pragma unroll
            for (int o = 8; o; o >>= 1) { gs += __shfl_xor_sync(0xffffffffu, gs, o); gq += __shfl_xor_sync(0xffffffffu, gq, o); }
            if ((lane & 15) == 0) {
              float* st = p.gn_stats + ((size_t)gimg * p.gn_groups + g) * 2;
              atomicAdd(st, gs); atomicAdd(st + 1, gq);
            }
          }
          gs = gq = 0.f;
        };
        uint32_t v[32];
        tmem_ld32(taddr, v);
        const int nchunks = min(N_TILE / 32, (p.N - t.n0 + 31) / 32);          // warp-uniform
        // second epilogue operand of this row (GELU-backward pre-activation, else the residual addend): the next chunk's 64
        // bytes are requested one chunk ahead — a dependent load per chunk exposed a full memory latency 8 times per tile
        const bool act_bwd = p.act == ACT_GELU_BWD || p.act == ACT_RELU_BWD;
        const __nv_bfloat16* pre = (EPI == E_STORE) ? (act_bwd ? p.aux_in : p.addend) : nullptr;
        // a gated data gradient with a residual has both: the addend is then the second pipelined operand
        constexpr bool TWO_OPS = (BM == B_WT);         // only the conv data gradients combine a gate with a residual
        const __nv_bfloat16* pre2 = (EPI == E_STORE && TWO_OPS && act_bwd) ? p.addend : nullptr;
        uint4 an[4] = {}, bn[4] = {};
        if (EPI == E_STORE && pre && rvalid) {
#pragma unroll
          for (int i = 0; i < 4; ++i) an[i] = reinterpret_cast<const uint4*>(pre + obase + t.n0)[i];
        }
        if (EPI == E_STORE && TWO_OPS && pre2 && rvalid) {
#pragma unroll
          for (int i = 0; i < 4; ++i) bn[i] = reinterpret_cast<const uint4*>(pre2 + obase + t.n0)[i];
        }
#pragma unroll 1
        for (int j = 0; j < nchunks; ++j) {
          const int nb = t.n0 + j * 32;
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          if (j + 1 < nchunks) tmem_ld32(taddr + (j + 1) * 32, v);           // next chunk in flight
          uint4 ac[4], bc[4];
          if (EPI == E_STORE && pre) {
#pragma unroll
            for (int i = 0; i < 4; ++i) ac[i] = an[i];
            if (j + 1 < nchunks && rvalid) {
#pragma unroll
              for (int i = 0; i < 4; ++i) an[i] = reinterpret_cast<const uint4*>(pre + obase + nb + 32)[i];
            }
          }
          if (EPI == E_STORE && TWO_OPS && pre2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) bc[i] = bn[i];
            if (j + 1 < nchunks && rvalid) {
#pragma unroll
              for (int i = 0; i < 4; ++i) bn[i] = reinterpret_cast<const uint4*>(pre2 + obase + nb + 32)[i];
            }
          }
          if (EPI == E_GNSTAT) {
            const int g = nb / p.gn_cpg;
            if (g != gcur) { if (gcur >= 0) flush(gcur); gcur = g; }
#pragma unroll
            for (int i = 0; i < 32; ++i) { gs += f[i]; gq += f[i] * f[i]; }
          }
          const size_t off = (obase + (size_t)nb) * 2;
          uint4 pk[4];
          if (EPI == E_STORE) {
            if (p.bias) {
              const float4* bp = reinterpret_cast<const float4*>(sbias + j * 32);
#pragma unroll
              for (int i = 0; i < 8; ++i) { const float4 b = bp[i]; f[4 * i] += b.x; f[4 * i + 1] += b.y; f[4 * i + 2] += b.z; f[4 * i + 3] += b.w; }
            }
            if (p.addend && rvalid) {
              if (act_bwd && TWO_OPS) {                        // both operands: the addend travels in the second buffer
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  float a8[8];
                  unpack8(bc[i], a8);
#pragma unroll
                  for (int k = 0; k < 8; ++k) f[8 * i + k] += a8[k];
                }
              } else if (act_bwd) {
                float a2[32];
                load32(p.addend + obase + nb, a2);
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] += a2[i];
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  float a8[8];
                  unpack8(ac[i], a8);
#pragma unroll
                  for (int k = 0; k < 8; ++k) f[8 * i + k] += a8[k];
                }
              }
            }
            if (p.act == ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
            } else if (p.act == ACT_GELU) {
              if (p.aux_out) {
                if (p.direct_store) { if (rvalid) store32(p.aux_out + obase + nb, f); }
                else { pack32(f, pk); stage_rows_out<4>(scratch, lane, pk, reinterpret_cast<char*>(p.aux_out), off, rvalid); }
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = gelu_f(f[i]);
            } else if (p.act == ACT_RELU_BWD) {
              // gradient gate of the ReLU that produced aux_in (the layer input this gradient belongs to): the consumer
              // (GroupNorm backward) then needs no mask pass of its own
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float a8[8];
                unpack8(ac[i], a8);
#pragma unroll
                for (int k = 0; k < 8; ++k) f[8 * i + k] = a8[k] > 0.f ? f[8 * i + k] : 0.f;
              }
            } else if (p.act == ACT_GELU_BWD) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float a8[8];
                unpack8(ac[i], a8);
#pragma unroll
                for (int k = 0; k < 8; ++k) f[8 * i + k] *= gelu_grad_f(a8[k]);   // rows beyond M are never stored: no per-element select
                                                                                  // (it compiled to a branch around every MUFU sequence)
              }
            }
          }
          // 32 rows x 64 bytes of the warp, written as full sectors (8 rows x 64 contiguous bytes per instruction)
          if (p.direct_store) { if (rvalid) store32(p.out + obase + nb, f); }
          else { pack32(f, pk); stage_rows_out<4>(scratch, lane, pk, reinterpret_cast<char*>(p.out), off, rvalid); }
        }
        if (EPI == E_GNSTAT && gcur >= 0) flush(gcur);
      } else if (EPI == E_SOFTMAX) {
        // scores tile: N_TILE == N; softmax over the row this thread owns (three passes over TMEM)
        const float sc = p.scale * 1.4426950408889634f;
        float mx = -3.0e38f;
#pragma unroll 1
        for (int j = 0; j < N_TILE / 32; ++j) {
          uint32_t v[32];
          tmem_ld32(taddr + j * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        float sum = 0.f;
#pragma unroll 1
        for (int j = 0; j < N_TILE / 32; ++j) {
          uint32_t v[32];
          tmem_ld32(taddr + j * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) sum += ex2_approx((__uint_as_float(v[i]) - mx) * sc);
        }
        const float inv = rcp_approx(sum);
#pragma unroll 1
        for (int j = 0; j < N_TILE / 32; ++j) {
          uint32_t v[32];
          tmem_ld32(taddr + j * 32, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = ex2_approx((__uint_as_float(v[i]) - mx) * sc) * inv;
          uint4 pk[4];
          pack32(f, pk);
          stage_rows_out<4>(scratch, lane, pk, reinterpret_cast<char*>(p.out), (obase + t.n0 + j * 32) * 2, rvalid);
        }
      } else if (EPI == E_DSOFT) {
        float dot = 0.f;
        // the probabilities of this row are requested one 64-byte chunk ahead of their use in both passes (the second
        // pass hits L1/L2): a dependent load per chunk exposed the memory latency eight times per tile
        const uint4* prow = reinterpret_cast<const uint4*>(p.aux_in + obase + t.n0);
        uint4 an[4] = {};
        if (rvalid) {
#pragma unroll
          for (int i = 0; i < 4; ++i) an[i] = prow[i];
        }
#pragma unroll 1
        for (int j = 0; j < N_TILE / 32; ++j) {
          uint32_t v[32];
          tmem_ld32(taddr + j * 32, v);
          uint4 ac[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) ac[i] = an[i];
          if (rvalid) {
            const int jn = (j + 1 < N_TILE / 32) ? j + 1 : 0;          // last chunk: start the second pass
#pragma unroll
            for (int i = 0; i < 4; ++i) an[i] = prow[jn * 4 + i];
          }
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float a8[8];
            unpack8(ac[i], a8);
#pragma unroll
            for (int k = 0; k < 8; ++k) dot = fmaf(__uint_as_float(v[8 * i + k]), a8[k], dot);
          }
        }
#pragma unroll 1
        for (int j = 0; j < N_TILE / 32; ++j) {
          uint32_t v[32];
          tmem_ld32(taddr + j * 32, v);
          uint4 ac[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) ac[i] = an[i];
          if (rvalid && j + 1 < N_TILE / 32) {
#pragma unroll
            for (int i = 0; i < 4; ++i) an[i] = prow[(j + 1) * 4 + i];
          }
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float a8[8];
            unpack8(ac[i], a8);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[8 * i + k] = a8[k] * (__uint_as_float(v[8 * i + k]) - dot) * p.scale;
          }
          uint4 pk[4];
          pack32(f, pk);
          stage_rows_out<4>(scratch, lane, pk, reinterpret_cast<char*>(p.out), (obase + t.n0 + j * 32) * 2, rvalid);
        }
      } else if (EPI == E_SGD_T) {
        // rows = (tap, cin) blocks of the weight, columns = output channels: W[b][cout][kidx].  For a fixed column the 32
        // lanes of a warp (consecutive cin) touch 128 contiguous bytes of the fp32 master and 64 of the bf16 copy, so the
        // read-modify-write runs on fully coalesced accesses with 32 independent loads in flight per thread
        const int blk = t.mt * 2 + (row >> 6), ch = row & 63;
        const int tp = blk / p.cchunks, c0 = (blk - tp * p.cchunks) * 64;
        const bool kvalid = blk < p.nblocks && c0 + ch < p.cin;
        const size_t kidx = (size_t)tp * p.cin + c0 + ch;
#pragma unroll 1
        for (int j = 0; j < N_TILE / 32; ++j) {
          const int nb = t.n0 + j * 32;
          if (nb >= p.N) break;
          uint32_t v[32];
          tmem_ld32(taddr + j * 32, v);
          tmem_ld_wait();
          if (!kvalid) continue;
          const size_t idx0 = (size_t)nb * p.w_ld + kidx;                  // + i * w_ld per output channel
          if (p.grad_out) {
            float* gp = p.grad_out + (size_t)t.bi * p.g_bs + idx0;
#pragma unroll
            for (int i = 0; i < 32; ++i) gp[(size_t)i * p.w_ld] = __uint_as_float(v[i]);
          } else {
            float* wp = p.w_master + (size_t)t.bi * p.w_bs + idx0;
            __nv_bfloat16* bp = p.w_bf16 + (size_t)t.bi * p.wb_bs + idx0;
            float w[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) w[i] = wp[(size_t)i * p.w_ld];    // all loads before any store
            if (p.w_global) {
              const float* ap = p.w_global + idx0;
#pragma unroll
              for (int i = 0; i < 32; ++i) w[i] -= p.lr * (__uint_as_float(v[i]) + p.mu * (w[i] - __ldg(ap + (size_t)i * p.w_ld)));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) w[i] -= p.lr * __uint_as_float(v[i]);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) { wp[(size_t)i * p.w_ld] = w[i]; bp[(size_t)i * p.w_ld] = __float2bfloat16_rn(w[i]); }
          }
        }
      } else {  // E_SGD: W[b][m][n] -= lr * (g + mu (w - w_global)), bf16 operand copy refreshed in the same pass
        uint32_t vn[32];
        tmem_ld32(taddr, vn);
#pragma unroll 1
        for (int j = 0; j < N_TILE / 32; ++j) {
          const int nb = t.n0 + j * 32;
          if (nb >= p.N) break;
          tmem_ld_wait();
          uint32_t v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = vn[i];
          if (j + 1 < N_TILE / 32 && nb + 32 < p.N) tmem_ld32(taddr + (j + 1) * 32, vn);   // next chunk in flight
          if (rvalid && p.grad_out) {
            float4* gp = reinterpret_cast<float4*>(p.grad_out + (size_t)t.bi * p.g_bs + (size_t)m * p.w_ld + nb);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              gp[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          } else if (rvalid) {
            float4* wp = reinterpret_cast<float4*>(p.w_master + (size_t)t.bi * p.w_bs + (size_t)m * p.w_ld + nb);
            float4 wv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) wv[i] = wp[i];                                    // all loads before any store
            float w[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) { w[4 * i] = wv[i].x; w[4 * i + 1] = wv[i].y; w[4 * i + 2] = wv[i].z; w[4 * i + 3] = wv[i].w; }
            if (p.w_global) {
              const float4* ap = reinterpret_cast<const float4*>(p.w_global + (size_t)m * p.w_ld + nb);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 a = __ldg(ap + i);
                w[4 * i] -= p.lr * (__uint_as_float(v[4 * i]) + p.mu * (w[4 * i] - a.x));
                w[4 * i + 1] -= p.lr * (__uint_as_float(v[4 * i + 1]) + p.mu * (w[4 * i + 1] - a.y));
                w[4 * i + 2] -= p.lr * (__uint_as_float(v[4 * i + 2]) + p.mu * (w[4 * i + 2] - a.z));
                w[4 * i + 3] -= p.lr * (__uint_as_float(v[4 * i + 3]) + p.mu * (w[4 * i + 3] - a.w));
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) w[i] -= p.lr * __uint_as_float(v[i]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) wp[i] = make_float4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
            store32(p.w_bf16 + (size_t)t.bi * p.wb_bs + (size_t)m * p.w_ld + nb, w);
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tempty[buf]);
      tphase ^= 1;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

}  // namespace bg
