// api.cu — C ABI of the sm_100a kernel library (loaded from python with ctypes; no torch headers, so a full
// rebuild is a few seconds).  Pointers are raw device pointers (tensor.data_ptr()), streams are cudaStream_t.
// Every entry point returns 0 on success or a negative/CUDA error code; olsb_last_error() gives the text.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "conv_tc.cuh"
#include "fl_kernels.cuh"
#include "linear_tc.cuh"

using namespace convtc;
using flk::Cnn2Layout;

namespace {
char g_err[512] = {0};
int g_sms = 0;

int fail(const char* what, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return -(int)e - 1000;
}
#define CK(call)                                   \
  do {                                             \
    cudaError_t _e = (call);                       \
    if (_e != cudaSuccess) return fail(#call, _e); \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int ensure_init() {
  if (g_encode && g_sms) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled not available from the driver");
    return -2;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  int dev = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}

CUtensorMapSwizzle swz_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// bf16 tensor, innermost dim first.  strides are in elements (innermost stride is implicitly 1).
int make_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_el,
             const uint32_t* box, int row_bytes) {
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_el[i] * 2;
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz_for(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, row_bytes %d)", (int)r,
             rank, row_bytes);
    return -3;
  }
  return 0;
}

// activation tensor [N][H][W][Cel] -> 4-D map, box (Cel, 16, box_h, 1)
int make_act_map(CUtensorMap* tm, const void* base, uint64_t n, int H, int W, int cel, int box_h, int box_w = 16) {
  const uint64_t dims[4] = {(uint64_t)cel, (uint64_t)W, (uint64_t)H, n};
  const uint64_t str[3] = {(uint64_t)cel, (uint64_t)W * cel, (uint64_t)H * W * cel};
  const uint32_t box[4] = {(uint32_t)cel, (uint32_t)box_w, (uint32_t)box_h, 1u};
  return make_map(tm, base, 4, dims, str, box, cel * 2);
}
// weight tensor [C][rows][k] -> 3-D map, box (box_k, box_rows, 1)
int make_w_map(CUtensorMap* tm, const void* base, uint64_t C, int rows, int k, int box_rows, int box_k) {
  const uint64_t dims[3] = {(uint64_t)k, (uint64_t)rows, C};
  const uint64_t str[2] = {(uint64_t)k, (uint64_t)rows * k};
  const uint32_t box[3] = {(uint32_t)box_k, (uint32_t)box_rows, 1u};
  return make_map(tm, base, 3, dims, str, box, box_k * 2);
}

// conv1 input X5 [N][32][32][16] stored pre-swizzled by the ingest kernel: 4-D map {256 el = 16 px, 2 halves, 32 rows, N}
// without TMA swizzling -> a 12-row halo box is 12 box rows of 512 B instead of 192 rows of 32 B
int make_x5_row_map(CUtensorMap* tm, const void* base, uint64_t n, int box_h, int box_planes = 1) {
  const uint64_t dims[4] = {256, 2, 32, n};
  const uint64_t str[3] = {256, 512, 32 * 512};
  const uint32_t box[4] = {256, (uint32_t)box_planes, (uint32_t)box_h, 1};
  cuuint64_t gd[4], gs[3];
  cuuint32_t bx[4], es[4];
  for (int i = 0; i < 4; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < 3; ++i) gs[i] = str[i] * 2;
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled (x5 row map) failed with CUresult %d", (int)r);
    return -3;
  }
  return 0;
}

constexpr int BATCH = 32;
// kernel instances of the CNN-2 family
//                          ROWB NKW BN  H   W   BATCH  FLIP   EPI         NSTAGE
#define CONV1_FWD conv_taps_kernel<32, 1, 32, 32, 32, BATCH, false, EPI_POOL, 12, 2, 2>
#define CONV2_FWD conv_fwd_fold_kernel<32, 64, 16, BATCH, 7>
#define CONV2_DGRAD conv_dgrad_fold_kernel<32, 64, 16, BATCH, 4>
using Conv1FwdCfg = FwdCfg<32, 1, 32, 32, 32, BATCH, false, EPI_POOL, 12, 2, 2>;
#define CONV1_WINDOW conv1_window_kernel<BATCH, 5>
using Conv1WinCfgT = Conv1WinCfg<BATCH, 5>;
using Conv2FwdCfg = FwdFoldCfg<32, 64, 16, BATCH, 7>;
using Conv2DgradCfg = DgradFoldCfg<32, 64, 16, BATCH, 4>;
#define CONV2_FWD_SHIFT conv_fwd_shift_kernel<32, 64, 16, 16, BATCH, 6>
using Conv2FwdShiftCfg = FwdShiftCfg<32, 64, 16, 16, BATCH, 6>;
#define CONV2_DGRAD_PAIR conv_dgrad_pair_kernel<32, 64, 16, BATCH, 6>
using Conv2DgradPairCfg = DgradPairCfg<32, 64, 16, BATCH, 6>;
//                              XROWB NKW YROWB H   W   BATCH NSS SYNTH_Y
#define CONV2_WGRAD conv_wgrad_kernel<64, 5, 128, 16, 16, BATCH, 2, false, 2>
#define CONV1_WGRAD conv_wgrad_kernel<32, 1, 64, 32, 32, BATCH, 8, true, 2>
// conv2 weight gradient with ONE 12x20 halo box per chunk (horizontal taps = 1-pixel operand shifts)
#define CONV2_WGRAD_WIDE conv_wgrad_kernel<64, 5, 128, 16, 16, BATCH, 6, false, 2, 20>
using Conv2WgradWideCfg = WgradCfg<64, 5, 128, 6, 20>;
using Conv2WgradCfg = WgradCfg<64, 5, 128, 2>;
using Conv1WgradCfg = WgradCfg<32, 1, 64, 8>;

// conv1 input layout (set before data is ingested): 0 plain rows fetched with 32-byte swizzled TMA rows, 1 pre-swizzled
// rows fetched as 512-byte box rows, 2 pre-swizzled column-parity planes (conv1_window_kernel: pooling by accumulator
// layout)
int g_x5_rows = 2;
// conv2 data gradient on CTA pairs (cta_group::2, weight operand split across the pair): 0 never, 1 always,
// -1 auto = when a rank holds fewer than 400 clients (measured: 139 vs 151 us at 160 clients, 380 vs 379 at 592)
int g_dgrad_pair = -1;
// conv2 forward as 25 shifted-view MMAs per tile (16x8 tiles, weight-stationary pairs): 0 never, 1 always, -1 auto =
// when a rank holds fewer than 400 clients (measured: 135 vs 149 us at 160 clients, 397 vs 393 us at 592)
int g_fwd2_shift = -1;
int g_wgrad2_ws = 1;     // 1: weight-stationary (B-collector) MMAs in the conv2 weight gradient
unsigned g_peer_spin_limit = 1u << 28;   // polls of a peer flag before a cross-GPU barrier reports the peer as lost
int g_wgrad2_mode = 1;   // 0: five 12x16 boxes per chunk, 1: one 12x20 box, 2: one box + swizzle base offset
bool g_attr_set = false;
int set_attrs() {
  if (g_attr_set) return 0;
  CK(cudaFuncSetAttribute(CONV1_FWD, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv1FwdCfg::SMEM_BYTES));
  CK(cudaFuncSetAttribute(CONV1_WINDOW, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv1WinCfgT::SMEM_BYTES));
  CK(cudaFuncSetAttribute(CONV2_FWD, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2FwdCfg::SMEM_BYTES));
  CK(cudaFuncSetAttribute(CONV2_DGRAD, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2DgradCfg::SMEM_BYTES));
  CK(cudaFuncSetAttribute(CONV2_FWD_SHIFT, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2FwdShiftCfg::SMEM_BYTES));
  CK(cudaFuncSetAttribute(CONV2_DGRAD_PAIR, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2DgradPairCfg::SMEM_BYTES));
  CK(cudaFuncSetAttribute(CONV2_WGRAD, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2WgradCfg::SMEM_BYTES));
  CK(cudaFuncSetAttribute(CONV1_WGRAD, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv1WgradCfg::SMEM_BYTES));
  CK(cudaFuncSetAttribute(CONV2_WGRAD_WIDE, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2WgradWideCfg::SMEM_BYTES));
  CK(cudaFuncSetAttribute(flk::head_fc_xent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, flk::kHeadSmem));
  g_attr_set = true;
  return 0;
}

int auto_split(int C, int tiles_per_client, int nsplit, int units = 0) {
  // Work items per client.  Cost model: waves(s)/s client-times, where waves = ceil(C*s / #SMs); every extra split
  // re-loads the client's weights (~100 KB), so ties (and near-ties within 6 %) go to the smaller split.
  if (nsplit > 0) return nsplit;
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= 8; s *= 2) {
    if (tiles_per_client % s) break;
    const int nu = units > 0 ? units : g_sms;      // schedulable units: SMs, or CTA pairs for cta_group::2 kernels
    const int waves = (C * s + nu - 1) / nu;
    const double cost = (double)waves / s * (1.0 + 0.06 * (s > 1) + 0.02 * (s > 2) + 0.02 * (s > 4));
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}
}  // namespace

extern "C" {

const char* olsb_last_error() { return g_err; }

int olsb_init() {
  int r = ensure_init();
  if (r) return r;
  return set_attrs();
}
int olsb_num_sms() { return g_sms; }
int olsb_round_record_len() { return flk::kRoundRec; }
int olsb_cnn2_param_count() { return Cnn2Layout::P; }
int olsb_cnn2_batch() { return BATCH; }

// ------------------------------------------------------------------------------------------------
int olsb_ingest_u8(const void* src_u8, void* dst_x5, long long nsamples, void* stream) {
  const size_t npix = (size_t)nsamples * 32 * 32;
  const int th = 256;
  flk::ingest_u8_kernel<<<(unsigned)((npix + th - 1) / th), th, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)src_u8, (__nv_bfloat16*)dst_x5, npix, g_x5_rows);
  CK(cudaGetLastError());
  return 0;
}

// conv1 forward: X5 [x_total][32][32][16] -> P1 [C*32][16][16][32] (+code1)
int olsb_conv1_fwd(const void* x5, long long x_total, int x_n_stride, int x_n_off, const void* w1_fwd,
                   const void* master, long long P, int C, const void* active, void* p1, void* code1, int nsplit,
                   void* stream) {
  if (int r = olsb_init()) return r;
  CUtensorMap tmA, tmW;
  if (g_x5_rows == 2) {
    // window form: one 36-row x 2-plane box per image, work items are whole images
    if (int r = make_x5_row_map(&tmA, x5, (uint64_t)x_total, Conv1WinCfgT::BOX_ROWS, 2)) return r;
    if (int r = make_w_map(&tmW, w1_fwd, (uint64_t)C, 32, 80, 32, 16)) return r;
    FwdParams p;
    memset(&p, 0, sizeof(p));
    p.num_clients = C; p.nsplit = auto_split(C, BATCH, nsplit);
    p.a_n_stride = x_n_stride; p.a_n_off = x_n_off; p.active = (const uint8_t*)active;
    p.bias = (const float*)master + Cnn2Layout::B1; p.bias_stride = (int)P;
    p.out = (__nv_bfloat16*)p1; p.code = (uint8_t*)code1;
    const int work = C * p.nsplit;
    const int grid = work < g_sms ? work : g_sms;
    CONV1_WINDOW<<<grid, Conv1WinCfgT::THREADS, Conv1WinCfgT::SMEM_BYTES, (cudaStream_t)stream>>>(tmA, tmW, p);
    CK(cudaGetLastError());
    return 0;
  }
  if (int r = g_x5_rows ? make_x5_row_map(&tmA, x5, (uint64_t)x_total, kBoxH)
                        : make_act_map(&tmA, x5, (uint64_t)x_total, 32, 32, 16, kBoxH)) return r;
  if (int r = make_w_map(&tmW, w1_fwd, (uint64_t)C, 32, 80, 32, 16)) return r;
  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.a_rowmap = g_x5_rows;
  p.num_clients = C; p.nsplit = auto_split(C, Conv1FwdCfg::TILES_PER_CLIENT, nsplit);
  p.a_n_stride = x_n_stride; p.a_n_off = x_n_off; p.active = (const uint8_t*)active;
  p.bias = (const float*)master + Cnn2Layout::B1; p.bias_stride = (int)P;
  p.out = (__nv_bfloat16*)p1; p.code = (uint8_t*)code1;
  const int work = C * p.nsplit;
  const int grid = work < g_sms ? work : g_sms;
  CONV1_FWD<<<grid, Conv1FwdCfg::THREADS, Conv1FwdCfg::SMEM_BYTES, (cudaStream_t)stream>>>(tmA, tmW, p);
  CK(cudaGetLastError());
  return 0;
}

// conv2 forward: P1 [C*32][16][16][32] -> P2 [C*32][8][8][64] (+code2)
int olsb_conv2_fwd(const void* p1, const void* w2_fwd, const void* master, long long P, int C, const void* active,
                   void* p2, void* code2, int nsplit, void* stream) {
  if (int r = olsb_init()) return r;
  CUtensorMap tmA, tmW;
  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.num_clients = C; p.nsplit = auto_split(C, Conv2FwdCfg::TILES_PER_CLIENT, nsplit);
  p.a_n_stride = BATCH; p.a_n_off = 0; p.active = (const uint8_t*)active;
  p.bias = (const float*)master + Cnn2Layout::B2; p.bias_stride = (int)P;
  p.out = (__nv_bfloat16*)p2; p.code = (uint8_t*)code2;
  const int work = C * p.nsplit;
  const int grid = work < g_sms ? work : g_sms;
  if (g_fwd2_shift > 0 || (g_fwd2_shift < 0 && C < 400)) {
    // 16x8-pixel tiles, one (16+4) x (8+4) halo box per tile; weights as 25 K-major [64][32] blocks
    using SC = Conv2FwdShiftCfg;
    if (int r = make_act_map(&tmA, p1, (uint64_t)C * BATCH, 16, 16, 32, SC::BOX_H, SC::BOX_W)) return r;
    if (int r = make_w_map(&tmW, w2_fwd, (uint64_t)C, 64, 800, 64, 32)) return r;
    if (SC::TILES_PER_CLIENT % (SC::NT * p.nsplit)) {
      snprintf(g_err, sizeof(g_err), "conv2 fwd (shift): tiles per work item must be a multiple of %d", SC::NT);
      return -4;
    }
    CONV2_FWD_SHIFT<<<grid, kFoldThreads, SC::SMEM_BYTES, (cudaStream_t)stream>>>(tmA, tmW, p);
    CK(cudaGetLastError());
    return 0;
  }
  if (int r = make_act_map(&tmA, p1, (uint64_t)C * BATCH, 16, 16, 32, kBoxH)) return r;
  if (int r = make_w_map(&tmW, w2_fwd, (uint64_t)C, 64, 800, 32, 32)) return r;
  CONV2_FWD<<<grid, kFoldThreads, Conv2FwdCfg::SMEM_BYTES, (cudaStream_t)stream>>>(tmA, tmW, p);
  CK(cudaGetLastError());
  return 0;
}

// conv2 data gradient fused with the ReLU-backward gate of layer 1 and the conv1-bias SGD step:
// dY2 [C*32][16][16][64] (x) W2 -> dP1 [C*32][16][16][32] (gradient w.r.t. the pooled conv1 activation, gated by P1>0)
int olsb_conv2_dgrad(const void* dy2, const void* w2_dgrad, void* master, const void* global_w, long long P, int C,
                     const void* active, const void* p1, void* dp1, float lr, float mu, void* dbias_out, int nsplit,
                     void* stream) {
  if (int r = olsb_init()) return r;
  CUtensorMap tmA, tmW;
  const bool pair = g_dgrad_pair > 0 || (g_dgrad_pair < 0 && C < 400);
  if (int r = make_act_map(&tmA, dy2, (uint64_t)C * BATCH, 16, 16, 64, kBoxH)) return r;
  // B operand = forward weight copy [C][64][800]: 32-channel (pair kernel: 16-channel) MN-major chunks per tap
  if (int r = make_w_map(&tmW, w2_dgrad /* = forward copy w2_fwd */, (uint64_t)C, 64, 800, 64, pair ? 16 : 32)) return r;
  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.num_clients = C;
  p.nsplit = auto_split(C, Conv2DgradCfg::TILES_PER_CLIENT, nsplit, pair ? g_sms / 2 : 0);
  p.a_n_stride = BATCH; p.a_n_off = 0; p.active = (const uint8_t*)active;
  p.pooled = (const __nv_bfloat16*)p1; p.dpool = (__nv_bfloat16*)dp1;
  p.bias_master = (float*)master + Cnn2Layout::B1; p.bias_stride = (int)P;
  p.bias_global = global_w ? (const float*)global_w + Cnn2Layout::B1 : nullptr;
  p.lr = lr; p.mu = mu; p.dbias_out = (float*)dbias_out;
  const int work = C * p.nsplit;
  if (pair) {
    if (p.nsplit > 32 || (Conv2DgradPairCfg::TILES_PER_CLIENT % (2 * p.nsplit))) {
      snprintf(g_err, sizeof(g_err), "dgrad pair kernel needs an even number of tiles per work item (nsplit %d)", p.nsplit);
      return -4;
    }
    const int npairs = work < g_sms / 2 ? work : g_sms / 2;
    CONV2_DGRAD_PAIR<<<2 * npairs, kFoldThreads, Conv2DgradPairCfg::SMEM_BYTES, (cudaStream_t)stream>>>(tmA, tmW, p);
    CK(cudaGetLastError());
    return 0;
  }
  const int grid = work < g_sms ? work : g_sms;
  CONV2_DGRAD<<<grid, kFoldThreads, Conv2DgradCfg::SMEM_BYTES, (cudaStream_t)stream>>>(tmA, tmW, p);
  CK(cudaGetLastError());
  return 0;
}

int olsb_conv2_wgrad(const void* p1, const void* dy2, void* master, const void* global_w, long long P, int C,
                     const void* active, void* w2_fwd, void* w2_dgrad, float lr, float mu, void* grad_out,
                     void* stream) {
  if (int r = olsb_init()) return r;
  CUtensorMap tmX, tmY;
  const bool wide = g_wgrad2_mode != 0;
  if (int r = make_act_map(&tmX, p1, (uint64_t)C * BATCH, 16, 16, 32, kBoxH, wide ? 20 : 16)) return r;
  if (int r = make_act_map(&tmY, dy2, (uint64_t)C * BATCH, 16, 16, 64, kTileH)) return r;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.num_clients = C; p.x_n_stride = BATCH; p.x_n_off = 0; p.active = (const uint8_t*)active;
  p.w_master = (float*)master; p.w_global = (const float*)global_w; p.p_stride = (size_t)P;
  p.w_off = Cnn2Layout::W2; p.w_fwd = (__nv_bfloat16*)w2_fwd; p.w_dgrad = (__nv_bfloat16*)w2_dgrad;
  p.lr = lr; p.mu = mu; p.grad_out = (float*)grad_out;
  const int grid = C < g_sms ? C : g_sms;
  p.desc_mode = g_wgrad2_mode == 2 ? 1 : 0;
  p.mma_ws = g_wgrad2_ws;
  if (wide)
    CONV2_WGRAD_WIDE<<<grid, 320, Conv2WgradWideCfg::SMEM_BYTES, (cudaStream_t)stream>>>(tmX, tmY, p);
  else
    CONV2_WGRAD<<<grid, 320, Conv2WgradCfg::SMEM_BYTES, (cudaStream_t)stream>>>(tmX, tmY, p);
  CK(cudaGetLastError());
  return 0;
}

// conv1 weight gradient; the un-pooled output gradient is synthesised in shared memory from dP1 + code1
int olsb_conv1_wgrad(const void* x5, long long x_total, int x_n_stride, int x_n_off, const void* dp1,
                     const void* code1, void* master, const void* global_w, long long P, int C, const void* active,
                     void* w1_fwd, float lr, float mu, void* grad_out, void* stream) {
  if (int r = olsb_init()) return r;
  CUtensorMap tmX;
  if (int r = g_x5_rows ? make_x5_row_map(&tmX, x5, (uint64_t)x_total, kBoxH)
                        : make_act_map(&tmX, x5, (uint64_t)x_total, 32, 32, 16, kBoxH)) return r;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.x_rowmap = g_x5_rows != 0;
  p.x_planes = g_x5_rows == 2;
  p.num_clients = C; p.x_n_stride = x_n_stride; p.x_n_off = x_n_off; p.active = (const uint8_t*)active;
  p.w_master = (float*)master; p.w_global = (const float*)global_w; p.p_stride = (size_t)P;
  p.w_off = Cnn2Layout::W1; p.w_fwd = (__nv_bfloat16*)w1_fwd; p.w_dgrad = nullptr;
  p.lr = lr; p.mu = mu; p.grad_out = (float*)grad_out;
  p.dpool = (const __nv_bfloat16*)dp1; p.pcode = (const uint8_t*)code1;
  const int grid = C < g_sms ? C : g_sms;
  CONV1_WGRAD<<<grid, 320, Conv1WgradCfg::SMEM_BYTES, (cudaStream_t)stream>>>(tmX, tmX, p);
  CK(cudaGetLastError());
  return 0;
}

int olsb_head(int C, const void* active, const void* p2, const void* code2, const void* labels, int label_stride,
              int label_off, void* master, const void* global_w, long long P, float lr, float mu, void* dy2,
              void* loss_sum, void* nonfinite, void* logits_out, int apply_update, void* dwfc_out, void* db2_out,
              void* stream) {
  if (int r = olsb_init()) return r;
  flk::HeadParams p;
  memset(&p, 0, sizeof(p));
  p.num_clients = C; p.active = (const uint8_t*)active; p.p2 = (const __nv_bfloat16*)p2;
  p.code2 = (const uint8_t*)code2; p.labels = (const int32_t*)labels; p.label_stride = label_stride;
  p.label_off = label_off; p.master = (float*)master; p.global_w = (const float*)global_w; p.P = (int)P;
  p.lr = lr; p.mu = mu; p.dy2 = (__nv_bfloat16*)dy2; p.loss_sum = (float*)loss_sum;
  p.nonfinite = (int32_t*)nonfinite; p.logits_out = (float*)logits_out; p.apply_update = apply_update;
  p.dwfc_out = (float*)dwfc_out; p.db2_out = (float*)db2_out;
  const int grid = C < 2 * g_sms ? C : 2 * g_sms;
  flk::head_fc_xent_kernel<<<grid, flk::kHeadThreads, flk::kHeadSmem, (cudaStream_t)stream>>>(p);
  CK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// One full local SGD step of the CNN-2 family for all resident clients (6 launches, no host sync).
struct OlsbCnn2Step {
  const void* x5; long long x_total; int x_n_stride; int x_n_off;
  const void* labels; int label_stride; int label_off;
  void* master; const void* global_w; long long P; int C;
  const void* active;
  void* w1_fwd; void* w2_fwd; void* w2_dgrad;
  void* p1; void* code1; void* p2; void* code2; void* dy2; void* dp1;
  void* loss_sum; void* nonfinite;
  float lr, mu; int use_prox; int nsplit;
};

int olsb_cnn2_local_step(const OlsbCnn2Step* s, void* stream) {
  const void* gw = s->use_prox ? s->global_w : nullptr;
  int r;
  if ((r = olsb_conv1_fwd(s->x5, s->x_total, s->x_n_stride, s->x_n_off, s->w1_fwd, s->master, s->P, s->C, s->active,
                          s->p1, s->code1, s->nsplit, stream))) return r;
  if ((r = olsb_conv2_fwd(s->p1, s->w2_fwd, s->master, s->P, s->C, s->active, s->p2, s->code2, s->nsplit, stream)))
    return r;
  if ((r = olsb_head(s->C, s->active, s->p2, s->code2, s->labels, s->label_stride, s->label_off, s->master, gw, s->P,
                     s->lr, s->mu, s->dy2, s->loss_sum, s->nonfinite, nullptr, 1, nullptr, nullptr, stream)))
    return r;
  if ((r = olsb_conv2_dgrad(s->dy2, s->w2_fwd, s->master, gw, s->P, s->C, s->active, s->p1, s->dp1, s->lr, s->mu,
                            nullptr, s->nsplit, stream))) return r;
  if ((r = olsb_conv2_wgrad(s->p1, s->dy2, s->master, gw, s->P, s->C, s->active, s->w2_fwd, s->w2_dgrad, s->lr, s->mu,
                            nullptr, stream))) return r;
  if ((r = olsb_conv1_wgrad(s->x5, s->x_total, s->x_n_stride, s->x_n_off, s->dp1, s->code1, s->master, gw, s->P,
                            s->C, s->active, s->w1_fwd, s->lr, s->mu, nullptr, stream))) return r;
  return 0;
}

// ------------------------------------------------------------------------------------------------
int olsb_bcast_scatter_cnn2(const void* const* global_peers, int nranks, int rank, int shard_len, void* global_local,
                            void* master, int C, const void* participate, void* w1_fwd, void* w2_fwd, void* w2_dgrad,
                            void* stream) {
  if (int r = olsb_init()) return r;
  flk::BcastParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < nranks && i < flk::kMaxPeers; ++i) p.global_peers.p[i] = const_cast<void*>(global_peers[i]);
  p.nranks = nranks; p.rank = rank; p.shard_len = shard_len; p.global_local = (float*)global_local;
  p.master = (float*)master; p.num_clients = C; p.participate = (const uint8_t*)participate;
  p.w1_fwd = (__nv_bfloat16*)w1_fwd; p.w2_fwd = (__nv_bfloat16*)w2_fwd; p.w2_dgrad = (__nv_bfloat16*)w2_dgrad;
  p.P = Cnn2Layout::P;
  const int th = 256;
  const int gx = (Cnn2Layout::P / 4 + th - 1) / th;
  int gy = (4 * g_sms + gx - 1) / gx;
  if (gy > C) gy = C;
  if (gy < 1) gy = 1;
  flk::bcast_scatter_cnn2_kernel<<<dim3(gx, gy), th, 0, (cudaStream_t)stream>>>(p);
  CK(cudaGetLastError());
  return 0;
}

int olsb_fedavg_partial(const void* master, const void* global_w, const void* weight, const void* nonfinite, int C,
                        int P, int groups, void* partial, void* wsum, void* stream) {
  if (int r = olsb_init()) return r;
  flk::PartialParams p;
  p.master = (const float*)master; p.global_w = (const float*)global_w; p.weight = (const float*)weight;
  p.nonfinite = (const int32_t*)nonfinite; p.num_clients = C; p.P = P; p.groups = groups;
  p.partial = (float*)partial; p.wsum = (float*)wsum;
  const int th = 128;
  const int gx = (P / 4 + th - 1) / th;
  flk::fedavg_partial_kernel<<<dim3(gx, groups), th, 0, (cudaStream_t)stream>>>(p);
  CK(cudaGetLastError());
  return 0;
}

int olsb_fedavg_finish(const void* const* partial_peers, const void* const* wsum_peers, const void* const* flag_peers,
                       int nranks, int rank, int groups, int P, int shard_len, unsigned epoch, void* global_w,
                       void* opt_m, void* opt_v, int opt, float server_lr, float beta1, float beta2, float eps, int step,
                       void* stats, void* arrive_ctr, const void* const* round_peers, void* round_out, int round_len,
                       const void* partial_mc, void* global_mc, void* stream) {
  if (int r = olsb_init()) return r;
  flk::FinishParams p;
  memset(&p, 0, sizeof(p));
  p.partial_mc = (const float*)partial_mc; p.global_mc = (float*)global_mc;
  p.spin_limit = g_peer_spin_limit;
  for (int i = 0; i < nranks && i < flk::kMaxPeers; ++i) {
    p.partial_peers.p[i] = const_cast<void*>(partial_peers[i]);
    p.wsum_peers.p[i] = const_cast<void*>(wsum_peers[i]);
    p.flags.p[i] = flag_peers ? const_cast<void*>(flag_peers[i]) : nullptr;
  }
  p.nranks = nranks; p.rank = rank; p.groups = groups; p.P = P; p.shard_len = shard_len; p.epoch = epoch;
  p.global_w = (float*)global_w; p.opt_m = (float*)opt_m; p.opt_v = (float*)opt_v; p.opt = opt;
  p.server_lr = server_lr; p.beta1 = beta1; p.beta2 = beta2; p.eps = eps; p.step = step;
  p.stats = (float*)stats; p.arrive_ctr = (uint32_t*)arrive_ctr;
  if (round_peers && round_out) {
    for (int i = 0; i < nranks && i < flk::kMaxPeers; ++i) p.round_peers.p[i] = const_cast<void*>(round_peers[i]);
    p.round_out = (float*)round_out;
    p.round_len = round_len > 0 ? (round_len < flk::kRoundRec ? round_len : flk::kRoundRec) : 8;
  }
  CK(cudaMemsetAsync(arrive_ctr, 0, 2 * sizeof(uint32_t), (cudaStream_t)stream));
  CK(cudaMemsetAsync(stats, 0, 4 * sizeof(float), (cudaStream_t)stream));
  const int th = 256;
  int grid = (shard_len / 4 + th - 1) / th;
  if (grid > g_sms) grid = g_sms;   // all CTAs must be co-resident for the in-kernel gates
  if (grid < 1) grid = 1;
  flk::fedavg_finish_kernel<<<grid, th, 0, (cudaStream_t)stream>>>(p);
  CK(cudaGetLastError());
  return 0;
}

// runtime switches (kernel variants under evaluation); returns the previous value, -1 for an unknown name
int olsb_set_option(const char* name, int value) {
  if (!strcmp(name, "x5_rows")) { const int o = g_x5_rows; g_x5_rows = value; return o; }
  if (!strcmp(name, "dgrad_pair")) { const int o = g_dgrad_pair; g_dgrad_pair = value; return o; }
  if (!strcmp(name, "fwd2_shift")) { const int o = g_fwd2_shift; g_fwd2_shift = value; return o; }
  if (!strcmp(name, "wgrad2_ws")) { const int o = g_wgrad2_ws; g_wgrad2_ws = value; return o; }
  if (!strcmp(name, "wgrad2_mode")) { const int o = g_wgrad2_mode; g_wgrad2_mode = value; return o; }
  if (!strcmp(name, "peer_spin_limit")) { const int o = (int)(g_peer_spin_limit >> 8); g_peer_spin_limit = (unsigned)value << 8; return o; }
  return -1;
}

int olsb_round_stats(const void* loss_sum, const void* nonfinite, const void* participate, const void* step_active,
                     const void* counters, int C, int S, float inv_steps, void* out, const void* class_id, int nclass,
                     void* stream) {
  flk::round_stats_kernel<<<1, 512, 0, (cudaStream_t)stream>>>((const float*)loss_sum, (const int*)nonfinite,
                                                               (const uint8_t*)participate, (const uint8_t*)step_active,
                                                               (const int*)counters, C, S, inv_steps, (float*)out,
                                                               (const int*)class_id, nclass);
  CK(cudaGetLastError());
  return 0;
}

int olsb_deviceflow_mask(int C, int client_base, unsigned long long seed, unsigned round, const void* p_online,
                         float drop_probability, const void* arrive_tick, int max_tick, const void* n_samples,
                         const void* local_steps, int max_steps, void* participate, void* step_active,
                         void* agg_weight, void* counters, void* stream) {
  flk::MaskParams p;
  p.num_clients = C; p.client_base = client_base; p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32);
  p.round = round; p.p_online = (const float*)p_online; p.drop_probability = drop_probability;
  p.arrive_tick = (const int32_t*)arrive_tick; p.max_tick = max_tick; p.n_samples = (const int32_t*)n_samples;
  p.local_steps = (const int32_t*)local_steps; p.max_steps = max_steps; p.participate = (uint8_t*)participate;
  p.step_active = (uint8_t*)step_active; p.agg_weight = (float*)agg_weight; p.counters = (int32_t*)counters;
  CK(cudaMemsetAsync(counters, 0, 4 * sizeof(int32_t), (cudaStream_t)stream));
  flk::deviceflow_mask_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p);
  CK(cudaGetLastError());
  return 0;
}

// flow-dispatch ticks of this rank's clients, computed on the device (see fl_kernels.cuh: flow_*_kernel)
int olsb_flow_ticks(int N, int c0, int C, unsigned long long seed, unsigned round, const void* p_online_all, float scale,
                    void* keys, void* hist, void* start, void* cursor, void* items, const void* cum, int nslots,
                    const void* dropmap, void* tick, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(hist, 0, flk::kFlowBuckets * sizeof(int), st));
  const int th = 256, gn = (N + th - 1) / th;
  flk::flow_keys_kernel<<<gn, th, 0, st>>>(N, (uint32_t)seed, (uint32_t)(seed >> 32), round, (const float*)p_online_all, scale,
                                            (uint32_t*)keys, (int*)hist);
  flk::flow_scan_kernel<<<1, 1024, 0, st>>>((const int*)hist, (int*)start, (int*)cursor);
  flk::flow_scatter_kernel<<<gn, th, 0, st>>>(N, (const uint32_t*)keys, (const int*)start, (int*)cursor, (int*)items);
  flk::flow_ticks_kernel<<<(C + th - 1) / th, th, 0, st>>>(c0, C, (const uint32_t*)keys, (const int*)start, (const int*)items,
                                                           (const int*)cum, nslots, (const uint8_t*)dropmap, (int32_t*)tick);
  CK(cudaGetLastError());
  return 0;
}

int olsb_rate_curve_area(const void* y, void* area, int bins, void* stream) {
  flk::rate_curve_area_kernel<<<(bins + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const double*)y, (double*)area,
                                                                                      bins);
  CK(cudaGetLastError());
  return 0;
}

int olsb_client_sgd(void* w, const void* g, const void* wg, const void* active, int C, int P, float lr, float mu,
                    float momentum, void* mom, void* stream) {
  const int th = 256;
  flk::client_sgd_kernel<<<dim3((P / 4 + th - 1) / th, C), th, 0, (cudaStream_t)stream>>>(
      (float*)w, (const float*)g, (const float*)wg, (const uint8_t*)active, C, P, lr, mu, momentum, (float*)mom);
  CK(cudaGetLastError());
  return 0;
}

// host-side Philox (parity tests)
void olsb_philox_host(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned* out) {
  flk::philox4x32_10(c0, c1, c2, c3, k0, k1, out);
}

// ------------------------------------------------------------------------------------------------
// Symmetric heap primitives: plain cudaMalloc slabs (torch's caching allocator may hand out blocks that are
// not exportable) shared between the one-process-per-GPU ranks through CUDA IPC handles.  The 64-byte handles
// travel over the bootstrap process group; after olsb_ipc_open every rank holds a device pointer into every
// peer's slab and the kernels above dereference those pointers directly over NVLink.
int olsb_sym_alloc(unsigned long long bytes, void** out) {
  void* p = nullptr;
  CK(cudaMalloc(&p, (size_t)bytes));
  CK(cudaMemset(p, 0, (size_t)bytes));
  *out = p;
  return 0;
}
int olsb_sym_free(void* p) { CK(cudaFree(p)); return 0; }
int olsb_ipc_get(void* p, unsigned char* handle64) {
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, p));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return 0;
}
int olsb_ipc_open(const unsigned char* handle64, void** out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *out = p;
  return 0;
}
int olsb_ipc_close(void* p) { CK(cudaIpcCloseMemHandle(p)); return 0; }
int olsb_can_access_peer(int dev, int peer) {
  int ok = 0;
  if (cudaDeviceCanAccessPeer(&ok, dev, peer) != cudaSuccess) return 0;
  return ok;
}
// pinned host staging buffers for the H2D input feeder / D2H metric read-back
int olsb_host_alloc(unsigned long long bytes, void** out) { CK(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault)); return 0; }
int olsb_host_free(void* p) { CK(cudaFreeHost(p)); return 0; }
int olsb_memcpy_h2d_async(void* dst, const void* src, unsigned long long bytes, void* stream) {
  CK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Generic per-client GEMM family (fully-connected layers).  See linear_tc.cuh.
struct OlsbOperand {            // bf16 tensor [C][rows][inner] with arbitrary strides (elements)
  const void* base; long long client_stride; long long row_stride; int rows; int inner;
};
struct OlsbGemm {
  int variant;                  // 0 forward (bias+act, transposed store), 1 dgrad (ReLU gate + bias SGD), 2 wgrad (+SGD)
  int C, M, N, K;
  OlsbOperand a, b;
  const void* active; float lr, mu;
  void* out_t; long long out_stride; int out_ld; const void* bias; long long bias_stride; int relu;
  const void* mask_t; void* b_master; long long b_stride; const void* b_global; void* db_out;
  void* w_master; long long w_stride; int w_ld; const void* w_global; void* w_bf16; long long wb_stride;
  void* grad_out; long long g_stride;
};

namespace {
int make_operand_map(CUtensorMap* tm, const OlsbOperand& o, int C, int box_rows) {
  const uint64_t dims[3] = {(uint64_t)o.inner, (uint64_t)o.rows, (uint64_t)C};
  const uint64_t str[2] = {(uint64_t)o.row_stride, (uint64_t)o.client_stride};
  const uint32_t box[3] = {64u, (uint32_t)box_rows, 1u};
  return make_map(tm, o.base, 3, dims, str, box, 128);
}
template <bool A_MN, bool B_MN, int N_TILE, int KC, int EPI, int NSTAGE>
int launch_gemm(const OlsbGemm* g, cudaStream_t stream) {
  using Cfg = lintc::GemmCfg<A_MN, B_MN, N_TILE, KC, NSTAGE>;
  auto kern = lintc::client_gemm_kernel<A_MN, B_MN, N_TILE, KC, EPI, NSTAGE>;
  static bool attr = false;
  if (!attr) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)); attr = true; }
  CUtensorMap tmA, tmB;
  if (int r = make_operand_map(&tmA, g->a, g->C, A_MN ? KC : 128)) return r;
  if (int r = make_operand_map(&tmB, g->b, g->C, B_MN ? KC : N_TILE)) return r;
  lintc::GemmParams p;
  memset(&p, 0, sizeof(p));
  p.num_clients = g->C; p.M = g->M; p.N = g->N; p.K = g->K;
  p.m_tiles = (g->M + 127) / 128; p.n_tiles = (g->N + N_TILE - 1) / N_TILE; p.k_chunks = (g->K + KC - 1) / KC;
  p.active = (const uint8_t*)g->active; p.lr = g->lr; p.mu = g->mu;
  p.out_t = (__nv_bfloat16*)g->out_t; p.out_stride = g->out_stride; p.out_ld = g->out_ld;
  p.bias = (const float*)g->bias; p.bias_stride = g->bias_stride; p.relu = g->relu;
  p.mask_t = (const __nv_bfloat16*)g->mask_t; p.b_master = (float*)g->b_master; p.b_stride = g->b_stride;
  p.b_global = (const float*)g->b_global; p.db_out = (float*)g->db_out;
  p.w_master = (float*)g->w_master; p.w_stride = g->w_stride; p.w_ld = g->w_ld; p.w_global = (const float*)g->w_global;
  p.w_bf16 = (__nv_bfloat16*)g->w_bf16; p.wb_stride = g->wb_stride; p.grad_out = (float*)g->grad_out; p.g_stride = g->g_stride;
  if (EPI == lintc::EPI_MASK_T && p.n_tiles != 1) { snprintf(g_err, sizeof(g_err), "dgrad epilogue needs the batch in one N tile"); return -4; }
  const int work = g->C * p.m_tiles * p.n_tiles;
  const int grid = work < g_sms ? work : g_sms;
  kern<<<grid, lintc::kThreads, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  CK(cudaGetLastError());
  return 0;
}
}  // namespace

extern "C" {
int olsb_client_gemm(const OlsbGemm* g, void* stream) {
  if (int r = olsb_init()) return r;
  cudaStream_t st = (cudaStream_t)stream;
  switch (g->variant) {
    case 0: return launch_gemm<false, false, 32, 64, lintc::EPI_BIAS_ACT_T, 6>(g, st);
    case 1: return launch_gemm<true, false, 32, 64, lintc::EPI_MASK_T, 6>(g, st);
    case 2: return launch_gemm<true, true, 256, 32, lintc::EPI_SGD, 6>(g, st);
    default: snprintf(g_err, sizeof(g_err), "unknown gemm variant %d", g->variant); return -5;
  }
}

int olsb_client_xent(int C, int batch, int classes, int ld, const void* active, const void* logits, const void* labels,
                     int label_stride, int label_off, void* dlogits, void* loss_sum, void* nonfinite, void* b_master,
                     long long b_stride, const void* b_global, float lr, float mu, void* db_out, void* stream) {
  lintc::XentParams p;
  p.num_clients = C; p.batch = batch; p.classes = classes; p.ld = ld; p.active = (const uint8_t*)active;
  p.logits = (const __nv_bfloat16*)logits; p.labels = (const int32_t*)labels; p.label_stride = label_stride;
  p.label_off = label_off; p.dlogits = (__nv_bfloat16*)dlogits; p.loss_sum = (float*)loss_sum;
  p.nonfinite = (int32_t*)nonfinite; p.b_master = (float*)b_master; p.b_stride = b_stride;
  p.b_global = (const float*)b_global; p.lr = lr; p.mu = mu; p.db_out = (float*)db_out;
  if (batch > 32 || classes > 16) { snprintf(g_err, sizeof(g_err), "xent kernel supports batch<=32, classes<=16"); return -6; }
  const int wpb = 4;
  lintc::client_xent_kernel<<<(C + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(p);
  CK(cudaGetLastError());
  return 0;
}

int olsb_bcast_scatter_generic(const void* const* global_peers, int nranks, int shard_len, int P, void* global_local,
                               void* master, void* master_bf16, int C, const void* participate, int P_bf16,
                               long long bf16_stride, void* stream) {
  if (int r = olsb_init()) return r;
  lintc::BcastGenericParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < nranks && i < 8; ++i) p.global_peers[i] = const_cast<void*>(global_peers[i]);
  p.nranks = nranks; p.shard_len = shard_len; p.P = P; p.global_local = (float*)global_local; p.master = (float*)master;
  p.master_bf16 = (__nv_bfloat16*)master_bf16; p.num_clients = C; p.participate = (const uint8_t*)participate;
  p.P_bf16 = P_bf16 > 0 ? P_bf16 : P; p.bf16_stride = bf16_stride > 0 ? bf16_stride : P;
  const int th = 256, gx = (P / 4 + th - 1) / th;
  int gy = (4 * g_sms + gx - 1) / gx; if (gy > C) gy = C; if (gy < 1) gy = 1;
  lintc::bcast_scatter_generic_kernel<<<dim3(gx, gy), th, 0, (cudaStream_t)stream>>>(p);
  CK(cudaGetLastError());
  return 0;
}
}  // extern "C"
