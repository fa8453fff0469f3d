// api_nn.cu — C ABI of the batched GEMM / implicit-conv family (bgemm_tc.cuh) and the neural-net support kernels
// (nn_kernels.cuh) used by the ResNet-18 and BERT-mini local steps.  Same conventions as api.cu: raw device pointers,
// the caller's stream, 0 on success, olsb_nn_last_error() for the text.  Parameter blocks are plain structs that
// ops/bgemm.py mirrors with ctypes; olsb_nn_sizes() lets the python side verify the mirror at import time.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "bgemm_tc.cuh"
#include "nn_kernels.cuh"

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
}  // namespace

// bf16 tensor map description: innermost dimension first, strides in elements (dims[0] is contiguous)
struct OlsbTmap {
  const void* base; int rank;
  unsigned long long dims[5]; unsigned long long strides[4]; unsigned box[5]; unsigned estr[5];
};
struct OlsbBgemm {
  int am, bm, n_tile, epi;
  OlsbTmap a, b;
  bg::Params p;
};

namespace {
int encode(CUtensorMap* tm, const OlsbTmap& d) {
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < d.rank; ++i) { gd[i] = d.dims[i]; bx[i] = d.box[i]; es[i] = d.estr[i] ? d.estr[i] : 1; }
  for (int i = 0; i < d.rank - 1; ++i) gs[i] = d.strides[i] * 2;
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d.rank, const_cast<void*>(d.base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err),
             "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims %llu %llu %llu %llu %llu strides %llu %llu %llu %llu "
             "box %u %u %u %u %u", (int)r, d.rank, d.dims[0], d.dims[1], d.dims[2], d.dims[3], d.dims[4], d.strides[0],
             d.strides[1], d.strides[2], d.strides[3], d.box[0], d.box[1], d.box[2], d.box[3], d.box[4]);
    return -3;
  }
  return 0;
}

template <int AM, int BM, int N_TILE, int EPI, int NSTAGE>
int launch(const OlsbBgemm* g, cudaStream_t stream) {
  using C = bg::Cfg<AM, BM, N_TILE, NSTAGE>;
  auto kern = bg::bgemm_kernel<AM, BM, N_TILE, EPI, NSTAGE>;
  static bool attr = false;
  if (!attr) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES)); attr = true; }
  CUtensorMap tmA, tmB;
  if (int r = encode(&tmA, g->a)) return r;
  if (int r = encode(&tmB, g->b)) return r;
  bg::Params p = g->p;
  if (p.act_div < 1) p.act_div = 1;
  if (p.o_bdiv < 1) p.o_bdiv = 1;
  if (p.bias_div < 1) p.bias_div = 1;
  if (p.cchunks < 1) p.cchunks = 1;
  if (p.kb_cpi < 1) p.kb_cpi = 1;
  if (p.src_cimgs < 1) p.src_cimgs = p.imgs;
  const long long work = (long long)p.batches * p.m_tiles * p.n_tiles;
  if (work <= 0) return 0;
  const int grid = work < g_sms ? (int)work : g_sms;
  kern<<<grid, bg::kThreads, C::SMEM_BYTES, stream>>>(tmA, tmB, p);
  CK(cudaGetLastError());
  return 0;
}
}  // namespace

extern "C" {

const char* olsb_nn_last_error() { return g_err; }

// sizes / offsets the ctypes mirror is checked against
int olsb_nn_sizes(int* out, int n) {
  const int v[] = {(int)sizeof(OlsbTmap), (int)sizeof(bg::Params), (int)sizeof(OlsbBgemm), (int)offsetof(OlsbBgemm, p),
                   (int)offsetof(bg::Params, out), (int)offsetof(bg::Params, w_master), (int)offsetof(bg::Params, lr),
                   (int)sizeof(nnk::GnApplyParams), (int)sizeof(nnk::GnBwdParams), (int)sizeof(nnk::PoolHeadParams),
                   (int)sizeof(nnk::LnParams), (int)sizeof(nnk::EmbBwdParams), (int)sizeof(nnk::ClsHeadParams),
                   (int)sizeof(nnk::ColsumParams)};
  const int k = (int)(sizeof(v) / sizeof(v[0]));
  for (int i = 0; i < n && i < k; ++i) out[i] = v[i];
  return k;
}

int olsb_bgemm(const OlsbBgemm* g, void* stream) {
  if (int r = ensure_init()) return r;
  cudaStream_t st = (cudaStream_t)stream;
  using namespace bg;
  const int key = g->am * 1000000 + g->bm * 100000 + g->n_tile * 100 + g->epi;
  switch (key) {
    // ---- plain batched GEMMs (fully-connected layers, attention)
    case A_K * 1000000 + B_K * 100000 + 256 * 100 + E_STORE: return launch<A_K, B_K, 256, E_STORE, 4>(g, st);
    case A_K * 1000000 + B_MN * 100000 + 256 * 100 + E_STORE: return launch<A_K, B_MN, 256, E_STORE, 4>(g, st);
    case A_MN * 1000000 + B_MN * 100000 + 256 * 100 + E_SGD: return launch<A_MN, B_MN, 256, E_SGD, 4>(g, st);
    case A_K * 1000000 + B_K * 100000 + 128 * 100 + E_SOFTMAX: return launch<A_K, B_K, 128, E_SOFTMAX, 6>(g, st);
    case A_K * 1000000 + B_K * 100000 + 128 * 100 + E_DSOFT: return launch<A_K, B_K, 128, E_DSOFT, 6>(g, st);
    case A_K * 1000000 + B_MN * 100000 + 64 * 100 + E_STORE: return launch<A_K, B_MN, 64, E_STORE, 8>(g, st);
    case A_MN * 1000000 + B_MN * 100000 + 64 * 100 + E_STORE: return launch<A_MN, B_MN, 64, E_STORE, 8>(g, st);
    case A_K * 1000000 + B_K * 100000 + 64 * 100 + E_STORE: return launch<A_K, B_K, 64, E_STORE, 8>(g, st);
    // ---- implicit-GEMM convolutions
    case A_CONV * 1000000 + B_K * 100000 + 64 * 100 + E_GNSTAT: return launch<A_CONV, B_K, 64, E_GNSTAT, 8>(g, st);
    case A_CONV * 1000000 + B_K * 100000 + 128 * 100 + E_GNSTAT: return launch<A_CONV, B_K, 128, E_GNSTAT, 6>(g, st);
    case A_CONV * 1000000 + B_K * 100000 + 256 * 100 + E_GNSTAT: return launch<A_CONV, B_K, 256, E_GNSTAT, 4>(g, st);
    case A_CONV * 1000000 + B_WT * 100000 + 64 * 100 + E_STORE: return launch<A_CONV, B_WT, 64, E_STORE, 8>(g, st);
    case A_CONV * 1000000 + B_WT * 100000 + 128 * 100 + E_STORE: return launch<A_CONV, B_WT, 128, E_STORE, 6>(g, st);
    case A_CONV * 1000000 + B_WT * 100000 + 256 * 100 + E_STORE: return launch<A_CONV, B_WT, 256, E_STORE, 4>(g, st);
    // ---- halo-box form (3x3, stride 1, 64 -> 64 channels): resident weights, one box per tile
    case A_HALO * 1000000 + B_K * 100000 + 64 * 100 + E_GNSTAT: return launch<A_HALO, B_K, 64, E_GNSTAT, 5>(g, st);
    case A_HALO * 1000000 + B_WT * 100000 + 64 * 100 + E_STORE: return launch<A_HALO, B_WT, 64, E_STORE, 5>(g, st);
    case A_CONVX * 1000000 + B_MN * 100000 + 128 * 100 + E_SGD_T: return launch<A_CONVX, B_MN, 128, E_SGD_T, 6>(g, st);
    case A_CONVX * 1000000 + B_MN * 100000 + 256 * 100 + E_SGD_T: return launch<A_CONVX, B_MN, 256, E_SGD_T, 4>(g, st);
    case A_HALOW * 1000000 + B_MN * 100000 + 192 * 100 + E_SGD_T: return launch<A_HALOW, B_MN, 192, E_SGD_T, 5>(g, st);
    case A_MN * 1000000 + B_CONVX * 100000 + 192 * 100 + E_SGD: return launch<A_MN, B_CONVX, 192, E_SGD, 5>(g, st);
    default:
      snprintf(g_err, sizeof(g_err), "bgemm: no instantiation for am %d bm %d n_tile %d epi %d", g->am, g->bm, g->n_tile, g->epi);
      return -5;
  }
}

int olsb_stem_im2col(const void* src_u8, void* dst, long long nsamples, void* stream) {
  const size_t npix = (size_t)nsamples * 32 * 32;
  nnk::stem_im2col_kernel<<<(unsigned)((npix + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const uint8_t*)src_u8,
                                                                                              (nnk::bf16*)dst, npix);
  CK(cudaGetLastError());
  return 0;
}

int olsb_gn_apply(const nnk::GnApplyParams* p, void* stream) {
  if (int r = ensure_init()) return r;
  if (p->groups > 8 || p->Ch % (8 * p->groups)) { snprintf(g_err, sizeof(g_err), "gn_apply: groups <= 8 and channels %% (8*groups) == 0"); return -6; }
  const int vpi = p->HW * p->Ch / 8;
  int gx = (vpi + 255) / 256;
  if (gx > 8) gx = 8;
  nnk::gn_apply_kernel<<<dim3(gx, p->imgs, p->C), 256, 0, (cudaStream_t)stream>>>(*p);
  CK(cudaGetLastError());
  return 0;
}

// GroupNorm backward: reduce pass (per-image sums, dgamma / dbeta) then apply pass (dx, optional dres)
static int g_gn_fused = 1;
int olsb_gn_set_fused(int v) { const int o = g_gn_fused; g_gn_fused = v; return o; }

int olsb_gn_bwd(const nnk::GnBwdParams* p, void* stream) {
  if (int r = ensure_init()) return r;
  if (p->groups > 8 || p->Ch > 512 || (256 % (p->Ch / 8))) { snprintf(g_err, sizeof(g_err), "gn_bwd: channels in {64,128,256,512}, groups <= 8"); return -6; }
  if (g_gn_fused) {
    // one kernel, masked gradient parked in shared memory between the reduction and the apply pass
    const int smem = p->HW * p->Ch * 2;
    static int attr = 0;
    if (smem > attr) { CK(cudaFuncSetAttribute(nnk::gn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = smem; }
    if (smem <= 200 * 1024 && (256 % (p->Ch / 8)) == 0) {
      // 128 KB images (one CTA per SM): 512 threads for bytes in flight; smaller images: 256 threads, several CTAs per SM
      const int th = smem > 64 * 1024 ? 512 : 256;
      nnk::gn_bwd_fused_kernel<<<dim3(p->imgs, p->C), th, smem, (cudaStream_t)stream>>>(*p);
      CK(cudaGetLastError());
      return 0;
    }
  }
  nnk::gn_bwd_reduce_kernel<<<dim3(p->imgs, p->C), 256, 0, (cudaStream_t)stream>>>(*p);
  CK(cudaGetLastError());
  const int vpi = p->HW * p->Ch / 8;
  int gx = (vpi + 255) / 256;
  if (gx > 8) gx = 8;
  nnk::gn_bwd_apply_kernel<<<dim3(gx, p->imgs, p->C), 256, 0, (cudaStream_t)stream>>>(*p);
  CK(cudaGetLastError());
  return 0;
}

int olsb_pool_head(const nnk::PoolHeadParams* p, void* stream) {
  if (p->classes > 16) { snprintf(g_err, sizeof(g_err), "pool head: classes <= 16"); return -6; }
  const int smem = (2 * p->classes * p->Ch + p->Ch + 32) * 4;
  static int attr = 0;
  if (smem > attr) { CK(cudaFuncSetAttribute(nnk::pool_fc_xent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = smem; }
  nnk::pool_fc_xent_kernel<<<p->C, 256, smem, (cudaStream_t)stream>>>(*p);
  CK(cudaGetLastError());
  return 0;
}

int olsb_ln_fwd(const nnk::LnParams* p, void* stream) {
  int gx = (p->T + 7) / 8;
  if (gx > 64) gx = 64;
  nnk::ln_fwd_kernel<<<dim3(gx, p->C), 256, 0, (cudaStream_t)stream>>>(*p);
  CK(cudaGetLastError());
  return 0;
}
int olsb_ln_bwd(const nnk::LnParams* p, void* stream) {
  int gx = (p->T + 127) / 128;            // 16 tokens per warp: register accumulators amortise the atomics
  if (gx < 1) gx = 1;
  nnk::ln_bwd_kernel<<<dim3(gx, p->C), 256, 0, (cudaStream_t)stream>>>(*p);
  CK(cudaGetLastError());
  return 0;
}
int olsb_emb_bwd(const nnk::EmbBwdParams* p, void* stream) {
  int gx = (p->T + 63) / 64;
  nnk::emb_bwd_kernel<<<dim3(gx, p->C), 256, 0, (cudaStream_t)stream>>>(*p);
  CK(cudaGetLastError());
  return 0;
}
int olsb_cls_head(const nnk::ClsHeadParams* p, void* stream) {
  if (p->classes > 16 || p->imgs > 32) { snprintf(g_err, sizeof(g_err), "cls head: classes <= 16, batch <= 32"); return -6; }
  // gradient of the last hidden state: zero except token 0 of every sample
  CK(cudaMemsetAsync(p->dy, 0, (size_t)p->C * p->imgs * p->S * nnk::kH * sizeof(nnk::bf16), (cudaStream_t)stream));
  nnk::cls_head_kernel<<<p->C, 256, 0, (cudaStream_t)stream>>>(*p);
  CK(cudaGetLastError());
  return 0;
}
int olsb_colsum(const nnk::ColsumParams* p, void* stream) {
  if (p->N % 8 || p->N / 8 > 256) { snprintf(g_err, sizeof(g_err), "colsum: N %% 8 == 0 and N <= 2048"); return -6; }
  const int tx = p->N / 8;
  int ty = 256 / tx; if (ty < 1) ty = 1; if (ty > 8) ty = 8;
  nnk::ColsumParams q = *p;
  if (q.chunk < 1) q.chunk = 256;
  const int gx = (q.T + q.chunk - 1) / q.chunk;
  nnk::colsum_kernel<<<dim3(gx, q.C), dim3(tx, ty), (size_t)ty * q.N * 4, (cudaStream_t)stream>>>(q);
  CK(cudaGetLastError());
  return 0;
}
int olsb_slice_sgd(void* master, long long P, int off, int n, const void* g, long long gs_stride, const void* global_w,
                   const void* active, int C, float lr, float mu, void* stream) {
  if (n <= 0) return 0;
  nnk::slice_sgd_kernel<<<dim3((n + 255) / 256, C), 256, 0, (cudaStream_t)stream>>>(
      (float*)master, P, off, n, (const float*)g, gs_stride, (const float*)global_w, (const uint8_t*)active, lr, mu);
  CK(cudaGetLastError());
  return 0;
}
int olsb_memset_async(void* ptr, int value, unsigned long long bytes, void* stream) {
  CK(cudaMemsetAsync(ptr, value, (size_t)bytes, (cudaStream_t)stream));
  return 0;
}

}  // extern "C"
