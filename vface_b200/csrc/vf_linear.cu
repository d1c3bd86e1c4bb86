// Projection with the residual add and the bias inside the library GEMM (vf_linear_residual).
//
//   out[r, :] = residual[r, :] + x[r, :] . W^T + bias        W (n, k) row-major (nn.Linear.weight), bias (n) or NULL
//
// Replaces  `x = ff(norm3(x)) + x`  (ldm/modules/attention.py:242, the feed-forward down-projection + residual) and
// `return x + x_in` after proj_out (attention.py:287-288) on the product path: cuBLASLt computes beta * C in its fp32
// epilogue, so the projection is never rounded to bf16 and written out just to be read back by an add kernel.  Per
// call that removes one write + one read of a (rows, n) tensor and one of the two bf16 roundings of the residual
// stream (measured on one 96-sample step: 0.343 -> 0.252 ms for 1280 -> 320 at 64x64, 0.230 -> 0.120 ms for 320 -> 320).
// A plain library GEMM (projections stay on cuBLAS); the caller owns the workspace.
#include "vf_common.cuh"

#include <cublasLt.h>
#include <dlfcn.h>
#include <map>
#include <mutex>
#include <tuple>

namespace vf {

// cuBLASLt is bound at RUN time, not link time: the host process (PyTorch) ships its own libcublas / libcublasLt
// pair, and a DT_NEEDED entry here would let the dynamic loader pull the SYSTEM libcublasLt in first whenever this
// library is loaded before torch -- torch's libcublas 12.8 then runs against a 12.9 libcublasLt and its first
// cublasGemmEx fails with CUBLAS_STATUS_INVALID_VALUE (seen in __graft_entry__: build() then smoke() in one
// process).  So: take the copy that is already loaded (RTLD_NOLOAD), and only load one ourselves if there is none.
struct LtApi {
  decltype(&cublasLtCreate) Create;
  decltype(&cublasLtMatmulDescCreate) MatmulDescCreate;
  decltype(&cublasLtMatmulDescSetAttribute) MatmulDescSetAttribute;
  decltype(&cublasLtMatrixLayoutCreate) MatrixLayoutCreate;
  decltype(&cublasLtMatrixLayoutSetAttribute) MatrixLayoutSetAttribute;
  decltype(&cublasLtMatrixLayoutDestroy) MatrixLayoutDestroy;
  decltype(&cublasLtMatmulPreferenceCreate) MatmulPreferenceCreate;
  decltype(&cublasLtMatmulPreferenceSetAttribute) MatmulPreferenceSetAttribute;
  decltype(&cublasLtMatmulPreferenceDestroy) MatmulPreferenceDestroy;
  decltype(&cublasLtMatmulAlgoGetHeuristic) MatmulAlgoGetHeuristic;
  decltype(&cublasLtMatmul) Matmul;
  bool ok = false;
};
static LtApi g_api;

static int load_lt_api() {           // call with g_lt_mutex held
  if (g_api.ok) return 0;
  void* h = dlopen("libcublasLt.so.12", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libcublasLt.so.12", RTLD_NOW | RTLD_LOCAL);
  if (!h) return fail("vf_linear_residual: libcublasLt.so.12 is not loaded and cannot be loaded (%s)", dlerror());
#define VF_LT_SYM(field, name)                                                            \
  g_api.field = reinterpret_cast<decltype(g_api.field)>(dlsym(h, #name));                 \
  if (!g_api.field) return fail("vf_linear_residual: symbol %s not found in libcublasLt.so.12", #name);
  VF_LT_SYM(Create, cublasLtCreate)
  VF_LT_SYM(MatmulDescCreate, cublasLtMatmulDescCreate)
  VF_LT_SYM(MatmulDescSetAttribute, cublasLtMatmulDescSetAttribute)
  VF_LT_SYM(MatrixLayoutCreate, cublasLtMatrixLayoutCreate)
  VF_LT_SYM(MatrixLayoutSetAttribute, cublasLtMatrixLayoutSetAttribute)
  VF_LT_SYM(MatrixLayoutDestroy, cublasLtMatrixLayoutDestroy)
  VF_LT_SYM(MatmulPreferenceCreate, cublasLtMatmulPreferenceCreate)
  VF_LT_SYM(MatmulPreferenceSetAttribute, cublasLtMatmulPreferenceSetAttribute)
  VF_LT_SYM(MatmulPreferenceDestroy, cublasLtMatmulPreferenceDestroy)
  VF_LT_SYM(MatmulAlgoGetHeuristic, cublasLtMatmulAlgoGetHeuristic)
  VF_LT_SYM(Matmul, cublasLtMatmul)
#undef VF_LT_SYM
  g_api.ok = true;
  return 0;
}

struct LtPlan {
  cublasLtMatmulDesc_t op = nullptr;
  cublasLtMatrixLayout_t a = nullptr, b = nullptr, c = nullptr;
  cublasLtMatmulAlgo_t algo;
  size_t ws = 0;
};

static cublasLtHandle_t g_lt = nullptr;
static std::mutex g_lt_mutex;
static std::map<std::tuple<long long, int, int, long long, long long, long long, int, int, int>, LtPlan> g_plans;

static int lt_fail(cublasStatus_t s, const char* what) { return fail("vf_linear_residual: %s failed with cublasStatus %d", what, (int)s); }

#define VF_LT_TRY(expr)                                                    \
  do {                                                                     \
    cublasStatus_t _s = (expr);                                            \
    if (_s != CUBLAS_STATUS_SUCCESS) return lt_fail(_s, #expr);            \
  } while (0)

static int set_batch(cublasLtMatrixLayout_t l, int batch, long long stride) {
  const int32_t b = batch;
  const int64_t st = stride;
  VF_LT_TRY(g_api.MatrixLayoutSetAttribute(l, CUBLASLT_MATRIX_LAYOUT_BATCH_COUNT, &b, sizeof(b)));
  VF_LT_TRY(g_api.MatrixLayoutSetAttribute(l, CUBLASLT_MATRIX_LAYOUT_STRIDED_BATCH_OFFSET, &st, sizeof(st)));
  return 0;
}

// batch > 1: `rows` rows per batch entry, entries contiguous (stride rows * ld), W shared, bias (batch, n).
static int make_plan(LtPlan& P, long long rows, int k, int n, long long ld_x, long long ld_res, long long ld_out, int dtype,
                     bool has_bias, size_t ws_bytes, int batch, uint32_t align) {
  const cudaDataType_t dt = dtype == VF_BF16 ? CUDA_R_16BF : CUDA_R_32F;
  VF_LT_TRY(g_api.MatmulDescCreate(&P.op, CUBLAS_COMPUTE_32F, CUDA_R_32F));
  // row-major out (rows, n) = x (rows, k) . W^T  <=>  column-major out^T (n, rows) = W (k, n)^T . x^T (k, rows)
  const cublasOperation_t ta = CUBLAS_OP_T, tb = CUBLAS_OP_N;
  VF_LT_TRY(g_api.MatmulDescSetAttribute(P.op, CUBLASLT_MATMUL_DESC_TRANSA, &ta, sizeof(ta)));
  VF_LT_TRY(g_api.MatmulDescSetAttribute(P.op, CUBLASLT_MATMUL_DESC_TRANSB, &tb, sizeof(tb)));
  if (has_bias) {
    const cublasLtEpilogue_t ep = CUBLASLT_EPILOGUE_BIAS;
    VF_LT_TRY(g_api.MatmulDescSetAttribute(P.op, CUBLASLT_MATMUL_DESC_EPILOGUE, &ep, sizeof(ep)));
    VF_LT_TRY(g_api.MatmulDescSetAttribute(P.op, CUBLASLT_MATMUL_DESC_BIAS_DATA_TYPE, &dt, sizeof(dt)));
  }
  VF_LT_TRY(g_api.MatrixLayoutCreate(&P.a, dt, (uint64_t)k, (uint64_t)n, (int64_t)k));
  VF_LT_TRY(g_api.MatrixLayoutCreate(&P.b, dt, (uint64_t)k, (uint64_t)rows, (int64_t)ld_x));
  VF_LT_TRY(g_api.MatrixLayoutCreate(&P.c, dt, (uint64_t)n, (uint64_t)rows, (int64_t)ld_res));
  cublasLtMatrixLayout_t d = nullptr;
  VF_LT_TRY(g_api.MatrixLayoutCreate(&d, dt, (uint64_t)n, (uint64_t)rows, (int64_t)ld_out));
  if (batch > 1) {
    if (int rc = set_batch(P.a, batch, 0)) return rc;
    if (int rc = set_batch(P.b, batch, rows * ld_x)) return rc;
    if (int rc = set_batch(P.c, batch, rows * ld_res)) return rc;
    if (int rc = set_batch(d, batch, rows * ld_out)) return rc;
    if (has_bias) {
      const int64_t bs = n;
      VF_LT_TRY(g_api.MatmulDescSetAttribute(P.op, CUBLASLT_MATMUL_DESC_BIAS_BATCH_STRIDE, &bs, sizeof(bs)));
    }
  }
  cublasLtMatmulPreference_t pref = nullptr;
  VF_LT_TRY(g_api.MatmulPreferenceCreate(&pref));
  VF_LT_TRY(g_api.MatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws_bytes, sizeof(ws_bytes)));
  // the heuristic assumes 256-byte aligned operands unless told otherwise; the entry point only asks for 16 bytes, so
  // the alignment class of THIS call's pointers/strides (part of the plan key) bounds what the algorithm may assume
  VF_LT_TRY(g_api.MatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_A_BYTES, &align, sizeof(align)));
  VF_LT_TRY(g_api.MatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_B_BYTES, &align, sizeof(align)));
  VF_LT_TRY(g_api.MatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_C_BYTES, &align, sizeof(align)));
  VF_LT_TRY(g_api.MatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_D_BYTES, &align, sizeof(align)));
  cublasLtMatmulHeuristicResult_t res;
  int found = 0;
  cublasStatus_t s = g_api.MatmulAlgoGetHeuristic(g_lt, P.op, P.a, P.b, P.c, d, pref, 1, &res, &found);
  g_api.MatmulPreferenceDestroy(pref);
  g_api.MatrixLayoutDestroy(d);
  if (s != CUBLAS_STATUS_SUCCESS || found == 0)
    return fail("vf_linear_residual: no cuBLASLt algorithm for rows=%lld k=%d n=%d (status %d)", rows, k, n, (int)s);
  P.algo = res.algo;
  P.ws = res.workspaceSize;
  return 0;
}

}  // namespace vf

static int linear_residual_impl(const void* x, const void* w, const void* bias, const void* residual, void* out, int batch,
                                long long rows, int k, int n, long long ld_x, long long ld_res, long long ld_out,
                                void* workspace, long long workspace_bytes, int dtype, void* stream) {
  using namespace vf;
  if (int rc = check_device()) return rc;
  if (!x || !w || !residual || !out) return fail("vf_linear_residual: null pointer");
  if (dtype != VF_BF16 && dtype != VF_F32) return fail("vf_linear_residual: bad dtype %d", dtype);
  if (rows <= 0 || k <= 0 || n <= 0 || batch < 1) return fail("vf_linear_residual: bad shape batch=%d rows=%lld k=%d n=%d", batch, rows, k, n);
  if (ld_x < k || ld_res < n || ld_out < n) return fail("vf_linear_residual: row strides smaller than the row length");
  const void* ptrs[4] = {x, w, residual, out};
  for (const void* p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & 15) return fail("vf_linear_residual: pointers must be 16-byte aligned");
  if (out == residual) return fail("vf_linear_residual: out must not alias residual");
  if (workspace_bytes < 0 || (workspace_bytes > 0 && !workspace)) return fail("vf_linear_residual: bad workspace");

  // alignment class: the largest power of two <= 256 dividing every operand address and every row/batch stride in bytes
  const size_t esz = dtype == VF_BF16 ? 2 : 4;
  uintptr_t bits = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(residual) |
                   reinterpret_cast<uintptr_t>(out) | (uintptr_t)(ld_x * esz) | (uintptr_t)(ld_res * esz) |
                   (uintptr_t)(ld_out * esz) | (uintptr_t)((size_t)k * esz) | 256u;
  if (bias) bits |= reinterpret_cast<uintptr_t>(bias);
  if (bias && batch > 1) bits |= (uintptr_t)((size_t)n * esz);
  const uint32_t align = (uint32_t)(bits & (~bits + 1));
  LtPlan plan;
  {
    std::lock_guard<std::mutex> lock(g_lt_mutex);
    if (int rc = load_lt_api()) return rc;
    if (!g_lt) VF_LT_TRY(g_api.Create(&g_lt));
    const auto key = std::make_tuple(rows, k, n, ld_x, ld_res, ld_out, dtype * 2 + (bias ? 1 : 0), (int)(workspace_bytes >> 20), batch + ((int)align << 16));
    auto it = g_plans.find(key);
    if (it == g_plans.end()) {
      LtPlan fresh;
      if (int rc = make_plan(fresh, rows, k, n, ld_x, ld_res, ld_out, dtype, bias != nullptr, (size_t)workspace_bytes, batch, align)) return rc;
      it = g_plans.emplace(key, fresh).first;
    }
    plan = it->second;
    // the bias pointer is per call (the descriptor is shared by all layers of one shape): set it under the lock and
    // enqueue under the lock as well -- cublasLtMatmul reads the descriptor at enqueue time
    if (bias) VF_LT_TRY(g_api.MatmulDescSetAttribute(plan.op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bias, sizeof(bias)));
    const float alpha = 1.0f, beta = 1.0f;
    cublasLtMatrixLayout_t d = plan.c;
    cublasLtMatrixLayout_t d_own = nullptr;
    if (ld_out != ld_res) {
      const cudaDataType_t dt = dtype == VF_BF16 ? CUDA_R_16BF : CUDA_R_32F;
      VF_LT_TRY(g_api.MatrixLayoutCreate(&d_own, dt, (uint64_t)n, (uint64_t)rows, (int64_t)ld_out));
      if (batch > 1) if (int rc = set_batch(d_own, batch, rows * ld_out)) return rc;
      d = d_own;
    }
    cublasStatus_t s = g_api.Matmul(g_lt, plan.op, &alpha, w, plan.a, x, plan.b, &beta, residual, plan.c, out, d, &plan.algo,
                                      workspace, (size_t)workspace_bytes, static_cast<cudaStream_t>(stream));
    if (d_own) g_api.MatrixLayoutDestroy(d_own);
    if (s != CUBLAS_STATUS_SUCCESS) return lt_fail(s, "cublasLtMatmul");
  }
  return 0;
}

extern "C" int vf_linear_residual(const void* x, const void* w, const void* bias, const void* residual, void* out,
                                  long long rows, int k, int n, long long ld_x, long long ld_res, long long ld_out,
                                  void* workspace, long long workspace_bytes, int dtype, void* stream) {
  return linear_residual_impl(x, w, bias, residual, out, 1, rows, k, n, ld_x, ld_res, ld_out, workspace, workspace_bytes, dtype, stream);
}

// Per-sample bias: x, residual, out are (batch, rows, .) contiguous per sample, bias is (batch, n).
extern "C" int vf_linear_residual_batched(const void* x, const void* w, const void* bias, const void* residual, void* out,
                                          int batch, long long rows, int k, int n, long long ld_x, long long ld_res,
                                          long long ld_out, void* workspace, long long workspace_bytes, int dtype, void* stream) {
  return linear_residual_impl(x, w, bias, residual, out, batch, rows, k, n, ld_x, ld_res, ld_out, workspace, workspace_bytes, dtype, stream);
}
