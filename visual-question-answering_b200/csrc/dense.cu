#include "dense.cuh"
#include "gemm_ffma.cuh"
#include "gemm_tc.cuh"
#include "util_kernels.cuh"

namespace hca {

namespace {
int pick_splitk(int M, int N, int K) {
  const int tiles = ((M + 127) / 128) * ((N + 127) / 128);
  int sk = (148 * 2 + tiles - 1) / tiles;
  const int maxk = (K + 255) / 256;
  if (sk > maxk) sk = maxk;
  if (sk > 64) sk = 64;
  return sk < 1 ? 1 : sk;
}

void fill_epi(GemmParams& g, const DenseEpi& e) {
  g.bias = e.bias;
  g.act_tanh = e.act_tanh;
  g.accumulate = e.accumulate;
  if (e.mulx) g.mulx = {e.mulx, 0, e.mulx_ld, 1, 0};
}
}  // namespace

static inline int64_t round8(int64_t x) { return (x + 7) / 8 * 8; }

// bf16 planes of both operands (up to 3 planes each); valid for all three layouts
size_t dense_scratch_bytes(int M, int N, int K) {
  const size_t a = (size_t)(M + 8) * (K + 8), b = (size_t)(N + 8) * (K + 8);
  return 3 * 2 * (a + b) + 4096;
}

namespace {
constexpr int TC_PLANES = 2;     // bf16x2 split: 3 MMAs, ~2^-16 operand precision (1e-5 end to end)

struct PlaneBuf {
  TcOperand op;
  bool ok = false;
};
// split an fp32 matrix [rows, cols] into bf16 planes carved from the workspace
PlaneBuf make_planes(const float* src, int64_t ld, int rows, int cols, bool mn_major, int P, Workspace& ws, cudaStream_t s, int* rc) {
  PlaneBuf b;
  const int64_t ldp = round8(cols);
  const int64_t stride = (int64_t)rows * ldp;
  __nv_bfloat16* buf = ws.take<__nv_bfloat16>((size_t)P * stride);
  if (!buf) {
    *rc = set_err(HCA_ERR_WORKSPACE, "dense: workspace too small for the bf16 operand planes (%d x %d)", rows, cols);
    return b;
  }
  *rc = launch_split_planes(src, ld, rows, cols, buf, ldp, stride, P, s);
  b.op.planes = buf; b.op.ld = ldp; b.op.plane_stride = stride; b.op.rows = rows; b.op.cols = cols; b.op.mn_major = mn_major;
  b.ok = (*rc == 0);
  return b;
}

int tc_splitk(int M, int N, int K) {
  const int tiles = ((M + 127) / 128) * ((N + 127) / 128);
  if (tiles >= 96) return 1;
  int sk = (148 + tiles - 1) / tiles;
  const int maxk = (K + 255) / 256;       // keep at least 4 k-blocks per split
  if (sk > maxk) sk = maxk;
  return sk < 1 ? 1 : sk;
}

bool want_tc(const DenseEpi* e) { return use_tc() && tc_available() && !(e && e->exact); }

// A: [a_rows, a_cols] fp32; B likewise; layouts given by the *_mn flags
int tc_gemm(const float* A, int64_t lda, int a_rows, int a_cols, bool a_mn, const float* B, int64_t ldb, int b_rows, int b_cols,
            bool b_mn, float* D, int64_t ldd, int M, int N, int K, const DenseEpi* e, int splitk, Workspace& ws, cudaStream_t s,
            int P = TC_PLANES) {
  const size_t mark = ws.off;
  int rc = 0;
  PlaneBuf pa = make_planes(A, lda, a_rows, a_cols, a_mn, P, ws, s, &rc);
  if (!pa.ok) { ws.off = mark; return rc; }
  PlaneBuf pb = make_planes(B, ldb, b_rows, b_cols, b_mn, P, ws, s, &rc);
  if (!pb.ok) { ws.off = mark; return rc; }
  TcEpilogue te;
  te.D = D; te.ldd = ldd;
  if (e) { te.bias = e->bias; te.act_tanh = e->act_tanh; te.mulx = e->mulx; te.mulx_ld = e->mulx_ld; te.accumulate = e->accumulate; }
  rc = launch_gemm_tc(pa.op, pb.op, P, M, N, K, te, splitk, s);
  ws.off = mark;                          // stream-ordered reuse: later kernels on the same stream run after this GEMM
  return rc;
}
}  // namespace

int dense_nt(const float* A, int64_t lda, const float* B, int64_t ldb, float* D, int64_t ldd, int M, int N, int K,
             const DenseEpi& e, Workspace& ws, cudaStream_t s) {
  if (want_tc(&e)) return tc_gemm(A, lda, M, K, false, B, ldb, N, K, false, D, ldd, M, N, K, &e, 1, ws, s);
  GemmParams g;
  g.A = {A, 0, lda, 1, 0};
  g.B = {B, 0, ldb, 1, 0};
  g.M = M; g.N = N; g.K = K;
  g.D = D; g.d_sm = ldd; g.d_sn = 1;
  fill_epi(g, e);
  return launch_gemm_ffma(g, M >= 128 && N >= 128, s);
}

int dense_nn(const float* A, int64_t lda, const float* B, int64_t ldb, float* D, int64_t ldd, int M, int N, int K,
             const DenseEpi& e, Workspace& ws, cudaStream_t s) {
  if (want_tc(&e)) return tc_gemm(A, lda, M, K, false, B, ldb, K, N, true, D, ldd, M, N, K, &e, 1, ws, s);
  GemmParams g;
  g.A = {A, 0, lda, 1, 0};
  g.B = {B, 0, 1, ldb, 0};      // B[n][k] = Bmat[k][n]
  g.M = M; g.N = N; g.K = K;
  g.D = D; g.d_sm = ldd; g.d_sn = 1;
  fill_epi(g, e);
  return launch_gemm_ffma(g, M >= 128 && N >= 128, s);
}

int dense_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* D, int64_t ldd, int M, int N, int K,
             bool zero_first, Workspace& ws, cudaStream_t s) {
  const bool tc = want_tc(nullptr);
  const int sk_tc = tc ? tc_splitk(M, N, K) : 1;
  if (zero_first && !(tc && sk_tc == 1)) {
    if (ldd == N) HCA_TRY(zero_async(D, (size_t)M * N * 4, s));
    else HCA_CUDA(cudaMemset2DAsync(D, ldd * 4, 0, (size_t)N * 4, M, s));
  }
  if (tc) {
    DenseEpi e;
    e.accumulate = zero_first ? 0 : 1;
    return tc_gemm(A, lda, K, M, true, B, ldb, K, N, true, D, ldd, M, N, K, &e, sk_tc, ws, s);
  }
  GemmParams g;
  g.A = {A, 0, 1, lda, 0};      // A[m][k] = Amat[k][m]
  g.B = {B, 0, 1, ldb, 0};
  g.M = M; g.N = N; g.K = K;
  g.D = D; g.d_sm = ldd; g.d_sn = 1;
  g.splitk = pick_splitk(M, N, K);
  if (g.splitk == 1) g.accumulate = 1;
  return launch_gemm_ffma(g, M >= 128 && N >= 128, s);
}

}  // namespace hca

// ---- test / profiling entry point of the C ABI -------------------------------------------------------
extern "C" size_t hca_gemm_workspace(int M, int N, int K) { return hca::dense_scratch_bytes(M, N, K) + 1024; }

extern "C" int hca_gemm(const float* A, const float* B, const float* bias, float* D, int M, int N, int K, int layout, int path,
                        void* ws, size_t ws_bytes, void* stream) {
  using namespace hca;
  HCA_CHECK_ARG(A && B && D, "gemm: null pointer");
  HCA_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: bad sizes");
  HCA_CHECK_ARG(layout >= 0 && layout <= 2 && path >= 0 && path <= 2, "gemm: layout in {0 nt,1 nn,2 tn}, path in {0,1,2}");
  HCA_CHECK_ARG(!(layout == 2 && bias), "gemm: the tn (weight-gradient) layout has no bias epilogue");
  Workspace w(ws, ws_bytes);
  cudaStream_t s = (cudaStream_t)stream;
  if (path == 0) {
    DenseEpi e;
    e.bias = bias;
    e.exact = 1;
    // the exact path is selected by e.exact inside dense_nt / dense_nn; dense_tn takes the option switch
    if (layout == 0) return dense_nt(A, K, B, K, D, N, M, N, K, e, w, s);
    if (layout == 1) return dense_nn(A, K, B, N, D, N, M, N, K, e, w, s);
    HCA_TRY(zero_async(D, (size_t)M * N * 4, s));
    GemmParams g;
    g.A = {A, 0, 1, M, 0};
    g.B = {B, 0, 1, N, 0};
    g.M = M; g.N = N; g.K = K;
    g.D = D; g.d_sm = N; g.d_sn = 1;
    g.splitk = pick_splitk(M, N, K);
    if (g.splitk == 1) g.accumulate = 1;
    return launch_gemm_ffma(g, M >= 128 && N >= 128, s);
  }
  HCA_CHECK_ARG(tc_available(), "gemm: the tensor-core path needs cuTensorMapEncodeTiled from the driver");
  const int P = path == 1 ? 2 : 3;
  DenseEpi e;
  e.bias = bias;
  if (layout == 0) return tc_gemm(A, K, M, K, false, B, K, N, K, false, D, N, M, N, K, &e, 1, w, s, P);
  if (layout == 1) return tc_gemm(A, K, M, K, false, B, N, K, N, true, D, N, M, N, K, &e, 1, w, s, P);
  const int sk = tc_splitk(M, N, K);
  if (sk > 1) HCA_TRY(zero_async(D, (size_t)M * N * 4, s));
  return tc_gemm(A, M, K, M, true, B, N, K, N, true, D, N, M, N, K, &e, sk, w, s, P);
}

// ---- the projection kernel on its own (bench.py's roofline leg): operands already split into bf16 hi/lo planes ------
extern "C" int hca_split_planes(const float* src, int64_t rows, int cols, void* planes, void* stream) {
  using namespace hca;
  HCA_CHECK_ARG(src && planes && rows > 0 && cols > 0 && cols % 8 == 0, "split_planes: bad arguments (cols %% 8 == 0 required)");
  return launch_split_planes(src, cols, rows, cols, (__nv_bfloat16*)planes, cols, rows * cols, 2, (cudaStream_t)stream);
}
extern "C" int hca_proj_planes(const void* a_planes, int64_t M, int K, const void* w_planes, int N, const float* bias, void* out_planes,
                               void* stream) {
  using namespace hca;
  HCA_CHECK_ARG(a_planes && w_planes && out_planes && M > 0 && M < (1LL << 31) && K > 0 && N > 0 && K % 8 == 0 && N % 8 == 0,
                "proj_planes: bad arguments (K, N %% 8 == 0 required)");
  HCA_CHECK_ARG(tc_available(), "proj_planes: cuTensorMapEncodeTiled is not available from the driver");
  TcOperand A, B;
  A.planes = (const __nv_bfloat16*)a_planes; A.ld = K; A.plane_stride = M * K; A.rows = (int)M; A.cols = K;
  B.planes = (const __nv_bfloat16*)w_planes; B.ld = K; B.plane_stride = (int64_t)N * K; B.rows = N; B.cols = K;
  TcEpilogue e;
  e.bias = bias;
  e.P.p = (__nv_bfloat16*)out_planes; e.P.ld = N; e.P.plane_stride = M * N; e.P.batch_stride = 0; e.P.nbatch = 1;
  return launch_gemm_tc(A, B, 2, (int)M, N, K, e, 1, (cudaStream_t)stream);
}

extern "C" int hca_debug_gemm_timeline(void* buf, int nctas) {
  hca::tc_set_timeline((long long*)buf, buf ? nctas : 0);
  return 0;
}
extern "C" int hca_debug_gemm_timeline_select(void* buf, int nctas, int launch_index) {
  hca::tc_set_timeline((long long*)buf, buf ? nctas : 0, launch_index);
  return 0;
}
