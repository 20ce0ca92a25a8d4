// Test / profiling entry points of the C ABI around the tensor-core GEMM (gemm_tc.cu): one dense (un-batched) contraction from
// fp32 operands in the three layouts the path needs, and the projection kernel on operands that are already bf16 planes.
//
//   nt : D[M,N] = A[M,K] . B[N,K]^T      forward of nn.Linear (x . W^T), B is a weight [out,in]
//   nn : D[M,N] = A[M,K] . B[K,N]        data gradient (dY . W)
//   tn : D[M,N] = A[K,M]^T . B[K,N]      weight gradient (dY^T . X), long K -> split-K
#include <algorithm>
#include "gemm_tc.cuh"
#include "util_kernels.cuh"

namespace hca {
namespace {
inline int64_t round8(int64_t x) { return (x + 7) / 8 * 8; }

// bf16 planes of both operands (up to 3 planes each); valid for all three layouts
size_t gemm_scratch_bytes(int M, int N, int K) {
  const size_t a = (size_t)(M + 8) * (K + 8), b = (size_t)(N + 8) * (K + 8);
  return 3 * 2 * (a + b) + 4096;
}

// split an fp32 matrix [rows, cols] into bf16 planes carved from the workspace
int make_planes(TcOperand& op, const float* src, int64_t ld, int rows, int cols, bool mn_major, int P, Workspace& ws, cudaStream_t s) {
  const int64_t ldp = round8(cols);
  const int64_t stride = (int64_t)rows * ldp;
  __nv_bfloat16* buf = ws.take<__nv_bfloat16>((size_t)P * stride);
  if (!buf) return set_err(HCA_ERR_WORKSPACE, "gemm: workspace too small for the bf16 operand planes (%d x %d)", rows, cols);
  HCA_TRY(launch_split_planes(src, ld, rows, cols, buf, ldp, stride, P, s));
  op.planes = buf; op.ld = ldp; op.plane_stride = stride; op.rows = rows; op.cols = cols; op.mn_major = mn_major;
  return 0;
}

}  // namespace
}  // namespace hca

extern "C" size_t hca_gemm_workspace(int M, int N, int K) { return hca::gemm_scratch_bytes(M, N, K) + 1024; }

extern "C" int hca_gemm(const float* A, const float* B, const float* bias, float* D, int M, int N, int K, int layout, int path,
                        void* ws, size_t ws_bytes, void* stream) {
  using namespace hca;
  HCA_CHECK_ARG(A && B && D, "gemm: null pointer");
  HCA_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: bad sizes");
  HCA_CHECK_ARG(layout >= 0 && layout <= 2 && (path == 1 || path == 2), "gemm: layout in {0 nt,1 nn,2 tn}, path in {1 bf16x2, 2 bf16x3}");
  HCA_CHECK_ARG(!(layout == 2 && bias), "gemm: the tn (weight-gradient) layout has no bias epilogue");
  HCA_CHECK_ARG(tc_available(), "gemm: cuTensorMapEncodeTiled is not available from the driver");
  Workspace w(ws, ws_bytes);
  cudaStream_t s = (cudaStream_t)stream;
  const int P = path == 1 ? 2 : 3;
  TcOperand a, b;
  int splitk = 1;
  if (layout == 0) {
    HCA_TRY(make_planes(a, A, K, M, K, false, P, w, s));
    HCA_TRY(make_planes(b, B, K, N, K, false, P, w, s));
  } else if (layout == 1) {
    HCA_TRY(make_planes(a, A, K, M, K, false, P, w, s));
    HCA_TRY(make_planes(b, B, N, K, N, true, P, w, s));
  } else {
    HCA_TRY(make_planes(a, A, M, K, M, true, P, w, s));
    HCA_TRY(make_planes(b, B, N, K, N, true, P, w, s));
    splitk = tc_splitk(M, N, K);
    if (splitk > 1) HCA_TRY(zero_async(D, (size_t)M * N * 4, s));
  }
  TcEpilogue e;
  e.D = D; e.ldd = N; e.bias = bias;
  return launch_gemm_tc(a, b, P, M, N, K, e, splitk, s);
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

// ---- a weight-gradient product on its own (bench.py's roofline legs): D[M,N] += A[K,M]^T . B[K,N], both operands MN-major planes --
extern "C" int hca_wgrad_planes(const void* a_planes, const void* b_planes, int M, int N, int64_t K, float* D, void* stream) {
  using namespace hca;
  HCA_CHECK_ARG(a_planes && b_planes && D && M > 0 && N > 0 && K > 0 && K < (1LL << 31) && M % 8 == 0 && N % 8 == 0,
                "wgrad_planes: bad arguments (M, N %% 8 == 0 required)");
  HCA_CHECK_ARG(tc_available(), "wgrad_planes: cuTensorMapEncodeTiled is not available from the driver");
  TcOperand A, B;
  A.planes = (const __nv_bfloat16*)a_planes; A.ld = M; A.plane_stride = K * M; A.rows = (int)K; A.cols = M; A.mn_major = true;
  B.planes = (const __nv_bfloat16*)b_planes; B.ld = N; B.plane_stride = K * N; B.rows = (int)K; B.cols = N; B.mn_major = true;
  TcEpilogue e;
  e.D = D; e.ldd = N; e.accumulate = 1;
  return launch_gemm_tc(A, B, 2, M, N, (int)K, e, std::max(2, tc_splitk(M, N, (int)K)), (cudaStream_t)stream);
}

extern "C" int hca_debug_gemm_timeline(void* buf, int nctas) {
  hca::tc_set_timeline((long long*)buf, buf ? nctas : 0);
  return 0;
}
extern "C" int hca_debug_gemm_timeline_select(void* buf, int nctas, int launch_index) {
  hca::tc_set_timeline((long long*)buf, buf ? nctas : 0, launch_index);
  return 0;
}
