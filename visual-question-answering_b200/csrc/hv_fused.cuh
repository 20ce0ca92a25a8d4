// Region-side hidden state Hv_l = tanh(PV + C_l^T PQ_l) of the co-attention for the three question levels in one kernel: the
// attention scores (forward) and dZv / dPV / dwv / dbv (backward).  See hv_fused.cu.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace hca {

struct HvPlanes {              // bf16 hi/lo planes of a row-major matrix: element (pl, row, col) at p[pl * ps + row * ld + col]
  __nv_bfloat16* p = nullptr;
  int64_t ld = 0, ps = 0;      // elements, multiples of 8
};

// C planes [B][3T][N] (C_all[b] = tanh(Q_all[b] V[b]^T)), PQ planes [B][3T][d], PV planes [B][N][d]; wv [d].
// sv [B][3][N] must be zeroed: sv[b][l][n] += tanh(PV[b] + C_l[b]^T PQ_l[b])[n,:] . wv
int launch_hv_scores(const HvPlanes& C, const HvPlanes& PQ, const HvPlanes& PV, const float* wv, float* sv, int B, int N, int T, int d,
                     cudaStream_t s);
// dsv [B][3][N]; dZq planes [B][3T][d].  Outputs: dZv planes [B][3][N][d], dPV planes [B][N][d]; dwv, dbv [d] must be zeroed.
int launch_hv_grads(const HvPlanes& C, const HvPlanes& PQ, const HvPlanes& PV, const HvPlanes& dZq, const float* wv, const float* dsv,
                    const HvPlanes& dZv, const HvPlanes& dPV, float* dwv, float* dbv, int B, int N, int T, int d, cudaStream_t s);

}  // namespace hca
