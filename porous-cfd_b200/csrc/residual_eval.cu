// The loss modules of the reference on EXPLICIT tensors (models/losses.py:154-156, 177-182, 209-217, 256-266, 301-311):
// `ContinuityLoss*.func(jacobian)` and `MomentumLoss*.func(internal_input, u, u_jac, u_laplace, p_grad)` as one
// elementwise kernel over the points, and `forward` = the mean of the squares per component (vector_loss :10-20 /
// mse_loss), reduced in a fixed order.  The training step does not come through here (its residual, losses and
// gradient are the fused pcfd_residual_loss); these entry points serve callers that hold the derivative tensors
// themselves, e.g. the reference's predict_step (models/model_base.py:241-246) and evaluation scripts.
#include "common.cuh"

namespace pcfd {

struct EvalArgs {
  const float *u, *jac, *lap, *pgrad, *zone, *dcoef, *fcoef;
  int64_t rows;
  pcfd_residual_params_t p;
  float *momentum, *div;
};

template <int D>
__global__ void __launch_bounds__(256) residual_eval_kernel(EvalArgs a) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.rows) return;
  const pcfd_residual_params_t& P = a.p;
  const bool manu = P.loss_kind == PCFD_LOSS_MANUFACTURED;
  float su[D], mu[D], sx[D];
  const float sp = manu ? 1.0f : P.p_std;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    su[d] = manu ? 1.0f : P.u_std[d];
    mu[d] = manu ? 0.0f : P.u_mean[d];
    sx[d] = manu ? 1.0f : P.c_std[d];
  }
  if (a.div != nullptr) {
    float div = 0.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) div += __ldg(a.jac + (t * D + d) * D + d) * su[d] / sx[d];
    a.div[t] = div;
  }
  if (a.momentum == nullptr) return;
  float ur[D], nrm2 = 0.0f;
#pragma unroll
  for (int d = 0; d < D; ++d) { ur[d] = su[d] * __ldg(a.u + t * D + d) + mu[d]; nrm2 += ur[d] * ur[d]; }
  const float nrm = sqrtf(nrm2);
  const float zone = a.zone != nullptr ? __ldg(a.zone + t) : 0.0f;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    float conv = 0.0f, visc = 0.0f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      conv += __ldg(a.jac + (t * D + c) * D + j) * (ur[j] / sx[j]);
      visc += __ldg(a.lap + (t * D + c) * D + j) * (1.0f / (sx[j] * sx[j]));
    }
    conv *= su[c];
    visc *= P.nu * su[c];
    const float pres = (sp / sx[c]) * __ldg(a.pgrad + t * D + c);
    float dc = P.d, fc = P.f;
    if (P.loss_kind == PCFD_LOSS_VARIABLE) {
      dc = P.d_min[c] + P.d_range[c] * __ldg(a.dcoef + t * D + c);
      fc = P.f_min[c] + P.f_range[c] * __ldg(a.fcoef + t * D + c);
    }
    const float source = ur[c] * (dc * P.nu + 0.5f * nrm * fc);
    float r = conv - visc + pres + source * zone;
    if (manu && a.fcoef != nullptr) r -= __ldg(a.fcoef + t * D + c);
    a.momentum[t * D + c] = r;
  }
}

// out[c] = mean over rows of x[row][c]^2: one CTA per column, fixed-order tree (deterministic)
__global__ void __launch_bounds__(1024) mean_squares_kernel(const float* __restrict__ x, int64_t rows, int cols,
                                                            float* __restrict__ out) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  float s = 0.0f;
  for (int64_t r = threadIdx.x; r < rows; r += 1024) { const float v = __ldg(x + r * cols + c); s += v * v; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[c] = s / (float)rows;
  }
}

}  // namespace pcfd

using namespace pcfd;

extern "C" int pcfd_residual_eval(const float* u, const float* jac, const float* lap, const float* p_grad, const float* zone,
                                  const float* dcoef, const float* fcoef, int64_t rows, const pcfd_residual_params_t* prm_host,
                                  float* momentum, float* div, void* stream) {
  if (!prm_host || rows <= 0 || (!momentum && !div) || !jac) return PCFD_ERR_ARG;
  if (momentum && (!u || !lap || !p_grad)) return PCFD_ERR_ARG;
  if (momentum && prm_host->loss_kind == PCFD_LOSS_VARIABLE && (!dcoef || !fcoef)) return PCFD_ERR_ARG;
  if (prm_host->dims != 2 && prm_host->dims != 3) return PCFD_ERR_ARG;
  const int rc = check_sm100();
  if (rc) return rc;
  EvalArgs a{u, jac, lap, p_grad, zone, dcoef, fcoef, rows, *prm_host, momentum, div};
  const unsigned grid = (unsigned)((rows + 255) / 256);
  if (prm_host->dims == 2) residual_eval_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  else residual_eval_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

extern "C" int pcfd_mean_squares(const float* x, int64_t rows, int32_t cols, float* out, void* stream) {
  if (!x || !out || rows <= 0 || cols <= 0) return PCFD_ERR_ARG;
  mean_squares_kernel<<<(unsigned)cols, 1024, 0, (cudaStream_t)stream>>>(x, rows, cols, out);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}
