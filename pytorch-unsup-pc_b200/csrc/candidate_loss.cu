// Candidate-selection projection loss, fused.
//
// Reference: models/model_pc_to.py:339-385 add_proj_loss (AvgPool2d of the
// ground-truth masks from G x G to the prediction's V x V, :349-356) and
// :410-440 proj_loss_pose_candidates (per-candidate sum of squared differences,
// argmin over the candidates, one-hot-masked loss, optional per-sample
// weights, / BV).  The reference runs this as ~10 ATen launches over P x V x V
// temporaries (tf_repeat_0 of gt, sq_diff, one_hot, loss_tensor ...).
//
// Forward: one CTA per (sample, view) bv.  The pooled ground truth of the view
// is built once in registers (the G x G mask is read once), each of the C
// candidate masks is read once, the C sums are block-reduced in fixed order
// (fp64 across threads), thread 0 takes the first minimum (torch.argmin) and
// writes the view's weighted loss.  Nothing of size P x V x V is written.
// Backward: one CTA per (bv, candidate): zeros for the losing candidates,
// -2 k w^2 (gt - pred) for the winner, with the pooled gt rebuilt on the fly.
#include <cooperative_groups.h>

#include "common.cuh"

namespace dpc {

constexpr int kLossThreads = 256;
constexpr int kLossSplit = 8;          // CTAs (one cluster) per view of the forward
constexpr int kMaxCandidates = 16;

// One thread-block CLUSTER of kLossSplit CTAs per view: every CTA takes a slice of the V x V
// pixels (a single CTA per view is a ~40 us latency chain at 64 views or fewer -- 16 CTAs on 148
// SMs), reduces its C partial sums in fixed order (fp32 per thread, fp64 across threads), and
// rank 0 adds the slices' sums in rank order through distributed shared memory.
__global__ void __launch_bounds__(kLossThreads)
candidate_loss_fwd_kernel(const float *__restrict__ gt, const float *__restrict__ pred,
                          const float *__restrict__ weights, int C, int V, int G,
                          float *__restrict__ all_loss, long long *__restrict__ min_idx,
                          float *__restrict__ view_loss, int *__restrict__ winners,
                          float *__restrict__ kcoef, float coeff) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank(), split = cluster.num_blocks();
  const int bv = blockIdx.y, tid = threadIdx.x, n = G / V, VV = V * V;
  const float *g = gt + (size_t)bv * G * G;
  float acc[kMaxCandidates];
#pragma unroll
  for (int c = 0; c < kMaxCandidates; ++c) acc[c] = 0.f;
  const int per = (VV + split - 1) / split, i_hi = min(VV, (int)(rank + 1) * per);
  for (int i = rank * per + tid; i < i_hi; i += kLossThreads) {
    const int y = i / V, x = i - y * V;
    const float gp = pooled_gt(g, G, n, y, x);
#pragma unroll
    for (int c = 0; c < kMaxCandidates; ++c)
      if (c < C) {
        const float d = gp - __ldg(pred + ((size_t)bv * C + c) * VV + i);
        acc[c] = fmaf(d, d, acc[c]);
      }
  }
  // fixed-order reduction: warp shuffles (fp64), then warps in index order, then CTAs in rank order
  __shared__ double red[kLossThreads / 32][kMaxCandidates];
  __shared__ double slice[kMaxCandidates];
#pragma unroll
  for (int c = 0; c < kMaxCandidates; ++c)
    if (c < C) {
      double v = (double)acc[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if ((tid & 31) == 0) red[tid >> 5][c] = v;
    }
  __syncthreads();
  if (tid < C) {
    double v = 0;
    for (int w = 0; w < kLossThreads / 32; ++w) v += red[w][tid];
    slice[tid] = v;
  }
  cluster.sync();
  if (rank == 0 && tid == 0) {
    double best = 0;
    int arg = 0;
    for (int c = 0; c < C; ++c) {
      double v = 0;
      for (unsigned r = 0; r < split; ++r) v += *cluster.map_shared_rank(&slice[c], r);
      all_loss[(size_t)bv * C + c] = (float)v;
      if (c == 0 || v < best) { best = v; arg = c; }   // first minimum, like torch.argmin
    }
    const double w = weights ? (double)weights[bv] : 1.0;
    min_idx[bv] = arg;
    view_loss[bv] = (float)(w * w * best);
    // render_loss: the winning projection of the view and the factor of its mask gradient,
    // rounded exactly as candidate_loss_bwd_kernel rounds it
    if (winners) winners[bv] = bv * C + arg;
    if (kcoef) {
      const float wf = weights ? weights[bv] : 1.f;
      kcoef[bv] = -2.f * wf * wf * coeff;
    }
  }
  cluster.sync();               // no CTA may exit while its slice can still be read
}

__global__ void __launch_bounds__(kLossThreads)
candidate_loss_bwd_kernel(const float *__restrict__ gt, const float *__restrict__ pred,
                          const float *__restrict__ weights, const long long *__restrict__ min_idx,
                          const float *__restrict__ upstream, float coeff, int C, int V, int G,
                          float *__restrict__ g_pred) {
  const int img = blockIdx.x, bv = img / C, c = img - bv * C, n = G / V, VV = V * V;
  float *out = g_pred + (size_t)img * VV;
  if ((long long)c != min_idx[bv]) {
    for (int i = threadIdx.x; i < VV; i += kLossThreads) out[i] = 0.f;
    return;
  }
  const float w = weights ? weights[bv] : 1.f;
  // d/dpred sum(((gt - pred) w)^2) * coeff * upstream = -2 w^2 coeff upstream (gt - pred)
  const float k = -2.f * w * w * coeff * (upstream ? __ldg(upstream) : 1.f);
  const float *g = gt + (size_t)bv * G * G;
  const float *p = pred + (size_t)img * VV;
  for (int i = threadIdx.x; i < VV; i += kLossThreads) {
    const int y = i / V, x = i - y * V;
    out[i] = k * (pooled_gt(g, G, n, y, x) - __ldg(p + i));
  }
}

int launch_candidate_loss_fwd(const float *gt, const float *pred, const float *weights, int BV,
                              int C, int V, int G, float *all_loss, long long *min_idx,
                              float *view_loss, cudaStream_t s, int *winners, float *kcoef,
                              float coeff) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kLossSplit, BV);
  cfg.blockDim = dim3(kLossThreads);
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kLossSplit;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, candidate_loss_fwd_kernel, gt, pred, weights, C, V, G, all_loss, min_idx,
                     view_loss, winners, kcoef, coeff);
  return check_launch("candidate_loss_fwd");
}

// loss = coeff * sum_bv view_loss[bv]: one warp, lanes stride over the views, fixed shuffle tree
__global__ void loss_total_kernel(const float *__restrict__ view_loss, int BV, float coeff,
                                  float *__restrict__ loss) {
  pdl_wait();             // view_loss comes from candidate_loss_fwd
  double v = 0;
  for (int i = threadIdx.x; i < BV; i += 32) v += (double)ld_dep(view_loss + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (threadIdx.x == 0) loss[0] = (float)(v * (double)coeff);
}

int launch_loss_total(const float *view_loss, int BV, float coeff, float *loss, cudaStream_t s) {
  launch_dep(loss_total_kernel, dim3(1), dim3(32), 0, s, view_loss, BV, coeff, loss);
  return check_launch("loss_total");
}

int launch_candidate_loss_bwd(const float *gt, const float *pred, const float *weights,
                              const long long *min_idx, const float *upstream, float coeff, int BV,
                              int C, int V, int G, float *g_pred, cudaStream_t s) {
  candidate_loss_bwd_kernel<<<BV * C, kLossThreads, 0, s>>>(gt, pred, weights, min_idx, upstream,
                                                            coeff, C, V, G, g_pred);
  return check_launch("candidate_loss_bwd");
}

int candidate_loss_max_candidates() { return kMaxCandidates; }

}  // namespace dpc
