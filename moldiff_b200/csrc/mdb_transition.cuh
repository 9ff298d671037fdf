// Fused reverse-transition step of the sampler: everything MolDiff.sample does between two
// denoiser evaluations (reference models/model.py:287-300,365-372) in ONE launch instead of ~100 elementwise ones:
//   positions:   x_{t-1} = coef_x0[t] x0_pred + coef_xt[t] x_t + std[t] z          (z ignored at t == 0)
//                                                  ContigousTransition.get_prev_from_recon, models/transition.py:44-63
//   categorical: log q(v_{t-1} | v_t, v0_pred) = log(exp(log_vt) Q_t^T) + log(softmax(pred) Qbar_{t-1}), normalised;
//                = log softmax(pred) at t == 0                GeneralCategoricalTransition.q_v_posterior, :285-315
//                v_{t-1} = argmax(log q + Gumbel(u))          log_sample_categorical, models/diffusion.py:79-85
//                one-hot encode; half-edge one-hots are written twice ([2 Eh, K]: the directed edge list of the next step)
// The uniform / normal variates come from the caller (torch's generator, same consumption order as the PyTorch path:
// z_pos, u_node, u_half), so on one device the fused step reproduces the unfused one.
// One thread per node / half-edge row; ~200 B of HBM traffic per row, no reuse: a pure streaming kernel.
//
// Included by mdb_forward.cu inside its anonymous namespace.
#pragma once

struct TransArgs {
  int n_nodes, n_half, kn, ke;
  const int64_t *batch_node, *batch_half, *t;        // t: [n_graphs]
  // positions
  const float *pos, *pred_pos, *z_pos;               // [N][3]
  const float *coef_x0, *coef_xt, *stdv;             // [T]
  float* pos_out;
  // node types
  const float *pred_node, *log_node, *u_node;        // [N][Kn]
  const float *qn_cum, *qn_stepT;                    // [T][Kn][Kn]: q_mats, transpopse_q_onestep_mats
  float *log_node_out, *h_node_out;
  // half-edge types
  const float *pred_half, *log_half, *u_half;        // [Eh][Ke]
  const float *qe_cum, *qe_stepT;
  float *log_half_out, *h_edge_out;                  // h_edge_out: [2 Eh][Ke]
  int64_t* half_type_out;                            // [Eh] sampled class ids (guidance objectives use them), may be null
};

constexpr float TR_TINY = 1e-30f, TR_LOG_FLOOR = -32.f;

// KFIX > 0: class count known at compile time (loops unroll, rows stay in registers); KFIX = 0: runtime K <= KMAX
template <int KMAX, int KFIX>
__device__ __forceinline__ int categorical_row(int Krt, int64_t t, const float* __restrict__ pred, const float* __restrict__ log_vt,
                                               const float* __restrict__ u, const float* __restrict__ q_cum,
                                               const float* __restrict__ q_stepT, float* __restrict__ log_out) {
  const int K = KFIX > 0 ? KFIX : Krt;
  float lv0[KMAX], lvt[KMAX], out[KMAX];
  float mx = -INFINITY;
  _Pragma("unroll") for (int k = 0; k < K; ++k) { lv0[k] = pred[k]; lvt[k] = log_vt[k]; mx = fmaxf(mx, lv0[k]); }
  float se = 0.f;
  _Pragma("unroll") for (int k = 0; k < K; ++k) se += expf(lv0[k] - mx);
  const float lse = mx + logf(se);
  _Pragma("unroll") for (int k = 0; k < K; ++k) lv0[k] -= lse;                                     // log_softmax(pred)
  if (t == 0) {
    _Pragma("unroll") for (int k = 0; k < K; ++k) out[k] = lv0[k];
  } else {
    const float* st = q_stepT + (size_t)t * K * K;
    const float* cu = q_cum + (size_t)(t - 1) * K * K;
    float f1[KMAX], f2[KMAX];
    _Pragma("unroll") for (int k = 0; k < K; ++k) { f1[k] = 0.f; f2[k] = 0.f; }
    _Pragma("unroll") for (int j = 0; j < K; ++j) {
      const float pt = expf(lvt[j]), p0 = expf(lv0[j]);
      _Pragma("unroll") for (int k = 0; k < K; ++k) { f1[k] = fmaf(pt, st[j * K + k], f1[k]); f2[k] = fmaf(p0, cu[j * K + k], f2[k]); }
    }
    float m2 = -INFINITY;
    _Pragma("unroll") for (int k = 0; k < K; ++k) {
      out[k] = fmaxf(logf(f1[k] + TR_TINY), TR_LOG_FLOOR) + fmaxf(logf(f2[k] + TR_TINY), TR_LOG_FLOOR);
      m2 = fmaxf(m2, out[k]);
    }
    float s2 = 0.f;
    _Pragma("unroll") for (int k = 0; k < K; ++k) s2 += expf(out[k] - m2);
    const float l2 = m2 + logf(s2);
    _Pragma("unroll") for (int k = 0; k < K; ++k) out[k] -= l2;
  }
  int best = 0;
  float bv = -INFINITY;
  _Pragma("unroll") for (int k = 0; k < K; ++k) {
    log_out[k] = out[k];
    const float g = -logf(-logf(u[k] + TR_TINY) + TR_TINY) + out[k];
    if (g > bv) { bv = g; best = k; }                                           // first maximum, like torch.argmax
  }
  return best;
}

template <int KN, int KE>
__global__ void __launch_bounds__(256) transition_step_kernel(const TransArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.n_nodes) {
    const int64_t t = a.t[a.batch_node[i]];
    const float c0 = a.coef_x0[t], ct = a.coef_xt[t], sd = (t == 0) ? 0.f : a.stdv[t];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float mu = c0 * a.pred_pos[i * 3 + d] + ct * a.pos[i * 3 + d];
      a.pos_out[i * 3 + d] = (t == 0) ? mu : mu + sd * a.z_pos[i * 3 + d];
    }
    const int K = KN > 0 ? KN : a.kn;
    const int cls = categorical_row<(KN > 0 ? KN : 16), KN>(K, t, a.pred_node + (size_t)i * K, a.log_node + (size_t)i * K, a.u_node + (size_t)i * K,
                                        a.qn_cum, a.qn_stepT, a.log_node_out + (size_t)i * K);
    _Pragma("unroll") for (int k = 0; k < K; ++k) a.h_node_out[(size_t)i * K + k] = (k == cls) ? 1.f : 0.f;
  } else if (i < a.n_nodes + a.n_half) {
    const int h = i - a.n_nodes;
    const int64_t t = a.t[a.batch_half[h]];
    const int K = KE > 0 ? KE : a.ke;
    const int cls = categorical_row<(KE > 0 ? KE : 16), KE>(K, t, a.pred_half + (size_t)h * K, a.log_half + (size_t)h * K, a.u_half + (size_t)h * K,
                                        a.qe_cum, a.qe_stepT, a.log_half_out + (size_t)h * K);
    _Pragma("unroll") for (int k = 0; k < K; ++k) {
      const float v = (k == cls) ? 1.f : 0.f;
      a.h_edge_out[(size_t)h * K + k] = v;
      a.h_edge_out[(size_t)(a.n_half + h) * K + k] = v;
    }
    if (a.half_type_out) a.half_type_out[h] = cls;
  }
}
