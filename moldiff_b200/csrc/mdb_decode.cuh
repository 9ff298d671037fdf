// On-device decode of the sampler's final predictions: per node / half-edge row
// softmax -> (argmax class, its probability), i.e. the arithmetic of FeaturizeMol.decode_output
// (reference utils/transforms.py:76-96: scipy softmax, np.argmax, np.max) for the whole batch in one launch, so that only
// 5 B/row instead of the logits (and none of the 1.9 GB trajectory) has to cross PCIe before RDKit reconstruction.
// Masking / bond filtering / per-molecule splitting are index bookkeeping and stay on the host (moldiff_b200/decode.py).
// Included by mdb_forward.cu inside its anonymous namespace.
#pragma once

__global__ void __launch_bounds__(256) decode_rows_kernel(int n_nodes, int kn, const float* __restrict__ pred_node, int n_half,
                                                          int ke, const float* __restrict__ pred_half,
                                                          uint8_t* __restrict__ node_type, float* __restrict__ node_prob,
                                                          uint8_t* __restrict__ half_type, float* __restrict__ half_prob) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool is_node = i < n_nodes;
  if (!is_node && i >= n_nodes + n_half) return;
  const int r = is_node ? i : i - n_nodes;
  const int K = is_node ? kn : ke;
  const float* x = (is_node ? pred_node : pred_half) + (size_t)r * K;
  float mx = x[0];
  int best = 0;
  for (int k = 1; k < K; ++k) {
    if (x[k] > mx) { mx = x[k]; best = k; }              // first maximum, like np.argmax
  }
  float se = 0.f;
  for (int k = 0; k < K; ++k) se += expf(x[k] - mx);
  (is_node ? node_type : half_type)[r] = (uint8_t)best;
  (is_node ? node_prob : half_prob)[r] = 1.f / se;       // softmax at the argmax: exp(0) / sum
}
