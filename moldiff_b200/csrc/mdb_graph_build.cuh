// Edge builders over 3D atom positions: brute-force radius_graph / knn_graph inside each molecule.
// The reference imports torch_geometric's radius_graph / knn_graph (models/graph.py:6) but only reaches them from dead
// code (graph.py:101 commented out, :466 in the unused NodeBondNet), and torch_cluster is not installable here, so their
// results are UNPINNED by the reference: semantics follow torch_cluster's documented behaviour (same-graph pairs only,
// `loop` controls self edges, radius: at most max_num_neighbors per centre in index order; knn: the k nearest by
// distance, ties by index) and are validated against a brute-force PyTorch restatement (tests/test_gpu_graph_build.py).
// Molecules have <= a few hundred atoms: one thread per centre node scans its own graph segment (positions are L1/L2
// resident); the output is a dense [n][max] neighbour table + counts, compacted to edge_index by the caller.
//
// Included by mdb_forward.cu inside its anonymous namespace.
#pragma once

// squared distance without FMA contraction, summed in the order a plain (dx*dx).sum(-1) uses: borderline pairs then fall on
// the same side of the cutoff as in the PyTorch restatement
__device__ __forceinline__ float dist2(float dx, float dy, float dz) {
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__global__ void __launch_bounds__(128) radius_graph_kernel(int n, const float* __restrict__ pos, const int* __restrict__ seg_lo,
                                                           const int* __restrict__ seg_hi, float r2, int loop, int max_nb,
                                                           int* __restrict__ counts, int* __restrict__ nbr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float xi = pos[i * 3], yi = pos[i * 3 + 1], zi = pos[i * 3 + 2];
  int c = 0;
  for (int j = seg_lo[i]; j < seg_hi[i] && c < max_nb; ++j) {
    if (j == i && !loop) continue;
    const float dx = pos[j * 3] - xi, dy = pos[j * 3 + 1] - yi, dz = pos[j * 3 + 2] - zi;
    if (dist2(dx, dy, dz) < r2) nbr[(size_t)i * max_nb + c++] = j;
  }
  counts[i] = c;
}

constexpr int KNN_MAX_K = 32;
__global__ void __launch_bounds__(128) knn_graph_kernel(int n, const float* __restrict__ pos, const int* __restrict__ seg_lo,
                                                        const int* __restrict__ seg_hi, int k, int loop,
                                                        int* __restrict__ counts, int* __restrict__ nbr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float xi = pos[i * 3], yi = pos[i * 3 + 1], zi = pos[i * 3 + 2];
  float bd[KNN_MAX_K];
  int bj[KNN_MAX_K];
  int c = 0;
  for (int j = seg_lo[i]; j < seg_hi[i]; ++j) {
    if (j == i && !loop) continue;
    const float dx = pos[j * 3] - xi, dy = pos[j * 3 + 1] - yi, dz = pos[j * 3 + 2] - zi;
    const float d = dist2(dx, dy, dz);
    if (c == k && !(d < bd[k - 1])) continue;          // not better than the current k-th (ties keep the lower index)
    int p = c < k ? c++ : k - 1;                        // insertion sort, ascending distance, stable in j
#pragma unroll 1
    for (; p > 0 && bd[p - 1] > d; --p) { bd[p] = bd[p - 1]; bj[p] = bj[p - 1]; }
    bd[p] = d; bj[p] = j;
  }
  for (int t = 0; t < c; ++t) nbr[(size_t)i * k + t] = bj[t];
  counts[i] = c;
}
