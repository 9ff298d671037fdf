/*
 * moldiff_b200 -- C-ABI of the B200-native MolDiff denoising hot path.
 *
 * The reference (pengxingang/MolDiff) is pure Python and has no FFI; its seam for this path is the
 * Python class surface.  This header is what a maintainer binds from Python
 * (ctypes / cffi) to replace the bodies of:
 *
 *   mdb_net_forward       <- NodeEdgeNet.forward                models/graph.py:348-367
 *   mdb_moldiff_forward   <- MolDiff.forward                    models/model.py:204-234
 *   mdb_bondpred_forward  <- BondPredictor.forward              models/bond_predictor.py:128-162
 *   mdb_bondpred_backward <- torch.autograd.grad(objective, pos_in) in the guidance block of
 *                            MolDiff.sample                      models/model.py:309-325
 *   mdb_plan              <- (new) one-off per-batch graph plan; the reference re-reads edge_index
 *                            in every scatter/gather (graph.py:37,275,385)
 *
 * Conventions: every pointer is a DEVICE pointer unless its name ends in _host; tensors are dense,
 * row-major, float32 / int32 / int64 as typed; `stream` is a cudaStream_t passed as void*.  All entry
 * points are asynchronous on `stream`, never synchronise the device, and return 0 on success or a
 * negative MDB_E* code (mdb_last_error() gives the message).  No torch types cross this boundary.
 */
#ifndef MOLDIFF_B200_H
#define MOLDIFF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDB_NODE_DIM 256 /* config.node_dim the kernels are specialised for (configs/train/*.yml) */
#define MDB_EDGE_DIM 64  /* config.edge_dim */
#define MDB_NUM_RBF 16   /* NodeEdgeNet num_gaussians default (graph.py:308-311) */
#define MDB_MAX_BLOCKS 16

#define MDB_OK 0
#define MDB_EINVAL (-1)  /* bad argument / unsupported shape */
#define MDB_ECUDA (-2)   /* CUDA runtime error */
#define MDB_EARCH (-3)   /* device is not sm_100 */

/*
 * Weight slots of ONE NodeEdgeNet block inside the packed fp32 blob (offsets in floats).  Every matrix
 * is stored [K][N] row-major, i.e. transposed from torch's Linear.weight [out=N][in=K], and split where
 * a first-layer Linear acts on a concatenation (so that the per-node part can be hoisted out of the
 * per-edge work -- SURVEY.md section 7.1).  X(name) order is the ABI: moldiff_b200/packing.py parses
 * this list.
 */
#define MDB_BLOCK_SLOTS(X)                                                                         \
  /* edge_embs.i : Linear(64+16 -> 64)                                        graph.py:338,357 */  \
  X(EE_W) X(EE_B)                                                                                  \
  /* node_blocks_with_edge.i (NodeBlock)                                      graph.py:12-55   */  \
  X(NB_NN1_W) X(NB_NN1_B) X(NB_NN1_G) X(NB_NN1_BE) X(NB_NN2_W) X(NB_NN2_B)                         \
  X(NB_EN1_W) X(NB_EN1_B) X(NB_EN1_G) X(NB_EN1_BE) X(NB_EN2_W) X(NB_EN2_B)                         \
  X(NB_MSG_W) X(NB_MSG_B)                                                                          \
  X(NB_GE_W) X(NB_GX_W) X(NB_GT_W) X(NB_G1_B) X(NB_G1_G) X(NB_G1_BE) X(NB_G2_W) X(NB_G2_B)         \
  X(NB_CEN_W) X(NB_CEN_B) X(NB_LN_G) X(NB_LN_BE) X(NB_OUT_W) X(NB_OUT_B)                           \
  /* edge_blocks.i.bond_ffn_left / right (BondFFN 64,256 -> 128 -> 64)        graph.py:122-141 */  \
  X(EL_BL_W) X(EL_NL_W) X(EL_I1_W) X(EL_I1_B) X(EL_I1_G) X(EL_I1_BE) X(EL_I2_W) X(EL_I2_B)         \
  X(EL_GB_W) X(EL_GN_W) X(EL_GT_W) X(EL_G1_B) X(EL_G1_G) X(EL_G1_BE) X(EL_G2_W) X(EL_G2_B)         \
  X(ER_BL_W) X(ER_NL_W) X(ER_I1_W) X(ER_I1_B) X(ER_I1_G) X(ER_I1_BE) X(ER_I2_W) X(ER_I2_B)         \
  X(ER_GB_W) X(ER_GN_W) X(ER_GT_W) X(ER_G1_B) X(ER_G1_G) X(ER_G1_BE) X(ER_G2_W) X(ER_G2_B)         \
  /* edge_blocks.i remaining (EdgeBlock)                                      graph.py:260-295 */  \
  X(EB_NFL_W) X(EB_NFL_B) X(EB_NFR_W) X(EB_NFR_B) X(EB_SELF_W) X(EB_SELF_B)                        \
  X(EB_LN_G) X(EB_LN_BE) X(EB_OUT_W) X(EB_OUT_B)                                                   \
  /* pos_blocks.i (PosUpdate; absent when update_pos = False)                 graph.py:377-396 */  \
  X(PU_LL1_W) X(PU_LL1_B) X(PU_LL1_G) X(PU_LL1_BE) X(PU_LL2_W) X(PU_LL2_B)                         \
  X(PU_RL1_W) X(PU_RL1_B) X(PU_RL1_G) X(PU_RL1_BE) X(PU_RL2_W) X(PU_RL2_B)                         \
  X(PU_PB_W) X(PU_PN_W) X(PU_I1_W) X(PU_I1_B) X(PU_I1_G) X(PU_I1_BE) X(PU_I2_W) X(PU_I2_B)         \
  X(PU_GB_W) X(PU_GN_W) X(PU_GT_W) X(PU_G1_B) X(PU_G1_G) X(PU_G1_BE) X(PU_G2_W) X(PU_G2_B)         \
  /* input-gradient backward (bond predictor only): the same Linears stored [out][in], i.e. [K][N]  */  \
  /* for dX = dY * W; EE split into its h_edge (64) and rbf (16, padded to 32) columns              */  \
  X(T_EEH) X(T_EEG)                                                                                \
  X(T_NB_NN1) X(T_NB_NN2) X(T_NB_EN1) X(T_NB_EN2) X(T_NB_MSG) X(T_NB_GE) X(T_NB_GX) X(T_NB_G2)     \
  X(T_NB_CEN) X(T_NB_OUT)                                                                          \
  X(T_EL_BL) X(T_EL_NL) X(T_EL_I1) X(T_EL_I2) X(T_EL_GB) X(T_EL_GN) X(T_EL_G2)                     \
  X(T_ER_BL) X(T_ER_NL) X(T_ER_I1) X(T_ER_I2) X(T_ER_GB) X(T_ER_GN) X(T_ER_G2)                     \
  X(T_EB_NFL) X(T_EB_NFR) X(T_EB_SELF) X(T_EB_OUT)

enum mdb_block_slot {
#define MDB_X(name) MDB_S_##name,
  MDB_BLOCK_SLOTS(MDB_X)
#undef MDB_X
  MDB_NUM_BLOCK_SLOTS
};

/* Slots of the model "head" (embedders / decoders) that surround the NodeEdgeNet. */
#define MDB_HEAD_SLOTS(X)                                                                          \
  X(RBF_OFFSET) X(RBF_COEFF)           /* {denoiser,encoder}.distance_expansion   common.py:216-237 */ \
  X(TIME_OFFSET) X(TIME_COEFF)         /* time_emb (linear GaussianSmearing)      model.py:34-36    */ \
  X(NODE_EMB_W)                        /* node_embedder  [Kn][256-time_dim]       model.py:32       */ \
  X(EDGE_EMB_W)                        /* edge_embedder  [Ke or 2Kn][64-time_dim] model.py:33, bond_predictor.py:29 */ \
  X(NDEC1_W) X(NDEC1_B) X(NDEC1_G) X(NDEC1_BE) X(NDEC2_W) X(NDEC2_B) /* node_decoder, out padded to 32 */ \
  X(EDEC1_W) X(EDEC1_B) X(EDEC1_G) X(EDEC1_BE) X(EDEC2_W) X(EDEC2_B) /* edge_decoder, out padded to 32 */ \
  X(EDEC1N_W)                          /* bond predictor: node half of edge_decoder.net.0 [256][64] */ \
  X(EDEC3_G) X(EDEC3_BE) X(EDEC3_W) X(EDEC3_B) /* bond predictor: third layer (LN + Linear, out padded to 32) */ \
  X(T_EDEC1) X(T_EDEC1N) X(T_EDEC2) X(T_EDEC3) /* bond predictor backward: edge_decoder Linears stored [out][in] */

enum mdb_head_slot {
#define MDB_X(name) MDB_H_##name,
  MDB_HEAD_SLOTS(MDB_X)
#undef MDB_X
  MDB_NUM_HEAD_SLOTS
};

/* Tensor-core operand images of one block (byte offsets into the tc blob): per-K-stage split-fp16 (hi | lo)
 * planes of the "N x K, K-major" B operand in the UMMA SWIZZLE_NONE canonical layout (csrc/tc_engine.cuh). */
#define MDB_TC_SLOTS(X)                                                                            \
  X(NB_EN1) X(NB_EN2) X(NB_MSG) X(NB_GE) X(NB_G2)          /* NodeBlock per-edge Linears, forward   */ \
  X(BT_NB_G2) X(BT_NB_GE) X(BT_NB_MSG) X(BT_NB_EN2) X(BT_NB_EN1) /* same, transposed use (backward) */ \
  X(PU_PB) X(PU_PN) X(PU_I1) X(PU_GB) X(PU_GN)              /* PosUpdate edge_lin Linears            */ \
  X(EB_SELF) X(EB_OUT)                                      /* EdgeBlock tail Linears (64 -> 64)     */ \
  X(EE)                                                     /* edge_embs (80 -> 64)                  */ \
  X(EL_BL) X(EL_GB) X(EL_I1) X(EL_G2) X(EL_I2)              /* bond_ffn_left per-edge Linears        */ \
  X(ER_BL) X(ER_GB) X(ER_I1) X(ER_G2) X(ER_I2)              /* bond_ffn_right                        */ \
  X(BT_EEH) X(BT_EEG)                                       /* edge_embs, transposed use (backward)  */ \
  X(BT_EL_G2) X(BT_EL_I2) X(BT_EL_GB) X(BT_EL_I1) X(BT_EL_BL) /* bond_ffn_left, transposed use       */ \
  X(BT_ER_G2) X(BT_ER_I2) X(BT_ER_GB) X(BT_ER_I1) X(BT_ER_BL)                                      \
  /* per-node Linears (tc_node_kernel): NodeBlock node path, PosUpdate node MLPs, hoisted projections */ \
  X(NB_OUT) X(PU_LL1) X(PU_LL2) X(PU_RL1) X(PU_RL2) X(NB_NN1) X(NB_NN2) X(NB_GX) X(NB_CEN)         \
  X(EL_NL) X(ER_NL) X(EL_GN) X(ER_GN) X(EB_NFL) X(EB_NFR)                                          \
  /* transposed use in the tensor-core backward of the EdgeBlock tail and of the per-node path (bond predictor only) */ \
  X(BT_EB_OUT) X(BT_EB_SELF)                                                                       \
  X(BT_EB_NFL) X(BT_EB_NFR) X(BT_EL_NL) X(BT_ER_NL) X(BT_EL_GN) X(BT_ER_GN)                        \
  X(BT_NB_GX) X(BT_NB_NN2) X(BT_NB_NN1) X(BT_NB_OUT) X(BT_NB_CEN)

enum mdb_tc_slot {
#define MDB_X(name) MDB_T_##name,
  MDB_TC_SLOTS(MDB_X)
#undef MDB_X
  MDB_NUM_TC_SLOTS
};

/* Tensor-core images of the head Linears applied per node (decoders). */
#define MDB_TC_HEAD_SLOTS(X) X(NDEC1) X(NDEC2) X(EDEC1N) X(BT_EDEC1N)

enum mdb_tc_head_slot {
#define MDB_X(name) MDB_TH_##name,
  MDB_TC_HEAD_SLOTS(MDB_X)
#undef MDB_X
  MDB_NUM_TC_HEAD_SLOTS
};

/* Static description of one packed network (host struct, passed by pointer). */
typedef struct mdb_net_desc {
  const float* blob;            /* device: packed fp32 weights                                   */
  int32_t num_blocks;           /* 6 (denoiser) / 8 (bond predictor)                              */
  int32_t update_pos;           /* NodeEdgeNet.update_pos                  graph.py:324-327       */
  float rbf_start, rbf_stop;    /* GaussianSmearing clamp = (start, cutoff) graph.py:316          */
  int32_t time_dim;             /* config.diff.time_dim (10 / 20)                                 */
  int32_t num_node_types;       /* width of h_node_pert                                           */
  int32_t num_edge_types;       /* width of h_edge_pert (MolDiff) / of the logits (BondPredictor) */
  float num_timesteps;          /* T; time features are t / T               model.py:221-222      */
  int32_t kind;                 /* 0 = bare NodeEdgeNet, 1 = MolDiff, 2 = BondPredictor           */
  int64_t head_off[MDB_NUM_HEAD_SLOTS];                        /* float offsets into blob, -1 = absent */
  int64_t block_off[MDB_MAX_BLOCKS][MDB_NUM_BLOCK_SLOTS];
  const float* blob_host;       /* host copy of `blob` (bias / LayerNorm vectors are passed to the tensor-core   */
                                /* kernels by value, i.e. through the constant bank); required when tc_blob set  */
  const void* tc_blob;          /* device: tensor-core operand images, or NULL = fp32 FFMA path only    */
  int64_t tc_block_off[MDB_MAX_BLOCKS][MDB_NUM_TC_SLOTS];      /* byte offsets into tc_blob, -1 = absent */
  int64_t tc_head_off[MDB_NUM_TC_HEAD_SLOTS];
} mdb_net_desc;

/*
 * Per-batch graph plan (constant over all T sampling steps).  Edges are processed in CSR order
 * (sorted by left node, ties by right node); `perm[q]` = index into the caller's edge_index of the
 * q-th sorted edge.  The host builds it once per batch (moldiff_b200/engine.py: GraphPlan) -- it is
 * plumbing outside the T-step loop, so a device sort from the host framework is used.
 */
typedef struct mdb_plan {
  int32_t n_nodes, n_edges, n_half; /* n_half = E/2 when edges are (half, flipped half) pairs, else 0 */
  const int32_t* left;     /* [E] sorted left node  */
  const int32_t* right;    /* [E] sorted right node */
  const int32_t* perm;     /* [E] sorted position -> original edge id */
  const int32_t* inv;      /* [E] original edge id  -> sorted position */
} mdb_plan;

/* Bytes of fp32 workspace one forward (with_backward = 0) or one guidance call (= 1) needs. */
size_t mdb_workspace_bytes(int64_t n_nodes, int64_t n_edges, int32_t with_backward, int32_t num_blocks);

/* NodeEdgeNet.forward (graph.py:348-367).  h_edge_in/out are in the CALLER's edge order.
 * node_time [N], edge_time [E] (the reference passes [N,1]/[E,1]).  pos_out may alias nothing. */
int mdb_net_forward(const mdb_net_desc* net, const mdb_plan* plan,
                    const float* h_node_in, const float* pos_in, const float* h_edge_in,
                    const float* node_time, const float* edge_time,
                    float* h_node_out, float* pos_out, float* h_edge_out,
                    float* workspace, size_t workspace_bytes, void* stream);

/* MolDiff.forward (model.py:204-234): h_node_pert [N][Kn], pos_pert [N][3], h_edge_pert [E][Ke]
 * (caller order; the two halves are what the caller concatenated), batch_node int64 [N], batch_edge
 * int64 [E], t int64 [n_graphs] -> pred_node [N][Kn], pred_pos [N][3], pred_halfedge [E/2][Ke]. */
int mdb_moldiff_forward(const mdb_net_desc* net, const mdb_plan* plan,
                        const float* h_node_pert, const float* pos_pert, const float* h_edge_pert,
                        const int64_t* batch_node, const int64_t* batch_edge, const int64_t* t,
                        float* pred_node, float* pred_pos, float* pred_halfedge,
                        float* workspace, size_t workspace_bytes, void* stream);

/* BondPredictor.forward (bond_predictor.py:128-162): h_node [N][Kn] (one-hot or soft), pos [N][3]
 * -> logits [E/2][num_edge_types].  save_for_backward != 0 keeps every block's inputs (h_node, h_edge)
 * in the workspace (sized with with_backward = 1) for mdb_bondpred_backward. */
int mdb_bondpred_forward(const mdb_net_desc* net, const mdb_plan* plan,
                         const float* h_node, const float* pos,
                         const int64_t* batch_node, const int64_t* batch_edge, const int64_t* t,
                         float* logits, int32_t save_for_backward,
                         float* workspace, size_t workspace_bytes, void* stream);

/* d(sum(logits * d_logits)) / d pos -- the input gradient the sampling guidance takes through the bond
 * predictor (models/model.py:312-325: torch.autograd.grad(objective, pos_in)).  Must follow a
 * mdb_bondpred_forward(save_for_backward = 1) on the same workspace and inputs.  Hand-written backward
 * kernels: per-edge activations are recomputed tile by tile, nothing but block inputs is stored.
 * d_logits [E/2][num_edge_types] -> d_pos [N][3]. */
int mdb_bondpred_backward(const mdb_net_desc* net, const mdb_plan* plan,
                          const float* h_node, const float* pos,
                          const int64_t* batch_node, const int64_t* batch_edge, const int64_t* t,
                          const float* d_logits, float* d_pos,
                          float* workspace, size_t workspace_bytes, void* stream);

/* Kernel classes, for the per-kernel device timing bench.py reports (order is the ABI). */
#define MDB_KERNEL_CLASSES(X)                                                                      \
  X(node_init) X(edge_init) X(node) X(edge_b) X(edge_d) X(edge_decode) X(edge_unsort)              \
  X(bwd_decode) X(bwd_node) X(bwd_edge_tail) X(bwd_edge_nodeblock) X(bwd_edge_bondffn) X(bwd_pos)   \
  X(tc_nodeblock) X(tc_nodeblock_bwd) X(tc_edge_d) X(tc_bondffn) X(tc_bondffn_bwd) X(tc_node) X(transition) X(graph_build) X(decode_rows) \
  X(tc_edge_tail_bwd) X(tc_node_bwd)

enum mdb_kernel_class {
#define MDB_X(name) MDB_K_##name,
  MDB_KERNEL_CLASSES(MDB_X)
#undef MDB_X
  MDB_NUM_KERNEL_CLASSES
};

/*
 * mdb_transition_step <- the posterior-sampling block of MolDiff.sample between two denoiser evaluations
 *   (models/model.py:287-300,365-372; models/transition.py:44-63,285-315; models/diffusion.py:79-85), discrete
 *   categorical space: positions via the Gaussian posterior, node / half-edge types via the categorical posterior +
 *   Gumbel-max, one-hot encodings of the sampled classes (half-edges doubled to the [2 Eh, Ke] directed list).
 * t [n_graphs] int64; pos, pred_pos, z_pos [N,3]; pred_node, log_node, u_node [N,Kn]; pred_half, log_half, u_half
 * [Eh,Ke]; coef_x0, coef_xt, std [T]; q*_cum = q_mats [T,K,K]; q*_stepT = transpopse_q_onestep_mats [T,K,K].
 * z_pos ~ N(0,1) and u_* ~ U[0,1) are drawn by the caller.  Outputs: pos_out [N,3], log_node_out / h_node_out [N,Kn],
 * log_half_out [Eh,Ke], h_edge_out [2 Eh,Ke], half_type_out [Eh] int64 (may be NULL).  Kn, Ke <= 16.
 */
int mdb_transition_step(int32_t n_nodes, int32_t n_half, int32_t kn, int32_t ke, const int64_t* batch_node,
                        const int64_t* batch_half, const int64_t* t, const float* pos, const float* pred_pos,
                        const float* z_pos, const float* coef_x0, const float* coef_xt, const float* std_, float* pos_out,
                        const float* pred_node, const float* log_node, const float* u_node, const float* qn_cum,
                        const float* qn_stepT, float* log_node_out, float* h_node_out, const float* pred_half,
                        const float* log_half, const float* u_half, const float* qe_cum, const float* qe_stepT,
                        float* log_half_out, float* h_edge_out, int64_t* half_type_out, void* stream);

/*
 * mdb_decode_rows <- the softmax / argmax / max arithmetic of FeaturizeMol.decode_output (utils/transforms.py:76-96) for a
 * whole batch: node_type[i] = argmax_k pred_node[i,k], node_prob[i] = softmax(pred_node[i])[node_type[i]], same for the
 * half-edges.  uint8 classes, fp32 probabilities.  The caller (moldiff_b200/decode.py) does the index bookkeeping
 * (masked atoms, bond filtering, per-molecule split of utils/sample.py:4-30) on the host.
 */
int mdb_decode_rows(int32_t n_nodes, int32_t kn, const float* pred_node, int32_t n_half, int32_t ke, const float* pred_half,
                    uint8_t* node_type, float* node_prob, uint8_t* half_type, float* half_prob, void* stream);

/*
 * Edge builders over atom positions (torch_geometric.nn.radius_graph / knn_graph, imported by models/graph.py:6 and
 * reached only from dead code there -- SURVEY.md 8f N4).  Nodes of one graph must be contiguous: seg_lo[i] / seg_hi[i]
 * are the first / one-past-last node index of node i's graph.  Output: counts[i] neighbours of centre i in
 * neighbors[i * max .. ) (max = max_num_neighbors or k): radius -- every j with |pos_j - pos_i| < radius in index
 * order, truncated at max_num_neighbors; knn -- the k nearest, ascending distance, ties by index.  loop != 0 keeps j == i.
 */
int mdb_radius_graph(int32_t n_nodes, const float* pos, const int32_t* seg_lo, const int32_t* seg_hi, float radius,
                     int32_t loop, int32_t max_num_neighbors, int32_t* counts, int32_t* neighbors, void* stream);
int mdb_knn_graph(int32_t n_nodes, const float* pos, const int32_t* seg_lo, const int32_t* seg_hi, int32_t k, int32_t loop,
                  int32_t* counts, int32_t* neighbors, void* stream);

/*
 * Range check of the tensor-core operands.  The split-fp16 operand planes saturate at +-65504 (cvt.satfinite); every GEMM input
 * is a LayerNorm / ReLU / sigmoid-bounded activation EXCEPT the two residual streams and what is derived from them linearly.
 * After a forward on `workspace` (with_backward = 0 layout) this returns max |v| of h_node, h_edge, e = edge_embs(.) and the
 * node_net table of the LAST block in amax4[0..3] (device floats; NaN / Inf report as Inf), so that a caller can refuse to
 * trust a checkpoint whose activations leave the fp16 range instead of saturating silently (moldiff_b200/engine.py:
 * check_operand_range, called by MolDiff.sample on the first step).
 */
int mdb_operand_amax(int64_t n_nodes, int64_t n_edges, const float* workspace, float* amax4, void* stream);

/* Profiling: between begin and end every kernel launch of this library is bracketed by CUDA events on
 * its own stream; end synchronises those events and returns summed milliseconds / launch counts per
 * kernel class (arrays of MDB_NUM_KERNEL_CLASSES).  Not for use inside a timed region. */
void mdb_profile_begin(void);
int mdb_profile_end(double* ms_per_class, int64_t* launches_per_class);
const char* mdb_kernel_class_name(int cls);
int mdb_num_kernel_classes(void);

/* Self-test of the tcgen05 GEMM pipeline (one 128-row tile): y[128][n] = x[128][k] * W, W given as the packed
 * split-fp16 stage images of moldiff_b200/packing.py:tc_image; `twice` bit 0 accumulates the product twice, bit 1 selects
 * the cross-first accumulation order (csrc/tc_pipe.cuh) that the bond predictor's forward kernels use. */
int mdb_tc_selftest(const float* x, const void* w_img, float* y, int32_t k, int32_t n, int32_t twice, void* stream);

/* Debug: device int64 buffer [n_tiles][32] that tc_nodeblock_fwd_kernel fills with clock64() phase stamps of its row
 * thread 0 (NULL = off, the default).  tools/tc_phase_times.py prints the per-phase breakdown. */
void mdb_debug_set_buffer(void* device_i64_buffer);
/* Which kernel writes the stamps: 0 = NodeBlock forward (default), 1 = NodeBlock backward, 2 = BondFFN forward. */
void mdb_debug_select(int32_t kernel);

/* Diagnostics. */
const char* mdb_last_error(void);
int mdb_version(void);
/* Number of kernel launches issued by this library since load (bench.py's gpu_launches counter). */
int64_t mdb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MOLDIFF_B200_H */
