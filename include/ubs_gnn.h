/* ubs_gnn.h — C ABI of the B200-native hetero-graph message-passing hot path of uav_bs_ctrl.
 *
 * Drop-in boundary: these entry points replace what the reference delegates to
 * DGL 0.9.0 / PyTorch 1.12 from
 *   - dglnn.GATv2Conv.forward          (call sites algos/madrqn/agents/gnn_agents.py:92-97,103-104,
 *                                        algos/drqn/agents/gnn_agents.py:17-18,27)
 *   - TarMAC.forward's edge ops         (algos/madrqn/agents/gnn_agents.py:261-266: apply_edges(u_dot_v),
 *                                        edge_softmax, update_all(u_mul_e, sum))
 *   - nn.GRUCell                        (gnn_agents.py:29,246,270; drqn gnn_agents.py:20,28)
 * and are what a ctypes / pybind binding on the reference side would bind.
 *
 * Conventions
 *   - plain C: pointers + sizes only; every pointer is a DEVICE pointer to fp32 / int32 / uint32 data
 *     owned by the caller (PyTorch's caching allocator in our host code); nothing is allocated, freed or
 *     retained past return; workspaces are caller-provided.
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it (no device sync).
 *   - return value 0 = success; non-zero = failure, message via ubs_last_error() (thread-local).
 *   - graphs are CSR-by-destination: in-edges of dst v are slots [indptr[v], indptr[v+1]);
 *     src_idx[slot] is the source node, or src_idx == NULL for the star layout (source id == slot id).
 *   - feature index of (head k, channel d) is k*D + d (head-major, what `.view(N,-1)` flattens,
 *     gnn_agents.py:103-104).
 */
#ifndef UBS_GNN_H_
#define UBS_GNN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UBS_GNN_VERSION 100

#if defined(__GNUC__)
#define UBS_API __attribute__((visibility("default")))
#else
#define UBS_API
#endif

/* flags for ubs_gatv2_* */
#define UBS_GAT_RESIDUAL 1 /* res_fc is a Linear(F_d, H) (always the case at the reference call sites) */
#define UBS_GAT_RELU 2     /* activation=nn.ReLU() */

UBS_API int ubs_version(void);
UBS_API const char* ubs_last_error(void);
/* Number of kernel launches issued through this library by the calling process (bench "gpu_launches"). */
UBS_API int64_t ubs_launch_count(void);
UBS_API void ubs_reset_launch_count(void);
/* Replaying a captured CUDA graph re-launches kernels without passing through this library: the host code that
 * replays a graph adds the number of library kernels the graph contains.                                        */
UBS_API void ubs_add_launch_count(int64_t n);

/* ---- GATv2 relation, fused projection (F_s <= 4, F_d <= 2, H = heads*D in {32,64,128}) -----------------
 * out[v, k, :] = act( sum_e alpha[e,k] * (W_src x_src[u_e] + b_src)[k,:] + W_res x_dst[v] + b_res )
 * alpha = edge_softmax_by_dst( attn[k,:] . leaky_relu(W_src x_u + b_src + W_dst x_v + b_dst)[k,:] )
 * smax / ssum (n_dst*heads each, may both be NULL for inference) receive the per-(dst,head) softmax max and
 * denominator that the backward needs.                                                                     */
UBS_API int ubs_gatv2_fwd(const float* x_src, const float* x_dst, const int32_t* indptr, const int32_t* src_idx,
                  const float* W_src, const float* b_src, const float* W_dst, const float* b_dst,
                  const float* attn, const float* W_res, const float* b_res,
                  float* out, float* smax, float* ssum,
                  int64_t n_dst, int64_t n_edges, int F_s, int F_d, int heads, int D,
                  float negative_slope, int flags, void* stream);

/* Size in floats of the workspace ubs_gatv2_bwd needs (per-CTA partial parameter gradients). */
UBS_API int64_t ubs_gatv2_bwd_workspace(int64_t n_dst, int F_s, int F_d, int heads, int D);

/* Backward of ubs_gatv2_fwd.  grad_params is ONE flat buffer laid out as
 *   [gW_src (H*F_s) | gb_src (H) | gW_dst (H*F_d) | gb_dst (H) | gattn (H) | gW_res (H*F_d) | gb_res (H)]
 * and is overwritten (not accumulated).  grad_x_src (n_src*F_s) / grad_x_dst (n_dst*F_d) may be NULL when the
 * observations are leaves (always the case in the env configs); when given, grad_x_src must be zero-filled by
 * the caller if src_idx != NULL (several edges may share a source).  Deterministic: two-stage reduction.    */
UBS_API int ubs_gatv2_bwd(const float* x_src, const float* x_dst, const int32_t* indptr, const int32_t* src_idx,
                  const float* W_src, const float* b_src, const float* W_dst, const float* b_dst,
                  const float* attn, const float* W_res, const float* b_res,
                  const float* out, const float* grad_out, const float* smax, const float* ssum,
                  float* grad_params, float* grad_x_src, float* grad_x_dst, float* workspace,
                  int64_t n_dst, int64_t n_edges, int64_t n_src, int F_s, int F_d, int heads, int D,
                  float negative_slope, int flags, void* stream);

/* Strided-segment variants: the same kernels over n_seg graphs laid out at fixed strides (one per timestep of a
 * sequence arena), in ONE launch.  Segment s reads x_src + s*st_xsrc, x_dst + s*st_xdst, indptr + s*st_ip (segment-
 * local offsets), src_idx + s*st_sidx (strides in elements); destination v of segment s is output row s*n_dst_seg + v
 * with row stride ld_out (so two relations can write the two halves of one (rows, 2H) buffer); smax / ssum rows
 * likewise.  n_edges is only a hint (mean degree) for the launch shape.                                          */
UBS_API int ubs_gatv2_seg_fwd(const float* x_src, const float* x_dst, const int32_t* indptr, const int32_t* src_idx,
                      const float* W_src, const float* b_src, const float* W_dst, const float* b_dst,
                      const float* attn, const float* W_res, const float* b_res,
                      float* out, float* smax, float* ssum, int64_t n_seg, int64_t n_dst_seg, int64_t n_edges,
                      int64_t st_xsrc, int64_t st_xdst, int64_t st_ip, int64_t st_sidx, int64_t ld_out,
                      int F_s, int F_d, int heads, int D, float negative_slope, int flags, void* stream);
UBS_API int ubs_gatv2_seg_bwd(const float* x_src, const float* x_dst, const int32_t* indptr, const int32_t* src_idx,
                      const float* W_src, const float* b_src, const float* W_dst, const float* b_dst,
                      const float* attn, const float* W_res, const float* b_res,
                      const float* out, const float* grad_out, const float* smax, const float* ssum,
                      float* grad_params, float* grad_x_src, float* grad_x_dst, float* workspace,
                      int64_t n_seg, int64_t n_dst_seg, int64_t n_edges, int64_t st_xsrc, int64_t st_xdst,
                      int64_t st_ip, int64_t st_sidx, int64_t ld_out, int64_t ld_gout,
                      int F_s, int F_d, int heads, int D, float negative_slope, int flags, void* stream);

/* Training pair with SAVED ATTENTION SCORES: the forward additionally writes the raw score of every (edge slot, head) —
 * scores[(s * st_scores + e) * heads + k] for CSR slot e of segment s (16 bytes per edge at 4 heads) — and the backward
 * reads them instead of recomputing the H-channel score of every edge (its pass over the edges then only forms
 * exp(score - max) / sum and the softmax gradient).  scores == NULL: same as ubs_gatv2_seg_fwd / _bwd.  16-byte aligned. */
UBS_API int ubs_gatv2_seg_fwd_scores(const float* x_src, const float* x_dst, const int32_t* indptr, const int32_t* src_idx,
                      const float* W_src, const float* b_src, const float* W_dst, const float* b_dst,
                      const float* attn, const float* W_res, const float* b_res,
                      float* out, float* smax, float* ssum, float* scores, int64_t st_scores,
                      int64_t n_seg, int64_t n_dst_seg, int64_t n_edges,
                      int64_t st_xsrc, int64_t st_xdst, int64_t st_ip, int64_t st_sidx, int64_t ld_out,
                      int F_s, int F_d, int heads, int D, float negative_slope, int flags, void* stream);
UBS_API int ubs_gatv2_seg_bwd_scores(const float* x_src, const float* x_dst, const int32_t* indptr, const int32_t* src_idx,
                      const float* W_src, const float* b_src, const float* W_dst, const float* b_dst,
                      const float* attn, const float* W_res, const float* b_res,
                      const float* out, const float* grad_out, const float* smax, const float* ssum,
                      const float* scores, int64_t st_scores,
                      float* grad_params, float* grad_x_src, float* grad_x_dst, float* workspace,
                      int64_t n_seg, int64_t n_dst_seg, int64_t n_edges, int64_t st_xsrc, int64_t st_xdst,
                      int64_t st_ip, int64_t st_sidx, int64_t ld_out, int64_t ld_gout,
                      int F_s, int F_d, int heads, int D, float negative_slope, int flags, void* stream);

/* ---- GATv2 attention + aggregation on pre-projected features, general CSR (wide inputs / synthetic sweep) ---------
 * After the library GEMMs el = fc_src(h_src) (n_src,H), er = fc_dst(h_dst) (n_dst,H), res = res_fc(h_dst) (nullable):
 *   out[v] = act( sum_e softmax_e( <attn_k, leaky_relu(el[u_e] + er[v])_k> ) * el[u_e] + res[v] ),  H = heads*D in
 * {32,64,128,256}.  One gather pass over the edges (coalesced full-row loads, online softmax); nothing of size E x H
 * is materialised.  Backward: grad_el (n_src,H; the caller zero-fills it when src_idx != NULL — vector atomics),
 * grad_er (n_dst,H), grad_res (n_dst,H, nullable) = grad_out * 1[out > 0], grad_attn (H) via a fixed-order two-stage
 * reduction (workspace: ubs_gat_aggr_bwd_workspace floats).                                                      */
UBS_API int ubs_gat_aggr_fwd(const float* el, const float* er, const float* res, const int32_t* indptr,
                     const int32_t* src_idx, const float* attn, float* out, float* smax, float* ssum,
                     int64_t n_dst, int64_t n_edges, int heads, int D, float negative_slope, int flags, void* stream);
UBS_API int64_t ubs_gat_aggr_bwd_workspace(int64_t n_dst, int heads, int D);
UBS_API int ubs_gat_aggr_bwd(const float* el, const float* er, const float* res, const int32_t* indptr,
                     const int32_t* src_idx, const float* attn, const float* out, const float* grad_out,
                     const float* smax, const float* ssum, float* grad_el, float* grad_er, float* grad_res,
                     float* grad_attn, float* workspace, int64_t n_dst, int64_t n_edges, int heads, int D,
                     float negative_slope, int flags, void* stream);

/* ---- Dense projection on the tensor cores: C[M,N] = act(A[M,K] W[N,K]^T + bias), fp32 in/out, 3xTF32 -----------
 * tcgen05.mma.kind::tf32 with TMEM accumulators; every operand is split hi + lo (13 low mantissa bits) and
 * hi*hi + hi*lo + lo*hi is accumulated in fp32, so the result is fp32-accurate (~2^-21) — single-pass TF32 would
 * break the 1e-5 parity bar.  A, W row-major (nn.Linear layout), K % 32 == 0, N % 16 == 0, N <= 256, 16-byte aligned
 * rows.  Returns 3 if W does not fit shared memory.  relu != 0 applies ReLU in the epilogue.                        */
UBS_API int ubs_tf32x3_gemm(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                    float* C, int64_t ldc, int64_t M, int N, int K, int relu, void* stream);

/* Weight-gradient products of the window:  C[Mo, No] = A[R, Mo]^T . B[R, No]  (fp32 in / out, 3xTF32 on tcgen05,
 * reduction over the R = T*N rows split into 256-row slices whose partial tiles are added by a second kernel in a
 * fixed order).  Replaces the `grad_out.t() @ input` GEMMs autograd runs for nn.Linear / GRUCell weights.
 * workspace: ubs_tf32x3_gemm_tn_workspace(R, Mo, No) floats, 16-byte aligned.  No multiple of 16, <= 128.          */
UBS_API int64_t ubs_tf32x3_gemm_tn_workspace(int64_t R, int Mo, int No);
UBS_API int ubs_tf32x3_gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                               float* workspace, int64_t R, int Mo, int No, void* stream);

/* Bias gradients of the window: out[c] = sum over the R = T*N rows of X[r, c] — what autograd's `grad.sum(0)` computes
 * for the biases of nn.Linear / GRUCell (reference algos/madrqn/agents/gnn_agents.py:65-67,101-107 run through
 * learner.py:157 `loss.backward()`).  Streaming two-stage reduction in a fixed order (deterministic).  C % 4 == 0,
 * C <= 1024, 16-byte aligned rows; workspace: ubs_colsum_workspace(C) floats.
 * ubs_relu_bwd_colsum additionally applies the ReLU backward of the aggregator layer in the same pass:
 * dX = X * (Y > 0) (dX may alias X), out = column sums of dX.                                                      */
UBS_API int64_t ubs_colsum_workspace(int C);
UBS_API int ubs_colsum(const float* X, int64_t ldx, int64_t R, int C, float* out, float* workspace, void* stream);
UBS_API int ubs_relu_bwd_colsum(const float* X, int64_t ldx, const float* Y, int64_t ldy, float* dX, int64_t ldd,
                                int64_t R, int C, float* out, float* workspace, void* stream);

/* ---- TarMAC attention over block-diagonal comm graphs ----------------------------------------------------
 * Nodes are grouped in consecutive blocks of `block` (= agents per env, <= 32) nodes; every edge stays inside a
 * block (batched per-env graphs).  mask[v] bit i set  <=>  edge (block_start(v)+i) -> v exists.
 *   e[u->v] = <s_u, q_v> * scale ; a = softmax over in-edges of v ; c_v = sum_u a[u->v] * val_u
 * s/q/val are row-strided views (ld_* in floats) so they may alias one (N, M+2K) projection buffer.
 * alpha (N*block, dense per dst, 0 where no edge) is written for the backward.                              */
UBS_API int ubs_block_attn_fwd(const float* s, int64_t ld_s, const float* q, int64_t ld_q, const float* val, int64_t ld_v,
                       const uint32_t* mask, float* c, float* alpha,
                       int64_t n_nodes, int block, int key_size, int msg_size, float scale, void* stream);

/* ds_work: workspace of n_nodes*block floats. grad_s / grad_q / grad_val are row-strided like the inputs and are
 * overwritten.                                                                                              */
UBS_API int ubs_block_attn_bwd(const float* s, int64_t ld_s, const float* q, int64_t ld_q, const float* val, int64_t ld_v,
                       const uint32_t* mask, const float* alpha, const float* grad_c,
                       float* grad_s, int64_t ld_gs, float* grad_q, int64_t ld_gq, float* grad_val, int64_t ld_gv,
                       float* ds_work, int64_t n_nodes, int block, int key_size, int msg_size, float scale,
                       void* stream);

/* ---- Mean aggregation over block-diagonal comm graphs (BaseComm / CommNet reduce step) ---------------------
 * c_v = mean over the in-neighbours u of v of msg[u] (0 without in-edges): update_all(udf_msg, mailbox.mean(1)) of
 * gnn_agents.py:125-133,144 / :205-214,224 for messages that depend on the source node only.  mask / block as in
 * ubs_block_attn_fwd.  msg / out / grads are row-strided (ld_* in floats).  Deterministic (no atomics).        */
UBS_API int ubs_block_mean_fwd(const float* msg, int64_t ld_msg, const uint32_t* mask, float* out, int64_t ld_out,
                               int64_t n, int block, int F, void* stream);
UBS_API int ubs_block_mean_bwd(const float* grad_out, int64_t ld_go, const uint32_t* mask, float* grad_msg, int64_t ld_gm,
                               int64_t n, int block, int F, void* stream);

/* ---- DiscreteComm message passing over block-diagonal comm graphs (gnn_agents.py:151-193) --------------------------
 * Per edge u->v and bit k: hard Gumbel-softmax (tau) of the source's encoder logits (n, 2*msg_size: f_enc output per
 * NODE) plus per-EDGE noise; per destination: element-wise max (OR) over the in-edges, zeros without in-edges.
 * expo (E, msg_size, 2): Exponential(1) draws in the reference's edge-id order (gumbel = -log expo: what
 * F.gumbel_softmax consumes); env_edge_offset (n / block) int64: first edge id of every env (exclusive prefix sum of
 * the envs' edge counts); mask / block as in ubs_block_attn_fwd.  winner (n, 2*msg_size) uint8 (NULL for inference):
 * local source index of the first maximal mailbox entry per (destination, feature) — torch.max(dim)'s one-winner
 * gradient — consumed by the backward, which writes grad_logits (n, 2*msg_size), deterministic, no atomics.      */
UBS_API int ubs_block_bitmax_fwd(const float* logits, int64_t ld_logits, const float* expo, const uint32_t* mask,
                                 const int64_t* env_edge_offset, float* out, int64_t ld_out, uint8_t* winner,
                                 int64_t n, int block, int msg_size, float tau, void* stream);
UBS_API int ubs_block_bitmax_bwd(const float* logits, int64_t ld_logits, const float* expo, const uint32_t* mask,
                                 const int64_t* env_edge_offset, const uint8_t* winner, const float* grad_out,
                                 int64_t ld_go, float* grad_logits, int64_t ld_gl, int64_t n, int block, int msg_size,
                                 float tau, void* stream);

/* ---- GRUCell gate math (nn.GRUCell, gate order r,z,n) given gi = W_ih x + b_ih, gh = W_hh h + b_hh --------- */
UBS_API int ubs_gru_gates_fwd(const float* gi, const float* gh, const float* h, float* h_out, int64_t n, int H, void* stream);
/* grad_gi, grad_gh: (n, 3H); grad_h_direct: (n, H) = grad_out * z (the part that does not go through W_hh). */
UBS_API int ubs_gru_gates_bwd(const float* gi, const float* gh, const float* h, const float* grad_out,
                      float* grad_gi, float* grad_gh, float* grad_h_direct, int64_t n, int H, void* stream);

/* ---- Fused recurrent agent step / sequence -------------------------------------------------------------------
 * Everything reference GnnAgent.forward does after the GATv2 relations (algos/madrqn/agents/gnn_agents.py:51-56):
 * f_aggr Linear(Fin,H)+ReLU (:99,106-107) [UBS_STEP_AGGR], TarMAC.forward (:248-271) [UBS_STEP_TARMAC] or a plain
 * GRUCell (:29,55; drqn gnn_agents.py:20,28), and the Linear Q head (:43-46).  One launch walks n_steps timesteps
 * of every row tile (<= 16 agent rows = whole envs; envs never exchange data), hidden state resident on chip.
 *   xin    (n_steps, n_rows, Fin)   [x_gt | x_ubs] (Fin = 2H) with AGGR, else the H-wide encoder output
 *   h0     (n_rows, H)              hidden state entering step 0
 *   mask   (n_steps, n_rows) uint32 talk-graph block masks (TARMAC), see ubs_block_attn_fwd
 *   h_out  (n_steps, n_rows, H), q (n_steps, n_rows, A), actions (n_steps, n_rows) int64 argmax (nullable)
 *   sv_*   saved activations for the backward (all NULL => inference):
 *          sv_xc (n_steps,n_rows,H[+M]) [x | c], sv_vsq (.., round4(M+2K)), sv_alpha (.., U), sv_gate (.., 4H)
 * `packed` comes from ubs_agent_pack (transposed + original copies of every weight, ubs_agent_pack_size floats)
 * and must be rebuilt whenever a parameter changes.                                                            */
#define UBS_STEP_AGGR 1
#define UBS_STEP_TARMAC 2
UBS_API int64_t ubs_agent_pack_size(int H, int M, int K, int A, int U, int Fin, int flags);
UBS_API int ubs_agent_pack(int H, int M, int K, int A, int U, int Fin, int flags,
                   const float* W_aggr, const float* b_aggr, const float* W_val, const float* b_val,
                   const float* W_sign, const float* b_sign, const float* W_que, const float* b_que,
                   const float* W_ih, const float* b_ih, const float* W_hh, const float* b_hh,
                   const float* W_out, const float* b_out, float* packed, void* stream);
UBS_API int ubs_agent_seq_fwd(int H, int M, int K, int A, int U, int Fin, int flags, const float* packed,
                      const float* xin, const float* h0, const uint32_t* mask,
                      float* h_out, float* q, int64_t* actions,
                      float* sv_xc, float* sv_vsq, float* sv_alpha, float* sv_gate,
                      int64_t n_rows, int n_steps, void* stream);
/* Same kernel with epsilon-greedy action selection fused in (reference learner.act, algos/madrqn/learner.py:75-78):
 * actions[t,r] = eg_u[t,r] <= *eg_eps ? eg_a[t,r] : argmax_a q[t,r,a].  eg_u (n_steps,n_rows) uniforms (one draw per
 * env, repeated for its agents), eg_a (n_steps,n_rows) int64 random actions, eg_eps device scalar; all NULL = greedy. */
UBS_API int ubs_agent_act_fwd(int H, int M, int K, int A, int U, int Fin, int flags, const float* packed,
                      const float* xin, const float* h0, const uint32_t* mask,
                      float* h_out, float* q, int64_t* actions,
                      const float* eg_u, const int64_t* eg_a, const float* eg_eps,
                      float* sv_xc, float* sv_vsq, float* sv_alpha, float* sv_gate,
                      int64_t n_rows, int n_steps, void* stream);
/* 1 if an inference call (no save buffers) of these dimensions runs the kernel that stages each layer's weights into
 * shared memory with bulk async copies (TMA) one layer ahead of the GEMM that consumes them, 0 if it streams weights
 * through registers (the configuration does not fit 227 KB, or UBS_ACT_TMA=0). */
UBS_API int ubs_agent_act_uses_tma(int H, int M, int K, int A, int U, int Fin, int flags);
/* ---- Act vector-step as ONE kernel: observation relations + agent step ------------------------------------------
 * ubs_agent_act_fwd with the two GATv2 star relations of GraphObservationEncoder (gnn_agents.py:92-97,103-104:
 * 'seen' gt->agent, 'near' ubs->agent) folded in, i.e. everything learner.act (algos/madrqn/learner.py:69-80) runs
 * per env step: GnnAgent.forward + argmax + epsilon-greedy.  The relations are read straight from an observation
 * packet: x_gt (rows, F_gt) / x_ubs (rows, F_ubs) compacted source rows in star layout (row id == CSR slot), ip_seen /
 * ip_near (n_rows + 1) cumulative in-degrees, x_agent (n_rows, F_d).  A CTA owns 16 consecutive agent rows whose source
 * rows are one contiguous range, staged into shared memory by bulk async copies (TMA); cap_gt / cap_ubs = the largest
 * possible in-degree (G, U-1), which sizes the staging area.  relpack = the per-relation constant tables of
 * ubs_gatv2_rel_pack ('seen' then 'near', ubs_gatv2_rel_pack_size floats each), rebuilt whenever a parameter of the
 * relation encoders changes.  packed / relpack / x_gt / x_ubs must be 16-byte aligned.
 * ubs_agent_act_rel_supported: 1 when the configuration fits the kernel (H = 32/64-class configs whose weights fit
 * 227 KB), else 0 — the caller then runs ubs_gatv2_seg_fwd x 2 + ubs_agent_act_fwd.
 * gat_flags: UBS_GAT_* of the relation encoders, optionally | UBS_ACT_PDL: launch with programmatic stream
 * serialization, so that inside a rollout (act, act, ... or act, env step, pack, act, ...) the grid's start-up — barrier
 * init, the first weight layer's bulk copy — overlaps the tail of the previous kernel; the kernel reads the hidden
 * state and the packet only after `griddepcontrol.wait`.  The caller guarantees that the kernel launched just before
 * on the same stream does not write `packed` (i.e. never set it right after ubs_agent_pack).                          */
#define UBS_ACT_PDL 0x100
UBS_API int64_t ubs_gatv2_rel_pack_size(int heads, int D);
UBS_API int ubs_gatv2_rel_pack(const float* W_src, const float* b_src, const float* W_dst, const float* b_dst,
                               const float* attn, const float* W_res, const float* b_res, int F_s, int F_d, int heads,
                               int D, float negative_slope, int flags, float* out, void* stream);
UBS_API int ubs_agent_act_rel_supported(int H, int M, int K, int A, int U, int flags, int heads, int F_gt, int cap_gt,
                                        int F_ubs, int cap_ubs, int F_d);
UBS_API int ubs_agent_act_rel_fwd(int H, int M, int K, int A, int U, int flags, const float* packed, const float* relpack,
                                  const float* x_gt, const int32_t* ip_seen, int F_gt, int cap_gt,
                                  const float* x_ubs, const int32_t* ip_near, int F_ubs, int cap_ubs,
                                  const float* x_agent, int F_d, int heads, int gat_flags,
                                  const float* h0, const uint32_t* mask, float* h_out, float* q, int64_t* actions,
                                  const float* eg_u, const int64_t* eg_a, const float* eg_eps, int64_t n_rows,
                                  void* stream);
/* Reverse-time walk.  dq (n_steps,n_rows,A) and dh_last (n_rows,H, nullable) come from the loss; outputs:
 * d_xin (n_steps,n_rows,Fin), d_h0 (nullable) and the stashes st_dgi/st_dgh (.., 3H), st_dvsq (.., round4(M+2K)),
 * st_dpre (.., H) from which the caller forms the parameter gradients with batched GEMMs over the whole sequence. */
UBS_API int ubs_agent_seq_bwd(int H, int M, int K, int A, int U, int Fin, int flags, const float* packed,
                      const float* h0, const float* h_out, const float* sv_xc, const float* sv_vsq,
                      const float* sv_alpha, const float* sv_gate, const float* dq, const float* dh_last,
                      float* d_xin, float* d_h0, float* st_dgi, float* st_dgh, float* st_dvsq, float* st_dpre,
                      int64_t n_rows, int n_steps, void* stream);

/* ---- Sequence path with recurrent weights resident in shared memory ----------------------------------------------
 * Same math as ubs_agent_seq_*, split along the one true dependency of a BPTT window.  The caller computes, for all
 * n_steps*n_rows rows at once (batched GEMMs; the encoder does not depend on h, gnn_agents.py:53):
 *     x = relu(W_aggr xin + b_aggr),  pv = W_vsq[:, :H] x + b_vsq  (n_steps,n_rows,round4(M+2K)),
 *     pg = W_ih[:, :H] x + b_ih  (n_steps,n_rows,3H)         [row strides ld_pv / ld_pg: one fused GEMM output]
 * and afterwards the Q head, the x-path gradients and every parameter gradient.  The kernel keeps
 *     wt_vsq_h = W_vsq[:, H:]^T (H, round4(M+2K)), wt_ih_c = W_ih[:, H:]^T (M, 3H), wt_hh = W_hh^T (H, 3H), b_hh
 * in shared memory for the whole sequence (backward: w_hh (3H,H) and w_ih_c = W_ih[:, H:] (3H,M), original layouts).
 * Returns 3 when the weights do not fit (ubs_agent_seq2_smem_bytes > 227 KB): use ubs_agent_seq_* then.            */
UBS_API int64_t ubs_agent_seq2_smem_bytes(int H, int M, int K, int U, int flags, int backward);
UBS_API int ubs_agent_seq2_fwd(int H, int M, int K, int U, int flags, const float* wt_vsq_h, const float* wt_ih_c,
                       const float* wt_hh, const float* b_hh, const float* pv, const float* pg, const float* h0,
                       const uint32_t* mask, float* h_out, float* sv_vsq, float* sv_alpha, float* sv_c,
                       float* sv_gate, int64_t ld_pv, int64_t ld_pg, int64_t n_rows, int n_steps, void* stream);
/* dhq (n_steps,n_rows,H) = dq W_out (+ grad of the last hidden state on the last step).  Outputs the stashes
 * st_dgi / st_dgh (.., 3H), st_dvsq (.., round4(M+2K)) — three column blocks of row stride ld_stash, so they can
 * share one buffer [dgi | dvsq | dgh] that feeds the batched GEMMs without copies — and d_h0 (nullable).            */
UBS_API int ubs_agent_seq2_bwd(int H, int M, int K, int U, int flags, const float* w_hh, const float* w_ih_c,
                       const float* h0, const float* h_out, const float* sv_vsq, const float* sv_alpha,
                       const float* sv_gate, const float* dhq, float* st_dgi, float* st_dgh, float* st_dvsq,
                       float* d_h0, int64_t ld_stash, int64_t n_rows, int n_steps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UBS_GNN_H_ */
