/*
 * linkteller_b200 — C ABI of the B200-native (sm_100a) LinkTeller influence-analysis hot path.
 *
 * The reference (AI-secure/LinkTeller) is pure Python and has no FFI layer of its own; its
 * boundary for this path is a set of Python call sites.  Every entry point below
 * names the reference interface it replaces (file:line relative to the upstream tree).  The
 * reference-side binding a maintainer would add is a ctypes stub — see INTEGRATION.md and
 * linkteller_b200/_lib.py.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in
 *    `_host`; the caller (PyTorch) owns all input and output buffers;
 *  - every call takes the caller's CUDA stream (as void*) and is asynchronous on it, except the
 *    calls documented to return a count through a `_host` pointer, which synchronise the stream;
 *  - return value: LT_OK (0) or a negative error code; lt_last_error() gives the thread-local
 *    message.  No C++ exception crosses the boundary;
 *  - scratch memory is stream-ordered (cudaMallocAsync) and released before the call returns;
 *    the only persistent allocation is the opaque lt_plan;
 *  - CSR matrices are (indptr int32[n+1], indices int32[nnz], values), rows sorted by column.
 */
#ifndef LINKTELLER_B200_H
#define LINKTELLER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LT_OK 0
#define LT_ERR_ARG (-1)
#define LT_ERR_CUDA (-2)
#define LT_ERR_NOMEM (-3)
#define LT_ERR_INTERNAL (-4)
#define LT_ERR_DATA (-5) /* input violates a documented precondition (unsorted COO, asymmetric Â …) */
#define LT_ERR_CAPACITY (-6)

/* adjacency normalisations in scope — fetch_normalization, utils/load.py:617-627 */
#define LT_NORM_FIRST_ORDER_GCN 0 /* I + D^-1/2 A D^-1/2        utils/load.py:572-578 */
#define LT_NORM_AUG_NORM_ADJ 1    /* D~^-1/2 (A+I) D~^-1/2      utils/load.py:562-569 */

/* value precision of the normalisation arithmetic */
#define LT_PREC_F32 0
#define LT_PREC_F64 1

typedef void* lt_stream_t; /* cudaStream_t */
typedef struct lt_plan lt_plan;
typedef struct lt_comm lt_comm;

const char* lt_last_error(void);
int lt_version(void);
/* sm count, compute capability and the device's memory size of the current device */
int lt_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host, int64_t* mem_bytes_host);

/* ---------------------------------------------------------------------------------------------
 * a1  graph_reader, utils/load.py:452-460:
 *       adj = csr_matrix((ones, (edges[:,0], edges[:,1])), shape=(n,n), dtype=float32); adj + adj.T
 *     edges: int64 [m,2].  Duplicates are SUMMED and self loops doubled, exactly like scipy; with
 *     binarize != 0 every stored value is 1 instead.  indices/values have capacity `cap` (2m is
 *     always enough).  Synchronises; the number of stored entries is returned in *nnz_host.
 * ------------------------------------------------------------------------------------------- */
int lt_csr_from_edges(const int64_t* edges, int64_t m, int32_t n, int binarize, int32_t* indptr,
                      int32_t* indices, float* values, int64_t cap, int64_t* nnz_host, lt_stream_t stream);

/* a3 inverse — the torch sparse COO tensor `adj` that GraphConvolution.forward receives
 * (utils/load.py:552-559; gcn/layers.py:30) → CSR.  row/col: int64 [nnz], row-major sorted and
 * duplicate-free (what sparse_mx_to_torch_sparse_tensor produces).  Returns LT_ERR_DATA through
 * lt_csr_check() if that does not hold. */
int lt_coo_to_csr(const int64_t* row, const int64_t* col, int64_t nnz, int32_t n, int32_t* indptr,
                  int32_t* indices, lt_stream_t stream);
/* 1 in *ok_host iff rows are sorted, columns strictly increasing within a row and in [0,n). Syncs. */
int lt_csr_check(const int32_t* indptr, const int32_t* indices, int32_t n, int64_t nnz, int* ok_host,
                 lt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a2  fetch_normalization(...)(adj), utils/load.py:562-578, followed by the float32 cast of
 *     sparse_mx_to_torch_sparse_tensor (utils/load.py:554).
 *     Step 1: row sums (degree vector) in the working precision, accumulated in column order.
 *     Step 2 (host): d^-1/2 with numpy exactly as the reference does (np.power, inf → 0).
 *     Step 3: scale, add the identity, drop zero products, emit Â with float32 values.
 *     `values`, `row_sums` and `d_inv_sqrt` are float32 for LT_PREC_F32 and float64 for
 *     LT_PREC_F64.  Output capacity nnz+n is always enough.  lt_adj_normalize synchronises.
 * ------------------------------------------------------------------------------------------- */
int lt_csr_row_sums(const int32_t* indptr, const void* values, int32_t n, int prec, void* row_sums,
                    lt_stream_t stream);
int lt_adj_normalize(const int32_t* indptr, const int32_t* indices, const void* values, int32_t n,
                     int kind, int prec, const void* d_inv_sqrt, int32_t* out_indptr,
                     int32_t* out_indices, float* out_values, int64_t cap, int64_t* nnz_host,
                     lt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a4  GraphConvolution.forward, gcn/layers.py:30-36:  support = mm(input, W); spmm(adj, support) + b
 *     lt_gemm_f32 : C[M,N] = A[M,K]·B (+bias[N]) (ReLU).  Row-major.  a_rows (nullable) gathers
 *                   A's rows: row m of the product uses A[a_rows[m], :].  transB != 0 reads B as
 *                   [N,K] (C = A·Bᵀ).  algo: 0 = auto, 1 = fp32 SIMT, 2 = 3xTF32 tensor cores.
 *     lt_spmm_csr_f32 : C[i,:] = Σ_k Â[r_i,k]·B[k,:] (+bias[d]) (ReLU), r_i = rows ? rows[i] : i;
 *                   Â is n x n, B has n rows.
 * ------------------------------------------------------------------------------------------- */
int lt_gemm_f32(const float* A, int64_t lda, const int32_t* a_rows, const float* B, int64_t ldb,
                int transB, const float* bias, int relu, float* C, int64_t ldc, int32_t M, int32_t N,
                int32_t K, int algo, lt_stream_t stream);
int lt_spmm_csr_f32(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n,
                    const int32_t* rows, int32_t n_rows, const float* B, int64_t ldb, int32_t d,
                    const float* bias, int relu, float* C, int64_t ldc, lt_stream_t stream);

/* a5  GCN.forward, gcn/models.py:19-24 (eval mode: dropout is the identity): raw logits [n,c]. */
int lt_gcn_forward(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n,
                   const float* X, int64_t ldx, int32_t f, const float* W1, const float* b1,
                   const float* W2, const float* b2, int32_t h, int32_t c, float* logits,
                   int64_t ldz, lt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a7/a8  Attacker.get_gradient_eps_mat (attacker.py:100-108) inside the loop of
 *        link_prediction_attack_efficient (attacker.py:209-229).
 *
 *  lt_plan_create builds the symbolic structure of the attack for the perturbed nodes
 *  test_nodes[lo:hi) observed at all T test nodes: the frontier rows R = ∪ N̂(v), the incidence
 *  list S (v, r, Â[r,v]), the observer lists B (r → test nodes adjacent to r) and the list of
 *  (perturbed, observer) pairs in the support with their contributions.  Â must have a symmetric
 *  pattern (both normalisations in scope do); otherwise LT_ERR_DATA.  Synchronises twice.
 *
 *  lt_influence_rows evaluates, for every v = test_nodes[lo+j]:
 *     out[j, i] = || ((f(X + δ·e_v x_vᵀ) − f(X)) / δ)[test_nodes[i]] ||₂          (0 outside 2 hops)
 *  from H_R = (Â·X·W1 + b1)[R] (row k ↔ lt_plan_rows()[k]) and P_v = X[v]·W1 given as
 *  P[p_index ? p_index[j] : j, :].  out is [hi-lo, T] float32 with leading dimension ldo and is
 *  fully overwritten (zeros included).
 *
 *  lt_attack_efficient = plan + restricted precompute of H_R / P + lt_influence_rows in one call.  When the
 *  products are large enough for the tensor cores it keeps only the signs of H_R (one bit per element and the
 *  row's smallest magnitude) and recomputes the rows where a ReLU can switch exactly; the precompute is enqueued
 *  on the caller's stream as soon as the frontier rows are known, while the rest of the plan runs on an
 *  internal high-priority stream that is joined back before the frontier kernels (to the caller all work
 *  stays ordered on the stream passed in).
 * ------------------------------------------------------------------------------------------- */
int lt_plan_create(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n,
                   const int32_t* test_nodes, int32_t T, int32_t lo, int32_t hi, lt_stream_t stream,
                   lt_plan** plan_host);
int lt_plan_destroy(lt_plan* plan, lt_stream_t stream);
/* info_host[0..7]: |R|, incidences E, observer entries, records, pairs in the support,
 * algorithmic bytes of one lt_influence_rows (for h,c below), Σ_{r∈R} nnz(Â[r,:]),
 * incidences whose ReLUs could switch and took the exact recomputed path (lt_attack_efficient only, else 0) */
int lt_plan_info(const lt_plan* plan, int32_t h, int32_t c, int64_t* info_host, lt_stream_t stream);
const int32_t* lt_plan_rows(const lt_plan* plan);
/* copies the |R| frontier row ids (ascending node ids) into dst (capacity info[0]) */
int lt_plan_copy_rows(const lt_plan* plan, int32_t* dst, lt_stream_t stream);
int lt_influence_rows(const lt_plan* plan, const float* H_R, int64_t ldh, const float* P, int64_t ldp,
                      const int32_t* p_index, const float* W2, int32_t h, int32_t c, float delta,
                      float* out, int64_t ldo, lt_stream_t stream);
int lt_attack_efficient(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n,
                        const float* X, int64_t ldx, int32_t f, const float* W1, const float* b1,
                        const float* W2, int32_t h, int32_t c, const int32_t* test_nodes, int32_t T,
                        int32_t lo, int32_t hi, float delta, float* out, int64_t ldo,
                        int64_t* info_host /* nullable, 8 entries as lt_plan_info */, lt_stream_t stream);

/* f3  --n-layer 3: GCN3.forward (gcn/models.py:40-46, eval mode) and the efficient attack on it
 *     (gcn_trainer.py:81-86 builds GCN3(nfeat, hidden1, hidden2, nclass); Attacker calls self.model
 *     generically, attacker.py:100-108).  W1 [f,h1], W2 [h1,h2], W3 [h2,c] row-major.
 *     lt_attack_efficient3: out[j,i] as lt_attack_efficient, exactly 0 outside THREE hops.  The perturbed
 *     nodes are processed in chunks whose second-hop expansions (Σ_{r∈N̂(v)} nnz(Â[r,:]) records) stay below
 *     max_chunk_records (<= 0: sized from the free device memory).  info_host (nullable, 8 entries):
 *     chunks, first-hop incidences, second-hop records, second-hop incidences, third-hop records,
 *     pairs in the support, 0, 0.  Synchronises several times per chunk. */
int lt_gcn3_forward(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n,
                    const float* X, int64_t ldx, int32_t f, const float* W1, const float* b1,
                    const float* W2, const float* b2, const float* W3, const float* b3, int32_t h1,
                    int32_t h2, int32_t c, float* logits, int64_t ldz, lt_stream_t stream);
int lt_attack_efficient3(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n,
                         const float* X, int64_t ldx, int32_t f, const float* W1, const float* b1,
                         const float* W2, const float* b2, const float* W3, int32_t h1, int32_t h2,
                         int32_t c, const int32_t* test_nodes, int32_t T, int32_t lo, int32_t hi,
                         float delta, float* out, int64_t ldo, int64_t max_chunk_records,
                         int64_t* info_host, lt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a6  _get_edge_sets_among_nodes, utils/load.py:304-326: label of every pair i<j in sample order
 *     (row-major) = test_nodes[j] ∈ row test_nodes[i] of adj_ori.  labels: uint8 [T(T-1)/2].
 * a8  pair score (attacker.py:233-245): score(i<j) = influence[j][i] (perturb the later node of
 *     the pair, observe the earlier one).  scores: float32 [T(T-1)/2], same order.
 * a9  compute_and_save (attacker.py:378-386): ROC-AUC with one threshold per distinct score
 *     (tie-aware Mann-Whitney) and sklearn's average precision; attack_stats_all.py:108-116:
 *     true positives among the global top-n_pos scores (tp_min/tp_max over tie-breaks).
 * ------------------------------------------------------------------------------------------- */
int lt_pair_labels(const int32_t* indptr, const int32_t* indices, int32_t n, const int32_t* test_nodes,
                   int32_t T, uint8_t* labels, int64_t* n_pos_host, lt_stream_t stream);
int lt_pair_scores(const float* influence, int64_t ldi, int32_t T, float* scores, lt_stream_t stream);
int lt_auc_ap(const float* scores, const uint8_t* labels, int64_t n_pairs, double* auc_host,
              double* ap_host, lt_stream_t stream);
int lt_topk_tp(const float* scores, const uint8_t* labels, int64_t n_pairs, const int64_t* n_pos_host,
               int32_t n_queries, int64_t* tp_min_host, int64_t* tp_max_host, lt_stream_t stream);

/* a9  the curves compute_and_save stores (attacker.py:382-385 → sklearn _binary_clf_curve): one point
 *     per distinct score in descending order — threshold, cumulative true and false positives.
 *     cap == 0 queries the number of points (returned in *n_points_host).  Synchronises. */
int lt_rank_curve(const float* scores, const uint8_t* labels, int64_t n_pairs, float* thresholds, int64_t* tps,
                  int64_t* fps, int64_t cap, int64_t* n_points_host, lt_stream_t stream);

/* a10 baseline_attack (LSA2-post / LSA2-attr), attacker.py:287-334: centred cosine similarity of
 *     the rows M[test_nodes[i], :d] after transform (0 none, 1 softmax, 2 sigmoid); dist [T,T]
 *     float32, upper triangle (i<j) filled, rest 0. */
int lt_lsa2_corr(const float* M, int64_t ldm, int32_t d, const int32_t* test_nodes, int32_t T,
                 int transform, float* dist, int64_t ldd, lt_stream_t stream);

/* f2  baseline_attack_balanced (attacker.py:337-375, --sample-type balanced-full): the same centred
 *     cosine similarity for LISTED pairs (pu[k], pv[k]); the mean runs over ALL n_rows rows of M
 *     (attacker.py:361), out[k] in list order. */
int lt_lsa2_pairs(const float* M, int64_t ldm, int32_t n_rows, int32_t d, int transform, const int32_t* pu,
                  const int32_t* pv, int64_t n_pairs, float* out, lt_stream_t stream);

/* f2/f4  link_prediction_attack_efficient_balanced (attacker.py:250-284) and the naive per-pair attack
 *     (attacker.py:143-160 → get_gradient_eps :87-97) read single entries grad_mat(perturbed)[observed]:
 *     out[k] = block[(rows[k] − row0)·ld + cols[k]] for the pairs whose perturbed node lies in the
 *     block of influence rows [row0, row0 + n_rows) that lt_attack_efficient just produced. */
int lt_pair_gather(const float* block, int64_t ld, int32_t row0, int32_t n_rows, const int32_t* rows,
                   const int32_t* cols, int64_t n_pairs, float* out, lt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a11 perturb_adj_discrete (EdgeRand), worker.py:213-278: the Bernoulli hits and their add/remove
 *     coin are drawn on the HOST from numpy's legacy stream (same seed ⇒ same flips); the device
 *     applies them: adj + add − minus with −1→0 and 2→1, explicit zeros kept (worker.py:264-267).
 *     hits: int64 [n_hits] cell ids r*n+c with r<c, ascending; coin: uint8 [n_hits] (1 = add).
 *     Output values float64, capacity nnz+2·n_hits is always enough.  Synchronises.
 * a12 perturb_adj_continuous (Lapgraph), worker.py:281-335: noise is the host-drawn Laplace
 *     matrix (float64 [n,n], row-major); the device adds tril(A,-1) into it (worker.py:298), selects
 *     the exact top-n_keep cells of the strict lower triangle (ties at the cut: lowest row-major
 *     cell first) and rebuilds the symmetric binary CSR (int32 ones upstream, float32 here).
 *     Capacity 2·n_keep.  Synchronises.
 * ------------------------------------------------------------------------------------------- */
int lt_dp_edgerand_apply(const int32_t* indptr, const int32_t* indices, const double* values, int32_t n,
                         const int64_t* hits, const uint8_t* coin, int64_t n_hits, int32_t* out_indptr,
                         int32_t* out_indices, double* out_values, int64_t cap, int64_t* nnz_host,
                         lt_stream_t stream);
int lt_dp_lapgraph_select(const int32_t* indptr, const int32_t* indices, const double* values, int32_t n,
                          double* noise /* consumed: tril(A,-1) is added in place */, int64_t n_keep, int32_t* out_indptr, int32_t* out_indices,
                          float* out_values, int64_t cap, int64_t* nnz_host, lt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Phase timers (measurement only).  When enabled, every kernel family launched by
 * the calls above is bracketed by CUDA events on the caller's stream.  lt_prof_report synchronises
 * the device and writes "name count total_ms\n" lines (aggregated by name, launch order) into buf;
 * returns LT_ERR_CAPACITY if buf is too small.  Disabled by default: no events are recorded.
 * ------------------------------------------------------------------------------------------- */
int64_t lt_launch_count(void); /* kernels launched by this library in this process so far */
int lt_prof_enable(int on);
int lt_prof_reset(void);
int lt_prof_report(char* buf_host, int64_t cap);

/* ---------------------------------------------------------------------------------------------
 * f2 / f4  link_prediction_attack_efficient_balanced (attacker.py:250-284, --sample-type balanced-full) and the naive per-pair
 *     attack (attacker.py:143-160 → get_gradient_eps, :87-97): one influence entry per LISTED (perturbed, observed) pair.
 *     test_nodes: the T distinct perturbed nodes; pair k perturbs test_nodes[pair_j[k]] and observes node pair_obs[k] (any
 *     node of the graph).  out[k] = ‖((f(X + δ·e_p x_pᵀ) − f(X))/δ)[o]‖₂.  Same precompute / F1 / float64 fix-up as
 *     lt_attack_efficient; the reduce intersects N̂(p) ∩ N̂(o) per pair — no row block over all observers is formed.
 * ------------------------------------------------------------------------------------------- */
int lt_pair_influence(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n, const float* X,
                      int64_t ldx, int32_t f, const float* W1, const float* b1, const float* W2, int32_t h, int32_t c,
                      const int32_t* test_nodes, int32_t T, const int32_t* pair_j, const int32_t* pair_obs, int64_t K,
                      float delta, float* out, lt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Several GPUs of one node (the reference has no multi-device code to replace: grep finds no
 * torch.distributed / NCCL / MPI call site).  One process per GPU.  lt_comm_create allocates this rank's symmetric buffer
 * (`bytes` of payload: the full T x T float32 influence matrix) and returns its 64-byte CUDA IPC handle through
 * lt_comm_get_handle; the caller exchanges the handles of all ranks (any host-side channel) and hands the world x 64 bytes,
 * in rank order, to lt_comm_open, which maps every peer buffer into this process.  lt_attack_efficient_comm is
 * lt_attack_efficient for rank `rank`: it perturbs test_nodes[lo:hi) and its reduce kernel stores the influence entries it
 * produces into EVERY rank's matrix (st.global on the mapped peer pointers, carried by NVLink / NVSwitch), ordered by
 * epoch flags with system-scope fences — there is no collective call and no dense all-gather.  On return (stream order)
 * lt_comm_buffer(comm) holds all T rows.  Every rank must make the call for every attack, also with hi == lo.  Waits are
 * bounded (≈ 2 s); lt_comm_timed_out reports (and clears) a wait that gave up.  Synchronising calls: create, open,
 * timed_out, destroy.
 * aux_bytes > 0 adds a second symmetric region: with n·(ceil(h/32) + 2)·4 bytes of it lt_attack_efficient_comm also SHARDS
 * THE RESTRICTED PRECOMPUTE — every rank marks the frontier rows of ALL T test nodes (the same set on every rank), takes
 * the rank-th 1/world of them through the SpMM and the sign GEMM, stores its 44-byte sign records (sign mask, min and max
 * |H|) into every rank's aux region and a third flag exchange completes them — instead of every rank re-computing the hub
 * rows its slice shares with all the others.  (Option "shard_rows" = 0, or an aux region that is too small, keeps the
 * per-slice precompute.)
 * ------------------------------------------------------------------------------------------- */
int lt_comm_create(int rank, int world, int64_t bytes, int64_t aux_bytes, lt_comm** comm_host);
int lt_comm_get_handle(const lt_comm* comm, void* handle64_host);
int lt_comm_open(lt_comm* comm, const void* all_handles_host);
void* lt_comm_buffer(const lt_comm* comm);
int lt_comm_timed_out(lt_comm* comm, int* flag_host);
int lt_comm_destroy(lt_comm* comm);
int lt_attack_efficient_comm(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n, const float* X,
                             int64_t ldx, int32_t f, const float* W1, const float* b1, const float* W2, int32_t h, int32_t c,
                             const int32_t* test_nodes, int32_t T, int32_t lo, int32_t hi, float delta, lt_comm* comm,
                             int64_t* info_host /* nullable */, lt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Process-wide tuning / debugging switches.  The reference has no counterpart (its only knobs are main.py's argparse
 * flags, main.py:17-93); these select between kernels that produce the same results and exist for tests and profiling.
 * Initial values are read from the environment ONCE (LT_NO_OVERLAP, LT_NO_SIGN_PATH, LT_TC_WS, LT_SPMM_PANEL[_MB|_LMIN],
 * LT_SPMM_HUB[_MB]); no compute call reads the environment.  Names: "overlap" (plan tail / big-group reduce on the
 * library's side stream, default 1), "sign_path" (sign-mask form of H_R, 1), "tc_ws" (warp-specialised tcgen05 GEMM, 1),
 * "spmm_panel" (column-panel SpMM, 0), "spmm_panel_mb" (48), "spmm_panel_lmin" (64), "spmm_hub" (hub-column L2 pinning
 * in the restricted SpMM, 0), "spmm_hub_mb" (size of the pinned hub panel in MB, 64), "fix_eps" (the float64 fix-up of
 * lt_attack_efficient treats |H| ≤ 2^-fix_eps·max(f,64)/64·max|H[r,:]| as undecidable in float32 and re-evaluates it; 19),
 * "spmm_stage" (cp.async shared-memory staging of the gathered rows, 0), "spmm_segmeta" (work-item table instead of a
 * binary search in the restricted SpMM, 1), "shard_rows" (several ranks: restricted precompute sharded by global frontier
 * rows, 1), "spmm_hint" (L2 eviction-priority hints for the gathers of the restricted SpMM, chosen from the entry values —
 * 1/sqrt(d_r·d_c) says how hot column c is; 0 off, 2 = hot evict_last / cold evict_first, default 2; results do not
 * depend on it), "spmm_hint_deg" (column degree from which a row counts as hot; 0 = measured from the degree histogram
 * so that the hot rows fill spmm_hub_mb, default 0).
 * ------------------------------------------------------------------------------------------- */
int lt_set_option(const char* name_host, int value);
int lt_get_option(const char* name_host, int* value_host);

/* primitives exported for the unit tests */
int lt_exclusive_scan_i32(const int32_t* in, int32_t* out /* n+1 */, int64_t n, lt_stream_t stream);
int lt_radix_sort_pairs_u64(uint64_t* keys, uint64_t* vals /* nullable */, int64_t n, int begin_bit,
                            int end_bit, lt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LINKTELLER_B200_H */
