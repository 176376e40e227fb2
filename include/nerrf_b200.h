/* nerrf_b200 -- C-ABI of the B200-native NERRF AI hot path.
 *
 * The reference (Itz-Agasta/nerrf @ a38ae13) has NO FFI / plugin interface for this path
 * (SURVEY.md 8b): the ai/ module it names (README.md:72-76) was never written.  The entry
 * points below are therefore the boundary a maintainer WOULD bind from the Python `ai/`
 * package the README names; each cites the reference interface (prose) it realises.
 *
 * Conventions:
 *   - plain pointers and sizes only; no torch / C++ types;
 *   - every entry point returns 0 on success, <0 on error; the message is available from
 *     nerrf_last_error() (thread-local); nothing throws across the ABI;
 *   - device-pointer entry points are asynchronous on `stream` (a cudaStream_t passed as
 *     void*; NULL = default stream), allocate nothing, and own nothing: the caller owns all
 *     inputs, outputs and workspaces;
 *   - *_host entry points take HOST pointers (pinned for async copies), do the H2D / D2H
 *     copies themselves and synchronise before returning; the only persistent allocations
 *     live inside an explicit opaque session handle (create / destroy).
 *   - all matrices row-major fp32; indices int32 (rowptr int32 or int64).
 */
#ifndef NERRF_B200_H
#define NERRF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NERRF_ABI_VERSION 2   /* 2: guard arrays (planner spec v1), MCTS host session, nerrf_trace_path_flags */

#define NERRF_OK 0
#define NERRF_ERR_INVALID (-1)   /* bad argument (shape, alignment, unsupported size)      */
#define NERRF_ERR_CUDA (-2)      /* a CUDA runtime call / kernel launch failed              */
#define NERRF_ERR_NODEVICE (-3)  /* no sm_100 device                                        */
#define NERRF_ERR_WORKSPACE (-4) /* workspace too small                                     */

/* algo selector for the GraphSAGE-T layer */
#define NERRF_SAGE_ALGO_AUTO 0
#define NERRF_SAGE_ALGO_FFMA 1   /* fused gather+aggregate -> fp32 CUDA-core GEMM           */
#define NERRF_SAGE_ALGO_UMMA 2   /* fused gather+aggregate -> tcgen05 GEMM, 3-term bf16 split
                                    (6 MMAs per K step; fp32-equivalent accuracy)           */
#define NERRF_SAGE_ALGO_UMMA2 3  /* same with a 2-term split (3 MMAs; ~1e-5 relative)       */

/* OR into `algo` of nerrf_sage_layer_fwd_ex: long_ws already holds the hub-row scan of THIS graph and row
 * range (it depends on rowptr only), e.g. layers 2..L of one forward: skip the re-scan. */
#define NERRF_SAGE_FLAG_REUSE_LONG_SCAN 0x100

typedef void* nerrf_stream_t;    /* cudaStream_t */

int nerrf_abi_version(void);
const char* nerrf_last_error(void);
/* sm count / compute capability of the current device; NERRF_ERR_NODEVICE if none. */
int nerrf_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------ GraphSAGE-T (rows a1-a3)
 * Replaces: ai/models/GraphSAGE-T.py `GraphSAGE_T.forward` (README.md:73, ROADMAP.md:127;
 * behaviour: docs/content/docs/architecture.mdx:49-53,157, threat-model.mdx:176-189).
 * Graph = CSR by destination: rowptr[n_nodes+1], col[E] source ids, ew[E] temporal weights. */

/* K1 alone: m[r - row_begin] = sum_e ew_e * x[col_e] / max(sum_e ew_e, 1e-12) for rows
 * [row_begin,row_end).  F in {32, 64, 128}.  m is [(row_end-row_begin), F]. */
int nerrf_sage_aggregate(const float* x, const void* rowptr, int rowptr_is64, const int32_t* col,
                         const float* ew, float* m, int64_t n_nodes, int64_t row_begin,
                         int64_t row_end, int F, nerrf_stream_t stream);

/* One fused layer: out[v] = act([x_v || m_v] @ W + b) for v in [row_begin,row_end); out is the
 * FULL [n_nodes, H] array (rows outside the range untouched).  W [2F, H], b [H]; H == 128. */
int nerrf_sage_layer_fwd(const float* x, const void* rowptr, int rowptr_is64, const int32_t* col,
                         const float* ew, const float* W, const float* b, float* out,
                         int64_t n_nodes, int64_t row_begin, int64_t row_end, int F, int H,
                         int relu, int algo, nerrf_stream_t stream);

/* The same layer with the node head fused into its epilogue (used for the last layer):
 * additionally score[v] = sigmoid(out_v . node_w + node_b) for v in [row_begin,row_end). */
int nerrf_sage_layer_head_fwd(const float* x, const void* rowptr, int rowptr_is64, const int32_t* col,
                              const float* ew, const float* W, const float* b, float* out,
                              int64_t n_nodes, int64_t row_begin, int64_t row_end, int F, int H,
                              int relu, int algo, const float* node_w, float node_b, float* score,
                              nerrf_stream_t stream);

/* Full-option layer call.  node_w == NULL: no fused head.  long_ws (optional device scratch, size from
 * nerrf_sage_long_rows_workspace_bytes): destination rows with more than 128 in-edges ("hub" rows: popular files, process nodes) are
 * pre-aggregated chunk-wise by many CTAs into it instead of being gathered by a single warp; rows that do
 * not fit in the scratch (or long_ws == NULL) are processed inline -- same result, slower. */
int nerrf_sage_long_rows_workspace_bytes(int64_t n_edges, size_t* bytes);
int nerrf_sage_layer_fwd_ex(const float* x, const void* rowptr, int rowptr_is64, const int32_t* col,
                            const float* ew, const float* W, const float* b, float* out,
                            int64_t n_nodes, int64_t row_begin, int64_t row_end, int F, int H,
                            int relu, int algo, const float* node_w, float node_b, float* score,
                            void* long_ws, size_t long_ws_bytes, float* const* peer_out, int n_peers,
                            const uint8_t* peer_need, nerrf_stream_t stream);
/* peer_out (host array of n_peers <= 7 device pointers, or NULL): peer-mapped [n_nodes, H] buffers of the other
 * ranks of a 1-D sharded forward (CUDA IPC / symmetric memory over NVLink).  The layer's epilogue stores each
 * output row to `out` AND to every peer buffer: the per-layer embedding exchange is fused into the layer
 * kernel.  peer_need (device uint8 [n_nodes], or NULL): bit i set <=> peer_out[i]'s rank references that row as
 * a source, so only the rows a peer actually reads are sent to it.  n_peers == -1: peer_out[0] is an NVSwitch MULTICAST address of the buffer on all ranks (this one
 * included): each element is stored once with multimem.st and replicated by the switch.
 * The caller provides the cross-rank barrier between layers. */

/* Heads.  score[v] = sigmoid(h_v . node_w + node_b).  If edge_W != NULL also writes
 * proj[v] = (h_v.We[0:H,0], h_v.We[0:H,1], h_v.We[H:2H,0], h_v.We[H:2H,1])  (proj [n,4]). */
int nerrf_sage_node_head(const float* h, const float* node_w, float node_b, float* score,
                         const float* edge_W, float* proj, int64_t row_begin, int64_t row_end,
                         int H, nerrf_stream_t stream);
/* edge_logit[e] = proj[col_e][0:2] + proj[dst_e][2:4] + edge_b   (== [h_src||h_dst] @ We + be) */
int nerrf_sage_edge_head(const float* proj, const void* rowptr, int rowptr_is64, const int32_t* col,
                         const float* edge_b, float* edge_logit, int64_t row_begin,
                         int64_t row_end, nerrf_stream_t stream);

/* Whole forward on device pointers: L layers + node head.  W[l] is [2F_l, H], b[l] is [H]
 * (host arrays of device pointers).  h_out [n_nodes,H]; score_out [n_nodes]; workspace must
 * hold n_nodes*H floats when L > 1; any bytes beyond that (256-aligned) are used as the hub-row scratch
 * (see nerrf_sage_long_rows_workspace_bytes). */
int nerrf_sage_forward(const float* x, const void* rowptr, int rowptr_is64, const int32_t* col,
                       const float* ew, int64_t n_nodes, int f_in, int hidden, int num_layers,
                       const float* const* W, const float* const* b, const float* node_w,
                       float node_b, float* h_out, float* score_out, float* workspace,
                       size_t workspace_bytes, int algo, nerrf_stream_t stream);

/* ------------------------------------------------------------------ GraphSAGE-T layer backward (training)
 * Replaces: the GraphSAGE-T half of ai/train.py "joint GNN+LSTM training script" (README.md:75; ROADMAP.md:62-69).
 * Given one layer's saved tensors -- input h_in [n,F], its aggregate m [n,F] (nerrf_sage_aggregate of h_in), output
 * y [n,H] -- and the gradient dy [n,H] of the loss w.r.t. y:
 *     dP = dy * [y > 0] (relu != 0) or dy;   db = sum_v dP[v];   dW = [h_in || m]^T dP  ([2F,H]);
 *     dZ = dP W^T;   dh = dZ[:, :F] + A^T dZ[:, F:]
 * A^T is a gather over the TRANSPOSED graph, CSR by source: t_rowptr [n+1], t_col [E] = destination of each out-edge,
 * t_w [E] = ew_e / max(weight sum of that destination, 1e-12).  dh == NULL skips the input gradient (first layer),
 * dW == NULL (with db) skips the weight gradient.  fp32, no atomics: bit-reproducible.  All pointers device,
 * 16-byte aligned; workspace from nerrf_sage_layer_bwd_workspace_bytes. */
int nerrf_sage_layer_bwd_workspace_bytes(int64_t n_nodes, int F, size_t* bytes);
int nerrf_sage_layer_bwd(const float* h_in, const float* m, const float* y, const float* dy, const float* W,
                         const void* t_rowptr, int t_rowptr_is64, const int32_t* t_col, const float* t_w,
                         float* dh, float* dW, float* db, float* workspace, size_t workspace_bytes,
                         int64_t n_nodes, int F, int H, int relu, nerrf_stream_t stream);

/* Host-buffer session (the e2e call): device buffers live in the handle. */
typedef struct nerrf_sage_session nerrf_sage_session;
int nerrf_sage_session_create(int64_t max_nodes, int64_t max_edges, int f_in, int hidden,
                              int num_layers, nerrf_sage_session** out);
/* weights are HOST pointers: W[l] [2F_l,H], b[l] [H], node_w [H]; copied to the device once. */
int nerrf_sage_session_set_weights(nerrf_sage_session* s, const float* const* W,
                                   const float* const* b, const float* node_w, float node_b);
/* HOST in / HOST out: copies the graph + features H2D, runs forward, copies score (and h if
 * h_out_host != NULL) D2H, synchronises. */
int nerrf_sage_session_forward_host(nerrf_sage_session* s, const float* x_host,
                                    const int32_t* rowptr_host, const int32_t* col_host,
                                    const float* ew_host, int64_t n_nodes, int64_t n_edges,
                                    float* score_out_host, float* h_out_host, int algo);
/* Pipelined form of the same call (a stream of graphs, e.g. one per sliding-window tick): submit queues H2D -> forward ->
 * D2H on the session's streams and returns at once with a ticket; wait blocks until that ticket's outputs are in the
 * host buffers.  Two steps may be in flight (two device buffer sets): the upload of step i+1 overlaps the layers of
 * step i.  A third submit blocks until the oldest ticket is complete.  The host buffers of a ticket must stay valid and
 * unmodified until its wait returns.  forward_host == submit + wait. */
int nerrf_sage_session_submit_host(nerrf_sage_session* s, const float* x_host,
                                   const int32_t* rowptr_host, const int32_t* col_host,
                                   const float* ew_host, int64_t n_nodes, int64_t n_edges,
                                   float* score_out_host, float* h_out_host, int algo,
                                   uint64_t* ticket);
int nerrf_sage_session_wait(nerrf_sage_session* s, uint64_t ticket);
int nerrf_sage_session_destroy(nerrf_sage_session* s);

/* ------------------------------------------------------------------ rewards.score (row a6)
 * Replaces: ai/planner/rewards.py (README.md:74,115 `Reward = -(data_loss + 0.1*downtime)`;
 * threat-model.mdx:205-223).  states uint32 [B, n_words], n_words = 32*NW, NW = 1/2/4 for
 * A <= 1024/2048/4096; action a = bit (a&31) of word (a>>5).  Bit-exact fixed-order fp32.
 * guard (device int32 [A], or NULL): planner spec v1 -- guard[a] = index (0..31) of the "kill process" action
 * (threat-model.mdx:212-215) that makes reverting a durable, or -1; an applied action whose guard is NOT applied
 * additionally loses p_guard * p_a * size_a (the live process re-encrypts the file).  NULL = spec v0. */
int nerrf_reward_score(const uint32_t* states, int64_t B, const float* p, const float* size,
                       const float* cost, const int32_t* guard, int A, float* out, nerrf_stream_t stream);

/* ------------------------------------------------------------------ planner.mcts.search (a5)
 * Replaces: ai/planner/mcts.py (README.md:74, ROADMAP.md:84; architecture.mdx:62-72).
 * Leaf-parallel UCT, one persistent cooperative kernel for all T iterations, ONE grid barrier per
 * iteration (every CTA keeps its own replica of the tree).
 * Workspace layout is private (it grows with the SM count); query its size first.  Outputs: root_n int32 [A_pad],
 * root_w fp32 [A_pad] (A_pad = 1024*NW), num_nodes int32 [1] -- all device pointers.
 * ln_table: device fp32 [T+2], ln_table[k] = fp32(ln(k*R)).  R must be a power of two. */
int nerrf_mcts_workspace_bytes(int A, int T, int R, size_t* bytes);
int nerrf_mcts_search(const float* p, const float* size, const float* cost, const int32_t* guard, int A,
                      const uint32_t* root_state, int R, int D, int T, uint64_t seed, float c,
                      float lo, float inv_range, const float* ln_table, int32_t* root_n,
                      float* root_w, int32_t* num_nodes, void* workspace, size_t workspace_bytes,
                      nerrf_stream_t stream);
/* HOST in / HOST out, one-shot (allocates + frees device memory internally, synchronous).
 * root_state_host and guard may be NULL. */
int nerrf_mcts_search_host(const float* p, const float* size, const float* cost, const int32_t* guard, int A,
                           const uint32_t* root_state_host, int R, int D, int T, uint64_t seed,
                           float c, float lo, float inv_range, const float* ln_table_host,
                           int32_t* root_n_host, float* root_w_host, int32_t* num_nodes_host);
/* plan(): validate-and-commit loop on the device (architecture.mdx:81-86 "sandbox validates, then apply"; the commit half of
 * ai/planner/mcts.py plan()).  cand[n_cand] (device int32, ranked root children of a search) are scored as score(state + c)
 * with the exact reward; the highest-ranked improving one is committed, the candidates that still improved stay in the list;
 * repeated up to max_commits times.  allow_tentative != 0: if nothing improves at the first round the first candidate is
 * committed anyway (spec-v1 lookahead).  state (device uint32 [32*NW]) is updated in place; actions_out / scores_out
 * [max_commits] and n_out[2] = {commits made, candidates left} are device arrays.  Asynchronous on `stream`; one cooperative
 * launch.  Decisions and scores are bit-identical to doing the loop with nerrf_reward_score on the host. */
int nerrf_plan_commit_workspace_bytes(int n_cand_max, size_t* bytes);
int nerrf_plan_commit(const float* p, const float* size, const float* cost, const int32_t* guard, int A,
                      uint32_t* state, const int32_t* cand, int n_cand, float cur, int max_commits,
                      int allow_tentative, int32_t* actions_out, float* scores_out, int32_t* n_out,
                      void* workspace, size_t workspace_bytes, nerrf_stream_t stream);

/* HOST in / HOST out through a session: the device buffers and a stream live in the handle (create once for the
 * largest A / T / R; a search is then six small H2D copies, one launch and three D2H copies, no allocation).
 * A must have the same word count NW as A_max. */
typedef struct nerrf_mcts_session nerrf_mcts_session;
int nerrf_mcts_session_create(int A_max, int T_max, int R_max, nerrf_mcts_session** out);
int nerrf_mcts_session_search_host(nerrf_mcts_session* s, const float* p, const float* size, const float* cost,
                                   const int32_t* guard, int A, const uint32_t* root_state_host, int R, int D, int T,
                                   uint64_t seed, float c, float lo, float inv_range, const float* ln_table_host,
                                   int32_t* root_n_host, float* root_w_host, int32_t* num_nodes_host);
int nerrf_mcts_session_destroy(nerrf_mcts_session* s);

/* ------------------------------------------------------------------ lstm.forward (row a4)
 * Replaces: ai/models/lstm.py (README.md:73; architecture.mdx:55-59; threat-model.mdx:191-203).
 * BiLSTM, torch.nn.LSTM gate order (i,f,g,o).  seq [B,T,D_in]; len int32 [B] (valid steps
 * t < len[b]).  Weights per layer l and direction d (index 2*l+d), PRE-TRANSPOSED by the host:
 *   Wih_t[2l+d] [D_l, 4H],  Whh_t[2l+d] [H, 4H],  bias[2l+d] [4H] (= b_ih + b_hh);
 * D_0 = D_in, D_l = 2H.  head_W [2, 2H], head_b [2].  out [B,2] = sigmoid(head).
 * workspace: 2 * B*T*2H floats (layer outputs ping-pong).  H must be 256; B any. */
int nerrf_lstm_workspace_bytes(int64_t B, int T, int H, size_t* bytes);
int nerrf_lstm_forward(const float* seq, const int32_t* len, int64_t B, int T, int D_in, int H,
                       int num_layers, const float* const* Wih_t, const float* const* Whh_t,
                       const float* const* bias, const float* head_W, const float* head_b,
                       float* out, void* workspace, size_t workspace_bytes,
                       nerrf_stream_t stream);

/* ------------------------------------------------------------------ temporal-graph constructor, device half
 * (SURVEY.md 8f rank 1 -- the step immediately before GraphSAGE_T.forward)
 * Replaces: the graph-constructor stage the reference documents but does not ship
 * (docs/content/docs/architecture.mdx:32-42 sliding window, inode dedup, "edge weight = causality
 * confidence"; node schema :144-160; input records proto/trace.proto:11-49).
 * Edge list (device, int32 src/dst in [0, n_nodes), fp32 event time t and confidence conf) ->
 * CSR-by-destination in exactly the layout nerrf_sage_forward reads:
 *   edges stably sorted by (dst, t);  col[i] = src;  ew[i] = conf * exp(-(t_ref - t) / tau);
 *   rowptr[v] = #edges with dst < v   (int32, or int64 when rowptr_is64; n_edges >= 2^31 needs int64).
 * rowptr and col are bit-exact against the host constructor (nerrf_b200/graph.py csr_from_edges);
 * ew differs only by the exp implementation (<= 2 ulp).  Deterministic (no atomics).
 * Limits: n_edges < 2^32, n_nodes < 2^31.  workspace: nerrf_graph_csr_workspace_bytes, 256-byte
 * aligned.  The call waits for `stream` once, to report out-of-range vertex ids as NERRF_ERR_INVALID. */
int nerrf_graph_csr_workspace_bytes(int64_t n_edges, int64_t n_nodes, int64_t* bytes);
int nerrf_graph_build_csr(const int32_t* src, const int32_t* dst, const float* t, const float* conf,
                          int64_t n_edges, int64_t n_nodes, float t_ref, float tau,
                          void* rowptr_out, int rowptr_is64, int32_t* col_out, float* ew_out,
                          void* workspace, int64_t workspace_bytes, nerrf_stream_t stream);
/* The same, and also perm_out[i] (device uint32 [n_edges], may be NULL) = index in the INPUT edge list of the edge at CSR
 * position i -- lets later stages (the per-file event sequences) go back from a CSR row to the events that made it. */
int nerrf_graph_build_csr_ex(const int32_t* src, const int32_t* dst, const float* t, const float* conf,
                             int64_t n_edges, int64_t n_nodes, float t_ref, float tau, void* rowptr_out,
                             int rowptr_is64, int32_t* col_out, float* ew_out, uint32_t* perm_out, void* workspace,
                             int64_t workspace_bytes, void* stream);
/* Per-file event sequences for lstm.forward on the device (architecture.mdx:55-59 "last 100 events per file"; spec =
 * nerrf_b200/ingest.py sequences_core): for every candidate file node its last t_max events, oldest first, as
 * seq_out [n_cand, t_max, 16] fp32 (zero padded) + len_out [n_cand].  Graph built from the window's events in time order
 * with TWO edges per event (process -> file, file -> process: merge_renames mode), perm from nerrf_graph_build_csr_ex, order
 * [n_window] = stored index of the window's events by rank; ts / event_slot / bytes / path_flags are the stored columns.
 * All pointers device. */
int nerrf_trace_sequences(const int64_t* cand_nodes, int n_cand, const void* rowptr, int rowptr_is64, const uint32_t* perm,
                          const int64_t* order, const double* ts, const uint8_t* event_slot, const int64_t* bytes,
                          const uint8_t* path_flags, double t0, double span, int t_max, float* seq_out,
                          int32_t* len_out, void* stream);

/* Per-node features on the device (rest of SURVEY.md 8f rank 1): event columns already mapped to node
 * ids (nerrf_trace_intern, events in time order) -> x [n_nodes, 32] in the layout GraphSAGE_T.forward
 * reads (node schema docs/content/docs/architecture.mdx:144-160, threat-model.mdx:154-184; column
 * meaning nerrf_b200/graph.py graph_from_events): 0 file, 1 process, 3/4 log1p(in/out degree),
 * 5..12 log1p(event-kind counts), 13 log1p(bytes)/20, 14 (last-first)/window, 15 first/window,
 * 16 .lockbit, 17 ransom-note, 18 /tmp|/proc, 19 completes/(starts+writes); label_out (0/1: the file
 * was encrypted) and size_mb_out are optional.  Device pointers; t = seconds since the first event
 * (>= 0, double); node_g may be NULL (no rename targets).  Integer atomics only -> deterministic;
 * equals the host constructor up to the rounding of log1p (<= 1 ulp of fp32).  Waits for `stream`
 * once to report malformed columns. */
int nerrf_graph_node_features_workspace_bytes(int64_t n_nodes, int64_t* bytes);
int nerrf_graph_node_features(const int32_t* node_p, const int32_t* node_f, const int32_t* node_g,
                              const double* t, const uint8_t* event_slot, const uint64_t* bytes,
                              const uint8_t* path_flags, int64_t n_events, const int8_t* node_kind,
                              int64_t n_nodes, double window, float* x_out, int32_t* label_out,
                              float* size_mb_out, void* workspace, int64_t workspace_bytes,
                              nerrf_stream_t stream);

/* ------------------------------------------------------------------ EventBatch ingest (HOST code)
 * (SURVEY.md 8f rank 2 -- the wire format on the input side of the path)
 * Replaces: the consumer of the tracker's gRPC stream (proto/trace.proto:11-49 Event, :47-49
 * EventBatch; producer tracker/cmd/tracker/main.go:229-252).  `buf` holds one serialized
 * nerrf.trace.EventBatch, or any number of them concatenated (repeated fields append).
 * nerrf_trace_scan counts the events (top-level hop) and gives an upper bound per string column
 * (string_bytes[4] = comm, syscall, path, new_path; exact totals are off[n_events] after decoding);
 * nerrf_trace_decode parses the events once and fills caller-owned arrays: scalars [n_events]; strings as offsets [n_events+1]
 * + packed bytes; event_slot = graph feature slot of the syscall name (0 file_created,
 * 1 file_encrypt_start, 2 file_encrypt_complete, 3 ransom_note_created, 4 openat, 5 write,
 * 6 rename, 7 other); path_flags = NERRF_PATH_* bits of `path`.  Malformed input ->
 * NERRF_ERR_INVALID with the byte offset in nerrf_last_error().
 * Events must name a file: the tracker zeroes `path` on write events (tracker/bpf/tracepoints.c:62-64); the host
 * side (nerrf_b200/ingest.py resolve_columns) attributes those to the pid's last non-empty path before interning.
 * nerrf_trace_intern: pid / path interning ("merge by inode", architecture.mdx:41): node ids in
 * order of first appearance over events visited in `order` (NULL = as stored); node tables sized
 * by the caller (node_capacity >= 2*n_events with merge_renames, 3*n_events without); with merge_renames a
 * non-empty new_path aliases its merge key to the event's file node (a real rename a.dat -> a.dat.lockbit3);
 * node_name_which: 0 = path, 1 = new_path of event node_name_event, 2 = the pid itself. */
#define NERRF_PATH_LOCKBIT   1   /* ".lockbit" in path */
#define NERRF_PATH_NOTE      2   /* README / RANSOM (ASCII case-insensitive) in path */
#define NERRF_PATH_TMP       4   /* path starts with /tmp or /proc */
#define NERRF_PATH_ENCRYPTED 8   /* path ends with ".lockbit3" (benchmarks/m1/scripts/m1_rollback.sh:95) */
int nerrf_trace_scan(const uint8_t* buf, int64_t len, int64_t* n_events, int64_t* string_bytes);
int nerrf_trace_decode(const uint8_t* buf, int64_t len, int64_t n_events,
                       int64_t* ts_sec, int32_t* ts_nanos, uint32_t* pid, uint32_t* tid,
                       int32_t* flags, int64_t* ret_val, uint64_t* bytes,
                       uint8_t* event_slot, uint8_t* path_flags,
                       int64_t* comm_off, uint8_t* comm_data, int64_t* syscall_off, uint8_t* syscall_data,
                       int64_t* path_off, uint8_t* path_data, int64_t* new_path_off, uint8_t* new_path_data);
/* NERRF_PATH_* bits for every string of a packed string column (off [n+1], data): the rule nerrf_trace_decode
 * applies to `path`, available for `new_path` (a rename target names the file from then on). */
int nerrf_trace_path_flags(const int64_t* off, const uint8_t* data, int64_t n, uint8_t* flags_out);

/* The same interning ON THE DEVICE ("inode/path dedup via hash"; architecture.mdx:39-41): identical
 * nodes, numbering and names for the same events in the same processing order, computed by data-parallel passes (64-bit
 * hash + open-addressing table with first-mention atomicMin, rename-alias forest, creating-mention scan).  All array
 * arguments are DEVICE pointers; n_nodes is a HOST pointer (the call synchronises the stream).  The columns hold n_stored
 * events; the n_events events order[0..n_events) of them are processed in that order (order == NULL: the first n_events in
 * stored order) -- a sliding window over a resident stream is a slice of its time-sorted index array.  node_p / node_f /
 * node_g are written BY STORED INDEX (arrays of n_stored entries, only the processed events' entries are touched).  A 64-bit hash collision between two different keys is detected (keys are compared with their first mention's
 * bytes) and reported as an error, never merged silently.  workspace: 256-byte aligned, size from
 * nerrf_trace_intern_device_workspace_bytes. */
int nerrf_trace_intern_device_workspace_bytes(int64_t n_events, int64_t node_capacity, int64_t* bytes);
int nerrf_trace_intern_device(int64_t n_events, int64_t n_stored, const int64_t* order, const uint32_t* pid, const int64_t* path_off,
                              const uint8_t* path_data, const int64_t* new_path_off, const uint8_t* new_path_data,
                              int merge_renames, int32_t* node_p, int32_t* node_f, int32_t* node_g, int64_t* n_nodes,
                              int8_t* node_kind, int64_t* node_name_event, int8_t* node_name_which,
                              int64_t node_capacity, void* workspace, int64_t workspace_bytes, void* stream);
/* hash_out[v] = 64-bit hash of node v's NAME (the path / new_path of its naming event; pid nodes: the pid), device arrays:
 * lets a caller match nodes against a set of names without moving the strings (nerrf_b200/ingest.py name_hash is the same
 * function on the host). */
int nerrf_trace_name_hash(int64_t n_nodes, const int64_t* node_name_event, const int8_t* node_name_which,
                          const uint32_t* pid, const int64_t* path_off, const uint8_t* path_data,
                          const int64_t* new_path_off, const uint8_t* new_path_data, uint64_t* hash_out, void* stream);
int nerrf_trace_intern(int64_t n_events, const int64_t* order, const uint32_t* pid,
                       const int64_t* path_off, const uint8_t* path_data,
                       const int64_t* new_path_off, const uint8_t* new_path_data, int merge_renames,
                       int32_t* node_p, int32_t* node_f, int32_t* node_g, int64_t* n_nodes,
                       int8_t* node_kind, int64_t* node_name_event, int8_t* node_name_which,
                       int64_t node_capacity);

#ifdef __cplusplus
}
#endif
#endif /* NERRF_B200_H */
