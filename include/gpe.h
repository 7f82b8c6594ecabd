/* gpe.h -- C ABI of libgpe.so, the B200-native replacement for GNN-PE's three hot paths.
 *
 * The reference (JamesWhiteSnow/GNN-PE) has no plugin or FFI interface: include/custom.h is
 * compiled into src/main.cpp's single translation unit.  The seams a maintainer would cut are
 * three call sites in src/main.cpp; each group of entry points below names the one it replaces
 * (file:line relative to the reference's GNN-PE/ directory).  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain pointers and sizes only; vertex / label ids are uint32_t (reference `ui`,
 *     include/configuration/types.h:13-17), embeddings are double, row counts are uint64_t
 *     (the reference's 32-bit `ui path_num`, custom.h:548, overflows at BASELINE.json's sizes);
 *   - every call returns 0 on success, non-zero on error; gpe_last_error() gives the message;
 *     the library never calls exit() (the reference's R-tree library does, functions.cpp:24-29);
 *   - host pointers passed in are caller-owned and not retained after the call returns;
 *   - a context owns one GPU and one CUDA stream and is not thread-safe: where the reference
 *     runs one OpenMP thread per partition (main.cpp:160-164) the caller makes ONE call;
 *   - there is no CPU fallback: without a usable sm_100 device gpe_create() fails.
 */
#ifndef GPE_H_
#define GPE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpe_ctx gpe_ctx;

#define GPE_OK 0
#define GPE_ERR_INVALID 1     /* bad argument / call order */
#define GPE_ERR_CUDA 2        /* CUDA runtime error (message has the detail) */
#define GPE_ERR_UNSUPPORTED 3 /* (L, e) or query size outside the compiled kernels */
#define GPE_ERR_IO 4

#define GPE_LIMIT_MAX 0xFFFFFFFFull /* `-n MAX` = UINT_MAX, main.cpp:62-65 */
#define GPE_MAX_QUERY_VERTICES 64   /* MAXIMUM_QUERY_GRAPH_SIZE, include/configuration/config.h:3 */

/* ---- lifecycle ------------------------------------------------------------------------ */
int gpe_create(int device, gpe_ctx **out);
void gpe_destroy(gpe_ctx *ctx);
const char *gpe_last_error(const gpe_ctx *ctx); /* ctx may be NULL: error of the last failed gpe_create */
int gpe_abi_version(void);

/* ---- host-side mirror of the reference's cheap serial steps (pure host code, no GPU) --- */

/* Static_Graph::loadGraphFromFile, libsrc/graph/graph.cpp:163-242.  Two-call pattern: with
 * offsets == NULL only V, E are returned.  nbrs come back sorted ascending per vertex (:231-233).  A file whose header,
 * declared degrees and edge list disagree is an error here (the reference reads or writes out of bounds). */
int gpe_host_load_graph(const char *path, uint32_t *V, uint32_t *E, uint32_t *offsets /*V+1*/,
                        uint32_t *nbrs /*2E*/, uint32_t *labels /*V*/);
/* gen_vde_x + gen_vde, custom.h:492-544.  x, vde: V x e, row-major. */
int gpe_host_gen_vde(uint32_t V, const uint32_t *offsets, const uint32_t *nbrs, const uint32_t *labels,
                     uint32_t e, double *x, double *vde);
/* dfs_query (custom.h:94-119, main.cpp:142-146) + gen_vde(query) + gen_query_pde (custom.h:574-633).
 * Writes at most cap plan paths (vids/labels/degs: n x L, pde: n x L*e) and returns the plan size in *n.  The query arrays
 * are checked first (offsets from 0 and monotone, ids in range, adjacency strictly ascending and symmetric), as in every
 * batch call. */
int gpe_host_query_plan(uint32_t nq, const uint32_t *q_offsets, const uint32_t *q_nbrs, const uint32_t *q_labels,
                        uint32_t L, uint32_t e, uint32_t cap, uint32_t *vids, uint32_t *labels, uint32_t *degs,
                        double *pde, uint32_t *n);

/* ---- data graph + embeddings ------------------------------------------------------------ */

/* The CSR Static_Graph holds (graph.h:61-63).  Simple graph required (no loops, no duplicate
 * edges): the reference's hash-set dedup (custom.h:68-79) is what a duplicate edge would need. */
int gpe_set_graph(gpe_ctx *ctx, uint32_t V, const uint32_t *offsets, const uint32_t *nbrs, const uint32_t *labels);
/* Per-vertex dominance embeddings, V x e row-major: the `vde` gen_vde returns (custom.h:536-540).
 * Host-computed so the FP64 compares see the reference's exact values. */
int gpe_set_embeddings(gpe_ctx *ctx, uint32_t e, const double *vde);

/* ---- seam S1: offline path enumeration (main.cpp:92-96 `dfs` loop, custom.h:66-92) ------ */

/* L = l + 1 vertices per path.  sorted_nodes = line order of membership.txt (main.cpp:80-85);
 * membership[v] in [0, p).  Counts only; returns the row count of every partition (a path belongs to
 * the partition of its FIRST vertex, custom.h:74) and the total.  Path ids are the reference's. */
int gpe_enumerate(gpe_ctx *ctx, uint32_t L, const uint32_t *sorted_nodes, const uint32_t *membership, uint32_t p,
                  uint64_t *rows_per_partition /*p*/, uint64_t *n_rows);
/* Rows [first, first+n) of all_paths.txt in the reference's order (main.cpp:110-119), n x L row-major. */
int gpe_dump_paths(gpe_ctx *ctx, uint64_t first, uint64_t n, uint32_t *vids);
/* First path id of every start vertex, in membership.txt order (V+1 entries); partition_paths.txt
 * (main.cpp:98-108) is the concatenation of [start_row[i], start_row[i+1]) over a partition's vertices. */
int gpe_start_rows(gpe_ctx *ctx, uint64_t *start_row /*V+1*/);

/* ---- seam S2: the dominance filter (gen_pde custom.h:546-572 + Partition ctor :205-266 +
 *      Partition::query :366-489 + the merge main.cpp:166-172) ------------------------------ */

/* Materialise the structure-of-arrays path table (labels, degrees, path embeddings; vertex ids beside
 * it) for the partitions with part_select[i] != 0 (NULL = all).  Needs gpe_enumerate + gpe_set_embeddings. */
int gpe_build_table(gpe_ctx *ctx, const uint8_t *part_select /*p or NULL*/, uint64_t *n_table_rows);
/* How the next gpe_build_table stores the rows.  GPE_TABLE_ROWS: labels, degrees and path embeddings materialised
 * (8L + 8Le bytes per row, 72 at l=2,e=2: what the TMA scan streams) next to the vertex ids (4L).  GPE_TABLE_IDS: the
 * vertex ids only; the scan gathers everything else from packed per-vertex records -- for tables that do not fit
 * materialised (l=3, e=4: 160-byte rows).  GPE_TABLE_AUTO (default): rows while they fit in 70 % of the free HBM.
 * Candidate sets and answers do not depend on the layout. */
#define GPE_TABLE_AUTO 0
#define GPE_TABLE_ROWS 1
#define GPE_TABLE_IDS 2
int gpe_set_table_layout(gpe_ctx *ctx, int layout);
/* Copy the table back in its physical (label-bucketed) row order, for parity checks: any may be NULL. */
int gpe_dump_table(gpe_ctx *ctx, uint64_t first, uint64_t n, uint32_t *vids /*n x L*/, uint32_t *labels /*n x L*/,
                   uint32_t *degs /*n x L*/, double *pde /*n x L*e*/);

#define GPE_FILTER_NO_PRUNE 1u /* streaming mode: every plan path is compared with every table row */
/* Exact mode (off by default because it changes results relative to the reference): every plan path is
 * also compared with the OTHER orientation of every stored row.  The reference stores one orientation per path
 * (custom.h:68-79) and never compares the reverse (:407-435), which loses candidates and under-counts (its quick start
 * prints 45,426 of 221,832 embeddings); with both orientations the candidate sets are complete and the answer is the
 * true number of embeddings.  Accepted by gpe_filter, gpe_batch_upload, gpe_query_batch(es) and the multi-GPU calls. */
#define GPE_FILTER_BOTH_ORIENTATIONS 2u

/* One query: the plan paths (Query_Plan Q of Partition::query) against the whole table.  The candidate
 * sets stay on the device; cand_offsets (nq+1) gives their sizes, survivors (n_qpaths, may be NULL) the
 * number of (plan path, data path) pairs that passed the leaf compare (custom.h:407-435). */
int gpe_filter(gpe_ctx *ctx, uint32_t n_qpaths, const uint32_t *q_vids, const uint32_t *q_labels,
               const uint32_t *q_degs, const double *q_pde, uint32_t nq, uint32_t flags,
               uint64_t *cand_offsets /*nq+1*/, uint64_t *survivors);
/* Sorted, duplicate-free candidate lists of the last gpe_filter, concatenated (std::set order). */
int gpe_get_candidates(gpe_ctx *ctx, uint32_t *cand);

/* ---- seam S3: refinement (custom.h:890-932, called at main.cpp:177) ---------------------- */

/* Matching order (generateGQLQueryPlan :670-722), backward neighbours (:724-755) and the enumeration
 * (:757-888) from caller-supplied candidate sets.  limit = MAX_LIMIT (main.cpp:62-69).  n_matches is
 * min(total, max(limit,1)) like the reference's early exit (:851-854).  order_out / pivot_out (nq, may
 * be NULL) return the plan.  matches (may be NULL) receives up to matches_cap embeddings, nq ids each,
 * indexed by query vertex, in no particular order. */
int gpe_refine(gpe_ctx *ctx, uint32_t nq, const uint32_t *q_offsets, const uint32_t *q_nbrs, const uint32_t *q_labels,
               const uint64_t *cand_offsets, const uint32_t *cand, uint64_t limit, uint64_t *n_matches,
               uint32_t *order_out, uint32_t *pivot_out, uint32_t *matches, uint64_t matches_cap);

/* ---- the whole online stage for a batch of queries (main.cpp:136-179 per query) ---------- */

/* Queries are concatenated: query i has vertices [q_vbase[i], q_vbase[i+1]) of q_labels, local CSR
 * offsets q_offsets[q_vbase[i] + i .. ] (nq_i + 1 entries, starting at 0) and adjacency
 * q_nbrs[q_ebase[i] ..] (local vertex ids).  answers[i] as gpe_refine's n_matches. */
typedef struct gpe_batch {
    uint32_t n_queries;
    const uint32_t *q_vbase;   /* n_queries + 1 */
    const uint32_t *q_ebase;   /* n_queries + 1 */
    const uint32_t *q_offsets; /* sum(nq_i + 1) */
    const uint32_t *q_nbrs;
    const uint32_t *q_labels;
    const uint64_t *limits;    /* n_queries, or NULL for GPE_LIMIT_MAX everywhere */
} gpe_batch;

/* Stage 1 (host): plan every query (gpe_host_query_plan) and copy the batch to the device. */
int gpe_batch_upload(gpe_ctx *ctx, const gpe_batch *batch, uint32_t flags);
/* Stage 2 (device only, asynchronous on the context's stream): tile selection, dominance scan,
 * candidate compaction.  With world > 1 this is the local shard's contribution. */
int gpe_batch_filter(gpe_ctx *ctx);
/* Stage 3 (device only): matching orders + join over the start candidates idx % world == rank. */
int gpe_batch_join(gpe_ctx *ctx, uint32_t rank, uint32_t world);
/* Stage 4: wait, copy the per-query counts back.  Raw totals (not yet clamped by the limit) so that
 * shards can be summed; gpe_clamp_answer applies the reference's limit rule.  A raw total is exact below 2^44
 * and saturated (some value in [2^44, 2^48]) beyond: the reference's answer is min(total, limit) with
 * limit <= UINT_MAX (custom.h:846-855), and every step of the counting join saturates instead of wrapping. */
int gpe_batch_download(gpe_ctx *ctx, uint64_t *raw_counts /*n_queries*/);
uint64_t gpe_clamp_answer(uint64_t raw_total, uint64_t limit);
/* All four stages for one GPU, host buffers in, answers out: the end-to-end call. */
int gpe_query_batch(gpe_ctx *ctx, const gpe_batch *batch, uint32_t flags, uint64_t *answers);

/* Several batches in one call, pipelined: the host plans batch i+1 while the GPU works on batch i (a context holds one
 * batch at a time, so nothing is double-buffered on the device).  answers[i] (n_queries of batch i) as gpe_query_batch.
 * Works with or without a communicator (gpe_comm_init): with one, every rank makes the same call.  gpe_stats.h2d_bytes /
 * d2h_bytes afterwards hold the totals over all the batches. */
int gpe_query_batches(gpe_ctx *ctx, uint32_t n_batches, const gpe_batch *batches, uint32_t flags, uint64_t *const *answers);

/* Candidate exchange between shards (replaces the serial merge main.cpp:166-172 when the table is
 * sharded over GPUs).  Device pointers: the caller moves them with NCCL. */
int gpe_batch_cand_info(gpe_ctx *ctx, uint64_t *n_slots, uint64_t *n_cand_total);
int gpe_batch_cand_export(gpe_ctx *ctx, void *d_counts_u32 /*n_slots*/, void *d_cand_u32 /*n_cand_total*/);
/* Replace the batch's candidate sets by the union of `world` shards' lists.  d_counts: world x n_slots
 * (u32), d_cand: world x stride (u32), shard r's lists concatenated at d_cand + r*stride.  The lists that
 * gpe_batch_cand_export writes are in the library's device-side id space (vertices numbered by (label, id): the same
 * numbering on every GPU that loaded the same graph), ascending inside a slot; they are meant for this call. */
int gpe_batch_cand_merge(gpe_ctx *ctx, uint32_t world, const void *d_counts, const void *d_cand, uint64_t stride);
/* The same exchange in bitmap form, the one the multi-GPU engine uses: fixed size (no count exchange, no host
 * sync), one all-gather, and the union is fused into the compaction's popcount pass.
 *   gpe_batch_scan         stage 2 without the compaction: tile selection + dominance scan of the local shard;
 *   gpe_batch_bitmap       device pointer and size of the batch's candidate bitmaps: n_slots x words, bit i of slot s
 *                          = the i-th vertex (ascending id) of the slot's label -- identical layout on every shard;
 *   gpe_batch_bitmap_merge d_all = `world` such bitmaps one after the other (the all-gather's output, device memory);
 *                          replaces the local bitmaps by their OR and builds the sorted candidate lists.
 * All three are asynchronous on the context's stream: no host sync (the candidate total stays on the device until
 * gpe_batch_download or a candidate getter needs it). */
int gpe_batch_scan(gpe_ctx *ctx);
int gpe_batch_bitmap(gpe_ctx *ctx, void **d_bitmap, uint64_t *n_bytes);
int gpe_batch_bitmap_merge(gpe_ctx *ctx, uint32_t world, const void *d_all);
/* Per-query-vertex candidate counts / lists of the current batch (after filter or merge), host side. */
int gpe_batch_get_candidates(gpe_ctx *ctx, uint64_t *cand_offsets /*n_slots+1*/, uint32_t *cand /*or NULL*/);
int gpe_batch_get_plan(gpe_ctx *ctx, uint32_t *order /*n_slots*/, uint32_t *pivot /*n_slots*/);

/* ---- multi-GPU (replaces the per-partition OpenMP loop + serial merge, main.cpp:160-172) -----------------------------
 * The path table is sharded by the reference's own partitions: a path belongs to the partition of its FIRST vertex
 * (custom.h:74), GPU r of N holds the partitions i % N == r (enumerate with p >= N).  Per batch there is ONE exchange --
 * an all-gather of the shards' candidate bitmaps, whose union is fused into the compaction -- then every GPU joins the
 * start candidates idx % N == r and the match counts are summed.  NCCL is bound at run time (libnccl.so.2); a
 * single-GPU user never loads it.  Two ways to run it:
 *   one process per GPU:  rank 0 calls gpe_comm_unique_id and hands the 128 bytes to the others (any transport);
 *                         every rank: gpe_comm_init, gpe_build_table_shard, then per batch gpe_batch_upload,
 *                         gpe_batch_step, gpe_batch_finish (collective calls: same order on every rank);
 *   one process, N GPUs:  N contexts, gpe_comm_init_all, gpe_build_table_shard on each, then gpe_multi_query_batch
 *                         (or its three stages) -- what `host/main -g N` does. */
#define GPE_COMM_ID_BYTES 128
int gpe_comm_unique_id(void *id_out /*GPE_COMM_ID_BYTES*/);
int gpe_comm_init(gpe_ctx *ctx, int rank, int world, const void *id /*GPE_COMM_ID_BYTES*/);
int gpe_comm_init_all(gpe_ctx **ctxs, int n);
int gpe_comm_destroy(gpe_ctx *ctx); /* also done by gpe_destroy */
int gpe_comm_info(gpe_ctx *ctx, int *rank, int *world, int *nccl_version /*e.g. 22703, 0 if NCCL is not loaded*/);
/* gpe_build_table for the partitions of this context's rank (all of them without a communicator). */
int gpe_build_table_shard(gpe_ctx *ctx, uint64_t *n_table_rows);
/* Stages 2 + 3 of an uploaded batch, asynchronous on the context's stream: scan of the local shard, candidate
 * exchange, compaction, join of this rank's start candidates.  Without a communicator: gpe_batch_filter + gpe_batch_join. */
int gpe_batch_step(gpe_ctx *ctx);
/* Stage 4: wait, sum the shards' counts (all-reduce), apply the limit rule: answers as gpe_query_batch returns them. */
int gpe_batch_finish(gpe_ctx *ctx, uint64_t *answers /*n_queries*/);
/* One thread driving N contexts (contexts of gpe_comm_init_all, in rank order): the batch is planned once on the host
 * and uploaded to every GPU; the GPUs run concurrently, NCCL calls are grouped. */
int gpe_multi_batch_upload(gpe_ctx **ctxs, int n, const gpe_batch *batch, uint32_t flags);
int gpe_multi_batch_step(gpe_ctx **ctxs, int n);
int gpe_multi_batch_finish(gpe_ctx **ctxs, int n, uint64_t *answers);
int gpe_multi_query_batch(gpe_ctx **ctxs, int n, const gpe_batch *batch, uint32_t flags, uint64_t *answers);

/* ---- GNN-PGE: the reference's sibling variant of the filter (GNN-PGE/src/main.cpp, GNN-PGE/include/custom.h) --------
 * One row per data VERTEX: the bounding box of the embeddings of all simple paths of `pl` vertices that start at it
 * ("path group", src/main.cpp:91-176), over the dominance embeddings and over the label embeddings.  A data vertex v
 * is a candidate of a query vertex u iff label and degree fit, the label boxes overlap and v's upper corner is not
 * below u's lower corner (include/custom.h:332-367).  Matching order and join are shared with the path filter.
 * Verified on B200 against the golden vectors of the unmodified GNN-PGE binary (tests/test_gpu_pge.py). */
/* x: the label embeddings per vertex (V x e) as gpe_host_gen_vde returns them; needs gpe_set_graph + gpe_set_embeddings.
 * pl = vertices per path (GNN-PGE's -l, default 2; 1..4 here). */
int gpe_pge_build(gpe_ctx *ctx, uint32_t pl, const double *x);
/* The same groups on the host (pure host code, no GPU): what the batch upload computes for every query graph, and a
 * CPU statement of gpe_pge_build + gpe_pge_dump_groups for any graph.  pg, plg: V x (pl*e) x [lo, hi]. */
int gpe_host_pge_groups(uint32_t V, const uint32_t *offsets, const uint32_t *nbrs, const uint32_t *labels, uint32_t pl,
                        uint32_t e, double *pg, double *plg, uint8_t *has);
/* The groups in vertex order, V x (pl*e) x [lo, hi] each, as src/main.cpp:179-194 writes them; has[v] = 0 marks a vertex
 * without any such path (its box is then [vde, vde | 0 ...]). */
int gpe_pge_dump_groups(gpe_ctx *ctx, double *pg, double *plg, uint8_t *has);
/* The batch stages with the GNN-PGE filter in place of the path filter; continue with gpe_batch_join /
 * gpe_batch_download (or gpe_batch_get_candidates). */
int gpe_pge_batch_upload(gpe_ctx *ctx, const gpe_batch *batch);
int gpe_pge_batch_filter(gpe_ctx *ctx);
int gpe_pge_query_batch(gpe_ctx *ctx, const gpe_batch *batch, uint64_t *answers);

/* ---- measurement hooks ------------------------------------------------------------------- */
typedef struct gpe_stats {
    uint64_t table_rows, table_tiles, tile_rows, row_bytes;  /* row_bytes = L*4 + L*4 + L*e*8 */
    uint64_t scan_items;        /* (query-path block, tile) pairs the last scan examined */
    uint64_t scan_items_unpruned;
    uint64_t scan_rows;         /* rows examined = scan_items * tile_rows (tail tile included) */
    uint64_t scan_launches, select_launches, compact_launches, join_launches, build_launches;
    float last_scan_ms, last_select_ms, last_compact_ms, last_join_ms, last_build_ms, last_enumerate_ms;
    uint64_t n_qpaths, n_qblocks, n_slots, n_candidates, join_items;
    uint64_t kernel_launches;          /* kernels launched by this context since creation */
    uint64_t h2d_bytes, d2h_bytes;     /* host<->device bytes of the last batch (upload .. download) */
    uint64_t join_exports, join_donations, join_steps;  /* last join: subtree hand-overs between threads, DFS steps summed over threads */
    uint64_t join_warp_iters, join_idle_polls;          /* last join: warp loop iterations with / without a busy lane (lane utilisation = steps / (32 x warp_iters)) */
    uint64_t join_bfs;        /* last join ran level-synchronously (counting, no answer limits) instead of depth-first */
    uint64_t join_fallbacks;  /* level-synchronous joins recomputed depth-first because a frontier outgrew its buffer */
    uint64_t join_reruns;     /* batches some of whose queries were joined a second time with their weighted counted leaves
                                 walked, because a weighted count met a saturated (>= 2^62) table entry */
    uint64_t table_ids_only;  /* 1: the table holds vertex ids only (GPE_TABLE_IDS) */
    uint64_t stored_row_bytes; /* bytes per row actually held in HBM: 4L + row_bytes, or 4L for an ids-only table */
    uint64_t exchange_bytes;  /* multi-GPU: bytes this GPU received in the last candidate exchange (dense bitmaps or sparse pairs) */
    uint64_t exchange_redos;  /* steps redone with the dense exchange because a shard outgrew the sparse buffer */
    /* last gpe_query_batches call, host wall clock summed over its batches: staging + H2D enqueue of a planned batch, kernel /
     * collective enqueue, planning of the NEXT batch (overlaps the GPU), waiting for the GPU + D2H + all-reduce */
    float host_upload_ms, host_enqueue_ms, host_plan_ms, host_finish_ms;
} gpe_stats;
int gpe_get_stats(gpe_ctx *ctx, gpe_stats *out);
/* The context's CUDA stream (cudaStream_t) so callers can bracket calls with their own events. */
void *gpe_stream(gpe_ctx *ctx);
int gpe_sync(gpe_ctx *ctx);
/* Per-stage CUDA-event timing on the context's stream.  mode 0: off (default).  mode 1: synchronous, each
 * stage is followed by an event sync and lands in gpe_stats.last_*_ms.  mode 2: deferred, event pairs are
 * queued without any extra sync and summed by gpe_collect_timings -- the way to time kernels inside a
 * timed region without perturbing it. */
int gpe_set_timing(gpe_ctx *ctx, int mode);
#define GPE_STAGE_SELECT 0
#define GPE_STAGE_SCAN 1
#define GPE_STAGE_COMPACT 2
#define GPE_STAGE_JOIN 3
#define GPE_STAGE_ENUMERATE 4
#define GPE_NUM_STAGES 5
/* Syncs the stream, returns per-stage summed milliseconds and launch counts since the last call. */
int gpe_collect_timings(gpe_ctx *ctx, double *sum_ms /*GPE_NUM_STAGES*/, uint64_t *count /*GPE_NUM_STAGES*/);

#ifdef __cplusplus
}
#endif
#endif /* GPE_H_ */
