// gpe_internal.h -- context, device buffers and launch declarations shared by the .cu files.
// Product code: nothing here may include or call anything under oracle/.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "gpe.h"
#include "host_ref.h"

typedef uint32_t u32;
typedef uint64_t u64;

namespace gpe {

// ---- compile-time geometry of the scan -------------------------------------------------------
constexpr int kTileRows = 256;   // rows per tile; one consumer thread per row
constexpr int kQB = 8;           // query paths per query-path block
constexpr int kStages = 4;       // TMA pipeline depth per CTA
constexpr int kMaxL = 4;         // path positions supported by the compiled kernels (l = 2, 3)
constexpr int kMaxE = 8;
constexpr u32 kKeyBudget = 1u << 22;  // max label-sequence buckets of the table directory
constexpr double kEps = 1e-6;    // custom.h:43
constexpr int kMaxDevices = 64;  // per-device caches of launch configurations

// Grow-only device buffer.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes + bytes / 4 + 256);
        if (e == cudaSuccess) cap = bytes + bytes / 4 + 256;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Device-side view of the data graph and the per-vertex tables (replicated on every GPU).
struct GraphView {
    u32 V;
    const u32 *off;    // V+1
    const u32 *nbr;    // 2E, ascending per vertex
    const u32 *label;  // V
    const u32 *deg;    // V
    const u32 *rank;   // V: position in membership.txt
    const u64 *ranklab;  // V: rank | label << 32 (the enumeration walks gather both with one load)
    const double *vde; // V x e
    u32 e;
    const u32 *lpos;   // V: position of a vertex inside its label class
    const u32 *lclass; // vertices by (label, id)
    const u32 *lcoff;  // labels + 1 class offsets
    const u32 *nbrG;   // adjacency grouped by neighbour label (ascending id inside a group); same offsets as nbr
};

// The join's own copy of the data graph, in CLASS ORDER (built on the device by k0_graph.cu): vertex v has the id
// v' = lcoff[label(v)] + lpos(v).  Sorted by v' an adjacency list is grouped by neighbour label with ids ascending inside
// every group, and "the i-th vertex of label l" (candidate bitmaps, subtree tables) is v' - lcoff[l] without a lookup.
struct JoinGraph {
    u32 V, nl;
    const u32 *nbrJ;            // adjacency entries, rows by v': narrow u32 (v' | min(degree,255) << 24) or wide uint2
    const unsigned char *gtab;  // group directory, row v' at gtab + v' * dir_row_bytes: narrow u32 base | u16 rel[nl+1], wide u32 abs[nl+1]
    u32 dir_row_bytes;
    bool wide_adj, wide_dir;
    const u32 *degJ, *labelJ;   // per vertex, class order
    const u32 *lcoff;           // nl + 1: first v' of every label
    const u32 *orig;            // v' -> the caller's vertex id (= lclass)
    const u64 *tpool;           // subtree tables (k3_tree_tables); the walk's view is shifted down by V entries (see k3_order)
    const u64 *bloom;           // edge filter: two bits per undirected edge, both inside one 64-bit word
    u64 bloom_word_mask;
};

// Blocked edge filter hash: both bits of an edge live in ONE 64-bit word (one 8-byte load per test), a handful of
// 32-bit multiplies (the r01k capture had 16 % of the join's instructions in two 64-bit mixers per test).
__host__ __device__ __forceinline__ void join_edge_probe(u32 a, u32 b, u64 word_mask, u64 &word, u64 &bits) {
    const u32 lo = a < b ? a : b, hi = a < b ? b : a;
    u32 x = lo * 0x9E3779B1u + hi * 0x85EBCA77u;
    x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12;
    u32 y = (lo ^ 0x68E31DA4u) * 0xB5297A4Du + hi * 0x1B56C4E9u;
    y ^= y >> 16;
    word = (u64)x & word_mask;
    bits = (1ull << (y & 63)) | (1ull << (y >> 6 & 63));
}

// Physical layout of the path table: tile-major blocked structure-of-arrays.
//   scan tile t (tile_bytes each): labels[L][R] u32 | degs[L][R] u32 | pde[L*e][R] f64
//   vids tile t:                   vids[L][R] u32   (separate array, only survivors read it): the POSITION of the
//                                  vertex inside its label class (lpos), which is the bit index of the candidate
//                                  bitmaps; the id is lclass[lcoff[label] + position]
struct TableView {
    u32 L, E, D;
    u64 n_rows, n_tiles;
    u32 tile_bytes;
    bool ids_only;  // no scan tiles and no summaries: `vids` holds the rows' vertex ids, everything else is gathered (k2_scan_ids)
    unsigned char *tiles;
    u32 *vids;
    // per-tile summaries (the analogue of Partition::build_auxiliary_index, custom.h:268-364)
    u32 *lab_min, *lab_max, *deg_max;  // [L][n_tiles]
    double *pde_max;                   // [D][n_tiles]
    // label-sequence directory
    u32 key_radix[kMaxL], key_stride[kMaxL];
    u32 n_keys;
    const u64 *bucket_start;  // n_keys + 1
};

// One query-path block: up to kQB plan paths that share a table bucket, laid out so that one
// cp.async.bulk brings it into shared memory next to the tile it is compared with.
template <int L, int E>
struct alignas(16) QBlockRec {
    u32 n;
    u32 first_qpath;  // index of the block's first plan path (survivor counters)
    u32 pad[2];
    u32 labels[kQB][L];
    u32 degs[kQB][L];
    u32 slot[kQB][L];   // candidate bitmap of (query, plan path vertex k)
    u32 qpath[kQB];
    double pde[kQB][L * E];
};

struct QBlockHost {  // layout-agnostic staging on the host
    u32 n, first_qpath;
    u32 t0, t1;  // tile range from the directory
};

// Per-depth join plan of one query (filled by the order kernel), in EXECUTION order: depth 0 is the
// reference's start vertex (generateGQLQueryPlan, custom.h:670-722); the remaining vertices are ordered by the
// same greedy rule, except that query leaves are moved to the end ("tail") where their completions are counted
// instead of walked.  The number of embeddings does not depend on the order below the start vertex.
struct alignas(16) JoinDepth {
    u32 u;            // query vertex matched at this depth
    u32 label;
    u32 deg;
    u32 pivot_depth;  // depth at which the pivot (first earlier query neighbour) was matched
    u64 bn_mask;      // depths of the other backward neighbours (generateBN, custom.h:724-755)
    u64 tree_off;     // table of everything that hangs below this vertex in peeled subtrees (pool offset), or kNoTree
    u64 tail_mask;    // [tail depths] prefix depths that may sit in this leaf's label group and need an edge test;  [walked depths >= 1]
                      // earlier depths of the same label (the only ones a candidate could coincide with);  [depth 0] walk starts from: 0 list, 1 class
    u32 kid_begin, kid_count;  // later depths (walked or tail) whose pivot is this depth, as a slice of the query's kid list
                               // (depth, label): their label groups are looked up when this depth is matched
    u64 units_mask;   // [walked depths] heads of the counted-tail units whose factor becomes computable at this depth
    u32 tail_k;       // [depth 0] number of tail depths;  [tail depths] operation, see kTail*
    u32 sure_used;    // [tail depths] prefix vertices known to sit in the group (same label, query-adjacent to the pivot)
};
static_assert(sizeof(JoinDepth) == 64, "JoinDepth is read as four 16-byte words");
constexpr u32 kTailMul = 0;    // multiply by the free members of the leaf's label group
constexpr u32 kTailFall = 1;   // same pivot and label as the previous tail depth: multiply by (previous factor - 1)
constexpr u32 kTailPairA = 2;  // two same-label leaves on different pivots: |A||B| - |A n B| (this depth and the next)
constexpr u32 kTailPairB = 3;
constexpr u32 kTailW = 0x100;  // flag: the leaf carries peeled subtrees; its members weigh N_u[y] (tree_off), the sum over a
                               // pivot's group is tabulated as S_u (units_mask), bn_mask names the prefix vertices surely in it
constexpr u64 kNoTree = ~0ull;

// One per query vertex slot: the table of a vertex with peeled children (level 0 = none).
struct TreeJob {
    u32 level;       // 1 + the highest level among its peeled children (children are tabulated first)
    u32 label, qdeg;
    u32 start_slot;  // this vertex is the peeled start vertex: candidate bitmap its matches must be in, else 0xffffffff
    u64 table_off;   // into the table pool, one u64 per vertex of the label class
    u32 child_begin, n_child;  // peeled children: slice of the child-slot array
};

// Work queue of the join (device memory): tickets [0, n_init) are the start-candidate items, later tickets are
// subtrees exported on demand by busy threads.
struct alignas(16) JoinQueue {
    unsigned long long head;     // tickets claimed by consumers
    unsigned long long tail;     // items produced
    long long pending;           // items published and not yet retired; 0 <=> the join is complete
    unsigned long long idle;     // warps without any busy lane
    unsigned long long n_init;
    unsigned long long steps;    // DFS steps, summed over threads
    unsigned long long exports;  // hand-overs between warps (through the queue)
    unsigned long long donations;  // hand-overs inside a warp (through shared memory)
    unsigned long long full;     // set once the item buffer overflowed (exports stop; result unaffected)
    unsigned long long warp_iters;  // loop iterations of warps that had at least one busy lane (32 x this = lane slots)
    unsigned long long lane_iters;  // unused
    unsigned long long idle_polls;  // loop iterations of warps without any busy lane
};

}  // namespace gpe

struct gpe_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int timing = 0;  // 0 off, 1 synchronous per stage, 2 deferred (events queued, read by gpe_collect_timings)
    struct Span { cudaEvent_t a, b; int stage; };
    std::vector<Span> spans;
    size_t spans_used = 0;
    std::string err;
    gpe_stats stats{};

    // graph
    u32 V = 0, n_adj = 0, n_labels = 0, max_degree = 0;
    double branching = 0.0;  // (sum d^2 / sum d) / labels: growth factor of a walk over unique-label query vertices
    gpe::DevBuf d_off, d_nbr, d_label, d_deg, d_rank, d_sorted, d_member, d_vde, d_vrec, d_ranklab;
    gpe::DevBuf d_nbrJ, d_gtab, d_offJ, d_degJ, d_labelJ, d_newid;  // the join's class-ordered copy of the graph (k0_graph.cu)
    gpe::DevBuf d_nbrG;  // adjacency grouped by neighbour label in the caller's ids (k1 histogram / fill)
    gpe::DevBuf d_lclass, d_lpos, d_lcoff;  // label classes: vertices by (label, id), position in class, class offsets
    bool wide_adj = false, wide_dir = false;
    u32 dir_row_bytes = 0;
    gpe::DevBuf d_tjobs, d_tchild, d_tpool, d_tcursor, d_tlist, d_qcur;  // subtree tables of the join (jobs, child lists, value pool, pool cursor)
    u32 max_class = 0;
    gpe::DevBuf d_bloom;  // edge filter of the join
    u64 bloom_bits = 0;
    gpe::DevBuf d_pge, d_pge_x, d_pge_q;  // GNN-PGE: path groups of the data vertices, label embeddings, query records
    u32 pge_pl = 0;
    bool have_pge = false, b_pge = false, b_pge_rows_pending = false;
    u32 b_rank = 0, b_world = 1;
    gpe::DevBuf d_items, d_ready, d_jq, d_init, d_kids;  // exported join work items, their publication flags, the queue header, start tickets
    u32 join_epoch = 0;
    u32 b_max_nq = 0;
    u32 e = 0;
    bool have_graph = false, have_emb = false, have_enum = false, have_table = false;

    // enumeration
    u32 L = 0, p = 0;
    gpe::DevBuf d_offr;       // u32 V+1: rank-ordered CSR offsets
    gpe::DevBuf d_ebase;      // u64 n_adj+1: first path id of every rank-ordered (a,b) slot
    gpe::DevBuf d_scan_tmp;   // scratch of the device-wide scan
    gpe::DevBuf d_start_rows; // u64 V+1 (+p): first path id of every start vertex, rank order
    u64 n_rows = 0;
    std::vector<u32> h_member;

    // table
    gpe::TableView tv{};
    int table_layout = 0;  // gpe_set_table_layout: 0 auto, 1 materialised rows, 2 ids only
    gpe::DevBuf d_tiles, d_vids, d_sum_u32, d_sum_f64, d_bucket, d_cursor;
    std::vector<u64> h_bucket_start;

    // batch / filter state
    u32 b_nq = 0;          // queries
    u32 b_slots = 0;       // sum of query vertices
    u32 b_qpaths = 0, b_qblocks = 0;
    u32 b_flags = 0;
    u64 b_words = 0;       // bitmap words per slot
    u64 b_items_cap = 0, b_items_unpruned = 0;
    u64 b_n_cand = 0;
    bool b_scanned = false, b_filtered = false, b_joined = false, b_cand_external = false, b_cand_clean = false;
    std::vector<u32> h_q_vbase, h_q_ebase, h_q_offsets, h_q_nbrs, h_q_labels;
    std::vector<u64> h_limits;
    std::vector<u32> h_slot_query;  // slot -> query
    gpe::DevBuf d_qblocks, d_qb_t0, d_qb_prefix, d_worklist, d_counters, d_bitmap, d_survivors, d_slot_label;
    gpe::DevBuf d_chunk_cnt, d_chunk_off, d_cand, d_cand_off;
    gpe::DevBuf d_q_vbase, d_q_ebase, d_q_offsets, d_q_nbrs, d_q_labels, d_limits;
    gpe::DevBuf d_order, d_pivot, d_jplan, d_item_base, d_answers, d_matches, d_match_cursor, d_qmode;
    gpe::PinnedBuf h_pin, h_pin2, h_pin3, h_pin_q;  // staging: query-path blocks, answers, compaction results, query arrays
    cudaEvent_t ev_upload = nullptr;                // the last upload's copies out of h_pin / h_pin_q are done
    u64 b_cand_cap = 0, cand_seen_max = 0;          // entries d_cand holds; largest candidate total of any batch so far
    bool b_cand_external_lists = false;             // candidate lists were supplied by the caller (gpe_refine): no compaction ran
    bool b_cand_known = false;                      // b_n_cand / scan counters have been read back for this batch
    // multi-GPU (gpe_comm_*): NCCL communicator of this context, its rank, the all-gathered shard bitmaps
    void *comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    gpe::DevBuf d_all_bitmaps, d_reduce, d_sparse;
    u64 sparse_cap = 0, sparse_seen_max = 0;  // pairs per shard buffer of the sparse exchange; most non-zero words a shard ever had
    bool b_sparse_used = false, force_dense = false, need_dense_redo = false;
    gpe::PinnedBuf h_pin4;                    // the shards' non-zero word counts of the last sparse exchange
    gpe::LabelTable label_table;  // label embeddings of the queries seen so far (host planning)
    u64 b_chunks_per_slot = 0;

    int fail(int code, const char *fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
};

#define GPE_CUDA(ctx, expr)                                                                          \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return (ctx)->fail(GPE_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,           \
                               cudaGetErrorString(e__));                                             \
    } while (0)

namespace gpe {

// ---- launchers (defined in the .cu files) ------------------------------------------------------
// device-wide exclusive scan of n u64 values, in place; total returned in d_total (device) if non-null
cudaError_t exclusive_scan_u64(u64 *d_data, u64 n, DevBuf &tmp, cudaStream_t s);
u64 exclusive_scan_launches(u64 n);

// K0: device-side construction of what gpe_set_graph derives from the CSR (k0_graph.cu)
// err3 (4 words): {smallest (code << 32 | vertex) or ~0, max label, max degree, sum of squared degrees}; codes 1 offsets, 2 id range, 3 self loop, 4 order
cudaError_t k0_validate(u32 V, u32 n_adj, const u32 *off, const u32 *nbr, const u32 *label, u32 *deg, u64 *err3, cudaStream_t s);
cudaError_t k0_build_classes(u32 V, u32 n_labels, const u32 *label, const u32 *deg, u32 *lcoff /*n_labels + 2*/, u32 *lclass,
                             u32 *lpos, u32 *newid, u32 *degJ, u32 *labelJ, u32 *offJ /*V + 1*/, u32 *max_class_dev, DevBuf &tmp,
                             cudaStream_t s);
cudaError_t k0_build_join_graph(u32 V, u32 n_adj, u32 n_labels, const u32 *off, const u32 *nbr, const u32 *lclass,
                                const u32 *newid, const u32 *degJ, const u32 *offJ, const u32 *lcoff, bool wide_adj,
                                bool wide_dir, u32 dir_row_bytes, u32 *nbrJ, u32 *nbrG, void *gtab, u64 *bloom, u64 bloom_bits,
                                DevBuf &tmp, int sm_count, cudaStream_t s);
cudaError_t k0_gather(u64 n, const u32 *map, const u32 *in, u32 *out, cudaStream_t s);  // out[i] = map[in[i]]
struct ZeroList {  // up to 8 small regions (32-bit words) zeroed by one launch
    u32 *p[8];
    u64 words[8];
    int n = 0;
    void add(void *ptr, size_t bytes) { p[n] = reinterpret_cast<u32 *>(ptr); words[n] = bytes / 4; n++; }
};
cudaError_t k0_zero(const ZeroList &z, cudaStream_t s);

// K1
cudaError_t k1_rank_labels(u32 V, const u32 *rank, const u32 *label, u64 *ranklab, cudaStream_t s);
cudaError_t k1_count(const GraphView &g, u32 L, const u32 *sorted, const u32 *offr, u64 *cnt_r, int sm_count,
                     cudaStream_t s);
cudaError_t k1_rows_per_partition(u32 V, const u32 *sorted, const u32 *offr, const u64 *ebase, const u32 *member,
                                  u64 *part_rows, u64 *start_rows, cudaStream_t s);
cudaError_t k1_dump(const GraphView &g, u32 L, const u32 *sorted, const u32 *offr, const u64 *ebase, u32 rank_lo,
                    u32 rank_hi, u64 first, u64 n, u32 *out, cudaStream_t s);
cudaError_t k1_histogram(const GraphView &g, const TableView &t, const u32 *sorted, const u32 *member,
                         const unsigned char *part_sel, u64 *hist, int sm_count, cudaStream_t s);
cudaError_t k1_fill(const GraphView &g, const TableView &t, const u32 *sorted, const u32 *member,
                    const unsigned char *part_sel, u64 *cursor, int sm_count, cudaStream_t s);
// vertex ids (written by k1_fill) -> scan tiles + class positions + per-tile summaries
// packed per-vertex records (label, degree, class position | embedding) that k1_expand gathers from
size_t k1_vertex_record_bytes(u32 V, u32 e);
cudaError_t k1_vertex_records(const GraphView &g, void *vrec, cudaStream_t s);
cudaError_t k1_expand(const TableView &t, const void *vrec, int sm_count, cudaStream_t s);
cudaError_t k1_dump_table(const TableView &t, const GraphView &g, const void *vrec, u64 first, u64 n, u32 *vids, u32 *labels,
                          u32 *degs, double *pde, cudaStream_t s);

// K2
size_t qblock_rec_bytes(u32 L, u32 E);
// labels/degs/slots: n x L, pde: n x L*E (already gathered in block order); qpath_ids: n
void qblock_pack(u32 L, u32 E, void *dst_rec, u32 n, u32 first_qpath, const u32 *qpath_ids, const u32 *labels,
                 const u32 *degs, const u32 *slots, const double *pde);
cudaError_t k2_select(const TableView &t, const void *qblocks, const u32 *qb_t0, const u64 *qb_prefix, u32 n_qblocks,
                      u64 n_items, bool prune, u64 *worklist, u64 *counters, cudaStream_t s);
// with_vids: the tiles' vertex ids travel with them through the TMA ring (pruned work lists, survivors common)
cudaError_t k2_scan(const TableView &t, const void *qblocks, const u64 *worklist, const u64 *counters, u32 *bitmap,
                    u64 words_per_slot, u64 *survivors, bool with_vids, int sm_count, cudaStream_t s);
// the same scan over an ids-only table: rows are gathered from the packed vertex records (k1_vertex_records)
cudaError_t k2_scan_ids(const TableView &t, const void *vrec, const void *qblocks, const u64 *worklist, const u64 *counters,
                        u32 *bitmap, u64 words_per_slot, u64 *survivors, int sm_count, cudaStream_t s);
bool k2_supported(u32 L, u32 E);

// K3 (candidate compaction, matching order, join)
constexpr u32 kChunkWords = 256;  // bitmap words per compaction chunk
cudaError_t k3_chunk_count(const u32 *bitmap, u64 words_per_slot, u64 chunks_per_slot, u32 n_slots, u64 *chunk_cnt,
                           cudaStream_t s);
// same, fused with the union of `world` all-gathered shard bitmaps (shard r at all + r * shard_words) into `bitmap`
cudaError_t k3_merge_count(const u32 *all, u64 shard_words, u32 world, u32 *bitmap, u64 words_per_slot,
                           u64 chunks_per_slot, u32 n_slots, u64 *chunk_cnt, cudaStream_t s);
// bit i of slot s stands for vertex lclass[lcoff[slot_label[s]] + i]
// (candidate lists on the device hold class-order ids: first id of the slot's label + bit position)
// cap: entries `cand` holds; a chunk that would write beyond it is dropped and *overflow set
cudaError_t k3_compact(const u32 *bitmap, u64 words_per_slot, u64 chunks_per_slot, u32 n_slots, const u64 *chunk_off,
                       const u32 *slot_label, const u32 *lcoff, u32 n_labels, u32 *cand, u64 *cand_off, u64 cap,
                       u64 *overflow, cudaStream_t s);
cudaError_t k3_scatter(const u32 *counts, const u32 *cand, u64 stride, u32 world, u32 n_slots, const u32 *slot_label,
                       const u32 *lcoff, u32 n_labels, u32 *bitmap, u64 words_per_slot, u64 *prefix_tmp /*world x n_slots*/,
                       cudaStream_t s);
// sparse candidate exchange: non-zero bitmap words as (index, word) pairs in a fixed-capacity buffer (u64 count | u64 | pairs)
cudaError_t k3_sparse_pack(const u32 *bitmap, u64 n_words, u64 cap, void *buf, int sm_count, cudaStream_t s);
cudaError_t k3_sparse_merge(const void *all, u64 stride_bytes, u32 world, u32 my_rank, u64 cap, u32 *bitmap, int sm_count,
                            cudaStream_t s);
cudaError_t k3_counts_from_offsets(const u64 *cand_off, u32 n_slots, u32 *counts, cudaStream_t s);
cudaError_t k3_order(u32 n_queries, u32 V, const u32 *q_vbase, const u32 *q_ebase, const u32 *q_offsets,
                     const u32 *q_nbrs, const u32 *q_labels, const u64 *cand_off, u32 *order, u32 *pivot,
                     JoinDepth *jplan, void *kids /*uint2 per query vertex*/, u64 *item_base, u32 rank, u32 world,
                     bool enumerate /*walk every vertex (matches wanted)*/, bool clean_start /*start candidates carry the
                     query label (they come from the filter)*/, u32 n_labels, const u32 *lcoff, TreeJob *tjobs, u32 *tchild,
                     u64 *tcursor, u32 *tcount /*kMaxTreeLevels, zeroed*/, u32 *tlist /*kMaxTreeLevels x 2 n_slots*/, u32 n_slots,
                     bool allow_weighted /*counted leaves may carry peeled subtrees (depth-first kernel only)*/,
                     const u32 *qmode /*per query or null: 0 as usual, 1 no weighted leaves, 2 leave the query out*/,
                     float branching /*growth factor of a walk per depth (graph statistic): decides whether tabulating peeled
                     subtrees over whole label classes pays for a query*/,
                     u64 pool_cap /*entries of the table pool: a query whose tables do not fit walks instead*/, cudaStream_t s);
// tables of the peeled subtrees, levels 1..max_level (one launch each)
constexpr u32 kMaxTreeLevels = GPE_MAX_QUERY_VERTICES;
cudaError_t k3_tree_tables(const JoinGraph &jv, u32 n_slots, u32 max_class, u32 max_level, const TreeJob *tjobs,
                           const u32 *tchild, const u32 *tcount, const u32 *tlist, const u32 *bitmap, u64 words_per_slot,
                           u64 *tpool, int sm_count, cudaStream_t s);
u32 k3_item_stride(u32 max_nq);  // u32 words per exported work item
// one ticket (query, position in cand[]) per start candidate of this shard; init: 8 bytes per ticket
cudaError_t k3_init_items(const JoinGraph &jv, u32 n_queries, const u32 *q_vbase, const JoinDepth *jplan,
                          const u64 *cand_off, const u32 *cand, const u64 *item_base, u32 rank, u32 world,
                          u32 heavy_deg /*roots of at least this degree are ticketed first*/, u64 *cursors /*6 per query, zeroed*/,
                          void *init, JoinQueue *jq, bool use_tables /*subtree tables are valid: dead roots get no ticket*/,
                          const u64 *cand_overflow /*device flag or null: set => the candidate lists are truncated, no tickets*/,
                          int sm_count, cudaStream_t s);
// one persistent launch: every thread runs work items (explicit-stack DFS) and exports subtrees when others starve
cudaError_t k3_dfs(const JoinGraph &jv, u32 max_nq, const u32 *q_vbase, const JoinDepth *jplan, const void *kids, const u32 *cand,
                   const void *init, const u64 *limits, u64 *answers, u32 *items, u64 export_cap, u32 *ready, u32 epoch,
                   JoinQueue *jq, u32 *matches, u64 matches_cap, u64 *match_cursor, u64 *inexact /*per query, zeroed: set when a
                   weighted count met a saturated operand*/, int sm_count, cudaStream_t s);

// K4: GNN-PGE (per-vertex path groups), see k4_pge.cu.  Rows in class order (index = lcoff[label] + lpos), columns by
// dimension: pg_lo/pg_hi/plg_lo/plg_hi [pde][V], deg [V], has [V].
struct PgeView {
    double *pg_lo, *pg_hi, *plg_lo, *plg_hi;
    u32 *deg;
    unsigned char *has;
};
size_t k4_pge_bytes(u32 V, u32 pde);
PgeView k4_pge_view(void *buf, u32 V, u32 pde);
cudaError_t k4_pge_groups(const GraphView &g, u32 pl, const double *d_x, const PgeView &p, int sm_count, cudaStream_t s);
cudaError_t k4_pge_scan(const PgeView &p, u32 V, u32 pde, u32 n_labels, const u32 *lcoff, const u32 *label_slot_off,
                        const u32 *slot_list, const u32 *q_deg, const double *q_pg_lo, const double *q_plg_lo,
                        const double *q_plg_hi, u32 *bitmap, u64 words_per_slot, u64 *survivors, u64 *rows_examined /*rows of the
                        label classes some query vertex asks for*/, int sm_count, cudaStream_t s);
bool k4_pge_supported(u32 pde);
cudaError_t k4_pge_dump(const PgeView &p, const GraphView &g, u32 pde, double *pg, double *plg, unsigned char *has,
                        cudaStream_t s);

}  // namespace gpe
