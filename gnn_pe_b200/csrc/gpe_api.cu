// gpe_api.cu -- the C ABI (include/gpe.h) and the host orchestration of the three kernel groups.
// Product code: no CPU fallback, nothing from oracle/.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <numeric>
#include <thread>

#include "gpe_internal.h"
#include "gpe_nccl.h"
#include "host_ref.h"

using namespace gpe;

static std::string g_create_err;

namespace {

enum Stage { kStageSelect = 0, kStageScan = 1, kStageCompact = 2, kStageJoin = 3, kStageEnumerate = 4, kNumStages_ = 5 };

struct StageTimer {
    gpe_ctx *c;
    float *dst;
    int span = -1;
    StageTimer(gpe_ctx *ctx, float *d, int stage) : c(ctx), dst(d) {
        if (c->timing == 1) cudaEventRecord(c->ev0, c->stream);
        if (c->timing == 2 && c->spans_used < (1u << 16)) {
            if (c->spans_used == c->spans.size()) {
                gpe_ctx::Span sp;
                cudaEventCreate(&sp.a);
                cudaEventCreate(&sp.b);
                c->spans.push_back(sp);
            }
            span = (int)c->spans_used++;
            c->spans[span].stage = stage;
            cudaEventRecord(c->spans[span].a, c->stream);
        }
    }
    ~StageTimer() {
        if (c->timing == 1) {
            cudaEventRecord(c->ev1, c->stream);
            cudaEventSynchronize(c->ev1);
            float ms = 0;
            cudaEventElapsedTime(&ms, c->ev0, c->ev1);
            *dst = ms;
        }
        if (span >= 0) cudaEventRecord(c->spans[span].b, c->stream);
    }
};

GraphView graph_view(const gpe_ctx *c) {
    GraphView g;
    g.V = c->V;
    g.off = c->d_off.as<u32>();
    g.nbr = c->d_nbr.as<u32>();
    g.label = c->d_label.as<u32>();
    g.deg = c->d_deg.as<u32>();
    g.rank = c->d_rank.as<u32>();
    g.ranklab = c->d_ranklab.as<u64>();
    g.vde = c->d_vde.as<double>();
    g.e = c->e;
    g.lpos = c->d_lpos.as<u32>();
    g.lclass = c->d_lclass.as<u32>();
    g.lcoff = c->d_lcoff.as<u32>();
    g.nbrG = c->d_nbrG.as<u32>();
    return g;
}

u32 host_key(const TableView &t, const u32 *labels) {
    u32 key = 0;
    for (u32 k = 0; k < t.L; k++)
        if (t.key_radix[k] > 1) key += (labels[k] % t.key_radix[k]) * t.key_stride[k];
    return key;
}

// Host-side staging of the plan paths of the current batch.
struct QPathSet {
    u32 L = 0, D = 0;
    std::vector<u32> slots, labels, degs;  // n x L
    std::vector<double> pde;               // n x D
    std::vector<u32> sid;                  // survivor counter of every path (empty: its own index)
    u32 n() const { return L ? (u32)(labels.size() / L) : 0; }
    // GPE_FILTER_BOTH_ORIENTATIONS: every plan path once more with its positions reversed -- comparing the reversed plan
    // path with a stored row is comparing the plan path with the row's other orientation, which the reference never
    // stored (custom.h:68-79) and never compares (:407-435).  Survivors of both count for the original path.
    void add_reversed() {
        const u32 n0 = n(), E = D / L;
        sid.resize(n0);
        for (u32 i = 0; i < n0; i++) sid[i] = i;
        for (u32 i = 0; i < n0; i++) {
            for (u32 k = 0; k < L; k++) {
                const size_t src = (size_t)i * L + (L - 1 - k);
                slots.push_back(slots[src]);
                labels.push_back(labels[src]);
                degs.push_back(degs[src]);
            }
            for (u32 k = 0; k < L; k++)
                for (u32 x = 0; x < E; x++) pde.push_back(pde[(size_t)i * D + (size_t)(L - 1 - k) * E + x]);
            sid.push_back(i);
        }
    }
};

// The host half of setting a batch up for the scan: plan paths grouped into query-path blocks by table bucket, the blocks'
// records packed, their tile ranges from the bucket directory, the label of every slot.  Reads the context (table
// directory), changes nothing: gpe_query_batches stages batch i+1 while the GPU works on batch i.
struct FilterStage {
    u32 n = 0, nb = 0, n_slots = 0, flags = 0;
    std::vector<unsigned char> recs;
    std::vector<u32> t0, slot_label;
    std::vector<u64> prefix;
};

void stage_filter(const gpe_ctx *c, const QPathSet &qp, u32 n_slots, u32 flags, FilterStage &st) {
    const TableView &t = c->tv;
    const u32 L = t.L, D = t.D, n = qp.n();
    const bool prune = !(flags & GPE_FILTER_NO_PRUNE);
    st.n = n;
    st.n_slots = n_slots;
    st.flags = flags;
    // group plan paths into blocks of <= kQB that read the same tiles
    std::vector<u32> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::vector<u32> keys(n, 0);
    if (prune) {
        for (u32 i = 0; i < n; i++) keys[i] = host_key(t, &qp.labels[(size_t)i * L]);
        std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return keys[a] < keys[b]; });
    }
    const size_t rec_bytes = qblock_rec_bytes(L, t.E);
    std::vector<QBlockHost> blocks;
    st.recs.clear();
    st.recs.reserve(((size_t)n / kQB + 64) * rec_bytes);
    std::vector<u32> ids(kQB), bl(kQB * kMaxL), bd(kQB * kMaxL), bs(kQB * kMaxL);
    std::vector<double> bp((size_t)kQB * kMaxL * kMaxE);
    for (u32 i = 0; i < n;) {
        u32 j = i;
        while (j < n && j - i < (u32)kQB && (!prune || keys[order[j]] == keys[order[i]])) j++;
        QBlockHost hb;
        hb.n = j - i;
        hb.first_qpath = qp.sid.empty() ? order[i] : qp.sid[order[i]];
        if (prune) {
            u64 r0 = c->h_bucket_start[keys[order[i]]], r1 = c->h_bucket_start[keys[order[i]] + 1];
            hb.t0 = (u32)(r0 / kTileRows);
            hb.t1 = r1 > r0 ? (u32)((r1 + kTileRows - 1) / kTileRows) : hb.t0;
        } else {
            hb.t0 = 0;
            hb.t1 = (u32)t.n_tiles;
        }
        for (u32 m = 0; m < hb.n; m++) {
            u32 q = order[i + m];
            ids[m] = qp.sid.empty() ? q : qp.sid[q];
            for (u32 k = 0; k < L; k++) {
                bl[m * L + k] = qp.labels[(size_t)q * L + k];
                bd[m * L + k] = qp.degs[(size_t)q * L + k];
                bs[m * L + k] = qp.slots[(size_t)q * L + k];
            }
            for (u32 d = 0; d < D; d++) bp[(size_t)m * D + d] = qp.pde[(size_t)q * D + d];
        }
        st.recs.resize(st.recs.size() + rec_bytes);
        qblock_pack(L, t.E, st.recs.data() + st.recs.size() - rec_bytes, hb.n, hb.first_qpath, ids.data(), bl.data(),
                    bd.data(), bs.data(), bp.data());
        blocks.push_back(hb);
        i = j;
    }
    const u32 nb = (u32)blocks.size();
    st.nb = nb;
    st.t0.assign(nb + 1, 0);
    st.prefix.assign(nb + 1, 0);
    for (u32 b = 0; b < nb; b++) {
        st.t0[b] = blocks[b].t0;
        st.prefix[b + 1] = st.prefix[b] + (blocks[b].t1 - blocks[b].t0);
    }
    // label of every slot, from the plan paths that cover it (an uncovered slot stays empty)
    st.slot_label.assign(std::max<u32>(n_slots, 1), 0xffffffffu);
    for (u32 i = 0; i < n; i++)
        for (u32 k = 0; k < L; k++) st.slot_label[qp.slots[(size_t)i * L + k]] = qp.labels[(size_t)i * L + k];
}

// The device half: buffers, the copies (through pinned memory, asynchronous), the batch state of the context.
int commit_filter(gpe_ctx *c, const FilterStage &st) {
    const u32 n = st.n, nb = st.nb, n_slots = st.n_slots;
    c->b_flags = st.flags;
    c->b_slots = n_slots;
    c->b_qpaths = n;
    c->b_qblocks = nb;
    c->b_items_unpruned = st.prefix[nb];
    // candidate bitmaps are local to the slot's label class: bit i of slot s = the i-th vertex (by id) of label(s).
    // 20x smaller than one bit per data vertex at 20 labels, so they stay in L2 while the table streams through
    c->b_words = ((u64)(c->max_class + 31) / 32 + kChunkWords - 1) / kChunkWords * kChunkWords;
    if (c->b_words == 0) c->b_words = kChunkWords;
    c->b_chunks_per_slot = c->b_words / kChunkWords;
    GPE_CUDA(c, c->d_slot_label.reserve(st.slot_label.size() * sizeof(u32)));
    GPE_CUDA(c, c->d_qblocks.reserve(std::max<size_t>(st.recs.size(), 16)));
    GPE_CUDA(c, c->d_qb_t0.reserve((nb + 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_qb_prefix.reserve((nb + 1) * sizeof(u64)));
    GPE_CUDA(c, c->d_worklist.reserve(std::max<u64>(st.prefix[nb], 1) * sizeof(u64)));
    GPE_CUDA(c, c->d_counters.reserve(8 * sizeof(u64)));
    GPE_CUDA(c, c->d_survivors.reserve(std::max<u32>(n, 1) * sizeof(u64)));
    GPE_CUDA(c, c->d_bitmap.reserve(std::max<u64>((u64)n_slots * c->b_words, 1) * sizeof(u32)));
    // stage through pinned memory so the copies are asynchronous
    const size_t sl_bytes = st.slot_label.size() * sizeof(u32);
    size_t o_rec = 0, o_t0 = (st.recs.size() + 15) / 16 * 16, o_pf = (o_t0 + (nb + 1) * sizeof(u32) + 15) / 16 * 16;
    const size_t o_sl = (o_pf + (nb + 1) * sizeof(u64) + 15) / 16 * 16;
    if (c->h_pin.cap < o_sl + sl_bytes) {  // (growing frees the old block: wait for copies still reading it)
        GPE_CUDA(c, cudaStreamSynchronize(c->stream));
        GPE_CUDA(c, c->h_pin.reserve(o_sl + sl_bytes));
    }
    unsigned char *pin = c->h_pin.as<unsigned char>();
    memcpy(pin + o_sl, st.slot_label.data(), sl_bytes);
    GPE_CUDA(c, cudaMemcpyAsync(c->d_slot_label.p, pin + o_sl, sl_bytes, cudaMemcpyHostToDevice, c->stream));
    memcpy(pin + o_rec, st.recs.data(), st.recs.size());
    memcpy(pin + o_t0, st.t0.data(), (nb + 1) * sizeof(u32));
    memcpy(pin + o_pf, st.prefix.data(), (nb + 1) * sizeof(u64));
    if (!st.recs.empty())
        GPE_CUDA(c, cudaMemcpyAsync(c->d_qblocks.p, pin + o_rec, st.recs.size(), cudaMemcpyHostToDevice, c->stream));
    GPE_CUDA(c, cudaMemcpyAsync(c->d_qb_t0.p, pin + o_t0, (nb + 1) * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
    GPE_CUDA(c, cudaMemcpyAsync(c->d_qb_prefix.p, pin + o_pf, (nb + 1) * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    c->stats.h2d_bytes += st.recs.size() + (nb + 1) * (sizeof(u32) + sizeof(u64)) + sl_bytes;
    GPE_CUDA(c, cudaEventRecord(c->ev_upload, c->stream));  // the staging block may be overwritten once this has passed
    c->b_scanned = false;
    c->b_filtered = false;
    c->b_joined = false;
    c->stats.n_qpaths = n;
    c->stats.n_qblocks = nb;
    c->stats.n_slots = n_slots;
    c->stats.scan_items_unpruned = st.prefix[nb];
    return GPE_OK;
}

int setup_filter(gpe_ctx *c, const QPathSet &qp, u32 n_slots, u32 flags) {
    FilterStage st;
    stage_filter(c, qp, n_slots, flags, st);
    return commit_filter(c, st);
}

// bitmap -> sorted candidate lists (d_cand, d_cand_off), entirely on the stream: NO host sync.  The total is only known on
// the device, so d_cand is sized from what earlier batches needed (or a bound); the expand kernel drops what does not fit
// and raises a flag that gpe_batch_download / cand_total() check once the stream has been waited for anyway -- the step
// is then redone with the right size (first batches only).  With d_all != null the bitmaps are first replaced by the
// union of `world` all-gathered shard bitmaps (same kernel as the popcount pass).
// pinned h_pin3: [0] candidate total, [1..2] scan counters, [3] overflow flag -- valid after the next stream sync.
int compact_candidates(gpe_ctx *c, const u32 *d_all = nullptr, u32 world = 0) {
    StageTimer tm(c, &c->stats.last_compact_ms, kStageCompact);
    const u64 n_chunks = c->b_chunks_per_slot * c->b_slots;
    GPE_CUDA(c, c->d_chunk_off.reserve((n_chunks + 1) * sizeof(u64)));
    GPE_CUDA(c, c->d_cand_off.reserve(((u64)c->b_slots + 1) * sizeof(u64)));
    GPE_CUDA(c, c->d_counters.reserve(8 * sizeof(u64)));
    if (d_all)
        GPE_CUDA(c, k3_merge_count(d_all, (u64)c->b_slots * c->b_words, world, c->d_bitmap.as<u32>(), c->b_words,
                                   c->b_chunks_per_slot, c->b_slots, c->d_chunk_off.as<u64>(), c->stream));
    else
        GPE_CUDA(c, k3_chunk_count(c->d_bitmap.as<u32>(), c->b_words, c->b_chunks_per_slot, c->b_slots,
                                   c->d_chunk_off.as<u64>(), c->stream));
    GPE_CUDA(c, exclusive_scan_u64(c->d_chunk_off.as<u64>(), n_chunks + 1, c->d_scan_tmp, c->stream));
    // capacity: twice the largest total seen so far, at least 4 M entries, never more than every slot's whole class
    const u64 bound = std::max<u64>((u64)c->b_slots * c->max_class, 1);
    const u64 want = std::min<u64>(bound, std::max<u64>(2 * c->cand_seen_max, 4ull << 20));
    if (c->d_cand.cap / sizeof(u32) < want) GPE_CUDA(c, c->d_cand.reserve(want * sizeof(u32)));
    c->b_cand_cap = c->d_cand.cap / sizeof(u32);
    if (const char *e = getenv("GPE_CAND_CAP")) c->b_cand_cap = std::min<u64>(c->b_cand_cap, strtoull(e, nullptr, 10));  // tests: force the repair path
    u64 *flag = c->d_counters.as<u64>() + 5;
    GPE_CUDA(c, cudaMemsetAsync(flag, 0, sizeof(u64), c->stream));
    if (n_chunks == 0) {
        GPE_CUDA(c, cudaMemsetAsync(c->d_cand_off.p, 0, sizeof(u64), c->stream));
    } else {
        GPE_CUDA(c, k3_compact(c->d_bitmap.as<u32>(), c->b_words, c->b_chunks_per_slot, c->b_slots,
                               c->d_chunk_off.as<u64>(), c->d_slot_label.as<u32>(), c->d_lcoff.as<u32>(), c->n_labels,
                               c->d_cand.as<u32>(), c->d_cand_off.as<u64>(), c->b_cand_cap, flag, c->stream));
    }
    GPE_CUDA(c, c->h_pin3.reserve(8 * sizeof(u64)));
    u64 *pin = c->h_pin3.as<u64>();
    GPE_CUDA(c, cudaMemcpyAsync(pin, c->d_chunk_off.as<u64>() + n_chunks, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    GPE_CUDA(c, cudaMemcpyAsync(pin + 1, c->d_counters.p, 2 * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    GPE_CUDA(c, cudaMemcpyAsync(pin + 3, flag, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    c->b_cand_known = false;
    c->b_cand_external_lists = false;
    c->stats.compact_launches += 3;
    c->stats.kernel_launches += 1 + exclusive_scan_launches(n_chunks + 1) + (n_chunks ? 1 : 0);
    c->stats.d2h_bytes += 4 * sizeof(u64);
    return GPE_OK;
}

// Wait for the stream and take over what the compaction left in pinned memory.  Returns 1 when the candidate lists did
// not fit d_cand (the caller redoes compaction + join; the capacity has been raised), 0 otherwise, < 0 never.
int cand_total(gpe_ctx *c, bool *overflow) {
    if (overflow) *overflow = false;
    if (c->b_cand_known) return GPE_OK;
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));
    const u64 *pin = c->h_pin3.as<u64>();
    c->b_n_cand = pin[0];
    c->cand_seen_max = std::max(c->cand_seen_max, pin[0]);
    c->stats.n_candidates = pin[0];
    c->stats.scan_items = pin[1];
    c->stats.scan_rows = pin[1] * kTileRows;
    if (c->b_pge_rows_pending) {  // GNN-PGE scan: rows = data vertices of the label classes some query vertex asked for
        c->stats.scan_rows = pin[4];
        c->stats.scan_items = (pin[4] + 255) / 256;
        c->b_pge_rows_pending = false;
    }
    c->b_cand_known = true;
    if (pin[3] && overflow) *overflow = true;
    return GPE_OK;
}

int run_join(gpe_ctx *c, u32 rank, u32 world, u32 *d_matches, u64 matches_cap, bool force_dfs = false,
             const u32 *d_qmode = nullptr);

// the lists did not fit: expand them again into a buffer of the right size (chunk offsets are still on the device)
int recompact(gpe_ctx *c) {
    // (sized as the next batches will ask for -- twice the largest total seen -- so that no later step reallocates)
    const u64 bound = std::max<u64>((u64)c->b_slots * c->max_class, 1);
    GPE_CUDA(c, c->d_cand.reserve(std::min<u64>(bound, std::max<u64>(2 * c->cand_seen_max, 4ull << 20)) * sizeof(u32)));
    c->b_cand_cap = c->d_cand.cap / sizeof(u32);
    u64 *flag = c->d_counters.as<u64>() + 5;
    GPE_CUDA(c, cudaMemsetAsync(flag, 0, sizeof(u64), c->stream));
    GPE_CUDA(c, k3_compact(c->d_bitmap.as<u32>(), c->b_words, c->b_chunks_per_slot, c->b_slots, c->d_chunk_off.as<u64>(),
                           c->d_slot_label.as<u32>(), c->d_lcoff.as<u32>(), c->n_labels, c->d_cand.as<u32>(),
                           c->d_cand_off.as<u64>(), c->b_cand_cap, flag, c->stream));
    c->stats.kernel_launches++;
    return GPE_OK;
}

// Candidate total known on the host and the lists complete; a join that already ran on truncated lists (it found the
// overflow flag and did nothing) is run again.
int settle_candidates(gpe_ctx *c) {
    bool overflow = false;
    int rc = cand_total(c, &overflow);
    if (rc || !overflow) return rc;
    if ((rc = recompact(c))) return rc;
    if (c->b_joined) rc = run_join(c, c->b_rank, c->b_world, nullptr, 0, false, nullptr);
    return rc;
}

// select + scan: the candidate bitmaps of this GPU's table (shard) are on the device afterwards
int run_scan(gpe_ctx *c) {
    const u64 n_items = c->b_items_unpruned;
    {
        ZeroList z;
        z.add(c->d_counters.p, 8 * sizeof(u64));
        z.add(c->d_survivors.p, std::max<u32>(c->b_qpaths, 1) * sizeof(u64));
        GPE_CUDA(c, k0_zero(z, c->stream));
    }
    GPE_CUDA(c, cudaMemsetAsync(c->d_bitmap.p, 0, std::max<u64>((u64)c->b_slots * c->b_words, 1) * sizeof(u32), c->stream));
    if (n_items > 0 && c->b_qblocks > 0) {
        {
            StageTimer tm(c, &c->stats.last_select_ms, kStageSelect);
            GPE_CUDA(c, k2_select(c->tv, c->d_qblocks.p, c->d_qb_t0.as<u32>(), c->d_qb_prefix.as<u64>(), c->b_qblocks,
                                  n_items, !(c->b_flags & GPE_FILTER_NO_PRUNE), c->d_worklist.as<u64>(),
                                  c->d_counters.as<u64>(), c->stream));
            c->stats.select_launches++;
            c->stats.kernel_launches++;
        }
        {
            StageTimer tm(c, &c->stats.last_scan_ms, kStageScan);
            if (c->tv.ids_only)
                GPE_CUDA(c, k2_scan_ids(c->tv, c->d_vrec.p, c->d_qblocks.p, c->d_worklist.as<u64>(), c->d_counters.as<u64>(),
                                        c->d_bitmap.as<u32>(), c->b_words, c->d_survivors.as<u64>(), c->sm_count, c->stream));
            else
            GPE_CUDA(c, k2_scan(c->tv, c->d_qblocks.p, c->d_worklist.as<u64>(), c->d_counters.as<u64>(),
                                c->d_bitmap.as<u32>(), c->b_words, c->d_survivors.as<u64>(),
                                !(c->b_flags & GPE_FILTER_NO_PRUNE), c->sm_count, c->stream));
            c->stats.scan_launches++;
            c->stats.kernel_launches++;
        }
    }
    c->b_scanned = true;
    c->b_filtered = false;
    return GPE_OK;
}

int run_filter(gpe_ctx *c) {
    c->b_sparse_used = false;
    int rc = run_scan(c);
    if (rc) return rc;
    if ((rc = compact_candidates(c))) return rc;
    c->b_filtered = true;
    c->b_cand_external = false;
    c->b_cand_clean = true;  // every candidate of a slot carries the slot's label, and the slot's bitmap is on the device
    return GPE_OK;
}

int upload_queries(gpe_ctx *c, u32 n_queries, const u32 *q_vbase, const u32 *q_ebase, const u32 *q_offsets,
                   const u32 *q_nbrs, const u32 *q_labels, const u64 *limits) {
    const u32 n_slots = q_vbase[n_queries], n_adj = q_ebase[n_queries];
    c->b_nq = n_queries;
    c->b_max_nq = 1;
    for (u32 q = 0; q < n_queries; q++) c->b_max_nq = std::max(c->b_max_nq, q_vbase[q + 1] - q_vbase[q]);
    c->stats.h2d_bytes = 0;
    c->stats.d2h_bytes = 0;
    c->h_q_vbase.assign(q_vbase, q_vbase + n_queries + 1);
    c->h_limits.assign(n_queries, GPE_LIMIT_MAX);
    if (limits) c->h_limits.assign(limits, limits + n_queries);
    size_t sz[6] = {(n_queries + 1) * sizeof(u32), (n_queries + 1) * sizeof(u32), ((size_t)n_slots + n_queries) * sizeof(u32),
                    std::max<size_t>(n_adj, 1) * sizeof(u32), std::max<size_t>(n_slots, 1) * sizeof(u32),
                    n_queries * sizeof(u64)};
    GPE_CUDA(c, c->d_q_vbase.reserve(sz[0]));
    GPE_CUDA(c, c->d_q_ebase.reserve(sz[1]));
    GPE_CUDA(c, c->d_q_offsets.reserve(sz[2]));
    GPE_CUDA(c, c->d_q_nbrs.reserve(sz[3]));
    GPE_CUDA(c, c->d_q_labels.reserve(sz[4]));
    GPE_CUDA(c, c->d_limits.reserve(std::max<size_t>(sz[5], 8)));
    {   // staged through pinned memory: the copies are asynchronous and the caller's buffers are free on return
        size_t off[7];
        const size_t bytes[6] = {sz[0], sz[1], sz[2], (size_t)n_adj * sizeof(u32), (size_t)n_slots * sizeof(u32), sz[5]};
        off[0] = 0;
        for (int i = 0; i < 6; i++) off[i + 1] = (off[i] + bytes[i] + 15) / 16 * 16;
        GPE_CUDA(c, cudaEventSynchronize(c->ev_upload));  // the previous batch's copies out of the staging blocks are done
        if (c->h_pin_q.cap < off[6]) {
            GPE_CUDA(c, cudaStreamSynchronize(c->stream));
            GPE_CUDA(c, c->h_pin_q.reserve(off[6]));
        }
        unsigned char *pin = c->h_pin_q.as<unsigned char>();
        const void *src[6] = {q_vbase, q_ebase, q_offsets, q_nbrs, q_labels, c->h_limits.data()};
        void *dst[6] = {c->d_q_vbase.p, c->d_q_ebase.p, c->d_q_offsets.p, c->d_q_nbrs.p, c->d_q_labels.p, c->d_limits.p};
        for (int i = 0; i < 6; i++) {
            if (!bytes[i]) continue;
            memcpy(pin + off[i], src[i], bytes[i]);
            GPE_CUDA(c, cudaMemcpyAsync(dst[i], pin + off[i], bytes[i], cudaMemcpyHostToDevice, c->stream));
        }
        GPE_CUDA(c, cudaEventRecord(c->ev_upload, c->stream));
    }
    GPE_CUDA(c, c->d_order.reserve(std::max<size_t>(n_slots, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_pivot.reserve(std::max<size_t>(n_slots, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_jplan.reserve(std::max<size_t>(n_slots, 1) * sizeof(JoinDepth)));
    GPE_CUDA(c, c->d_kids.reserve(std::max<size_t>(n_slots, 1) * 2 * sizeof(u32)));
    GPE_CUDA(c, c->d_item_base.reserve(((size_t)n_queries + 1) * sizeof(u64)));
    GPE_CUDA(c, c->d_answers.reserve((2 * (size_t)n_queries + 8) * sizeof(u64)));
    GPE_CUDA(c, c->d_match_cursor.reserve(2 * sizeof(u64)));
    c->stats.h2d_bytes += sz[0] + sz[1] + sz[2] + (size_t)n_adj * sizeof(u32) + (size_t)n_slots * sizeof(u32) + sz[5];
    return GPE_OK;
}

constexpr u64 kJoinExportBytes = 256ull << 20;  // room for exported work items

// d_answers: [0, nq + 8) raw counts, [nq + 8, 2 nq + 8) per-query "inexact" flags (see kSat in k3_join.cu)
int run_join(gpe_ctx *c, u32 rank, u32 world, u32 *d_matches, u64 matches_cap, bool force_dfs, const u32 *d_qmode) {
    StageTimer tm(c, &c->stats.last_join_ms, kStageJoin);
    c->b_rank = rank;
    c->b_world = world;
    const u32 nq = c->b_nq;
    u64 *answers = c->d_answers.as<u64>();
    ZeroList zl;  // everything the join wants zeroed, one launch (filled below)
    zl.add(c->d_answers.p, (2 * (size_t)nq + 8) * sizeof(u64));
    zl.add(c->d_match_cursor.p, 2 * sizeof(u64));
    GPE_CUDA(c, c->d_jq.reserve(sizeof(JoinQueue)));
    // Subtree tables are only sound when the start vertex's candidates carry its label; that holds for the
    // filter's own candidate sets (their bitmaps are on the device), not for caller-supplied ones (gpe_refine).
    const bool enumerate = d_matches != nullptr;
    const bool clean_start = c->b_cand_clean && !enumerate;
    const u32 n_slots = c->b_slots;
    // (two table jobs per slot at most: N_v of a vertex with peeled children, S_u of a weighted counted leaf)
    GPE_CUDA(c, c->d_tjobs.reserve(2 * std::max<size_t>(n_slots, 1) * sizeof(TreeJob)));
    GPE_CUDA(c, c->d_tchild.reserve(2 * std::max<size_t>(n_slots, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_tcursor.reserve(4 * sizeof(u64)));
    // table pool: room for a table per slot and a sum table per slot at most, capped at 1 Gi entries (8 GB) and a quarter
    // of the free memory -- a query whose tables do not fit walks instead (k3_order reserves per query)
    u64 pool_cap = 2 * std::max<u64>((u64)n_slots * c->max_class, 1);
    {
        const u64 have = c->d_tpool.cap / sizeof(u64), hard_cap = 1ull << 30;
        if (have < pool_cap && have < hard_cap) {  // batches are still getting bigger: grow, but never thrash -- the target
            size_t free_b = 0, total_b = 0;        // depends on the free memory of the moment, so only a clear gain counts
            cudaMemGetInfo(&free_b, &total_b);
            const u64 target = std::min<u64>(pool_cap, std::min<u64>(hard_cap, (free_b + c->d_tpool.cap) / 4 / sizeof(u64)));
            if (target > have + have / 4) GPE_CUDA(c, c->d_tpool.reserve(target * sizeof(u64)));
        }
        pool_cap = std::min<u64>(pool_cap, c->d_tpool.cap / sizeof(u64));
        if (const char *e = getenv("GPE_JOIN_POOL")) pool_cap = std::min<u64>(pool_cap, std::max<u64>(strtoull(e, nullptr, 10), 1));  // tests
    }
    zl.add(c->d_tcursor.p, 4 * sizeof(u64));
    GPE_CUDA(c, c->d_tlist.reserve(((size_t)kMaxTreeLevels * 2 * std::max<u32>(n_slots, 1) + kMaxTreeLevels) * sizeof(u32)));
    bool allow_weighted = true;  // counted leaves may carry peeled subtrees
    if (const char *e = getenv("GPE_JOIN_WEIGHTED")) allow_weighted = atoi(e) != 0;
    (void)force_dfs;
    u32 *tcount = c->d_tlist.as<u32>(), *tlist = tcount + kMaxTreeLevels;
    zl.add(tcount, kMaxTreeLevels * sizeof(u32));
    GPE_CUDA(c, c->d_qcur.reserve(((size_t)nq + 1) * 6 * sizeof(u64)));
    zl.add(c->d_qcur.p, ((size_t)nq + 1) * 6 * sizeof(u64));
    GPE_CUDA(c, k0_zero(zl, c->stream));
    GPE_CUDA(c, k3_order(nq, c->V, c->d_q_vbase.as<u32>(), c->d_q_ebase.as<u32>(), c->d_q_offsets.as<u32>(),
                         c->d_q_nbrs.as<u32>(), c->d_q_labels.as<u32>(), c->d_cand_off.as<u64>(), c->d_order.as<u32>(),
                         c->d_pivot.as<u32>(), c->d_jplan.as<JoinDepth>(), c->d_kids.p, c->d_item_base.as<u64>(), rank, world,
                         enumerate, clean_start, c->n_labels, c->d_lcoff.as<u32>(), c->d_tjobs.as<TreeJob>(),
                         c->d_tchild.as<u32>(), c->d_tcursor.as<u64>(), tcount, tlist, n_slots, allow_weighted, d_qmode,
                         (float)(getenv("GPE_JOIN_PEEL") ? (atoi(getenv("GPE_JOIN_PEEL")) ? 1e9 : 0.0) : c->branching), pool_cap,
                         c->stream));
    // (the kernels that walk read tpool[tree_off + v'] with tree_off = table offset + V - lcoff[label]: pointer shifted by -V;
    //  k3_tree_tables writes through the unshifted pointer it is given separately)
    JoinGraph jv{c->V, c->n_labels, c->d_nbrJ.as<u32>(), c->d_gtab.as<unsigned char>(), c->dir_row_bytes, c->wide_adj,
                 c->wide_dir, c->d_degJ.as<u32>(), c->d_labelJ.as<u32>(), c->d_lcoff.as<u32>(), c->d_lclass.as<u32>(),
                 c->d_tpool.as<u64>() - c->V, c->d_bloom.as<u64>(), (c->bloom_bits >> 6) - 1};
    u32 tree_launches = 0;
    if (!enumerate && c->b_max_nq >= 2) {
        tree_launches = c->b_max_nq - 1;
        GPE_CUDA(c, k3_tree_tables(jv, n_slots, c->max_class, tree_launches, c->d_tjobs.as<TreeJob>(), c->d_tchild.as<u32>(),
                                   tcount, tlist, clean_start ? c->d_bitmap.as<u32>() : nullptr, c->b_words,
                                   c->d_tpool.as<u64>(), c->sm_count, c->stream));
    }
    // tickets [0, n_init) are the root candidates of this shard (the candidate capacity + one label class per query bounds them);
    // later tickets are subtrees exported by busy threads
    const u32 stride = k3_item_stride(c->b_max_nq);
    const u64 cap = kJoinExportBytes / (stride * sizeof(u32));
    GPE_CUDA(c, c->d_init.reserve((std::max<u64>(c->b_cand_cap, 1) + (u64)nq * c->max_class) * 2 * sizeof(u32)));
    GPE_CUDA(c, c->d_items.reserve(cap * stride * sizeof(u32)));
    if (c->d_ready.cap < cap * sizeof(u32) || c->join_epoch == 0xffffffffu) {
        GPE_CUDA(c, c->d_ready.reserve(cap * sizeof(u32)));
        GPE_CUDA(c, cudaMemsetAsync(c->d_ready.p, 0, c->d_ready.cap, c->stream));
        c->join_epoch = 0;
    }
    const u32 epoch = ++c->join_epoch;
    JoinQueue *jq = c->d_jq.as<JoinQueue>();
    const u32 heavy_deg = std::max<u32>(32, c->V ? (u32)(4ull * c->n_adj / c->V) : 32);  // 4 x the mean degree
    GPE_CUDA(c, k3_init_items(jv, nq, c->d_q_vbase.as<u32>(), c->d_jplan.as<JoinDepth>(), c->d_cand_off.as<u64>(),
                              c->d_cand.as<u32>(), c->d_item_base.as<u64>(), rank, world, heavy_deg,
                              c->d_qcur.as<u64>(), c->d_init.p, jq, tree_launches > 0,
                              c->b_cand_external_lists ? nullptr : c->d_counters.as<u64>() + 5, c->sm_count, c->stream));
    GPE_CUDA(c, k3_dfs(jv, c->b_max_nq, c->d_q_vbase.as<u32>(), c->d_jplan.as<JoinDepth>(), c->d_kids.p, c->d_cand.as<u32>(),
                       c->d_init.p, c->d_limits.as<u64>(), answers, c->d_items.as<u32>(), cap, c->d_ready.as<u32>(), epoch,
                       jq, d_matches, matches_cap, c->d_match_cursor.as<u64>(), answers + nq + 8, c->sm_count, c->stream));
    c->stats.kernel_launches += tree_launches ? 1 : 0;  // all table levels in one cooperative launch
    c->stats.join_launches += tree_launches ? 1 : 0;
    c->stats.kernel_launches += 5;
    c->stats.join_launches += 5;
    c->b_joined = true;
    return GPE_OK;
}

// queue counters of the last join (items produced, DFS steps); needs the stream to be idle
int read_join_stats(gpe_ctx *c) {
    JoinQueue h;
    GPE_CUDA(c, cudaMemcpyAsync(&h, c->d_jq.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stats.join_items = h.tail;
    c->stats.join_exports = h.exports;
    c->stats.join_donations = h.donations;
    c->stats.join_steps = h.steps;
    c->stats.join_warp_iters = h.warp_iters;
    c->stats.join_idle_polls = h.idle_polls;
    return GPE_OK;
}

// the device's candidate lists hold class-order ids (ascending inside a slot, like the caller's ids: a slot's
// candidates share a label and class order is by id inside a label); the host gets the caller's ids
int download_candidates(gpe_ctx *c, u32 *cand) {
    if (!c->b_n_cand) return GPE_OK;
    DevBuf tmp;
    GPE_CUDA(c, tmp.reserve(c->b_n_cand * sizeof(u32)));
    cudaError_t e = k0_gather(c->b_n_cand, c->d_lclass.as<u32>(), c->d_cand.as<u32>(), tmp.as<u32>(), c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(cand, tmp.p, c->b_n_cand * sizeof(u32), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    tmp.release();
    GPE_CUDA(c, e);
    return GPE_OK;
}

bool check_query(gpe_ctx *c, u32 nq, const u32 *off, const u32 *nbr, std::string &why) {
    if (nq > GPE_MAX_QUERY_VERTICES) { why = "query has more than GPE_MAX_QUERY_VERTICES vertices"; return false; }
    if (nq == 0) { why = "empty query graph"; return false; }
    if (!query_csr_ok(nq, off, nbr, why)) return false;
    if (!query_connected(nq, off, nbr)) {
        // custom.h:684-704 reads an uninitialised `next_vertex` for a disconnected query
        why = "disconnected query graph (undefined behaviour in the reference)";
        return false;
    }
    (void)c;
    return true;
}

}  // namespace

// ================================================================================================================
extern "C" {

int gpe_abi_version(void) { return 1; }

const char *gpe_last_error(const gpe_ctx *ctx) {
    if (ctx) return ctx->err.c_str();
    return g_create_err.c_str();
}

int gpe_create(int device, gpe_ctx **out) {
    if (!out) return GPE_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_err = std::string("no CUDA device available (libgpe has no CPU fallback): ") +
                       (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return GPE_ERR_CUDA;
    }
    if (device < 0 || device >= n) { g_create_err = "device index out of range"; return GPE_ERR_INVALID; }
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_create_err = cudaGetErrorString(e);
        return GPE_ERR_CUDA;
    }
    if (prop.major != 10) {
        g_create_err = std::string("libgpe is built for sm_100a (B200) only; device is ") + prop.name;
        return GPE_ERR_UNSUPPORTED;
    }
    gpe_ctx *c = new gpe_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->ev_upload, cudaEventDisableTiming)) != cudaSuccess) {
        g_create_err = cudaGetErrorString(e);
        delete c;
        return GPE_ERR_CUDA;
    }
    *out = c;
    return GPE_OK;
}

void gpe_destroy(gpe_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->comm) nccl_api().CommDestroy((ncclComm_t)c->comm);
    c->d_all_bitmaps.release();
    c->d_reduce.release();
    c->d_sparse.release();
    DevBuf *bufs[] = {&c->d_off, &c->d_nbr, &c->d_label, &c->d_deg, &c->d_nbrJ, &c->d_gtab, &c->d_offJ, &c->d_degJ, &c->d_labelJ, &c->d_newid, &c->d_nbrG, &c->d_lclass, &c->d_lpos, &c->d_lcoff, &c->d_bloom, &c->d_tjobs, &c->d_tchild, &c->d_tpool, &c->d_tcursor, &c->d_tlist, &c->d_qcur, &c->d_items, &c->d_ready, &c->d_jq, &c->d_init, &c->d_kids, &c->d_rank, &c->d_sorted, &c->d_member, &c->d_vde, &c->d_vrec, &c->d_ranklab,
                      &c->d_offr, &c->d_ebase, &c->d_start_rows, &c->d_scan_tmp, &c->d_tiles, &c->d_vids, &c->d_sum_u32, &c->d_sum_f64,
                      &c->d_bucket, &c->d_cursor, &c->d_qblocks, &c->d_qb_t0, &c->d_qb_prefix, &c->d_worklist,
                      &c->d_counters, &c->d_bitmap, &c->d_survivors, &c->d_slot_label, &c->d_chunk_cnt, &c->d_chunk_off, &c->d_cand,
                      &c->d_cand_off, &c->d_q_vbase, &c->d_q_ebase, &c->d_q_offsets, &c->d_q_nbrs, &c->d_q_labels,
                      &c->d_limits, &c->d_order, &c->d_pivot, &c->d_jplan, &c->d_item_base, &c->d_answers,
                      &c->d_matches, &c->d_match_cursor, &c->d_qmode, &c->d_pge, &c->d_pge_x, &c->d_pge_q};
    for (DevBuf *b : bufs) b->release();
    c->h_pin.release();
    c->h_pin2.release();
    c->h_pin3.release();
    c->h_pin4.release();
    c->h_pin_q.release();
    if (c->ev_upload) cudaEventDestroy(c->ev_upload);
    for (auto &sp : c->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    delete c;
}

void *gpe_stream(gpe_ctx *c) { return c ? (void *)c->stream : nullptr; }

int gpe_sync(gpe_ctx *c) {
    if (!c) return GPE_ERR_INVALID;
    GPE_CUDA(c, cudaSetDevice(c->device));
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));
    return GPE_OK;
}

int gpe_set_timing(gpe_ctx *c, int mode) {
    if (!c || mode < 0 || mode > 2) return GPE_ERR_INVALID;
    c->timing = mode;
    return GPE_OK;
}

int gpe_collect_timings(gpe_ctx *c, double *sum_ms, uint64_t *count) {
    if (!c || !sum_ms || !count) return GPE_ERR_INVALID;
    GPE_CUDA(c, cudaSetDevice(c->device));
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < GPE_NUM_STAGES; i++) { sum_ms[i] = 0; count[i] = 0; }
    for (size_t i = 0; i < c->spans_used; i++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->spans[i].a, c->spans[i].b) == cudaSuccess) {
            sum_ms[c->spans[i].stage] += ms;
            count[c->spans[i].stage]++;
        }
    }
    c->spans_used = 0;
    return GPE_OK;
}

int gpe_get_stats(gpe_ctx *c, gpe_stats *out) {
    if (!c || !out) return GPE_ERR_INVALID;
    c->stats.table_rows = c->tv.n_rows;
    c->stats.table_tiles = c->tv.n_tiles;
    c->stats.tile_rows = kTileRows;
    c->stats.row_bytes = c->tv.L ? (u64)c->tv.L * 8 + (u64)c->tv.D * 8 : 0;
    c->stats.table_ids_only = c->tv.ids_only ? 1 : 0;
    c->stats.stored_row_bytes = c->tv.L ? (c->tv.ids_only ? (u64)c->tv.L * 4 : (u64)c->tv.L * 12 + (u64)c->tv.D * 8) : 0;
    *out = c->stats;
    return GPE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
int gpe_set_graph(gpe_ctx *c, uint32_t V, const uint32_t *offsets, const uint32_t *nbrs, const uint32_t *labels) try {
    if (!c || !offsets || !labels || (V && !nbrs && offsets[V])) return c ? c->fail(GPE_ERR_INVALID, "null argument") : GPE_ERR_INVALID;
    GPE_CUDA(c, cudaSetDevice(c->device));
    if (offsets[0] != 0) return c->fail(GPE_ERR_INVALID, "offsets[0] must be 0");
    const u32 n_adj = offsets[V];
    c->have_graph = c->have_emb = c->have_enum = c->have_table = c->have_pge = false;
    // the CSR goes to the device as it is; validation, degrees, label classes and the join's copy are built there
    GPE_CUDA(c, c->d_off.reserve(((size_t)V + 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_nbr.reserve(std::max<size_t>(n_adj, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_label.reserve(std::max<size_t>(V, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_deg.reserve(std::max<size_t>(V, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_counters.reserve(8 * sizeof(u64)));
    GPE_CUDA(c, cudaMemcpyAsync(c->d_off.p, offsets, ((size_t)V + 1) * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
    if (n_adj) GPE_CUDA(c, cudaMemcpyAsync(c->d_nbr.p, nbrs, (size_t)n_adj * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
    if (V) GPE_CUDA(c, cudaMemcpyAsync(c->d_label.p, labels, (size_t)V * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
    GPE_CUDA(c, k0_validate(V, n_adj, c->d_off.as<u32>(), c->d_nbr.as<u32>(), c->d_label.as<u32>(), c->d_deg.as<u32>(),
                            c->d_counters.as<u64>(), c->stream));
    u64 err3[4];
    GPE_CUDA(c, cudaMemcpyAsync(err3, c->d_counters.p, sizeof err3, cudaMemcpyDeviceToHost, c->stream));
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));
    if (err3[0] != ~0ull) {
        const u32 v = (u32)err3[0];
        switch (err3[0] >> 32) {
            case 1: return c->fail(GPE_ERR_INVALID, "offsets not monotone at vertex %u", v);
            case 2: return c->fail(GPE_ERR_INVALID, "neighbour id out of range at vertex %u", v);
            case 3: return c->fail(GPE_ERR_INVALID, "self loop at vertex %u (simple graphs only)", v);
            default: return c->fail(GPE_ERR_INVALID, "adjacency of vertex %u not strictly ascending (sorted, no duplicate edges)", v);
        }
    }
    const u32 max_label = (u32)err3[1], max_deg = (u32)err3[2];
    // label ids index dense per-label arrays (class offsets, the table's bucket directory, the group directory rows)
    if (max_label >= (1u << 22))
        return c->fail(GPE_ERR_UNSUPPORTED, "label id %u: label ids must be below 2^22 (dense per-label tables)", max_label);
    c->V = V;
    c->n_adj = n_adj;
    c->n_labels = V ? max_label + 1 : 0;
    c->max_degree = max_deg;
    const u32 nl = c->n_labels;
    // expected number of neighbours of ONE given label of a vertex reached over an edge: (sum d^2 / sum d) / labels.  Above ~1
    // a walk over unique-label query vertices grows (power-law graphs), below it dies out (the join's plan uses it)
    c->branching = n_adj && nl ? (double)err3[3] / (double)n_adj / (double)nl : 0.0;
    // group directory: one row per vertex with the start of every label group of its adjacency; 16-bit offsets from the
    // row's base while degrees allow it (46 bytes per vertex at 20 labels instead of 84)
    c->wide_adj = V > (1u << 24);
    c->wide_dir = max_deg >= 65536u;
    if (const char *e = getenv("GPE_JOIN_DIR")) c->wide_dir = c->wide_dir || !strcmp(e, "wide");  // measurement switches
    if (const char *e = getenv("GPE_JOIN_ADJ")) c->wide_adj = c->wide_adj || !strcmp(e, "wide");
    c->dir_row_bytes = c->wide_dir ? (nl + 1) * 4 : (4 + (nl + 1) * 2 + 3) / 4 * 4;
    const u64 dir_bytes = (u64)V * c->dir_row_bytes;
    {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (dir_bytes > free_b / 3)
            return c->fail(GPE_ERR_UNSUPPORTED, "group directory of %llu bytes (V x (labels+1) entries) does not fit; "
                                                "a sparse directory is not built yet", (unsigned long long)dir_bytes);
    }
    GPE_CUDA(c, c->d_lcoff.reserve(((size_t)nl + 2) * sizeof(u32)));
    GPE_CUDA(c, c->d_lclass.reserve(std::max<size_t>(V, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_lpos.reserve(std::max<size_t>(V, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_newid.reserve(std::max<size_t>(V, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_degJ.reserve(std::max<size_t>(V, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_labelJ.reserve(std::max<size_t>(V, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_offJ.reserve(((size_t)V + 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_nbrJ.reserve(std::max<size_t>(n_adj, 1) * (c->wide_adj ? 8 : 4)));
    GPE_CUDA(c, c->d_nbrG.reserve(std::max<size_t>(n_adj, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_gtab.reserve(std::max<u64>(dir_bytes, 16)));
    u64 bits = 1024;  // edge filter: 16 bits per undirected edge, two of them set
    while (bits < 8ull * n_adj) bits <<= 1;
    c->bloom_bits = bits;
    GPE_CUDA(c, c->d_bloom.reserve(bits / 8));
    DevBuf tmp;
    u32 *max_class_dev = reinterpret_cast<u32 *>(c->d_counters.as<u64>() + 4);
    cudaError_t e = k0_build_classes(V, nl, c->d_label.as<u32>(), c->d_deg.as<u32>(), c->d_lcoff.as<u32>(), c->d_lclass.as<u32>(),
                                     c->d_lpos.as<u32>(), c->d_newid.as<u32>(), c->d_degJ.as<u32>(), c->d_labelJ.as<u32>(),
                                     c->d_offJ.as<u32>(), max_class_dev, tmp, c->stream);
    if (e == cudaSuccess)
        e = k0_build_join_graph(V, n_adj, nl, c->d_off.as<u32>(), c->d_nbr.as<u32>(), c->d_lclass.as<u32>(), c->d_newid.as<u32>(),
                                c->d_degJ.as<u32>(), c->d_offJ.as<u32>(), c->d_lcoff.as<u32>(), c->wide_adj, c->wide_dir,
                                c->dir_row_bytes, c->d_nbrJ.as<u32>(), c->d_nbrG.as<u32>(), c->d_gtab.p, c->d_bloom.as<u64>(), bits,
                                tmp, c->sm_count, c->stream);
    u32 max_class = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&max_class, max_class_dev, sizeof(u32), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    tmp.release();
    GPE_CUDA(c, e);
    c->max_class = max_class;
    c->stats.kernel_launches += 12;
    c->have_graph = true;
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_set_graph", ex.what()) : GPE_ERR_INVALID;
}

int gpe_set_embeddings(gpe_ctx *c, uint32_t e, const double *vde) try {
    if (!c || !vde) return c ? c->fail(GPE_ERR_INVALID, "null argument") : GPE_ERR_INVALID;
    if (!c->have_graph) return c->fail(GPE_ERR_INVALID, "gpe_set_graph first");
    if (e == 0 || e > (u32)kMaxE) return c->fail(GPE_ERR_UNSUPPORTED, "embedding dimension %u not in 1..%d", e, kMaxE);
    GPE_CUDA(c, cudaSetDevice(c->device));
    c->e = e;
    GPE_CUDA(c, c->d_vde.reserve(std::max<size_t>((size_t)c->V * e, 1) * sizeof(double)));
    if (c->V) GPE_CUDA(c, cudaMemcpyAsync(c->d_vde.p, vde, (size_t)c->V * e * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GPE_CUDA(c, c->d_vrec.reserve(k1_vertex_record_bytes(c->V, e)));
    GPE_CUDA(c, k1_vertex_records(graph_view(c), c->d_vrec.p, c->stream));
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stats.kernel_launches++;
    c->have_emb = true;
    c->have_table = false;
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_set_embeddings", ex.what()) : GPE_ERR_INVALID;
}

// ---------------------------------------------------------------------------------------------------------------
int gpe_enumerate(gpe_ctx *c, uint32_t L, const uint32_t *sorted_nodes, const uint32_t *membership, uint32_t p,
                  uint64_t *rows_per_partition, uint64_t *n_rows) try {
    if (!c || !sorted_nodes || !membership || p == 0) return c ? c->fail(GPE_ERR_INVALID, "null argument") : GPE_ERR_INVALID;
    if (!c->have_graph) return c->fail(GPE_ERR_INVALID, "gpe_set_graph first");
    if (L != 3 && L != 4)
        return c->fail(GPE_ERR_UNSUPPORTED, "path length l=%u unsupported: kernels are built for l=2 and l=3 "
                                            "(the reference itself only handles l=2, SURVEY.md F5)", L - 1);
    GPE_CUDA(c, cudaSetDevice(c->device));
    const u32 V = c->V;
    std::vector<u32> rank(V, 0xffffffffu), offr((size_t)V + 1, 0), h_off((size_t)V + 1);
    GPE_CUDA(c, cudaMemcpy(h_off.data(), c->d_off.p, ((size_t)V + 1) * sizeof(u32), cudaMemcpyDeviceToHost));
    for (u32 i = 0; i < V; i++) {
        u32 v = sorted_nodes[i];
        if (v >= V || rank[v] != 0xffffffffu) return c->fail(GPE_ERR_INVALID, "sorted_nodes is not a permutation of the vertices");
        rank[v] = i;
        if (membership[v] >= p) return c->fail(GPE_ERR_INVALID, "membership[%u]=%u outside [0,%u)", v, membership[v], p);
        offr[i + 1] = offr[i] + (h_off[v + 1] - h_off[v]);
    }
    c->L = L;
    c->p = p;
    c->h_member.assign(membership, membership + V);
    GPE_CUDA(c, c->d_rank.reserve(std::max<size_t>(V, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_ranklab.reserve(std::max<size_t>(V, 1) * sizeof(u64)));
    GPE_CUDA(c, c->d_sorted.reserve(std::max<size_t>(V, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_member.reserve(std::max<size_t>(V, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_offr.reserve(((size_t)V + 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_ebase.reserve(((size_t)c->n_adj + 1) * sizeof(u64)));
    GPE_CUDA(c, c->d_start_rows.reserve(((size_t)V + 1 + p) * sizeof(u64)));  // start_rows | part_rows
    if (V) {
        GPE_CUDA(c, cudaMemcpyAsync(c->d_rank.p, rank.data(), (size_t)V * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
        GPE_CUDA(c, cudaMemcpyAsync(c->d_sorted.p, sorted_nodes, (size_t)V * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
        GPE_CUDA(c, cudaMemcpyAsync(c->d_member.p, membership, (size_t)V * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
    }
    GPE_CUDA(c, cudaMemcpyAsync(c->d_offr.p, offr.data(), ((size_t)V + 1) * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
    GPE_CUDA(c, k1_rank_labels(V, c->d_rank.as<u32>(), c->d_label.as<u32>(), c->d_ranklab.as<u64>(), c->stream));
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));

    u64 *ebase = c->d_ebase.as<u64>();
    u64 *start_rows = c->d_start_rows.as<u64>();
    u64 *part_rows = start_rows + V + 1;
    {
        StageTimer tm(c, &c->stats.last_enumerate_ms, kStageEnumerate);
        GPE_CUDA(c, cudaMemsetAsync(ebase + c->n_adj, 0, sizeof(u64), c->stream));
        GPE_CUDA(c, cudaMemsetAsync(part_rows, 0, p * sizeof(u64), c->stream));
        GPE_CUDA(c, k1_count(graph_view(c), L, c->d_sorted.as<u32>(), c->d_offr.as<u32>(), ebase, c->sm_count, c->stream));
        GPE_CUDA(c, exclusive_scan_u64(ebase, (u64)c->n_adj + 1, c->d_scan_tmp, c->stream));
        GPE_CUDA(c, k1_rows_per_partition(V, c->d_sorted.as<u32>(), c->d_offr.as<u32>(), ebase, c->d_member.as<u32>(),
                                          part_rows, start_rows, c->stream));
    }
    c->stats.build_launches += 3;
    c->stats.kernel_launches += 2 + exclusive_scan_launches((u64)c->n_adj + 1);
    std::vector<u64> host((size_t)V + 1 + p);
    GPE_CUDA(c, cudaMemcpyAsync(host.data(), start_rows, host.size() * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));
    c->h_bucket_start.clear();
    c->n_rows = host[V];
    if (rows_per_partition) std::copy(host.begin() + V + 1, host.end(), rows_per_partition);
    if (n_rows) *n_rows = c->n_rows;
    c->have_enum = true;
    c->have_table = false;
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_enumerate", ex.what()) : GPE_ERR_INVALID;
}

int gpe_start_rows(gpe_ctx *c, uint64_t *start_row) try {
    if (!c || !start_row) return GPE_ERR_INVALID;
    if (!c->have_enum) return c->fail(GPE_ERR_INVALID, "gpe_enumerate first");
    GPE_CUDA(c, cudaSetDevice(c->device));
    GPE_CUDA(c, cudaMemcpy(start_row, c->d_start_rows.p, ((size_t)c->V + 1) * sizeof(u64), cudaMemcpyDeviceToHost));
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_start_rows", ex.what()) : GPE_ERR_INVALID;
}

int gpe_dump_paths(gpe_ctx *c, uint64_t first, uint64_t n, uint32_t *vids) try {
    if (!c || (!vids && n)) return GPE_ERR_INVALID;
    if (!c->have_enum) return c->fail(GPE_ERR_INVALID, "gpe_enumerate first");
    if (first + n > c->n_rows) return c->fail(GPE_ERR_INVALID, "rows [%llu,%llu) outside the table of %llu rows",
                                              (unsigned long long)first, (unsigned long long)(first + n), (unsigned long long)c->n_rows);
    if (n == 0) return GPE_OK;
    GPE_CUDA(c, cudaSetDevice(c->device));
    std::vector<u64> start((size_t)c->V + 1);
    GPE_CUDA(c, cudaMemcpy(start.data(), c->d_start_rows.p, start.size() * sizeof(u64), cudaMemcpyDeviceToHost));
    u32 lo = (u32)(std::upper_bound(start.begin(), start.end(), first) - start.begin() - 1);
    u32 hi = (u32)(std::upper_bound(start.begin(), start.end(), first + n - 1) - start.begin() - 1);
    if (hi >= c->V) hi = c->V - 1;
    DevBuf out;
    GPE_CUDA(c, out.reserve(n * c->L * sizeof(u32)));
    cudaError_t e = k1_dump(graph_view(c), c->L, c->d_sorted.as<u32>(), c->d_offr.as<u32>(), c->d_ebase.as<u64>(), lo, hi,
                            first, n, out.as<u32>(), c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(vids, out.p, n * c->L * sizeof(u32), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    out.release();
    GPE_CUDA(c, e);
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_dump_paths", ex.what()) : GPE_ERR_INVALID;
}

// ---------------------------------------------------------------------------------------------------------------
int gpe_build_table(gpe_ctx *c, const uint8_t *part_select, uint64_t *n_table_rows) try {
    if (!c) return GPE_ERR_INVALID;
    if (!c->have_enum || !c->have_emb) return c->fail(GPE_ERR_INVALID, "gpe_enumerate and gpe_set_embeddings first");
    if (!k2_supported(c->L, c->e))
        return c->fail(GPE_ERR_UNSUPPORTED, "no scan kernel compiled for L=%u, e=%u (L in {3,4}, e in {1,2,3,4,8})", c->L, c->e);
    GPE_CUDA(c, cudaSetDevice(c->device));
    TableView t{};
    t.L = c->L;
    t.E = c->e;
    t.D = c->L * c->e;
    t.tile_bytes = kTileRows * (8 * t.L + 8 * t.D);
    // label-sequence directory: mixed radix over the label alphabet, truncated to kKeyBudget buckets
    u64 budget = kKeyBudget;
    for (u32 k = 0; k < (u32)kMaxL; k++) t.key_radix[k] = 1;
    for (u32 k = 0; k < t.L; k++) {
        u64 r = std::min<u64>(std::max<u32>(c->n_labels, 1), budget);
        t.key_radix[k] = (u32)std::max<u64>(r, 1);
        budget = std::max<u64>(budget / t.key_radix[k], 1);
    }
    u64 stride = 1;
    for (int k = (int)t.L - 1; k >= 0; k--) { t.key_stride[k] = (u32)stride; stride *= t.key_radix[k]; }
    for (u32 k = t.L; k < (u32)kMaxL; k++) t.key_stride[k] = 0;
    t.n_keys = (u32)stride;

    DevBuf d_sel;
    const unsigned char *sel = nullptr;
    if (part_select) {
        GPE_CUDA(c, d_sel.reserve(c->p));
        GPE_CUDA(c, cudaMemcpy(d_sel.p, part_select, c->p, cudaMemcpyHostToDevice));
        sel = d_sel.as<unsigned char>();
    }
    GPE_CUDA(c, c->d_bucket.reserve(((size_t)t.n_keys + 1) * sizeof(u64)));
    GPE_CUDA(c, c->d_cursor.reserve(((size_t)t.n_keys + 1) * sizeof(u64)));
    u64 *bucket = c->d_bucket.as<u64>();
    GraphView g = graph_view(c);
    cudaEvent_t b0, b1;
    // build time = the kernels (histogram + scan, then fill + expand); the allocation of the table between the two parts
    // (tens to hundreds of ms of cudaMalloc for tens of GB, box dependent) is not a kernel and is left out
    cudaEvent_t b0b, b1b;
    cudaEventCreate(&b0);
    cudaEventCreate(&b1);
    cudaEventCreate(&b0b);
    cudaEventCreate(&b1b);
    cudaEventRecord(b0, c->stream);
    GPE_CUDA(c, cudaMemsetAsync(bucket, 0, ((size_t)t.n_keys + 1) * sizeof(u64), c->stream));
    GPE_CUDA(c, k1_histogram(g, t, c->d_sorted.as<u32>(), c->d_member.as<u32>(), sel, bucket, c->sm_count, c->stream));
    GPE_CUDA(c, exclusive_scan_u64(bucket, (u64)t.n_keys + 1, c->d_scan_tmp, c->stream));
    cudaEventRecord(b1b, c->stream);
    c->h_bucket_start.resize((size_t)t.n_keys + 1);
    GPE_CUDA(c, cudaMemcpyAsync(c->h_bucket_start.data(), bucket, ((size_t)t.n_keys + 1) * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));
    t.n_rows = c->h_bucket_start[t.n_keys];
    t.n_tiles = (t.n_rows + kTileRows - 1) / kTileRows;
    if (t.n_tiles >= 0xffffffffull) return c->fail(GPE_ERR_UNSUPPORTED, "table has more than 2^32 tiles");
    size_t tile_bytes_total = std::max<u64>(t.n_tiles, 1) * t.tile_bytes;
    const size_t vids_bytes = std::max<u64>(t.n_tiles, 1) * t.L * kTileRows * sizeof(u32);
    // Layout: materialised scan rows (8L + 8Le bytes each, what k2_scan streams) next to the ids, or the ids alone
    // (4L bytes per row, everything else gathered by k2_scan_ids) when the rows would not fit -- config 4's l=3, e=4 table
    // is 160 bytes x 2 x 10^10 rows materialised.  Auto: materialise while table + ids stay below 70 % of the free HBM.
    {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        free_b += c->d_tiles.cap + c->d_vids.cap + c->d_sum_u32.cap + c->d_sum_f64.cap;  // what a rebuild would reuse
        t.ids_only = c->table_layout == 2 || (c->table_layout == 0 && (double)(tile_bytes_total + vids_bytes) > 0.7 * (double)free_b);
    }
    if (t.ids_only) {
        c->d_tiles.release();
        c->d_sum_u32.release();
        c->d_sum_f64.release();
        tile_bytes_total = 0;
    }
    cudaError_t e1 = t.ids_only ? cudaSuccess : c->d_tiles.reserve(tile_bytes_total);
    cudaError_t e2 = e1 == cudaSuccess ? c->d_vids.reserve(vids_bytes) : e1;
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        cudaGetLastError();
        return c->fail(GPE_ERR_CUDA, "path table of %llu rows needs %.1f GB of HBM: %s", (unsigned long long)t.n_rows,
                       (tile_bytes_total + vids_bytes) / 1e9, cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    }
    if (!t.ids_only) {
        GPE_CUDA(c, c->d_sum_u32.reserve(std::max<u64>(t.n_tiles, 1) * 3 * t.L * sizeof(u32)));
        GPE_CUDA(c, c->d_sum_f64.reserve(std::max<u64>(t.n_tiles, 1) * t.D * sizeof(double)));
    }
    t.tiles = c->d_tiles.as<unsigned char>();
    t.vids = c->d_vids.as<u32>();
    t.lab_min = c->d_sum_u32.as<u32>();
    t.lab_max = t.lab_min + t.n_tiles * t.L;
    t.deg_max = t.lab_max + t.n_tiles * t.L;
    t.pde_max = c->d_sum_f64.as<double>();
    t.bucket_start = bucket;
    if (t.n_tiles) {
        // only the tail of the last tile is never written
        if (!t.ids_only) GPE_CUDA(c, cudaMemsetAsync(t.tiles + (t.n_tiles - 1) * t.tile_bytes, 0, t.tile_bytes, c->stream));
        GPE_CUDA(c, cudaMemsetAsync(t.vids + (t.n_tiles - 1) * t.L * kTileRows, 0, t.L * kTileRows * sizeof(u32), c->stream));
    }
    cudaEventRecord(b0b, c->stream);
    GPE_CUDA(c, cudaMemcpyAsync(c->d_cursor.p, bucket, ((size_t)t.n_keys + 1) * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
    GPE_CUDA(c, k1_fill(g, t, c->d_sorted.as<u32>(), c->d_member.as<u32>(), sel, c->d_cursor.as<u64>(), c->sm_count, c->stream));
    if (!t.ids_only) GPE_CUDA(c, k1_expand(t, c->d_vrec.p, c->sm_count, c->stream));
    cudaEventRecord(b1, c->stream);
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));
    {
        float part1 = 0, part2 = 0;
        cudaEventElapsedTime(&part1, b0, b1b);
        cudaEventElapsedTime(&part2, b0b, b1);
        c->stats.last_build_ms = part1 + part2;
    }
    cudaEventDestroy(b0b);
    cudaEventDestroy(b1b);
    cudaEventDestroy(b0);
    cudaEventDestroy(b1);
    d_sel.release();
    c->stats.build_launches += 5;
    c->stats.kernel_launches += 3 + exclusive_scan_launches((u64)t.n_keys + 1);
    c->tv = t;
    c->have_table = true;
    if (n_table_rows) *n_table_rows = t.n_rows;
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_build_table", ex.what()) : GPE_ERR_INVALID;
}

int gpe_set_table_layout(gpe_ctx *c, int layout) {
    if (!c || layout < 0 || layout > 2) return c ? c->fail(GPE_ERR_INVALID, "layout must be 0 (auto), 1 (rows) or 2 (ids only)") : GPE_ERR_INVALID;
    c->table_layout = layout;
    return GPE_OK;
}

int gpe_dump_table(gpe_ctx *c, uint64_t first, uint64_t n, uint32_t *vids, uint32_t *labels, uint32_t *degs, double *pde) try {
    if (!c) return GPE_ERR_INVALID;
    if (!c->have_table) return c->fail(GPE_ERR_INVALID, "gpe_build_table first");
    if (first + n > c->tv.n_rows) return c->fail(GPE_ERR_INVALID, "row range outside the table");
    if (n == 0) return GPE_OK;
    GPE_CUDA(c, cudaSetDevice(c->device));
    const u32 L = c->tv.L, D = c->tv.D;
    DevBuf dv, dl, dd, dp;
    cudaError_t e = cudaSuccess;
    if (vids && e == cudaSuccess) e = dv.reserve(n * L * sizeof(u32));
    if (labels && e == cudaSuccess) e = dl.reserve(n * L * sizeof(u32));
    if (degs && e == cudaSuccess) e = dd.reserve(n * L * sizeof(u32));
    if (pde && e == cudaSuccess) e = dp.reserve(n * D * sizeof(double));
    if (e == cudaSuccess)
        e = k1_dump_table(c->tv, graph_view(c), c->d_vrec.p, first, n, vids ? dv.as<u32>() : nullptr, labels ? dl.as<u32>() : nullptr,
                          degs ? dd.as<u32>() : nullptr, pde ? dp.as<double>() : nullptr, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (vids && e == cudaSuccess) e = cudaMemcpy(vids, dv.p, n * L * sizeof(u32), cudaMemcpyDeviceToHost);
    if (labels && e == cudaSuccess) e = cudaMemcpy(labels, dl.p, n * L * sizeof(u32), cudaMemcpyDeviceToHost);
    if (degs && e == cudaSuccess) e = cudaMemcpy(degs, dd.p, n * L * sizeof(u32), cudaMemcpyDeviceToHost);
    if (pde && e == cudaSuccess) e = cudaMemcpy(pde, dp.p, n * D * sizeof(double), cudaMemcpyDeviceToHost);
    dv.release(); dl.release(); dd.release(); dp.release();
    GPE_CUDA(c, e);
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_dump_table", ex.what()) : GPE_ERR_INVALID;
}

// ---------------------------------------------------------------------------------------------------------------
int gpe_filter(gpe_ctx *c, uint32_t n_qpaths, const uint32_t *q_vids, const uint32_t *q_labels, const uint32_t *q_degs,
               const double *q_pde, uint32_t nq, uint32_t flags, uint64_t *cand_offsets, uint64_t *survivors) try {
    if (!c) return GPE_ERR_INVALID;
    if (!c->have_table) return c->fail(GPE_ERR_INVALID, "gpe_build_table first");
    if (n_qpaths && (!q_vids || !q_labels || !q_degs || !q_pde)) return c->fail(GPE_ERR_INVALID, "null plan arrays");
    GPE_CUDA(c, cudaSetDevice(c->device));
    const u32 L = c->tv.L, D = c->tv.D;
    QPathSet qp;
    qp.L = L;
    qp.D = D;
    for (u32 i = 0; i < n_qpaths * L; i++)
        if (q_vids[i] >= nq) return c->fail(GPE_ERR_INVALID, "plan path vertex %u outside the query (nq=%u)", q_vids[i], nq);
    qp.slots.assign(q_vids, q_vids + (size_t)n_qpaths * L);
    qp.labels.assign(q_labels, q_labels + (size_t)n_qpaths * L);
    qp.degs.assign(q_degs, q_degs + (size_t)n_qpaths * L);
    qp.pde.assign(q_pde, q_pde + (size_t)n_qpaths * D);
    c->b_nq = 1;
    if (flags & GPE_FILTER_BOTH_ORIENTATIONS) qp.add_reversed();
    int rc = setup_filter(c, qp, nq, flags);
    if (rc) return rc;
    rc = run_filter(c);
    if (rc) return rc;
    if ((rc = settle_candidates(c))) return rc;
    if (cand_offsets)
        GPE_CUDA(c, cudaMemcpyAsync(cand_offsets, c->d_cand_off.p, ((size_t)nq + 1) * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    if (survivors && n_qpaths)
        GPE_CUDA(c, cudaMemcpyAsync(survivors, c->d_survivors.p, (size_t)n_qpaths * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_filter", ex.what()) : GPE_ERR_INVALID;
}

int gpe_get_candidates(gpe_ctx *c, uint32_t *cand) try {
    if (!c || !cand) return GPE_ERR_INVALID;
    if (!c->b_filtered) return c->fail(GPE_ERR_INVALID, "no filter result");
    GPE_CUDA(c, cudaSetDevice(c->device));
    if (int rc = settle_candidates(c)) return rc;
    return download_candidates(c, cand);
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_get_candidates", ex.what()) : GPE_ERR_INVALID;
}

// ---------------------------------------------------------------------------------------------------------------
uint64_t gpe_clamp_answer(uint64_t raw_total, uint64_t limit) {
    // custom.h:846-855: the count stops growing once it reaches the limit, and the test comes after the
    // increment, so a limit of 0 still reports one match.
    if (limit == 0) limit = 1;
    return raw_total < limit ? raw_total : limit;
}

int gpe_refine(gpe_ctx *c, uint32_t nq, const uint32_t *q_offsets, const uint32_t *q_nbrs, const uint32_t *q_labels,
               const uint64_t *cand_offsets, const uint32_t *cand, uint64_t limit, uint64_t *n_matches,
               uint32_t *order_out, uint32_t *pivot_out, uint32_t *matches, uint64_t matches_cap) try {
    if (!c || !q_offsets || !q_labels || !cand_offsets) return c ? c->fail(GPE_ERR_INVALID, "null argument") : GPE_ERR_INVALID;
    if (!c->have_graph) return c->fail(GPE_ERR_INVALID, "gpe_set_graph first");
    GPE_CUDA(c, cudaSetDevice(c->device));
    std::string why;
    if (!check_query(c, nq, q_offsets, q_nbrs, why)) return c->fail(GPE_ERR_INVALID, "%s", why.c_str());
    const u64 total = cand_offsets[nq];
    for (u64 i = 0; i < total; i++)
        if (cand[i] >= c->V) return c->fail(GPE_ERR_INVALID, "candidate id out of range");
    u32 vbase[2] = {0, nq}, ebase[2] = {0, q_offsets[nq]};
    int rc = upload_queries(c, 1, vbase, ebase, q_offsets, q_nbrs, q_labels, &limit);
    if (rc) return rc;
    if (!c->d_rank.p) {  // the join never reads rank/vde, but the view wants valid pointers
        GPE_CUDA(c, c->d_rank.reserve(16));
    }
    c->b_slots = nq;
    c->b_n_cand = total;
    c->b_cand_known = true;
    c->b_cand_external_lists = true;  // no compaction ran: nothing to settle
    c->b_cand_clean = false;  // caller-supplied sets: the reference takes them as they are
    GPE_CUDA(c, c->d_cand.reserve(std::max<u64>(total, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_cand_off.reserve(((size_t)nq + 1) * sizeof(u64)));
    if (total) {  // the caller's ids -> class order (the join's id space)
        DevBuf tmp;
        GPE_CUDA(c, tmp.reserve(total * sizeof(u32)));
        cudaError_t e = cudaMemcpyAsync(tmp.p, cand, total * sizeof(u32), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = k0_gather(total, c->d_newid.as<u32>(), tmp.as<u32>(), c->d_cand.as<u32>(), c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        tmp.release();
        GPE_CUDA(c, e);
    }
    GPE_CUDA(c, cudaMemcpyAsync(c->d_cand_off.p, cand_offsets, ((size_t)nq + 1) * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    u32 *d_matches = nullptr;
    if (matches && matches_cap) {
        GPE_CUDA(c, c->d_matches.reserve(matches_cap * nq * sizeof(u32)));
        d_matches = c->d_matches.as<u32>();
    }
    rc = run_join(c, 0, 1, d_matches, d_matches ? matches_cap : 0);
    if (rc) return rc;
    u64 raw = 0, n_emitted = 0;
    GPE_CUDA(c, cudaMemcpyAsync(&raw, c->d_answers.p, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    GPE_CUDA(c, cudaMemcpyAsync(&n_emitted, c->d_match_cursor.p, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    if (order_out) GPE_CUDA(c, cudaMemcpyAsync(order_out, c->d_order.p, nq * sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    if (pivot_out) GPE_CUDA(c, cudaMemcpyAsync(pivot_out, c->d_pivot.p, nq * sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));
    if (d_matches) {
        u64 n = std::min<u64>(n_emitted, matches_cap);
        if (n) GPE_CUDA(c, cudaMemcpy(matches, d_matches, n * nq * sizeof(u32), cudaMemcpyDeviceToHost));
    }
    if (n_matches) *n_matches = gpe_clamp_answer(raw, limit);
    if ((rc = read_join_stats(c))) return rc;
    c->b_filtered = false;
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_refine", ex.what()) : GPE_ERR_INVALID;
}

// ---------------------------------------------------------------------------------------------------------------
// Host planning of a batch (dfs_query + gen_vde(query) + gen_query_pde per query, main.cpp:142-158): independent of the
// GPU, so one plan serves every context of a multi-GPU run, and it can be computed while a previous batch is in flight.
struct gpe_plan {
    u32 L = 0, E = 0;
    std::vector<QueryPlan> plans;
    std::string err;
    FilterStage stage;            // the batch staged for the scan of `staged_for` (stage_planned), or nothing
    const gpe_ctx *staged_for = nullptr;
};

static int plan_batch(const gpe_batch *b, u32 L, u32 E, LabelTable &table, gpe_plan &out, int ranks_on_host = 1) {
    out.L = L;
    out.E = E;
    std::string why;
    for (u32 q = 0; q < b->n_queries; q++) {
        const u32 vb = b->q_vbase[q], nq = b->q_vbase[q + 1] - vb;
        if (!check_query(nullptr, nq, b->q_offsets + vb + q, b->q_nbrs + b->q_ebase[q], why)) {
            out.err = "query " + std::to_string(q) + ": " + why;
            return GPE_ERR_INVALID;
        }
    }
    // the reference plans one query at a time; here the queries of a batch are planned by a few host threads, label
    // embeddings come from a table filled once per batch
    table.fill(b->q_labels, b->q_vbase[b->n_queries], E);
    out.plans.assign(b->n_queries, QueryPlan());
    auto plan_range = [&](u32 q0, u32 q1) {
        for (u32 q = q0; q < q1; q++) {
            const u32 vb = b->q_vbase[q], nq = b->q_vbase[q + 1] - vb;
            query_plan(nq, b->q_offsets + vb + q, b->q_nbrs + b->q_ebase[q], b->q_labels + vb, L, E, out.plans[q], &table);
        }
    };
    // (with one process per GPU every rank plans the whole batch: the ranks share the host's cores)
    const u32 hw = std::max<u32>(1, std::thread::hardware_concurrency());
    const u32 n_thr = b->n_queries >= 64 ? std::min<u32>(8, std::max<u32>(1, ranks_on_host > 1 ? hw / (u32)ranks_on_host : hw / 2)) : 1;
    if (n_thr <= 1) {
        plan_range(0, b->n_queries);
    } else {
        std::vector<std::thread> pool;
        const u32 per = (b->n_queries + n_thr - 1) / n_thr;
        for (u32 t = 1; t < n_thr; t++) pool.emplace_back(plan_range, std::min(t * per, b->n_queries), std::min((t + 1) * per, b->n_queries));
        plan_range(0, std::min(per, b->n_queries));
        for (auto &th : pool) th.join();
    }
    return GPE_OK;
}

// plans -> plan-path arrays -> staged filter, host only (the part of an upload that can run while the GPU is busy)
static void stage_planned(const gpe_ctx *c, const gpe_batch *b, gpe_plan &pl, uint32_t flags) {
    const u32 L = c->tv.L, D = c->tv.D;
    QPathSet qp;
    qp.L = L;
    qp.D = D;
    size_t total = 0;
    for (const QueryPlan &plan : pl.plans) total += plan.n;
    qp.slots.reserve(total * L * 2); qp.labels.reserve(total * L * 2); qp.degs.reserve(total * L * 2); qp.pde.reserve(total * D * 2);
    for (u32 q = 0; q < b->n_queries; q++) {
        const QueryPlan &plan = pl.plans[q];
        const u32 vb = b->q_vbase[q];
        for (u32 i = 0; i < plan.n * L; i++) qp.slots.push_back(vb + plan.vids[i]);
        qp.labels.insert(qp.labels.end(), plan.labels.begin(), plan.labels.end());
        qp.degs.insert(qp.degs.end(), plan.degs.begin(), plan.degs.end());
        qp.pde.insert(qp.pde.end(), plan.pde.begin(), plan.pde.end());
    }
    if (flags & GPE_FILTER_BOTH_ORIENTATIONS) qp.add_reversed();
    stage_filter(c, qp, b->q_vbase[b->n_queries], flags, pl.stage);
    pl.staged_for = c;
}

static int upload_planned(gpe_ctx *c, const gpe_batch *b, const gpe_plan &pl, uint32_t flags) {
    const u32 L = c->tv.L, D = c->tv.D;
    if (pl.L != L || pl.E != c->tv.E || pl.plans.size() != b->n_queries) return c->fail(GPE_ERR_INVALID, "plan does not belong to this batch / table");
    if (pl.staged_for == c && pl.stage.flags == flags) {  // staged ahead: only the copies are left
        int rc = upload_queries(c, b->n_queries, b->q_vbase, b->q_ebase, b->q_offsets, b->q_nbrs, b->q_labels, b->limits);
        if (rc) return rc;
        c->b_pge = false;
        return commit_filter(c, pl.stage);
    }
    QPathSet qp;
    qp.L = L;
    qp.D = D;
    for (u32 q = 0; q < b->n_queries; q++) {
        const QueryPlan &plan = pl.plans[q];
        const u32 vb = b->q_vbase[q];
        for (u32 i = 0; i < plan.n * L; i++) qp.slots.push_back(vb + plan.vids[i]);
        qp.labels.insert(qp.labels.end(), plan.labels.begin(), plan.labels.end());
        qp.degs.insert(qp.degs.end(), plan.degs.begin(), plan.degs.end());
        qp.pde.insert(qp.pde.end(), plan.pde.begin(), plan.pde.end());
    }
    if (flags & GPE_FILTER_BOTH_ORIENTATIONS) qp.add_reversed();
    int rc = upload_queries(c, b->n_queries, b->q_vbase, b->q_ebase, b->q_offsets, b->q_nbrs, b->q_labels, b->limits);
    if (rc) return rc;
    c->b_pge = false;
    return setup_filter(c, qp, b->q_vbase[b->n_queries], flags);
}

int gpe_batch_upload(gpe_ctx *c, const gpe_batch *b, uint32_t flags) try {
    if (!c || !b || !b->q_vbase || !b->q_ebase || !b->q_offsets || !b->q_labels) return c ? c->fail(GPE_ERR_INVALID, "null argument") : GPE_ERR_INVALID;
    if (!c->have_table) return c->fail(GPE_ERR_INVALID, "gpe_build_table first");
    GPE_CUDA(c, cudaSetDevice(c->device));
    gpe_plan pl;
    if (int rc = plan_batch(b, c->tv.L, c->tv.E, c->label_table, pl, c->comm_world)) return c->fail(rc, "%s", pl.err.c_str());
    return upload_planned(c, b, pl, flags);
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_batch_upload", ex.what()) : GPE_ERR_INVALID;
}

int gpe_batch_filter(gpe_ctx *c) {
    if (!c) return GPE_ERR_INVALID;
    if (!c->have_table) return c->fail(GPE_ERR_INVALID, "gpe_build_table first");
    GPE_CUDA(c, cudaSetDevice(c->device));
    return run_filter(c);
}

int gpe_batch_join(gpe_ctx *c, uint32_t rank, uint32_t world) {
    if (!c) return GPE_ERR_INVALID;
    if (!c->b_filtered) return c->fail(GPE_ERR_INVALID, "gpe_batch_filter (or gpe_batch_cand_merge) first");
    if (world == 0 || rank >= world) return c->fail(GPE_ERR_INVALID, "bad rank/world");
    GPE_CUDA(c, cudaSetDevice(c->device));
    return run_join(c, rank, world, nullptr, 0);
}

int gpe_batch_download(gpe_ctx *c, uint64_t *raw_counts) try {
    if (!c || !raw_counts) return GPE_ERR_INVALID;
    if (!c->b_joined) return c->fail(GPE_ERR_INVALID, "gpe_batch_join first");
    GPE_CUDA(c, cudaSetDevice(c->device));
    const size_t nq = c->b_nq;
    GPE_CUDA(c, c->h_pin2.reserve((2 * std::max<size_t>(nq, 16) + 32) * sizeof(u64)));
    u64 *pin = c->h_pin2.as<u64>();
    u64 *pin_flags = pin + nq + 16;  // behind the queue header
    GPE_CUDA(c, cudaMemcpyAsync(pin, c->d_answers.p, nq * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    GPE_CUDA(c, cudaMemcpyAsync(pin + nq, c->d_jq.p, sizeof(JoinQueue), cudaMemcpyDeviceToHost, c->stream));
    static_assert(sizeof(JoinQueue) <= 16 * sizeof(u64), "pinned layout");
    if (nq) GPE_CUDA(c, cudaMemcpyAsync(pin_flags, c->d_answers.as<u64>() + nq + 8, nq * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));
    c->need_dense_redo = false;
    if (c->b_sparse_used) {  // sparse exchange: did every shard's non-zero words fit the buffer?  (same numbers on every rank)
        const u64 *h = c->h_pin4.as<u64>();
        u64 mx = 0;
        for (int r = 0; r < c->comm_world; r++) mx = std::max(mx, h[2 * r]);
        c->sparse_seen_max = std::max(c->sparse_seen_max, mx);
        c->need_dense_redo = mx > c->sparse_cap;  // gpe_batch_finish / gpe_multi_batch_finish redo the step (a collective)
        if (c->need_dense_redo) {
            for (u32 q = 0; q < c->b_nq; q++) raw_counts[q] = 0;
            return GPE_OK;
        }
    }
    if (!c->b_cand_known) {  // first look at what the compaction reported: did the candidate lists fit their buffer?
        bool overflow = false;
        if (int rc = cand_total(c, &overflow)) return rc;
        if (overflow) {  // the join saw the flag and did nothing: expand the lists again, join again, read again
            if (int rc = recompact(c)) return rc;
            if (int rc = run_join(c, c->b_rank, c->b_world, nullptr, 0)) return rc;
            return gpe_batch_download(c, raw_counts);
        }
    }
    // raw counts are exact below 2^44 and saturated beyond (k3_join.cu kSat); capped so that shards can still be summed
    for (size_t q = 0; q < nq; q++) pin[q] = std::min<u64>(pin[q], 1ull << 48);
    memcpy(raw_counts, pin, (size_t)c->b_nq * sizeof(u64));
    JoinQueue jq;
    memcpy(&jq, pin + c->b_nq, sizeof jq);
    {   // queries whose weighted counted leaves met a saturated table entry (a difference of two numbers beyond 2^62):
        // joined once more with those leaves walked -- everything that is left saturates monotonically, so
        // min(count, limit) is exact.  Local to this shard: each rank repairs the share of its own start candidates.
        std::vector<u32> qmode;
        for (size_t q = 0; q < nq; q++)
            if (pin_flags[q]) { qmode.assign(nq, 2u); break; }
        if (!qmode.empty()) {
            for (size_t q = 0; q < nq; q++)
                if (pin_flags[q]) qmode[q] = 1u;
            GPE_CUDA(c, c->d_qmode.reserve(nq * sizeof(u32)));
            GPE_CUDA(c, cudaMemcpyAsync(c->d_qmode.p, qmode.data(), nq * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
            int rc = run_join(c, c->b_rank, c->b_world, nullptr, 0, /*force_dfs=*/true, c->d_qmode.as<u32>());
            if (rc) return rc;
            std::vector<u64> second(nq);
            GPE_CUDA(c, cudaMemcpyAsync(second.data(), c->d_answers.p, nq * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
            GPE_CUDA(c, cudaStreamSynchronize(c->stream));
            for (size_t q = 0; q < nq; q++)
                if (qmode[q] == 1u) raw_counts[q] = std::min<u64>(second[q], 1ull << 48);
            c->stats.join_reruns++;
        }
    }
    c->stats.join_items = jq.tail;
    c->stats.join_exports = jq.exports;
    c->stats.join_donations = jq.donations;
    c->stats.join_steps = jq.steps;
    c->stats.join_warp_iters = jq.warp_iters;
    c->stats.join_idle_polls = jq.idle_polls;
    c->stats.d2h_bytes += (size_t)c->b_nq * sizeof(u64) + sizeof(JoinQueue);
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_batch_download", ex.what()) : GPE_ERR_INVALID;
}

int gpe_query_batch(gpe_ctx *c, const gpe_batch *b, uint32_t flags, uint64_t *answers) try {
    if (!c || !answers) return GPE_ERR_INVALID;
    int rc = gpe_batch_upload(c, b, flags);
    if (rc) return rc;
    if ((rc = run_filter(c))) return rc;
    if ((rc = run_join(c, 0, 1, nullptr, 0))) return rc;
    if ((rc = gpe_batch_download(c, answers))) return rc;
    for (u32 q = 0; q < b->n_queries; q++) answers[q] = gpe_clamp_answer(answers[q], c->h_limits[q]);
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_query_batch", ex.what()) : GPE_ERR_INVALID;
}

int gpe_batch_cand_info(gpe_ctx *c, uint64_t *n_slots, uint64_t *n_cand_total) {
    if (!c) return GPE_ERR_INVALID;
    if (!c->b_filtered) return c->fail(GPE_ERR_INVALID, "no filter result");
    GPE_CUDA(c, cudaSetDevice(c->device));
    if (int rc = settle_candidates(c)) return rc;
    if (n_slots) *n_slots = c->b_slots;
    if (n_cand_total) *n_cand_total = c->b_n_cand;
    return GPE_OK;
}

int gpe_batch_cand_export(gpe_ctx *c, void *d_counts_u32, void *d_cand_u32) {
    if (!c || !d_counts_u32) return GPE_ERR_INVALID;
    if (!c->b_filtered) return c->fail(GPE_ERR_INVALID, "no filter result");
    GPE_CUDA(c, cudaSetDevice(c->device));
    if (int rc = settle_candidates(c)) return rc;
    GPE_CUDA(c, k3_counts_from_offsets(c->d_cand_off.as<u64>(), c->b_slots, (u32 *)d_counts_u32, c->stream));
    if (c->b_n_cand && d_cand_u32)
        GPE_CUDA(c, cudaMemcpyAsync(d_cand_u32, c->d_cand.p, c->b_n_cand * sizeof(u32), cudaMemcpyDeviceToDevice, c->stream));
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));
    return GPE_OK;
}

int gpe_batch_cand_merge(gpe_ctx *c, uint32_t world, const void *d_counts, const void *d_cand, uint64_t stride) {
    if (!c || !d_counts || world == 0) return GPE_ERR_INVALID;
    if (!c->b_slots && !c->b_filtered) return c->fail(GPE_ERR_INVALID, "no batch");
    GPE_CUDA(c, cudaSetDevice(c->device));
    GPE_CUDA(c, cudaMemsetAsync(c->d_bitmap.p, 0, std::max<u64>((u64)c->b_slots * c->b_words, 1) * sizeof(u32), c->stream));
    GPE_CUDA(c, c->d_chunk_cnt.reserve(std::max<size_t>((size_t)world * c->b_slots, 1) * sizeof(u64)));
    GPE_CUDA(c, k3_scatter((const u32 *)d_counts, (const u32 *)d_cand, stride, world, c->b_slots, c->d_slot_label.as<u32>(),
                           c->d_lcoff.as<u32>(), c->n_labels, c->d_bitmap.as<u32>(), c->b_words, c->d_chunk_cnt.as<u64>(),
                           c->stream));
    int rc = compact_candidates(c);
    if (rc) return rc;
    c->b_filtered = true;
    c->b_cand_external = true;
    c->b_cand_clean = true;  // a union of filter outputs
    return GPE_OK;
}

int gpe_batch_scan(gpe_ctx *c) {
    if (!c) return GPE_ERR_INVALID;
    if (!c->have_table) return c->fail(GPE_ERR_INVALID, "gpe_build_table first");
    GPE_CUDA(c, cudaSetDevice(c->device));
    return run_scan(c);
}

int gpe_batch_bitmap(gpe_ctx *c, void **d_bitmap, uint64_t *n_bytes) {
    if (!c) return GPE_ERR_INVALID;
    if (!c->b_scanned && !c->b_filtered) return c->fail(GPE_ERR_INVALID, "gpe_batch_scan first");
    if (d_bitmap) *d_bitmap = c->d_bitmap.p;
    if (n_bytes) *n_bytes = (u64)c->b_slots * c->b_words * sizeof(u32);
    return GPE_OK;
}

int gpe_batch_bitmap_merge(gpe_ctx *c, uint32_t world, const void *d_all) {
    if (!c || !d_all || world == 0) return c ? c->fail(GPE_ERR_INVALID, "null argument") : GPE_ERR_INVALID;
    if (!c->b_scanned && !c->b_filtered) return c->fail(GPE_ERR_INVALID, "gpe_batch_scan first");
    GPE_CUDA(c, cudaSetDevice(c->device));
    int rc = compact_candidates(c, (const u32 *)d_all, world);
    if (rc) return rc;
    c->b_filtered = true;
    c->b_cand_external = true;
    c->b_cand_clean = true;  // a union of filter outputs
    return GPE_OK;
}

int gpe_batch_get_candidates(gpe_ctx *c, uint64_t *cand_offsets, uint32_t *cand) try {
    if (!c || !cand_offsets) return GPE_ERR_INVALID;
    if (!c->b_filtered) return c->fail(GPE_ERR_INVALID, "no filter result");
    GPE_CUDA(c, cudaSetDevice(c->device));
    if (int rc = settle_candidates(c)) return rc;
    GPE_CUDA(c, cudaMemcpy(cand_offsets, c->d_cand_off.p, ((size_t)c->b_slots + 1) * sizeof(u64), cudaMemcpyDeviceToHost));
    if (cand) return download_candidates(c, cand);
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_batch_get_candidates", ex.what()) : GPE_ERR_INVALID;
}

int gpe_batch_get_plan(gpe_ctx *c, uint32_t *order, uint32_t *pivot) {
    if (!c) return GPE_ERR_INVALID;
    if (!c->b_joined) return c->fail(GPE_ERR_INVALID, "gpe_batch_join first");
    GPE_CUDA(c, cudaSetDevice(c->device));
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));
    if (order) GPE_CUDA(c, cudaMemcpy(order, c->d_order.p, (size_t)c->b_slots * sizeof(u32), cudaMemcpyDeviceToHost));
    if (pivot) GPE_CUDA(c, cudaMemcpy(pivot, c->d_pivot.p, (size_t)c->b_slots * sizeof(u32), cudaMemcpyDeviceToHost));
    return GPE_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// GNN-PGE variant of seam S2 (see k4_pge.cu for its status).  Shares candidate bitmaps, compaction, matching order and
// join with the path filter; only the table (one row per data vertex) and the scan differ.
int gpe_pge_build(gpe_ctx *c, uint32_t pl, const double *x) try {
    if (!c || !x) return c ? c->fail(GPE_ERR_INVALID, "null argument") : GPE_ERR_INVALID;
    if (!c->have_graph || !c->have_emb) return c->fail(GPE_ERR_INVALID, "gpe_set_graph and gpe_set_embeddings first");
    if (pl < 1 || pl > (u32)kMaxL) return c->fail(GPE_ERR_UNSUPPORTED, "GNN-PGE path groups are built for 1..%d vertices per path", kMaxL);
    GPE_CUDA(c, cudaSetDevice(c->device));
    const u32 pde = pl * c->e;
    if (!k4_pge_supported(pde)) return c->fail(GPE_ERR_UNSUPPORTED, "no GNN-PGE scan kernel compiled for pl*e = %u", pde);
    GPE_CUDA(c, c->d_pge_x.reserve(std::max<size_t>((size_t)c->V * c->e, 1) * sizeof(double)));
    GPE_CUDA(c, c->d_pge.reserve(k4_pge_bytes(c->V, pde)));
    if (c->V) GPE_CUDA(c, cudaMemcpyAsync(c->d_pge_x.p, x, (size_t)c->V * c->e * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GPE_CUDA(c, k4_pge_groups(graph_view(c), pl, c->d_pge_x.as<double>(), k4_pge_view(c->d_pge.p, c->V, pde), c->sm_count, c->stream));
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stats.kernel_launches++;
    c->pge_pl = pl;
    c->have_pge = true;
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_pge_build", ex.what()) : GPE_ERR_INVALID;
}

int gpe_pge_dump_groups(gpe_ctx *c, double *pg, double *plg, uint8_t *has) try {
    if (!c || !pg || !plg || !has) return GPE_ERR_INVALID;
    if (!c->have_pge) return c->fail(GPE_ERR_INVALID, "gpe_pge_build first");
    GPE_CUDA(c, cudaSetDevice(c->device));
    const u32 pde = c->pge_pl * c->e;
    DevBuf a, b, h;
    GPE_CUDA(c, a.reserve(std::max<size_t>((size_t)c->V * pde * 2, 1) * sizeof(double)));
    GPE_CUDA(c, b.reserve(std::max<size_t>((size_t)c->V * pde * 2, 1) * sizeof(double)));
    GPE_CUDA(c, h.reserve(std::max<size_t>(c->V, 1)));
    cudaError_t e = k4_pge_dump(k4_pge_view(c->d_pge.p, c->V, pde), graph_view(c), pde, a.as<double>(), b.as<double>(),
                                h.as<unsigned char>(), c->stream);
    if (e == cudaSuccess && c->V) e = cudaMemcpyAsync(pg, a.p, (size_t)c->V * pde * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && c->V) e = cudaMemcpyAsync(plg, b.p, (size_t)c->V * pde * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && c->V) e = cudaMemcpyAsync(has, h.p, c->V, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    a.release(); b.release(); h.release();
    GPE_CUDA(c, e);
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_pge_dump_groups", ex.what()) : GPE_ERR_INVALID;
}

int gpe_pge_batch_upload(gpe_ctx *c, const gpe_batch *b) try {
    if (!c || !b || !b->q_vbase || !b->q_ebase || !b->q_offsets || !b->q_labels) return c ? c->fail(GPE_ERR_INVALID, "null argument") : GPE_ERR_INVALID;
    if (!c->have_pge) return c->fail(GPE_ERR_INVALID, "gpe_pge_build first");
    GPE_CUDA(c, cudaSetDevice(c->device));
    const u32 E = c->e, pl = c->pge_pl, pde = pl * E, n_slots = b->q_vbase[b->n_queries], nl = c->n_labels;
    std::string why;
    for (u32 q = 0; q < b->n_queries; q++) {
        const u32 vb = b->q_vbase[q], nq = b->q_vbase[q + 1] - vb;
        if (!check_query(c, nq, b->q_offsets + vb + q, b->q_nbrs + b->q_ebase[q], why))
            return c->fail(GPE_ERR_INVALID, "query %u: %s", q, why.c_str());
    }
    // query records: the reference computes them inline per query (GNN-PGE/src/main.cpp:226-297)
    c->label_table.fill(b->q_labels, n_slots, E);
    std::vector<u32> q_deg(std::max<u32>(n_slots, 1)), slot_label(std::max<u32>(n_slots, 1), 0xffffffffu);
    std::vector<double> q_pg_lo((size_t)std::max<u32>(n_slots, 1) * pde), q_plg_lo(q_pg_lo.size()), q_plg_hi(q_pg_lo.size());
    std::vector<unsigned char> q_has(std::max<u32>(n_slots, 1), 0);
    for (u32 q = 0; q < b->n_queries; q++) {
        const u32 vb = b->q_vbase[q], nq = b->q_vbase[q + 1] - vb;
        const u32 *off = b->q_offsets + vb + q, *nbr = b->q_nbrs + b->q_ebase[q], *lab = b->q_labels + vb;
        std::vector<double> x((size_t)nq * E), vde((size_t)nq * E), pg((size_t)nq * 2 * pde), plg((size_t)nq * 2 * pde);
        std::vector<unsigned char> has(nq);
        gen_vde(nq, off, nbr, lab, E, x.data(), vde.data(), &c->label_table);
        pge_groups(nq, off, nbr, pl, E, x.data(), vde.data(), pg.data(), plg.data(), has.data());
        for (u32 u = 0; u < nq; u++) {
            q_deg[vb + u] = off[u + 1] - off[u];
            slot_label[vb + u] = lab[u];
            q_has[vb + u] = has[u];
            for (u32 d = 0; d < pde; d++) {
                q_pg_lo[(size_t)(vb + u) * pde + d] = pg[((size_t)u * pde + d) * 2];
                q_plg_lo[(size_t)(vb + u) * pde + d] = plg[((size_t)u * pde + d) * 2];
                q_plg_hi[(size_t)(vb + u) * pde + d] = plg[((size_t)u * pde + d) * 2 + 1];
            }
        }
    }
    // slots grouped by label; a query vertex without a path group gets no candidates
    std::vector<u32> ls_off((size_t)nl + 2, 0), slot_list(std::max<u32>(n_slots, 1));
    for (u32 s2 = 0; s2 < n_slots; s2++)
        if (q_has[s2] && slot_label[s2] < nl) ls_off[slot_label[s2] + 1]++;
    for (u32 l = 0; l < nl; l++) ls_off[l + 1] += ls_off[l];
    {
        std::vector<u32> at(ls_off.begin(), ls_off.begin() + nl);
        for (u32 s2 = 0; s2 < n_slots; s2++)
            if (q_has[s2] && slot_label[s2] < nl) slot_list[at[slot_label[s2]]++] = s2;
    }
    int rc = upload_queries(c, b->n_queries, b->q_vbase, b->q_ebase, b->q_offsets, b->q_nbrs, b->q_labels, b->limits);
    if (rc) return rc;
    c->b_flags = 0;
    c->b_slots = n_slots;
    c->b_qpaths = 0;
    c->b_qblocks = 0;
    c->b_items_unpruned = 0;
    c->b_words = ((u64)(c->max_class + 31) / 32 + kChunkWords - 1) / kChunkWords * kChunkWords;
    if (c->b_words == 0) c->b_words = kChunkWords;
    c->b_chunks_per_slot = c->b_words / kChunkWords;
    // one device block: q_deg | slot_label | ls_off | slot_list | q_pg_lo | q_plg_lo | q_plg_hi
    const size_t n1 = std::max<u32>(n_slots, 1);
    const size_t o_deg = 0, o_lab = o_deg + n1 * 4, o_off = o_lab + n1 * 4, o_list = o_off + ((size_t)nl + 2) * 4;
    const size_t o_d0 = (o_list + n1 * 4 + 15) / 16 * 16, dsz = n1 * pde * sizeof(double);
    GPE_CUDA(c, c->d_pge_q.reserve(o_d0 + 3 * dsz));
    unsigned char *base = c->d_pge_q.as<unsigned char>();
    GPE_CUDA(c, cudaMemcpyAsync(base + o_deg, q_deg.data(), n1 * 4, cudaMemcpyHostToDevice, c->stream));
    GPE_CUDA(c, cudaMemcpyAsync(base + o_lab, slot_label.data(), n1 * 4, cudaMemcpyHostToDevice, c->stream));
    GPE_CUDA(c, cudaMemcpyAsync(base + o_off, ls_off.data(), ((size_t)nl + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    GPE_CUDA(c, cudaMemcpyAsync(base + o_list, slot_list.data(), n1 * 4, cudaMemcpyHostToDevice, c->stream));
    GPE_CUDA(c, cudaMemcpyAsync(base + o_d0, q_pg_lo.data(), dsz, cudaMemcpyHostToDevice, c->stream));
    GPE_CUDA(c, cudaMemcpyAsync(base + o_d0 + dsz, q_plg_lo.data(), dsz, cudaMemcpyHostToDevice, c->stream));
    GPE_CUDA(c, cudaMemcpyAsync(base + o_d0 + 2 * dsz, q_plg_hi.data(), dsz, cudaMemcpyHostToDevice, c->stream));
    GPE_CUDA(c, c->d_slot_label.reserve(n1 * sizeof(u32)));
    GPE_CUDA(c, cudaMemcpyAsync(c->d_slot_label.p, slot_label.data(), n1 * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
    GPE_CUDA(c, c->d_bitmap.reserve(std::max<u64>((u64)n_slots * c->b_words, 1) * sizeof(u32)));
    GPE_CUDA(c, c->d_counters.reserve(8 * sizeof(u64)));
    GPE_CUDA(c, c->d_survivors.reserve(8 * sizeof(u64)));
    GPE_CUDA(c, cudaStreamSynchronize(c->stream));  // the staging vectors go out of scope
    c->stats.h2d_bytes += o_d0 + 3 * dsz;
    c->b_scanned = c->b_filtered = c->b_joined = false;
    c->b_pge = true;
    c->stats.n_slots = n_slots;
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_pge_batch_upload", ex.what()) : GPE_ERR_INVALID;
}

int gpe_pge_batch_filter(gpe_ctx *c) {
    if (!c) return GPE_ERR_INVALID;
    if (!c->have_pge || !c->b_pge) return c->fail(GPE_ERR_INVALID, "gpe_pge_batch_upload first");
    GPE_CUDA(c, cudaSetDevice(c->device));
    const u32 pde = c->pge_pl * c->e, nl = c->n_labels;
    const size_t n1 = std::max<u32>(c->b_slots, 1);
    const size_t o_deg = 0, o_lab = o_deg + n1 * 4, o_off = o_lab + n1 * 4, o_list = o_off + ((size_t)nl + 2) * 4;
    const size_t o_d0 = (o_list + n1 * 4 + 15) / 16 * 16, dsz = n1 * pde * sizeof(double);
    unsigned char *base = c->d_pge_q.as<unsigned char>();
    GPE_CUDA(c, cudaMemsetAsync(c->d_counters.p, 0, 8 * sizeof(u64), c->stream));
    GPE_CUDA(c, cudaMemsetAsync(c->d_survivors.p, 0, 8 * sizeof(u64), c->stream));
    GPE_CUDA(c, cudaMemsetAsync(c->d_bitmap.p, 0, std::max<u64>((u64)c->b_slots * c->b_words, 1) * sizeof(u32), c->stream));
    {
        StageTimer tm(c, &c->stats.last_scan_ms, kStageScan);
        GPE_CUDA(c, k4_pge_scan(k4_pge_view(c->d_pge.p, c->V, pde), c->V, pde, nl, c->d_lcoff.as<u32>(),
                                reinterpret_cast<const u32 *>(base + o_off), reinterpret_cast<const u32 *>(base + o_list),
                                reinterpret_cast<const u32 *>(base + o_deg), reinterpret_cast<const double *>(base + o_d0),
                                reinterpret_cast<const double *>(base + o_d0 + dsz),
                                reinterpret_cast<const double *>(base + o_d0 + 2 * dsz), c->d_bitmap.as<u32>(), c->b_words,
                                c->d_survivors.as<u64>(), c->d_survivors.as<u64>() + 1, c->sm_count, c->stream));
        c->stats.scan_launches++;
        c->stats.kernel_launches++;
    }
    GPE_CUDA(c, c->h_pin3.reserve(8 * sizeof(u64)));
    GPE_CUDA(c, cudaMemcpyAsync(c->h_pin3.as<u64>() + 4, c->d_survivors.as<u64>() + 1, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    int rc = compact_candidates(c);
    if (rc) return rc;
    c->b_pge_rows_pending = true;  // cand_total() turns h_pin3[4] into stats.scan_rows (rows of the classes asked for)
    c->b_filtered = true;
    c->b_cand_external = false;
    c->b_cand_clean = true;  // label equal and degree >= hold for every candidate; the bitmaps are on the device
    return GPE_OK;
}

int gpe_pge_query_batch(gpe_ctx *c, const gpe_batch *b, uint64_t *answers) try {
    if (!c || !answers) return GPE_ERR_INVALID;
    int rc = gpe_pge_batch_upload(c, b);
    if (rc) return rc;
    if ((rc = gpe_pge_batch_filter(c))) return rc;
    if ((rc = run_join(c, 0, 1, nullptr, 0))) return rc;
    if ((rc = gpe_batch_download(c, answers))) return rc;
    for (u32 q = 0; q < b->n_queries; q++) answers[q] = gpe_clamp_answer(answers[q], c->h_limits[q]);
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    return c ? c->fail(GPE_ERR_INVALID, "%s: %s", "gpe_pge_query_batch", ex.what()) : GPE_ERR_INVALID;
}


// ---------------------------------------------------------------------------------------------------------------
// Multi-GPU: the path table sharded by the reference's partitions (a path belongs to the partition of its first
// vertex, custom.h:74; GPU r of N holds partitions i % N == r), one candidate exchange per batch in place of the serial
// merge (main.cpp:166-172), the join split by start candidate.  NCCL inside the library, no PyTorch.
#define GPE_NCCL(ctx, expr)                                                                                       \
    do {                                                                                                          \
        ncclResult_t r__ = (expr);                                                                                \
        if (r__ != ncclSuccess)                                                                                   \
            return (ctx)->fail(GPE_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, nccl_api().GetErrorString(r__)); \
    } while (0)

int gpe_comm_unique_id(void *id_out) {
    if (!id_out) return GPE_ERR_INVALID;
    static_assert(sizeof(ncclUniqueId) == GPE_COMM_ID_BYTES, "ncclUniqueId size");
    NcclApi &n = nccl_api();
    if (!n.ok) { g_create_err = n.err; return GPE_ERR_UNSUPPORTED; }
    ncclUniqueId id;
    if (n.GetUniqueId(&id) != ncclSuccess) { g_create_err = "ncclGetUniqueId failed"; return GPE_ERR_CUDA; }
    memcpy(id_out, &id, sizeof id);
    return GPE_OK;
}

int gpe_comm_init(gpe_ctx *c, int rank, int world, const void *id_in) {
    if (!c || !id_in || world < 1 || rank < 0 || rank >= world) return c ? c->fail(GPE_ERR_INVALID, "bad rank/world/id") : GPE_ERR_INVALID;
    NcclApi &n = nccl_api();
    if (!n.ok) return c->fail(GPE_ERR_UNSUPPORTED, "%s", n.err.c_str());
    if (c->comm) return c->fail(GPE_ERR_INVALID, "context already has a communicator");
    GPE_CUDA(c, cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, id_in, sizeof id);
    ncclComm_t comm = nullptr;
    GPE_NCCL(c, n.CommInitRank(&comm, world, id, rank));
    c->comm = comm;
    c->comm_rank = rank;
    c->comm_world = world;
    return GPE_OK;
}

int gpe_comm_init_all(gpe_ctx **ctxs, int n_ctx) {
    if (!ctxs || n_ctx < 1 || n_ctx > kMaxDevices) return GPE_ERR_INVALID;
    NcclApi &n = nccl_api();
    if (!n.ok) return ctxs[0] ? ctxs[0]->fail(GPE_ERR_UNSUPPORTED, "%s", n.err.c_str()) : GPE_ERR_UNSUPPORTED;
    int devs[kMaxDevices];
    ncclComm_t comms[kMaxDevices];
    for (int i = 0; i < n_ctx; i++) {
        if (!ctxs[i] || ctxs[i]->comm) return GPE_ERR_INVALID;
        devs[i] = ctxs[i]->device;
    }
    GPE_NCCL(ctxs[0], n.CommInitAll(comms, n_ctx, devs));
    for (int i = 0; i < n_ctx; i++) {
        ctxs[i]->comm = comms[i];
        ctxs[i]->comm_rank = i;
        ctxs[i]->comm_world = n_ctx;
    }
    return GPE_OK;
}

int gpe_comm_destroy(gpe_ctx *c) {
    if (!c) return GPE_ERR_INVALID;
    if (c->comm) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        nccl_api().CommDestroy((ncclComm_t)c->comm);
        c->comm = nullptr;
    }
    c->comm_rank = 0;
    c->comm_world = 1;
    return GPE_OK;
}

int gpe_comm_info(gpe_ctx *c, int *rank, int *world, int *nccl_version) {
    if (!c) return GPE_ERR_INVALID;
    if (rank) *rank = c->comm_rank;
    if (world) *world = c->comm_world;
    if (nccl_version) {
        *nccl_version = 0;
        if (nccl_api().ok) nccl_api().GetVersion(nccl_version);
    }
    return GPE_OK;
}

int gpe_build_table_shard(gpe_ctx *c, uint64_t *n_table_rows) {
    if (!c) return GPE_ERR_INVALID;
    if (!c->have_enum) return c->fail(GPE_ERR_INVALID, "gpe_enumerate first");
    if (c->comm_world <= 1) return gpe_build_table(c, nullptr, n_table_rows);
    if (c->p < (u32)c->comm_world) return c->fail(GPE_ERR_INVALID, "%u partitions for %d GPUs: enumerate with p >= the number of GPUs", c->p, c->comm_world);
    std::vector<uint8_t> sel(c->p);
    for (u32 i = 0; i < c->p; i++) sel[i] = (int)(i % (u32)c->comm_world) == c->comm_rank;
    return gpe_build_table(c, sel.data(), n_table_rows);
}

namespace {
// stage 2 + 3 of one context up to the exchange / after it (everything asynchronous on the context's stream)
// The exchange has two forms (both ONE fixed-size all-gather, no size exchange, no host sync):
//   dense:  the bitmaps as they are; the union is fused into the compaction's popcount pass;
//   sparse: the non-zero words as (index, word) pairs in a buffer whose capacity follows what earlier batches needed --
//           chosen when that is at most a quarter of the dense size (large label alphabets: config 3 / 5).  A shard that
//           outgrows the capacity is noticed after the step (gpe_batch_download) and the step is redone dense.
bool sparse_choice(gpe_ctx *c, u64 &cap) {
    const u64 words = (u64)c->b_slots * c->b_words;
    cap = std::max<u64>(2 * c->sparse_seen_max, 1ull << 16);
    if (const char *e = getenv("GPE_SPARSE_CAP")) cap = std::max<u64>(strtoull(e, nullptr, 10), 1);  // tests: force the dense redo
    if (const char *e = getenv("GPE_EXCHANGE")) {
        if (!strcmp(e, "dense")) return false;
        if (!strcmp(e, "sparse")) return !c->force_dense && words < (1ull << 32);
    }
    return !c->force_dense && words < (1ull << 32) && cap * 8 + 16 <= words * 4 / 4;
}
int sharded_before_exchange(gpe_ctx *c) {  // after the scan: pack the local bitmaps when the exchange is sparse
    u64 cap = 0;
    c->b_sparse_used = sparse_choice(c, cap);
    if (!c->b_sparse_used) return GPE_OK;
    c->sparse_cap = cap;
    const u64 stride = 16 + cap * 8;
    GPE_CUDA(c, c->d_sparse.reserve(stride));
    GPE_CUDA(c, c->d_all_bitmaps.reserve(stride * c->comm_world));
    GPE_CUDA(c, k3_sparse_pack(c->d_bitmap.as<u32>(), (u64)c->b_slots * c->b_words, cap, c->d_sparse.p, c->sm_count, c->stream));
    c->stats.kernel_launches++;
    return GPE_OK;
}
int sharded_exchange_enqueue(gpe_ctx *c) {  // between ncclGroupStart/End when one thread drives several contexts
    if (c->b_sparse_used) {
        GPE_NCCL(c, nccl_api().AllGather(c->d_sparse.p, c->d_all_bitmaps.p, 16 + c->sparse_cap * 8, ncclUint8, (ncclComm_t)c->comm, c->stream));
        c->stats.exchange_bytes = (16 + c->sparse_cap * 8) * (u64)c->comm_world;
        return GPE_OK;
    }
    const u64 bytes = (u64)c->b_slots * c->b_words * sizeof(u32);
    GPE_CUDA(c, c->d_all_bitmaps.reserve(std::max<u64>(bytes * c->comm_world, 16)));
    if (bytes) GPE_NCCL(c, nccl_api().AllGather(c->d_bitmap.p, c->d_all_bitmaps.p, bytes, ncclUint8, (ncclComm_t)c->comm, c->stream));
    c->stats.exchange_bytes = bytes * (u64)c->comm_world;
    return GPE_OK;
}
int sharded_after_exchange(gpe_ctx *c) {
    int rc;
    if (c->b_sparse_used) {
        const u64 stride = 16 + c->sparse_cap * 8;
        GPE_CUDA(c, k3_sparse_merge(c->d_all_bitmaps.p, stride, (u32)c->comm_world, (u32)c->comm_rank, c->sparse_cap,
                                    c->d_bitmap.as<u32>(), c->sm_count, c->stream));
        c->stats.kernel_launches++;
        GPE_CUDA(c, c->h_pin4.reserve((size_t)kMaxDevices * 2 * sizeof(u64)));
        GPE_CUDA(c, cudaMemcpy2DAsync(c->h_pin4.p, 16, c->d_all_bitmaps.p, stride, 16, (size_t)c->comm_world, cudaMemcpyDeviceToHost, c->stream));
        rc = compact_candidates(c);
    } else {
        rc = compact_candidates(c, c->d_all_bitmaps.as<u32>(), (u32)c->comm_world);
    }
    if (rc) return rc;
    c->b_filtered = true;
    c->b_cand_external = true;
    c->b_cand_clean = true;  // a union of filter outputs
    return run_join(c, (u32)c->comm_rank, (u32)c->comm_world, nullptr, 0);
}
}  // namespace

int gpe_batch_step(gpe_ctx *c) {
    if (!c) return GPE_ERR_INVALID;
    if (!c->have_table) return c->fail(GPE_ERR_INVALID, "gpe_build_table first");
    GPE_CUDA(c, cudaSetDevice(c->device));
    if (!c->comm) {
        if (int rc = run_filter(c)) return rc;
        return run_join(c, 0, 1, nullptr, 0);
    }
    // (a communicator of one rank takes the same route: the exchange is then a copy)
    if (int rc = run_scan(c)) return rc;
    if (int rc = sharded_before_exchange(c)) return rc;
    if (int rc = sharded_exchange_enqueue(c)) return rc;
    return sharded_after_exchange(c);
}

int gpe_batch_finish(gpe_ctx *c, uint64_t *answers) {
    if (!c || !answers) return GPE_ERR_INVALID;
    if (int rc = gpe_batch_download(c, answers)) return rc;  // this shard's raw counts (repaired locally if need be)
    if (c->need_dense_redo) {  // a shard outgrew the sparse exchange buffer (every rank sees that): once more, dense
        c->force_dense = true;
        int rc = gpe_batch_step(c);
        c->force_dense = false;
        if (rc) return rc;
        if ((rc = gpe_batch_download(c, answers))) return rc;
        c->stats.exchange_redos++;
    }
    const u32 nq = c->b_nq;
    if (c->comm && nq) {  // C2: sum of the shards' counts (each at most 2^48: no wrap)
        GPE_CUDA(c, c->d_reduce.reserve((size_t)nq * sizeof(u64)));
        GPE_CUDA(c, c->h_pin2.reserve((2 * std::max<size_t>(nq, 16) + 32) * sizeof(u64)));
        u64 *pin = c->h_pin2.as<u64>();
        memcpy(pin, answers, (size_t)nq * sizeof(u64));
        GPE_CUDA(c, cudaMemcpyAsync(c->d_reduce.p, pin, (size_t)nq * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
        GPE_NCCL(c, nccl_api().AllReduce(c->d_reduce.p, c->d_reduce.p, nq, ncclUint64, ncclSum, (ncclComm_t)c->comm, c->stream));
        GPE_CUDA(c, cudaMemcpyAsync(pin, c->d_reduce.p, (size_t)nq * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
        GPE_CUDA(c, cudaStreamSynchronize(c->stream));
        memcpy(answers, pin, (size_t)nq * sizeof(u64));
        c->stats.h2d_bytes += (size_t)nq * sizeof(u64);
        c->stats.d2h_bytes += (size_t)nq * sizeof(u64);
    }
    for (u32 q = 0; q < nq; q++) answers[q] = gpe_clamp_answer(answers[q], c->h_limits[q]);
    return GPE_OK;
}

// Several batches, pipelined on one host thread: while the GPU works on batch i (everything up to the download is
// asynchronous), the host plans batch i+1 (dfs_query + gen_vde + gen_query_pde of every query, custom.h:94-119, :574-633
// -- what the reference does serially per query before it touches its index, main.cpp:142-158).  With a communicator
// the collective calls happen in the same order on every rank.
static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int gpe_query_batches(gpe_ctx *c, uint32_t n_batches, const gpe_batch *batches, uint32_t flags, uint64_t *const *answers) try {
    if (!c || (n_batches && (!batches || !answers))) return c ? c->fail(GPE_ERR_INVALID, "null argument") : GPE_ERR_INVALID;
    if (!c->have_table) return c->fail(GPE_ERR_INVALID, "gpe_build_table first");
    GPE_CUDA(c, cudaSetDevice(c->device));
    gpe_plan cur, next;
    if (n_batches)
        if (int rc = plan_batch(&batches[0], c->tv.L, c->tv.E, c->label_table, cur, c->comm_world)) return c->fail(rc, "batch 0: %s", cur.err.c_str());
    u64 h2d = 0, d2h = 0;
    double t_plan = 0, t_upload = 0, t_enqueue = 0, t_finish = 0;  // host wall clock per phase, summed over the batches
    for (uint32_t i = 0; i < n_batches; i++) {
        double t0 = now_ms();
        if (int rc = upload_planned(c, &batches[i], cur, flags)) return rc;
        double t1 = now_ms();
        if (int rc = gpe_batch_step(c)) return rc;
        double t2 = now_ms();
        if (i + 1 < n_batches) {
            next = gpe_plan();
            if (int rc = plan_batch(&batches[i + 1], c->tv.L, c->tv.E, c->label_table, next, c->comm_world)) return c->fail(rc, "batch %u: %s", i + 1, next.err.c_str());
            stage_planned(c, &batches[i + 1], next, flags);
        }
        double t3 = now_ms();
        if (int rc = gpe_batch_finish(c, answers[i])) return rc;
        t_upload += t1 - t0; t_enqueue += t2 - t1; t_plan += t3 - t2; t_finish += now_ms() - t3;
        h2d += c->stats.h2d_bytes;
        d2h += c->stats.d2h_bytes;
        std::swap(cur, next);
    }
    c->stats.h2d_bytes = h2d;  // of all the batches of this call
    c->stats.d2h_bytes = d2h;
    c->stats.host_upload_ms = (float)t_upload;
    c->stats.host_enqueue_ms = (float)t_enqueue;
    c->stats.host_plan_ms = (float)t_plan;
    c->stats.host_finish_ms = (float)t_finish;
    return GPE_OK;
} catch (const std::exception &ex) {
    return c ? c->fail(GPE_ERR_INVALID, "gpe_query_batches: %s", ex.what()) : GPE_ERR_INVALID;
}

// ---- one process, several contexts (host/main -g N): one thread enqueues every GPU's work, NCCL calls grouped --------
int gpe_multi_batch_upload(gpe_ctx **ctxs, int n, const gpe_batch *b, uint32_t flags) try {
    if (!ctxs || n < 1 || !b || !ctxs[0]) return GPE_ERR_INVALID;
    gpe_ctx *c0 = ctxs[0];
    if (!c0->have_table) return c0->fail(GPE_ERR_INVALID, "gpe_build_table first");
    gpe_plan pl;
    if (int rc = plan_batch(b, c0->tv.L, c0->tv.E, c0->label_table, pl)) return c0->fail(rc, "%s", pl.err.c_str());
    for (int i = 0; i < n; i++) {
        GPE_CUDA(ctxs[i], cudaSetDevice(ctxs[i]->device));
        if (int rc = upload_planned(ctxs[i], b, pl, flags)) return rc;
    }
    return GPE_OK;
} catch (const std::exception &ex) {
    return ctxs && ctxs[0] ? ctxs[0]->fail(GPE_ERR_INVALID, "gpe_multi_batch_upload: %s", ex.what()) : GPE_ERR_INVALID;
}

int gpe_multi_batch_step(gpe_ctx **ctxs, int n) {
    if (!ctxs || n < 1) return GPE_ERR_INVALID;
    if (n == 1) return gpe_batch_step(ctxs[0]);
    for (int i = 0; i < n; i++) {
        GPE_CUDA(ctxs[i], cudaSetDevice(ctxs[i]->device));
        if (int rc = run_scan(ctxs[i])) return rc;
        if (int rc = sharded_before_exchange(ctxs[i])) return rc;
    }
    GPE_NCCL(ctxs[0], nccl_api().GroupStart());
    for (int i = 0; i < n; i++) {
        GPE_CUDA(ctxs[i], cudaSetDevice(ctxs[i]->device));
        if (int rc = sharded_exchange_enqueue(ctxs[i])) { nccl_api().GroupEnd(); return rc; }
    }
    GPE_NCCL(ctxs[0], nccl_api().GroupEnd());
    for (int i = 0; i < n; i++) {
        GPE_CUDA(ctxs[i], cudaSetDevice(ctxs[i]->device));
        if (int rc = sharded_after_exchange(ctxs[i])) return rc;
    }
    return GPE_OK;
}

int gpe_multi_batch_finish(gpe_ctx **ctxs, int n, uint64_t *answers) try {
    if (!ctxs || n < 1 || !answers) return GPE_ERR_INVALID;
    const u32 nq = ctxs[0]->b_nq;
    std::vector<u64> part(nq);
    for (int attempt = 0; attempt < 2; attempt++) {
        for (u32 q = 0; q < nq; q++) answers[q] = 0;
        bool redo = false;
        for (int i = 0; i < n; i++) {  // the GPUs ran concurrently; the host sum replaces the all-reduce (one process)
            if (int rc = gpe_batch_download(ctxs[i], part.data())) return rc;
            redo = redo || ctxs[i]->need_dense_redo;
            for (u32 q = 0; q < nq; q++) answers[q] += part[q];
        }
        if (!redo) break;
        // a shard outgrew the sparse exchange buffer: the step once more with the dense exchange, on every GPU
        for (int i = 0; i < n; i++) ctxs[i]->force_dense = true;
        int rc = gpe_multi_batch_step(ctxs, n);
        for (int i = 0; i < n; i++) { ctxs[i]->force_dense = false; ctxs[i]->stats.exchange_redos++; }
        if (rc) return rc;
    }
    for (u32 q = 0; q < nq; q++) answers[q] = gpe_clamp_answer(answers[q], ctxs[0]->h_limits[q]);
    return GPE_OK;
} catch (const std::exception &ex) {
    return ctxs && ctxs[0] ? ctxs[0]->fail(GPE_ERR_INVALID, "gpe_multi_batch_finish: %s", ex.what()) : GPE_ERR_INVALID;
}

int gpe_multi_query_batch(gpe_ctx **ctxs, int n, const gpe_batch *b, uint32_t flags, uint64_t *answers) {
    if (int rc = gpe_multi_batch_upload(ctxs, n, b, flags)) return rc;
    if (int rc = gpe_multi_batch_step(ctxs, n)) return rc;
    return gpe_multi_batch_finish(ctxs, n, answers);
}

}  // extern "C"
