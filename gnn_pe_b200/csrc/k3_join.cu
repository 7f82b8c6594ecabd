// k3_join.cu -- candidate compaction, matching order and the backtracking join (hot path 3).
//
// Replaces main.cpp:166-172 (std::set merge), refinement (custom.h:890-932), generateGQLQueryPlan
// (:670-722), generateBN (:724-755), generateValidCandidates (:757-797) and exploreQuickSIStyle (:799-888).
//
//   * candidate bitmaps -> sorted duplicate-free lists: popcount per chunk, device-wide scan, expand
//     (a std::set iterates ascending; so does a bitmap);
//   * matching order: one thread per query, the reference's greedy rule restated on bitmasks;
//   * join: one thread per work item (a partial embedding + a candidate range), explicit-stack DFS with a
//     step budget and continuation export between bounded kernel rounds (see "the join" below).  The
//     reference's per-depth candidate buffers (valid_candidate[depth], sized by max label frequency) are
//     not materialised: a depth keeps a cursor into the pivot's label group.  Edge tests are the
//     reference's binary search in the shorter adjacency list (graph.h:215-236).  As in the reference,
//     candidate sets are only used for the start vertex and for ordering.
//
// Latency / divergence bound (L2-resident CSR gathers), no bandwidth roofline, no tensor cores.
#include <cooperative_groups.h>

#include <cstdlib>

#include "gpe_internal.h"

namespace gpe {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxNQ = GPE_MAX_QUERY_VERTICES;

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Counting arithmetic.  The reference counts one embedding at a time and stops at the answer limit, which is at most
// UINT_MAX (custom.h:846-855, main.cpp:62-69), so what has to be exact is min(total, limit) with limit < 2^32.  The
// factorised count multiplies and adds table entries instead; every such operation SATURATES at kSat = 2^62 (sums and
// products of non-negative numbers: min(x, kSat) is closed under both), so a count beyond 2^64 can never wrap to a
// small number.  The only subtractions are those of the weighted counted leaves (S_u[pivot] - N_u[prefix vertex],
// W(A) W(B) - overlap); when one of their operands is saturated the difference is unknown, the query is flagged in
// `inexact[]` and the host walks those leaves instead (run_join with qmode, gpe_api.cu).
constexpr u64 kSat = 1ull << 62;
constexpr u64 kFlushCap = 1ull << 44;   // a lane's contribution to answers[] per flush; totals below it are exact
constexpr u64 kAnswerFull = 1ull << 48; // answers[] at or beyond this are not added to any more
__device__ __forceinline__ u64 sat_mul(u64 a, u64 b) {
    const u64 lo = a * b;
    return (__umul64hi(a, b) != 0 || lo > kSat) ? kSat : lo;
}
__device__ __forceinline__ u64 sat_add(u64 a, u64 b) {  // a, b <= kSat
    const u64 s = a + b;
    return s > kSat ? kSat : s;
}
__device__ __forceinline__ void flush_answer(u64 *answers, u32 q, u64 acc) {
    if (*(volatile u64 *)&answers[q] < kAnswerFull)
        atomicAdd((unsigned long long *)&answers[q], (unsigned long long)(acc < kFlushCap ? acc : kFlushCap));
}

// ---- bitmap -> per-chunk popcounts ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k3_chunk_count_kernel(const u32 *__restrict__ bitmap, u64 words_per_slot,
                                                             u64 chunks_per_slot, u64 n_chunks,
                                                             u64 *__restrict__ chunk_cnt) {
    const int lane = threadIdx.x & 31;
    u64 w = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_chunks) {
        if (w == n_chunks && lane == 0) chunk_cnt[n_chunks] = 0;  // sentinel so the scan yields the total
        return;
    }
    u64 slot = w / chunks_per_slot, c = w % chunks_per_slot;
    const u32 *p = bitmap + slot * words_per_slot + c * kChunkWords;
    u32 n = 0;
#pragma unroll
    for (int i = 0; i < (int)kChunkWords / 32; i++) n += __popc(p[i * 32 + lane]);
#pragma unroll
    for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(kFull, n, o);
    if (lane == 0) chunk_cnt[w] = n;
}

// Multi-GPU merge fused with the popcount pass: word = OR over the `world` shards' bitmaps (all-gathered, one after
// the other, `shard_words` apart), stored into this GPU's bitmap and counted per chunk.  A path belongs to the shard of
// its FIRST vertex (custom.h:74) but sets bits for all its vertices, so the shards' sets overlap: the union is an OR.
__global__ void __launch_bounds__(256) k3_merge_count_kernel(const u32 *__restrict__ all, u64 shard_words, u32 world,
                                                             u32 *__restrict__ bitmap, u64 words_per_slot,
                                                             u64 chunks_per_slot, u64 n_chunks,
                                                             u64 *__restrict__ chunk_cnt) {
    const int lane = threadIdx.x & 31;
    u64 w = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_chunks) {
        if (w == n_chunks && lane == 0) chunk_cnt[n_chunks] = 0;
        return;
    }
    const u64 slot = w / chunks_per_slot, c = w % chunks_per_slot;
    const u64 at = slot * words_per_slot + c * kChunkWords;
    u32 n = 0;
#pragma unroll
    for (int i = 0; i < (int)kChunkWords / 32; i++) {
        u32 word = 0;
        for (u32 r = 0; r < world; r++) word |= __ldcs(all + (u64)r * shard_words + at + i * 32 + lane);
        bitmap[at + i * 32 + lane] = word;
        n += __popc(word);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(kFull, n, o);
    if (lane == 0) chunk_cnt[w] = n;
}

__global__ void __launch_bounds__(256) k3_compact_kernel(const u32 *__restrict__ bitmap, u64 words_per_slot,
                                                         u64 chunks_per_slot, u64 n_chunks, u32 n_slots,
                                                         const u64 *__restrict__ chunk_off,
                                                         const u32 *__restrict__ slot_label,
                                                         const u32 *__restrict__ lcoff, u32 n_labels,
                                                         u32 *__restrict__ cand, u64 *__restrict__ cand_off, u64 cap,
                                                         u64 *overflow) {
    const int lane = threadIdx.x & 31;
    u64 w = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_chunks) return;
    u64 slot = w / chunks_per_slot, c = w % chunks_per_slot;
    u64 out = chunk_off[w];
    if (c == 0 && lane == 0) {
        cand_off[slot] = out;
        if (slot == 0) cand_off[n_slots] = chunk_off[n_chunks];
    }
    if (chunk_off[w + 1] == out) return;
    if (chunk_off[w + 1] > cap) {  // the list buffer was sized before the total was known: the host redoes this step
        if (lane == 0) *overflow = 1;
        return;
    }
    const u32 *p = bitmap + slot * words_per_slot + c * kChunkWords;
    u32 vbase = (u32)(c * kChunkWords * 32);
    // bit -> vertex in class order: first id of the slot's label + the position (ascending position = ascending id)
    const u32 sl = slot_label[slot];
    const u32 cls = sl < n_labels ? lcoff[sl] : 0;
    for (int i = 0; i < (int)kChunkWords / 32; i++) {
        u32 word = p[i * 32 + lane];
        u32 n = __popc(word), inc = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 tt = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += tt;
        }
        u64 my = out + inc - n;
        u32 v0 = vbase + (i * 32 + lane) * 32;
        while (word) {
            int b = __ffs(word) - 1;
            word &= word - 1;
            cand[my++] = cls + v0 + b;
        }
        out += __shfl_sync(kFull, inc, 31);
    }
}

// ---- sparse form of the candidate exchange ---------------------------------------------------------------------
// On large label alphabets the class-local bitmaps are almost empty (config 5: 10^4 slots x 25 KB, a few hundred
// candidates each), and all-gathering them dense would dominate the step.  A shard packs its non-zero words as
// (word index, word) pairs into a buffer of fixed capacity -- the all-gather stays ONE fixed-size collective without a
// size exchange -- and every GPU ORs the other shards' pairs into its own bitmaps.  header[0] counts the non-zero
// words; more than the capacity means the step is redone with the dense exchange (checked on the host after the step).
__global__ void __launch_bounds__(256) k3_sparse_pack_kernel(const u32 *__restrict__ bitmap, u64 n_words, u64 cap,
                                                             unsigned long long *header, uint2 *pairs) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const u64 n_round = (n_words + 31) / 32 * 32;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += (u64)gridDim.x * blockDim.x) {
        const u32 w = i < n_words ? __ldcs(bitmap + i) : 0u;
        const unsigned m = __ballot_sync(kFull, w != 0);
        if (!m) continue;
        u64 base = 0;
        if (lane == 0) base = atomicAdd(header, (unsigned long long)__popc(m));
        base = __shfl_sync(kFull, base, 0) + __popc(m & lt);
        if (w && base < cap) pairs[base] = make_uint2((u32)i, w);
    }
}

// all: `world` exchange buffers one after the other (stride_bytes apart): u64 count | u64 | cap x (index, word)
__global__ void __launch_bounds__(256) k3_sparse_merge_kernel(const unsigned char *__restrict__ all, u64 stride_bytes,
                                                              u32 world, u32 my_rank, u64 cap, u32 *bitmap) {
    for (u32 r = blockIdx.y; r < world; r += gridDim.y) {
        if (r == my_rank) continue;  // my own bits are already in place
        const unsigned char *buf = all + (u64)r * stride_bytes;
        const u64 n = min(*reinterpret_cast<const u64 *>(buf), cap);
        const uint2 *pairs = reinterpret_cast<const uint2 *>(buf + 16);
        for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
            const uint2 pw = pairs[i];
            atomicOr(bitmap + pw.x, pw.y);
        }
    }
}

// ---- union of several shards' candidate lists into the bitmaps (multi-GPU merge) ---------------------------
__global__ void k3_scatter_prefix_kernel(const u32 *__restrict__ counts, u32 world, u32 n_slots, u64 *prefix) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= world) return;
    u64 run = 0;
    for (u32 s = 0; s < n_slots; s++) {
        prefix[(u64)r * n_slots + s] = run;
        run += counts[(u64)r * n_slots + s];
    }
}

__global__ void __launch_bounds__(256) k3_scatter_kernel(const u32 *__restrict__ counts, const u32 *__restrict__ cand,
                                                         const u64 *__restrict__ prefix, u64 stride, u32 world,
                                                         u32 n_slots, const u32 *__restrict__ slot_label,
                                                         const u32 *__restrict__ lcoff, u32 n_labels, u32 *bitmap,
                                                         u64 words_per_slot) {
    for (u32 job = blockIdx.x; job < world * n_slots; job += gridDim.x) {
        u32 r = job / n_slots, slot = job % n_slots;
        u32 n = counts[(u64)r * n_slots + slot];
        const u32 *list = cand + (u64)r * stride + prefix[(u64)r * n_slots + slot];
        const u32 sl = slot_label[slot];
        const u32 first = sl < n_labels ? lcoff[sl] : 0;
        for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
            u32 v = list[i] - first;  // candidates of a slot all carry the slot's label: the bit is the class position
            atomicOr(bitmap + (u64)slot * words_per_slot + (v >> 5), 1u << (v & 31));
        }
    }
}

__global__ void k3_counts_kernel(const u64 *__restrict__ cand_off, u32 n_slots, u32 *counts) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_slots) counts[i] = (u32)(cand_off[i + 1] - cand_off[i]);
}

// ---- matching order: generateGQLQueryPlan (custom.h:670-722) + generateBN (:724-755) -----------------------
__global__ void __launch_bounds__(256) k3_order_kernel(u32 n_queries, u32 V, const u32 *__restrict__ q_vbase,
                                                       const u32 *__restrict__ q_ebase,
                                                       const u32 *__restrict__ q_offsets,
                                                       const u32 *__restrict__ q_nbrs, const u32 *__restrict__ q_labels,
                                                       const u64 *__restrict__ cand_off, u32 *order, u32 *pivot,
                                                       JoinDepth *jplan, uint2 *kids, u64 *item_base, u32 rank,
                                                       u32 world, u32 per_ticket, bool enumerate, bool clean_start,
                                                       u32 n_labels, const u32 *__restrict__ lcoff, TreeJob *tjobs,
                                                       u32 *tchild, u64 *tcursor, u32 *tcount, u32 *tlist,
                                                       u32 n_slots /*stride of tlist: 2 x slots*/, u32 sjob_base /*slots*/,
                                                       bool allow_weighted_all, const u32 *__restrict__ qmode,
                                                       float branching, u64 pool_cap) {
    // One query per WARP, lane 0 only: the plan is serial, branchy code -- 32 different queries in the lanes of one warp
    // execute it 32 times over (the r01y launch list had this kernel at 94 us for 100 queries in a single CTA) -- and
    // spread over the SMs 1000 queries take as long as 100.
    const u32 q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q < n_queries && (threadIdx.x & 31) == 0) {
        // qmode (second pass of a batch some of whose weighted counts saturated, see kSat): 0 as usual, 1 plan this query
        // without weighted counted leaves, 2 leave this query out (no tickets, no tables)
        const u32 mode = qmode ? qmode[q] : 0;
        const bool allow_weighted = allow_weighted_all && mode == 0;
        const u32 vb = q_vbase[q], nq = mode == 2 ? 0 : q_vbase[q + 1] - vb;
        const u32 *off = q_offsets + vb + q;  // nq + 1 local offsets
        const u32 *nbr = q_nbrs + q_ebase[q];
        const u32 *qlab = q_labels + vb;
        const u64 *co = cand_off + vb;
        u32 *ord = order + vb, *piv = pivot + vb;
        // per-vertex candidate counts, degrees and adjacency masks once, up front (independent loads); the selection loops
        // below ask for them O(nq^2) times
        u32 cnt_[kMaxNQ], qd_[kMaxNQ];
        u64 adjm[kMaxNQ];
        for (u32 u = 0; u < nq; u++) {
            cnt_[u] = (u32)(co[u + 1] - co[u]);
            qd_[u] = off[u + 1] - off[u];
            u64 m = 0;
            for (u32 j = off[u]; j < off[u + 1]; j++) m |= 1ull << nbr[j];
            adjm[u] = m;
        }
        auto count = [&](u32 u) { return cnt_[u]; };
        auto qdeg = [&](u32 u) { return qd_[u]; };
        u64 items = 0;
        if (nq > 0) {
            // ---- the reference's plan (reported through gpe_refine / gpe_batch_get_plan) ----
            u32 start = 0;  // selectGQLStartVertex, custom.h:635-654
            for (u32 i = 1; i < nq; i++) {
                if (count(i) < count(start)) start = i;
                else if (count(i) == count(start) && qdeg(i) > qdeg(start)) start = i;
            }
            u64 visited = 0, adjacent = 0;
            auto mark = [&](u32 u) {
                visited |= 1ull << u;
                adjacent |= adjm[u];
            };
            ord[0] = start;
            piv[0] = 0xffffffffu;
            mark(start);
            for (u32 i = 1; i < nq; i++) {
                u32 next = 0, best = V + 1;
                for (u32 u = 0; u < nq; u++) {
                    if ((visited >> u & 1) || !(adjacent >> u & 1)) continue;
                    if (count(u) < best) { best = count(u); next = u; }
                    else if (count(u) == best && qdeg(u) > qdeg(next)) next = u;
                }
                mark(next);
                ord[i] = next;
                u32 pv = 0xffffffffu;
                for (u32 j = 0; j < i; j++)
                    if (adjm[next] >> ord[j] & 1) { pv = ord[j]; break; }
                piv[i] = pv;
            }

            // ---- the execution plan ---------------------------------------------------------------------------
            // Counting mode: pendant subtrees whose vertices all carry labels that are unique in the query cannot
            // collide with anything, so the number of ways to complete one below a matched vertex x is a function
            // of x alone.  They are peeled off here (leaves first), tabulated per data vertex by k3_tree_tables and
            // enter the walk as one factor at their attachment vertex.  What remains (the core: cycles, repeated
            // labels and whatever connects them) is walked; its own query leaves are counted as before.
            // Enumeration mode (matches wanted): nothing is peeled, the root is the start vertex.
            u64 alive = nq >= 64 ? ~0ull : (1ull << nq) - 1;
            u64 uniq = 0;
            u32 par[kMaxNQ], remdeg[kMaxNQ], peel[kMaxNQ], n_peel = 0;
            for (u32 u = 0; u < nq; u++) {
                remdeg[u] = qdeg(u);
                par[u] = 0xffffffffu;
                bool un = true;
                for (u32 v = 0; v < nq; v++) un = un && (v == u || qlab[v] != qlab[u]);
                if (un) uniq |= 1ull << u;
            }
            // (a start candidate of the wrong label -- possible with caller-supplied sets, which the reference takes as
            //  they are -- could collide with a peeled vertex: no peeling then)
            if (!enumerate && clean_start) {
                for (int pass = 0; pass < 2; pass++) {  // the start vertex goes last
                    bool changed = true;
                    while (changed) {
                        changed = false;
                        for (u32 u = 0; u < nq; u++) {
                            if (!(alive >> u & 1) || remdeg[u] != 1 || !(uniq >> u & 1)) continue;
                            if (u == start && pass == 0) continue;
                            u32 p = 0;
                            for (u32 j = off[u]; j < off[u + 1]; j++)
                                if (alive >> nbr[j] & 1) p = nbr[j];
                            par[u] = p;
                            alive &= ~(1ull << u);
                            remdeg[p]--;
                            peel[n_peel++] = u;
                            changed = true;
                        }
                    }
                }
            }
            // Does tabulating pay?  A table costs one entry per vertex of a whole label class; walking the peeled vertices
            // instead costs about roots x (b + b^2 + ...) steps, b = the graph's growth factor per depth.  With a selective
            // filter on a large label alphabet (BASELINE.json configs 3 / 5: a few hundred start candidates, classes of
            // 200 K vertices, b = 0.4) the walk is thousands of times cheaper; on a power-law graph (b > 1) it explodes.
            // Tables are dropped only when the walk is estimated at least 16 times cheaper.
            u64 pool_at = 0;  // this query's slice of the table pool (sub-allocated below without atomics)
            if (n_peel) {
                double table_cost = 0.0;
                for (u32 u = 0; u < nq; u++) {
                    bool has_child = false;
                    for (u32 k = 0; k < n_peel; k++) has_child = has_child || par[peel[k]] == u;
                    if (has_child && qlab[u] < n_labels) table_cost += (double)(lcoff[qlab[u] + 1] - lcoff[qlab[u]]);
                }
                double g = 0.0, bp = 1.0;
                for (u32 i = 1; i < nq; i++) { bp *= (double)branching; g += bp; if (g > 1e18) break; }
                const double walk_cost = (double)count(start) * (1.0 + g);
                // room in the table pool for everything this query may tabulate, reserved with ONE atomic: a table per
                // vertex with peeled children plus, at most, a sum table per core leaf that carries peeled subtrees (sized by
                // the largest class a pivot could have).  No room (the pool is capped) -> this query walks instead.
                u64 need = 0, biggest = 0;
                u32 n_sum = 0;
                for (u32 u = 0; u < nq; u++) {
                    const u64 cls = qlab[u] < n_labels ? lcoff[qlab[u] + 1] - lcoff[qlab[u]] : 0;
                    bool has_child = false;
                    for (u32 k = 0; k < n_peel; k++) has_child = has_child || par[peel[k]] == u;
                    if (has_child) need += cls;
                    if (alive >> u & 1) {
                        biggest = max(biggest, cls);
                        if (remdeg[u] == 1 && qdeg(u) != 1) n_sum++;
                    }
                }
                need += (u64)n_sum * biggest;
                bool keep = !(walk_cost * 16.0 < table_cost);
                if (keep) {
                    pool_at = atomicAdd((unsigned long long *)tcursor, (unsigned long long)need);
                    keep = pool_at + need <= pool_cap;
                }
                if (!keep) {
                    n_peel = 0;
                    alive = nq >= 64 ? ~0ull : (1ull << nq) - 1;
                    for (u32 u = 0; u < nq; u++) { remdeg[u] = qdeg(u); par[u] = 0xffffffffu; }
                }
            }
            const u32 nK = nq - n_peel;
            const bool root_is_start = alive >> start & 1;
            u32 root = start;
            if (!root_is_start) {  // the walk starts at the core vertex with the smallest label class
                u32 best = 0xffffffffu;
                for (u32 u = 0; u < nq; u++) {
                    if (!(alive >> u & 1)) continue;
                    const u32 sz = qlab[u] < n_labels ? lcoff[qlab[u] + 1] - lcoff[qlab[u]] : 0;
                    if (sz < best || (sz == best && qdeg(u) > qdeg(root))) { best = sz; root = u; }
                }
            }
            // Counted leaves: leaves of the CORE whose only core neighbour (their pivot) is walked.  A leaf that is also a
            // leaf of the query contributes the number of free members of its pivot's label group; a leaf that carries
            // peeled subtrees (it has a repeated label, so it could not be peeled itself) contributes the same sum
            // WEIGHTED by its own table N_u: sum over free members y of N_u[y] = S_u[pivot image] - (N_u of the prefix
            // vertices inside the group), with S_u tabulated like any parent-of-a-peeled-child table (k3_tree_tables).
            u64 tail_set = 0, weighted = 0;
            u32 tail_list[kMaxNQ], n_tail = 0, cpar[kMaxNQ];
            for (u32 u = 0; u < nq; u++) {
                cpar[u] = 0xffffffffu;
                if (!(alive >> u & 1) || remdeg[u] != 1) continue;
                for (u32 j = off[u]; j < off[u + 1]; j++)
                    if (alive >> nbr[j] & 1) cpar[u] = nbr[j];
            }
            for (u32 u = 0; u < nq && nK >= 3 && n_tail + 2 < nK; u++) {
                if (u == root || !(alive >> u & 1) || remdeg[u] != 1) continue;
                const bool wu = qdeg(u) != 1;  // carries peeled subtrees
                if (wu && !allow_weighted) continue;
                const u32 pv = cpar[u];
                u32 r = 0, p0 = 0;
                bool mixed = false, any_w = wu;
                for (u32 k = 0; k < n_tail; k++) {
                    const u32 t = tail_list[k];
                    if (qlab[t] != qlab[u]) continue;
                    const u32 pt = cpar[t];
                    if (r == 0) p0 = pt; else if (pt != p0) mixed = true;
                    any_w = any_w || (weighted >> t & 1);
                    r++;
                }
                // a label may appear on one tail leaf, on several plain leaves of ONE pivot, or on two leaves of two pivots
                if (r == 0 || (!mixed && pv == p0 && !any_w) || (r == 1 && pv != p0)) {
                    tail_list[n_tail++] = u;
                    tail_set |= 1ull << u;
                    if (wu) weighted |= 1ull << u;
                }
            }
            const u32 n_walk = nK - n_tail;
            u32 xo[kMaxNQ], depth_of[kMaxNQ], lab[kMaxNQ], top[kMaxNQ];
            xo[0] = root;
            visited = 0;
            adjacent = 0;
            mark(root);
            for (u32 i = 1; i < n_walk; i++) {
                u32 next = 0, best = V + 1;
                for (u32 u = 0; u < nq; u++) {
                    if ((visited >> u & 1) || !(adjacent >> u & 1) || (tail_set >> u & 1) || !(alive >> u & 1)) continue;
                    if (count(u) < best) { best = count(u); next = u; }
                    else if (count(u) == best && qdeg(u) > qdeg(next)) next = u;
                }
                mark(next);
                xo[i] = next;
            }
            u32 n_exec = n_walk;
            {   // tail depths, grouped by label
                u64 placed = 0;
                u32 at = n_walk;
                for (u32 k = 0; k < n_tail; k++) {
                    const u32 t = tail_list[k];
                    if (placed >> t & 1) continue;
                    placed |= 1ull << t;
                    const u32 first_at = at;
                    xo[at] = t;
                    top[at++] = kTailMul;
                    for (u32 k2 = k + 1; k2 < n_tail; k2++) {
                        const u32 t2 = tail_list[k2];
                        if ((placed >> t2 & 1) || qlab[t2] != qlab[t]) continue;
                        placed |= 1ull << t2;
                        xo[at] = t2;
                        if (cpar[t2] == cpar[t]) {
                            top[at++] = kTailFall;
                        } else {
                            top[first_at] = kTailPairA;
                            top[at++] = kTailPairB;
                        }
                    }
                }
                n_exec = at;
            }
            // peeled subtrees: every vertex with peeled children gets a table over its own label class (k3_tree_tables);
            // the core vertices' tables are the factors of the walk
            u32 tlevel[kMaxNQ];
            {
                u32 at = 0;
                for (u32 u = 0; u < nq; u++) {  // peeled children of u, contiguous in tchild
                    const u32 first = at;
                    for (u32 k = 0; k < n_peel; k++)
                        if (par[peel[k]] == u) tchild[vb + at++] = vb + peel[k];
                    TreeJob &tj = tjobs[vb + u];
                    tj.child_begin = vb + first;
                    tj.n_child = at - first;
                    tj.level = 0;
                    tj.label = qlab[u];
                    tj.qdeg = qdeg(u);
                    tj.start_slot = (u == start && !(alive >> u & 1)) ? vb + start : 0xffffffffu;
                    tj.table_off = 0;
                    tlevel[u] = 0;
                }
            }
            auto make_table = [&](u32 v) {  // children are final (peel order: children before parents)
                TreeJob &tj = tjobs[vb + v];
                if (!tj.n_child) return;
                u32 lvl = 1;
                for (u32 k = 0; k < n_peel; k++)
                    if (par[peel[k]] == v && tlevel[peel[k]] + 1 > lvl) lvl = tlevel[peel[k]] + 1;
                tlevel[v] = lvl;
                tj.level = lvl;
                const u32 sz = qlab[v] < n_labels ? lcoff[qlab[v] + 1] - lcoff[qlab[v]] : 0;
                tj.table_off = pool_at;
                pool_at += sz;
                tlist[(u64)lvl * n_slots + atomicAdd(&tcount[lvl], 1u)] = vb + v;  // this level's launch works on it
            };
            for (u32 k = 0; k < n_peel; k++) make_table(peel[k]);
            for (u32 u = 0; u < nq; u++)
                if (alive >> u & 1) make_table(u);
            // weighted counted leaves: S_u over the pivot's label class, S_u[x] = sum over y in N(x) of u's label and
            // degree of N_u[y] -- the table of a job whose only child is u (second half of the job / child arrays)
            u64 stab[kMaxNQ];
            for (u32 k = 0; k < n_tail; k++) {
                const u32 u = tail_list[k];
                stab[u] = 0;
                if (!(weighted >> u & 1)) continue;
                const u32 pv = cpar[u], ji = sjob_base + vb + u;
                tchild[ji] = vb + u;
                TreeJob &sj = tjobs[ji];
                sj.child_begin = ji;
                sj.n_child = 1;
                sj.level = tlevel[u] + 1;
                sj.label = qlab[pv];
                sj.qdeg = qdeg(pv);
                sj.start_slot = 0xffffffffu;
                const u32 sz = qlab[pv] < n_labels ? lcoff[qlab[pv] + 1] - lcoff[qlab[pv]] : 0;
                sj.table_off = pool_at;
                pool_at += sz;
                tlist[(u64)sj.level * n_slots + atomicAdd(&tcount[sj.level], 1u)] = ji;
                stab[u] = sj.table_off + V - (qlab[pv] < n_labels ? lcoff[qlab[pv]] : 0);  // read as tpool[.. + v' of the pivot's image]
            }
            for (u32 u = 0; u < nq; u++) depth_of[u] = 0xffffffffu;
            for (u32 i = 0; i < n_exec; i++) { depth_of[xo[i]] = i; lab[i] = qlab[xo[i]]; }
            for (u32 i = 0; i < n_exec; i++) {
                const u32 u = xo[i];
                JoinDepth jd;
                jd.u = u;
                jd.label = lab[i];
                jd.deg = qdeg(u);
                jd.pivot_depth = 0;
                jd.bn_mask = 0;
                // (tables are indexed by class position; the walk holds class-order ids v' = lcoff[label] + position and reads
                //  tpool[tree_off + v'] from a pool pointer shifted down by V, so the offset stays non-negative)
                jd.tree_off = tjobs[vb + u].level ? tjobs[vb + u].table_off + V - (qlab[u] < n_labels ? lcoff[qlab[u]] : 0) : kNoTree;
                jd.tail_mask = 0;
                jd.tail_k = 0;
                jd.sure_used = 0;
                jd.kid_begin = 0;
                jd.kid_count = 0;
                jd.units_mask = 0;
                if (i > 0) {
                    u32 pd = 0xffffffffu;  // pivot: the earliest-matched query neighbour
                    for (u32 j = off[u]; j < off[u + 1]; j++) {
                        const u32 dw = depth_of[nbr[j]];
                        if (dw < i && dw < pd) pd = dw;
                    }
                    jd.pivot_depth = pd == 0xffffffffu ? 0 : pd;
                    for (u32 j = off[u]; j < off[u + 1]; j++) {
                        const u32 dw = depth_of[nbr[j]];
                        if (dw < i && dw != jd.pivot_depth) jd.bn_mask |= 1ull << dw;
                    }
                }
                if (i > 0 && i < n_walk) {
                    // [walked depths] earlier depths a candidate of this one could coincide with: only those of the same
                    // label (candidates come from a label group); depth 0 too when its data label is only known at run time
                    u64 sm = (root_is_start && !clean_start) ? 1ull : 0ull;
                    for (u32 t = 0; t < i; t++)
                        if (lab[t] == lab[i]) sm |= 1ull << t;
                    jd.tail_mask = sm;
                }
                if (i >= n_walk) {
                    const bool wu = weighted >> u & 1;
                    jd.tail_k = top[i] | (wu ? kTailW : 0u);
                    const u32 pvu = xo[jd.pivot_depth];
                    u64 sure_mask = 0;
                    for (u32 t = 0; t < n_walk; t++) {
                        if (t == jd.pivot_depth) continue;
                        if (t == 0 && root_is_start && !clean_start) { jd.tail_mask |= 1ull; continue; }  // its data label is only known at run time
                        if (lab[t] != jd.label) continue;
                        if (adjm[xo[t]] >> pvu & 1) { jd.sure_used++; sure_mask |= 1ull << t; } else jd.tail_mask |= 1ull << t;
                    }
                    if (wu) {  // (a core leaf has no backward neighbour besides its pivot: both words are free)
                        jd.bn_mask = sure_mask;   // WHICH prefix vertices surely sit in the group: their weights are subtracted
                        jd.units_mask = stab[u];  // table S_u (pool offset)
                    }
                }
                jplan[vb + i] = jd;
            }
            jplan[vb].tail_k = n_exec - n_walk;              // [depth 0] number of tail depths
            jplan[vb].sure_used = n_exec;                    // [depth 0] depths of the execution plan
            jplan[vb].tail_mask = root_is_start ? 0 : 1;     // [depth 0] the walk starts from: 0 the start vertex's candidate list, 1 the root's label class
            {   // depths that draw their candidates from the vertex matched at depth t, as (depth, label) lists
                u32 at = 0;
                for (u32 t = 0; t < n_exec; t++) {
                    jplan[vb + t].kid_begin = at;
                    for (u32 i = t + 1; i < n_exec; i++)
                        if (jplan[vb + i].pivot_depth == t) kids[vb + at++] = make_uint2(i, jplan[vb + i].label);
                    jplan[vb + t].kid_count = at - jplan[vb + t].kid_begin;
                }
            }
            // counted-tail units (a Mul leaf with its Fall followers, or a PairA/PairB couple) are
            // evaluated at the shallowest depth at which every vertex they depend on is matched
            for (u32 i = n_walk; i < n_exec; i++) {
                const u32 op = top[i];
                if (op == kTailFall || op == kTailPairB) continue;
                auto msb = [](u64 m) { u32 r = 0; while (m >>= 1) r++; return r; };
                u32 dep = jplan[vb + i].pivot_depth;
                if (jplan[vb + i].tail_mask) dep = max(dep, msb(jplan[vb + i].tail_mask));
                // (a weighted leaf subtracts the weights of the prefix vertices that surely sit in its group: they must
                //  be matched by then; a plain leaf only counts them)
                if ((jplan[vb + i].tail_k & kTailW) && jplan[vb + i].bn_mask) dep = max(dep, msb(jplan[vb + i].bn_mask));
                if (op == kTailPairA) {
                    dep = max(dep, jplan[vb + i + 1].pivot_depth);
                    if (jplan[vb + i + 1].tail_mask) dep = max(dep, msb(jplan[vb + i + 1].tail_mask));
                    if ((jplan[vb + i + 1].tail_k & kTailW) && jplan[vb + i + 1].bn_mask)
                        dep = max(dep, msb(jplan[vb + i + 1].bn_mask));
                    for (u32 t = 1; t < n_walk; t++)
                        if (lab[t] == lab[i]) dep = max(dep, t);  // its vertex may sit in both groups
                }
                jplan[vb + dep].units_mask |= 1ull << i;
            }
            u64 total;
            if (root_is_start) total = count(start);
            else total = qlab[root] < n_labels ? lcoff[qlab[root] + 1] - lcoff[qlab[root]] : 0;
            items = total > rank ? (total - rank + world - 1) / world : 0;
        }
        item_base[q + 1] = (items + per_ticket - 1) / per_ticket;  // tickets of this query; turned into a prefix by the next kernel
    }
}

// item_base[q + 1] (tickets of query q) -> inclusive prefix, item_base[0] = 0; one warp, chunks of 32 queries
__global__ void k3_order_prefix_kernel(u32 n_queries, u64 *item_base) {
    const int lane = threadIdx.x;
    u64 run = 0;
    for (u32 q0 = 0; q0 < n_queries; q0 += 32) {
        const u32 q = q0 + lane;
        u64 v = q < n_queries ? item_base[q + 1] : 0, inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u64 t = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += t;
        }
        if (q < n_queries) item_base[q + 1] = run + inc;
        run += __shfl_sync(kFull, inc, 31);
    }
    if (lane == 0) item_base[0] = 0;
}

// ---- the join ------------------------------------------------------------------------------------------------
//
// Work item = a partial embedding (the first `level` vertices of the execution order) plus a range [lo, hi) of
// the candidate segment of that level.  One THREAD runs one item as an explicit-stack DFS (stack in shared
// memory, [level][thread] so accesses never bank-conflict); a candidate segment is the label group of the
// pivot's adjacency (JoinGraph::nbrL / gtab), so a step touches only neighbours that already carry the right
// label (level 0 walks the start vertex's candidate list instead).  Label groups are short (degree / labels),
// which is why a warp per segment would idle.
//
// What a step costs is the chain of dependent loads, so everything a matched vertex c decides is looked up ONCE,
// when c is matched, and kept on the stack:
//   * the label groups of c that later depths draw their candidates from (S0/E0 of every depth pivoting on c) --
//     all in c's row of gtab, fetched together; an empty group rejects c on the spot;
//   * the counted-tail factors (below) that become computable at this depth, folded into a running product
//     PROD(level); a zero factor rejects c.
// Descending is then a copy of two stack words, and the deepest walked level only adds PROD to the count.
//
// Counting shortcut (results unchanged): the tail depths are query leaves whose only neighbour is in the walked
// prefix.  Their completions are counted, not walked: a leaf contributes the size of its pivot's label group
// minus the prefix vertices inside it; leaves with different labels multiply; r same-label leaves on one pivot
// give a falling factorial; two same-label leaves on different pivots give |A||B| - |A n B|.  A factor is
// evaluated at the shallowest depth at which every vertex it depends on is matched (JoinDepth::units_mask).
//
// One persistent launch, two levels of load balancing (subtree sizes are heavy-tailed):
//   * inside a warp: a lane without work takes the upper half of the shallowest unexplored sibling range of
//     a busy lane, straight out of the donor's shared-memory stack column (no global traffic);
//   * between warps: tickets [0, n_init) are the start candidates (JoinInit), later tickets are exported items
//     in a linear buffer.  Lanes claim tickets while published items are available; a warp without any busy lane
//     registers as idle and polls the queue header with back-off; busy warps look at the header every
//     2^kExportEveryLog2 iterations (the loads are issued one iteration ahead of their use) and, when idle warps
//     outnumber the published items, one lane hands over the unexplored siblings of its shallowest level as new
//     items.  JoinQueue::pending counts published items whose work has not been retired; a warp retires what it
//     claimed whenever all its lanes run dry, so pending == 0 means the join is complete.  Exporting is only load
//     balancing: if the buffer fills up, threads simply keep their work.
//
// Edge tests (checkEdgeExistence, graph.h:215-236) are binary searches too, but inside the label group of one
// endpoint instead of its whole adjacency list: same answer, a fraction of the dependent loads.
// (JoinGraph -- the join's copy of the data graph in class order -- is declared in gpe_internal.h and built by k0_graph.cu)

constexpr int kItemHdr = 8;  // q, level, lo, hi, prod (2 words), label of the start vertex, pad; then EMB | S0 | E0
constexpr u32 kSplit = 8;
constexpr int kExportEveryLog2 = 1;  // log2 of the rounds between two looks at the queue header (config 2, ms per batch:
                                     // every 8 rounds 15.60, 4: 15.28, 2: 15.22, 1: 15.71 -- the end of the join is a few
                                     // generations of subtree hand-overs, each waiting for a busy warp's next look)
// (round 2, measured with tools/join_split_bench.py -- the join of one rank of an N-way split on one GPU, config 2, ms for
//  N = 1 / 2 / 4 / 8: steps per round 2 -> 1: 13.9 / 8.2 / 5.4 / 4.1 -> 14.0 / 8.3 / 5.2 / 3.5; tail batch 8 -> 16 on top:
//  13.7 / 8.0 / 5.1 / 3.5; export lanes 4 -> 8: 3.4-3.6 at N = 8.  Shorter rounds shorten the hand-over generations at
//  the end of a launch, which is most of a small share's time)
constexpr int kStepsPerRound = 1;   // DFS steps of a lane between two rounds of scheduling (tickets, donation, export)
constexpr int kExportLanes = 8;  // lanes of a warp that may hand work over in one round
constexpr int kTailBatch = 16;   // parked lanes that trigger a joint evaluation of their counted-tail factors

__host__ __device__ constexpr u32 item_stride(u32 m) { return kItemHdr + 3 * m; }

// Directory rows and binary-search probes are random reads without reuse inside an SM: they go around L1
// (ld.global.cg) so that the small, hot join plan stays resident in what the stack leaves of it.
struct DirRow {  // the group directory row of one vertex
    const unsigned char *p;
    u32 base;  // narrow rows: start of the vertex's adjacency; the row holds 16-bit offsets from it
};
__device__ __forceinline__ DirRow dir_row(const JoinGraph &g, u32 v) {
    DirRow r;
    r.p = g.gtab + (u64)v * g.dir_row_bytes;
    r.base = g.wide_dir ? 0u : __ldcg(reinterpret_cast<const u32 *>(r.p));
    return r;
}
__device__ __forceinline__ void row_range(const JoinGraph &g, const DirRow &r, u32 label, u32 &lo, u32 &hi) {
    if (g.wide_dir) {
        lo = __ldcg(reinterpret_cast<const u32 *>(r.p) + label);
        hi = __ldcg(reinterpret_cast<const u32 *>(r.p) + label + 1);
    } else {
        const unsigned short *h = reinterpret_cast<const unsigned short *>(r.p + 4) + label;
        lo = r.base + __ldcg(h);
        hi = r.base + __ldcg(h + 1);
    }
}
__device__ __forceinline__ void group_range(const JoinGraph &g, u32 v, u32 label, u32 &lo, u32 &hi) {
    if (label >= g.nl) { lo = hi = 0; return; }
    row_range(g, dir_row(g, v), label, lo, hi);
}
// adjacency entry: neighbour (class-order id) and its degree saturated at 255 (query degrees are < 64, so every
// `degree >= query degree` test is exact)
__device__ __forceinline__ void adj_entry(const JoinGraph &g, u32 at, u32 &c, u32 &cdeg) {
    if (g.wide_adj) {
        const uint2 e = reinterpret_cast<const uint2 *>(g.nbrJ)[at];
        c = e.x;
        cdeg = e.y;
    } else {
        const u32 e = g.nbrJ[at];
        c = e & 0xffffffu;
        cdeg = e >> 24;
    }
}
__device__ __forceinline__ u32 adj_id_cg(const JoinGraph &g, u32 at) {
    return g.wide_adj ? __ldcg(g.nbrJ + 2 * (u64)at) : __ldcg(g.nbrJ + at) & 0xffffffu;
}

// Edge filter: "no" is exact, "maybe" has to be confirmed by a search.  Most membership / edge tests of the join
// fail (a prefix vertex is rarely adjacent to the pivot), and a failed test costs one 8-byte load here instead of a
// chain of binary-search probes (join_edge_probe, gpe_internal.h: both bits of an edge live in ONE 64-bit word).
template <bool CG>
__device__ __forceinline__ bool edge_maybe(const JoinGraph &g, u32 a, u32 b) {
    u64 word, bits;
    join_edge_probe(a, b, g.bloom_word_mask, word, bits);
    const u64 *p = g.bloom + word;
    return ((CG ? __ldcg(p) : __ldg(p)) & bits) == bits;  // CG: around L1 (random, no reuse), which holds the plans
}

// is v a member of the group [s, e) of the adjacency?  (ids ascending)
__device__ __forceinline__ bool in_group(const JoinGraph &g, u32 s, u32 e, u32 v) {
    while (s < e) {
        const u32 mid = s + ((e - s) >> 1);
        const u32 x = adj_id_cg(g, mid);
        if (x == v) return true;
        if (x < v) s = mid + 1; else e = mid;
    }
    return false;
}

__device__ __forceinline__ u32 ld_acquire_u32(const u32 *p) {
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(u32 *p, u32 v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void ld_relaxed_2xu64(const void *p, u64 &a, u64 &b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

// One ticket per root candidate of this shard (idx % world == rank): (query, position in cand[] or in lclass[]).
// Subtree sizes are heavy-tailed and grow with the root's degree, so high-degree roots are ticketed first (longest
// jobs first: the end of the join is then made of small items).  Ticket order: the heavy roots of query 0, 1, ...,
// then the light roots of query 0, 1, ... -- neighbouring tickets belong to one query, so the lanes of a warp mostly
// read the same plan.  Three small launches: count the heavy roots per query, prefix, place.
struct RootItem { u32 q, pos; bool heavy, alive; };

// A root that fails the tests the walk would apply to it first -- its degree (class starts only: listed start candidates
// are taken as they are, custom.h:827-830) and the subtree table of the root's query vertex, which for a peeled start
// vertex carries the candidate bitmap -- never gets a ticket: the test runs here, coalesced over the class, instead of
// costing a ticket claim and a divergent first step in the walk.
__device__ __forceinline__ RootItem root_item(u64 item, u32 n_queries, const u32 *q_vbase, const JoinDepth *jplan,
                                              const u64 *cand_off, const u32 *cand, const u32 *deg /*class order*/,
                                              const u32 *lcoff, u32 n_labels, const u64 *item_base, u32 rank, u32 world,
                                              u32 heavy_deg, const u64 *tpool /*shifted by -V*/) {
    u32 lo = 0, hi = n_queries;
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (item_base[mid] <= item) lo = mid; else hi = mid;
    }
    RootItem r;
    r.q = lo;
    const u32 vb = q_vbase[r.q];
    const u64 idx = (item - item_base[r.q]) * world + rank;
    const JoinDepth &j0 = jplan[vb];
    const u64 first = j0.tail_mask ? (j0.label < n_labels ? lcoff[j0.label] : 0) : cand_off[vb + j0.u];
    r.pos = (u32)(first + idx);
    const u32 c = j0.tail_mask ? r.pos : cand[r.pos];  // a class position IS the class-order id
    const u32 dg = deg[c];
    r.heavy = dg >= heavy_deg;
    r.alive = !j0.tail_mask || dg >= j0.deg;
    if (r.alive && tpool && j0.tree_off != kNoTree) r.alive = tpool[j0.tree_off + c] != 0;
    return r;
}

// qcur: per query [0] heavy count, [1] heavy base, [2] light base, [3] heavy cursor, [4] light cursor, [5] light count
constexpr int kQCur = 6;
__global__ void __launch_bounds__(256) k3_init_count_kernel(u32 n_queries, const u32 *__restrict__ q_vbase,
                                                            const JoinDepth *__restrict__ jplan,
                                                            const u64 *__restrict__ cand_off,
                                                            const u32 *__restrict__ cand,
                                                            const u32 *__restrict__ deg,
                                                            const u32 *__restrict__ lcoff, u32 n_labels,
                                                            const u64 *__restrict__ item_base, u32 rank, u32 world,
                                                            u32 heavy_deg, const u64 *__restrict__ tpool, u64 *qcur,
                                                            const u64 *__restrict__ cand_overflow) {
    if (cand_overflow && *cand_overflow) return;  // truncated candidate lists: the host redoes the step
    const u64 n_items = item_base[n_queries], n_round = (n_items + 31) / 32 * 32;
    const int lane = threadIdx.x & 31;
    for (u64 item = (u64)blockIdx.x * blockDim.x + threadIdx.x; item < n_round; item += (u64)gridDim.x * blockDim.x) {
        RootItem r{0xffffffffu, 0, false, false};
        if (item < n_items)
            r = root_item(item, n_queries, q_vbase, jplan, cand_off, cand, deg, lcoff, n_labels, item_base, rank, world,
                          heavy_deg, tpool);
        // one atomic per (warp, query, class): consecutive items belong to one query almost always
        const unsigned peers = __match_any_sync(kFull, r.alive ? (r.q << 1 | (r.heavy ? 1u : 0u)) : 0xffffffffu);
        if (r.alive && lane == __ffs(peers) - 1)
            atomicAdd((unsigned long long *)&qcur[kQCur * (u64)r.q + (r.heavy ? 0 : 5)], (unsigned long long)__popc(peers));
    }
}

__global__ void k3_init_prefix_kernel(u32 n_queries, u64 *qcur, JoinQueue *jq) {  // one warp
    const int lane = threadIdx.x;
    u64 run = 0;
    for (int pass = 0; pass < 2; pass++)  // heavy roots of query 0, 1, ... then the light ones
        for (u32 q0 = 0; q0 < n_queries; q0 += 32) {
            const u32 q = q0 + lane;
            const u64 v = q < n_queries ? qcur[kQCur * (u64)q + (pass ? 5 : 0)] : 0;
            u64 inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u64 t = __shfl_up_sync(kFull, inc, o);
                if (lane >= o) inc += t;
            }
            if (q < n_queries) qcur[kQCur * (u64)q + (pass ? 2 : 1)] = run + inc - v;
            run += __shfl_sync(kFull, inc, 31);
        }
    if (lane != 0) return;
    const u64 n_items = run;  // live roots = tickets
    jq->head = 0;
    jq->tail = n_items;
    jq->pending = (long long)n_items;
    jq->idle = 0;
    jq->n_init = n_items;
    jq->steps = 0;
    jq->exports = 0;
    jq->donations = 0;
    jq->full = 0;
    jq->warp_iters = 0;
    jq->lane_iters = 0;
    jq->idle_polls = 0;
}

__global__ void __launch_bounds__(256) k3_init_items_kernel(u32 n_queries, const u32 *__restrict__ q_vbase,
                                                            const JoinDepth *__restrict__ jplan,
                                                            const u64 *__restrict__ cand_off,
                                                            const u32 *__restrict__ cand,
                                                            const u32 *__restrict__ deg,
                                                            const u32 *__restrict__ lcoff, u32 n_labels,
                                                            const u64 *__restrict__ item_base, u32 rank, u32 world,
                                                            u32 heavy_deg, const u64 *__restrict__ tpool, u64 *qcur,
                                                            uint2 *init, const u64 *__restrict__ cand_overflow) {
    if (cand_overflow && *cand_overflow) return;
    const u64 n_items = item_base[n_queries], n_round = (n_items + 31) / 32 * 32;
    const int lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    for (u64 item = (u64)blockIdx.x * blockDim.x + threadIdx.x; item < n_round; item += (u64)gridDim.x * blockDim.x) {
        RootItem r{0xffffffffu, 0, false, false};
        if (item < n_items)
            r = root_item(item, n_queries, q_vbase, jplan, cand_off, cand, deg, lcoff, n_labels, item_base, rank, world,
                          heavy_deg, tpool);
        const bool valid = r.alive;
        // one atomic per (warp, query, class)
        const unsigned peers = __match_any_sync(kFull, valid ? (r.q << 1 | (r.heavy ? 1u : 0u)) : 0xffffffffu);
        const int leader = __ffs(peers) - 1;
        u64 base = 0;
        if (valid && lane == leader) {
            u64 *qc = qcur + kQCur * (u64)r.q;
            base = r.heavy ? qc[1] + atomicAdd((unsigned long long *)&qc[3], (unsigned long long)__popc(peers))
                           : qc[2] + atomicAdd((unsigned long long *)&qc[4], (unsigned long long)__popc(peers));
        }
        base = __shfl_sync(kFull, base, leader);
        if (valid) init[base + __popc(peers & lt)] = make_uint2(r.q, r.pos);
    }
}

// ---- subtree tables ---------------------------------------------------------------------------------------------
// N_v[x], for a query vertex v with peeled children and every data vertex x of v's label: the number of ways to map
// everything that hangs below v in peeled subtrees when v -> x,
//     N_v[x] = prod over peeled children c of v ( sum over y in N(x), label(y) = label(c), deg(y) >= deg(c)
//                                                  [, y in C(start) when c is the start vertex] of N_c[y] ),
// with N_c = 1 for a child without children.  All labels of a peeled subtree are unique in the query, so these maps
// are injective and disjoint from the rest by construction.  ONE cooperative launch for all levels (children before
// parents, a grid-wide barrier between two levels): a batch without any table -- large label alphabets, where the plan
// walks instead (k3_order) -- costs one launch that reads one counter, not one launch per possible level.
__global__ void __launch_bounds__(256) k3_tree_tables_kernel(JoinGraph g, const TreeJob *__restrict__ tjobs,
                                                             const u32 *__restrict__ tchild, u32 max_level,
                                                             const u32 *__restrict__ tcount,
                                                             const u32 *__restrict__ tlist, u32 n_slots,
                                                             const u32 *__restrict__ bitmap, u64 words_per_slot,
                                                             u64 *tpool, u32 max_class) {
 cooperative_groups::grid_group grid = cooperative_groups::this_grid();
 for (u32 level = 1; level <= max_level; level++) {
  // work unit = (table of this level, block of 256 class positions).  Levels are filled bottom-up: a table of level k has
  // a child table of level k-1, so the first empty level ends the kernel (every block reads the same counter)
  const u32 n_jobs = tcount[level];
  if (n_jobs == 0) break;
  if (level > 1) grid.sync();  // the tables of the level below are complete
  // (the table's description and its children's are block-uniform loads that stay in L1; no shared memory, no block
  //  barrier: warps of a block run ahead of each other -- a capture with per-table barriers spent half its samples in them)
  const u32 chunks = (max_class + 255) / 256;
  const int lane = threadIdx.x & 31;
  for (u64 unit = blockIdx.x; unit < (u64)n_jobs * chunks; unit += gridDim.x) {
    const u32 ji = (u32)(unit / chunks), chunk = (u32)(unit % chunks);
    const TreeJob job = tjobs[tlist[(u64)level * n_slots + ji]];
    if (job.label >= g.nl) continue;
    const u32 c0 = g.lcoff[job.label], n = g.lcoff[job.label + 1] - c0;
    const u32 pos = chunk * 256 + threadIdx.x;
    if (chunk * 256 + (threadIdx.x & ~31u) >= n) continue;  // warp-uniform: nothing of this warp's 32 positions is in the class
    const bool valid = pos < n;
    DirRow row{};
    if (valid) row = dir_row(g, c0 + pos);  // class-order id of the class's pos-th vertex: rows are contiguous
    u64 val = valid ? 1 : 0;
    for (u32 k = 0; k < job.n_child; k++) {  // warp-uniform loop: a lane whose product is already 0 just idles
        const TreeJob cj = tjobs[tchild[job.child_begin + k]];
        u32 s = 0, e = 0;
        if (val && cj.label < g.nl) row_range(g, row, cj.label, s, e);
        u64 sum;
        if (cj.level == 0 && cj.start_slot == 0xffffffffu) {
            sum = e - s;  // a plain leaf: every neighbour of the label (its degree is >= 1)
        } else {
            const u32 *bm = cj.start_slot == 0xffffffffu ? nullptr : bitmap + (u64)cj.start_slot * words_per_slot;
            const u32 cfirst = cj.label < g.nl ? g.lcoff[cj.label] : 0;
            auto weight = [&](u32 at) -> u64 {
                u32 y, ydeg;
                adj_entry(g, at, y, ydeg);
                if (ydeg < cj.qdeg) return 0;
                const u32 yp = y - cfirst;
                if (bm && !(bm[yp >> 5] >> (yp & 31) & 1)) return 0;
                return cj.level ? tpool[cj.table_off + yp] : 1;
            };
            // Group sizes are heavy-tailed (a hub has 50 neighbours of one label where the average vertex has one): a lane
            // sums a short group itself, a long one is summed by the whole warp -- otherwise one hub row holds up 31 lanes
            constexpr u32 kLong = 8;
            sum = 0;
            if (e - s <= kLong)
                for (u32 at = s; at < e; at++) sum = sat_add(sum, weight(at));
            unsigned big = __ballot_sync(kFull, e - s > kLong);
            while (big) {
                const int src = __ffs(big) - 1;
                big &= big - 1;
                const u32 ss = __shfl_sync(kFull, s, src), ee = __shfl_sync(kFull, e, src);
                u64 part = 0;
                for (u32 at = ss + lane; at < ee; at += 32) part = sat_add(part, weight(at));
#pragma unroll
                for (int o = 16; o; o >>= 1) part = sat_add(part, __shfl_xor_sync(kFull, part, o));
                if (lane == src) sum = part;
            }
        }
        val = sat_mul(val, sum);
    }
    if (valid) tpool[job.table_off + pos] = val;
  }
 }
}

template <int M, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k3_dfs_kernel(JoinGraph g, const u32 *__restrict__ q_vbase,
                                                         const JoinDepth *__restrict__ jplan,
                                                         const uint2 *__restrict__ kids,
                                                         const u32 *__restrict__ cand, const uint2 *__restrict__ init,
                                                         const u64 *__restrict__ limits, u64 *answers, u32 *items,
                                                         u64 export_cap, u32 *ready, u32 epoch, JoinQueue *jq,
                                                         u32 *matches, u64 matches_cap, u64 *match_cursor, u32 flags,
                                                         u64 *inexact) {
    constexpr u32 stride = item_stride(M);
    const bool cgl = flags & 2u;  // bloom words and class positions go around L1
    const u32 exp_mask = (1u << (flags >> 4 & 7u)) - 1;  // rounds between two looks at the queue header, minus one (kExportEvery)
    const int tail_batch = (int)(flags >> 8 & 0xffu);  // parked lanes that trigger a joint evaluation (kTailBatch)
    const int spr = (int)(flags >> 16 & 0xffu);        // DFS steps of a lane between two rounds of scheduling (kStepsPerRound)
    const u32 split = flags >> 24 & 0xfu;              // pieces an exported sibling range is cut into (kSplit)
    const int export_lanes = (int)(flags >> 28 & 0xfu);  // lanes of a warp that may hand work over in one round (kExportLanes)
    const u32 backoff_max = 512u << (flags >> 2 & 3u);   // longest sleep of a warp without work between two looks at the queue (ns)
    extern __shared__ u64 s_stack64[];  // prod [M][THREADS] u64 | emb | cur | end | s0 | e0, each [M][THREADS] u32
    u64 *prod = s_stack64 + threadIdx.x;
    u32 *emb = reinterpret_cast<u32 *>(s_stack64 + M * THREADS) + threadIdx.x;
    u32 *cur = emb + M * THREADS;
    u32 *end = cur + M * THREADS;
    u32 *s0 = end + M * THREADS;
    u32 *e0 = s0 + M * THREADS;
    unsigned char *w_slot = reinterpret_cast<unsigned char *>(s_stack64 + (size_t)M * THREADS * 7 / 2) + (threadIdx.x & ~31u);
#define PROD(t) prod[(t) * THREADS]
#define EMB(t) emb[(t) * THREADS]
#define CUR(t) cur[(t) * THREADS]
#define END(t) end[(t) * THREADS]
#define S0(t) s0[(t) * THREADS]
#define E0(t) e0[(t) * THREADS]
    const int lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const u64 n_init = jq->n_init;  // written by k3_init_items_kernel, the previous launch on this stream
    bool have = false, ticketed = false, can_export = true;
    bool pend = false;  // parked: EMB(d) holds a candidate that passed the test and waits for its counted-tail factors
    u64 pend_p = 0;     // its product so far
    bool registered = false;  // warp-uniform: counted in JoinQueue::idle
    u64 ticket = 0;
    u32 q = 0, vb = 0, nq = 0, base = 0, d = 0, lab0 = 0, tail_at = 0;
    u32 acc_q = 0xffffffffu;
    u32 lim32 = 0xffffffffu;  // answer limit of the lane's query (custom.h:851-854); 0xffffffff = `-n MAX`: count everything
    u64 acc = 0, my_steps = 0, my_exports = 0, my_donations = 0;
    u64 w_iters = 0, w_polls = 0;                       // warp-uniform: iterations with / without a busy lane
    u32 claimed = 0, iter = 0, backoff = 64;            // warp-uniform
    u64 h_head = 0, h_tail = n_init, h_idle = 0;        // warp-uniform cached copy of the queue header
    long long h_pending = 1;
    u64 pf_a = 0, pf_b = 0, pf_c = 0, pf_d = 0;         // lane 0: header words in flight (issued last iteration)
    bool pf_valid = false;                              // warp-uniform

    for (;;) {
        ++iter;
        unsigned busy = __ballot_sync(kFull, have);
        unsigned tick = __ballot_sync(kFull, ticketed);
        if (!busy && claimed) {  // everything this warp took from the queue is finished
            if (lane == 0) atomicAdd((unsigned long long *)&jq->pending, 0ull - (unsigned long long)claimed);
            claimed = 0;
        }
        unsigned free_m = ~(busy | tick);
        // ---- queue header: a warp without work reads it now; a busy warp consumes the copy it asked for one
        //      iteration ago, so the round trip to L2 hides behind a DFS step
        bool fresh = false;
        if (!busy) {
            if (lane == 0) {
                ld_relaxed_2xu64(&jq->head, pf_a, pf_b);
                ld_relaxed_2xu64(&jq->pending, pf_c, pf_d);
            }
            pf_valid = true;
        }
        if (pf_valid) {
            h_head = __shfl_sync(kFull, pf_a, 0);
            h_tail = __shfl_sync(kFull, pf_b, 0);
            h_pending = (long long)__shfl_sync(kFull, pf_c, 0);
            h_idle = __shfl_sync(kFull, pf_d, 0);
            pf_valid = false;
            fresh = true;
            if (h_pending <= 0 && !busy) break;  // join complete
        }
        if (busy && ((iter & exp_mask) == 0 || ((free_m | tick) && (iter & 7) == 0))) {
            if (lane == 0) {
                ld_relaxed_2xu64(&jq->head, pf_a, pf_b);
                ld_relaxed_2xu64(&jq->pending, pf_c, pf_d);
            }
            pf_valid = true;
        }
        // ---- tickets: lanes without work claim published items (far from the end of the queue the cached
        //      header is good enough; near the end only a fresh one is trusted)
        if (free_m && h_head < h_tail && (fresh || h_head + 65536 < h_tail)) {
            const u64 avail = h_tail - h_head;
            // scarce work (fewer tickets than lanes on the GPU: small batches, selective filters) is spread one ticket per
            // warp -- 32 different roots in the lanes of one warp run 32 divergent walks one after the other while
            // thousands of warps sit idle; the lanes of the claiming warp are fed by donation instead
            // (decided once per launch from the number of start tickets: the tail of a long queue is still claimed 32 at a time)
            const u64 n_warps = (u64)gridDim.x * (THREADS / 32);
            const u32 share = n_init >= n_warps * 32 ? 32u : (u32)max((n_init + n_warps - 1) / n_warps, (u64)1);
            const u32 n_want = (u32)min((u64)min((u32)__popc(free_m), share), avail);
            u64 b0 = 0;
            if (lane == 0) b0 = atomicAdd(&jq->head, (unsigned long long)n_want);
            b0 = __shfl_sync(kFull, b0, 0);
            h_head = b0 + n_want;
            const u32 r = __popc(free_m & lt);
            if ((free_m >> lane & 1) && r < n_want) {
                ticket = b0 + r;
                ticketed = true;
            }
        }
        // ---- ticketed lanes: start candidates are always there, exported items once their flag is up
        bool got = false;
        if (ticketed) {
            if (ticket < n_init) {
                ticketed = false;
                got = true;
                const uint2 it = init[ticket];
                const u32 iq = it.x;
                if (iq != acc_q) {
                    if (acc) flush_answer(answers, acc_q, acc);
                    acc = 0;
                    acc_q = iq;
                }
                u64 limit = limits ? limits[iq] : GPE_LIMIT_MAX;
                if (limit == 0) limit = 1;  // the reference tests the limit only after counting a match (:851)
                lim32 = limit >= GPE_LIMIT_MAX ? 0xffffffffu : (u32)limit;
                q = iq;
                vb = q_vbase[q];
                nq = jplan[vb].sure_used;  // depths of the execution plan (peeled subtrees are not walked)
                if (limit >= GPE_LIMIT_MAX || *(volatile u64 *)&answers[q] < limit) {  // (no limit: nothing to read)
                    tail_at = matches ? nq : nq - jplan[vb].tail_k;  // depth at which the counting shortcut takes over
                    base = d = 0;
                    CUR(0) = it.y;
                    END(0) = it.y + 1;
                    have = true;
                }
            } else if (fresh && ticket - n_init < export_cap && ld_acquire_u32(ready + (ticket - n_init)) == epoch) {
                ticketed = false;
                got = true;
                const u32 *it = items + (ticket - n_init) * stride;
                const u32 iq = it[0];
                if (iq != acc_q) {
                    if (acc) flush_answer(answers, acc_q, acc);
                    acc = 0;
                    acc_q = iq;
                }
                u64 limit = limits ? limits[iq] : GPE_LIMIT_MAX;
                if (limit == 0) limit = 1;
                lim32 = limit >= GPE_LIMIT_MAX ? 0xffffffffu : (u32)limit;
                q = iq;
                vb = q_vbase[q];
                nq = jplan[vb].sure_used;
                base = it[1];
                if (it[2] < it[3] && (limit >= GPE_LIMIT_MAX || *(volatile u64 *)&answers[q] < limit)) {
                    lab0 = it[6];
                    tail_at = matches ? nq : nq - jplan[vb].tail_k;
                    for (u32 t = 0; t < base; t++) EMB(t) = it[kItemHdr + t];
                    for (u32 t = base; t < nq; t++) {
                        S0(t) = it[kItemHdr + M + t];
                        E0(t) = it[kItemHdr + 2 * M + t];
                    }
                    if (base) PROD(base - 1) = (u64)it[4] | ((u64)it[5] << 32);
                    d = base;
                    CUR(d) = it[2];
                    END(d) = it[3];
                    have = true;
                }
            }
        }
        {
            const unsigned got_m = __ballot_sync(kFull, got);
            if (got_m) claimed += __popc(got_m);
        }
        busy = __ballot_sync(kFull, have);
        if (!busy) {
            if (!registered) {
                registered = true;
                if (lane == 0) atomicAdd(&jq->idle, 1ull);
            }
            __nanosleep(backoff);
            if (backoff < backoff_max) backoff <<= 1;
            w_polls++;
            continue;
        }
        if (registered) {
            registered = false;
            backoff = 64;
            if (lane == 0) atomicAdd(&jq->idle, 0ull - 1ull);
        }

        // ---- inside the warp: lanes without work take half of a busy lane's shallowest sibling range ----
        free_m = ~(busy | __ballot_sync(kFull, ticketed));
        if (free_m) {
            u32 l = 0, rem = 0;
            if (have) {
                for (l = base; l <= d; l++) {
                    const u32 r = END(l) - CUR(l);  // CUR <= END always
                    if (r >= 2 || (r == 1 && l < d)) { rem = r; break; }
                }
            }
            const unsigned don_m = __ballot_sync(kFull, rem > 0);
            const int n_pairs = min(__popc(free_m), __popc(don_m));
            if (n_pairs) {
                const bool is_free = free_m >> lane & 1;
                const int my_rank = is_free ? __popc(free_m & lt) : __popc(don_m & lt);
                const bool give = rem > 0 && my_rank < n_pairs;
                const bool take = is_free && my_rank < n_pairs;
                if (give) w_slot[my_rank] = (unsigned char)lane;
                __syncwarp();
                const int partner = take ? (int)w_slot[my_rank] : lane;
                u32 lo_g = 0, hi_g = 0;
                if (give) {
                    hi_g = END(l);
                    lo_g = hi_g - (rem == 1 ? 1u : rem / 2);  // keep the lower part, give the upper one
                    END(l) = lo_g;
                }
                __syncwarp();
                const u32 p_q = __shfl_sync(kFull, q, partner), p_vb = __shfl_sync(kFull, vb, partner);
                const u32 p_nq = __shfl_sync(kFull, nq, partner), p_lab0 = __shfl_sync(kFull, lab0, partner);
                const u32 p_tail = __shfl_sync(kFull, tail_at, partner), p_l = __shfl_sync(kFull, l, partner);
                const u32 p_lo = __shfl_sync(kFull, lo_g, partner), p_hi = __shfl_sync(kFull, hi_g, partner);
                const u32 p_lim = __shfl_sync(kFull, lim32, partner);
                if (take) {
                    if (p_q != acc_q) {
                        if (acc) flush_answer(answers, acc_q, acc);
                        acc = 0;
                        acc_q = p_q;
                    }
                    q = p_q; vb = p_vb; nq = p_nq; lab0 = p_lab0; tail_at = p_tail; lim32 = p_lim;
                    base = d = p_l;
                    const int col = partner - lane;  // the donor's stack column, relative to mine
                    for (u32 t = 0; t < p_l; t++) EMB(t) = emb[t * THREADS + col];
                    for (u32 t = p_l; t < p_nq; t++) {
                        S0(t) = s0[t * THREADS + col];
                        E0(t) = e0[t * THREADS + col];
                    }
                    if (p_l) PROD(p_l - 1) = prod[(p_l - 1) * THREADS + col];
                    CUR(d) = p_lo;
                    END(d) = p_hi;
                    have = true;
                    my_donations++;
                }
                __syncwarp();
            }
        }

        // answer limit (`-n N`): the reference stops enumerating once the count reaches it (custom.h:851-854); a lane of a
        // limited query banks straight into answers[] (below) and drops its work as soon as the query's count is there
        if (have && lim32 != 0xffffffffu && *(volatile u64 *)&answers[q] >= lim32) {
            have = false;
            pend = false;
        }
        w_iters += spr;
        for (int rep = 0; rep < spr; rep++) {
          // A step has two parts with very different lane populations: the candidate test (most busy lanes) and the
          // counted-tail factors of a candidate that passed it (r01k capture: 40 % of the kernel's instructions with
          // 3-4 active lanes).  A lane whose candidate needs factors PARKS (pend) instead of evaluating them on the
          // spot; the warp evaluates all parked lanes together once kTailBatch of them wait or nobody else can move.
          u64 fin_p = 0;  // product to bank (leaf of the walk) or to descend with; 0 = nothing to do
          if (have && !pend) {
            // ---- one DFS step: test the next candidate of level d ----
            // (measured and dropped, config 2: letting a lane retry siblings until one passes the cheap tests -- 2/4/8 tries
            //  cost +0.6/+1.2/+1.4 ms, a retry lengthens the warp's critical chain; software prefetch of the next
            //  sibling's rows +0.9 ms; group lookups issued ahead of the subtree-table test: no change)
            my_steps++;
            const u32 at = CUR(d);
            CUR(d) = at + 1;
            const JoinDepth *jd = jplan + vb + d;
            u32 c, cdeg;
            if (d == 0) {
                if (jd->tail_mask) {  // the walk starts from the root's label class (the start vertex was peeled)
                    c = at;           // a position in the class-ordered vertex list is the id itself
                    cdeg = g.degJ[c];
                    lab0 = jd->label;
                } else {  // start candidates are taken as they are (the reference never checks them, custom.h:827-830)
                    c = cand[at];
                    cdeg = 0xffffffffu;
                    lab0 = g.labelJ[c];
                }
            } else {
                adj_entry(g, at, c, cdeg);
            }
            // (bit scans -- ffs/popc -- run on the quarter-rate XU pipe, which a first version of this loop saturated:
            //  the masks of the plan are walked with shifts instead)
            bool ok = cdeg >= jd->deg;
            // everything that hangs below this query vertex in peeled subtrees: one factor per data vertex
            const u64 tree_off = matches ? kNoTree : jd->tree_off;
            u64 tree_f = 1;
            if (ok && tree_off != kNoTree) {
                tree_f = __ldcg(g.tpool + tree_off + c);
                ok = tree_f != 0;
            }
            // injective: only earlier depths of the same label could collide (k3_order); enumeration mode also walks the
            // tail depths, whose tail_mask means something else: there every earlier depth is compared
            u64 sm = d ? (matches ? (1ull << d) - 1 : jd->tail_mask) : 0;
            for (u32 t = 0; sm; t++, sm >>= 1)
                if (sm & 1) ok = ok && EMB(t) != c;
            const DirRow row = dir_row(g, c);
            u64 bn = d ? jd->bn_mask : 0;
            for (u32 t = 0; ok && bn; t++, bn >>= 1) {  // the other backward neighbours: edge (c, EMB(t)) must exist
                if (!(bn & 1)) continue;
                const u32 lt_ = t ? jplan[vb + t].label : lab0;
                if (!(cgl ? edge_maybe<true>(g, c, EMB(t)) : edge_maybe<false>(g, c, EMB(t)))) {
                    ok = false;
                } else if (cdeg <= 64) {  // search c's (short) group of label(EMB(t))
                    u32 s = 0, e = 0;
                    if (lt_ < g.nl) row_range(g, row, lt_, s, e);
                    ok = in_group(g, s, e, EMB(t));
                } else {           // search EMB(t)'s group of c's label
                    u32 s, e;
                    group_range(g, EMB(t), jd->label, s, e);
                    ok = in_group(g, s, e, c);
                }
            }
            if (ok) {
                EMB(d) = c;
                // (1) label groups of c that later depths draw from: all in c's gtab row, two lookups in flight
                const uint2 *kl = kids + vb + jd->kid_begin;
                const u32 kn = jd->kid_count;
                for (u32 k = 0; k < kn; k += 2) {
                    const uint2 k0 = kl[k], k1 = kl[min(k + 1, kn - 1)];  // (depth, label)
                    const u32 i0 = k0.x, l0 = k0.y, i1 = k1.x, l1 = k1.y;
                    u32 a0 = 0, b0 = 0, a1 = 0, b1 = 0;
                    if (l0 < g.nl) row_range(g, row, l0, a0, b0);
                    if (l1 < g.nl) row_range(g, row, l1, a1, b1);
                    S0(i0) = a0; E0(i0) = b0;
                    S0(i1) = a1; E0(i1) = b1;
                    if (a0 >= b0 || a1 >= b1) { ok = false; break; }  // nothing to draw from: no match below c
                }
            }
            if (ok) {
                const u64 p = d ? sat_mul(PROD(d - 1), tree_f) : tree_f;
                const u64 um = tail_at < nq ? jd->units_mask >> tail_at : 0;
                if (um && p) { pend = true; pend_p = p; } else fin_p = p;
            }
          }
          // ---- (2) counted-tail factors that close at this depth, for all parked lanes at once ----
          const unsigned pend_m = __ballot_sync(kFull, pend);
          if (pend_m && (__popc(pend_m) >= tail_batch || !__ballot_sync(kFull, have && !pend))) {
            if (pend) {
                pend = false;
                const JoinDepth *jd = jplan + vb + d;
                u64 p = pend_p;
                u64 um = jd->units_mask >> tail_at;
                for (u32 i = tail_at; um && p; i++, um >>= 1) {
                    if (!(um & 1)) continue;
                    const JoinDepth *ld = jplan + vb + i;
                    // (weighted) number of free members of a counted leaf's group: its size minus the prefix vertices inside
                    // it -- or, for a leaf that carries peeled subtrees, S_u[pivot] minus the N_u of those prefix vertices
                    auto leaf_free = [&](const JoinDepth *lf, u32 s, u32 e) -> u64 {
                        const u32 llab = lf->label, pvx = EMB(lf->pivot_depth);
                        if (!(lf->tail_k & kTailW)) {
                            u32 used = lf->sure_used;
                            u64 m = lf->tail_mask;
                            for (u32 t = 0; m; t++, m >>= 1)
                                if ((m & 1) && (t != 0 || lab0 == llab) &&
                                    (cgl ? edge_maybe<true>(g, pvx, EMB(t)) : edge_maybe<false>(g, pvx, EMB(t))) &&
                                    in_group(g, s, e, EMB(t))) used++;
                            return (u64)((e - s) - used);
                        }
                        u64 wsum = __ldcg(g.tpool + lf->units_mask + pvx);
                        if (wsum >= kSat) inexact[q] = 1;  // saturated minuend: the difference below is unknown
                        const u64 sure = lf->bn_mask;
                        u64 m = lf->tail_mask | sure;
                        for (u32 t = 0; m; t++, m >>= 1) {
                            if (!(m & 1)) continue;
                            const u32 y = EMB(t);
                            const bool in = (sure >> t & 1) ||
                                            ((t != 0 || lab0 == llab) &&
                                             (cgl ? edge_maybe<true>(g, pvx, y) : edge_maybe<false>(g, pvx, y)) && in_group(g, s, e, y));
                            if (in && g.degJ[y] >= lf->deg)
                                wsum -= lf->tree_off != kNoTree ? __ldcg(g.tpool + lf->tree_off + y) : 1ull;
                        }
                        return wsum;
                    };
                    // weight of one group member (entry of nbrL) as image of a counted leaf
                    auto leaf_weight = [&](const JoinDepth *lf, u32 id, u32 dg) -> u64 {
                        if (!(lf->tail_k & kTailW)) return 1ull;
                        if (dg < lf->deg) return 0ull;
                        if (lf->tree_off == kNoTree) return 1ull;
                        return __ldcg(g.tpool + lf->tree_off + id);
                    };
                    const u32 s = S0(i), e = E0(i);
                    const u64 n_free = leaf_free(ld, s, e);
                    if ((ld->tail_k & 0xffu) == kTailMul) {
                        u64 f = n_free;
                        u64 n_run = n_free;  // (followers are plain leaves on the same pivot: falling factorial)
                        for (u32 k = i + 1; k < nq && (jplan[vb + k].tail_k & 0xffu) == kTailFall; k++) {
                            n_run = n_run ? n_run - 1 : 0;
                            f = sat_mul(f, n_run);
                        }
                        p = sat_mul(p, f);
                    } else {  // kTailPairA: leaf i and leaf i+1, same label, different pivots
                        const JoinDepth *lb = ld + 1;
                        const u32 s2 = S0(i + 1), e2 = E0(i + 1);
                        const u64 n_free2 = leaf_free(lb, s2, e2);
                        // ordered pairs of distinct vertices: W(A) W(B) - sum over the free members of both groups of
                        // wA wB; the groups are ascending id lists, so the intersection is a merge
                        u64 inter = 0;
                        u32 x = s, y = s2;
                        while (x < e && y < e2) {
                            u32 idx, dgx, idy, dgy;
                            adj_entry(g, x, idx, dgx);
                            adj_entry(g, y, idy, dgy);
                            if (idx == idy) {
                                bool is_used = false;
                                for (u32 t = 0; t <= d; t++) is_used = is_used || EMB(t) == idx;
                                if (!is_used) inter += leaf_weight(ld, idx, dgx) * leaf_weight(lb, idy, dgy);
                                x++;
                                y++;
                            } else if (idx < idy) {
                                x++;
                            } else {
                                y++;
                            }
                        }
                        // (plain leaves: group sizes, the product fits; weighted ones can exceed 64 bits)
                        if (__umul64hi(n_free, n_free2) != 0 || n_free * n_free2 > kSat) inexact[q] = 1;
                        p = sat_mul(p, n_free * n_free2 - inter);
                    }
                }
                fin_p = p;
            }
          }
          {
            {
                const u64 p = fin_p;
                if (p) {
                    if (d + 1 == tail_at) {
                        if (lim32 != 0xffffffffu) {  // limited query: bank at once so that every lane sees the limit reached
                            const u64 pc = p < kFlushCap ? p : kFlushCap;
                            if (atomicAdd((unsigned long long *)&answers[q], (unsigned long long)pc) + pc >= lim32) have = false;
                        } else {
                            acc = sat_add(acc, p);
                        }
                        if (matches) {  // tail_at == nq here: every vertex was walked
                            u64 pos = atomicAdd((unsigned long long *)match_cursor, 1ull);
                            if (pos < matches_cap) {
                                u32 *out = matches + pos * nq;
                                for (u32 t = 0; t <= d; t++) out[jplan[vb + t].u] = g.orig[EMB(t)];  // back to the caller's ids
                            }
                        }
                    } else {
                        PROD(d) = p;
                        d++;
                        CUR(d) = S0(d);
                        END(d) = E0(d);
                    }
                }
            }
            // pop exhausted levels (a parked lane keeps its depth: its candidate is still to be descended from)
            while (have && !pend && CUR(d) >= END(d)) {
                if (d == base) have = false; else d--;
            }
          }
        }

        // ---- between warps: when idle warps outnumber the published items, the lanes with the shallowest unexplored
        //      sibling ranges give them away ----
        if ((iter & exp_mask) == (1u & exp_mask) && h_idle > 0 && (long long)(h_tail - h_head) < (long long)(h_idle * 8)) {
            u32 key = 0xffffffffu, l = 0;
            if (have && can_export) {
                for (l = base; l <= d; l++)
                    if (CUR(l) < END(l)) { key = (l << 5) | (u32)lane; break; }
            }
            for (int round = 0; round < export_lanes; round++) {
                const u32 best = __reduce_min_sync(kFull, key);
                if (best == 0xffffffffu) break;
                if (key != best) continue;
                key = 0xffffffffu;
                const u32 c0 = CUR(l), len = END(l) - c0, np = min(len, split);
                const u64 o = atomicAdd(&jq->tail, (unsigned long long)np) - n_init;
                if (o + np > export_cap) {
                    can_export = false;  // those tickets are never published; their holders idle until the end
                    jq->full = 1;
                } else {
                    atomicAdd((unsigned long long *)&jq->pending, (unsigned long long)np);
                    const u64 pp = l ? PROD(l - 1) : 1;
                    for (u32 k = 0; k < np; k++) {
                        u32 *it = items + (o + k) * stride;
                        it[0] = q;
                        it[1] = l;
                        it[2] = c0 + (u32)((u64)len * k / np);
                        it[3] = c0 + (u32)((u64)len * (k + 1) / np);
                        it[4] = (u32)pp;
                        it[5] = (u32)(pp >> 32);
                        it[6] = lab0;
                        for (u32 t = 0; t < l; t++) it[kItemHdr + t] = EMB(t);
                        for (u32 t = l; t < nq; t++) {
                            it[kItemHdr + M + t] = S0(t);
                            it[kItemHdr + 2 * M + t] = E0(t);
                        }
                    }
                    __threadfence();
                    for (u32 k = 0; k < np; k++) st_relaxed_u32(ready + o + k, epoch);
                    END(l) = c0;
                    my_exports++;
                }
            }
            while (have && !pend && CUR(d) >= END(d)) {  // the exported range may have been this lane's current level
                if (d == base) have = false; else d--;
            }
        }
    }
    if (acc) flush_answer(answers, acc_q, acc);
    for (int o = 16; o; o >>= 1) {
        my_steps += __shfl_xor_sync(kFull, my_steps, o);
        my_exports += __shfl_xor_sync(kFull, my_exports, o);
        my_donations += __shfl_xor_sync(kFull, my_donations, o);
    }
    if (lane == 0 && my_steps) atomicAdd(&jq->steps, (unsigned long long)my_steps);
    if (lane == 0 && my_exports) atomicAdd(&jq->exports, (unsigned long long)my_exports);
    if (lane == 0 && my_donations) atomicAdd(&jq->donations, (unsigned long long)my_donations);
    if (lane == 0 && w_iters) atomicAdd(&jq->warp_iters, (unsigned long long)w_iters);
    if (lane == 0 && w_polls) atomicAdd(&jq->idle_polls, (unsigned long long)w_polls);
#undef PROD
#undef EMB
#undef CUR
#undef END
#undef S0
#undef E0
}

}  // namespace

cudaError_t k3_chunk_count(const u32 *bitmap, u64 words_per_slot, u64 chunks_per_slot, u32 n_slots, u64 *chunk_cnt,
                           cudaStream_t s) {
    u64 n_chunks = chunks_per_slot * n_slots;
    u64 warps = n_chunks + 1;
    k3_chunk_count_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, s>>>(bitmap, words_per_slot, chunks_per_slot,
                                                                              n_chunks, chunk_cnt);
    return cudaGetLastError();
}

cudaError_t k3_merge_count(const u32 *all, u64 shard_words, u32 world, u32 *bitmap, u64 words_per_slot,
                           u64 chunks_per_slot, u32 n_slots, u64 *chunk_cnt, cudaStream_t s) {
    u64 n_chunks = chunks_per_slot * n_slots;
    u64 warps = n_chunks + 1;
    k3_merge_count_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, s>>>(all, shard_words, world, bitmap, words_per_slot,
                                                                              chunks_per_slot, n_chunks, chunk_cnt);
    return cudaGetLastError();
}

cudaError_t k3_compact(const u32 *bitmap, u64 words_per_slot, u64 chunks_per_slot, u32 n_slots, const u64 *chunk_off,
                       const u32 *slot_label, const u32 *lcoff, u32 n_labels, u32 *cand, u64 *cand_off, u64 cap, u64 *overflow,
                       cudaStream_t s) {
    u64 n_chunks = chunks_per_slot * n_slots;
    if (n_chunks == 0) return cudaSuccess;
    k3_compact_kernel<<<(unsigned)((n_chunks * 32 + 255) / 256), 256, 0, s>>>(bitmap, words_per_slot, chunks_per_slot,
                                                                             n_chunks, n_slots, chunk_off, slot_label,
                                                                             lcoff, n_labels, cand, cand_off, cap, overflow);
    return cudaGetLastError();
}

cudaError_t k3_sparse_pack(const u32 *bitmap, u64 n_words, u64 cap, void *buf /*16 + cap x 8 bytes*/, int sm_count, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(buf, 0, 16, s);
    if (e != cudaSuccess) return e;
    if (n_words == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)std::min<u64>((n_words + 255) / 256, (u64)sm_count * 16);
    k3_sparse_pack_kernel<<<blocks, 256, 0, s>>>(bitmap, n_words, cap, reinterpret_cast<unsigned long long *>(buf),
                                                 reinterpret_cast<uint2 *>(reinterpret_cast<unsigned char *>(buf) + 16));
    return cudaGetLastError();
}

cudaError_t k3_sparse_merge(const void *all, u64 stride_bytes, u32 world, u32 my_rank, u64 cap, u32 *bitmap, int sm_count,
                            cudaStream_t s) {
    if (world <= 1) return cudaSuccess;
    dim3 grid((unsigned)sm_count * 2, world);
    k3_sparse_merge_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const unsigned char *>(all), stride_bytes, world, my_rank, cap, bitmap);
    return cudaGetLastError();
}

cudaError_t k3_scatter(const u32 *counts, const u32 *cand, u64 stride, u32 world, u32 n_slots, const u32 *slot_label,
                       const u32 *lcoff, u32 n_labels, u32 *bitmap, u64 words_per_slot, u64 *prefix_tmp, cudaStream_t s) {
    if (world * n_slots == 0) return cudaSuccess;
    k3_scatter_prefix_kernel<<<(world + 31) / 32, 32, 0, s>>>(counts, world, n_slots, prefix_tmp);
    unsigned blocks = std::min<unsigned>(world * n_slots, 148 * 8);
    k3_scatter_kernel<<<blocks, 256, 0, s>>>(counts, cand, prefix_tmp, stride, world, n_slots, slot_label, lcoff, n_labels, bitmap,
                                             words_per_slot);
    return cudaGetLastError();
}

cudaError_t k3_counts_from_offsets(const u64 *cand_off, u32 n_slots, u32 *counts, cudaStream_t s) {
    if (n_slots == 0) return cudaSuccess;
    k3_counts_kernel<<<(n_slots + 255) / 256, 256, 0, s>>>(cand_off, n_slots, counts);
    return cudaGetLastError();
}

cudaError_t k3_order(u32 n_queries, u32 V, const u32 *q_vbase, const u32 *q_ebase, const u32 *q_offsets,
                     const u32 *q_nbrs, const u32 *q_labels, const u64 *cand_off, u32 *order, u32 *pivot,
                     JoinDepth *jplan, void *kids, u64 *item_base, u32 rank, u32 world, bool enumerate, bool clean_start,
                     u32 n_labels, const u32 *lcoff, TreeJob *tjobs, u32 *tchild, u64 *tcursor, u32 *tcount, u32 *tlist,
                     u32 n_slots, bool allow_weighted, const u32 *qmode, float branching, u64 pool_cap, cudaStream_t s) {
    // jobs / child lists / per-level job lists hold 2 x n_slots entries: [0, n_slots) the tables N_v of vertices with
    // peeled children, [n_slots, 2 n_slots) the tables S_u of weighted counted leaves
    k3_order_kernel<<<(n_queries * 32 + 127) / 128 + 1, 128, 0, s>>>(n_queries, V, q_vbase, q_ebase, q_offsets, q_nbrs, q_labels, cand_off, order,
                                      pivot, jplan, reinterpret_cast<uint2 *>(kids), item_base, rank, world, 1, enumerate,
                                      clean_start, n_labels, lcoff, tjobs, tchild, tcursor, tcount, tlist, 2 * n_slots,
                                      n_slots, allow_weighted, qmode, branching, pool_cap);
    k3_order_prefix_kernel<<<1, 32, 0, s>>>(n_queries, item_base);
    return cudaGetLastError();
}

static u32 join_m(u32 max_nq) { return max_nq <= 8 ? 8 : max_nq <= 16 ? 16 : max_nq <= 32 ? 32 : 64; }

u32 k3_item_stride(u32 max_nq) { return item_stride(join_m(max_nq)); }

cudaError_t k3_init_items(const JoinGraph &jv, u32 n_queries, const u32 *q_vbase, const JoinDepth *jplan,
                          const u64 *cand_off, const u32 *cand, const u64 *item_base, u32 rank, u32 world, u32 heavy_deg,
                          u64 *cursors, void *init, JoinQueue *jq, bool use_tables, const u64 *cand_overflow, int sm_count,
                          cudaStream_t s) {
    // (enumeration mode walks every vertex and ignores the tables: roots are then only filtered by degree)
    const u64 *tp = use_tables ? jv.tpool : nullptr;
    k3_init_count_kernel<<<sm_count * 4, 256, 0, s>>>(n_queries, q_vbase, jplan, cand_off, cand, jv.degJ, jv.lcoff,
                                                     jv.nl, item_base, rank, world, heavy_deg, tp, cursors, cand_overflow);
    k3_init_prefix_kernel<<<1, 32, 0, s>>>(n_queries, cursors, jq);
    k3_init_items_kernel<<<sm_count * 4, 256, 0, s>>>(n_queries, q_vbase, jplan, cand_off, cand, jv.degJ, jv.lcoff,
                                                     jv.nl, item_base, rank, world, heavy_deg, tp, cursors,
                                                     reinterpret_cast<uint2 *>(init), cand_overflow);
    return cudaGetLastError();
}

cudaError_t k3_tree_tables(const JoinGraph &g, u32 n_slots, u32 max_class, u32 max_level, const TreeJob *tjobs,
                           const u32 *tchild, const u32 *tcount, const u32 *tlist, const u32 *bitmap, u64 words_per_slot,
                           u64 *tpool, int sm_count, cudaStream_t s) {
    if (n_slots == 0 || max_class == 0) return cudaSuccess;
    // one persistent 1-D grid per level over (table, block of 256 class positions) units: a level with a handful of tables
    // still fills the GPU, and a level without any costs a launch of blocks that read one counter and exit
    int dev = 0;
    cudaGetDevice(&dev);
    static int per_sm[kMaxDevices] = {};
    if (!per_sm[dev % kMaxDevices]) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k3_tree_tables_kernel, 256, 0) != cudaSuccess || n < 1) n = 1;
        per_sm[dev % kMaxDevices] = std::min(n, 8);
    }
    const u32 grid = (u32)sm_count * per_sm[dev % kMaxDevices];  // all blocks resident: the kernel synchronises grid-wide
    JoinGraph gg = g;
    u32 list_stride = 2 * n_slots;
    void *args[] = {&gg, (void *)&tjobs, (void *)&tchild, &max_level, (void *)&tcount, (void *)&tlist, &list_stride, (void *)&bitmap,
                    &words_per_slot, &tpool, &max_class};
    return cudaLaunchCooperativeKernel((const void *)k3_tree_tables_kernel, dim3(grid), dim3(256), args, 0, s);
}

cudaError_t k3_dfs(const JoinGraph &g, u32 max_nq, const u32 *q_vbase, const JoinDepth *jplan, const void *kids, const u32 *cand,
                   const void *init, const u64 *limits, u64 *answers, u32 *items, u64 export_cap, u32 *ready, u32 epoch,
                   JoinQueue *jq, u32 *matches, u64 matches_cap, u64 *match_cursor, u64 *inexact, int sm_count, cudaStream_t s) {
    // stack bytes per thread: M x (8 + 5 x 4); threads per CTA chosen so that ~30 warps fit in an SM's shared memory
    // (launch configurations are cached per device: function attributes and occupancy are per-device properties)
    int dev = 0;
    cudaGetDevice(&dev);
#define LAUNCH(M, T, B)                                                                                                \
    static int per_sm_arr_##M##_##B[kMaxDevices] = {};                                                                      \
    int &per_sm_##M##_##B = per_sm_arr_##M##_##B[dev % kMaxDevices];                                                        \
    const size_t smem_##M##_##B = (size_t)M * T * 28 + T;                                                                       \
    if (!per_sm_##M##_##B) {                                                                                                \
        cudaFuncSetAttribute(k3_dfs_kernel<M, T, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_##M##_##B);        \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_##M##_##B, k3_dfs_kernel<M, T, B>, T, smem_##M##_##B) !=           \
                cudaSuccess || per_sm_##M##_##B < 1)                                                                        \
            per_sm_##M##_##B = 1;                                                                                           \
    }                                                                                                                 \
    k3_dfs_kernel<M, T, B><<<sm_count * per_sm_##M##_##B, T, smem_##M##_##B, s>>>(g, q_vbase, jplan, reinterpret_cast<const uint2 *>(kids), cand, \
                                            reinterpret_cast<const uint2 *>(init), limits, answers, items, export_cap, \
                                            ready, epoch, jq, matches, matches_cap, match_cursor, flags, inexact)
    static int env_tb = -1;
    if (env_tb < 0) { const char *e = getenv("GPE_JOIN_TAILBATCH"); env_tb = e ? atoi(e) : kTailBatch; if (env_tb < 1 || env_tb > 32) env_tb = kTailBatch; }
    static int env_spr = -1;
    if (env_spr < 0) { const char *e = getenv("GPE_JOIN_SPR"); env_spr = e ? atoi(e) : kStepsPerRound; if (env_spr < 1 || env_spr > 64) env_spr = kStepsPerRound; }
    static int env_cg = -1;
    if (env_cg < 0) { const char *e = getenv("GPE_JOIN_CG"); env_cg = e ? atoi(e) : 1; }
    static int env_exp = -1;
    if (env_exp < 0) { const char *e = getenv("GPE_JOIN_EXPORT_LOG2"); env_exp = e ? atoi(e) : kExportEveryLog2; if (env_exp < 0 || env_exp > 7) env_exp = kExportEveryLog2; }
    static int env_split = -1;
    if (env_split < 0) { const char *e = getenv("GPE_JOIN_SPLIT"); env_split = e ? atoi(e) : (int)kSplit; if (env_split < 1 || env_split > 15) env_split = (int)kSplit; }
    static int env_xl = -1;
    if (env_xl < 0) { const char *e = getenv("GPE_JOIN_EXPORT_LANES"); env_xl = e ? atoi(e) : kExportLanes; if (env_xl < 1 || env_xl > 15) env_xl = kExportLanes; }
    static int env_bo = -1;  // 0..3: idle warps sleep at most 512 / 1024 / 2048 / 4096 ns between polls
    if (env_bo < 0) { const char *e = getenv("GPE_JOIN_BACKOFF"); env_bo = e ? atoi(e) : 3; if (env_bo < 0 || env_bo > 3) env_bo = 3; }
    const u32 flags = (env_cg ? 2u : 0u) | ((u32)env_bo << 2) | ((u32)env_exp << 4) | ((u32)env_tb << 8) | ((u32)env_spr << 16) | ((u32)env_split << 24) |
                      ((u32)env_xl << 28);
    // 8-vertex stacks, CTAs of 128 threads per SM (config 2, ms per batch): 5 (96 registers) 15.97, 6 (80 registers, 92 bytes of
    // spills) 15.61, 7 (72 registers) 19.9 -- beyond 6 the stacks leave too little of the SM's memory to L1, which holds the plans
    static int env_blocks = -1;
    if (env_blocks < 0) { const char *e = getenv("GPE_JOIN_BLOCKS"); env_blocks = e ? atoi(e) : 5; }  // (round 2, one step per round: 5 <= 6)
    if (max_nq <= 8 && env_blocks == 7) { LAUNCH(8, 128, 7); }
    else if (max_nq <= 8 && env_blocks == 6) { LAUNCH(8, 128, 6); }
    else if (max_nq <= 8) { LAUNCH(8, 128, 5); }
    else if (max_nq <= 16) { LAUNCH(16, 128, 4); }
    else if (max_nq <= 32) { LAUNCH(32, 128, 2); }
    else { LAUNCH(64, 128, 1); }
#undef LAUNCH
    return cudaGetLastError();
}

}  // namespace gpe
