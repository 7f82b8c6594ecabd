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

__global__ void __launch_bounds__(256) k3_compact_kernel(const u32 *__restrict__ bitmap, u64 words_per_slot,
                                                         u64 chunks_per_slot, u64 n_chunks, u32 n_slots,
                                                         const u64 *__restrict__ chunk_off, u32 *__restrict__ cand,
                                                         u64 *__restrict__ cand_off) {
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
    const u32 *p = bitmap + slot * words_per_slot + c * kChunkWords;
    u32 vbase = (u32)(c * kChunkWords * 32);
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
            cand[my++] = v0 + b;
        }
        out += __shfl_sync(kFull, inc, 31);
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
                                                         u32 n_slots, u32 *bitmap, u64 words_per_slot) {
    for (u32 job = blockIdx.x; job < world * n_slots; job += gridDim.x) {
        u32 r = job / n_slots, slot = job % n_slots;
        u32 n = counts[(u64)r * n_slots + slot];
        const u32 *list = cand + (u64)r * stride + prefix[(u64)r * n_slots + slot];
        for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
            u32 v = list[i];
            atomicOr(bitmap + (u64)slot * words_per_slot + (v >> 5), 1u << (v & 31));
        }
    }
}

__global__ void k3_counts_kernel(const u64 *__restrict__ cand_off, u32 n_slots, u32 *counts) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_slots) counts[i] = (u32)(cand_off[i + 1] - cand_off[i]);
}

// ---- matching order: generateGQLQueryPlan (custom.h:670-722) + generateBN (:724-755) -----------------------
__device__ bool q_edge(const u32 *off, const u32 *nbr, u32 u, u32 v) {
    for (u32 j = off[u]; j < off[u + 1]; j++)
        if (nbr[j] == v) return true;
    return false;
}

__global__ void __launch_bounds__(256) k3_order_kernel(u32 n_queries, u32 V, const u32 *__restrict__ q_vbase,
                                                       const u32 *__restrict__ q_ebase,
                                                       const u32 *__restrict__ q_offsets,
                                                       const u32 *__restrict__ q_nbrs, const u32 *__restrict__ q_labels,
                                                       const u64 *__restrict__ cand_off, u32 *order, u32 *pivot,
                                                       JoinDepth *jplan, u64 *item_base, u32 rank, u32 world) {
    for (u32 q = threadIdx.x; q < n_queries; q += blockDim.x) {
        const u32 vb = q_vbase[q], nq = q_vbase[q + 1] - vb;
        const u32 *off = q_offsets + vb + q;  // nq + 1 local offsets
        const u32 *nbr = q_nbrs + q_ebase[q];
        const u64 *co = cand_off + vb;
        u32 *ord = order + vb, *piv = pivot + vb;
        auto count = [&](u32 u) { return (u32)(co[u + 1] - co[u]); };
        auto qdeg = [&](u32 u) { return off[u + 1] - off[u]; };
        u64 items = 0;
        if (nq > 0) {
            u32 start = 0;  // selectGQLStartVertex, custom.h:635-654
            for (u32 i = 1; i < nq; i++) {
                if (count(i) < count(start)) start = i;
                else if (count(i) == count(start) && qdeg(i) > qdeg(start)) start = i;
            }
            u64 visited = 0, adjacent = 0;
            auto mark = [&](u32 u) {
                visited |= 1ull << u;
                for (u32 j = off[u]; j < off[u + 1]; j++) adjacent |= 1ull << nbr[j];
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
            }
            u32 depth_of[kMaxNQ];
            for (u32 i = 0; i < nq; i++) depth_of[ord[i]] = i;
            u32 pvd[kMaxNQ], lab[kMaxNQ];
            bool fast[kMaxNQ];
            for (u32 i = 0; i < nq; i++) {
                u32 u = ord[i];
                JoinDepth jd;
                jd.u = u;
                jd.label = q_labels[vb + u];
                jd.deg = qdeg(u);
                jd.pivot_depth = 0;
                jd.bn_mask = 0;
                jd.same_mask = 0;
                jd.tail_k = 0;
                jd.tail_mode = 0;
                if (i > 0) {
                    u32 pv = 0xffffffffu;
                    for (u32 j = 0; j < i; j++)
                        if (q_edge(off, nbr, u, ord[j])) { pv = ord[j]; break; }
                    piv[i] = pv;
                    jd.pivot_depth = pv == 0xffffffffu ? 0 : depth_of[pv];
                    for (u32 j = off[u]; j < off[u + 1]; j++) {
                        u32 w = nbr[j];
                        if (depth_of[w] < i && w != pv) jd.bn_mask |= 1ull << depth_of[w];
                    }
                    // depths whose vertex can equal a candidate of this depth: same query label; depth 0 always,
                    // because caller-supplied start candidates (gpe_refine) are not label-checked
                    jd.same_mask = 1ull;
                    for (u32 j = 1; j < i; j++)
                        if (lab[j] == jd.label) jd.same_mask |= 1ull << j;
                }
                lab[i] = jd.label;
                pvd[i] = jd.pivot_depth;
                fast[i] = i > 0 && jd.bn_mask == 0 && jd.deg <= 1;
                jplan[vb + i] = jd;
            }
            // trailing leaves that can be counted instead of walked (see "the join")
            u32 k = 0;
            while (k + 1 < nq) {
                u32 t = nq - 1 - k;  // candidate new tail member; the prefix would be depths [0, t)
                bool ok = fast[t];
                for (u32 i = t; i < nq && ok; i++) ok = pvd[i] < t;
                for (u32 i = t + 1; i < nq && ok; i++) ok = lab[i] != lab[t];
                if (!ok) break;
                k++;
            }
            u32 mode = k ? 1 : 0;
            if (k == 1 && nq >= 3) {
                u32 t = nq - 2;
                if (fast[t] && pvd[t] < t && pvd[nq - 1] < t && lab[t] == lab[nq - 1]) { k = 2; mode = 2; }
            }
            jplan[vb].tail_k = k;
            jplan[vb].tail_mode = mode;
            u64 total = count(start);
            items = total > rank ? (total - rank + world - 1) / world : 0;
        }
        item_base[q + 1] = items;  // turned into a prefix below
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 run = 0;
        item_base[0] = 0;
        for (u32 q = 0; q < n_queries; q++) {
            run += item_base[q + 1];
            item_base[q + 1] = run;
        }
    }
}

// ---- the join ------------------------------------------------------------------------------------------------
//
// Work item = a partial embedding (the first `depth` vertices of the matching order) plus a range [lo, hi)
// of the candidate segment for the next vertex.  One THREAD runs one item as an explicit-stack DFS (stack in
// shared memory, [level][thread] so accesses never bank-conflict); a candidate segment is the label group of
// the pivot's adjacency (JoinGraph::nbrL / gtab), so a step touches only neighbours that already carry the
// right label.  Parallelism in the reference's own decomposition is small (|C(order[0])| start candidates)
// and subtree sizes are heavy-tailed, so every item gets a step budget: a thread that exhausts it writes its
// continuation -- for every stack level the unexplored sibling range is an independent subtree -- as new
// items for the next round.  Rounds are plain bounded kernel launches: no spinning, no device-side queue.
//
// Counting shortcut (results unchanged): when the last k vertices of the matching order are leaves of the
// query that only constrain label (query degree 1, no backward neighbour besides the pivot, pivot in the
// first n-k vertices) and carry pairwise different labels, the completions of an (n-k)-prefix are counted as
// the product of the sizes of their label groups minus the members already used, instead of being walked.
// Two trailing leaves with the SAME label are handled by |A||B| - |A n B|.
struct JoinGraph {
    const u32 *off, *nbr, *deg;  // id-sorted CSR: edge tests (graph.h:215-236)
    const u32 *label;
    const u32 *nbrL;             // the same adjacency grouped by neighbour label, ascending id inside a group
    const u32 *gtab;             // V x (nl+1): start of every label group of every vertex (absolute, into nbrL)
    u32 V, nl;
};

constexpr int kItemHdr = 4;  // q, depth, lo, hi
constexpr u32 kSplit = 8;
constexpr int kDfsThreads = 256;

__device__ __forceinline__ bool has_edge(const JoinGraph &g, u32 u, u32 v) {
    // graph.h:215-236: search for the larger-degree endpoint in the smaller list
    u32 du = g.deg[u], dv = g.deg[v];
    if (du < dv) { u32 t = u; u = v; v = t; dv = du; }
    const u32 *a = g.nbr + g.off[v];
    int lo = 0, hi = (int)dv - 1;
    while (lo <= hi) {
        int mid = lo + ((hi - lo) >> 1);
        u32 x = a[mid];
        if (x == u) return true;
        if (x > u) hi = mid - 1; else lo = mid + 1;
    }
    return false;
}

__device__ __forceinline__ void group_range(const JoinGraph &g, u32 v, u32 label, u32 &lo, u32 &hi) {
    if (label >= g.nl) { lo = hi = 0; return; }
    const u32 *row = g.gtab + (u64)v * (g.nl + 1) + label;
    lo = row[0];
    hi = row[1];
}

// depth-1 items from the start candidates of this shard (idx % world == rank)
__global__ void __launch_bounds__(256) k3_init_items_kernel(JoinGraph g, u32 n_queries, const u32 *__restrict__ q_vbase,
                                                            const JoinDepth *__restrict__ jplan,
                                                            const u64 *__restrict__ cand_off, const u32 *__restrict__ cand,
                                                            const u64 *__restrict__ item_base, u32 rank, u32 world,
                                                            u32 *items, u32 stride, u64 *answers, u64 *n_items_out) {
    const u64 n_items = item_base[n_queries];
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_items_out = n_items;
    for (u64 item = (u64)blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += (u64)gridDim.x * blockDim.x) {
        u32 lo = 0, hi = n_queries;
        while (hi - lo > 1) {
            u32 mid = (lo + hi) >> 1;
            if (item_base[mid] <= item) lo = mid; else hi = mid;
        }
        const u32 q = lo, vb = q_vbase[q], nq = q_vbase[q + 1] - vb;
        const u64 idx = (item - item_base[q]) * world + rank;
        const u32 v0 = cand[cand_off[vb + jplan[vb].u] + idx];
        u32 *it = items + item * stride;
        it[0] = q;
        it[1] = 1;
        it[kItemHdr] = v0;
        if (nq == 1) {
            it[2] = it[3] = 0;  // nothing below the start vertex: the candidate itself is the match
            atomicAdd((unsigned long long *)&answers[q], 1ull);
        } else {
            u32 s, e;
            group_range(g, v0, jplan[vb + 1].label, s, e);  // pivot of depth 1 is always the start vertex
            it[2] = s;
            it[3] = e;
        }
    }
}

template <int MAXNQ>
__global__ void __launch_bounds__(kDfsThreads) k3_dfs_kernel(JoinGraph g, const u32 *__restrict__ q_vbase,
                                                             const JoinDepth *__restrict__ jplan,
                                                             const u64 *__restrict__ limits, u64 *answers,
                                                             const u32 *__restrict__ items_in,
                                                             const u64 *__restrict__ n_in_ptr, u32 *items_out,
                                                             u64 *out_count, u64 out_cap, u64 *fetch_counter, u32 budget,
                                                             u32 *matches, u64 matches_cap, u64 *match_cursor,
                                                             u64 *step_counter) {
    constexpr u32 stride = MAXNQ + kItemHdr;
    extern __shared__ u32 s_stack[];  // emb | cur | end, each [MAXNQ][kDfsThreads]
    u32 *emb = s_stack + threadIdx.x;
    u32 *cur = emb + MAXNQ * kDfsThreads;
    u32 *end = cur + MAXNQ * kDfsThreads;
#define EMB(t) emb[(t) * kDfsThreads]
#define CUR(t) cur[(t) * kDfsThreads]
#define END(t) end[(t) * kDfsThreads]
    const int lane = threadIdx.x & 31;
    const u64 n_in = *n_in_ptr;
    bool have = false, exhausted = false;
    u32 q = 0, vb = 0, nq = 0, base = 0, d = 0, steps = 0, lab0 = 0, tail_at = 0, tail_mode = 0;
    u32 acc_q = 0xffffffffu;
    u64 acc = 0, my_steps = 0;

    for (;;) {
        // ---- fetch: lanes without an item claim consecutive indices with one atomic per warp ----
        unsigned need = __ballot_sync(kFull, !have && !exhausted);
        if (need) {
            int leader = __ffs(need) - 1;
            u64 b = 0;
            if (lane == leader) b = atomicAdd((unsigned long long *)fetch_counter, (unsigned long long)__popc(need));
            b = __shfl_sync(kFull, b, leader);
            if (!have && !exhausted) {
                u64 idx = b + __popc(need & lanemask_lt());
                if (idx >= n_in) {
                    exhausted = true;
                } else {
                    const u32 *it = items_in + idx * stride;
                    u32 iq = it[0];
                    if (iq != acc_q) {
                        if (acc) atomicAdd((unsigned long long *)&answers[acc_q], (unsigned long long)acc);
                        acc = 0;
                        acc_q = iq;
                    }
                    u64 limit = limits ? limits[iq] : GPE_LIMIT_MAX;
                    if (limit == 0) limit = 1;  // the reference tests the limit only after counting a match (:851)
                    q = iq;
                    vb = q_vbase[q];
                    nq = q_vbase[q + 1] - vb;
                    base = it[1];
                    if (base < nq && it[2] < it[3] && *(volatile u64 *)&answers[q] < limit) {
                        for (u32 t = 0; t < base; t++) EMB(t) = it[kItemHdr + t];
                        lab0 = g.label[it[kItemHdr]];  // caller-supplied start candidates need not carry the query label
                        const JoinDepth j0 = jplan[vb];
                        tail_mode = matches ? 0 : j0.tail_mode;
                        tail_at = tail_mode ? nq - j0.tail_k : nq;  // depth at which the counting shortcut takes over
                        d = base;
                        CUR(d) = it[2];
                        END(d) = it[3];
                        steps = 0;
                        have = true;
                    }
                }
            }
        }
        if (__ballot_sync(kFull, have) == 0) break;
        if (!have) continue;

        // ---- one DFS step: test the next candidate of level d ----
        my_steps++;
        {
            const u32 at = CUR(d);
            const u32 c = g.nbrL[at];
            CUR(d) = at + 1;
            const JoinDepth jd = jplan[vb + d];
            bool ok = jd.deg <= 1 || g.deg[c] >= jd.deg;  // a neighbour always has degree >= 1
            u64 sm = jd.same_mask;                          // earlier depths that can hold a vertex of this label
            while (ok && sm) {
                int t = __ffsll((long long)sm) - 1;
                sm &= sm - 1;
                ok = EMB(t) != c;
            }
            u64 bn = jd.bn_mask;
            while (ok && bn) {
                int t = __ffsll((long long)bn) - 1;
                bn &= bn - 1;
                ok = has_edge(g, c, EMB(t));
            }
            if (ok) {
                if (d == nq - 1) {
                    acc++;
                    if (matches) {
                        u64 pos = atomicAdd((unsigned long long *)match_cursor, 1ull);
                        if (pos < matches_cap) {
                            u32 *row = matches + pos * nq;
                            for (u32 t = 0; t < d; t++) row[jplan[vb + t].u] = EMB(t);
                            row[jd.u] = c;
                        }
                    }
                } else {
                    EMB(d) = c;
                    const u32 nd = d + 1;
                    if (nd == tail_at) {
                        // counting shortcut over the trailing leaves nd .. nq-1
                        u64 total = 1;
                        u32 sz[2] = {0, 0}, gs[2] = {0, 0}, ge[2] = {0, 0}, pv[2] = {0, 0};
                        for (u32 i = nd; i < nq && total; i++) {
                            const JoinDepth ld = jplan[vb + i];
                            const u32 p = EMB(ld.pivot_depth);
                            u32 s, e;
                            group_range(g, p, ld.label, s, e);
                            u32 used = 0;  // prefix vertices that sit in this group
                            u64 m = ld.same_mask & ((1ull << nd) - 1);
                            while (m) {
                                int t = __ffsll((long long)m) - 1;
                                m &= m - 1;
                                if ((u32)t != ld.pivot_depth && (t != 0 || lab0 == ld.label) && has_edge(g, p, EMB(t))) used++;
                            }
                            const u32 n_free = (e - s) - used;
                            if (tail_mode == 2) {
                                sz[i - nd] = n_free; gs[i - nd] = s; ge[i - nd] = e; pv[i - nd] = p;
                            } else {
                                total *= n_free;
                            }
                        }
                        if (tail_mode == 2) {
                            // both leaves carry the same label: ordered pairs of distinct vertices
                            u64 inter;
                            if (pv[0] == pv[1]) {
                                inter = sz[0];
                            } else {
                                inter = 0;  // |G(p0) n G(p1)| without the used members: merge two ascending id lists
                                u32 x = gs[0], y = gs[1];
                                while (x < ge[0] && y < ge[1]) {
                                    const u32 vx = g.nbrL[x], vy = g.nbrL[y];
                                    if (vx == vy) {
                                        bool is_used = false;
                                        for (u32 t = 0; t < nd; t++) is_used = is_used || EMB(t) == vx;
                                        inter += is_used ? 0 : 1;
                                        x++;
                                        y++;
                                    } else if (vx < vy) {
                                        x++;
                                    } else {
                                        y++;
                                    }
                                }
                            }
                            total = (u64)sz[0] * sz[1] - inter;
                        }
                        acc += total;
                    } else {
                        const JoinDepth ndj = jplan[vb + nd];
                        u32 s, e;
                        group_range(g, EMB(ndj.pivot_depth), ndj.label, s, e);
                        d = nd;
                        CUR(d) = s;
                        END(d) = e;
                    }
                }
            }
        }
        // ---- pop exhausted levels ----
        while (have && CUR(d) >= END(d)) {
            if (d == base) have = false; else d--;
        }

        // ---- budget: hand the unexplored sibling ranges of every stack level to the next round ----
        if (have && ++steps >= budget) {
            u32 pieces = 0;
            for (u32 l = base; l <= d; l++) pieces += min(END(l) - CUR(l), kSplit);
            u64 o = atomicAdd((unsigned long long *)out_count, (unsigned long long)pieces);
            if (o + pieces > out_cap) {
                atomicAdd((unsigned long long *)out_count, (unsigned long long)(0ull - pieces));  // undo; keep running
                steps = 0;
            } else {
                for (u32 l = base; l <= d; l++) {
                    u32 len = END(l) - CUR(l), np = min(len, kSplit);
                    for (u32 k = 0; k < np; k++) {
                        u32 *it = items_out + o * stride;
                        it[0] = q;
                        it[1] = l;
                        it[2] = CUR(l) + (u32)((u64)len * k / np);
                        it[3] = CUR(l) + (u32)((u64)len * (k + 1) / np);
                        for (u32 t = 0; t < l; t++) it[kItemHdr + t] = EMB(t);
                        o++;
                    }
                }
                have = false;
            }
        }
    }
    if (acc) atomicAdd((unsigned long long *)&answers[acc_q], (unsigned long long)acc);
    for (int o = 16; o; o >>= 1) my_steps += __shfl_xor_sync(kFull, my_steps, o);
    if (lane == 0 && my_steps) atomicAdd((unsigned long long *)step_counter, (unsigned long long)my_steps);
#undef EMB
#undef CUR
#undef END
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

cudaError_t k3_compact(const u32 *bitmap, u64 words_per_slot, u64 chunks_per_slot, u32 n_slots, const u64 *chunk_off,
                       u32 *cand, u64 *cand_off, cudaStream_t s) {
    u64 n_chunks = chunks_per_slot * n_slots;
    if (n_chunks == 0) return cudaSuccess;
    k3_compact_kernel<<<(unsigned)((n_chunks * 32 + 255) / 256), 256, 0, s>>>(bitmap, words_per_slot, chunks_per_slot,
                                                                             n_chunks, n_slots, chunk_off, cand,
                                                                             cand_off);
    return cudaGetLastError();
}

cudaError_t k3_scatter(const u32 *counts, const u32 *cand, u64 stride, u32 world, u32 n_slots, u32 *bitmap,
                       u64 words_per_slot, u64 *prefix_tmp, cudaStream_t s) {
    if (world * n_slots == 0) return cudaSuccess;
    k3_scatter_prefix_kernel<<<(world + 31) / 32, 32, 0, s>>>(counts, world, n_slots, prefix_tmp);
    unsigned blocks = std::min<unsigned>(world * n_slots, 148 * 8);
    k3_scatter_kernel<<<blocks, 256, 0, s>>>(counts, cand, prefix_tmp, stride, world, n_slots, bitmap, words_per_slot);
    return cudaGetLastError();
}

cudaError_t k3_counts_from_offsets(const u64 *cand_off, u32 n_slots, u32 *counts, cudaStream_t s) {
    if (n_slots == 0) return cudaSuccess;
    k3_counts_kernel<<<(n_slots + 255) / 256, 256, 0, s>>>(cand_off, n_slots, counts);
    return cudaGetLastError();
}

cudaError_t k3_order(u32 n_queries, u32 V, const u32 *q_vbase, const u32 *q_ebase, const u32 *q_offsets,
                     const u32 *q_nbrs, const u32 *q_labels, const u64 *cand_off, u32 *order, u32 *pivot,
                     JoinDepth *jplan, u64 *item_base, u32 rank, u32 world, cudaStream_t s) {
    k3_order_kernel<<<1, 256, 0, s>>>(n_queries, V, q_vbase, q_ebase, q_offsets, q_nbrs, q_labels, cand_off, order,
                                      pivot, jplan, item_base, rank, world);
    return cudaGetLastError();
}

u32 k3_item_stride(u32 max_nq) {
    u32 m = max_nq <= 8 ? 8 : max_nq <= 16 ? 16 : max_nq <= 32 ? 32 : 64;
    return m + kItemHdr;
}

cudaError_t k3_init_items(const JoinView &jv, u32 n_queries, const u32 *q_vbase, const JoinDepth *jplan,
                          const u64 *cand_off, const u32 *cand, const u64 *item_base, u32 rank, u32 world, u32 *items,
                          u32 stride, u64 *answers, u64 *n_items_out, int sm_count, cudaStream_t s) {
    JoinGraph g{jv.off, jv.nbr, jv.deg, jv.label, jv.nbrL, jv.gtab, jv.V, jv.nl};
    k3_init_items_kernel<<<sm_count * 4, 256, 0, s>>>(g, n_queries, q_vbase, jplan, cand_off, cand, item_base, rank, world,
                                                     items, stride, answers, n_items_out);
    return cudaGetLastError();
}

cudaError_t k3_dfs_round(const JoinView &jv, u32 max_nq, const u32 *q_vbase, const JoinDepth *jplan, const u64 *limits,
                         u64 *answers, const u32 *items_in, const u64 *n_in, u32 *items_out, u64 *out_count, u64 out_cap,
                         u64 *fetch_counter, u32 budget, u32 *matches, u64 matches_cap, u64 *match_cursor,
                         u64 *step_counter, int sm_count, cudaStream_t s) {
    JoinGraph g{jv.off, jv.nbr, jv.deg, jv.label, jv.nbrL, jv.gtab, jv.V, jv.nl};
#define LAUNCH(M)                                                                                                     \
    static int per_sm_##M = 0;                                                                                        \
    const size_t smem_##M = (size_t)3 * M * kDfsThreads * sizeof(u32);                                                \
    if (!per_sm_##M) {                                                                                                \
        cudaFuncSetAttribute(k3_dfs_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_##M);           \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_##M, k3_dfs_kernel<M>, kDfsThreads, smem_##M) !=    \
                cudaSuccess || per_sm_##M < 1)                                                                        \
            per_sm_##M = 1;                                                                                           \
    }                                                                                                                 \
    k3_dfs_kernel<M><<<sm_count * per_sm_##M, kDfsThreads, smem_##M, s>>>(g, q_vbase, jplan, limits, answers, items_in, n_in, items_out, out_count, \
                                            out_cap, fetch_counter, budget, matches, matches_cap, match_cursor, step_counter)
    if (max_nq <= 8) { LAUNCH(8); }
    else if (max_nq <= 16) { LAUNCH(16); }
    else if (max_nq <= 32) { LAUNCH(32); }
    else { LAUNCH(64); }
#undef LAUNCH
    return cudaGetLastError();
}

}  // namespace gpe
