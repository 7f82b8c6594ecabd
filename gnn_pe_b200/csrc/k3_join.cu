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
            for (u32 i = 0; i < nq; i++) {
                u32 u = ord[i];
                JoinDepth jd;
                jd.u = u;
                jd.label = q_labels[vb + u];
                jd.deg = qdeg(u);
                jd.pivot_depth = 0;
                jd.bn_mask = 0;
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
                }
                jplan[vb + i] = jd;
            }
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
// of the candidate segment for the next vertex.  One THREAD runs one item as an explicit-stack DFS; a
// candidate segment is the label group of the pivot's adjacency (JoinGraph::nbrL / gtab), so a step touches
// only neighbours that already carry the right label.  Parallelism in the reference's own decomposition is
// tiny (|C(order[0])| start candidates, often < 100) and subtree sizes are heavy-tailed, so every item gets
// a step budget: a thread that exhausts it writes its continuation -- for every stack level the unexplored
// sibling range is an independent subtree -- as new items for the next round.  Rounds are plain bounded
// kernel launches: no spinning, no device-side queue.
struct JoinGraph {
    const u32 *off, *nbr, *deg;  // id-sorted CSR: edge tests (graph.h:215-236)
    const u32 *label;
    const u32 *nbrL;             // the same adjacency grouped by neighbour label, ascending id inside a group
    const u32 *gtab;             // V x (nl+1): start of every label group of every vertex (absolute, into nbrL)
    u32 V, nl;
};

constexpr int kItemHdr = 4;  // q, depth, lo, hi
constexpr u32 kSplit = 8;

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
__global__ void __launch_bounds__(256) k3_dfs_kernel(JoinGraph g, const u32 *__restrict__ q_vbase,
                                                     const JoinDepth *__restrict__ jplan, const u64 *__restrict__ limits,
                                                     u64 *answers, const u32 *__restrict__ items_in,
                                                     const u64 *__restrict__ n_in_ptr, u32 *items_out, u64 *out_count,
                                                     u64 out_cap, u64 *fetch_counter, u32 budget, u32 *matches,
                                                     u64 matches_cap, u64 *match_cursor) {
    constexpr u32 stride = MAXNQ + kItemHdr;
    const int lane = threadIdx.x & 31;
    const u64 n_in = *n_in_ptr;
    u32 emb[MAXNQ], cur[MAXNQ], end[MAXNQ];
    bool have = false, exhausted = false;
    u32 q = 0, vb = 0, nq = 0, base = 0, d = 0, steps = 0, lab0 = 0;
    u32 acc_q = 0xffffffffu;
    u64 acc = 0;

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
                    if (base < nq && *(volatile u64 *)&answers[q] < limit) {
                        for (u32 t = 0; t < base; t++) emb[t] = it[kItemHdr + t];
                        lab0 = g.label[emb[0]];  // caller-supplied start candidates need not carry the query label
                        d = base;
                        cur[d] = it[2];
                        end[d] = it[3];
                        steps = 0;
                        have = true;
                    }
                }
            }
        }
        if (__ballot_sync(kFull, have) == 0) break;
        if (!have) continue;

        // ---- one DFS step ----
        if (cur[d] < end[d]) {
            const u32 c = g.nbrL[cur[d]++];
            const JoinDepth jd = jplan[vb + d];
            bool ok = jd.deg <= 1 || g.deg[c] >= jd.deg;  // a neighbour has degree >= 1
            for (u32 t = 0; ok && t < d; t++) ok = emb[t] != c;
            u64 bn = jd.bn_mask;
            while (ok && bn) {
                int t = __ffsll((long long)bn) - 1;
                bn &= bn - 1;
                ok = has_edge(g, c, emb[t]);
            }
            if (ok) {
                if (d == nq - 1) {
                    acc++;
                    if (matches) {
                        u64 pos = atomicAdd((unsigned long long *)match_cursor, 1ull);
                        if (pos < matches_cap) {
                            u32 *row = matches + pos * nq;
                            for (u32 t = 0; t < d; t++) row[jplan[vb + t].u] = emb[t];
                            row[jd.u] = c;
                        }
                    }
                } else {
                    emb[d] = c;
                    d++;
                    const JoinDepth nd = jplan[vb + d];
                    u32 s, e;
                    group_range(g, emb[nd.pivot_depth], nd.label, s, e);
                    if (d == nq - 1 && nd.bn_mask == 0 && nd.deg <= 1 && !matches) {
                        // leaf fast path: every member of the group matches unless it is already used; the used
                        // vertices inside the group are the embedded ones with this label adjacent to the pivot
                        u32 used = 0;
                        const u32 p = emb[nd.pivot_depth];
                        for (u32 t = 0; t < d; t++)
                            if (t != nd.pivot_depth && (t ? jplan[vb + t].label : lab0) == nd.label && has_edge(g, p, emb[t])) used++;
                        acc += (e - s) - used;
                        d--;
                    } else {
                        cur[d] = s;
                        end[d] = e;
                    }
                }
            }
        } else if (d == base) {
            have = false;
        } else {
            d--;
        }

        // ---- budget: hand the unexplored sibling ranges of every stack level to the next round ----
        if (have && ++steps >= budget) {
            u32 pieces = 0;
            for (u32 l = base; l <= d; l++) pieces += min(end[l] - cur[l], kSplit);
            if (pieces == 0) {
                have = false;
            } else {
                u64 o = atomicAdd((unsigned long long *)out_count, (unsigned long long)pieces);
                if (o + pieces > out_cap) {
                    atomicAdd((unsigned long long *)out_count, (unsigned long long)(0ull - pieces));  // undo; keep running
                    steps = 0;
                } else {
                    for (u32 l = base; l <= d; l++) {
                        u32 len = end[l] - cur[l], np = min(len, kSplit);
                        for (u32 k = 0; k < np; k++) {
                            u32 *it = items_out + o * stride;
                            it[0] = q;
                            it[1] = l;
                            it[2] = cur[l] + (u32)((u64)len * k / np);
                            it[3] = cur[l] + (u32)((u64)len * (k + 1) / np);
                            for (u32 t = 0; t < l; t++) it[kItemHdr + t] = emb[t];
                            o++;
                        }
                    }
                    have = false;
                }
            }
        }
    }
    if (acc) atomicAdd((unsigned long long *)&answers[acc_q], (unsigned long long)acc);
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
                         u64 *fetch_counter, u32 budget, u32 *matches, u64 matches_cap, u64 *match_cursor, int sm_count,
                         cudaStream_t s) {
    JoinGraph g{jv.off, jv.nbr, jv.deg, jv.label, jv.nbrL, jv.gtab, jv.V, jv.nl};
#define LAUNCH(M)                                                                                                     \
    static int per_sm_##M = 0;                                                                                        \
    if (!per_sm_##M) {                                                                                                \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_##M, k3_dfs_kernel<M>, 256, 0) != cudaSuccess ||    \
            per_sm_##M < 1)                                                                                           \
            per_sm_##M = 4;                                                                                           \
    }                                                                                                                 \
    k3_dfs_kernel<M><<<sm_count * per_sm_##M, 256, 0, s>>>(g, q_vbase, jplan, limits, answers, items_in, n_in, items_out, out_count, \
                                            out_cap, fetch_counter, budget, matches, matches_cap, match_cursor)
    if (max_nq <= 8) { LAUNCH(8); }
    else if (max_nq <= 16) { LAUNCH(16); }
    else if (max_nq <= 32) { LAUNCH(32); }
    else { LAUNCH(64); }
#undef LAUNCH
    return cudaGetLastError();
}

}  // namespace gpe
