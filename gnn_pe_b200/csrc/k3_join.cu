// k3_join.cu -- candidate compaction, matching order and the backtracking join (hot path 3).
//
// Replaces main.cpp:166-172 (std::set merge), refinement (custom.h:890-932), generateGQLQueryPlan
// (:670-722), generateBN (:724-755), generateValidCandidates (:757-797) and exploreQuickSIStyle (:799-888).
//
//   * candidate bitmaps -> sorted duplicate-free lists: popcount per chunk, device-wide scan, expand
//     (a std::set iterates ascending; so does a bitmap);
//   * matching order: one thread per query, the reference's greedy rule restated on bitmasks;
//   * join: one warp per (query, start candidate).  The reference's per-depth candidate buffers
//     (valid_candidate[depth], sized by max label frequency) are not materialised: a depth keeps a cursor
//     into the pivot's adjacency list and a 32-bit mask of the lanes of the current 32-wide chunk that
//     passed label / degree / not-yet-used / backward-edge tests, so a warp's whole DFS state is a few
//     hundred bytes of shared memory.  Edge tests are the reference's binary search in the shorter
//     adjacency list (graph.h:215-236).  As in the reference, candidate sets are only used for the start
//     vertex and for ordering.
//
// Latency / divergence bound (L2-resident CSR gathers), no bandwidth roofline, no tensor cores.
#include "gpe_internal.h"

namespace gpe {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kJoinWarps = 8;
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
struct WarpState {
    u32 emb[kMaxNQ];
    u32 pos[kMaxNQ];
    u32 msk[kMaxNQ];
    u32 lab[kMaxNQ];
    u32 deg[kMaxNQ];
    u32 pvd[kMaxNQ];
    u32 qv[kMaxNQ];
    u64 bn[kMaxNQ];
};

__device__ __forceinline__ bool has_edge(const GraphView &g, u32 u, u32 v) {
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

__global__ void __launch_bounds__(kJoinWarps * 32) k3_join_kernel(
    GraphView g, u32 n_queries, const u32 *__restrict__ q_vbase, const JoinDepth *__restrict__ jplan,
    const u64 *__restrict__ cand_off, const u32 *__restrict__ cand, const u64 *__restrict__ item_base,
    const u64 *__restrict__ limits, u64 *answers, u64 *work_counter, u32 rank, u32 world, u32 *matches,
    u64 matches_cap, u64 *match_cursor) {
    __shared__ WarpState s_state[kJoinWarps];
    WarpState &st = s_state[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const u64 n_items = item_base[n_queries];

    while (true) {
        u64 item = 0;
        if (lane == 0) item = atomicAdd((unsigned long long *)work_counter, 1ull);
        item = __shfl_sync(kFull, item, 0);
        if (item >= n_items) break;
        u32 lo = 0, hi = n_queries;  // last q with item_base[q] <= item
        while (hi - lo > 1) {
            u32 mid = (lo + hi) >> 1;
            if (item_base[mid] <= item) lo = mid; else hi = mid;
        }
        const u32 q = lo;
        const u32 vb = q_vbase[q], nq = q_vbase[q + 1] - vb;
        u64 limit = limits ? limits[q] : GPE_LIMIT_MAX;
        if (limit == 0) limit = 1;  // the reference tests the limit only after counting a match (:851)
        if (*(volatile u64 *)&answers[q] >= limit) continue;

        __syncwarp();
        for (u32 d = lane; d < nq; d += 32) {
            JoinDepth jd = jplan[vb + d];
            st.lab[d] = jd.label;
            st.deg[d] = jd.deg;
            st.pvd[d] = jd.pivot_depth;
            st.qv[d] = jd.u;
            st.bn[d] = jd.bn_mask;
        }
        __syncwarp();
        const u64 idx = (item - item_base[q]) * world + rank;
        const u32 v0 = cand[cand_off[vb + st.qv[0]] + idx];
        u64 found = 0;
        if (nq == 1) {
            found = 1;
            if (matches && lane == 0) {
                u64 p = atomicAdd((unsigned long long *)match_cursor, 1ull);
                if (p < matches_cap) matches[p] = v0;
            }
        } else {
            if (lane == 0) { st.emb[0] = v0; st.pos[1] = 0; st.msk[1] = 0; }
            __syncwarp();
            int d = 1;
            while (d >= 1) {
                u32 m = st.msk[d];
                const u32 p = st.emb[st.pvd[d]];
                const u32 beg = g.off[p], end = g.off[p + 1];
                if (m == 0) {
                    const u32 at = beg + st.pos[d];
                    if (at >= end) { d--; continue; }
                    const u32 i = at + lane;
                    bool ok = false;
                    u32 c = 0;
                    if (i < end) {
                        c = g.nbr[i];
                        ok = g.label[c] == st.lab[d] && g.deg[c] >= st.deg[d];
                        for (int t = 0; ok && t < d; t++) ok = st.emb[t] != c;
                        u64 bn = st.bn[d];
                        while (ok && bn) {
                            int t = __ffsll((long long)bn) - 1;
                            bn &= bn - 1;
                            ok = has_edge(g, c, st.emb[t]);
                        }
                    }
                    m = __ballot_sync(kFull, ok);
                    __syncwarp();
                    if (lane == 0) st.pos[d] += 32;
                    if (d == (int)nq - 1) {
                        if (m) {
                            found += __popc(m);
                            if (matches) {
                                u64 mb = 0;
                                if (lane == 0) mb = atomicAdd((unsigned long long *)match_cursor, (unsigned long long)__popc(m));
                                mb = __shfl_sync(kFull, mb, 0);
                                u64 mine = mb + __popc(m & lanemask_lt());
                                if (ok && mine < matches_cap) {
                                    u32 *row = matches + mine * nq;
                                    for (int t = 0; t < d; t++) row[st.qv[t]] = st.emb[t];
                                    row[st.qv[d]] = c;
                                }
                            }
                        }
                        __syncwarp();
                        continue;
                    }
                    if (lane == 0) st.msk[d] = m;
                    __syncwarp();
                    if (m == 0) continue;
                }
                // descend into the lowest remaining lane of this chunk
                const int b = __ffs(m) - 1;
                const u32 c = g.nbr[beg + st.pos[d] - 32 + b];
                __syncwarp();
                if (lane == 0) {
                    st.msk[d] = m & (m - 1);
                    st.emb[d] = c;
                    st.pos[d + 1] = 0;
                    st.msk[d + 1] = 0;
                }
                __syncwarp();
                d++;
            }
        }
        if (lane == 0 && found) atomicAdd((unsigned long long *)&answers[q], (unsigned long long)found);
    }
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

cudaError_t k3_join(const GraphView &g, u32 n_queries, const u32 *q_vbase, const JoinDepth *jplan, const u64 *cand_off,
                    const u32 *cand, const u64 *item_base, const u64 *limits, u64 *answers, u64 *work_counter,
                    u32 rank, u32 world, u32 *matches, u64 matches_cap, u64 *match_cursor, int sm_count,
                    cudaStream_t s) {
    k3_join_kernel<<<sm_count * 4, kJoinWarps * 32, 0, s>>>(g, n_queries, q_vbase, jplan, cand_off, cand, item_base,
                                                           limits, answers, work_counter, rank, world, matches,
                                                           matches_cap, match_cursor);
    return cudaGetLastError();
}

}  // namespace gpe
