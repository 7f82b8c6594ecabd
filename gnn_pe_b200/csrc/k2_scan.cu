// k2_scan.cu -- the dominance scan (hot path 1).
//
// Replaces Partition::query (custom.h:366-489): the best-first R*-tree traversal becomes a streaming
// compare of query-path blocks against tiles of the structure-of-arrays path table.
//   * the per-row test is the reference's leaf compare, custom.h:407-435, in FP64:
//       accept iff for every position k: q.label[k] == p.label[k] and q.degree[k] <= p.degree[k]
//              and for every dimension d: not (q.pde[d] > p.pde[d] and |q.pde[d] - p.pde[d]| > 1e-6)
//     (q > p  =>  |q-p| = q-p, so the second line is  not (q.pde[d] - p.pde[d] > 1e-6));
//   * internal-node routing (custom.h:439-484) becomes k2_select: a (block, tile) pair is dropped when the
//     tile's label range, max degrees or max-corner rule out every plan path of the block.  Pruning never
//     changes the result, only the number of tiles read;
//   * candidates[q.vids[k]].insert(p.vids[k]) (custom.h:429-432) becomes a test-and-set in a per-
//     (query vertex) bitmap over data vertices.
//
// HBM-bound: 72 B/row at l=2,e=2.  Each CTA is a TMA pipeline: one producer lane issues
// cp.async.bulk copies of (tile, query-path block) into a kStages-deep shared-memory ring guarded by
// mbarriers; 256 consumer threads compare one row each.  No tensor cores: nothing here is a contraction.
#include <cstdlib>

#include "gpe_internal.h"

namespace gpe {

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---- mbarrier / bulk-copy PTX -------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64 *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ u64 make_evict_first_policy() {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load(void *dst, const void *src, u32 bytes, u64 *bar, u64 policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void bulk_load_nohint(void *dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

template <int L, int E, bool VIDS, int NSREQ = kStages>
struct ScanGeom {
    static constexpr int D = L * E;
    static constexpr int kTileBytes = kTileRows * (8 * L + 8 * D);
    static constexpr int kVidBytes = VIDS ? kTileRows * 4 * L : 0;  // the tile's vertex ids ride along when survivors are the rule
    static constexpr int kRecBytes = (int)sizeof(QBlockRec<L, E>);
    static constexpr int kStageBytes = ((kTileBytes + kVidBytes + kRecBytes + 127) / 128) * 128;
    static constexpr int kNumStages = (NSREQ * kStageBytes <= 200 * 1024) ? NSREQ
                                       : ((200 * 1024) / kStageBytes >= 2 ? (200 * 1024) / kStageBytes : 2);
    static constexpr int kSmemBytes = kNumStages * kStageBytes + 1024;  // ring + barriers/meta + alignment slack
};

// ---- tile selection -------------------------------------------------------------------------------------
template <int L, int E>
__global__ void __launch_bounds__(256) k2_select_kernel(TableView t, const QBlockRec<L, E> *__restrict__ qblocks,
                                                        const u32 *__restrict__ qb_t0,
                                                        const u64 *__restrict__ qb_prefix, u32 n_qblocks, u64 n_items,
                                                        bool prune, u64 *__restrict__ worklist, u64 *counters) {
    constexpr int D = L * E;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 base = (u64)blockIdx.x * blockDim.x; base < n_items; base += stride) {
        u64 idx = base + threadIdx.x;
        bool pass = false;
        u32 tile = 0, b = 0;
        if (idx < n_items) {
            u32 lo = 0, hi = n_qblocks;  // last b with qb_prefix[b] <= idx
            while (hi - lo > 1) {
                u32 mid = (lo + hi) >> 1;
                if (qb_prefix[mid] <= idx) lo = mid; else hi = mid;
            }
            b = lo;
            tile = qb_t0[b] + (u32)(idx - qb_prefix[b]);
            if (!prune || t.ids_only) {  // (an ids-only table keeps no per-tile summaries: the bucket directory is the pruning)
                pass = true;
            } else {
                const QBlockRec<L, E> &rec = qblocks[b];
                for (u32 j = 0; j < rec.n && !pass; j++) {
                    bool ok = true;
#pragma unroll
                    for (int k = 0; k < L; k++) {
                        u32 ql = rec.labels[j][k];
                        ok = ok && ql >= t.lab_min[k * t.n_tiles + tile] && ql <= t.lab_max[k * t.n_tiles + tile] &&
                             rec.degs[j][k] <= t.deg_max[k * t.n_tiles + tile];
                    }
                    if (ok) {
#pragma unroll
                        for (int d = 0; d < D; d++)
                            ok = ok && !(rec.pde[j][d] - t.pde_max[d * t.n_tiles + tile] > kEps);
                    }
                    pass = ok;
                }
            }
        }
        unsigned m = __ballot_sync(kFull, pass);
        if (m) {
            u64 wbase = 0;
            int leader = __ffs(m) - 1;
            if ((int)(threadIdx.x & 31) == leader)
                wbase = atomicAdd((unsigned long long *)&counters[0], (unsigned long long)__popc(m));
            wbase = __shfl_sync(kFull, wbase, leader);
            if (pass) worklist[wbase + __popc(m & lanemask_lt())] = ((u64)b << 32) | tile;
        }
    }
}

// ---- the scan ----------------------------------------------------------------------------------------------
// VIDS = true  (pruned work list: the tiles come from the plan paths' own label buckets, most rows pass the label
//               test and survivors are common): the tile's vertex ids are part of the stage, no dependent global load;
// VIDS = false (streaming, every row against every plan path: survivors are rare): vertex ids are fetched from
//               global memory by the few rows that need them.
// Survivors set their bits with fire-and-forget reductions (RED.OR): nothing in the consumer loop waits on L2.
template <int L, int E, bool VIDS, int NSREQ = kStages>
__global__ void __launch_bounds__(kTileRows + 32, 1)
k2_scan_kernel(TableView t, const QBlockRec<L, E> *__restrict__ qblocks, const u64 *__restrict__ worklist,
               const u64 *__restrict__ counters, u32 *__restrict__ bitmap, u64 words_per_slot,
               u64 *__restrict__ survivors, int red_mode, u32 stage_words) {
    using G = ScanGeom<L, E, VIDS, NSREQ>;
    constexpr int D = G::D;
    constexpr int NS = G::kNumStages;
    // (no manual re-alignment through an integer cast: it would hide the shared address space from the compiler and
    //  turn every read of the stage into a generic load; bulk copies need 16 bytes, the declaration asks for 128)
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *ring = smem;
    u64 *full_bar = reinterpret_cast<u64 *>(smem + NS * G::kStageBytes);
    u64 *empty_bar = full_bar + NS;
    u64 *meta = empty_bar + NS;  // work item of every stage
    // Survivor staging (pruned work lists): a CTA works on a CONTIGUOUS chunk of the work list, i.e. on one query-path
    // block for hundreds of tiles, and a slot's class-local bitmap is a few KB: survivors set their bits in a shared-
    // memory copy (ATOMS) and the CTA flushes the non-zero words with one RED each when the block changes.  Without it
    // the REDs of the c-position (one per surviving row, no runs to deduplicate) cost 0.26 of the scan's 0.74 ms.
    int *stg_off = reinterpret_cast<int *>(meta + NS);           // [kQB * L] word offset of (path j, position k), or -1
    u32 *stg_slot = reinterpret_cast<u32 *>(stg_off + kQB * L);  // [kQB * L] slot of every staged segment
    u32 *stg_used = stg_slot + kQB * L;                          // staged words in use
    u32 *stg = reinterpret_cast<u32 *>(smem + G::kSmemBytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kConsumerWarps = kTileRows / 32;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const u64 n_items = counters[0];
    // streaming: items dealt round-robin; pruned lists: one contiguous chunk per CTA (see the staging above)
    const u64 per = (n_items + gridDim.x - 1) / gridDim.x;
    const u64 first = VIDS ? min((u64)blockIdx.x * per, n_items) : (u64)blockIdx.x;
    const u64 step = VIDS ? 1 : (u64)gridDim.x;
    const u64 n_my = VIDS ? min(per, n_items - first) : (first < n_items ? (n_items - first + gridDim.x - 1) / gridDim.x : 0);
    const bool staging = VIDS && stage_words >= words_per_slot && red_mode == 0;
    if (staging) {
        for (u32 i = threadIdx.x; i < stage_words; i += blockDim.x) stg[i] = 0;
        if (threadIdx.x == 0) *stg_used = 0;
        __syncthreads();
    }

    if (warp == kConsumerWarps) {
        // ---------------- producer warp: lane 0 issues the copies, the warp prefetches work items ---------
        const u64 policy = make_evict_first_policy();
        int stage = 0;
        u32 phase = 0;
        for (u64 k0 = 0; k0 < n_my; k0 += 32) {
            u64 mine = (k0 + lane < n_my) ? worklist[first + (k0 + lane) * step] : 0;
            int cnt = (int)min((u64)32, n_my - k0);
            for (int i = 0; i < cnt; i++) {
                u64 item = __shfl_sync(kFull, mine, i);
                if (lane == 0) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    u32 tile = (u32)item, b = (u32)(item >> 32);
                    unsigned char *dst = ring + stage * G::kStageBytes;
                    meta[stage] = item;
                    mbar_arrive_expect_tx(&full_bar[stage], G::kTileBytes + G::kVidBytes + G::kRecBytes);
                    bulk_load(dst, t.tiles + (u64)tile * G::kTileBytes, G::kTileBytes, &full_bar[stage], policy);
                    if (VIDS)
                        bulk_load(dst + G::kTileBytes, t.vids + (u64)tile * L * kTileRows, G::kVidBytes, &full_bar[stage], policy);
                    bulk_load_nohint(dst + G::kTileBytes + G::kVidBytes, &qblocks[b], G::kRecBytes, &full_bar[stage]);
                }
                if (++stage == NS) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ---------------- consumers: one row per thread ---------------------------------------------------
        const u32 r = threadIdx.x;
        int stage = 0;
        u32 phase = 0;
        // survivor counters: lane j < kQB of every warp counts plan path j of the block the warp is working on
        u32 cur_b = 0xffffffffu, my_cnt = 0, my_qpath = 0;
        for (u64 k = 0; k < n_my; k++) {
            mbar_wait(&full_bar[stage], phase);
            const unsigned char *buf = ring + stage * G::kStageBytes;
            const u64 item = meta[stage];
            const u32 tile = (u32)item, b = (u32)(item >> 32);
            const u32 *s_lab = reinterpret_cast<const u32 *>(buf);
            const u32 *s_deg = reinterpret_cast<const u32 *>(buf + 4 * L * kTileRows);
            const double *s_pde = reinterpret_cast<const double *>(buf + 8 * L * kTileRows);
            const u32 *s_vid = reinterpret_cast<const u32 *>(buf + G::kTileBytes);
            const QBlockRec<L, E> *rec = reinterpret_cast<const QBlockRec<L, E> *>(buf + G::kTileBytes + G::kVidBytes);
            const bool valid = (u64)tile * kTileRows + r < t.n_rows;
            if (b != cur_b) {
                if (my_cnt) atomicAdd((unsigned long long *)&survivors[my_qpath], (unsigned long long)my_cnt);
                my_cnt = 0;
                cur_b = b;
                my_qpath = lane < kQB ? rec->qpath[lane] : 0;
                if (staging) {  // every consumer warp passes here at the same item: named barrier of the 256 consumers
                    const u32 wps = (u32)words_per_slot;
                    asm volatile("bar.sync 1, %0;" ::"n"(kTileRows) : "memory");  // the previous block's tiles are done
                    const u32 used = *stg_used;
                    for (u32 i = threadIdx.x; i < used; i += kTileRows) {
                        const u32 v = stg[i];
                        if (v) {
                            const u32 seg = i / wps;
                            atomicOr(bitmap + (u64)stg_slot[seg] * words_per_slot + (i - seg * wps), v);
                            stg[i] = 0;
                        }
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(kTileRows) : "memory");
                    if (threadIdx.x == 0) {  // the c-position first: its bits have no runs to deduplicate
                        const u32 cap = stage_words / wps;
                        u32 n_seg = 0;
                        for (int kk = L - 1; kk >= 0; kk--)
                            for (u32 j = 0; j < kQB; j++) {
                                int off = -1;
                                if (j < rec->n && n_seg < cap) {
                                    off = (int)(n_seg * wps);
                                    stg_slot[n_seg++] = rec->slot[j][kk];
                                }
                                stg_off[j * L + kk] = off;
                            }
                        *stg_used = n_seg * wps;
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(kTileRows) : "memory");
                }
            }

            u32 lab[L], dg[L];
            double pde[D];
#pragma unroll
            for (int kk = 0; kk < L; kk++) {
                lab[kk] = s_lab[kk * kTileRows + r];
                dg[kk] = s_deg[kk * kTileRows + r];
            }
#pragma unroll
            for (int d = 0; d < D; d++) pde[d] = s_pde[d * kTileRows + r];
            const u32 nq = rec->n;
            u32 okm = 0;  // plan paths of the block this row survives
            for (u32 j = 0; j < nq; j++) {
                // label/degree test and embedding test are each branch-free (broadcast shared-memory reads, no
                // serialising short-circuits); the FP64 part is skipped by rows that already failed -- in streaming
                // mode that is nearly all of them, and the FP64 pipe would otherwise bound the scan
                bool ok = valid;
#pragma unroll
                for (int kk = 0; kk < L; kk++) ok &= (rec->labels[j][kk] == lab[kk]) & (rec->degs[j][kk] <= dg[kk]);
                if (ok) {
#pragma unroll
                    for (int d = 0; d < D; d++) ok &= !(rec->pde[j][d] - pde[d] > kEps);
                }
                const unsigned m = __ballot_sync(kFull, ok);
                if (lane == (int)j) my_cnt += __popc(m);
                okm |= (u32)ok << j;
            }
            if (__any_sync(kFull, okm != 0)) {
                u32 v[L];
#pragma unroll
                for (int kk = 0; kk < L; kk++)
                    v[kk] = !okm ? 0xffffffffu
                                 : VIDS ? s_vid[kk * kTileRows + r] : t.vids[((u64)tile * L + kk) * kTileRows + r];
                // rows of one start vertex are neighbours in the table: a lane whose left neighbour sets the same bits
                // for the same vertex stays silent
                const u32 okm_left = __shfl_up_sync(kFull, okm, 1);
#pragma unroll
                for (int kk = 0; kk < L; kk++) {
                    const u32 v_left = __shfl_up_sync(kFull, v[kk], 1);
                    u32 todo = okm;
                    if (lane > 0 && v_left == v[kk]) todo &= ~okm_left;
                    while (todo) {
                        const int j = __ffs(todo) - 1;
                        todo &= todo - 1;
                        u32 *w = bitmap + (u64)rec->slot[j][kk] * words_per_slot + (v[kk] >> 5);
                        const u32 bit = 1u << (v[kk] & 31);
                        // red_mode 1: look before setting.  Bits only ever get set, so a stale (L1) copy can at worst
                        // cost a redundant RED; most survivors of a power-law graph name a vertex that is already in
                        if (red_mode == 1 && (__ldca(w) & bit)) continue;
                        if (red_mode == 2) continue;  // measurement only: no bitmap traffic at all
                        const int so = staging ? stg_off[j * L + kk] : -1;
                        if (so >= 0) atomicOr(stg + so + (v[kk] >> 5), bit);
                        else atomicOr(w, bit);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        if (my_cnt) atomicAdd((unsigned long long *)&survivors[my_qpath], (unsigned long long)my_cnt);
        if (staging) {
            const u32 wps = (u32)words_per_slot;
            asm volatile("bar.sync 1, %0;" ::"n"(kTileRows) : "memory");
            const u32 used = *stg_used;
            for (u32 i = threadIdx.x; i < used; i += kTileRows) {
                const u32 v = stg[i];
                if (v) {
                    const u32 seg = i / wps;
                    atomicOr(bitmap + (u64)stg_slot[seg] * words_per_slot + (i - seg * wps), v);
                }
            }
        }
    }
}

// ---- the scan of an ids-only table ---------------------------------------------------------------------------
// Layout for tables that do not fit materialised (BASELINE.json config 4: l=3, e=4 -> 160-byte rows, ~2 x 10^10 of them):
// a row is its L vertex ids (4L bytes); label, degree, class position and embedding of every vertex come from the packed
// per-vertex records k1_expand gathers from (16 + 8e bytes each, L2-resident up to a few million vertices).  The
// row-level test is the same; what streams from HBM is 4L bytes per row plus the gathers that miss L2, and neighbouring
// rows of a bucket share all but their last vertex, so most gathers of a warp are one broadcast line.
// One row per thread, a contiguous chunk of the work list per CTA, the query-path block of the chunk in shared memory.
template <int L, int E>
__global__ void __launch_bounds__(kTileRows) k2_scan_ids_kernel(TableView t, const uint4 *__restrict__ vrec,
                                                                const QBlockRec<L, E> *__restrict__ qblocks,
                                                                const u64 *__restrict__ worklist,
                                                                const u64 *__restrict__ counters, u32 *__restrict__ bitmap,
                                                                u64 words_per_slot, u64 *__restrict__ survivors) {
    constexpr int D = L * E;
    constexpr int RQ = 1 + (E + 1) / 2;
    __shared__ QBlockRec<L, E> s_rec;
    const int lane = threadIdx.x & 31;
    const u32 r = threadIdx.x;
    const u64 n_items = counters[0];
    const u64 per = (n_items + gridDim.x - 1) / gridDim.x;
    const u64 first = min((u64)blockIdx.x * per, n_items), n_my = min(per, n_items - first);
    u32 cur_b = 0xffffffffu, my_cnt = 0, my_qpath = 0;
    for (u64 k = 0; k < n_my; k++) {
        const u64 item = worklist[first + k];
        const u32 tile = (u32)item, b = (u32)(item >> 32);
        if (b != cur_b) {  // block-uniform
            if (my_cnt) atomicAdd((unsigned long long *)&survivors[my_qpath], (unsigned long long)my_cnt);
            my_cnt = 0;
            cur_b = b;
            __syncthreads();  // everybody is done with the previous block's record
            const u32 *src = reinterpret_cast<const u32 *>(&qblocks[b]);
            u32 *dst = reinterpret_cast<u32 *>(&s_rec);
            for (u32 i = threadIdx.x; i < sizeof(QBlockRec<L, E>) / 4; i += blockDim.x) dst[i] = src[i];
            __syncthreads();
            my_qpath = lane < kQB ? s_rec.qpath[lane] : 0;
        }
        const bool valid = (u64)tile * kTileRows + r < t.n_rows;
        u32 v[L];
#pragma unroll
        for (int kk = 0; kk < L; kk++) v[kk] = valid ? __ldcs(t.vids + ((u64)tile * L + kk) * kTileRows + r) : 0u;
        uint4 head[L];
        double pde[D];
#pragma unroll
        for (int kk = 0; kk < L; kk++) {
            const uint4 *rec = vrec + (u64)v[kk] * RQ;
            head[kk] = __ldg(rec);
#pragma unroll
            for (int x = 0; x < E; x += 2) {
                const uint4 q = __ldg(rec + 1 + x / 2);
                pde[kk * E + x] = __hiloint2double((int)q.y, (int)q.x);
                if (x + 1 < E) pde[kk * E + x + 1] = __hiloint2double((int)q.w, (int)q.z);
            }
        }
        const u32 nq = s_rec.n;
        u32 okm = 0;
        for (u32 j = 0; j < nq; j++) {
            bool ok = valid;
#pragma unroll
            for (int kk = 0; kk < L; kk++) ok &= (s_rec.labels[j][kk] == head[kk].x) & (s_rec.degs[j][kk] <= head[kk].y);
            if (ok) {
#pragma unroll
                for (int d = 0; d < D; d++) ok &= !(s_rec.pde[j][d] - pde[d] > kEps);
            }
            const unsigned m = __ballot_sync(kFull, ok);
            if (lane == (int)j) my_cnt += __popc(m);
            okm |= (u32)ok << j;
        }
        if (__any_sync(kFull, okm != 0)) {
            const u32 okm_left = __shfl_up_sync(kFull, okm, 1);
#pragma unroll
            for (int kk = 0; kk < L; kk++) {
                const u32 pos = okm ? head[kk].z : 0xffffffffu, lab = head[kk].x;  // class position: the bit index
                const u32 pos_left = __shfl_up_sync(kFull, pos, 1), lab_left = __shfl_up_sync(kFull, lab, 1);
                u32 todo = okm;
                if (lane > 0 && pos_left == pos && lab_left == lab) todo &= ~okm_left;  // the left neighbour sets the same bits
                while (todo) {
                    const int j = __ffs(todo) - 1;
                    todo &= todo - 1;
                    atomicOr(bitmap + (u64)s_rec.slot[j][kk] * words_per_slot + (pos >> 5), 1u << (pos & 31));
                }
            }
        }
    }
    if (my_cnt) atomicAdd((unsigned long long *)&survivors[my_qpath], (unsigned long long)my_cnt);
}

template <int L, int E>
cudaError_t launch_select(const TableView &t, const void *qblocks, const u32 *qb_t0, const u64 *qb_prefix,
                          u32 n_qblocks, u64 n_items, bool prune, u64 *worklist, u64 *counters, cudaStream_t s) {
    if (n_items == 0) return cudaSuccess;
    unsigned blocks = (unsigned)std::min<u64>((n_items + 255) / 256, 148 * 16);
    k2_select_kernel<L, E><<<blocks, 256, 0, s>>>(t, reinterpret_cast<const QBlockRec<L, E> *>(qblocks), qb_t0,
                                                  qb_prefix, n_qblocks, n_items, prune, worklist, counters);
    return cudaGetLastError();
}

inline int scan_red_mode() {
    static int mode = -1;
    if (mode < 0) { const char *e = getenv("GPE_SCAN_RED"); mode = e ? atoi(e) : 0; }
    return mode;
}

inline bool scan_stage_on() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("GPE_SCAN_STAGE"); on = e ? atoi(e) : 1; }
    return on != 0;
}

inline int scan_env(const char *name, int dflt, int lo, int hi) {
    const char *e = getenv(name);
    const int v = e ? atoi(e) : dflt;
    return v < lo || v > hi ? dflt : v;
}

template <int L, int E, bool VIDS, int NSREQ>
cudaError_t launch_scan_ns(const TableView &t, const void *qblocks, const u64 *worklist, const u64 *counters,
                           u32 *bitmap, u64 words_per_slot, u64 *survivors, int sm_count, int want_ctas, cudaStream_t s) {
    using G = ScanGeom<L, E, VIDS, NSREQ>;
    // Staging capacity: what `want_ctas` CTAs per SM leave next to the ring, in whole slots (at most one per (path, position)
    // of a block); when not even one slot fits (huge label classes) the kernel falls back to direct REDs.
    // (cached per device: cudaFuncSetAttribute and the occupancy are per-device properties, and one process may hold
    //  contexts on several GPUs)
    struct Cfg { int ctas_per_sm = 0; u64 cfg_words = ~0ull; u32 stage_words = 0; size_t smem_bytes = 0; };
    static Cfg cfgs[kMaxDevices];
    int dev = 0;
    cudaGetDevice(&dev);
    Cfg &cf = cfgs[dev % kMaxDevices];
    int &ctas_per_sm = cf.ctas_per_sm;
    u64 &cfg_words = cf.cfg_words;
    u32 &stage_words = cf.stage_words;
    size_t &smem_bytes = cf.smem_bytes;
    if (ctas_per_sm == 0 || cfg_words != words_per_slot) {
        stage_words = 0;
        if (VIDS && words_per_slot > 0 && scan_stage_on()) {
            const size_t budget = (size_t)(227 * 1024) / want_ctas - 1024;  // per CTA (1 KB reserved each)
            if (budget > (size_t)G::kSmemBytes) {
                const u64 slots = std::min<u64>((budget - G::kSmemBytes) / 4 / words_per_slot, (u64)kQB * L);
                stage_words = (u32)(slots * words_per_slot);
            }
        }
        smem_bytes = (size_t)G::kSmemBytes + (size_t)stage_words * 4;
        cudaError_t e = cudaFuncSetAttribute(k2_scan_kernel<L, E, VIDS, NSREQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem_bytes);
        if (e != cudaSuccess) return e;
        int n = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k2_scan_kernel<L, E, VIDS, NSREQ>, kTileRows + 32, smem_bytes);
        if (e != cudaSuccess) return e;
        ctas_per_sm = n < 1 ? 1 : n;
        cfg_words = words_per_slot;
    }
    unsigned blocks = (unsigned)(sm_count * ctas_per_sm);
    k2_scan_kernel<L, E, VIDS, NSREQ><<<blocks, kTileRows + 32, smem_bytes, s>>>(
        t, reinterpret_cast<const QBlockRec<L, E> *>(qblocks), worklist, counters, bitmap, words_per_slot, survivors,
        scan_red_mode(), stage_words);
    return cudaGetLastError();
}

template <int L, int E, bool VIDS>
cudaError_t launch_scan_v(const TableView &t, const void *qblocks, const u64 *worklist, const u64 *counters,
                          u32 *bitmap, u64 words_per_slot, u64 *survivors, int sm_count, cudaStream_t s) {
    // Bucketed work lists (VIDS): a 2-deep ring and as many CTAs per SM as fit (3 at l=2, e=2) instead of a 4-deep ring and
    // 2 CTAs -- the consumers' FP64 compare chains were latency bound at 27 % occupancy: in-step scan 0.670 -> 0.558 ms on
    // config 2 = 0.61 -> 0.73 of the HBM peak (gpurun_out/r02y_ab*: 3 stages x 2 CTAs 0.674, 2 x 2 0.696).  The streaming mode
    // (every row against every plan path, 0.97 of the peak) keeps the deep ring.  GPE_SCAN_STAGES / GPE_SCAN_CTAS override.
    if constexpr (VIDS) {
        using G2 = ScanGeom<L, E, VIDS, 2>;
        static const int ns = scan_env("GPE_SCAN_STAGES", 2, 2, 4);
        static const int fit = (int)std::max<size_t>(1, std::min<size_t>(3, (size_t)(227 * 1024) / ((size_t)G2::kSmemBytes + 4096)));
        static const int ctas = scan_env("GPE_SCAN_CTAS", ns == 2 ? fit : 2, 1, 4);
        if (ns == 2) return launch_scan_ns<L, E, VIDS, 2>(t, qblocks, worklist, counters, bitmap, words_per_slot, survivors, sm_count, ctas, s);
        return launch_scan_ns<L, E, VIDS, kStages>(t, qblocks, worklist, counters, bitmap, words_per_slot, survivors, sm_count, ctas, s);
    }
    return launch_scan_ns<L, E, VIDS, kStages>(t, qblocks, worklist, counters, bitmap, words_per_slot, survivors, sm_count, 2, s);
}

template <int L, int E>
cudaError_t launch_scan(const TableView &t, const void *qblocks, const u64 *worklist, const u64 *counters,
                        u32 *bitmap, u64 words_per_slot, u64 *survivors, bool with_vids, int sm_count, cudaStream_t s) {
    return with_vids ? launch_scan_v<L, E, true>(t, qblocks, worklist, counters, bitmap, words_per_slot, survivors, sm_count, s)
                     : launch_scan_v<L, E, false>(t, qblocks, worklist, counters, bitmap, words_per_slot, survivors, sm_count, s);
}

template <int L, int E>
void pack_rec(void *dst, u32 n, u32 first_qpath, const u32 *qpath_ids, const u32 *labels, const u32 *degs,
              const u32 *slots, const double *pde) {
    QBlockRec<L, E> rec;
    memset(&rec, 0, sizeof rec);
    rec.n = n;
    rec.first_qpath = first_qpath;
    for (u32 j = 0; j < n; j++) {
        for (int k = 0; k < L; k++) {
            rec.labels[j][k] = labels[j * L + k];
            rec.degs[j][k] = degs[j * L + k];
            rec.slot[j][k] = slots[j * L + k];
        }
        rec.qpath[j] = qpath_ids[j];
        for (int d = 0; d < L * E; d++) rec.pde[j][d] = pde[(size_t)j * L * E + d];
    }
    memcpy(dst, &rec, sizeof rec);
}

}  // namespace

#define GPE_DISPATCH_LE(L_, E_, CALL)                                    \
    do {                                                                 \
        if ((L_) == 3 && (E_) == 1) { CALL(3, 1); }                      \
        else if ((L_) == 3 && (E_) == 2) { CALL(3, 2); }                 \
        else if ((L_) == 3 && (E_) == 3) { CALL(3, 3); }                 \
        else if ((L_) == 3 && (E_) == 4) { CALL(3, 4); }                 \
        else if ((L_) == 3 && (E_) == 8) { CALL(3, 8); }                 \
        else if ((L_) == 4 && (E_) == 1) { CALL(4, 1); }                 \
        else if ((L_) == 4 && (E_) == 2) { CALL(4, 2); }                 \
        else if ((L_) == 4 && (E_) == 3) { CALL(4, 3); }                 \
        else if ((L_) == 4 && (E_) == 4) { CALL(4, 4); }                 \
        else if ((L_) == 4 && (E_) == 8) { CALL(4, 8); }                 \
    } while (0)

bool k2_supported(u32 L, u32 E) {
    return (L == 3 || L == 4) && (E == 1 || E == 2 || E == 3 || E == 4 || E == 8);
}

size_t qblock_rec_bytes(u32 L, u32 E) {
    size_t r = 0;
#define CALL(l, e) r = sizeof(QBlockRec<l, e>)
    GPE_DISPATCH_LE(L, E, CALL);
#undef CALL
    return r;
}

void qblock_pack(u32 L, u32 E, void *dst, u32 n, u32 first_qpath, const u32 *qpath_ids, const u32 *labels,
                 const u32 *degs, const u32 *slots, const double *pde) {
#define CALL(l, e) pack_rec<l, e>(dst, n, first_qpath, qpath_ids, labels, degs, slots, pde)
    GPE_DISPATCH_LE(L, E, CALL);
#undef CALL
}

cudaError_t k2_select(const TableView &t, const void *qblocks, const u32 *qb_t0, const u64 *qb_prefix, u32 n_qblocks,
                      u64 n_items, bool prune, u64 *worklist, u64 *counters, cudaStream_t s) {
    cudaError_t e = cudaErrorInvalidValue;
#define CALL(l, e_) e = launch_select<l, e_>(t, qblocks, qb_t0, qb_prefix, n_qblocks, n_items, prune, worklist, counters, s)
    GPE_DISPATCH_LE(t.L, t.E, CALL);
#undef CALL
    return e;
}

template <int L, int E>
cudaError_t launch_scan_ids(const TableView &t, const void *vrec, const void *qblocks, const u64 *worklist, const u64 *counters,
                            u32 *bitmap, u64 words_per_slot, u64 *survivors, int sm_count, cudaStream_t s) {
    k2_scan_ids_kernel<L, E><<<(unsigned)sm_count * 8, kTileRows, 0, s>>>(t, reinterpret_cast<const uint4 *>(vrec),
                                                                        reinterpret_cast<const QBlockRec<L, E> *>(qblocks), worklist,
                                                                        counters, bitmap, words_per_slot, survivors);
    return cudaGetLastError();
}

cudaError_t k2_scan_ids(const TableView &t, const void *vrec, const void *qblocks, const u64 *worklist, const u64 *counters,
                        u32 *bitmap, u64 words_per_slot, u64 *survivors, int sm_count, cudaStream_t s) {
    cudaError_t e = cudaErrorInvalidValue;
#define CALL(l, e_) e = launch_scan_ids<l, e_>(t, vrec, qblocks, worklist, counters, bitmap, words_per_slot, survivors, sm_count, s)
    GPE_DISPATCH_LE(t.L, t.E, CALL);
#undef CALL
    return e;
}

cudaError_t k2_scan(const TableView &t, const void *qblocks, const u64 *worklist, const u64 *counters, u32 *bitmap,
                    u64 words_per_slot, u64 *survivors, bool with_vids, int sm_count, cudaStream_t s) {
    cudaError_t e = cudaErrorInvalidValue;
#define CALL(l, e_) e = launch_scan<l, e_>(t, qblocks, worklist, counters, bitmap, words_per_slot, survivors, with_vids, sm_count, s)
    GPE_DISPATCH_LE(t.L, t.E, CALL);
#undef CALL
    return e;
}

}  // namespace gpe
