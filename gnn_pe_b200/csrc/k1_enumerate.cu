// k1_enumerate.cu -- offline path enumeration and path-table construction (hot path 2).
//
// Replaces the reference's serial `dfs` + hash-set dedup (custom.h:66-92, driven by main.cpp:92-96)
// and the per-path gather of gen_pde (custom.h:546-572).  The hash set is not needed: on a simple
// graph the reference's output equals, in content and order,
//     for a in membership order, b in N(a) ascending, c in N(b) ascending [, d in N(c) ascending]:
//         emit iff the vertices are distinct and rank[a] < rank[last]
// so enumeration is count -> exclusive scan -> write, one warp per start vertex, lanes over the
// innermost adjacency list with ballot/popc compaction.
//
// Bound by HBM write bandwidth (fill) and L2 gathers of rank[]/label[] (count): no tensor cores.
#include "gpe_internal.h"

namespace gpe {

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ------------------------------------------------------------------------------------------------
// device-wide exclusive scan (u64), three-phase: block reduce, recursive scan of block sums, block scan
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanBlock = kScanThreads * kScanItems;

__device__ __forceinline__ u64 block_exclusive_scan(u64 v, u64 *total, u64 *s_warp /*>=8*/) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u64 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u64 t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        u64 ws = lane < (kScanThreads / 32) ? s_warp[lane] : 0;
        u64 winc = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u64 t = __shfl_up_sync(kFull, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < (kScanThreads / 32)) s_warp[lane] = winc - ws;
        if (lane == (kScanThreads / 32) - 1) s_warp[kScanThreads / 32] = winc;
    }
    __syncthreads();
    if (total) *total = s_warp[kScanThreads / 32];
    return inc - v + s_warp[w];
}

__global__ void scan_reduce_kernel(const u64 *data, u64 n, u64 *block_sums) {
    __shared__ u64 s_warp[kScanThreads / 32 + 1];
    u64 base = (u64)blockIdx.x * kScanBlock;
    u64 sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        u64 idx = base + (u64)i * kScanThreads + threadIdx.x;
        if (idx < n) sum += data[idx];
    }
    u64 total;
    block_exclusive_scan(sum, &total, s_warp);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void scan_apply_kernel(u64 *data, u64 n, const u64 *block_offsets) {
    __shared__ u64 s_warp[kScanThreads / 32 + 1];
    u64 base = (u64)blockIdx.x * kScanBlock + (u64)threadIdx.x * kScanItems;
    u64 v[kScanItems];
    u64 sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        v[i] = (base + i < n) ? data[base + i] : 0;
        sum += v[i];
    }
    u64 excl = block_exclusive_scan(sum, nullptr, s_warp) + (block_offsets ? block_offsets[blockIdx.x] : 0);
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        if (base + i < n) data[base + i] = excl;
        excl += v[i];
    }
}

// ------------------------------------------------------------------------------------------------
// The walk: one warp, one start vertex `a` (rank ra).  F is a functor with warp-uniform hooks.
// ------------------------------------------------------------------------------------------------
// GROUPED: the innermost adjacency list is read in label-grouped order (nbrG) instead of id order.  The set of rows
// is the same; consecutive lanes then mostly share a label, hence a table bucket, and the fill writes runs of
// consecutive rows (full sectors) instead of one row per bucket.  Only for the order-insensitive passes (histogram,
// fill): count and dump reproduce the reference's order and walk by id.
template <int L, bool GROUPED, class F>
__device__ __forceinline__ void walk_start_vertex(const GraphView &g, u32 a, u32 ra, int lane, F &f) {
    const u32 a0 = g.off[a], a1 = g.off[a + 1];
    for (u32 s = a0; s < a1; ++s) {
        const u32 b = g.nbr[s];
        f.begin_slot(s - a0, b);
        const u32 b0 = g.off[b], b1 = g.off[b + 1];
        if (L == 3) {
            for (u32 j = b0; j < b1; j += 32) {
                u32 jj = j + lane;
                bool in = jj < b1;
                u32 c = in ? (GROUPED ? g.nbrG[jj] : g.nbr[jj]) : 0u;
                const u64 rl = in ? g.ranklab[c] : 0ull;  // rank | label << 32: one gather for both
                bool valid = in && (u32)rl > ra;  // c != a follows from the strict rank test
                f.chunk(b, c, 0u, valid, (u32)(rl >> 32));
            }
        } else {
            for (u32 t = b0; t < b1; ++t) {
                const u32 c = g.nbr[t];
                if (c == a) continue;
                const u32 c0 = g.off[c], c1 = g.off[c + 1];
                for (u32 j = c0; j < c1; j += 32) {
                    u32 jj = j + lane;
                    bool in = jj < c1;
                    u32 d = in ? (GROUPED ? g.nbrG[jj] : g.nbr[jj]) : 0u;
                    const u64 rl = in ? g.ranklab[d] : 0ull;
                    bool valid = in && d != b && (u32)rl > ra;  // d != a by rank, d != c: no loops
                    f.chunk(b, c, d, valid, (u32)(rl >> 32));
                }
            }
        }
        f.end_slot(s - a0);
    }
}

// ---- count --------------------------------------------------------------------------------------
struct CountF {
    u64 *out;  // this start vertex's slots
    u64 cnt;
    int lane;
    __device__ void begin_slot(u32, u32) { cnt = 0; }
    __device__ void chunk(u32, u32, u32, bool valid, u32) { cnt += __popc(__ballot_sync(kFull, valid)); }
    __device__ void end_slot(u32 si) { if (lane == 0) out[si] = cnt; }
};

template <int L>
__global__ void __launch_bounds__(256) k1_count_kernel(GraphView g, const u32 *__restrict__ sorted,
                                                       const u32 *__restrict__ offr, u64 *__restrict__ cnt_r) {
    const int lane = threadIdx.x & 31;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < g.V; i += nwarps) {
        u32 a = sorted[i];
        CountF f{cnt_r + offr[i], 0, lane};
        walk_start_vertex<L, false>(g, a, i, lane, f);
    }
}

// ---- rank-ordered CSR offsets, per-partition row counts ---------------------------------------------
__global__ void k1_rows_kernel(u32 V, const u32 *__restrict__ sorted, const u32 *__restrict__ offr,
                               const u64 *__restrict__ ebase, const u32 *__restrict__ member,
                               u64 *part_rows, u64 *start_rows) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > V) return;
    u64 first = ebase[offr[i]];
    start_rows[i] = first;
    if (i < V) {
        u64 rows = ebase[offr[i + 1]] - first;
        if (rows) atomicAdd((unsigned long long *)&part_rows[member[sorted[i]]], (unsigned long long)rows);
    }
}

// ---- dump in the reference's order -----------------------------------------------------------------
template <int L>
struct DumpF {
    const u64 *ebase;  // this start vertex's slots
    u32 a;
    u64 first, n;
    u32 *out;
    u64 id;
    int lane;
    __device__ void begin_slot(u32 si, u32) { id = ebase[si]; }
    __device__ void chunk(u32 b, u32 c, u32 d, bool valid, u32) {
        unsigned m = __ballot_sync(kFull, valid);
        u64 my = id + __popc(m & lanemask_lt());
        if (valid && my >= first && my < first + n) {
            u32 *row = out + (my - first) * L;
            row[0] = a;
            row[1] = b;
            row[2] = c;
            if (L == 4) row[3] = d;
        }
        id += __popc(m);
    }
    __device__ void end_slot(u32) {}
};

template <int L>
__global__ void __launch_bounds__(256) k1_dump_kernel(GraphView g, const u32 *__restrict__ sorted,
                                                      const u32 *__restrict__ offr, const u64 *__restrict__ ebase,
                                                      u32 rank_lo, u32 rank_hi, u64 first, u64 n, u32 *out) {
    const int lane = threadIdx.x & 31;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 i = rank_lo + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); i <= rank_hi && i < g.V; i += nwarps) {
        DumpF<L> f{ebase + offr[i], sorted[i], first, n, out, 0, lane};
        walk_start_vertex<L, false>(g, sorted[i], i, lane, f);
    }
}

// ---- label-sequence key of a row ----------------------------------------------------------------------
struct KeyParams {
    u32 radix[kMaxL], stride[kMaxL];
};

__device__ __forceinline__ u32 key_term(const KeyParams &kp, int k, u32 label) {
    return kp.radix[k] <= 1 ? 0u : (label % kp.radix[k]) * kp.stride[k];
}

// ---- histogram of rows per bucket ----------------------------------------------------------------------
template <int L>
struct HistF {
    const GraphView &g;
    const KeyParams &kp;
    u64 *hist;
    u32 key_a;
    u32 key_ab;
    __device__ void begin_slot(u32, u32 b) { key_ab = key_a + key_term(kp, 1, g.label[b]); }
    __device__ void chunk(u32, u32 c, u32 d, bool valid, u32 last_label) {  // last_label: label of the path's last vertex
        unsigned act = __ballot_sync(kFull, valid);
        if (!valid) return;
        u32 key = key_ab + (L == 4 ? key_term(kp, 2, g.label[c]) + key_term(kp, 3, last_label) : key_term(kp, 2, last_label));
        unsigned peers = __match_any_sync(act, key);
        if ((peers & lanemask_lt()) == 0)
            atomicAdd((unsigned long long *)&hist[key], (unsigned long long)__popc(peers));
    }
    __device__ void end_slot(u32) {}
};

template <int L>
__global__ void __launch_bounds__(256) k1_hist_kernel(GraphView g, KeyParams kp, const u32 *__restrict__ sorted,
                                                      const u32 *__restrict__ member,
                                                      const unsigned char *__restrict__ part_sel, u64 *hist) {
    const int lane = threadIdx.x & 31;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < g.V; i += nwarps) {
        u32 a = sorted[i];
        if (part_sel && !part_sel[member[a]]) continue;
        HistF<L> f{g, kp, hist, key_term(kp, 0, g.label[a]), 0};
        walk_start_vertex<L, true>(g, a, i, lane, f);
    }
}

// ---- fill: every row (its vertex ids) goes to the next free position of its bucket --------------------
template <int L>
struct FillF {
    const GraphView &g;
    const KeyParams &kp;
    const TableView &t;
    u64 *cursor;
    u32 a, key_a, key_ab;
    __device__ void begin_slot(u32, u32 b) { key_ab = key_a + key_term(kp, 1, g.label[b]); }
    // Only the vertex ids are scattered (12 bytes per row at l=2): the 72-byte scan row is materialised afterwards by
    // k1_expand, tile by tile with full-line stores.  Scattering whole rows cost 21 partial-sector stores per row
    // (r01q capture: 3.7 G write sectors and 29 GB of DRAM read-for-fill for 43 GB of payload, 88 ms).
    __device__ void put(u64 row, int k, u32 v) const {
        t.vids[((row / kTileRows) * L + k) * kTileRows + (u32)(row % kTileRows)] = v;
    }
    __device__ void chunk(u32 b, u32 c, u32 d, bool valid, u32 last_label) {
        unsigned act = __ballot_sync(kFull, valid);
        if (!valid) return;
        u32 key = key_ab + (L == 4 ? key_term(kp, 2, g.label[c]) + key_term(kp, 3, last_label) : key_term(kp, 2, last_label));
        unsigned peers = __match_any_sync(act, key);
        int leader = __ffs(peers) - 1;
        u64 base = 0;
        if ((peers & lanemask_lt()) == 0)
            base = atomicAdd((unsigned long long *)&cursor[key], (unsigned long long)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        u64 row = base + __popc(peers & lanemask_lt());
        put(row, 0, a);
        put(row, 1, b);
        put(row, 2, c);
        if (L == 4) put(row, 3, d);
    }
    __device__ void end_slot(u32) {}
};

template <int L>
__global__ void __launch_bounds__(256) k1_fill_kernel(GraphView g, KeyParams kp, TableView t,
                                                      const u32 *__restrict__ sorted, const u32 *__restrict__ member,
                                                      const unsigned char *__restrict__ part_sel, u64 *cursor) {
    const int lane = threadIdx.x & 31;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < g.V; i += nwarps) {
        u32 a = sorted[i];
        if (part_sel && !part_sel[member[a]]) continue;
        FillF<L> f{g, kp, t, cursor, a, key_term(kp, 0, g.label[a]), 0};
        walk_start_vertex<L, true>(g, a, i, lane, f);
    }
}

// ---- expand: vertex ids -> scan tiles, and the per-tile summaries (label range, max degree, max-corner of the path
//      embeddings: the GPU analogue of build_auxiliary_index, custom.h:268-364) in the same pass ----------
// One CTA per tile, one thread per row: reads the row's vertex ids (coalesced), gathers label / degree / embedding /
// class position of every vertex (28 bytes per vertex, the whole per-vertex table is L2-resident), writes the tile's
// structure-of-arrays columns with full-line stores, replaces the ids by class positions (the bit index of the
// candidate bitmaps), and reduces the tile's label range, max degrees and max-corner.  HBM-write bound:
// (8L + 8Le + 4L) bytes per row.
// The per-vertex gathers come from ONE packed record per vertex (label, degree, class position, pad | embedding: 16 + 8e
// bytes, built by k1_vertex_records): one or two 32-byte sectors per vertex instead of one each from four arrays -- the
// kernel was bound by L2 sector requests (6 G four-byte gathers for 515 M rows), not by its stores.
template <int L, int E>
__global__ void __launch_bounds__(kTileRows) k1_expand_kernel(TableView t, const uint4 *__restrict__ vrec) {
    constexpr int W = kTileRows / 32;
    constexpr int RQ = 1 + (E + 1) / 2;         // 16-byte words per record
    __shared__ u32 s_u[2][3 * L][W];            // per-warp partials: label min, label max, max degree per position
    __shared__ double s_d[2][L * E][W];         // per-warp partials of the max-corner; double-buffered by tile parity,
    const u32 r = threadIdx.x;                  // so one barrier per tile is enough
    const int lane = r & 31, warp = r >> 5;
    int buf = 0;
    for (u64 tile = blockIdx.x; tile < t.n_tiles; tile += gridDim.x, buf ^= 1) {
        const bool valid = tile * kTileRows + r < t.n_rows;
        unsigned char *base = t.tiles + tile * t.tile_bytes;
        u32 *lab = reinterpret_cast<u32 *>(base);
        u32 *dg = reinterpret_cast<u32 *>(base + 4u * L * kTileRows);
        double *pde = reinterpret_cast<double *>(base + 8u * L * kTileRows);
        u32 *vid = t.vids + tile * L * kTileRows;
        // all of a row's loads are issued before anything depends on them: ids first, then the records
        u32 v[L];
#pragma unroll
        for (int k = 0; k < L; k++) v[k] = valid ? __ldcs(vid + k * kTileRows + r) : 0u;  // streaming: the table is touched once,
        uint4 head[L];                                                                  // the vertex records stay in L2
        double emb[L][E];
#pragma unroll
        for (int k = 0; k < L; k++) {
            const uint4 *rec = vrec + (u64)v[k] * RQ;
            head[k] = valid ? __ldg(rec) : make_uint4(0xffffffffu, 0u, 0u, 0u);
#pragma unroll
            for (int x = 0; x < E; x += 2) {
                const uint4 q = valid ? __ldg(rec + 1 + x / 2) : make_uint4(0u, 0u, 0xbff00000u, 0u);
                emb[k][x] = valid ? __hiloint2double((int)q.y, (int)q.x) : -1.0;
                if (x + 1 < E) emb[k][x + 1] = valid ? __hiloint2double((int)q.w, (int)q.z) : -1.0;
            }
        }
#pragma unroll
        for (int k = 0; k < L; k++) {
#pragma unroll
            for (int x = 0; x < E; x++) {
                double e = emb[k][x];
                if (valid) __stcs(pde + (k * E + x) * kTileRows + r, e);
#pragma unroll
                for (int o = 16; o; o >>= 1) e = fmax(e, __shfl_xor_sync(kFull, e, o));
                if (lane == 0) s_d[buf][k * E + x][warp] = e;
            }
        }
#pragma unroll
        for (int k = 0; k < L; k++) {
            const u32 lk = head[k].x, dk = head[k].y;
            if (valid) {
                __stcs(lab + k * kTileRows + r, lk);
                __stcs(dg + k * kTileRows + r, dk);
                __stcs(vid + k * kTileRows + r, head[k].z);
            }
            u32 mn = lk, mx = valid ? lk : 0u, dm = dk;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                mn = min(mn, __shfl_xor_sync(kFull, mn, o));
                mx = max(mx, __shfl_xor_sync(kFull, mx, o));
                dm = max(dm, __shfl_xor_sync(kFull, dm, o));
            }
            if (lane == 0) { s_u[buf][3 * k][warp] = mn; s_u[buf][3 * k + 1][warp] = mx; s_u[buf][3 * k + 2][warp] = dm; }
        }
        __syncthreads();
        if (r < 3 * L) {
            const int k = r / 3, what = r % 3;
            u32 acc = s_u[buf][r][0];
            for (int w = 1; w < W; w++) acc = what == 0 ? min(acc, s_u[buf][r][w]) : max(acc, s_u[buf][r][w]);
            u32 *dst = what == 0 ? t.lab_min : what == 1 ? t.lab_max : t.deg_max;
            dst[k * t.n_tiles + tile] = acc;
        } else if (r >= 32 && r - 32 < t.D) {
            const u32 dd = r - 32;
            double acc = s_d[buf][dd][0];
            for (int w = 1; w < W; w++) acc = fmax(acc, s_d[buf][dd][w]);
            t.pde_max[dd * t.n_tiles + tile] = acc;
        }
    }
}

// label | degree | class position | 0, then the embedding padded to an even number of doubles
__global__ void __launch_bounds__(256) k1_vertex_records_kernel(GraphView g, u32 rq, uint4 *vrec) {
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g.V) return;
    uint4 *rec = vrec + (u64)v * rq;
    rec[0] = make_uint4(g.label[v], g.deg[v], g.lpos[v], 0u);
    for (u32 x = 0; x < g.e; x += 2) {
        const double a = g.vde[(u64)v * g.e + x], b = x + 1 < g.e ? g.vde[(u64)v * g.e + x + 1] : 0.0;
        rec[1 + x / 2] = make_uint4((u32)__double2loint(a), (u32)__double2hiint(a), (u32)__double2loint(b), (u32)__double2hiint(b));
    }
}

__global__ void k1_dump_table_kernel(TableView t, GraphView g, const uint4 *__restrict__ vrec, u64 first, u64 n, u32 *vids,
                                     u32 *labels, u32 *degs, double *pde) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 row = first + i;
    u64 tile = row / kTileRows;
    u32 r = (u32)(row % kTileRows);
    if (t.ids_only) {  // the row is its vertex ids; the columns are what the scan gathers
        const u32 rq = 1 + (t.E + 1) / 2;
        for (u32 k = 0; k < t.L; k++) {
            const u32 v = t.vids[(tile * t.L + k) * kTileRows + r];
            const uint4 h = vrec[(u64)v * rq];
            if (vids) vids[i * t.L + k] = v;
            if (labels) labels[i * t.L + k] = h.x;
            if (degs) degs[i * t.L + k] = h.y;
            if (pde)
                for (u32 x = 0; x < t.E; x++) pde[i * t.D + k * t.E + x] = g.vde[(u64)v * t.E + x];
        }
        return;
    }
    const unsigned char *base = t.tiles + tile * t.tile_bytes;
    const u32 *lab = reinterpret_cast<const u32 *>(base);
    const u32 *dg = reinterpret_cast<const u32 *>(base + 4u * t.L * kTileRows);
    const double *pd = reinterpret_cast<const double *>(base + 8u * t.L * kTileRows);
    for (u32 k = 0; k < t.L; k++) {
        if (vids) vids[i * t.L + k] = g.lclass[g.lcoff[lab[k * kTileRows + r]] + t.vids[(tile * t.L + k) * kTileRows + r]];
        if (labels) labels[i * t.L + k] = lab[k * kTileRows + r];
        if (degs) degs[i * t.L + k] = dg[k * kTileRows + r];
    }
    if (pde)
        for (u32 d = 0; d < t.D; d++) pde[i * t.D + d] = pd[d * kTileRows + r];
}

KeyParams key_params(const TableView &t) {
    KeyParams kp;
    for (int k = 0; k < kMaxL; k++) { kp.radix[k] = t.key_radix[k]; kp.stride[k] = t.key_stride[k]; }
    return kp;
}

int walk_grid(int sm_count) { return sm_count * 8; }  // 8 CTAs of 8 warps per SM

}  // namespace

u64 exclusive_scan_launches(u64 n) {
    if (n == 0) return 0;
    u64 levels = 0;
    while (n > (u64)kScanBlock) { n = (n + kScanBlock - 1) / kScanBlock; levels++; }
    return 1 + 2 * levels;
}

cudaError_t exclusive_scan_u64(u64 *d_data, u64 n, DevBuf &tmp, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    // sizes of every level
    std::vector<u64> level_n;
    u64 cur = n;
    while (cur > (u64)kScanBlock) {
        cur = (cur + kScanBlock - 1) / kScanBlock;
        level_n.push_back(cur);
    }
    u64 total = 0;
    for (u64 x : level_n) total += x;
    cudaError_t e = tmp.reserve((total + 1) * sizeof(u64));
    if (e != cudaSuccess) return e;
    std::vector<u64 *> level_ptr;
    u64 *ptr = tmp.as<u64>();
    for (u64 x : level_n) { level_ptr.push_back(ptr); ptr += x; }
    // up-sweep
    u64 *src = d_data;
    u64 src_n = n;
    for (size_t lv = 0; lv < level_n.size(); lv++) {
        scan_reduce_kernel<<<(unsigned)level_n[lv], kScanThreads, 0, s>>>(src, src_n, level_ptr[lv]);
        src = level_ptr[lv];
        src_n = level_n[lv];
    }
    // top level fits one block
    scan_apply_kernel<<<1, kScanThreads, 0, s>>>(src, src_n, nullptr);
    // down-sweep
    for (size_t lv = level_n.size(); lv-- > 0;) {
        u64 *dst = lv == 0 ? d_data : level_ptr[lv - 1];
        u64 dst_n = lv == 0 ? n : level_n[lv - 1];
        scan_apply_kernel<<<(unsigned)level_n[lv], kScanThreads, 0, s>>>(dst, dst_n, level_ptr[lv]);
    }
    return cudaGetLastError();
}

namespace {
__global__ void __launch_bounds__(256) k1_rank_labels_kernel(u32 V, const u32 *__restrict__ rank, const u32 *__restrict__ label, u64 *out) {
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < V) out[v] = (u64)rank[v] | ((u64)label[v] << 32);
}
}  // namespace

cudaError_t k1_rank_labels(u32 V, const u32 *rank, const u32 *label, u64 *ranklab, cudaStream_t s) {
    if (V) k1_rank_labels_kernel<<<(V + 255) / 256, 256, 0, s>>>(V, rank, label, ranklab);
    return cudaGetLastError();
}

cudaError_t k1_count(const GraphView &g, u32 L, const u32 *sorted, const u32 *offr, u64 *cnt_r, int sm_count,
                     cudaStream_t s) {
    if (L == 3) k1_count_kernel<3><<<walk_grid(sm_count), 256, 0, s>>>(g, sorted, offr, cnt_r);
    else k1_count_kernel<4><<<walk_grid(sm_count), 256, 0, s>>>(g, sorted, offr, cnt_r);
    return cudaGetLastError();
}

cudaError_t k1_rows_per_partition(u32 V, const u32 *sorted, const u32 *offr, const u64 *ebase, const u32 *member,
                                  u64 *part_rows, u64 *start_rows, cudaStream_t s) {
    k1_rows_kernel<<<(V + 1 + 255) / 256, 256, 0, s>>>(V, sorted, offr, ebase, member, part_rows, start_rows);
    return cudaGetLastError();
}

cudaError_t k1_dump(const GraphView &g, u32 L, const u32 *sorted, const u32 *offr, const u64 *ebase, u32 rank_lo,
                    u32 rank_hi, u64 first, u64 n, u32 *out, cudaStream_t s) {
    u32 nv = rank_hi - rank_lo + 1;
    unsigned blocks = (unsigned)std::min<u64>(((u64)nv * 32 + 255) / 256, 148 * 8);
    if (L == 3) k1_dump_kernel<3><<<blocks, 256, 0, s>>>(g, sorted, offr, ebase, rank_lo, rank_hi, first, n, out);
    else k1_dump_kernel<4><<<blocks, 256, 0, s>>>(g, sorted, offr, ebase, rank_lo, rank_hi, first, n, out);
    return cudaGetLastError();
}

cudaError_t k1_histogram(const GraphView &g, const TableView &t, const u32 *sorted, const u32 *member,
                         const unsigned char *part_sel, u64 *hist, int sm_count, cudaStream_t s) {
    KeyParams kp = key_params(t);
    if (t.L == 3) k1_hist_kernel<3><<<walk_grid(sm_count), 256, 0, s>>>(g, kp, sorted, member, part_sel, hist);
    else k1_hist_kernel<4><<<walk_grid(sm_count), 256, 0, s>>>(g, kp, sorted, member, part_sel, hist);
    return cudaGetLastError();
}

cudaError_t k1_fill(const GraphView &g, const TableView &t, const u32 *sorted, const u32 *member,
                    const unsigned char *part_sel, u64 *cursor, int sm_count, cudaStream_t s) {
    KeyParams kp = key_params(t);
    if (t.L == 3) k1_fill_kernel<3><<<walk_grid(sm_count), 256, 0, s>>>(g, kp, t, sorted, member, part_sel, cursor);
    else k1_fill_kernel<4><<<walk_grid(sm_count), 256, 0, s>>>(g, kp, t, sorted, member, part_sel, cursor);
    return cudaGetLastError();
}

size_t k1_vertex_record_bytes(u32 V, u32 e) { return (size_t)std::max<u32>(V, 1) * (1 + (e + 1) / 2) * sizeof(uint4); }

cudaError_t k1_vertex_records(const GraphView &g, void *vrec, cudaStream_t s) {
    if (g.V == 0) return cudaSuccess;
    k1_vertex_records_kernel<<<(g.V + 255) / 256, 256, 0, s>>>(g, 1 + (g.e + 1) / 2, reinterpret_cast<uint4 *>(vrec));
    return cudaGetLastError();
}

cudaError_t k1_expand(const TableView &t, const void *vrec, int sm_count, cudaStream_t s) {
    if (t.n_tiles == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)std::min<u64>(t.n_tiles, (u64)sm_count * 8 * 16);
    const uint4 *vr = reinterpret_cast<const uint4 *>(vrec);
#define EXPAND(L_, E_) k1_expand_kernel<L_, E_><<<blocks, kTileRows, 0, s>>>(t, vr)
    if (t.L == 3) {
        switch (t.E) { case 1: EXPAND(3, 1); break; case 2: EXPAND(3, 2); break; case 3: EXPAND(3, 3); break;
                       case 4: EXPAND(3, 4); break; case 8: EXPAND(3, 8); break; default: return cudaErrorInvalidValue; }
    } else {
        switch (t.E) { case 1: EXPAND(4, 1); break; case 2: EXPAND(4, 2); break; case 3: EXPAND(4, 3); break;
                       case 4: EXPAND(4, 4); break; case 8: EXPAND(4, 8); break; default: return cudaErrorInvalidValue; }
    }
#undef EXPAND
    return cudaGetLastError();
}

cudaError_t k1_dump_table(const TableView &t, const GraphView &g, const void *vrec, u64 first, u64 n, u32 *vids, u32 *labels,
                          u32 *degs, double *pde, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    k1_dump_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(t, g, reinterpret_cast<const uint4 *>(vrec), first, n, vids,
                                                                    labels, degs, pde);
    return cudaGetLastError();
}

}  // namespace gpe
