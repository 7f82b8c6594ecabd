// k4_pge.cu -- GNN-PGE, the reference's sibling variant of the filter (citations relative to
// the reference's GNN-PGE/ directory).  One row per data VERTEX instead of one per path:
//   * offline (src/main.cpp:91-176): every vertex gets the bounding box of the embeddings of all simple paths of `pl`
//     vertices that start at it ("path group") -- over the dominance embeddings (pg) and over the label embeddings (plg);
//   * online (include/custom.h:292-367, leaf test :332-367): a data vertex v is a candidate of a query vertex u iff
//     label(u) == label(v), deg(u) <= deg(v), the label boxes overlap in every dimension, and v's pg upper corner is not
//     below u's pg lower corner in any dimension (FP64, no epsilon at the leaf).
// The R*-tree over the vertex boxes becomes a scan of the label class of every query vertex.  Candidates land in the
// same class-local bitmaps as the path filter's, so compaction, matching order and join are shared.
//
// Verified on B200 against the golden vectors of the unmodified GNN-PGE binary (tests/test_gpu_pge.py).
//
// Kernels: the group build is one WARP per vertex (like k1: lanes over the innermost adjacency list, ballots decide which
// outer vertices lie on a full-length path; min / max are exact in any order, so the boxes are bit-identical to the
// reference's serial fold).  The scan is HBM-bound on paper -- (3 * pl * e * 8 + 4) bytes per data vertex of a label
// class that some query vertex asks for (the lower corner of pg is never tested) -- with the query records of the
// label staged in shared memory and one bitmap word per warp and slot instead of one atomic per survivor.
#include "gpe_internal.h"

namespace gpe {

namespace {

constexpr unsigned kFull = 0xffffffffu;

struct Box {  // running bounding boxes of one position, all lanes (outer positions: warp-uniform values)
    double lo[kMaxE], hi[kMaxE], llo[kMaxE], lhi[kMaxE];
};

__device__ __forceinline__ void box_init(Box &b, u32 e) {
    for (u32 k = 0; k < e; k++) { b.lo[k] = b.llo[k] = 1e300; b.hi[k] = b.lhi[k] = -1e300; }
}
__device__ __forceinline__ void box_fold(Box &b, const GraphView &g, const double *__restrict__ x, u32 v, u32 e) {
    for (u32 k = 0; k < e; k++) {
        const double a = g.vde[(u64)v * e + k], l = x[(u64)v * e + k];
        b.lo[k] = fmin(b.lo[k], a);
        b.hi[k] = fmax(b.hi[k], a);
        b.llo[k] = fmin(b.llo[k], l);
        b.lhi[k] = fmax(b.lhi[k], l);
    }
}
__device__ __forceinline__ void box_reduce(Box &b, u32 e) {
    for (u32 k = 0; k < e; k++)
        for (int o = 16; o; o >>= 1) {
            b.lo[k] = fmin(b.lo[k], __shfl_xor_sync(kFull, b.lo[k], o));
            b.hi[k] = fmax(b.hi[k], __shfl_xor_sync(kFull, b.hi[k], o));
            b.llo[k] = fmin(b.llo[k], __shfl_xor_sync(kFull, b.llo[k], o));
            b.lhi[k] = fmax(b.lhi[k], __shfl_xor_sync(kFull, b.lhi[k], o));
        }
}

// one warp per vertex v: all simple paths (v, b [, c [, d]]) of PL vertices; position j's box takes a vertex iff it lies on
// at least one full-length path (src/main.cpp:91-176 folds complete paths only)
template <int PL>
__global__ void __launch_bounds__(256) k4_pge_groups_kernel(GraphView g, const double *__restrict__ x /*V x e*/, PgeView p) {
    const u32 e = g.e;
    const int lane = threadIdx.x & 31;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < g.V; v += nwarps) {
        Box box[PL > 1 ? PL - 1 : 1];  // positions 1 .. PL-1 (position 0 is v itself)
        for (int j = 0; j < PL - 1; j++) box_init(box[j], e);
        bool any = PL == 1;
        const u32 v0 = g.off[v], v1 = g.off[v + 1];
        if (PL == 2) {
            for (u32 j = v0 + lane; j < v1; j += 32) { box_fold(box[0], g, x, g.nbr[j], e); any = true; }
        } else if (PL >= 3) {
            for (u32 s = v0; s < v1; ++s) {
                const u32 b = g.nbr[s];
                const u32 b0 = g.off[b], b1 = g.off[b + 1];
                bool any_b = false;
                if (PL == 3) {
                    for (u32 j = b0 + lane; j < b1; j += 32) {
                        const u32 c = g.nbr[j];
                        if (c == v) continue;
                        box_fold(box[1], g, x, c, e);
                        any_b = true;
                    }
                } else {
                    for (u32 t = b0; t < b1; ++t) {
                        const u32 c = g.nbr[t];
                        if (c == v) continue;
                        bool any_c = false;
                        for (u32 j = g.off[c] + lane; j < g.off[c + 1]; j += 32) {
                            const u32 d = g.nbr[j];
                            if (d == v || d == b) continue;
                            box_fold(box[PL - 2], g, x, d, e);
                            any_c = true;
                        }
                        if (__any_sync(kFull, any_c)) { box_fold(box[PL >= 4 ? 1 : 0], g, x, c, e); any_b = true; }
                    }
                }
                if (__any_sync(kFull, any_b)) { box_fold(box[0], g, x, b, e); any = true; }
            }
        }
        any = __any_sync(kFull, any);
        if (PL > 1) box_reduce(box[PL - 2], e);  // the innermost position was folded lane by lane
        const u32 lab = g.label[v];
        const u64 idx = (u64)g.lcoff[lab] + g.lpos[v];  // class order: a label class is one contiguous run of rows
        if (lane == 0) {
            p.deg[idx] = g.deg[v];
            p.has[idx] = (any && PL > 1) || PL == 1 ? 1 : 0;
        }
        // position 0 is v; without any full path the box is [vde, vde | 0 ...] and [x, x | 0 ...] (src/main.cpp:103-121)
        for (u32 d = lane; d < (u32)PL * e; d += 32) {
            const u32 j = d / e, k = d % e;
            double lo, hi, llo, lhi;
            if (j == 0) {
                lo = hi = g.vde[(u64)v * e + k];
                llo = lhi = x[(u64)v * e + k];
            } else if (!any) {
                lo = hi = llo = lhi = 0.0;
            } else {
                // (a lane cannot index a register array by a run-time position: select)
                lo = hi = llo = lhi = 0.0;
                for (int jj = 0; jj < PL - 1; jj++)
                    for (u32 kk = 0; kk < e; kk++)
                        if ((u32)jj + 1 == j && kk == k) { lo = box[jj].lo[kk]; hi = box[jj].hi[kk]; llo = box[jj].llo[kk]; lhi = box[jj].lhi[kk]; }
            }
            p.pg_lo[(u64)d * g.V + idx] = lo;
            p.pg_hi[(u64)d * g.V + idx] = hi;
            p.plg_lo[(u64)d * g.V + idx] = llo;
            p.plg_hi[(u64)d * g.V + idx] = lhi;
        }
    }
}

// Scan: a CTA takes tiles of 256 consecutive rows (class order, so a tile holds one label, two at a class boundary); a
// thread keeps its row's columns in registers, the query records of the row's label come through shared memory in chunks,
// and the 32 rows of a warp -- consecutive bit positions of the slot's class-local bitmap -- leave as at most two words.
constexpr int kPgeChunk = 16;  // query vertex slots staged at a time
template <int PDE>
__global__ void __launch_bounds__(256) k4_pge_scan_kernel(PgeView p, u32 V, u32 n_labels, const u32 *__restrict__ lcoff,
                                                          const u32 *__restrict__ label_slot_off /*n_labels + 1*/,
                                                          const u32 *__restrict__ slot_list,
                                                          const u32 *__restrict__ q_deg, const double *__restrict__ q_pg_lo,
                                                          const double *__restrict__ q_plg_lo,
                                                          const double *__restrict__ q_plg_hi /*slot x pde*/,
                                                          u32 *bitmap, u64 words_per_slot, u64 *survivors, u64 *rows_examined) {
    __shared__ double s_q[kPgeChunk][3 * PDE];
    __shared__ u32 s_deg[kPgeChunk], s_slot[kPgeChunk];
    const int lane = threadIdx.x & 31;
    u64 my = 0, my_rows = 0;
    const u32 n_tiles = (V + 255) / 256;
    for (u32 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const u32 idx = tile * 256 + threadIdx.x, first = tile * 256, last = min(first + 255, V - 1);
        // labels of the tile's first and last row: last l with lcoff[l] <= row
        u32 l_first, l_last;
        {
            u32 lo = 0, hi = n_labels;
            while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (lcoff[mid] <= first) lo = mid; else hi = mid; }
            l_first = lo;
            hi = n_labels;
            while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (lcoff[mid] <= last) lo = mid; else hi = mid; }
            l_last = lo;
        }
        bool any_slots = false;
        for (u32 l = l_first; l <= l_last; l++) any_slots = any_slots || label_slot_off[l + 1] > label_slot_off[l];
        if (!any_slots) continue;  // nobody asks for this label: the rows are not read
        const bool in = idx < V;
        double vl[PDE], vh[PDE], ph[PDE];
        u32 dg = 0;
        if (in) {
            dg = p.deg[idx];
#pragma unroll
            for (int k = 0; k < PDE; k++) {
                vl[k] = __ldcs(p.plg_lo + (u64)k * V + idx);
                vh[k] = __ldcs(p.plg_hi + (u64)k * V + idx);
                ph[k] = __ldcs(p.pg_hi + (u64)k * V + idx);
            }
        }
        for (u32 l = l_first; l <= l_last; l++) {
            const u32 c0 = lcoff[l], c1 = lcoff[l + 1];
            const bool mine = in && idx >= c0 && idx < c1;
            if (mine) my_rows++;
            for (u32 i0 = label_slot_off[l]; i0 < label_slot_off[l + 1]; i0 += kPgeChunk) {
                const u32 n = min((u32)kPgeChunk, label_slot_off[l + 1] - i0);
                __syncthreads();
                for (u32 t = threadIdx.x; t < n * 3 * PDE; t += blockDim.x) {
                    const u32 j = t / (3 * PDE), r = t % (3 * PDE), s = slot_list[i0 + j];
                    const double *src = r < PDE ? q_plg_lo : r < 2 * PDE ? q_plg_hi : q_pg_lo;
                    s_q[j][r] = src[(u64)s * PDE + r % PDE];
                }
                if (threadIdx.x < n) { s_slot[threadIdx.x] = slot_list[i0 + threadIdx.x]; s_deg[threadIdx.x] = q_deg[slot_list[i0 + threadIdx.x]]; }
                __syncthreads();
                for (u32 j = 0; j < n; j++) {
                    bool ok = mine && s_deg[j] <= dg;
#pragma unroll
                    for (int k = 0; k < PDE; k++)  // label boxes overlap, upper corner not below the query's lower corner
                        ok = ok && !(vh[k] < s_q[j][k] || vl[k] > s_q[j][PDE + k]) && !(ph[k] < s_q[j][2 * PDE + k]);
                    const unsigned m = __ballot_sync(kFull, ok);
                    if (!m) continue;
                    my += ok ? 1 : 0;
                    // the warp's 32 rows are consecutive bit positions base_pos .. base_pos + 31 of the slot's class-local
                    // bitmap (negative / beyond the class for rows of a neighbouring class, whose `ok` is false): they fall
                    // into at most two words, written by lane 0 -- no per-survivor atomics, no match instruction
                    if (lane == 0) {
                        const int base_pos = (int)(idx) - (int)c0;  // lane 0's position
                        const int sh = ((base_pos % 32) + 32) % 32;
                        const int w_first = (base_pos - sh) / 32;   // floor(base_pos / 32)
                        const unsigned m0 = sh ? m & ((1u << (32 - sh)) - 1) : m, m1 = sh ? m >> (32 - sh) : 0u;
                        u32 *dst = bitmap + (u64)s_slot[j] * words_per_slot;
                        if (m0) atomicOr(dst + w_first, m0 << sh);
                        if (m1) atomicOr(dst + w_first + 1, m1);
                    }
                }
            }
        }
    }
    for (int o = 16; o; o >>= 1) { my += __shfl_xor_sync(kFull, my, o); my_rows += __shfl_xor_sync(kFull, my_rows, o); }
    if (lane == 0 && my) atomicAdd((unsigned long long *)survivors, (unsigned long long)my);
    if (lane == 0 && my_rows) atomicAdd((unsigned long long *)rows_examined, (unsigned long long)my_rows);
}

__global__ void k4_pge_dump_kernel(PgeView p, GraphView g, u32 pde, double *pg, double *plg, unsigned char *has) {
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g.V) return;
    const u64 idx = (u64)g.lcoff[g.label[v]] + g.lpos[v];
    has[v] = p.has[idx];
    for (u32 d = 0; d < pde; d++) {
        pg[((u64)v * pde + d) * 2] = p.pg_lo[(u64)d * g.V + idx];
        pg[((u64)v * pde + d) * 2 + 1] = p.pg_hi[(u64)d * g.V + idx];
        plg[((u64)v * pde + d) * 2] = p.plg_lo[(u64)d * g.V + idx];
        plg[((u64)v * pde + d) * 2 + 1] = p.plg_hi[(u64)d * g.V + idx];
    }
}

}  // namespace

size_t k4_pge_bytes(u32 V, u32 pde) { return (size_t)V * (4 + 1 + 3) + (size_t)4 * pde * V * sizeof(double) + 64; }

PgeView k4_pge_view(void *buf, u32 V, u32 pde) {
    PgeView p;
    unsigned char *b = reinterpret_cast<unsigned char *>(buf);
    p.pg_lo = reinterpret_cast<double *>(b); b += (size_t)pde * V * sizeof(double);
    p.pg_hi = reinterpret_cast<double *>(b); b += (size_t)pde * V * sizeof(double);
    p.plg_lo = reinterpret_cast<double *>(b); b += (size_t)pde * V * sizeof(double);
    p.plg_hi = reinterpret_cast<double *>(b); b += (size_t)pde * V * sizeof(double);
    p.deg = reinterpret_cast<u32 *>(b); b += (size_t)V * sizeof(u32);
    p.has = b;
    return p;
}

cudaError_t k4_pge_groups(const GraphView &g, u32 pl, const double *d_x, const PgeView &p, int sm_count, cudaStream_t s) {
    if (g.V == 0) return cudaSuccess;
    const unsigned grid = (unsigned)sm_count * 8;
    switch (pl) {
        case 1: k4_pge_groups_kernel<1><<<grid, 256, 0, s>>>(g, d_x, p); break;
        case 2: k4_pge_groups_kernel<2><<<grid, 256, 0, s>>>(g, d_x, p); break;
        case 3: k4_pge_groups_kernel<3><<<grid, 256, 0, s>>>(g, d_x, p); break;
        case 4: k4_pge_groups_kernel<4><<<grid, 256, 0, s>>>(g, d_x, p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t k4_pge_scan(const PgeView &p, u32 V, u32 pde, u32 n_labels, const u32 *lcoff, const u32 *label_slot_off,
                        const u32 *slot_list, const u32 *q_deg, const double *q_pg_lo, const double *q_plg_lo,
                        const double *q_plg_hi, u32 *bitmap, u64 words_per_slot, u64 *survivors, u64 *rows_examined, int sm_count,
                        cudaStream_t s) {
    if (V == 0) return cudaSuccess;
    const unsigned grid = std::min<unsigned>((V + 255) / 256, (unsigned)sm_count * 8);
#define PGE_SCAN(N)                                                                                                      \
    case N:                                                                                                              \
        k4_pge_scan_kernel<N><<<grid, 256, 0, s>>>(p, V, n_labels, lcoff, label_slot_off, slot_list, q_deg, q_pg_lo, q_plg_lo, \
                                                   q_plg_hi, bitmap, words_per_slot, survivors, rows_examined);          \
        break;
    switch (pde) {
        PGE_SCAN(1) PGE_SCAN(2) PGE_SCAN(3) PGE_SCAN(4) PGE_SCAN(6) PGE_SCAN(8) PGE_SCAN(9) PGE_SCAN(12) PGE_SCAN(16)
        default: return cudaErrorInvalidValue;
    }
#undef PGE_SCAN
    return cudaGetLastError();
}

bool k4_pge_supported(u32 pde) {
    return pde == 1 || pde == 2 || pde == 3 || pde == 4 || pde == 6 || pde == 8 || pde == 9 || pde == 12 || pde == 16;
}

cudaError_t k4_pge_dump(const PgeView &p, const GraphView &g, u32 pde, double *pg, double *plg, unsigned char *has,
                        cudaStream_t s) {
    if (g.V == 0) return cudaSuccess;
    k4_pge_dump_kernel<<<(g.V + 255) / 256, 256, 0, s>>>(p, g, pde, pg, plg, has);
    return cudaGetLastError();
}

}  // namespace gpe
