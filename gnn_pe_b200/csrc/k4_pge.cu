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
// STATUS: written at the end of round 1 after the GPU budget was spent -- compiles for sm_100a, parity against the
// (pinned) oracle is NOT yet verified on a GPU; the tests that exercise it run in a subprocess and are marked xfail
// until they have been seen green.  Nothing on the default path calls into this file.
#include "gpe_internal.h"

namespace gpe {

namespace {

// thread per vertex: depth-first walk over the simple paths of pl vertices from v, folding min/max per dimension
__global__ void __launch_bounds__(128) k4_pge_groups_kernel(GraphView g, u32 pl, const double *__restrict__ x /*V x e*/,
                                                            PgeView p) {
    const u32 e = g.e, pde = pl * e;
    for (u32 v = blockIdx.x * blockDim.x + threadIdx.x; v < g.V; v += gridDim.x * blockDim.x) {
        double lo[kMaxL * kMaxE], hi[kMaxL * kMaxE], llo[kMaxL * kMaxE], lhi[kMaxL * kMaxE];
        u32 path[kMaxL], cur[kMaxL + 1];
        bool first = true;
        path[0] = v;
        u32 len = 1;
        cur[1] = g.off[v];
        while (len >= 1) {
            if (len == pl) {
                for (u32 j = 0; j < pl; j++)
                    for (u32 k = 0; k < e; k++) {
                        const double a = g.vde[(u64)path[j] * e + k], b = x[(u64)path[j] * e + k];
                        const u32 d = j * e + k;
                        if (first) { lo[d] = hi[d] = a; llo[d] = lhi[d] = b; }
                        else {
                            if (lo[d] > a) lo[d] = a;
                            if (hi[d] < a) hi[d] = a;
                            if (llo[d] > b) llo[d] = b;
                            if (lhi[d] < b) lhi[d] = b;
                        }
                    }
                first = false;
                len--;
                continue;
            }
            const u32 node = path[len - 1];
            if (cur[len] < g.off[node + 1]) {
                const u32 nb = g.nbr[cur[len]++];
                bool seen = false;
                for (u32 t = 0; t < len; t++) seen = seen || path[t] == nb;
                if (seen) continue;
                path[len] = nb;
                len++;
                if (len < pl) cur[len] = g.off[nb];
            } else {
                len--;
            }
        }
        if (first)  // no path of pl vertices: [vde, vde | 0 ...] and [x, x | 0 ...] (src/main.cpp:103-121)
            for (u32 d = 0; d < pde; d++) {
                lo[d] = hi[d] = d < e ? g.vde[(u64)v * e + d] : 0.0;
                llo[d] = lhi[d] = d < e ? x[(u64)v * e + d] : 0.0;
            }
        const u32 lab = g.label[v];
        const u64 idx = (u64)g.lcoff[lab] + g.lpos[v];  // class order: a label class is one contiguous run of rows
        p.deg[idx] = g.deg[v];
        p.has[idx] = first ? 0 : 1;
        for (u32 d = 0; d < pde; d++) {
            p.pg_lo[(u64)d * g.V + idx] = lo[d];
            p.pg_hi[(u64)d * g.V + idx] = hi[d];
            p.plg_lo[(u64)d * g.V + idx] = llo[d];
            p.plg_hi[(u64)d * g.V + idx] = lhi[d];
        }
    }
}

// thread per data vertex (class order); the query vertex slots of its label are tested one after the other
__global__ void __launch_bounds__(256) k4_pge_scan_kernel(PgeView p, u32 V, u32 pde, u32 n_labels,
                                                          const u32 *__restrict__ lcoff,
                                                          const u32 *__restrict__ label_slot_off /*n_labels + 1*/,
                                                          const u32 *__restrict__ slot_list,
                                                          const u32 *__restrict__ q_deg, const double *__restrict__ q_pg_lo,
                                                          const double *__restrict__ q_plg_lo,
                                                          const double *__restrict__ q_plg_hi /*slot x pde*/,
                                                          u32 *bitmap, u64 words_per_slot, u64 *survivors) {
    u64 my = 0;
    for (u32 idx = blockIdx.x * blockDim.x + threadIdx.x; idx < V; idx += gridDim.x * blockDim.x) {
        u32 lo = 0, hi = n_labels;  // label of this row: last l with lcoff[l] <= idx
        while (hi - lo > 1) {
            const u32 mid = (lo + hi) >> 1;
            if (lcoff[mid] <= idx) lo = mid; else hi = mid;
        }
        const u32 lab = lo, pos = idx - lcoff[lab], dg = p.deg[idx];
        for (u32 i = label_slot_off[lab]; i < label_slot_off[lab + 1]; i++) {
            const u32 s = slot_list[i];
            bool ok = q_deg[s] <= dg;
            for (u32 k = 0; ok && k < pde; k++) {
                const double vl = p.plg_lo[(u64)k * V + idx], vh = p.plg_hi[(u64)k * V + idx];
                ok = !(vh < q_plg_lo[(u64)s * pde + k] || vl > q_plg_hi[(u64)s * pde + k]);
            }
            for (u32 k = 0; ok && k < pde; k++) ok = !(p.pg_hi[(u64)k * V + idx] < q_pg_lo[(u64)s * pde + k]);
            if (ok) {
                atomicOr(bitmap + (u64)s * words_per_slot + (pos >> 5), 1u << (pos & 31));
                my++;
            }
        }
    }
    for (int o = 16; o; o >>= 1) my += __shfl_xor_sync(0xffffffffu, my, o);
    if ((threadIdx.x & 31) == 0 && my) atomicAdd((unsigned long long *)survivors, (unsigned long long)my);
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
    k4_pge_groups_kernel<<<sm_count * 8, 128, 0, s>>>(g, pl, d_x, p);
    return cudaGetLastError();
}

cudaError_t k4_pge_scan(const PgeView &p, u32 V, u32 pde, u32 n_labels, const u32 *lcoff, const u32 *label_slot_off,
                        const u32 *slot_list, const u32 *q_deg, const double *q_pg_lo, const double *q_plg_lo,
                        const double *q_plg_hi, u32 *bitmap, u64 words_per_slot, u64 *survivors, int sm_count, cudaStream_t s) {
    if (V == 0) return cudaSuccess;
    k4_pge_scan_kernel<<<sm_count * 8, 256, 0, s>>>(p, V, pde, n_labels, lcoff, label_slot_off, slot_list, q_deg, q_pg_lo,
                                                    q_plg_lo, q_plg_hi, bitmap, words_per_slot, survivors);
    return cudaGetLastError();
}

cudaError_t k4_pge_dump(const PgeView &p, const GraphView &g, u32 pde, double *pg, double *plg, unsigned char *has,
                        cudaStream_t s) {
    if (g.V == 0) return cudaSuccess;
    k4_pge_dump_kernel<<<(g.V + 255) / 256, 256, 0, s>>>(p, g, pde, pg, plg, has);
    return cudaGetLastError();
}

}  // namespace gpe
