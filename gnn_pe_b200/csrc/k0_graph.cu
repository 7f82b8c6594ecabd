// k0_graph.cu -- device-side construction of everything gpe_set_graph derives from the CSR (Static_Graph, graph.h:61-63):
// validation, degrees, label classes and the join's own copy of the graph.  Setup code, not a hot path: it exists so
// that loading a 10M-vertex / 100M-edge graph takes a fraction of a second per GPU instead of minutes on one host core.
//
// The join works in CLASS ORDER: vertex v gets the id  v' = lcoff[label(v)] + lpos(v)  (its position among all vertices
// sorted by (label, id)), so that
//   * a vertex's adjacency sorted by v' is grouped by neighbour label with ids ascending inside every group -- the
//     structure the reference builds in Static_Graph::BuildLabelOffset (graph.cpp:126-160) and never uses;
//   * "the i-th vertex of label l" (bit i of a candidate bitmap, entry i of a subtree table) is v' - lcoff[l]: no lookup;
//   * the rows of the group directory and the table entries of one label class are contiguous.
// Adjacency entry (4 bytes when V <= 2^24): v' | min(degree, 255) << 24; else 8 bytes (v', min(degree, 255)).
// Group directory row of v' (narrow, max degree < 65536): u32 base | u16 rel[labels + 1]; wide: u32 abs[labels + 1].
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "gpe_internal.h"

namespace gpe {

namespace {

constexpr unsigned kFull = 0xffffffffu;

// err[0]: smallest (code << 32 | vertex) seen, err[1]: max label, err[2]: max degree, err[3]: sum of squared degrees
__global__ void __launch_bounds__(256) k0_validate_kernel(u32 V, u32 n_adj, const u32 *__restrict__ off,
                                                          const u32 *__restrict__ nbr, const u32 *__restrict__ label,
                                                          u32 *__restrict__ deg, unsigned long long *err) {
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    u32 my_label = 0, my_deg = 0;
    if (v < V) {
        const u32 a = off[v], b = off[v + 1];
        unsigned long long bad = ~0ull;
        if (b < a || b > n_adj) {
            bad = (1ull << 32) | v;
        } else {
            my_deg = b - a;
            u32 prev = 0;
            for (u32 j = a; j < b; j++) {
                const u32 w = nbr[j];
                if (w >= V) { bad = (2ull << 32) | v; break; }
                if (w == v) { bad = (3ull << 32) | v; break; }
                if (j > a && prev >= w) { bad = (4ull << 32) | v; break; }
                prev = w;
            }
        }
        deg[v] = my_deg;
        my_label = label[v];
        if (bad != ~0ull) atomicMin(err, bad);
    }
    unsigned long long sq = (unsigned long long)my_deg * my_deg;
    for (int o = 16; o; o >>= 1) {
        my_label = max(my_label, __shfl_xor_sync(kFull, my_label, o));
        my_deg = max(my_deg, __shfl_xor_sync(kFull, my_deg, o));
        sq += __shfl_xor_sync(kFull, sq, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(err + 1, (unsigned long long)my_label);
        atomicMax(err + 2, (unsigned long long)my_deg);
        atomicAdd(err + 3, sq);
    }
}

__global__ void __launch_bounds__(256) k0_label_hist_kernel(u32 V, const u32 *__restrict__ label, u32 *hist /*n_labels + 1*/) {
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = v < V;
    const u32 l = in ? label[v] : 0xffffffffu;
    const unsigned peers = __match_any_sync(kFull, l);
    if (in && (peers & ((1u << (threadIdx.x & 31)) - 1)) == 0) atomicAdd(&hist[l], (u32)__popc(peers));
}

__global__ void __launch_bounds__(256) k0_iota_kernel(u32 n, u32 *out) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

__global__ void __launch_bounds__(256) k0_max_class_kernel(u32 n_labels, const u32 *__restrict__ lcoff, u32 *out) {
    u32 m = 0;
    for (u32 l = blockIdx.x * blockDim.x + threadIdx.x; l < n_labels; l += gridDim.x * blockDim.x) m = max(m, lcoff[l + 1] - lcoff[l]);
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(kFull, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// lclass (vertices by (label, id)) -> newid, lpos, and the per-vertex arrays in class order
__global__ void __launch_bounds__(256) k0_classes_kernel(u32 V, const u32 *__restrict__ lclass, const u32 *__restrict__ label,
                                                         const u32 *__restrict__ deg, const u32 *__restrict__ lcoff,
                                                         u32 *newid, u32 *lpos, u32 *degJ, u32 *labelJ, u32 *offJ_in) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) {
        if (i == V) offJ_in[V] = 0;
        return;
    }
    const u32 v = lclass[i], l = label[v];
    newid[v] = i;
    lpos[v] = i - lcoff[l];
    degJ[i] = deg[v];
    labelJ[i] = l;
    offJ_in[i] = deg[v];  // exclusive-scanned into the row starts
}

// one warp per vertex: (v' << 32 | w') for every adjacency entry
__global__ void __launch_bounds__(256) k0_keys_kernel(u32 V, const u32 *__restrict__ off, const u32 *__restrict__ nbr,
                                                      const u32 *__restrict__ newid, u64 *keys) {
    const int lane = threadIdx.x & 31;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < V; v += nwarps) {
        const u64 hi = (u64)newid[v] << 32;
        for (u32 j = off[v] + lane; j < off[v + 1]; j += 32) keys[j] = hi | newid[nbr[j]];
    }
}

// sorted keys -> join adjacency entries, the label-grouped adjacency in original ids (k1's histogram / fill walk it),
// and the edge filter
__global__ void __launch_bounds__(256) k0_entries_kernel(u64 n_adj, const u64 *__restrict__ keys, const u32 *__restrict__ offJ,
                                                         const u32 *__restrict__ off, const u32 *__restrict__ lclass,
                                                         const u32 *__restrict__ degJ, bool wide, u32 *nbrJ, u32 *nbrG,
                                                         u64 *bloom, u64 bloom_word_mask) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_adj) return;
    const u64 key = keys[i];
    const u32 a = (u32)(key >> 32), b = (u32)key;
    const u32 d8 = min(degJ[b], 255u);
    if (wide) reinterpret_cast<uint2 *>(nbrJ)[i] = make_uint2(b, d8);
    else nbrJ[i] = b | (d8 << 24);
    nbrG[off[lclass[a]] + (u32)(i - offJ[a])] = lclass[b];
    if (a < b) {
        u64 word, bits;
        join_edge_probe(a, b, bloom_word_mask, word, bits);
        atomicOr((unsigned long long *)&bloom[word], (unsigned long long)bits);
    }
}

// group directory: thread per (v', label boundary); the row of v' is ascending in w', i.e. label-major
__global__ void __launch_bounds__(256) k0_gtab_kernel(u32 V, u32 nl, const u32 *__restrict__ offJ, const u32 *__restrict__ nbrJ,
                                                      bool wide_adj, const u32 *__restrict__ lcoff, bool wide_dir,
                                                      u32 row_bytes, unsigned char *gtab) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 v = (u32)(t / (nl + 1)), l = (u32)(t % (nl + 1));
    if (v >= V) return;
    const u32 base = offJ[v];
    u32 lo = base, hi = offJ[v + 1];
    const u32 first = lcoff[l];  // first v' of label l (lcoff[nl] = V: past every entry)
    while (lo < hi) {
        const u32 mid = lo + ((hi - lo) >> 1);
        const u32 w = wide_adj ? nbrJ[2 * (u64)mid] : nbrJ[mid] & 0xffffffu;
        if (w < first) lo = mid + 1; else hi = mid;
    }
    unsigned char *row = gtab + (u64)v * row_bytes;
    if (wide_dir) {
        reinterpret_cast<u32 *>(row)[l] = lo;
    } else {
        if (l == 0) *reinterpret_cast<u32 *>(row) = base;
        reinterpret_cast<unsigned short *>(row + 4)[l] = (unsigned short)(lo - base);
    }
}

// caller-supplied candidate ids (gpe_refine) -> class order, and back for the host
__global__ void __launch_bounds__(256) k0_gather_kernel(u64 n, const u32 *__restrict__ map, const u32 *__restrict__ in, u32 *out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = map[in[i]];
}

}  // namespace

namespace {
__global__ void __launch_bounds__(256) k0_zero_kernel(ZeroList z) {
    for (int r = 0; r < z.n; r++)
        for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < z.words[r]; i += (u64)gridDim.x * blockDim.x) z.p[r][i] = 0;
}
}  // namespace

// several small regions zeroed by ONE launch (a step of a small batch is bound by its launch count)
cudaError_t k0_zero(const ZeroList &z, cudaStream_t s) {
    u64 most = 0;
    for (int r = 0; r < z.n; r++) most = std::max(most, z.words[r]);
    if (z.n == 0 || most == 0) return cudaSuccess;
    k0_zero_kernel<<<(unsigned)std::min<u64>((most + 255) / 256, 64), 256, 0, s>>>(z);
    return cudaGetLastError();
}

cudaError_t k0_gather(u64 n, const u32 *map, const u32 *in, u32 *out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    k0_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, map, in, out);
    return cudaGetLastError();
}

cudaError_t k0_validate(u32 V, u32 n_adj, const u32 *off, const u32 *nbr, const u32 *label, u32 *deg, u64 *err3, cudaStream_t s) {
    const u64 init[4] = {~0ull, 0, 0, 0};
    cudaError_t e = cudaMemcpyAsync(err3, init, sizeof init, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    if (V) k0_validate_kernel<<<(V + 255) / 256, 256, 0, s>>>(V, n_adj, off, nbr, label, deg, reinterpret_cast<unsigned long long *>(err3));
    return cudaGetLastError();
}

static int bits_for(u64 n) {  // bits needed to represent values < n
    int b = 1;
    while (b < 64 && (1ull << b) < n) b++;
    return b;
}

cudaError_t k0_build_classes(u32 V, u32 n_labels, const u32 *label, const u32 *deg, u32 *lcoff, u32 *lclass, u32 *lpos,
                             u32 *newid, u32 *degJ, u32 *labelJ, u32 *offJ, u32 *max_class_dev, DevBuf &tmp, cudaStream_t s) {
    cudaError_t e;
    // label histogram -> class offsets
    if ((e = cudaMemsetAsync(lcoff, 0, ((size_t)n_labels + 2) * sizeof(u32), s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(max_class_dev, 0, sizeof(u32), s)) != cudaSuccess) return e;
    if (V) k0_label_hist_kernel<<<(V + 255) / 256, 256, 0, s>>>(V, label, lcoff);
    size_t scan_bytes = 0, sort_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, lcoff, lcoff, (int)n_labels + 1, s);
    size_t scan2_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan2_bytes, offJ, offJ, (int)V + 1, s);
    const int lbits = bits_for(std::max<u64>(n_labels, 2));
    u32 *ids_in = nullptr, *keys_out = nullptr;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, label, keys_out, ids_in, lclass, (int)V, 0, lbits, s);
    const size_t a = (std::max(std::max(scan_bytes, scan2_bytes), sort_bytes) + 255) / 256 * 256;
    if ((e = tmp.reserve(a + 2 * ((size_t)V + 1) * sizeof(u32))) != cudaSuccess) return e;
    unsigned char *base = tmp.as<unsigned char>();
    ids_in = reinterpret_cast<u32 *>(base + a);
    keys_out = ids_in + V + 1;
    if ((e = cub::DeviceScan::ExclusiveSum(base, scan_bytes, lcoff, lcoff, (int)n_labels + 1, s)) != cudaSuccess) return e;
    if (n_labels) k0_max_class_kernel<<<std::min<u32>((n_labels + 255) / 256, 1024), 256, 0, s>>>(n_labels, lcoff, max_class_dev);
    if (V) {
        k0_iota_kernel<<<(V + 255) / 256, 256, 0, s>>>(V, ids_in);
        // stable: ids stay ascending inside a label class
        if ((e = cub::DeviceRadixSort::SortPairs(base, sort_bytes, label, keys_out, ids_in, lclass, (int)V, 0, lbits, s)) != cudaSuccess) return e;
    }
    k0_classes_kernel<<<(V + 1 + 255) / 256, 256, 0, s>>>(V, lclass, label, deg, lcoff, newid, lpos, degJ, labelJ, offJ);
    if ((e = cub::DeviceScan::ExclusiveSum(base, scan2_bytes, offJ, offJ, (int)V + 1, s)) != cudaSuccess) return e;
    return cudaGetLastError();
}

cudaError_t k0_build_join_graph(u32 V, u32 n_adj, u32 n_labels, const u32 *off, const u32 *nbr, const u32 *lclass,
                                const u32 *newid, const u32 *degJ, const u32 *offJ, const u32 *lcoff, bool wide_adj,
                                bool wide_dir, u32 dir_row_bytes, u32 *nbrJ, u32 *nbrG, void *gtab, u64 *bloom, u64 bloom_bits,
                                DevBuf &tmp, int sm_count, cudaStream_t s) {
    cudaError_t e;
    if ((e = cudaMemsetAsync(bloom, 0, bloom_bits / 8, s)) != cudaSuccess) return e;
    if (n_adj) {
        size_t sort_bytes = 0;
        u64 *k_in = nullptr, *k_out = nullptr;
        const int end_bit = 32 + bits_for(std::max<u64>(V, 2));
        cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, k_in, k_out, (int)n_adj, 0, end_bit, s);
        const size_t a = (sort_bytes + 255) / 256 * 256;
        if ((e = tmp.reserve(a + 2 * (size_t)n_adj * sizeof(u64))) != cudaSuccess) return e;
        k_in = reinterpret_cast<u64 *>(tmp.as<unsigned char>() + a);
        k_out = k_in + n_adj;
        k0_keys_kernel<<<sm_count * 8, 256, 0, s>>>(V, off, nbr, newid, k_in);
        if ((e = cub::DeviceRadixSort::SortKeys(tmp.p, sort_bytes, k_in, k_out, (int)n_adj, 0, end_bit, s)) != cudaSuccess) return e;
        k0_entries_kernel<<<(unsigned)(((u64)n_adj + 255) / 256), 256, 0, s>>>(n_adj, k_out, offJ, off, lclass, degJ, wide_adj, nbrJ,
                                                                             nbrG, bloom, (bloom_bits >> 6) - 1);
    }
    if (V) {
        const u64 threads = (u64)V * (n_labels + 1);
        k0_gtab_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(V, n_labels, offJ, nbrJ, wide_adj, lcoff, wide_dir,
                                                                        dir_row_bytes, reinterpret_cast<unsigned char *>(gtab));
    }
    return cudaGetLastError();
}

}  // namespace gpe
