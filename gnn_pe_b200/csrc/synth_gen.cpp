// synth_gen.cpp -- seeded synthetic data graphs at BASELINE.json's larger sizes (10 M vertices / 100 M edges), built
// into gnn_pe_b200/libgpe_synth.so.  Measurement infrastructure: the reference ships no generator, and the numpy one
// (synth.py) needs minutes at this size, which on a multi-GPU box is paid once per GPU-minute.
//
// The graph is a pure function of (V, draws, n_labels, seed): edge i joins hash(seed, 2i) % V and hash(seed, 2i+1) % V;
// self loops and duplicate edges are dropped (the reference wants a simple graph), so the edge count is `draws` minus a
// few hundred.  Labels are hash(seed ^ c, v) % n_labels.  The result does not depend on the number of threads.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <parallel/algorithm>
#include <thread>
#include <vector>

namespace {

inline uint64_t mix(uint64_t x) {  // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct Synth {
    uint32_t V = 0;
    std::vector<uint32_t> offsets, nbrs, labels;
};

template <class F>
void parallel_for(uint64_t n, int threads, F f) {
    threads = std::max(1, threads);
    std::vector<std::thread> pool;
    const uint64_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        const uint64_t lo = std::min<uint64_t>(n, t * per), hi = std::min<uint64_t>(n, lo + per);
        if (lo < hi) pool.emplace_back([=] { f(lo, hi); });
    }
    for (auto &th : pool) th.join();
}

}  // namespace

extern "C" {

void *gpe_synth_uniform(uint32_t V, uint64_t draws, uint32_t n_labels, uint64_t seed, int threads) {
    if (V < 2 || n_labels == 0) return nullptr;
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    Synth *s = new Synth();
    s->V = V;
    std::vector<uint64_t> keys(draws);
    const uint64_t s1 = mix(seed), none = ~0ull;
    parallel_for(draws, threads, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; i++) {
            const uint64_t a = mix(s1 ^ (2 * i)) % V, b = mix(s1 ^ (2 * i + 1)) % V;
            keys[i] = a == b ? none : std::min(a, b) * V + std::max(a, b);
        }
    });
    __gnu_parallel::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    if (!keys.empty() && keys.back() == none) keys.pop_back();
    const uint64_t E = keys.size();
    if (2 * E >= (1ull << 32)) { delete s; return nullptr; }  // the reference's offsets are 32-bit (graph.h:61)
    std::vector<std::atomic<uint32_t>> deg(V);
    parallel_for(V, threads, [&](uint64_t lo, uint64_t hi) { for (uint64_t v = lo; v < hi; v++) deg[v].store(0, std::memory_order_relaxed); });
    parallel_for(E, threads, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; i++) {
            deg[keys[i] / V].fetch_add(1, std::memory_order_relaxed);
            deg[keys[i] % V].fetch_add(1, std::memory_order_relaxed);
        }
    });
    s->offsets.resize((size_t)V + 1);
    s->offsets[0] = 0;
    for (uint32_t v = 0; v < V; v++) s->offsets[v + 1] = s->offsets[v] + deg[v].load(std::memory_order_relaxed);
    s->nbrs.resize(2 * E);
    parallel_for(V, threads, [&](uint64_t lo, uint64_t hi) { for (uint64_t v = lo; v < hi; v++) deg[v].store(0, std::memory_order_relaxed); });
    parallel_for(E, threads, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; i++) {
            const uint32_t a = (uint32_t)(keys[i] / V), b = (uint32_t)(keys[i] % V);
            s->nbrs[s->offsets[a] + deg[a].fetch_add(1, std::memory_order_relaxed)] = b;
            s->nbrs[s->offsets[b] + deg[b].fetch_add(1, std::memory_order_relaxed)] = a;
        }
    });
    parallel_for(V, threads, [&](uint64_t lo, uint64_t hi) {  // ascending adjacency (graph.cpp:231-233); also what makes the result
        for (uint64_t v = lo; v < hi; v++)                     // independent of the order the threads filled it in
            std::sort(s->nbrs.begin() + s->offsets[v], s->nbrs.begin() + s->offsets[v + 1]);
    });
    s->labels.resize(V);
    const uint64_t s2 = mix(seed ^ 0x6C6162656C73ull);
    parallel_for(V, threads, [&](uint64_t lo, uint64_t hi) { for (uint64_t v = lo; v < hi; v++) s->labels[v] = (uint32_t)(mix(s2 ^ v) % n_labels); });
    return s;
}

uint64_t gpe_synth_adjacency_entries(void *h) { return h ? ((Synth *)h)->nbrs.size() : 0; }

void gpe_synth_copy(void *h, uint32_t *offsets, uint32_t *nbrs, uint32_t *labels) {
    Synth *s = (Synth *)h;
    std::memcpy(offsets, s->offsets.data(), s->offsets.size() * sizeof(uint32_t));
    if (!s->nbrs.empty()) std::memcpy(nbrs, s->nbrs.data(), s->nbrs.size() * sizeof(uint32_t));
    std::memcpy(labels, s->labels.data(), s->labels.size() * sizeof(uint32_t));
}

void gpe_synth_free(void *h) { delete (Synth *)h; }

}  // extern "C"
