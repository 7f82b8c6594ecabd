// host_ref.cpp -- host-side mirror of the reference's cheap, serial steps: the `.graph` loader,
// vertex embeddings and the per-query plan.  Pure host C++ (no CUDA); these run in microseconds to
// milliseconds and feed the kernels.  Product code: never includes anything from oracle/.
//
// Citations are relative to the reference's GNN-PE/ directory.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <numeric>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include "gpe.h"
#include "host_ref.h"

namespace gpe {

// libsrc/graph/graph.cpp:163-242.  Same accepted grammar: 't V E', 'v id label degree', 'e u v';
// CSR offsets come from the declared degrees (graph.cpp:192), adjacency sorted ascending (:231-233).
int load_graph_file(const char *path, HostGraph &g, std::string &err) {
    std::ifstream in(path);
    if (!in.is_open()) {
        err = std::string("Can not open the graph file ") + path + " .";  // graph.cpp:167
        return GPE_ERR_IO;
    }
    char tag;
    uint32_t V = 0, E = 0;
    in >> tag >> V >> E;
    if (!in || tag != 't') { err = std::string(path) + " does not start with a 't V E' line"; return GPE_ERR_INVALID; }
    {   // a 'v' line takes at least 7 bytes and an 'e' line 5: refuse a header the file cannot back before allocating for it
        const std::streampos here = in.tellg();
        in.seekg(0, std::ios::end);
        const uint64_t bytes = (uint64_t)in.tellg();
        in.seekg(here);
        if ((uint64_t)V * 7 + (uint64_t)E * 5 > bytes) { err = "the header promises more vertices and edges than the file holds"; return GPE_ERR_INVALID; }
    }
    g.offsets.assign((size_t)V + 1, 0);
    g.nbrs.assign((size_t)E * 2, 0);
    g.labels.assign(V, 0);
    std::vector<uint32_t> used(V, 0);
    while (in >> tag) {
        if (tag == 'v') {
            uint32_t id, label, degree;
            if (!(in >> id >> label >> degree)) { err = "truncated 'v' line"; return GPE_ERR_INVALID; }
            if (id >= V) { err = "vertex id out of range"; return GPE_ERR_INVALID; }
            g.labels[id] = label;
            g.offsets[id + 1] = g.offsets[id] + degree;
        } else if (tag == 'e') {
            uint32_t u, v;
            if (!(in >> u >> v)) { err = "truncated 'e' line"; return GPE_ERR_INVALID; }
            if (u >= V || v >= V) { err = "edge endpoint out of range"; return GPE_ERR_INVALID; }
            size_t pu = (size_t)g.offsets[u] + used[u], pv = (size_t)g.offsets[v] + used[v];
            if (pu >= g.nbrs.size() || pv >= g.nbrs.size() || used[u] >= g.offsets[u + 1] - g.offsets[u] ||
                used[v] >= g.offsets[v + 1] - g.offsets[v]) {
                err = "declared vertex degrees do not match the edge list";
                return GPE_ERR_INVALID;
            }
            g.nbrs[pu] = v;
            g.nbrs[pv] = u;
            used[u]++;
            used[v]++;
        }
    }
    // The reference trusts the header and the declared degrees (graph.cpp:176-229) and reads or writes out of bounds when
    // they disagree with the edge list; here that is an error.
    for (uint32_t v = 0; v < V; v++)
        if (g.offsets[v + 1] < g.offsets[v] || used[v] != g.offsets[v + 1] - g.offsets[v]) {
            err = "declared vertex degrees do not match the edge list";
            return GPE_ERR_INVALID;
        }
    if (g.offsets[V] != g.nbrs.size()) { err = "the header's edge count does not match the vertex degrees"; return GPE_ERR_INVALID; }
    for (uint32_t v = 0; v < V; v++) std::sort(g.nbrs.begin() + g.offsets[v], g.nbrs.begin() + g.offsets[v + 1]);
    return GPE_OK;
}

// custom.h:492-511: mt19937 seeded with the label, `e` draws of uniform_real_distribution<double>(0,1),
// normalised by their sum.  Uses the very libstdc++ facilities the reference uses.
static void label_embedding(uint32_t label, uint32_t e, double *out) {
    std::mt19937 gen(label);
    std::uniform_real_distribution<double> dis(0.0, 1.0);
    for (uint32_t i = 0; i < e; i++) out[i] = dis(gen);
    double sum = std::accumulate(out, out + e, 0.0);
    for (uint32_t i = 0; i < e; i++) out[i] = out[i] / sum;
}

void LabelTable::fill(const uint32_t *labels, size_t n, uint32_t e_) {
    if (e_ != e) { e = e_; x.clear(); have.clear(); }
    for (size_t i = 0; i < n; i++) {
        const uint32_t lab = labels[i];
        if (lab >= (1u << 24)) continue;  // sparse huge labels: computed on the fly
        if (lab >= have.size()) { have.resize((size_t)lab + 1, 0); x.resize(((size_t)lab + 1) * e, 0.0); }
        if (!have[lab]) { label_embedding(lab, e, &x[(size_t)lab * e]); have[lab] = 1; }
    }
}

// custom.h:513-544.
void gen_vde(uint32_t V, const uint32_t *off, const uint32_t *nbr, const uint32_t *labels, uint32_t e, double *x,
             double *vde, const LabelTable *cache) {
    if (cache && cache->e == e) {
        bool all = true;
        for (uint32_t v = 0; v < V && all; v++) {
            const double *src = cache->get(labels[v]);
            if (src) std::memcpy(x + (size_t)v * e, src, sizeof(double) * e); else all = false;
        }
        if (!all) cache = nullptr;
    } else {
        cache = nullptr;
    }
    uint32_t max_label = 0;
    for (uint32_t v = 0; v < V; v++) max_label = std::max(max_label, labels[v]);
    // gen_vde_x accepts any 32-bit label (it only seeds the generator with it): a dense table while the label
    // alphabet is small, a hash map for sparse huge label ids (never an allocation proportional to the largest id)
    if (V && !cache && max_label < (1u << 24)) {
        std::vector<double> table(((size_t)max_label + 1) * e);
        std::vector<char> have((size_t)max_label + 1, 0);
        for (uint32_t v = 0; v < V; v++) {
            uint32_t lab = labels[v];
            if (!have[lab]) { label_embedding(lab, e, &table[(size_t)lab * e]); have[lab] = 1; }
            std::memcpy(x + (size_t)v * e, &table[(size_t)lab * e], sizeof(double) * e);
        }
    } else if (V && !cache) {
        std::unordered_map<uint32_t, std::vector<double>> table;
        for (uint32_t v = 0; v < V; v++) {
            auto it = table.find(labels[v]);
            if (it == table.end()) {
                it = table.emplace(labels[v], std::vector<double>(e)).first;
                label_embedding(labels[v], e, it->second.data());
            }
            std::memcpy(x + (size_t)v * e, it->second.data(), sizeof(double) * e);
        }
    }
    for (uint32_t v = 0; v < V; v++) {
        double *out = vde + (size_t)v * e;
        for (uint32_t k = 0; k < e; k++) out[k] = 0.0;
        for (uint32_t j = off[v]; j < off[v + 1]; j++) {  // ascending neighbour id, FP64 adds in that order
            const double *xn = x + (size_t)nbr[j] * e;
            for (uint32_t k = 0; k < e; k++) out[k] += xn[k];
        }
        for (uint32_t k = 0; k < e; k++) out[k] = x[(size_t)v * e + k] + out[k];
    }
}

// A query graph as the C ABI takes it: CSR offsets from 0, adjacency strictly ascending (simple), ids in range, no self
// loops, every edge listed from both ends.  The reference builds its CSR from a file and never checks (graph.cpp:163-242).
bool query_csr_ok(uint32_t nq, const uint32_t *off, const uint32_t *nbr, std::string &why) {
    if (!off || (nq && off[nq] && !nbr)) { why = "null query arrays"; return false; }
    if (off[0] != 0) { why = "query offsets must start at 0"; return false; }
    for (uint32_t u = 0; u < nq; u++) {
        if (off[u + 1] < off[u]) { why = "query offsets not monotone"; return false; }
        if (off[u + 1] - off[u] >= nq) { why = "query degree out of range"; return false; }
    }
    for (uint32_t u = 0; u < nq; u++)
        for (uint32_t j = off[u]; j < off[u + 1]; j++) {
            if (nbr[j] >= nq) { why = "query neighbour id out of range"; return false; }
            if (nbr[j] == u) { why = "self loop in query graph"; return false; }
            if (j > off[u] && nbr[j - 1] >= nbr[j]) { why = "query adjacency must be strictly ascending (simple graph)"; return false; }
        }
    for (uint32_t u = 0; u < nq; u++)
        for (uint32_t j = off[u]; j < off[u + 1]; j++) {
            const uint32_t v = nbr[j];
            if (!std::binary_search(nbr + off[v], nbr + off[v + 1], u)) { why = "query adjacency is not symmetric"; return false; }
        }
    return true;
}

bool query_connected(uint32_t nq, const uint32_t *off, const uint32_t *nbr) {
    if (nq == 0) return true;
    std::vector<char> seen(nq, 0);
    std::vector<uint32_t> stack(1, 0);
    seen[0] = 1;
    uint32_t n = 1;
    while (!stack.empty()) {
        uint32_t v = stack.back();
        stack.pop_back();
        for (uint32_t j = off[v]; j < off[v + 1]; j++)
            if (!seen[nbr[j]]) { seen[nbr[j]] = 1; n++; stack.push_back(nbr[j]); }
    }
    return n == nq;
}

// The query's simple paths of L vertices in dfs_query order (custom.h:94-119, main.cpp:142-146): start
// vertices 0..nq-1, neighbours ascending, a path kept iff its reverse was not kept earlier, i.e. iff
// first < last (the closed form of the hash-set dedup on a simple graph).
static void query_paths(uint32_t nq, const uint32_t *off, const uint32_t *nbr, uint32_t L, std::vector<uint32_t> &rows) {
    uint32_t path[GPE_MAX_QUERY_VERTICES];
    struct Rec {
        static void go(uint32_t len, uint32_t L, uint32_t *path, const uint32_t *off, const uint32_t *nbr,
                       std::vector<uint32_t> &rows) {
            if (len == L) {
                if (path[0] < path[L - 1]) rows.insert(rows.end(), path, path + L);
                return;
            }
            uint32_t last = path[len - 1];
            for (uint32_t j = off[last]; j < off[last + 1]; j++) {
                uint32_t nb = nbr[j];
                bool dup = false;
                for (uint32_t k = 0; k < len; k++) dup = dup || path[k] == nb;
                if (dup) continue;
                path[len] = nb;
                go(len + 1, L, path, off, nbr, rows);
            }
        }
    };
    for (uint32_t s = 0; s < nq; s++) {
        path[0] = s;
        Rec::go(1, L, path, off, nbr, rows);
    }
}

// custom.h:574-633.  std::sort by weight (sum of degrees) descending -- the reference's comparator on the
// reference's initial order, with libstdc++'s std::sort, so ties fall exactly as they do there (SURVEY.md
// Q4) -- then the greedy vertex cover.
void query_plan(uint32_t nq, const uint32_t *off, const uint32_t *nbr, const uint32_t *labels, uint32_t L, uint32_t e,
                QueryPlan &plan, const LabelTable *table) {
    std::vector<uint32_t> rows;
    query_paths(nq, off, nbr, L, rows);
    std::vector<double> x((size_t)nq * e), vde((size_t)nq * e);
    gen_vde(nq, off, nbr, labels, e, x.data(), vde.data(), table);

    struct Item { uint32_t weight, row; };
    size_t n = rows.size() / L;
    std::vector<Item> items(n);
    for (size_t i = 0; i < n; i++) {
        uint32_t w = 0;
        for (uint32_t k = 0; k < L; k++) { uint32_t v = rows[i * L + k]; w += off[v + 1] - off[v]; }
        items[i] = Item{w, (uint32_t)i};
    }
    std::sort(items.begin(), items.end(), [](const Item &a, const Item &b) { return a.weight > b.weight; });

    std::vector<char> covered(nq, 0);
    uint32_t n_covered = 0;
    plan.n = 0;
    plan.L = L;
    plan.e = e;
    plan.vids.clear(); plan.labels.clear(); plan.degs.clear(); plan.pde.clear();
    plan.n_query_paths = (uint32_t)n;
    for (size_t i = 0; i < n; i++) {
        const uint32_t *row = &rows[(size_t)items[i].row * L];
        uint32_t inside = 0;
        for (uint32_t k = 0; k < L; k++) inside += covered[row[k]] ? 1 : 0;
        if (inside != L) {
            for (uint32_t k = 0; k < L; k++) {
                uint32_t v = row[k];
                if (!covered[v]) { covered[v] = 1; n_covered++; }
                plan.vids.push_back(v);
                plan.labels.push_back(labels[v]);
                plan.degs.push_back(off[v + 1] - off[v]);
                for (uint32_t d = 0; d < e; d++) plan.pde.push_back(vde[(size_t)v * e + d]);
            }
            plan.n++;
        }
        if (n_covered == nq) break;
    }
}

static void pge_walk(const uint32_t *off, const uint32_t *nbr, uint32_t pl, uint32_t e, const double *x, const double *vde,
                     uint32_t *path, uint32_t len, double *pg, double *plg, bool &first) {
    if (len == pl) {
        for (uint32_t j = 0; j < pl; j++)
            for (uint32_t k = 0; k < e; k++) {
                const double a = vde[(size_t)path[j] * e + k], b = x[(size_t)path[j] * e + k];
                const uint32_t d = j * e + k;
                if (first) { pg[2 * d] = pg[2 * d + 1] = a; plg[2 * d] = plg[2 * d + 1] = b; }
                else {
                    if (pg[2 * d] > a) pg[2 * d] = a;
                    if (pg[2 * d + 1] < a) pg[2 * d + 1] = a;
                    if (plg[2 * d] > b) plg[2 * d] = b;
                    if (plg[2 * d + 1] < b) plg[2 * d + 1] = b;
                }
            }
        first = false;
        return;
    }
    const uint32_t node = path[len - 1];
    for (uint32_t j = off[node]; j < off[node + 1]; j++) {
        const uint32_t nb = nbr[j];
        bool seen = false;
        for (uint32_t t = 0; t < len; t++) seen = seen || path[t] == nb;
        if (seen) continue;
        path[len] = nb;
        pge_walk(off, nbr, pl, e, x, vde, path, len + 1, pg, plg, first);
    }
}

void pge_groups(uint32_t V, const uint32_t *off, const uint32_t *nbr, uint32_t pl, uint32_t e, const double *x,
                const double *vde, double *pg, double *plg, unsigned char *has) {
    const uint32_t pde = pl * e;
    uint32_t path[GPE_MAX_QUERY_VERTICES];
    for (uint32_t v = 0; v < V; v++) {
        double *a = pg + (size_t)v * 2 * pde, *b = plg + (size_t)v * 2 * pde;
        bool first = true;
        path[0] = v;
        pge_walk(off, nbr, pl, e, x, vde, path, 1, a, b, first);
        has[v] = first ? 0 : 1;
        if (first)
            for (uint32_t d = 0; d < pde; d++) {
                a[2 * d] = a[2 * d + 1] = d < e ? vde[(size_t)v * e + d] : 0.0;
                b[2 * d] = b[2 * d + 1] = d < e ? x[(size_t)v * e + d] : 0.0;
            }
    }
}

}  // namespace gpe

// ---- C ABI ---------------------------------------------------------------------------------------------------
static thread_local std::string g_host_err;
const char *gpe_host_last_error_internal() { return g_host_err.c_str(); }

extern "C" int gpe_host_load_graph(const char *path, uint32_t *V, uint32_t *E, uint32_t *offsets, uint32_t *nbrs,
                                   uint32_t *labels) try {
    if (!path) { g_host_err = "null path"; return GPE_ERR_INVALID; }
    gpe::HostGraph g;
    int rc = gpe::load_graph_file(path, g, g_host_err);
    if (rc) return rc;
    if (V) *V = (uint32_t)g.labels.size();
    if (E) *E = (uint32_t)(g.nbrs.size() / 2);
    if (offsets) std::copy(g.offsets.begin(), g.offsets.end(), offsets);
    if (nbrs) std::copy(g.nbrs.begin(), g.nbrs.end(), nbrs);
    if (labels) std::copy(g.labels.begin(), g.labels.end(), labels);
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    g_host_err = ex.what();
    return GPE_ERR_INVALID;
}

// Memory safety of a caller's CSR before host code walks it (gpe_set_graph checks the rest -- order, symmetry, simple --
// on the device).
static bool data_csr_in_bounds(uint32_t V, const uint32_t *off, const uint32_t *nbr, std::string &why) {
    if (off[0] != 0) { why = "offsets must start at 0"; return false; }
    for (uint32_t v = 0; v < V; v++)
        if (off[v + 1] < off[v]) { why = "offsets not monotone"; return false; }
    if (off[V] && !nbr) { why = "null neighbour array"; return false; }
    for (size_t j = 0; j < off[V]; j++)
        if (nbr[j] >= V) { why = "neighbour id out of range"; return false; }
    return true;
}

extern "C" int gpe_host_gen_vde(uint32_t V, const uint32_t *offsets, const uint32_t *nbrs, const uint32_t *labels,
                                uint32_t e, double *x, double *vde) try {
    if (!offsets || !labels || !x || !vde || e == 0) return GPE_ERR_INVALID;
    if (!data_csr_in_bounds(V, offsets, nbrs, g_host_err)) return GPE_ERR_INVALID;
    gpe::gen_vde(V, offsets, nbrs, labels, e, x, vde);
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    g_host_err = ex.what();
    return GPE_ERR_INVALID;
}

extern "C" int gpe_host_query_plan(uint32_t nq, const uint32_t *q_offsets, const uint32_t *q_nbrs,
                                   const uint32_t *q_labels, uint32_t L, uint32_t e, uint32_t cap, uint32_t *vids,
                                   uint32_t *labels, uint32_t *degs, double *pde, uint32_t *n) try {
    if (nq > GPE_MAX_QUERY_VERTICES || L < 2 || L > GPE_MAX_QUERY_VERTICES || e == 0 || !q_offsets || !q_labels) return GPE_ERR_INVALID;
    if (!gpe::query_csr_ok(nq, q_offsets, q_nbrs, g_host_err)) return GPE_ERR_INVALID;
    gpe::QueryPlan plan;
    gpe::query_plan(nq, q_offsets, q_nbrs, q_labels, L, e, plan);
    uint32_t m = std::min(plan.n, cap);
    if (vids) std::copy(plan.vids.begin(), plan.vids.begin() + (size_t)m * L, vids);
    if (labels) std::copy(plan.labels.begin(), plan.labels.begin() + (size_t)m * L, labels);
    if (degs) std::copy(plan.degs.begin(), plan.degs.begin() + (size_t)m * L, degs);
    if (pde) std::copy(plan.pde.begin(), plan.pde.begin() + (size_t)m * L * e, pde);
    if (n) *n = plan.n;
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    g_host_err = ex.what();
    return GPE_ERR_INVALID;
}

extern "C" int gpe_host_pge_groups(uint32_t V, const uint32_t *offsets, const uint32_t *nbrs, const uint32_t *labels,
                                   uint32_t pl, uint32_t e, double *pg, double *plg, uint8_t *has) try {
    if (!offsets || !labels || !pg || !plg || !has || e == 0 || pl == 0 || pl > GPE_MAX_QUERY_VERTICES) return GPE_ERR_INVALID;
    if (!data_csr_in_bounds(V, offsets, nbrs, g_host_err)) return GPE_ERR_INVALID;
    std::vector<double> x((size_t)V * e), vde((size_t)V * e);
    gpe::gen_vde(V, offsets, nbrs, labels, e, x.data(), vde.data());
    gpe::pge_groups(V, offsets, nbrs, pl, e, x.data(), vde.data(), pg, plg, has);
    return GPE_OK;
} catch (const std::exception &ex) {  // e.g. std::bad_alloc: an error code, never an abort through the C ABI
    g_host_err = ex.what();
    return GPE_ERR_INVALID;
}
