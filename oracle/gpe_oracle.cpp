// gpe_oracle.cpp -- CPU restatement of GNN-PE's three hot paths.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (gnn_pe_b200/, host/) may include,
// link or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs load it, and only as the checker.
//
// Parity status: PINNED.  Every function below is checked (tests/test_oracle_golden.py)
// against outputs of the unmodified reference built from /root/reference by
// oracle/Makefile (oracle/_ref/main, oracle/_ref/probe); the generated vectors are
// committed under tests/golden/ together with tests/golden/make_golden.py.
// The only unpinned piece is the METIS partition assignment (pymetis is absent here);
// path tables, candidate sets and answers do not depend on it.
// The GNN-PGE section (orc_pge_*, citations relative to /root/reference/GNN-PGE/) is pinned the same
// way: oracle/_ref/pge_main, oracle/_ref/pge_probe, tests/golden/make_golden_pge.py,
// tests/test_oracle_pge.py.
//
// All file:line citations are relative to /root/reference/GNN-PE/.
// Flat arrays, no classes from the reference, nothing copied: each routine restates
// the algorithm the cited lines implement.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <set>
#include <string>
#include <unordered_set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint32_t u32;
typedef uint64_t u64;

namespace {

// ---------------------------------------------------------------------------------------
// Graph store.  include/graph/graph.h:51-239, libsrc/graph/graph.cpp:163-242.
// ---------------------------------------------------------------------------------------
struct OGraph {
    u32 V = 0, E = 0, labels_count = 0, max_degree = 0, max_label_freq = 0;
    std::vector<u32> off, nbr, label;
    u32 deg(u32 v) const { return off[v + 1] - off[v]; }
};

// graph.cpp:163-242: "t V E", "v id label degree" in id order, "e u v"; adjacency sorted
// ascending afterwards (:231-233); labels_count = max(#distinct, max label + 1) (:223).
bool load_graph_text(const char *path, OGraph &g) {
    std::ifstream in(path);
    if (!in.is_open()) return false;
    char type;
    in >> type >> g.V >> g.E;
    g.off.assign(g.V + 1, 0);
    g.nbr.assign((size_t)g.E * 2, 0);
    g.label.assign(g.V, 0);
    std::vector<u32> fill(g.V, 0);
    std::vector<u32> freq;
    u32 max_label = 0, distinct = 0;
    while (in >> type) {
        if (type == 'v') {
            u32 id, lab, d;
            in >> id >> lab >> d;
            g.label[id] = lab;
            g.off[id + 1] = g.off[id] + d;
            if (d > g.max_degree) g.max_degree = d;
            if (lab >= freq.size()) freq.resize(lab + 1, 0);
            if (freq[lab] == 0) distinct++;
            freq[lab]++;
            if (lab > max_label) max_label = lab;
        } else if (type == 'e') {
            u32 a, b;
            in >> a >> b;
            g.nbr[g.off[a] + fill[a]++] = b;
            g.nbr[g.off[b] + fill[b]++] = a;
        }
    }
    g.labels_count = std::max(distinct, max_label + 1);
    for (u32 f : freq) g.max_label_freq = std::max(g.max_label_freq, f);
    for (u32 v = 0; v < g.V; v++) std::sort(g.nbr.begin() + g.off[v], g.nbr.begin() + g.off[v + 1]);
    return true;
}

void graph_from_csr(OGraph &g, u32 V, const u32 *off, const u32 *nbr, const u32 *label) {
    g.V = V;
    g.E = off[V] / 2;
    g.off.assign(off, off + V + 1);
    g.nbr.assign(nbr, nbr + off[V]);
    g.label.assign(label, label + V);
    std::vector<u32> freq;
    u32 distinct = 0, max_label = 0;
    for (u32 v = 0; v < V; v++) {
        g.max_degree = std::max(g.max_degree, g.deg(v));
        u32 lab = label[v];
        if (lab >= freq.size()) freq.resize(lab + 1, 0);
        if (freq[lab]++ == 0) distinct++;
        max_label = std::max(max_label, lab);
    }
    g.labels_count = V ? std::max(distinct, max_label + 1) : 0;
    for (u32 f : freq) g.max_label_freq = std::max(g.max_label_freq, f);
}

// graph.h:215-236: binary search for the larger-degree endpoint inside the smaller list.
bool edge_exists(const OGraph &g, u32 u, u32 v) {
    if (g.deg(u) < g.deg(v)) std::swap(u, v);
    int lo = 0, hi = (int)g.deg(v) - 1;
    const u32 *a = g.nbr.data() + g.off[v];
    while (lo <= hi) {
        int mid = lo + ((hi - lo) >> 1);
        if (a[mid] == u) return true;
        if (a[mid] > u) hi = mid - 1; else lo = mid + 1;
    }
    return false;
}

// ---------------------------------------------------------------------------------------
// Label embedding.  custom.h:492-511.  mt19937 and generate_canonical<double,53> are
// written out by hand so the oracle does not lean on libstdc++:
//   r = (g() + g() * 2^32) / 2^64, clamped below 1; then divide by the left-to-right sum.
// ---------------------------------------------------------------------------------------
struct MT19937 {
    u32 mt[624];
    int idx;
    explicit MT19937(u32 seed) {
        mt[0] = seed;
        for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (u32)i;
        idx = 624;
    }
    u32 next() {
        if (idx >= 624) {
            for (int i = 0; i < 624; i++) {
                u32 y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        u32 y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
};

void label_embedding(u32 label, u32 e, double *out) {
    MT19937 gen(label);
    for (u32 i = 0; i < e; i++) {
        double lo = (double)gen.next();
        double hi = (double)gen.next();
        double r = (lo + hi * 4294967296.0) / 18446744073709551616.0;
        if (r >= 1.0) r = std::nextafter(1.0, 0.0);
        out[i] = r;
    }
    double sum = 0.0;  // std::accumulate(.., 0.0), custom.h:504
    for (u32 i = 0; i < e; i++) sum += out[i];
    for (u32 i = 0; i < e; i++) out[i] = out[i] / sum;
}

// custom.h:513-544: x[v] = label embedding; nx[v] = sum over neighbours in adjacency
// (ascending id) order, starting from 0.0; vde = x + nx.
void vertex_embeddings(const OGraph &g, u32 e, double *x, double *vde) {
    std::vector<double> per_label((size_t)g.labels_count * e);
    std::vector<char> have(g.labels_count, 0);
    for (u32 v = 0; v < g.V; v++) {
        u32 lab = g.label[v];
        if (!have[lab]) { label_embedding(lab, e, &per_label[(size_t)lab * e]); have[lab] = 1; }
        for (u32 k = 0; k < e; k++) x[(size_t)v * e + k] = per_label[(size_t)lab * e + k];
    }
    for (u32 v = 0; v < g.V; v++) {
        for (u32 k = 0; k < e; k++) {
            double nx = 0.0;
            for (u32 j = g.off[v]; j < g.off[v + 1]; j++) nx += x[(size_t)g.nbr[j] * e + k];
            vde[(size_t)v * e + k] = x[(size_t)v * e + k] + nx;
        }
    }
}

// gnnpe.py:71-72: vertices in stable ascending-degree order (python sorted() is stable,
// dict order is node order = id order for the shipped pickle).
void degree_order(const OGraph &g, u32 *sorted_nodes) {
    std::vector<u32> ord(g.V);
    for (u32 v = 0; v < g.V; v++) ord[v] = v;
    std::stable_sort(ord.begin(), ord.end(), [&](u32 a, u32 b) { return g.deg(a) < g.deg(b); });
    std::copy(ord.begin(), ord.end(), sorted_nodes);
}

// ---------------------------------------------------------------------------------------
// Path enumeration, literal: custom.h:52-92 + main.cpp:92-96.  A DFS from every start
// vertex in membership.txt order, an unordered_set of whole paths, keep a path iff neither
// it nor its reverse was seen.  `start_depth` is main.cpp:95's `path_length - 2` (so the
// unmodified reference always walks L=3) or 1 for the patched l!=2 oracle.
// ---------------------------------------------------------------------------------------
struct VecHash {
    size_t operator()(const std::vector<u32> &p) const {
        size_t h = 0;
        for (u32 v : p) h ^= std::hash<u32>()(v) + 0x9e3779b9 + (h << 6) + (h >> 2);
        return h;
    }
};

void enumerate_literal(const OGraph &g, u32 L, const u32 *sorted_nodes, std::vector<u32> &rows) {
    std::unordered_set<std::vector<u32>, VecHash> seen;
    std::vector<u32> path;
    std::function<void(u32)> walk = [&](u32 node) {
        if (path.size() == L) {
            if (seen.count(path)) return;
            std::vector<u32> rev(path.rbegin(), path.rend());
            if (seen.count(rev)) return;
            rows.insert(rows.end(), path.begin(), path.end());
            seen.insert(path);
            return;
        }
        for (u32 j = g.off[node]; j < g.off[node + 1]; j++) {
            u32 nb = g.nbr[j];
            if (std::find(path.begin(), path.end(), nb) != path.end()) continue;
            path.push_back(nb);
            walk(nb);
            path.pop_back();
        }
    };
    for (u32 i = 0; i < g.V; i++) {
        path.assign(1, sorted_nodes[i]);
        walk(sorted_nodes[i]);
    }
}

// Closed form of the same table: simple graph assumed.  Emit the
// simple path (v0..v_{L-1}) found from start v0 iff rank[v0] < rank[v_{L-1}], rank being
// the position in membership.txt.  Content and order equal enumerate_literal's.
void enumerate_closed(const OGraph &g, u32 L, const u32 *sorted_nodes, std::vector<u32> &rows,
                      std::vector<u64> *rows_before_start /* V+1, by rank */) {
    std::vector<u32> rank(g.V);
    for (u32 i = 0; i < g.V; i++) rank[sorted_nodes[i]] = i;
    u32 path[16];
    u64 n = 0;
    std::function<void(u32)> walk = [&](u32 len) {
        if (len == L) {
            if (rank[path[0]] < rank[path[L - 1]]) { rows.insert(rows.end(), path, path + L); n++; }
            return;
        }
        u32 node = path[len - 1];
        for (u32 j = g.off[node]; j < g.off[node + 1]; j++) {
            u32 nb = g.nbr[j];
            bool dup = false;
            for (u32 k = 0; k < len; k++) dup |= (path[k] == nb);
            if (dup) continue;
            path[len] = nb;
            walk(len + 1);
        }
    };
    if (rows_before_start) rows_before_start->assign(g.V + 1, 0);
    for (u32 i = 0; i < g.V; i++) {
        if (rows_before_start) (*rows_before_start)[i] = n;
        path[0] = sorted_nodes[i];
        walk(1);
    }
    if (rows_before_start) (*rows_before_start)[g.V] = n;
}

// ---------------------------------------------------------------------------------------
// Query plan.  custom.h:94-119 (dfs_query from vertices 0..nq-1), :574-633 (gen_query_pde).
// ---------------------------------------------------------------------------------------
struct QPath {
    std::vector<u32> vids, labels, degrees;
    std::vector<double> pde;
    double key;
    u32 weight;
};

void query_plan(const OGraph &q, u32 L, u32 e, bool literal_paths, std::vector<QPath> &plan,
                std::vector<u32> *all_query_paths) {
    std::vector<u32> ident(q.V);
    for (u32 i = 0; i < q.V; i++) ident[i] = i;
    std::vector<u32> rows;
    if (literal_paths) enumerate_literal(q, L, ident.data(), rows);
    else enumerate_closed(q, L, ident.data(), rows, nullptr);
    if (all_query_paths) *all_query_paths = rows;
    std::vector<double> x((size_t)q.V * e), vde((size_t)q.V * e);
    vertex_embeddings(q, e, x.data(), vde.data());

    size_t n = rows.size() / L;
    std::vector<QPath> qp(n);
    for (size_t i = 0; i < n; i++) {
        qp[i].weight = 0;
        for (u32 j = 0; j < L; j++) {
            u32 v = rows[i * L + j];
            qp[i].vids.push_back(v);
            qp[i].labels.push_back(q.label[v]);
            qp[i].degrees.push_back(q.deg(v));
            qp[i].weight += q.deg(v);
            for (u32 k = 0; k < e; k++) qp[i].pde.push_back(vde[(size_t)v * e + k]);
        }
        qp[i].key = 0;
        for (double d : qp[i].pde) qp[i].key -= d;
    }
    // custom.h:601-605: std::sort, weight descending, unstable above 16 elements (Q4).
    // The permutation only depends on comparator outcomes, so sorting these records with
    // libstdc++'s std::sort reproduces the reference's order.
    std::sort(qp.begin(), qp.end(), [](const QPath &a, const QPath &b) { return a.weight > b.weight; });
    // custom.h:607-628: greedy cover.
    std::set<u32> covered;
    plan.clear();
    for (size_t i = 0; i < qp.size(); i++) {
        u32 inside = 0;
        for (u32 v : qp[i].vids) inside += covered.count(v) ? 1 : 0;
        if (inside != L) {
            covered.insert(qp[i].vids.begin(), qp[i].vids.end());
            plan.push_back(qp[i]);
        }
        if (covered.size() == q.V) break;
    }
}

// ---------------------------------------------------------------------------------------
// Filter.  The leaf compare of custom.h:407-435 applied to every (plan path, data path)
// pair; SURVEY.md F2 shows the index traversal returns exactly this.
// ---------------------------------------------------------------------------------------
const double kEps = 1e-6;  // custom.h:43

inline bool row_accepts(const OGraph &g, const double *vde, u32 e, u32 L, const QPath &q, const u32 *row) {
    for (u32 k = 0; k < L; k++) {
        u32 v = row[k];
        if (q.labels[k] != g.label[v] || q.degrees[k] > g.deg(v)) return false;   // :412
    }
    for (u32 k = 0; k < L; k++) {
        const double *pv = vde + (size_t)row[k] * e;
        for (u32 d = 0; d < e; d++) {
            double qd = q.pde[k * e + d], pd = pv[d];
            if (qd > pd && std::fabs(qd - pd) > kEps) return false;               // :422
        }
    }
    return true;
}

void filter_rows(const OGraph &g, const double *vde, u32 e, u32 L, const std::vector<QPath> &plan,
                 const u32 *rows, u64 n_rows, std::vector<std::set<u32>> &cand, u64 *survivors_per_qpath) {
    for (u64 r = 0; r < n_rows; r++) {
        const u32 *row = rows + r * L;
        for (size_t j = 0; j < plan.size(); j++) {
            if (!row_accepts(g, vde, e, L, plan[j], row)) continue;
            if (survivors_per_qpath) survivors_per_qpath[j]++;
            for (u32 k = 0; k < L; k++) cand[plan[j].vids[k]].insert(row[k]);      // :429-432
        }
    }
}

// ---------------------------------------------------------------------------------------
// Refinement.  custom.h:635-722 (order + pivots), :724-755 (backward neighbours),
// :757-888 (enumeration), :890-932 (driver).
// ---------------------------------------------------------------------------------------
bool query_connected(const OGraph &q) {
    if (q.V == 0) return true;
    std::vector<char> seen(q.V, 0);
    std::vector<u32> st(1, 0);
    seen[0] = 1;
    u32 n = 1;
    while (!st.empty()) {
        u32 v = st.back(); st.pop_back();
        for (u32 j = q.off[v]; j < q.off[v + 1]; j++)
            if (!seen[q.nbr[j]]) { seen[q.nbr[j]] = 1; n++; st.push_back(q.nbr[j]); }
    }
    return n == q.V;
}

void matching_order(const OGraph &g, const OGraph &q, const u32 *count, u32 *order, u32 *pivot) {
    u32 nq = q.V;
    std::vector<char> visited(nq, 0), adjacent(nq, 0);
    u32 start = 0;                                                   // :635-654
    for (u32 i = 1; i < nq; i++) {
        if (count[i] < count[start]) start = i;
        else if (count[i] == count[start] && q.deg(i) > q.deg(start)) start = i;
    }
    auto mark = [&](u32 u) {                                         // :656-668
        visited[u] = 1;
        for (u32 j = q.off[u]; j < q.off[u + 1]; j++) adjacent[q.nbr[j]] = 1;
    };
    order[0] = start;
    mark(start);
    for (u32 i = 1; i < nq; i++) {                                   // :682-705
        u32 next = 0, best = g.V + 1;
        for (u32 u = 0; u < nq; u++) {
            if (visited[u] || !adjacent[u]) continue;
            if (count[u] < best) { best = count[u]; next = u; }
            else if (count[u] == best && q.deg(u) > q.deg(next)) next = u;
        }
        mark(next);
        order[i] = next;
    }
    pivot[0] = 0xffffffffu;  // uninitialised in the reference, never read
    for (u32 i = 1; i < nq; i++)                                     // :707-719
        for (u32 j = 0; j < i; j++)
            if (edge_exists(q, order[i], order[j])) { pivot[i] = order[j]; break; }
}

// Returns the number of embeddings whose image of order[0] lies in start_cands, stopping
// once `limit` is reached (custom.h:846-855).  If `dump` is non-null every embedding is
// appended to it as nq vertex ids indexed by query vertex.
u64 enumerate_matches(const OGraph &g, const OGraph &q, const u32 *order, const u32 *pivot,
                      const u32 *start_cands, u32 n_start, u64 limit, std::vector<u32> *dump) {
    u32 nq = q.V;
    if (nq == 0) return 0;
    std::vector<std::vector<u32>> bn(nq);                            // :724-755
    {
        std::vector<char> seen(nq, 0);
        seen[order[0]] = 1;
        for (u32 i = 1; i < nq; i++) {
            u32 u = order[i];
            for (u32 j = q.off[u]; j < q.off[u + 1]; j++) {
                u32 w = q.nbr[j];
                if (seen[w] && w != pivot[i]) bn[i].push_back(w);
            }
            seen[u] = 1;
        }
    }
    std::vector<u32> emb(nq, 0), idx(nq, 0);
    std::vector<std::vector<u32>> valid(nq);
    std::vector<char> used(g.V, 0);
    valid[0].assign(start_cands, start_cands + n_start);
    u64 found = 0;
    int depth = 0;
    while (true) {                                                   // :836-870
        while (idx[depth] < valid[depth].size()) {
            u32 u = order[depth], v = valid[depth][idx[depth]++];
            emb[u] = v;
            used[v] = 1;
            if (depth == (int)nq - 1) {
                found++;
                if (dump) dump->insert(dump->end(), emb.begin(), emb.end());
                used[v] = 0;
                if (found >= limit) return found;
            } else {
                depth++;
                idx[depth] = 0;
                valid[depth].clear();                                // :757-797
                u32 w = order[depth], p = emb[pivot[depth]];
                for (u32 j = g.off[p]; j < g.off[p + 1]; j++) {
                    u32 c = g.nbr[j];
                    if (used[c] || g.label[c] != q.label[w] || q.deg(w) > g.deg(c)) continue;
                    bool ok = true;
                    for (u32 b : bn[depth]) if (!edge_exists(g, c, emb[b])) { ok = false; break; }
                    if (ok) valid[depth].push_back(c);
                }
            }
        }
        depth--;
        if (depth < 0) break;
        used[emb[order[depth]]] = 0;
    }
    return found;
}

// ---------------------------------------------------------------------------------------
// GNN-PGE (the reference's sibling variant; citations relative to the
// reference's GNN-PGE/ directory).  Per-VERTEX filter: every vertex carries the bounding box of the
// embeddings of all simple paths of `pl` vertices that start at it ("path group", src/main.cpp:91-176
// for the data graph, :226-291 for the query), over the dominance embeddings (pg) and over the label
// embeddings (plg); both as [lo0, hi0, lo1, hi1, ...] with pde = e * pl dimensions.
// ---------------------------------------------------------------------------------------
static void pge_walk(const OGraph &g, u32 pl, u32 e, const double *x, const double *vde, u32 *path, u32 len,
                     double *pg, double *plg, bool &first) {
    if (len == pl) {  // include/custom.h:52-57 + the min/max folds of main.cpp:145-176
        for (u32 j = 0; j < pl; j++)
            for (u32 k = 0; k < e; k++) {
                const double a = vde[(size_t)path[j] * e + k], b = x[(size_t)path[j] * e + k];
                const u32 d = j * e + k;
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
    const u32 node = path[len - 1];
    for (u32 j = g.off[node]; j < g.off[node + 1]; j++) {  // include/custom.h:59-70: simple paths only
        const u32 nb = g.nbr[j];
        bool seen = false;
        for (u32 t = 0; t < len; t++) seen = seen || path[t] == nb;
        if (seen) continue;
        path[len] = nb;
        pge_walk(g, pl, e, x, vde, path, len + 1, pg, plg, first);
    }
}

// has[v] = 0 for a vertex without any such path: the data side then stores [vde, vde | 0 ...] and
// [x, x | 0 ...] (main.cpp:103-121); the query side leaves the group empty (main.cpp:249-252)
static void pge_groups(const OGraph &g, u32 pl, u32 e, const double *x, const double *vde, double *pg, double *plg,
                       unsigned char *has) {
    const u32 pde = pl * e;
#pragma omp parallel for schedule(dynamic, 64)
    for (long long vv = 0; vv < (long long)g.V; vv++) {
        const u32 v = (u32)vv;
        u32 path[64];
        double *a = pg + (size_t)v * 2 * pde, *b = plg + (size_t)v * 2 * pde;
        bool first = true;
        path[0] = v;
        pge_walk(g, pl, e, x, vde, path, 1, a, b, first);
        has[v] = first ? 0 : 1;
        if (first) {
            for (u32 d = 0; d < pde; d++) {
                const double va = d < e ? vde[(size_t)v * e + d] : 0.0, vb = d < e ? x[(size_t)v * e + d] : 0.0;
                a[2 * d] = a[2 * d + 1] = va;
                b[2 * d] = b[2 * d + 1] = vb;
            }
        }
    }
}

// The leaf test of Partition::query, include/custom.h:332-367: label equal, query degree <= data degree,
// (the per-vertex embedding loop at :339-346 is dead: it starts at k = vde_dim), the label boxes overlap in
// every dimension, and the data box's upper corner is not below the query box's lower corner in any
// dimension (no epsilon here; the epsilon of :395 only routes inside the R*-tree).
static inline bool pge_leaf_test(u32 pde, u32 qlabel, u32 qdeg, const double *qpg, const double *qplg, u32 vlabel,
                                 u32 vdeg, const double *vpg, const double *vplg) {
    if (!(qdeg <= vdeg && qlabel == vlabel)) return false;
    for (u32 k = 0; k < pde; k++)
        if (vplg[2 * k + 1] < qplg[2 * k] || vplg[2 * k] > qplg[2 * k + 1]) return false;
    for (u32 k = 0; k < pde; k++)
        if (vpg[2 * k + 1] < qpg[2 * k]) return false;
    return true;
}

struct Handle {
    OGraph g;
    std::vector<u32> rows;       // enumerated path table, row-major n x L
    u32 L = 0;
    std::vector<u64> rows_before_start;
};

}  // namespace

// =========================================================================================
// C interface for ctypes (tests/, bench.py cpu_baseline).  Every call returns 0 on success.
// =========================================================================================
extern "C" {

void *orc_graph_load(const char *path) {
    Handle *h = new Handle();
    if (!load_graph_text(path, h->g)) { delete h; return nullptr; }
    return h;
}

void *orc_graph_from_csr(u32 V, const u32 *off, const u32 *nbr, const u32 *label) {
    Handle *h = new Handle();
    graph_from_csr(h->g, V, off, nbr, label);
    return h;
}

void orc_graph_free(void *h) { delete (Handle *)h; }

void orc_graph_meta(void *h, u32 *out5) {
    OGraph &g = ((Handle *)h)->g;
    out5[0] = g.V; out5[1] = g.E; out5[2] = g.labels_count; out5[3] = g.max_degree; out5[4] = g.max_label_freq;
}

void orc_graph_csr(void *h, u32 *off, u32 *nbr, u32 *label) {
    OGraph &g = ((Handle *)h)->g;
    std::copy(g.off.begin(), g.off.end(), off);
    std::copy(g.nbr.begin(), g.nbr.end(), nbr);
    std::copy(g.label.begin(), g.label.end(), label);
}

void orc_label_embedding(u32 label, u32 e, double *out) { label_embedding(label, e, out); }

void orc_vertex_embeddings(void *h, u32 e, double *x, double *vde) {
    vertex_embeddings(((Handle *)h)->g, e, x, vde);
}

void orc_degree_order(void *h, u32 *sorted_nodes) { degree_order(((Handle *)h)->g, sorted_nodes); }

// mode 0 = closed form, 1 = literal DFS + hash set.  Returns row count; table kept in the handle.
u64 orc_enumerate(void *hv, u32 L, const u32 *sorted_nodes, int literal) {
    Handle *h = (Handle *)hv;
    h->rows.clear();
    h->L = L;
    if (literal) enumerate_literal(h->g, L, sorted_nodes, h->rows);
    else enumerate_closed(h->g, L, sorted_nodes, h->rows, &h->rows_before_start);
    return h->rows.size() / L;
}

void orc_paths_copy(void *hv, u64 first, u64 n, u32 *out) {
    Handle *h = (Handle *)hv;
    std::copy(h->rows.begin() + first * h->L, h->rows.begin() + (first + n) * h->L, out);
}

// rows per partition: a path belongs to the partition of its FIRST vertex (custom.h:74, T6).
void orc_rows_per_partition(void *hv, const u32 *membership, u32 p, u64 *out) {
    Handle *h = (Handle *)hv;
    for (u32 i = 0; i < p; i++) out[i] = 0;
    for (u64 r = 0; r < h->rows.size() / h->L; r++) out[membership[h->rows[r * h->L]]]++;
}

// Query plan of query graph `qv`: writes up to cap plan paths; returns the plan size.
// vids/labels/degrees: n x L, pde: n x L*e, weight: n.
u32 orc_query_plan(void *qv, u32 L, u32 e, int literal, u32 cap, u32 *vids, u32 *labels, u32 *degrees,
                   double *pde, u32 *weight, u32 *n_query_paths) {
    std::vector<QPath> plan;
    std::vector<u32> all;
    query_plan(((Handle *)qv)->g, L, e, literal != 0, plan, &all);
    if (n_query_paths) *n_query_paths = (u32)(all.size() / L);
    for (u32 i = 0; i < plan.size() && i < cap; i++) {
        for (u32 k = 0; k < L; k++) {
            vids[i * L + k] = plan[i].vids[k];
            labels[i * L + k] = plan[i].labels[k];
            degrees[i * L + k] = plan[i].degrees[k];
        }
        for (u32 k = 0; k < L * e; k++) pde[(size_t)i * L * e + k] = plan[i].pde[k];
        weight[i] = plan[i].weight;
    }
    return (u32)plan.size();
}

// Brute-force filter of the handle's enumerated table (all rows) for query `qv`.
// literal_plan: bit 0 = the literal dfs enumeration of the query paths, bit 1 = both orientations (exact mode).
// cand_off: nq+1 offsets into cand (sorted ascending per query vertex); returns total size
// (call with cand == NULL first to size the buffer).  survivors: per plan path, may be NULL.
u64 orc_filter(void *gv, void *qv, u32 e, int literal_plan, u64 *cand_off, u32 *cand, u64 cand_cap,
               u64 *survivors, u32 survivors_cap) {
    Handle *h = (Handle *)gv;
    OGraph &q = ((Handle *)qv)->g;
    u32 L = h->L;
    std::vector<QPath> plan;
    query_plan(q, L, e, (literal_plan & 1) != 0, plan, nullptr);
    const size_t n_plan = plan.size();
    if (literal_plan & 2) {
        // Exact mode (not the reference's behaviour): the reference stores one orientation of every data
        // path (custom.h:68-79) and compares plan paths with that one only (:407-435).  Comparing the reversed plan path
        // with the stored row is comparing the plan path with the row's other orientation.
        for (size_t j = 0; j < n_plan; j++) {
            QPath r = plan[j];
            std::reverse(r.vids.begin(), r.vids.end());
            std::reverse(r.labels.begin(), r.labels.end());
            std::reverse(r.degrees.begin(), r.degrees.end());
            for (u32 k = 0; k < L; k++)
                for (u32 d = 0; d < e; d++) r.pde[k * e + d] = plan[j].pde[(L - 1 - k) * e + d];
            plan.push_back(r);
        }
    }
    std::vector<double> x((size_t)h->g.V * e), vde((size_t)h->g.V * e);
    vertex_embeddings(h->g, e, x.data(), vde.data());
    std::vector<std::set<u32>> cs(q.V);
    std::vector<u64> surv(plan.size(), 0);
    filter_rows(h->g, vde.data(), e, L, plan, h->rows.data(), h->rows.size() / L, cs, surv.data());
    if (literal_plan & 2) {  // survivors of both orientations count for the plan path
        for (size_t j = 0; j < n_plan; j++) surv[j] += surv[n_plan + j];
        surv.resize(n_plan);
        plan.resize(n_plan);
    }
    u64 total = 0;
    for (u32 u = 0; u < q.V; u++) {
        cand_off[u] = total;
        for (u32 v : cs[u]) { if (cand && total < cand_cap) cand[total] = v; total++; }
    }
    cand_off[q.V] = total;
    if (survivors) for (u32 j = 0; j < plan.size() && j < survivors_cap; j++) survivors[j] = surv[j];
    return total;
}

int orc_query_connected(void *qv) { return query_connected(((Handle *)qv)->g) ? 1 : 0; }

void orc_matching_order(void *gv, void *qv, const u32 *cand_count, u32 *order, u32 *pivot) {
    matching_order(((Handle *)gv)->g, ((Handle *)qv)->g, cand_count, order, pivot);
}

// custom.h:890-932 given candidate sets (CSR form, sorted).  limit: main.cpp:62-69 (UINT_MAX for MAX).
// matches (optional): up to matches_cap embeddings of nq ids each, in the reference's DFS order.
u64 orc_refine(void *gv, void *qv, const u64 *cand_off, const u32 *cand, u64 limit, u32 *matches,
               u64 matches_cap) {
    OGraph &g = ((Handle *)gv)->g;
    OGraph &q = ((Handle *)qv)->g;
    u32 nq = q.V;
    std::vector<u32> cnt(nq), order(nq), pivot(nq);
    for (u32 u = 0; u < nq; u++) cnt[u] = (u32)(cand_off[u + 1] - cand_off[u]);
    matching_order(g, q, cnt.data(), order.data(), pivot.data());
    std::vector<u32> dump;
    u32 s = order[0];
    u64 n = enumerate_matches(g, q, order.data(), pivot.data(), cand + cand_off[s], cnt[s], limit,
                              matches ? &dump : nullptr);
    if (matches) std::copy(dump.begin(), dump.begin() + std::min<u64>(dump.size(), matches_cap * nq), matches);
    return n;
}

// Whole online stage for one query against the enumerated table (main.cpp:122-179).
u64 orc_online(void *gv, void *qv, u32 e, u64 limit) {
    Handle *h = (Handle *)gv;
    OGraph &q = ((Handle *)qv)->g;
    std::vector<u64> off(q.V + 1);
    u64 total = orc_filter(gv, qv, e, 0, off.data(), nullptr, 0, nullptr, 0);
    std::vector<u32> cand(total ? total : 1);
    orc_filter(gv, qv, e, 0, off.data(), cand.data(), total, nullptr, 0);
    (void)h;
    return orc_refine(gv, qv, off.data(), cand.data(), limit, nullptr, 0);
}

// ---------------------------------------------------------------------------------------
// CPU baseline at scale ("port"): the same all-pairs compare,
// but data paths are walked on the fly from the CSR instead of a materialised table, with
// OpenMP over start vertices.  Returns the answer; timings (seconds) in t3 = {plan, filter, refine}.
// Supports L = 3 and L = 4.
// ---------------------------------------------------------------------------------------
u64 orc_online_streaming(void *gv, void *qv, u32 L, u32 e, const u32 *sorted_nodes, const double *vde,
                         u64 limit, double *t3, int threads) {
    Handle *h = (Handle *)gv;
    OGraph &g = h->g;
    OGraph &q = ((Handle *)qv)->g;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
    double t0 = omp_get_wtime();
#else
    double t0 = 0;
#endif
    std::vector<QPath> plan;
    query_plan(q, L, e, false, plan, nullptr);
    std::vector<u32> rank(g.V);
    for (u32 i = 0; i < g.V; i++) rank[sorted_nodes[i]] = i;
#ifdef _OPENMP
    double t1 = omp_get_wtime();
#else
    double t1 = 0;
#endif
    u32 nq = q.V;
    std::vector<std::vector<u32>> cand(nq);
#pragma omp parallel
    {
        std::vector<std::vector<u32>> local(nq);
        u32 row[4];
#pragma omp for schedule(dynamic, 256)
        for (u32 a = 0; a < g.V; a++) {
            row[0] = a;
            for (u32 i = g.off[a]; i < g.off[a + 1]; i++) {
                u32 b = g.nbr[i];
                row[1] = b;
                for (u32 j = g.off[b]; j < g.off[b + 1]; j++) {
                    u32 c = g.nbr[j];
                    if (c == a) continue;
                    row[2] = c;
                    if (L == 3) {
                        if (rank[a] >= rank[c]) continue;
                        for (size_t t = 0; t < plan.size(); t++)
                            if (row_accepts(g, vde, e, L, plan[t], row))
                                for (u32 k = 0; k < L; k++) local[plan[t].vids[k]].push_back(row[k]);
                    } else {
                        for (u32 m = g.off[c]; m < g.off[c + 1]; m++) {
                            u32 d = g.nbr[m];
                            if (d == a || d == b || rank[a] >= rank[d]) continue;
                            row[3] = d;
                            for (size_t t = 0; t < plan.size(); t++)
                                if (row_accepts(g, vde, e, L, plan[t], row))
                                    for (u32 k = 0; k < L; k++) local[plan[t].vids[k]].push_back(row[k]);
                        }
                    }
                }
            }
        }
#pragma omp critical
        for (u32 u = 0; u < nq; u++) cand[u].insert(cand[u].end(), local[u].begin(), local[u].end());
    }
    std::vector<u64> off(nq + 1, 0);
    std::vector<u32> flat;
    for (u32 u = 0; u < nq; u++) {
        std::sort(cand[u].begin(), cand[u].end());
        cand[u].erase(std::unique(cand[u].begin(), cand[u].end()), cand[u].end());
        off[u] = flat.size();
        flat.insert(flat.end(), cand[u].begin(), cand[u].end());
    }
    off[nq] = flat.size();
    if (flat.empty()) flat.push_back(0);
#ifdef _OPENMP
    double t2 = omp_get_wtime();
#else
    double t2 = 0;
#endif
    u64 n = orc_refine(gv, qv, off.data(), flat.data(), limit, nullptr, 0);
#ifdef _OPENMP
    double t3e = omp_get_wtime();
#else
    double t3e = 0;
#endif
    if (t3) { t3[0] = t1 - t0; t3[1] = t2 - t1; t3[2] = t3e - t2; }
    return n;
}


// ---- GNN-PGE ------------------------------------------------------------------------------------------
// pg, plg: V x 2*pl*e; has: V
void orc_pge_groups(void *hv, u32 pl, u32 e, double *pg, double *plg, unsigned char *has) {
    OGraph &g = ((Handle *)hv)->g;
    std::vector<double> x((size_t)g.V * e), vde((size_t)g.V * e);
    vertex_embeddings(g, e, x.data(), vde.data());
    pge_groups(g, pl, e, x.data(), vde.data(), pg, plg, has);
}

// Candidate sets of every query vertex (ascending ids, like the std::set merge of main.cpp:331-338): all data
// vertices that pass the leaf test.  Two-call pattern: cand == NULL returns the sizes in cand_off only.
u64 orc_pge_filter(void *gv, void *qv, u32 pl, u32 e, u64 *cand_off, u32 *cand, u64 cand_cap) {
    OGraph &g = ((Handle *)gv)->g;
    OGraph &q = ((Handle *)qv)->g;
    const u32 pde = pl * e;
    std::vector<double> gpg((size_t)g.V * 2 * pde), gplg((size_t)g.V * 2 * pde), qpg((size_t)q.V * 2 * pde),
        qplg((size_t)q.V * 2 * pde);
    std::vector<unsigned char> ghas(g.V), qhas(q.V);
    orc_pge_groups(gv, pl, e, gpg.data(), gplg.data(), ghas.data());
    orc_pge_groups(qv, pl, e, qpg.data(), qplg.data(), qhas.data());
    u64 total = 0;
    for (u32 u = 0; u < q.V; u++) {
        cand_off[u] = total;
        if (!qhas[u]) continue;  // (the reference would read an empty vector here: undefined; connected queries of
                                 //  at least pl vertices per path always have a group)
        for (u32 v = 0; v < g.V; v++)
            if (pge_leaf_test(pde, q.label[u], q.deg(u), &qpg[(size_t)u * 2 * pde], &qplg[(size_t)u * 2 * pde], g.label[v],
                              g.deg(v), &gpg[(size_t)v * 2 * pde], &gplg[(size_t)v * 2 * pde])) {
                if (cand && total < cand_cap) cand[total] = v;
                total++;
            }
    }
    cand_off[q.V] = total;
    return total;
}

// filter + the (shared) refinement: the number the reference prints as "Answer Num" (main.cpp:343-346)
u64 orc_pge_online(void *gv, void *qv, u32 pl, u32 e, u64 limit) {
    OGraph &q = ((Handle *)qv)->g;
    std::vector<u64> off(q.V + 1);
    const u64 total = orc_pge_filter(gv, qv, pl, e, off.data(), nullptr, 0);
    std::vector<u32> cand(total ? total : 1);
    orc_pge_filter(gv, qv, pl, e, off.data(), cand.data(), total);
    return orc_refine(gv, qv, off.data(), cand.data(), limit, nullptr, 0);
}

}  // extern "C"
