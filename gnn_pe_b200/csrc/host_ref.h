// host_ref.h -- host-side mirror of the reference's serial steps (see host_ref.cpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace gpe {

struct HostGraph {
    std::vector<uint32_t> offsets, nbrs, labels;
};

struct QueryPlan {
    uint32_t n = 0, L = 0, e = 0, n_query_paths = 0;
    std::vector<uint32_t> vids, labels, degs;  // n x L
    std::vector<double> pde;                   // n x L*e
};

// Label embeddings (gen_vde_x, custom.h:492-511) depend on (label, e) only: a table filled on first use saves a
// mt19937 seeding per query vertex when thousands of queries are planned.  Thread-safe for concurrent readers once a
// label is present; fill() is called up front for the labels of a batch.
struct LabelTable {
    uint32_t e = 0;
    std::vector<double> x;    // (label) x e
    std::vector<char> have;
    void fill(const uint32_t *labels, size_t n, uint32_t e_);
    const double *get(uint32_t label) const { return label < have.size() && have[label] ? &x[(size_t)label * e] : nullptr; }
};

int load_graph_file(const char *path, HostGraph &g, std::string &err);
void gen_vde(uint32_t V, const uint32_t *off, const uint32_t *nbr, const uint32_t *labels, uint32_t e, double *x,
             double *vde, const LabelTable *table = nullptr);
bool query_csr_ok(uint32_t nq, const uint32_t *off, const uint32_t *nbr, std::string &why);
bool query_connected(uint32_t nq, const uint32_t *off, const uint32_t *nbr);
void query_plan(uint32_t nq, const uint32_t *off, const uint32_t *nbr, const uint32_t *labels, uint32_t L, uint32_t e,
                QueryPlan &plan, const LabelTable *table = nullptr);

// GNN-PGE (GNN-PGE/src/main.cpp:226-291 for the query, :91-176 for the data graph): bounding boxes of the embeddings of all
// simple paths of pl vertices from every vertex; pg, plg: V x pl*e x 2 ([lo, hi] per dimension), has[v] = 0 when v starts
// no such path (the box is then [vde, vde | 0 ...] / [x, x | 0 ...] as on the data side).
void pge_groups(uint32_t V, const uint32_t *off, const uint32_t *nbr, uint32_t pl, uint32_t e, const double *x,
                const double *vde, double *pg, double *plg, unsigned char *has);

}  // namespace gpe

const char *gpe_host_last_error_internal();
