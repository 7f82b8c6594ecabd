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

int load_graph_file(const char *path, HostGraph &g, std::string &err);
void gen_vde(uint32_t V, const uint32_t *off, const uint32_t *nbr, const uint32_t *labels, uint32_t e, double *x,
             double *vde);
bool query_connected(uint32_t nq, const uint32_t *off, const uint32_t *nbr);
void query_plan(uint32_t nq, const uint32_t *off, const uint32_t *nbr, const uint32_t *labels, uint32_t L, uint32_t e,
                QueryPlan &plan);

}  // namespace gpe

const char *gpe_host_last_error_internal();
