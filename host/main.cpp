// main.cpp -- the reference's command line (GNN-PE/src/main.cpp:38-184) on top of libgpe.so.
//
//   main -f <dataset dir/> -d <data.graph> -q <query.graph> -m offline|online -p <partitions>
//        -l <path length> -e <embedding dim> -n <MAX|answers>
//
// Same flags (short and long forms, main.cpp:46-54), same defaults (main.cpp:40-44, custom.h:47-50), same
// files (gnn-pe/membership.txt in, gnn-pe/all_paths.txt and partitions/partition-i/partition_paths.txt
// out) and the same stdout lines (graph.cpp:244-247, custom.h:630, main.cpp:179).  Everything
// data-parallel goes through the C ABI in include/gpe.h; there is no CPU fallback.
//
// Differences, all deliberate:
//   * online mode does not parse all_paths.txt / build index.dat: it re-enumerates on the GPU from the
//     graph and membership.txt (milliseconds) and scans instead of traversing an R*-tree;
//   * the offline -> online handoff is a small versioned binary manifest (gnn-pe/paths.gpe: hashes of the graph and of
//     membership.txt, l, p, row counts) next to the reference's text files.  Online mode checks it and refuses to run
//     against outputs of another graph / membership / -l / -p -- the reference reuses a stale index.dat silently
//     (custom.h:218-258);
//   * missing input files are errors (the reference reads zeros silently, main.cpp:80-85);
//   * -l other than 2 follows the patched-oracle semantics of SURVEY.md F5 (only l=2 and l=3 are built);
//   * -g N (--gpus N) shards the online stage over N GPUs of this machine, one process: GPU r holds the path table of
//     the partitions i % N == r (needs -p >= N), the candidate sets are exchanged with NCCL inside libgpe and the join
//     is split by start candidate (gpe_comm_init_all / gpe_multi_query_batch).  Answers are those of -g 1;
//   * --filter pge (-F pge) runs the sibling variant GNN-PGE instead (GNN-PGE/src/main.cpp:38-363: same flags, -l is
//     the number of VERTICES per path there, files under gnn-pge/): offline writes gnn-pge/data_vertices.bin byte for
//     byte as the reference does (:179-194), online prints its `Answer Num:` line (:361).  One GPU;
//   * -q may name a DIRECTORY: every *.graph file in it (sorted by name) is answered in one batch -- one
//     `<file>: Answer Number: N` line per query, then the batch's total time and queries/s (BASELINE.json config 5).
#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <sys/stat.h>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "gpe.h"

namespace {

struct Options {
    std::string file = "../Test/", data = "../Test/data_graph.graph", query = "../Test/query_graph.graph";
    std::string mode = "offline", answers = "MAX", filter = "pe";
    uint32_t partitions = 5, length = 2, embedding = 2, gpus = 1;
};

bool parse(int argc, char **argv, Options &o) {
    struct Opt { const char *s, *l; int id; };
    static const Opt opts[] = {{"-f", "--file", 0}, {"-d", "--data", 1}, {"-q", "--query", 2}, {"-m", "--mode", 3},
                               {"-p", "--partition", 4}, {"-l", "--length", 5}, {"-e", "--embedding", 6},
                               {"-n", "--answers", 7}, {"-g", "--gpus", 8}, {"-F", "--filter", 9}};
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i], val;
        int id = -1;
        for (const Opt &op : opts) {
            std::string s = op.s, l = op.l;
            if (a == s || a == l) { id = op.id; if (i + 1 < argc) val = argv[++i]; else return false; break; }
            if (a.rfind(l + "=", 0) == 0) { id = op.id; val = a.substr(l.size() + 1); break; }
            if (a.rfind(s, 0) == 0 && a.size() > 2 && a[1] != '-') { id = op.id; val = a.substr(a[2] == '=' ? 3 : 2); break; }
        }
        if (a == "-h" || a == "--help") {
            std::puts("usage: main [-f dir/] [-d data.graph] [-q query.graph] [-m offline|online] [-p N] [-l N] [-e N] [-n MAX|N] [-g GPUS] [-F pe|pge]");
            std::exit(0);
        }
        if (id < 0) { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return false; }
        switch (id) {
            case 0: o.file = val; break;
            case 1: o.data = val; break;
            case 2: o.query = val; break;
            case 3: o.mode = val; break;
            case 4: o.partitions = (uint32_t)std::stoul(val); break;
            case 5: o.length = (uint32_t)std::stoul(val); break;
            case 6: o.embedding = (uint32_t)std::stoul(val); break;
            case 7: o.answers = val; break;
            case 8: o.gpus = (uint32_t)std::stoul(val); break;
            case 9: o.filter = val; break;
        }
    }
    return true;
}

struct Graph {
    uint32_t V = 0, E = 0;
    std::vector<uint32_t> off, nbr, lab;
};

bool load(const std::string &path, Graph &g) {
    if (gpe_host_load_graph(path.c_str(), &g.V, &g.E, nullptr, nullptr, nullptr) != GPE_OK) {
        std::cout << "Can not open the graph file " << path << " ." << std::endl;  // graph.cpp:166-168
        return false;
    }
    g.off.assign((size_t)g.V + 1, 0);
    g.nbr.assign(std::max<size_t>((size_t)g.E * 2, 1), 0);
    g.lab.assign(std::max<size_t>(g.V, 1), 0);
    return gpe_host_load_graph(path.c_str(), &g.V, &g.E, g.off.data(), g.nbr.data(), g.lab.data()) == GPE_OK;
}

void print_meta(const Graph &g) {  // Static_Graph::printGraphMetaData, graph.cpp:244-247
    std::unordered_map<uint32_t, uint32_t> freq;
    uint32_t max_label = 0, max_deg = 0, max_freq = 0;
    for (uint32_t v = 0; v < g.V; v++) {
        max_label = std::max(max_label, g.lab[v]);
        max_deg = std::max(max_deg, g.off[v + 1] - g.off[v]);
        max_freq = std::max(max_freq, ++freq[g.lab[v]]);
    }
    uint32_t labels_count = g.V ? std::max<uint32_t>((uint32_t)freq.size(), max_label + 1) : 0;  // graph.cpp:223
    std::cout << "|V|: " << g.V << ", |E|: " << g.E << ", |Σ|: " << labels_count << std::endl;
    std::cout << "Max Degree: " << max_deg << ", Max Label Frequency: " << max_freq << std::endl;
}

bool read_membership(const std::string &path, uint32_t V, std::vector<uint32_t> &sorted, std::vector<uint32_t> &member) {
    std::ifstream fin(path);
    if (!fin.is_open()) { std::fprintf(stderr, "cannot open %s (run gnnpe.py first)\n", path.c_str()); return false; }
    sorted.assign(V, 0);
    member.assign(V, 0);
    for (uint32_t i = 0; i < V; i++) {  // main.cpp:81-84
        uint32_t v, part;
        if (!(fin >> v >> part) || v >= V) { std::fprintf(stderr, "%s: bad line %u\n", path.c_str(), i); return false; }
        sorted[i] = v;
        member[v] = part;
    }
    return true;
}

// ---- offline -> online manifest ---------------------------------------------------------------------------------
struct Manifest {
    char magic[8];       // "GPEPATHS"
    uint32_t version;    // 1
    uint32_t L, p, V, E, reserved;
    uint64_t graph_hash, membership_hash, n_rows;
};

uint64_t fnv1a(const void *data, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char *b = static_cast<const unsigned char *>(data);
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

Manifest make_manifest(const Graph &g, const std::vector<uint32_t> &sorted, const std::vector<uint32_t> &member, uint32_t L,
                       uint32_t p, uint64_t n_rows) {
    Manifest m{};
    std::memcpy(m.magic, "GPEPATHS", 8);
    m.version = 1;
    m.L = L; m.p = p; m.V = g.V; m.E = g.E;
    uint64_t h = fnv1a(g.off.data(), ((size_t)g.V + 1) * 4);
    h = fnv1a(g.nbr.data(), (size_t)g.off[g.V] * 4, h);
    m.graph_hash = fnv1a(g.lab.data(), (size_t)g.V * 4, h);
    m.membership_hash = fnv1a(member.data(), member.size() * 4, fnv1a(sorted.data(), sorted.size() * 4));
    m.n_rows = n_rows;
    return m;
}

// 0 = matches, 1 = no manifest (outputs of the reference's own offline run, or none: nothing to check), 2 = stale
int check_manifest(const std::string &path, const Manifest &now, const std::vector<uint64_t> &rows_pp, std::string &why) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return 1;
    Manifest m{};
    std::vector<uint64_t> pp(now.p);
    bool ok = std::fread(&m, sizeof m, 1, f) == 1 && std::memcmp(m.magic, "GPEPATHS", 8) == 0 && m.version == 1;
    if (ok && m.p == now.p) ok = std::fread(pp.data(), sizeof(uint64_t), now.p, f) == now.p;
    std::fclose(f);
    if (!ok) { why = "unreadable or from another version"; return 2; }
    if (m.graph_hash != now.graph_hash || m.V != now.V || m.E != now.E) { why = "written for a different data graph"; return 2; }
    if (m.membership_hash != now.membership_hash) { why = "written for a different membership.txt"; return 2; }
    if (m.L != now.L) { why = "written for -l " + std::to_string(m.L - 1); return 2; }
    if (m.p != now.p) { why = "written for -p " + std::to_string(m.p); return 2; }
    if (m.n_rows != now.n_rows || pp != rows_pp) { why = "row counts differ"; return 2; }
    return 0;
}

// A directory of query graphs as one batch (every *.graph file, sorted by name).
struct QueryDir {
    std::vector<std::string> files;
    std::vector<uint32_t> vbase{0}, ebase{0}, offs, nbrs, labs;
    std::vector<uint64_t> limits;
    gpe_batch batch{};
};

// 0 = loaded, 1 = no query files, -1 = a file could not be read (message printed)
int load_query_dir(const std::string &path, uint64_t limit, QueryDir &qd) {
    if (DIR *dir = opendir(path.c_str())) {
        while (dirent *e = readdir(dir)) {
            std::string n = e->d_name;
            if (n.size() > 6 && n.compare(n.size() - 6, 6, ".graph") == 0) qd.files.push_back(n);
        }
        closedir(dir);
    }
    std::sort(qd.files.begin(), qd.files.end());
    if (qd.files.empty()) { std::fprintf(stderr, "no *.graph files in %s\n", path.c_str()); return 1; }
    for (const std::string &n : qd.files) {
        Graph Qi;
        if (!load(path + "/" + n, Qi)) return -1;
        qd.offs.insert(qd.offs.end(), Qi.off.begin(), Qi.off.end());
        qd.nbrs.insert(qd.nbrs.end(), Qi.nbr.begin(), Qi.nbr.begin() + Qi.off[Qi.V]);
        qd.labs.insert(qd.labs.end(), Qi.lab.begin(), Qi.lab.begin() + Qi.V);
        qd.vbase.push_back(qd.vbase.back() + Qi.V);
        qd.ebase.push_back(qd.ebase.back() + Qi.off[Qi.V]);
    }
    qd.nbrs.push_back(0);
    qd.limits.assign(qd.files.size(), limit);
    qd.batch = gpe_batch{(uint32_t)qd.files.size(), qd.vbase.data(), qd.ebase.data(), qd.offs.data(), qd.nbrs.data(),
                         qd.labs.data(), qd.limits.data()};
    return 0;
}

bool is_dir(const std::string &path) {
    struct stat st{};
    return stat(path.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

#define CK(ctx, call)                                                                  \
    do {                                                                               \
        if ((call) != GPE_OK) {                                                        \
            std::fprintf(stderr, "libgpe: %s\n", gpe_last_error(ctx));                 \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

// ---- GNN-PGE (GNN-PGE/src/main.cpp:38-363) ------------------------------------------------------------------------
// The reference also reads gnn-pge/membership.txt (:77-88), but only to deal the vertices to its per-partition R*-trees;
// there is one row per data vertex here and one scan over all of them, so the file is not needed.
int run_pge(const Options &o, uint64_t limit) {
    if (o.gpus > 1) { std::fprintf(stderr, "--filter pge runs on one GPU (one row per data vertex: nothing to shard)\n"); return 1; }
    const uint32_t pl = o.length, e = o.embedding;  // GNN-PGE's -l counts vertices (custom.h:47, main.cpp:58)
    Graph G;
    if (!load(o.data, G)) return -1;
    gpe_ctx *ctx = nullptr;
    if (gpe_create(0, &ctx) != GPE_OK) { std::fprintf(stderr, "libgpe: %s\n", gpe_last_error(nullptr)); return 1; }
    std::vector<double> x((size_t)G.V * e), vde((size_t)G.V * e);
    CK(ctx, gpe_host_gen_vde(G.V, G.off.data(), G.nbr.data(), G.lab.data(), e, x.data(), vde.data()));
    CK(ctx, gpe_set_graph(ctx, G.V, G.off.data(), G.nbr.data(), G.lab.data()));
    CK(ctx, gpe_set_embeddings(ctx, e, vde.data()));
    CK(ctx, gpe_pge_build(ctx, pl, x.data()));  // main.cpp:91-176 for every data vertex

    if (o.mode == "offline") {  // gnn-pge/data_vertices.bin, main.cpp:179-194
        const size_t W = (size_t)2 * pl * e;
        std::vector<double> pg((size_t)G.V * W), plg((size_t)G.V * W), nx(e);
        std::vector<uint8_t> has(std::max<size_t>(G.V, 1));
        CK(ctx, gpe_pge_dump_groups(ctx, pg.data(), plg.data(), has.data()));
        const std::string name = o.file + "gnn-pge/data_vertices.bin";
        FILE *f = std::fopen(name.c_str(), "wb");
        if (!f) { std::fprintf(stderr, "cannot write %s\n", name.c_str()); return 1; }
        bool ok = std::fwrite(&G.V, 4, 1, f) == 1;
        for (uint32_t v = 0; v < G.V && ok; v++) {
            const uint32_t head[3] = {v, G.lab[v], G.off[v + 1] - G.off[v]};
            const double key = 0;  // never set for data vertices: the value-initialised field (custom.h:73-88)
            std::fill(nx.begin(), nx.end(), 0.0);  // the neighbours' label embeddings, custom.h:460-471
            for (uint32_t j = G.off[v]; j < G.off[v + 1]; j++)
                for (uint32_t k = 0; k < e; k++) nx[k] += x[(size_t)G.nbr[j] * e + k];
            ok = std::fwrite(head, 4, 3, f) == 3 && std::fwrite(&key, 8, 1, f) == 1 &&
                 std::fwrite(&x[(size_t)v * e], 8, e, f) == e && std::fwrite(nx.data(), 8, e, f) == e &&
                 std::fwrite(&vde[(size_t)v * e], 8, e, f) == e && std::fwrite(&pg[v * W], 8, W, f) == W &&
                 std::fwrite(&plg[v * W], 8, W, f) == W;
        }
        if (std::fclose(f) != 0 || !ok) { std::fprintf(stderr, "cannot write %s\n", name.c_str()); return 1; }
    }

    if (o.mode == "online") {
        if (is_dir(o.query)) {
            QueryDir qd;
            if (int rc = load_query_dir(o.query, limit, qd)) return rc;
            std::vector<uint64_t> answers(qd.files.size(), 0);
            auto t0 = std::chrono::high_resolution_clock::now();
            CK(ctx, gpe_pge_query_batch(ctx, &qd.batch, answers.data()));
            auto t1 = std::chrono::high_resolution_clock::now();
            double ms = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count() / 1e6;
            for (size_t i = 0; i < qd.files.size(); i++)
                std::cout << qd.files[i] << ": Answer Num: " << (uint32_t)answers[i] << std::endl;
            std::cout << "Queries: " << qd.files.size() << " Query Time (ms): " << ms << " Queries/s: " << qd.files.size() / (ms / 1e3) << std::endl;
        } else {
            Graph Q;
            if (!load(o.query, Q)) return -1;
            uint32_t vbase[2] = {0, Q.V}, ebase[2] = {0, Q.off[Q.V]};
            gpe_batch batch{1, vbase, ebase, Q.off.data(), Q.nbr.data(), Q.lab.data(), &limit};
            uint64_t answer = 0;
            auto t0 = std::chrono::high_resolution_clock::now();  // query groups + filter + refinement, main.cpp:251-361
            CK(ctx, gpe_pge_query_batch(ctx, &batch, &answer));
            auto t1 = std::chrono::high_resolution_clock::now();
            double ms = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count() / 1e6;
            std::cout << "Answer Num: " << (uint32_t)answer << " Query Time (ms): " << ms << std::endl;  // main.cpp:361
        }
    }
    gpe_destroy(ctx);
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    Options o;
    if (!parse(argc, argv, o)) return 2;
    const uint32_t L = o.length + 1;  // main.cpp:58
    uint64_t limit = GPE_LIMIT_MAX;   // main.cpp:62-69
    if (o.answers != "MAX") limit = (uint64_t)(uint32_t)std::stoi(o.answers);
    if (o.filter == "pge") return run_pge(o, limit);
    if (o.filter != "pe") { std::fprintf(stderr, "--filter is pe or pge\n"); return 2; }
    const std::string partitions_path = o.file + "gnn-pe/partitions/";

    Graph G;
    if (!load(o.data, G)) return -1;
    print_meta(G);

    std::vector<uint32_t> sorted, member;
    if (!read_membership(o.file + "gnn-pe/membership.txt", G.V, sorted, member)) return 1;
    for (uint32_t v = 0; v < G.V; v++)
        if (member[v] >= o.partitions) { std::fprintf(stderr, "membership.txt names partition %u but -p is %u\n", member[v], o.partitions); return 1; }

    gpe_ctx *ctx = nullptr;
    if (gpe_create(0, &ctx) != GPE_OK) { std::fprintf(stderr, "libgpe: %s\n", gpe_last_error(nullptr)); return 1; }
    CK(ctx, gpe_set_graph(ctx, G.V, G.off.data(), G.nbr.data(), G.lab.data()));
    std::vector<uint64_t> rows_pp(o.partitions);
    uint64_t n_rows = 0;
    CK(ctx, gpe_enumerate(ctx, L, sorted.data(), member.data(), o.partitions, rows_pp.data(), &n_rows));

    if (o.mode == "offline") {
        if (n_rows > UINT_MAX) { std::fprintf(stderr, "%llu paths do not fit the reference's 32-bit text format\n", (unsigned long long)n_rows); return 1; }
        std::vector<uint64_t> start((size_t)G.V + 1);
        CK(ctx, gpe_start_rows(ctx, start.data()));
        for (uint32_t i = 0; i < o.partitions; i++) {  // main.cpp:98-108
            std::string name = partitions_path + "partition-" + std::to_string(i) + "/partition_paths.txt";
            FILE *f = std::fopen(name.c_str(), "w");
            if (!f) { std::fprintf(stderr, "cannot write %s\n", name.c_str()); return 1; }
            std::fprintf(f, "%llu\n", (unsigned long long)rows_pp[i]);
            for (uint32_t r = 0; r < G.V; r++)
                if (member[sorted[r]] == i)
                    for (uint64_t id = start[r]; id < start[r + 1]; id++) std::fprintf(f, "%llu\n", (unsigned long long)id);
            std::fclose(f);
        }
        std::string name = o.file + "/gnn-pe/all_paths.txt";  // main.cpp:110-119
        FILE *f = std::fopen(name.c_str(), "w");
        if (!f) { std::fprintf(stderr, "cannot write %s\n", name.c_str()); return 1; }
        std::fprintf(f, "%llu\n", (unsigned long long)n_rows);
        const uint64_t chunk = 1u << 22;
        std::vector<uint32_t> rows;
        for (uint64_t first = 0; first < n_rows; first += chunk) {
            uint64_t n = std::min(chunk, n_rows - first);
            rows.resize(n * L);
            CK(ctx, gpe_dump_paths(ctx, first, n, rows.data()));
            for (uint64_t r = 0; r < n; r++) {
                for (uint32_t k = 0; k < L; k++) std::fprintf(f, "%u ", rows[r * L + k]);
                std::fputc('\n', f);
            }
        }
        std::fclose(f);
        const Manifest m = make_manifest(G, sorted, member, L, o.partitions, n_rows);
        name = o.file + "gnn-pe/paths.gpe";
        f = std::fopen(name.c_str(), "wb");
        if (!f || std::fwrite(&m, sizeof m, 1, f) != 1 ||
            std::fwrite(rows_pp.data(), sizeof(uint64_t), rows_pp.size(), f) != rows_pp.size()) {
            std::fprintf(stderr, "cannot write %s\n", name.c_str());
            return 1;
        }
        std::fclose(f);
    }

    if (o.mode == "online") {
        std::string why;
        if (check_manifest(o.file + "gnn-pe/paths.gpe", make_manifest(G, sorted, member, L, o.partitions, n_rows), rows_pp, why) == 2) {
            std::fprintf(stderr, "%sgnn-pe/paths.gpe is stale (%s): run -m offline again\n", o.file.c_str(), why.c_str());
            return 1;
        }
        std::vector<double> x((size_t)G.V * o.embedding), vde((size_t)G.V * o.embedding);
        CK(ctx, gpe_host_gen_vde(G.V, G.off.data(), G.nbr.data(), G.lab.data(), o.embedding, x.data(), vde.data()));
        // -g N: one context per GPU (context 0 is the one that enumerated above), one communicator, sharded tables
        const int n_gpu = (int)std::max<uint32_t>(o.gpus, 1);
        if ((uint32_t)n_gpu > o.partitions) { std::fprintf(stderr, "-g %d needs -p >= %d (GPU r holds the partitions i %% %d == r)\n", n_gpu, n_gpu, n_gpu); return 1; }
        std::vector<gpe_ctx *> ctxs(1, ctx);
        for (int d = 1; d < n_gpu; d++) {
            gpe_ctx *cd = nullptr;
            if (gpe_create(d, &cd) != GPE_OK) { std::fprintf(stderr, "libgpe: GPU %d: %s\n", d, gpe_last_error(nullptr)); return 1; }
            ctxs.push_back(cd);
        }
        if (n_gpu > 1) CK(ctx, gpe_comm_init_all(ctxs.data(), n_gpu));
        {
            std::vector<int> rcs(n_gpu, GPE_OK);
            std::vector<std::thread> pool;
            for (int d = 0; d < n_gpu; d++)
                pool.emplace_back([&, d] {
                    gpe_ctx *cd = ctxs[d];
                    int rc = GPE_OK;
                    if (d > 0) {
                        rc = gpe_set_graph(cd, G.V, G.off.data(), G.nbr.data(), G.lab.data());
                        if (!rc) rc = gpe_enumerate(cd, L, sorted.data(), member.data(), o.partitions, nullptr, nullptr);
                    }
                    if (!rc) rc = gpe_set_embeddings(cd, o.embedding, vde.data());
                    uint64_t table_rows = 0;
                    if (!rc) rc = gpe_build_table_shard(cd, &table_rows);
                    rcs[d] = rc;
                });
            for (auto &t : pool) t.join();
            for (int d = 0; d < n_gpu; d++)
                if (rcs[d] != GPE_OK) { std::fprintf(stderr, "libgpe: GPU %d: %s\n", d, gpe_last_error(ctxs[d])); return 1; }
        }
        auto run_batch = [&](const gpe_batch *b, uint64_t *out) {
            return n_gpu == 1 ? gpe_query_batch(ctx, b, 0, out) : gpe_multi_query_batch(ctxs.data(), n_gpu, b, 0, out);
        };
        auto report = [&](int rc) {
            if (rc == GPE_OK) return false;
            for (gpe_ctx *cd : ctxs)
                if (gpe_last_error(cd)[0]) std::fprintf(stderr, "libgpe: %s\n", gpe_last_error(cd));
            return true;
        };
        auto destroy_all = [&] { for (gpe_ctx *cd : ctxs) gpe_destroy(cd); };

        if (is_dir(o.query)) {  // a directory of queries: one batch
            QueryDir qd;
            if (int rc = load_query_dir(o.query, limit, qd)) return rc;
            std::vector<uint64_t> answers(qd.files.size(), 0);
            auto t0 = std::chrono::high_resolution_clock::now();
            if (report(run_batch(&qd.batch, answers.data()))) return 1;
            auto t1 = std::chrono::high_resolution_clock::now();
            double ms = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count() / 1e6;
            for (size_t i = 0; i < qd.files.size(); i++)
                std::cout << qd.files[i] << ": Answer Number: " << (uint32_t)answers[i] << std::endl;
            std::cout << "Queries: " << qd.files.size() << " Query Time (ms): " << ms << " Queries/s: " << qd.files.size() / (ms / 1e3) << std::endl;
            destroy_all();
            return 0;
        }
        Graph Q;
        if (!load(o.query, Q)) return -1;
        uint32_t vbase[2] = {0, Q.V}, ebase[2] = {0, Q.off[Q.V]};
        gpe_batch batch{1, vbase, ebase, Q.off.data(), Q.nbr.data(), Q.lab.data(), &limit};
        uint32_t plan_size = 0;
        CK(ctx, gpe_host_query_plan(Q.V, Q.off.data(), Q.nbr.data(), Q.lab.data(), L, o.embedding, 0, nullptr, nullptr,
                                    nullptr, nullptr, &plan_size));
        std::cout << plan_size << std::endl;  // custom.h:630
        uint64_t answer = 0;
        auto t0 = std::chrono::high_resolution_clock::now();  // plan + filter + refinement, as main.cpp:148-179 sums
        if (report(run_batch(&batch, &answer))) return 1;
        auto t1 = std::chrono::high_resolution_clock::now();
        double ms = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count() / 1e6;
        std::cout << "Answer Number: " << (uint32_t)answer << " Query Time (ms): " << ms << std::endl;  // main.cpp:179
        destroy_all();
        return 0;
    }
    gpe_destroy(ctx);
    return 0;
}
