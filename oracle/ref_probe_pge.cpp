// ref_probe_pge.cpp -- drives the UNMODIFIED GNN-PGE header (/root/reference/GNN-PGE/include/custom.h): the
// reference's own Partition (R*-tree build + auxiliary index), Partition::query and refinement, and prints the
// candidate sets and the answer as JSON, so the oracle's GNN-PGE restatement is pinned on more than the one
// "Answer Num" line the reference binary prints.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_ref/pge_probe from the sources where they lie
// under /root/reference (nothing is copied into this repo); used by tests/golden/make_golden_pge.py.
//
// The reference computes the path groups inline in main() (src/main.cpp:91-176 data, :226-297 query), so they
// cannot be called; the probe takes them from files the unmodified reference binary wrote itself:
//   <dataset_dir>/gnn-pge/data_vertices.bin        from `pge_main -m offline` on the data graph
//   <query_dir>/gnn-pge/data_vertices.bin          from `pge_main -m offline` run with the QUERY graph as data graph
// (record layout: src/main.cpp:179-194).  Only the three lines that derive a query vertex's key from its group
// (src/main.cpp:293-297) are restated here.
//
// usage: pge_probe <dataset_dir/> <data.graph> <query_dir/> <query.graph> <p> <path_vertices> <e> [limit]
#include "./rtree/rtree.h"
#include "./rtree/rtnode.h"
#include "./rtree/entry.h"
#include "./blockfile/blk_file.h"
#include "./blockfile/cache.h"
#include "./linlist/linlist.h"
#include "./rtree/rtree_cmd.h"
#include "rand.h"
#include "cdf.h"

#include "./graph/graph.h"
#include "custom.h"

#define NOMINMAX
#undef min
#undef max

#include <cstdio>
#include <sstream>
#include <unistd.h>

static bool read_vertices(const string &path, vector<Vertex> &out)
{
	ifstream fin(path, std::ios::binary);
	if (!fin.is_open())
		return false;
	ui count = 0;
	fin.read(reinterpret_cast<char *>(&count), sizeof(ui));
	for (ui i = 0; i < count; i++)
	{
		Vertex v;
		fin.read(reinterpret_cast<char *>(&v.vid), sizeof(ui));
		fin.read(reinterpret_cast<char *>(&v.label), sizeof(ui));
		fin.read(reinterpret_cast<char *>(&v.degree), sizeof(ui));
		fin.read(reinterpret_cast<char *>(&v.key), sizeof(double));
		v.x.resize(vde_dim);
		v.nx.resize(vde_dim);
		v.vde.resize(vde_dim);
		v.path_group.resize(pde_dim * 2);
		v.path_label_group.resize(pde_dim * 2);
		fin.read(reinterpret_cast<char *>(v.x.data()), vde_dim * sizeof(double));
		fin.read(reinterpret_cast<char *>(v.nx.data()), vde_dim * sizeof(double));
		fin.read(reinterpret_cast<char *>(v.vde.data()), vde_dim * sizeof(double));
		fin.read(reinterpret_cast<char *>(v.path_group.data()), pde_dim * 2 * sizeof(double));
		fin.read(reinterpret_cast<char *>(v.path_label_group.data()), pde_dim * 2 * sizeof(double));
		out.push_back(v);
	}
	return (bool)fin;
}

int main(int argc, char **argv)
{
	if (argc < 8)
	{
		fprintf(stderr, "usage: pge_probe dir data qdir query p path_vertices e [limit]\n");
		return 2;
	}
	string dataset_path = argv[1], data_name = argv[2], query_dir = argv[3], query_name = argv[4];
	partition_num = atoi(argv[5]);
	path_length = atoi(argv[6]);
	vde_dim = atoi(argv[7]);
	pde_dim = vde_dim * path_length;
	if (argc > 8)
		MAX_LIMIT = (ui)stoi(argv[8]);

	// the reference prints its own lines on stdout (graph meta, R-tree sizes): keep them out of the JSON
	std::stringstream captured;
	std::streambuf *old = cout.rdbuf(captured.rdbuf());
	FILE *real_stdout = fdopen(dup(fileno(stdout)), "w");
	freopen("/dev/null", "w", stdout);

	Static_Graph *G = new Static_Graph(true);
	G->loadGraphFromFile(data_name);
	Static_Graph *Q = new Static_Graph(true);
	Q->loadGraphFromFile(query_name);

	vector<ui> membership(G->getVerticesCount()), sorted_nodes(G->getVerticesCount());
	{
		ifstream fin(dataset_path + "gnn-pge/membership.txt");
		for (ui i = 0; i < G->getVerticesCount(); i++)
			fin >> sorted_nodes[i] >> membership[sorted_nodes[i]];
	}
	vector<vector<ui>> partition_vertices(partition_num);
	for (ui node : sorted_nodes)
		partition_vertices[membership[node]].push_back(node);

	vector<Vertex> data_vertices, query_vertices;
	if (!read_vertices(dataset_path + "gnn-pge/data_vertices.bin", data_vertices) ||
		!read_vertices(query_dir + "gnn-pge/data_vertices.bin", query_vertices))
	{
		fprintf(stderr, "cannot read data_vertices.bin (run pge_main -m offline on both graphs first)\n");
		return 1;
	}
	for (Vertex &v : query_vertices)  // src/main.cpp:293-297
	{
		v.key = 0;
		for (ui i = 0; i < pde_dim; i++)
			v.key -= v.path_group[2 * i];
	}

	double index_build_time = 0;
	vector<Partition> partitions;
	for (ui i = 0; i < partition_num; i++)
	{
		string partition_path = dataset_path + "gnn-pge/partitions/partition-" + to_string(i) + "/";
		Partition partition(data_vertices, partition_path, partition_vertices[i], index_build_time);
		partitions.push_back(partition);
	}

	Query_Plan plan(query_vertices);
	vector<vector<set<ui>>> candidate_sets(partition_num, vector<set<ui>>(Q->getVerticesCount()));
	vector<set<ui>> candidate_set(Q->getVerticesCount());
	for (ui pid = 0; pid < partition_num; pid++)
		partitions[pid].query(candidate_sets[pid], plan);
	for (ui pid = 0; pid < partition_num; pid++)
		for (ui i = 0; i < Q->getVerticesCount(); i++)
			candidate_set[i].insert(candidate_sets[pid][i].begin(), candidate_sets[pid][i].end());

	ui answer_num = 0;
	refinement(G, Q, candidate_set, answer_num);

	cout.rdbuf(old);
	fprintf(real_stdout, "{\"answer\": %u, \"candidates\": [", answer_num);
	for (ui i = 0; i < Q->getVerticesCount(); i++)
	{
		fprintf(real_stdout, "%s[", i ? ", " : "");
		bool first = true;
		for (ui v : candidate_set[i])
		{
			fprintf(real_stdout, "%s%u", first ? "" : ", ", v);
			first = false;
		}
		fprintf(real_stdout, "]");
	}
	fprintf(real_stdout, "]}\n");
	fclose(real_stdout);
	return 0;
}
