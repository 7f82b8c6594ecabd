// ref_probe.cpp -- drives the UNMODIFIED reference header (/root/reference/GNN-PE/include/custom.h)
// and prints its intermediate values as JSON, so the oracle restatement can be pinned on more
// than the single "Answer Number" line the reference binary prints.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_ref/probe from the sources
// where they lie under /root/reference (nothing is copied into this repo); used by
// tests/golden/make_golden.py to generate the committed golden vectors.
//
// The include order below is the one src/main.cpp:1-13 uses (gendef.h defines min/max macros,
// SURVEY.md Q12).  The probe replaces main.cpp's driver only: everything it calls is the
// reference's own code.
//
// usage: probe <dataset_dir/> <data.graph> <query.graph> <p> <l> <e> <start_depth|-1> [limit]
//   start_depth -1  -> main.cpp:95/145's own `path_length - 2`
//   start_depth 1   -> the patched variant SURVEY.md F5 describes for l != 2
// The dataset dir must already hold gnn-pe/all_paths.txt and partition_paths.txt files
// (from `main -m offline`, or written by make_golden.py for patched l).
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

static void print_vec_u(const char *name, const std::vector<ui> &v, bool comma = true)
{
	printf("\"%s\": [", name);
	for (size_t i = 0; i < v.size(); i++)
		printf("%s%u", i ? ", " : "", v[i]);
	printf("]%s\n", comma ? "," : "");
}

static void print_hex_d(const std::vector<double> &v)
{
	printf("[");
	for (size_t i = 0; i < v.size(); i++)
	{
		unsigned long long bits;
		memcpy(&bits, &v[i], 8);
		printf("%s\"%016llx\"", i ? ", " : "", bits);
	}
	printf("]");
}

int main(int argc, char **argv)
{
	// offline mode: main.cpp:77-119 with a caller-chosen DFS start depth (the l != 2 patch of
	// SURVEY.md F5); `dfs` itself is the reference's.
	//   probe offline <dataset_dir/> <data.graph> <p> <l> <start_depth>
	if (argc == 7 && string(argv[1]) == "offline")
	{
		string dir = argv[2];
		partition_num = atoi(argv[4]);
		path_length = atoi(argv[5]) + 1;
		ui sd = atoi(argv[6]);
		Static_Graph *G = new Static_Graph(true);
		G->loadGraphFromFile(argv[3]);
		vector<ui> membership(G->getVerticesCount()), sorted_nodes(G->getVerticesCount());
		ifstream fin(dir + "gnn-pe/membership.txt");
		for (ui i = 0; i < G->getVerticesCount(); i++)
			fin >> sorted_nodes[i] >> membership[sorted_nodes[i]];
		vector<vector<ui>> partitions_paths(partition_num), all_paths;
		unordered_set<vector<ui>, VectorHash> all_paths_set;
		for (ui node : sorted_nodes)
		{
			vector<ui> path = {node};
			dfs(node, sd, path, G, all_paths, all_paths_set, partitions_paths[membership[node]]);
		}
		for (ui i = 0; i < partition_num; i++)
		{
			ofstream fout(dir + "gnn-pe/partitions/partition-" + to_string(i) + "/partition_paths.txt");
			fout << partitions_paths[i].size() << endl;
			for (ui id : partitions_paths[i])
				fout << id << endl;
		}
		ofstream fout(dir + "/gnn-pe/all_paths.txt");
		fout << all_paths.size() << endl;
		for (auto &row : all_paths)
		{
			for (ui j = 0; j < path_length; j++)
				fout << row[j] << " ";
			fout << endl;
		}
		return 0;
	}
	if (argc < 8)
	{
		fprintf(stderr, "usage: probe dir data query p l e start_depth [limit]\n");
		return 2;
	}
	string dataset_path = argv[1];
	string data_name = argv[2];
	string query_name = argv[3];
	partition_num = atoi(argv[4]);
	path_length = atoi(argv[5]);
	vde_dim = atoi(argv[6]);
	int start_depth_arg = atoi(argv[7]);
	if (argc > 8)
		MAX_LIMIT = (ui)stoi(argv[8]);

	path_length += 1;
	pde_dim = vde_dim * path_length;
	ui start_depth = start_depth_arg < 0 ? path_length - 2 : (ui)start_depth_arg;

	// stdout belongs to the JSON; the reference prints its own lines there too, so
	// collect those in a string stream and emit them as a field.
	std::stringstream captured;
	std::streambuf *old = cout.rdbuf(captured.rdbuf());

	Static_Graph *G = new Static_Graph(true);
	G->loadGraphFromFile(data_name);
	Static_Graph *Q = new Static_Graph(true);
	Q->loadGraphFromFile(query_name);

	vector<Vertex> data_vertices = gen_vde(G);
	vector<Path> data_paths = gen_pde(data_vertices, dataset_path + "/gnn-pe/all_paths.txt");

	vector<Partition> partitions;
	string partitions_path = dataset_path + "gnn-pe/partitions/";
	for (ui i = 0; i < partition_num; i++)
	{
		Partition partition(data_paths, partitions_path + "partition-" + to_string(i) + "/");
		partitions.push_back(partition);
	}

	vector<vector<ui>> all_paths;
	unordered_set<vector<ui>, VectorHash> all_paths_set;
	for (ui node = 0; node < Q->getVerticesCount(); node++)
	{
		vector<ui> path = {node};
		dfs_query(node, start_depth, path, Q, all_paths, all_paths_set);
	}
	vector<Vertex> query_vertices = gen_vde(Q);
	vector<Query_Path> plan = gen_query_pde(query_vertices, all_paths);
	Query_Plan QP(plan);

	ui nq = Q->getVerticesCount();
	vector<vector<set<ui>>> candidate_sets(partition_num, vector<set<ui>>(nq));
	vector<set<ui>> candidate_set(nq);
	for (ui pid = 0; pid < partition_num; pid++)
		partitions[pid].query(nq, candidate_sets[pid], QP);
	for (ui pid = 0; pid < partition_num; pid++)
		for (ui i = 0; i < nq; i++)
			candidate_set[i].insert(candidate_sets[pid][i].begin(), candidate_sets[pid][i].end());

	// all-pairs leaf compare over the reference's own Path records (custom.h:407-435 verbatim semantics)
	vector<unsigned long long> survivors(plan.size(), 0);
	vector<set<ui>> brute(nq);
	for (size_t r = 0; r < data_paths.size(); r++)
		for (size_t j = 0; j < plan.size(); j++)
		{
			ui k = 0;
			for (; k < path_length; k++)
				if (plan[j].labels[k] != data_paths[r].labels[k] || plan[j].degrees[k] > data_paths[r].degrees[k])
					break;
			if (k != path_length)
				continue;
			for (k = 0; k < pde_dim; k++)
				if (plan[j].pde[k] > data_paths[r].pde[k] && abs(plan[j].pde[k] - data_paths[r].pde[k]) > epsilon)
					break;
			if (k != pde_dim)
				continue;
			survivors[j]++;
			for (k = 0; k < path_length; k++)
				brute[plan[j].vids[k]].insert(data_paths[r].vids[k]);
		}
	bool index_equals_brute = true;
	for (ui i = 0; i < nq; i++)
		index_equals_brute = index_equals_brute && (brute[i] == candidate_set[i]);

	vector<ui> counts(nq);
	for (ui i = 0; i < nq; i++)
		counts[i] = candidate_set[i].size();
	ui *order = NULL, *pivot = NULL;
	generateGQLQueryPlan(G, Q, counts.data(), order, pivot);

	ui answer = 0;
	refinement(G, Q, candidate_set, answer);

	cout.rdbuf(old);

	printf("\n=====JSON=====\n{\n");
	printf("\"V\": %u, \"E\": %u, \"labels_count\": %u, \"max_degree\": %u, \"max_label_freq\": %u,\n",
		   G->getVerticesCount(), G->getEdgesCount(), G->getLabelsCount(), G->getGraphMaxDegree(), G->getGraphMaxLabelFrequency());
	printf("\"p\": %u, \"L\": %u, \"e\": %u, \"start_depth\": %u,\n", partition_num, path_length, vde_dim, start_depth);
	printf("\"n_data_paths\": %zu, \"n_query_paths\": %zu, \"plan_size\": %zu,\n", data_paths.size(), all_paths.size(), plan.size());
	printf("\"query_paths\": [");
	for (size_t i = 0; i < all_paths.size(); i++)
	{
		printf("%s[", i ? ", " : "");
		for (size_t k = 0; k < all_paths[i].size(); k++)
			printf("%s%u", k ? ", " : "", all_paths[i][k]);
		printf("]");
	}
	printf("],\n\"plan\": [\n");
	for (size_t j = 0; j < plan.size(); j++)
	{
		printf("  {\"vids\": [");
		for (ui k = 0; k < path_length; k++)
			printf("%s%u", k ? ", " : "", plan[j].vids[k]);
		printf("], \"weight\": %u, \"survivors\": %llu, \"pde\": ", plan[j].weight, survivors[j]);
		print_hex_d(plan[j].pde);
		printf("}%s\n", j + 1 < plan.size() ? "," : "");
	}
	printf("],\n");
	printf("\"label_x\": {");
	{
		set<ui> labs;
		for (ui v = 0; v < nq; v++)
			labs.insert(Q->getVertexLabel(v));
		labs.insert(0);
		labs.insert(1);
		bool first = true;
		for (ui lab : labs)
		{
			printf("%s\"%u\": ", first ? "" : ", ", lab);
			print_hex_d(gen_vde_x(lab));
			first = false;
		}
	}
	printf("},\n\"data_vde_sample\": {");
	{
		ui V = G->getVerticesCount();
		ui picks[6] = {0, 1, V / 3, V / 2, V - 2, V - 1};
		set<ui> done;
		bool first = true;
		for (ui t = 0; t < 6; t++)
		{
			ui v = picks[t] < V ? picks[t] : 0;
			if (!done.insert(v).second)
				continue;
			printf("%s\"%u\": ", first ? "" : ", ", v);
			print_hex_d(data_vertices[v].vde);
			first = false;
		}
	}
	printf("},\n\"candidates\": [\n");
	for (ui i = 0; i < nq; i++)
	{
		printf("  [");
		bool first = true;
		for (ui v : candidate_set[i])
		{
			printf("%s%u", first ? "" : ", ", v);
			first = false;
		}
		printf("]%s\n", i + 1 < nq ? "," : "");
	}
	printf("],\n");
	print_vec_u("candidate_counts", counts);
	printf("\"index_equals_brute\": %s,\n", index_equals_brute ? "true" : "false");
	vector<ui> ord(order, order + nq), piv(pivot, pivot + nq);
	piv[0] = 4294967295u;
	print_vec_u("order", ord);
	print_vec_u("pivot", piv);
	printf("\"answer\": %u\n}\n", answer);
	return 0;
}
