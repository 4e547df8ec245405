// Command line of the reference (main.cc:10-139) in front of the B200 engine: same flag names,
// same derived pruning parameters, same index-file naming, same console output.
//   hs_main --dataset=sift --solve_strategy=hnsw_slim --k=10 --m=16 --ef_construction=200 --ef_search=100
// Extra flags: --data_dir (default ../data), --index_dir (default ../statistics/index), --device, and for
// corpora sharded over the GPUs of one box: --shards S --gpus N [--batch B] with --solve_strategy=hnsw_slim
// (S sub-graphs built on the GPUs, queries in batches of B through one hs_shardgroup per GPU);
// --solve_strategy=hnsw_slim_serve [--threads T --max_batch B --max_wait_us U --patches a.bin,b.bin
// --patch_inline_last --index_path P]: the query set as single queries from T concurrent callers through hs_service
// (the server's /query handler), after applying the reference server's delta patches (gpu_service.h).
#include <cmath>
#include <cstring>
#include <map>

#include "gpu_service.h"
#include "gpu_strategies.h"

static std::map<std::string, std::string> parse_flags(int argc, char **argv) {   // gflags syntax: --name=value | --name value
  std::map<std::string, std::string> f;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a.rfind("--", 0) != 0 && a.rfind("-", 0) == 0) a = "-" + a;
    if (a.rfind("--", 0) != 0) {
      std::cout << "unexpected argument: " << a << std::endl;
      exit(1);
    }
    a = a.substr(2);
    const size_t eq = a.find('=');
    if (eq != std::string::npos) {
      f[a.substr(0, eq)] = a.substr(eq + 1);
    } else if (i + 1 < argc && std::strncmp(argv[i + 1], "--", 2) != 0) {
      f[a] = argv[++i];
    } else {
      f[a] = "true";
    }
  }
  return f;
}

int main(int argc, char **argv) {
  auto flags = parse_flags(argc, argv);
  auto str = [&](const char *n, const std::string &d) { return flags.count(n) ? flags[n] : d; };
  auto i64 = [&](const char *n, long long d) { return flags.count(n) ? std::stoll(flags[n]) : d; };
  auto f64 = [&](const char *n, double d) { return flags.count(n) ? std::stod(flags[n]) : d; };
  static const char *known[] = {"dataset", "solve_strategy", "k", "m", "m0", "ef_construction", "ef_search",
                                "branching_factor", "threshold_level", "top_degree_percent0", "top_degree_percent",
                                "top_M0", "low_m0", "top_M", "low_m", "level_ratio", "Mm_ratio", "min_indegree0",
                                "min_indegree", "data_dir", "index_dir", "device", "shards", "gpus", "batch", "index_path",
                                "threads", "max_batch", "max_wait_us", "patches", "patch_inline_last"};
  for (auto &kv : flags) {
    bool ok = false;
    for (const char *k : known) ok |= kv.first == k;
    if (!ok) {
      std::cout << "ERROR: unknown command line flag '" << kv.first << "'" << std::endl;   // as gflags does
      return 1;
    }
  }

  const std::string dataset = str("dataset", "sift");                  // main.cc:10-38
  std::string solve_strategy = str("solve_strategy", "hnsw_slim");
  K = (size_t)i64("k", (long long)K);
  M = (size_t)i64("m", (long long)M);
  M0 = (size_t)i64("m0", (long long)M0);
  EF_CONSTRUCTION = (size_t)i64("ef_construction", 128);
  EF_SEARCH = (size_t)i64("ef_search", 128);
  BRANCHING_FACTOR = str("branching_factor", BRANCHING_FACTOR);
  THRESHOLD_LEVEL = (size_t)i64("threshold_level", (long long)THRESHOLD_LEVEL);
  const int device = (int)i64("device", 0);

  const size_t level_ratio = (size_t)i64("level_ratio", 50), Mm_ratio = (size_t)i64("Mm_ratio", 25);
  const double ratio = 1.0 * level_ratio / 100.0;                      // main.cc:58-70
  PruneParams pp;
  pp.top_degree_percent0 = (float)f64("top_degree_percent0", 0.02);
  pp.top_degree_percent = pp.top_degree_percent0;
  pp.top_M0 = (size_t)i64("top_M0", 32);
  pp.low_m0 = pp.top_M0 * Mm_ratio / 100;
  pp.top_M = (size_t)(ratio * pp.top_M0);
  pp.low_m = (size_t)(ratio * pp.low_m0);
  pp.threshold_level = (int)THRESHOLD_LEVEL;

  // README / usage text spell two families with hyphens (README.md:98,114; main.cc:136)
  if (solve_strategy == "hnsw-slimq") solve_strategy = "hnsw_slimq";
  if (solve_strategy == "hnsw-slim") solve_strategy = "hnsw_slim";
  if (solve_strategy == "hnsw-slimzero") solve_strategy = "hnsw_slimzero";

  std::string suffix = solve_strategy + "_";                           // main.cc:80-100
  suffix += std::to_string(EF_CONSTRUCTION) + "_";
  suffix += std::to_string(M) + "_";
  suffix += BRANCHING_FACTOR;
  std::string tmp_suffix = "";
  tmp_suffix += "_" + std::to_string(pp.threshold_level);
  tmp_suffix += "_" + std::to_string(pp.top_degree_percent0);
  tmp_suffix += "_" + std::to_string(pp.top_degree_percent);
  tmp_suffix += "_" + std::to_string(pp.top_M0);
  tmp_suffix += "_" + std::to_string(pp.low_m0);
  tmp_suffix += "_" + std::to_string(pp.top_M);
  tmp_suffix += "_" + std::to_string(pp.low_m);
  suffix += tmp_suffix;
  suffix += ".graph";

  const std::string data_dir = str("data_dir", "../data"), index_dir = str("index_dir", "../statistics/index");
  const std::string source_path = data_dir + "/" + dataset + "/" + dataset + "_base.fvecs";
  const std::string query_path = data_dir + "/" + dataset + "/" + dataset + "_query.fvecs";
  const std::string gt_path = data_dir + "/" + dataset + "/" + dataset + "_groundtruth.ivecs";
  const std::string index_path = str("index_path", index_dir + "/" + dataset + "/" + suffix);

  std::cout << "Index path: " << index_path << std::endl;
  std::cout << "gt path: " << gt_path << std::endl;
  std::cout << "Running with param: " << "alpha0%: " << pp.top_degree_percent0 << ", " << "alpha%: "
            << pp.top_degree_percent << ", " << "top_m0: " << pp.top_M0 << ", " << "top_m: " << pp.top_M << ", "
            << "low_m0: " << pp.low_m0 << ", " << "low_m: " << pp.low_m << ", " << std::endl;

  try {
    SolveStrategy *strategy = nullptr;
    const size_t shards = (size_t)i64("shards", 0), gpus = (size_t)i64("gpus", 1), batch = (size_t)i64("batch", 10000);
    if (solve_strategy == "hnsw_slim" && shards > 0) {
      strategy = new HnswSlimShardedGpuStrategy(source_path, query_path, index_path, pp, shards, gpus, batch, device);
    } else if (solve_strategy == "hnsw_slim_serve") {      // the server's /query handler under concurrent callers + client-side patches
      strategy = new HnswSlimServeGpuStrategy(source_path, query_path, index_path, (size_t)i64("threads", 32),
                                              (size_t)i64("max_batch", 4096), (unsigned)i64("max_wait_us", 0),
                                              str("patches", ""), str("patch_inline_last", "false") == "true", device);
    } else if (solve_strategy == "hnsw_slim") {
      strategy = new HnswSlimGpuStrategy(source_path, query_path, index_path, pp, device);
    } else if (solve_strategy == "hnsw_slimq") {
      strategy = new HnswSlimQGpuStrategy(source_path, query_path, index_path, pp, device);
    } else if (solve_strategy == "hnsw") {
      strategy = new HnswGpuStrategy(source_path, query_path, index_path, device);
    } else if (solve_strategy == "hnsw_slimzero") {
      strategy = new HnswSlimZeroGpuStrategy(source_path, query_path, index_path, device);
    } else if (solve_strategy == "bruteforce") {
      strategy = new BruteForceGpu(source_path, query_path, index_path, gt_path, 100, device);
    } else {
      std::cout << "Unknown strategy: " << solve_strategy << std::endl;
      std::cout << "['hnsw', 'hnsw_slim', 'bruteforce', 'hnsw-slimq', 'hnsw-slimzero']" << std::endl;
      return 1;
    }
    strategy->solve();
    std::cout << "Solve strategy: " + solve_strategy << std::endl;
    strategy->recall(gt_path);
    std::cout << "Recall: " + gt_path << std::endl;
    delete strategy;
  } catch (const std::exception &e) {          // the reference lets std::runtime_error terminate the process
    std::cerr << "terminate called after throwing an instance of 'std::runtime_error'\n  what():  " << e.what()
              << std::endl;
    return 134;
  }
  return 0;
}
