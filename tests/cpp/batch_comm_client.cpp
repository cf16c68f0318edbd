// C++ client of the batch driver and of the NCCL all-gather behind the C ABI (include/b200reg.h, b200reg_batch_* /
// b200reg_comm_* / b200reg_allgather_results).  One rank per GPU as HOST THREADS of one process (the ABI's contexts are
// per device, so the per-device kernel attributes are exercised too):
//   batch_comm_client <src.bin> <dst.bin> <n_src> <n_dst> <world>
// every rank registers the same pair through its own batch driver, all-gathers the result records and checks that
// every gathered record is byte-identical.  Prints one JSON line.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/b200reg.h"

static std::vector<float> read_bin(const char* path, size_t n) {
  std::vector<float> v(n * 4);
  FILE* f = fopen(path, "rb");
  if (!f || fread(v.data(), sizeof(float), v.size(), f) != v.size()) {
    fprintf(stderr, "cannot read %s\n", path);
    exit(2);
  }
  fclose(f);
  return v;
}

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  const size_t ns = strtoul(argv[3], nullptr, 10), nd = strtoul(argv[4], nullptr, 10);
  const int world = atoi(argv[5]);
  std::vector<float> src = read_bin(argv[1], ns), dst = read_bin(argv[2], nd);
  unsigned char id[B200REG_UNIQUE_ID_BYTES];
  if (b200reg_comm_unique_id(id)) {
    fprintf(stderr, "unique id: %s\n", b200reg_last_error());
    return 1;
  }
  std::vector<int> status(world, 0);
  std::vector<std::vector<b200reg_result>> gathered(world);
  std::vector<double> latency(world, 0.0);
  auto rank_main = [&](int rank) {
    b200reg_batch* batch = nullptr;
    b200reg_ctx* ctx = nullptr;
    int rc = b200reg_batch_create(rank, 2, &batch);
    if (!rc) rc = b200reg_ctx_create(rank, &ctx);
    if (!rc) rc = b200reg_comm_init(ctx, id, rank, world);
    b200reg_gicp_params gp;
    b200reg_default_gicp_params(&gp);
    const float* sp[2] = {src.data(), src.data()};
    const float* dp[2] = {dst.data(), dst.data()};
    const size_t sn[2] = {ns, ns}, dn[2] = {nd, nd};
    b200reg_result local[2];
    if (!rc) {
      const int64_t t = b200reg_batch_submit_icp(batch, 2, sp, sn, dp, dn, 16, 0, &gp, local);
      rc = t < 0 ? (int)t : b200reg_batch_wait(batch, t, &latency[rank]);
    }
    gathered[rank].resize(2 * (size_t)world);
    if (!rc) rc = b200reg_allgather_results(ctx, local, 2, gathered[rank].data());
    if (rc) fprintf(stderr, "rank %d: %s\n", rank, b200reg_last_error());
    status[rank] = rc;
    b200reg_comm_destroy(ctx);
    b200reg_ctx_destroy(ctx);
    b200reg_batch_destroy(batch);
  };
  std::vector<std::thread> th;
  for (int r = 0; r < world; r++) th.emplace_back(rank_main, r);
  for (auto& t : th) t.join();
  int bad = 0;
  for (int r = 0; r < world; r++) bad |= status[r];
  bool same = !bad;
  for (int r = 0; r < world && same; r++)
    for (int k = 0; k < 2 * world; k++) same &= memcmp(&gathered[r][k], &gathered[0][0], sizeof(b200reg_result)) == 0;
  const b200reg_result& g = gathered[0][0];
  printf("{\"status\": %d, \"world\": %d, \"identical\": %s, \"converged\": %d, \"fitness\": %.17g, \"latency_ms\": %.3f, \"T\": [", bad, world,
         same ? "true" : "false", g.converged, g.fitness, latency[0]);
  for (int i = 0; i < 16; i++) printf("%.17g%s", g.T[i], i < 15 ? ", " : "");
  printf("]}\n");
  return bad || !same;
}
