// common.hpp -- declarations shared by the host translation units of libbsb200.so.
#pragma once
#include "../../include/bsb200.h"

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <thread>
#include <vector>

namespace bsb {

extern thread_local std::string g_last_error;
extern std::atomic<int> g_cancel;
extern std::atomic<long long> g_launches;
extern bsb200_poll_fn g_poll;
extern void *g_poll_user;
// bsb200_cancel() seen, or the registered interrupt poll asks to stop (include/bsb200.h)
bool cancel_requested();
// snapshot sink (chain_file.cpp): rows of z / weights / Y as they are produced, mu / lambda / plogLik / Ychange at the end
bool sink_active();
int sink_write(const char *param, int64_t row, const void *data, int64_t count, int is_f64, const int64_t *dims, int ndims);

// Records the message bsb200_last_error() returns and hands back `code`.
int fail(int code, const char *fmt, ...);

int lattice_neighbors(const int32_t *row, const int32_t *col, int64_t n, int platform,
                      std::vector<std::vector<int32_t>> &adj);
int radius_neighbors(const double *pos, int64_t n, double radius, int metric,
                     std::vector<std::vector<int32_t>> &adj);
int parse_neighbor_string(const char *s, std::vector<int32_t> &out);
int subspot_neighbors(const double *pos, int64_t n, int64_t n0, const int64_t *sptr,
                      const int32_t *sidx, double dist, std::vector<std::vector<int32_t>> &adj);
int colour_graph(const int64_t *ptr, const int32_t *idx, int64_t n, std::vector<int32_t> &colour,
                 int32_t *ncolours);
bool colouring_is_proper(const int64_t *ptr, const int32_t *idx, int64_t n, const int32_t *colour);

// Host threads for the O(n) set-up loops (validation, ordering, ELL staging): chunks [lo, hi) of [0, n) run on up to
// BSB200_HOST_THREADS (default: min(hardware threads, 16)) threads; small n runs inline.  fn must be thread-safe.
int host_threads();
void set_host_share(int ranks_on_this_host);   // sharded chains: host threads per rank = host_threads() / ranks
template <typename F>
inline void parallel_for(int64_t n, F fn, int64_t min_chunk = 1 << 18) {
  int nt = host_threads();
  if ((int64_t) nt > n / min_chunk) nt = (int) (n / min_chunk);
  if (nt <= 1) { fn((int64_t) 0, n, 0); return; }
  std::vector<std::thread> th;
  th.reserve((size_t) nt - 1);
  for (int t = 1; t < nt; t++) th.emplace_back([=]() { fn(n * t / nt, n * (t + 1) / nt, t); });
  fn((int64_t) 0, n / nt, 0);
  for (auto &x : th) x.join();
}
int host_threads_max();
inline int parallel_chunks(int64_t n, int64_t min_chunk = 1 << 18) {   // upper bound of the chunks parallel_for will use
  int nt = host_threads_max();
  if ((int64_t) nt > n / min_chunk) nt = (int) (n / min_chunk);
  return nt < 1 ? 1 : nt;
}

}   // namespace bsb
