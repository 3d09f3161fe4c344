// common.hpp -- declarations shared by the host translation units of libbsb200.so.
#pragma once
#include "../../include/bsb200.h"

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace bsb {

extern thread_local std::string g_last_error;
extern std::atomic<int> g_cancel;
extern std::atomic<long long> g_launches;
extern bsb200_poll_fn g_poll;
extern void *g_poll_user;
// bsb200_cancel() seen, or the registered interrupt poll asks to stop (include/bsb200.h)
bool cancel_requested();

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

}   // namespace bsb
