// graph_host.cpp -- host side of the neighbour construction behind the samplers and the graph
// colouring the chromatic scan needs.  Pure C++ (no CUDA): these run once per chain.
//
// Reference behaviour replaced (file:line under /root/reference):
//   .find_neighbors          R/spatialCluster.R:227-302
//   find_neighbors           R/utils.R:13-26          (O(n^2) there; a cell grid here, same result)
//   convert_neighbors        src/cluster.cpp:134-145  (+ src/utils.h:10-62, src/neighbor.h:67-80)
//   find_subspot_neighbors   src/cluster.cpp:161-215
//   convert_neighbors2mtx    src/cluster.cpp:147-159
#include "common.hpp"

#include <atomic>

#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace bsb {

thread_local std::string g_last_error;
std::atomic<int> g_cancel{0};
bsb200_poll_fn g_poll = nullptr;
void *g_poll_user = nullptr;
bool cancel_requested() {
  if (g_poll && g_poll(g_poll_user) != 0) g_cancel.store(1);
  return g_cancel.load() != 0;
}

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

// Writes a vector-of-lists adjacency into caller CSR buffers with the two-pass sizing protocol.
static int emit_csr(const std::vector<std::vector<int32_t>> &adj, int64_t *ptr, int32_t *idx,
                    int64_t cap, int64_t *nnz_out) {
  int64_t nnz = 0;
  for (const auto &v : adj) nnz += (int64_t) v.size();
  if (nnz_out) *nnz_out = nnz;
  if (ptr) {
    int64_t p = 0;
    for (size_t j = 0; j < adj.size(); j++) {
      ptr[j] = p;
      p += (int64_t) adj[j].size();
    }
    ptr[adj.size()] = p;
  }
  if (!idx || cap < nnz) {
    if (idx || cap > 0) return fail(BSB200_ERR_BUFFER, "neighbour buffer too small: need %lld entries", (long long) nnz);
    return BSB200_OK;   // sizing pass
  }
  int64_t p = 0;
  for (const auto &v : adj)
    for (int32_t u : v) idx[p++] = u;
  return BSB200_OK;
}

int lattice_neighbors(const int32_t *row, const int32_t *col, int64_t n, int platform,
                      std::vector<std::vector<int32_t>> &adj) {
  static const int hx[6] = {-2, 2, -1, 1, -1, 1}, hy[6] = {0, 0, -1, -1, 1, 1};
  static const int sx[4] = {0, 1, 0, -1}, sy[4] = {-1, 0, 1, 0};
  if (platform != BSB200_PLATFORM_VISIUM && platform != BSB200_PLATFORM_SQUARE)
    return fail(BSB200_ERR_INVALID, ".find_neighbors: Unsupported platform \"%d\".", platform);
  const int no = platform == BSB200_PLATFORM_VISIUM ? 6 : 4;
  const int *ox = platform == BSB200_PLATFORM_VISIUM ? hx : sx;
  const int *oy = platform == BSB200_PLATFORM_VISIUM ? hy : sy;
  adj.assign((size_t) n, {});
  if (n == 0) return BSB200_OK;
  int32_t rmin = row[0], rmax = row[0], cmin = col[0], cmax = col[0];
  for (int64_t j = 1; j < n; j++) {
    rmin = std::min(rmin, row[j]); rmax = std::max(rmax, row[j]);
    cmin = std::min(cmin, col[j]); cmax = std::max(cmax, col[j]);
  }
  const int64_t H = (int64_t) rmax - rmin + 1, W = (int64_t) cmax - cmin + 1;
  bool dense = H * W <= 64 * n + 1024;
  std::vector<int32_t> grid;
  if (dense) {
    grid.assign((size_t) (H * W), -1);
    for (int64_t j = 0; j < n && dense; j++) {
      int32_t &g = grid[(size_t) ((row[j] - rmin) * W + (col[j] - cmin))];
      if (g >= 0) dense = false;   // duplicated coordinates: merge() semantics need the sorted path
      g = (int32_t) j;
    }
  }
  if (dense) {
    for (int64_t j = 0; j < n; j++) {
      auto &v = adj[(size_t) j];
      for (int o = 0; o < no; o++) {
        const int64_t r = (int64_t) row[j] + oy[o] - rmin, c = (int64_t) col[j] + ox[o] - cmin;
        if (r < 0 || r >= H || c < 0 || c >= W) continue;
        const int32_t u = grid[(size_t) (r * W + c)];
        if (u >= 0) v.push_back(u);
      }
      std::sort(v.begin(), v.end());
    }
    return BSB200_OK;
  }
  // sparse / duplicated coordinates: sorted (col,row) table + binary search, keeps every match
  std::vector<std::pair<std::pair<int32_t, int32_t>, int32_t>> tab((size_t) n);
  for (int64_t j = 0; j < n; j++) tab[(size_t) j] = {{col[j], row[j]}, (int32_t) j};
  std::sort(tab.begin(), tab.end());
  for (int64_t j = 0; j < n; j++) {
    auto &v = adj[(size_t) j];
    for (int o = 0; o < no; o++) {
      const std::pair<int32_t, int32_t> key{col[j] + ox[o], row[j] + oy[o]};
      auto it = std::lower_bound(tab.begin(), tab.end(), std::make_pair(key, (int32_t) -1));
      for (; it != tab.end() && it->first == key; ++it) v.push_back(it->second);
    }
    std::sort(v.begin(), v.end());
  }
  return BSB200_OK;
}

int radius_neighbors(const double *pos, int64_t n, double radius, int metric,
                     std::vector<std::vector<int32_t>> &adj) {
  if (metric != BSB200_METRIC_MANHATTAN && metric != BSB200_METRIC_EUCLIDEAN)
    return fail(BSB200_ERR_INVALID, "find_neighbors: 'arg' should be one of \"manhattan\", \"euclidean\"");
  adj.assign((size_t) n, {});
  if (n == 0 || !(radius > 0)) return BSB200_OK;
  const double *x = pos, *y = pos + n;
  double xmin = x[0], ymin = y[0];
  for (int64_t j = 1; j < n; j++) { xmin = std::min(xmin, x[j]); ymin = std::min(ymin, y[j]); }
  auto cell = [&](double v, double lo) { return (int64_t) std::floor((v - lo) / radius); };
  auto key = [](int64_t cx, int64_t cy) { return (uint64_t) cx * 0x9E3779B97F4A7C15ull ^ ((uint64_t) cy + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full; };
  std::unordered_map<uint64_t, std::vector<int32_t>> cells;
  cells.reserve((size_t) n);
  for (int64_t j = 0; j < n; j++) cells[key(cell(x[j], xmin), cell(y[j], ymin))].push_back((int32_t) j);
  for (int64_t i = 0; i < n; i++) {
    auto &v = adj[(size_t) i];
    const int64_t cx = cell(x[i], xmin), cy = cell(y[i], ymin);
    for (int64_t ax = cx - 1; ax <= cx + 1; ax++)
      for (int64_t ay = cy - 1; ay <= cy + 1; ay++) {
        auto it = cells.find(key(ax, ay));
        if (it == cells.end()) continue;
        for (int32_t j : it->second) {
          if (cell(x[j], xmin) != ax || cell(y[j], ymin) != ay) continue;   // hash collision guard
          const double dx = x[i] - x[j], dy = y[i] - y[j];
          const double dist = metric == BSB200_METRIC_MANHATTAN ? std::fabs(dx) + std::fabs(dy)
                                                                : std::sqrt(dx * dx + dy * dy);
          if (dist <= radius && dist > 0) v.push_back(j);
        }
      }
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
  }
  return BSB200_OK;
}

// "3,17,18" (1-based) -> {2,16,17}; NULL -> {}.  Empty tokens are skipped like str_split does.
int parse_neighbor_string(const char *s, std::vector<int32_t> &out) {
  out.clear();
  if (!s) return BSB200_OK;
  const char *p = s;
  while (*p) {
    if (*p == ',') { p++; continue; }
    if (*p < '0' || *p > '9') return fail(BSB200_ERR_PARSE, "Unconvertable value: %s", s);
    uint64_t v = 0;
    while (*p >= '0' && *p <= '9') {
      v = v * 10 + (uint64_t) (*p - '0');
      if (v > 0x7fffffffull) return fail(BSB200_ERR_PARSE, "Value out of range: %s", s);
      p++;
    }
    if (*p && *p != ',') return fail(BSB200_ERR_PARSE, "Unconvertable value: %s", s);
    out.push_back((int32_t) ((int64_t) v - 1));
  }
  return BSB200_OK;
}

int subspot_neighbors(const double *pos, int64_t n, int64_t n0, const int64_t *sptr,
                      const int32_t *sidx, double dist, std::vector<std::vector<int32_t>> &adj) {
  if (n0 <= 0 || n <= 0) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  const int64_t S = n / n0;
  if (std::fabs((double) n / (double) n0 - (double) S) > 1e-6) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  adj.assign((size_t) n, {});
  const double *x = pos, *y = pos + n;
  for (int64_t s = 0; s < n; s++) {
    const int64_t j0 = s % n0;
    auto &v = adj[(size_t) s];
    const int64_t dg = sptr[j0 + 1] - sptr[j0];
    for (int64_t pi = 0; pi <= dg; pi++) {
      const int64_t p = pi < dg ? sidx[sptr[j0] + pi] : j0;
      if (p < 0 || p >= n0) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
      for (int64_t r = 0; r < S; r++) {
        const int64_t c = r * n0 + p;
        const double m = std::fabs(x[c] - x[s]) + std::fabs(y[c] - y[s]);
        if (m < dist && m > 1e-6) v.push_back((int32_t) c);
      }
    }
    if (v.size() < 2) return fail(BSB200_ERR_NEIGHBORS, "Error in finding neighbors of subspots!");
  }
  return BSB200_OK;
}

// Colouring of the symmetrised graph (self loops ignored): same-colour spots never read each
// other's label, which is all the chromatic scan needs.  DSATUR (Brelaz 1979) with ties broken by
// degree, then index: exact on bipartite graphs (square lattices, subspot graphs: 2 colours) and it
// finds the 3-colouring of triangulated hex lattices.  Deterministic.  Above 2M vertices the heap is
// skipped in favour of plain greedy in index order (callers with lattices pass closed-form colours).
int colour_graph(const int64_t *ptr, const int32_t *idx, int64_t n, std::vector<int32_t> &colour,
                 int32_t *ncolours) {
  for (int64_t i = 0; i < n; i++)
    for (int64_t e = ptr[i]; e < ptr[i + 1]; e++) {
      const int32_t j = idx[e];
      if (j < 0 || j >= n) return fail(BSB200_ERR_INVALID, "neighbour index %d out of range", (int) j);
    }
  // reverse edges that are missing from the forward lists
  bool asym = false;
  std::vector<std::vector<int32_t>> extra;
  for (int64_t i = 0; i < n; i++)
    for (int64_t e = ptr[i]; e < ptr[i + 1]; e++) {
      const int32_t j = idx[e];
      if (j == i) continue;
      bool back = false;
      for (int64_t f = ptr[j]; f < ptr[j + 1] && !back; f++) back = idx[f] == i;
      if (!back) {
        if (!asym) { extra.assign((size_t) n, {}); asym = true; }
        extra[(size_t) j].push_back((int32_t) i);
      }
    }
  auto for_each_nb = [&](int64_t i, auto &&fn) {
    for (int64_t e = ptr[i]; e < ptr[i + 1]; e++) if (idx[e] != i) fn(idx[e]);
    if (asym) for (int32_t j : extra[(size_t) i]) fn(j);
  };
  colour.assign((size_t) n, -1);
  std::vector<uint64_t> mask((size_t) n, 0);   // colours seen among neighbours (first 64)
  int32_t nc = 0;
  auto lowest_free = [&](int64_t i) {
    int32_t c = 0;
    if (~mask[(size_t) i]) c = __builtin_ctzll(~mask[(size_t) i]);
    else {   // more than 64 colours around a vertex: exact scan
      std::vector<uint8_t> used((size_t) nc + 1, 0);
      for_each_nb(i, [&](int32_t j) { if (colour[(size_t) j] >= 0) used[(size_t) colour[(size_t) j]] = 1; });
      c = 64;
      while (c < nc && used[(size_t) c]) c++;
    }
    return c;
  };
  auto assign = [&](int64_t i, int32_t c) {
    colour[(size_t) i] = c;
    if (c >= nc) nc = c + 1;
  };
  if (n > 2000000) {
    for (int64_t i = 0; i < n; i++) {
      for_each_nb(i, [&](int32_t j) { if (colour[(size_t) j] >= 0 && colour[(size_t) j] < 64) mask[(size_t) i] |= 1ull << colour[(size_t) j]; });
      assign(i, lowest_free(i));
    }
  } else {
    std::vector<int32_t> deg((size_t) n, 0), sat((size_t) n, 0);
    for (int64_t i = 0; i < n; i++) for_each_nb(i, [&](int32_t) { deg[(size_t) i]++; });
    // max-heap on (saturation, degree, -index) with lazy invalidation
    typedef std::pair<std::pair<int32_t, int32_t>, int32_t> Item;
    std::vector<Item> heap;
    heap.reserve((size_t) n * 2);
    for (int64_t i = 0; i < n; i++) heap.push_back({{0, deg[(size_t) i]}, (int32_t) -i});
    std::make_heap(heap.begin(), heap.end());
    while (!heap.empty()) {
      std::pop_heap(heap.begin(), heap.end());
      const Item it = heap.back();
      heap.pop_back();
      const int64_t i = -it.second;
      if (colour[(size_t) i] >= 0 || it.first.first != sat[(size_t) i]) continue;   // stale entry
      const int32_t c = lowest_free(i);
      assign(i, c);
      for_each_nb(i, [&](int32_t j) {
        if (colour[(size_t) j] >= 0 || c >= 64) return;
        if (!(mask[(size_t) j] & (1ull << c))) {
          mask[(size_t) j] |= 1ull << c;
          sat[(size_t) j]++;
          heap.push_back({{sat[(size_t) j], deg[(size_t) j]}, (int32_t) -j});
          std::push_heap(heap.begin(), heap.end());
        }
      });
    }
  }
  if (ncolours) *ncolours = nc;
  return BSB200_OK;
}

// true when no directed edge joins two spots of one colour (self loops excepted)
static std::atomic<int> g_host_share{1};
void set_host_share(int ranks_on_this_host) { g_host_share.store(ranks_on_this_host < 1 ? 1 : ranks_on_this_host); }
int host_threads_max() {
  static const int nt = []() {
    int v = (int) std::thread::hardware_concurrency();
    if (const char *e = std::getenv("BSB200_HOST_THREADS")) v = std::atoi(e);
    return v < 1 ? 1 : (v > 16 ? 16 : v);
  }();
  return nt;
}
int host_threads() {
  const int nt = host_threads_max();
  // The ranks of a sharded chain share the host's cores, but their O(n) loops rarely coincide: dividing the threads
  // by the rank count made the set-up at 8 ranks slower (halo plan 15 -> 40 ms), so a rank keeps at least half of them.
  const int share = g_host_share.load(std::memory_order_relaxed);
  return share > 1 ? std::max(nt / 2, nt / share) : nt;
}

bool colouring_is_proper(const int64_t *ptr, const int32_t *idx, int64_t n, const int32_t *colour) {
  std::atomic<int> bad{0};
  parallel_for(n, [&](int64_t lo, int64_t hi, int) {
    for (int64_t i = lo; i < hi && !bad.load(std::memory_order_relaxed); i++)
      for (int64_t e = ptr[i]; e < ptr[i + 1]; e++)
        if (idx[e] != i && colour[idx[e]] == colour[i]) { bad.store(1); return; }
  });
  return bad.load() == 0;
}

}   // namespace bsb

using namespace bsb;

extern "C" {

int bsb200_version(void) { return BSB200_VERSION; }
const char *bsb200_last_error(void) { return g_last_error.c_str(); }
void bsb200_cancel(void) { g_cancel.store(1); }
void bsb200_set_interrupt_poll(bsb200_poll_fn poll, void *user) {
  bsb::g_poll_user = user;
  bsb::g_poll = poll;
}

int bsb200_neighbors_lattice(const int32_t *array_row, const int32_t *array_col, int64_t n,
                             int32_t platform, int64_t *ptr, int32_t *idx, int64_t cap, int64_t *nnz) {
  if (n < 0 || (n > 0 && (!array_row || !array_col))) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  if (platform != BSB200_PLATFORM_VISIUM && platform != BSB200_PLATFORM_SQUARE)
    return fail(BSB200_ERR_INVALID, ".find_neighbors: Unsupported platform \"%d\".", platform);
  // Fast path for compact lattices without duplicated coordinates (the Visium-HD / 10M-spot case):
  // a dense (row, col) -> spot table and direct CSR emission, no per-spot allocations.
  if (n > 0) {
    static const int hx[6] = {-2, 2, -1, 1, -1, 1}, hy[6] = {0, 0, -1, -1, 1, 1};
    static const int sx[4] = {0, 1, 0, -1}, sy[4] = {-1, 0, 1, 0};
    const int no = platform == BSB200_PLATFORM_VISIUM ? 6 : 4;
    const int *ox = platform == BSB200_PLATFORM_VISIUM ? hx : sx, *oy = platform == BSB200_PLATFORM_VISIUM ? hy : sy;
    int32_t rmin = array_row[0], rmax = rmin, cmin = array_col[0], cmax = cmin;
    for (int64_t j = 1; j < n; j++) {
      rmin = std::min(rmin, array_row[j]); rmax = std::max(rmax, array_row[j]);
      cmin = std::min(cmin, array_col[j]); cmax = std::max(cmax, array_col[j]);
    }
    const int64_t H = (int64_t) rmax - rmin + 1, W = (int64_t) cmax - cmin + 1;
    bool dense = H * W <= 64 * n + 1024;
    std::vector<int32_t> grid;
    if (dense) {
      grid.assign((size_t) (H * W), -1);
      for (int64_t j = 0; j < n && dense; j++) {
        int32_t &g = grid[(size_t) ((array_row[j] - rmin) * W + (array_col[j] - cmin))];
        if (g >= 0) dense = false;
        g = (int32_t) j;
      }
    }
    if (dense) {
      const bool fill = idx && ptr;
      int64_t total = 0;
      for (int64_t j = 0; j < n; j++) {
        int32_t nb[6];
        int cnt = 0;
        for (int o = 0; o < no; o++) {
          const int64_t r = (int64_t) array_row[j] + oy[o] - rmin, c = (int64_t) array_col[j] + ox[o] - cmin;
          if (r < 0 || r >= H || c < 0 || c >= W) continue;
          const int32_t u = grid[(size_t) (r * W + c)];
          if (u >= 0) nb[cnt++] = u;
        }
        for (int a = 1; a < cnt; a++) {   // at most 6 entries: insertion sort (ascending ids, R/spatialCluster.R:263-267)
          const int32_t v = nb[a];
          int b = a - 1;
          for (; b >= 0 && nb[b] > v; b--) nb[b + 1] = nb[b];
          nb[b + 1] = v;
        }
        if (ptr) ptr[j] = total;
        if (fill && total + cnt <= cap)
          for (int e = 0; e < cnt; e++) idx[total + e] = nb[e];
        total += cnt;
      }
      if (ptr) ptr[n] = total;
      if (nnz) *nnz = total;
      if ((idx || cap > 0) && cap < total) return fail(BSB200_ERR_BUFFER, "neighbour buffer too small: need %lld entries", (long long) total);
      return BSB200_OK;
    }
  }
  std::vector<std::vector<int32_t>> adj;
  int rc = lattice_neighbors(array_row, array_col, n, platform, adj);
  return rc ? rc : emit_csr(adj, ptr, idx, cap, nnz);
}

int bsb200_neighbors_radius(const double *positions, int64_t n, double radius, int32_t metric,
                            int64_t *ptr, int32_t *idx, int64_t cap, int64_t *nnz) {
  if (n < 0 || (n > 0 && !positions)) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  std::vector<std::vector<int32_t>> adj;
  int rc = radius_neighbors(positions, n, radius, metric, adj);
  return rc ? rc : emit_csr(adj, ptr, idx, cap, nnz);
}

int bsb200_parse_neighbor_strings(const char *const *strings, int64_t n0, int64_t *ptr, int32_t *idx,
                                  int64_t cap, int64_t *nnz) {
  if (n0 < 0 || (n0 > 0 && !strings)) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  std::vector<std::vector<int32_t>> adj((size_t) n0);
  for (int64_t j = 0; j < n0; j++) {
    int rc = parse_neighbor_string(strings[j], adj[(size_t) j]);
    if (rc) return rc;
  }
  return emit_csr(adj, ptr, idx, cap, nnz);
}

int bsb200_subspot_neighbors(const double *positions, int64_t n, int64_t n0, const int64_t *spot_ptr,
                             const int32_t *spot_idx, double dist, int64_t *ptr, int32_t *idx,
                             int64_t cap, int64_t *nnz) {
  if (!positions || !spot_ptr) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  std::vector<std::vector<int32_t>> adj;
  int rc = subspot_neighbors(positions, n, n0, spot_ptr, spot_idx, dist, adj);
  return rc ? rc : emit_csr(adj, ptr, idx, cap, nnz);
}

int bsb200_neighbors_to_matrix4(const int64_t *ptr, const int32_t *idx, int64_t n, int64_t *out) {
  if (!ptr || !out) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  std::memset(out, 0, sizeof(int64_t) * 4 * (size_t) n);
  for (int64_t j = 0; j < n; j++) {
    const int64_t dg = ptr[j + 1] - ptr[j];
    if (dg > 4) return fail(BSB200_ERR_UNSUPPORTED, "convert_neighbors2mtx: spot %lld has %lld > 4 neighbours", (long long) j, (long long) dg);
    for (int64_t e = 0; e < dg; e++) out[e * n + j] = (int64_t) idx[ptr[j] + e] + 1;
  }
  return BSB200_OK;
}

int bsb200_colour_graph(const int64_t *ptr, const int32_t *idx, int64_t n, int32_t *colour,
                        int32_t *ncolours) {
  if (!ptr || !colour || n < 0) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  std::vector<int32_t> c;
  int rc = colour_graph(ptr, idx, n, c, ncolours);
  if (rc) return rc;
  std::memcpy(colour, c.data(), sizeof(int32_t) * (size_t) n);
  return BSB200_OK;
}

int bsb200_colour_lattice(const int32_t *array_row, const int32_t *array_col, int64_t n,
                          int32_t platform, int32_t *colour, int32_t *ncolours) {
  if (n < 0 || (n > 0 && (!array_row || !array_col || !colour))) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  if (platform == BSB200_PLATFORM_SQUARE) {
    for (int64_t j = 0; j < n; j++) colour[j] = (int32_t) (((int64_t) array_row[j] + array_col[j]) & 1);
    if (ncolours) *ncolours = n > 0 ? 2 : 0;
    return BSB200_OK;
  }
  if (platform == BSB200_PLATFORM_VISIUM) {
    for (int64_t j = 0; j < n; j++) {
      const int64_t t = (int64_t) array_col[j] - 3 * (int64_t) array_row[j];
      if (t & 1) return fail(BSB200_ERR_INVALID, "hex lattice needs array_col == array_row (mod 2)");
      colour[j] = (int32_t) ((((t / 2) % 3) + 3) % 3);
    }
    if (ncolours) *ncolours = n > 0 ? 3 : 0;
    return BSB200_OK;
  }
  return fail(BSB200_ERR_INVALID, ".find_neighbors: Unsupported platform \"%d\".", platform);
}

}   // extern "C"
