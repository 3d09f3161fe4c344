// shard.cpp -- host-side planning of the spot-sharded chain (see shard.hpp).  Pure C++.
#include "shard.hpp"

#include <algorithm>

#include "common.hpp"

namespace bsb {

int ShardPlan::owner(int64_t j) const { return group_of(j) / (groups / world); }

int make_shard_plan(int64_t n, const int64_t *ptr, const int32_t *idx, int groups, int world, int rank,
                    ShardPlan &plan) {
  if (world < 1 || rank < 0 || rank >= world) return fail(BSB200_ERR_INVALID, "rank %d / world %d invalid", rank, world);
  if (groups < 1 || groups % world) return fail(BSB200_ERR_INVALID, "world size %d must divide the %d virtual shards", world, groups);
  if (n < groups) return fail(BSB200_ERR_INVALID, "n = %lld is too small to shard", (long long) n);
  plan.world = world; plan.rank = rank; plan.groups = groups; plan.n = n;
  plan.g0 = rank * (groups / world);
  plan.g1 = (rank + 1) * (groups / world);
  plan.group_lo.assign((size_t) groups + 1, 0);
  for (int g = 0; g <= groups; g++) plan.group_lo[(size_t) g] = ((int64_t) g * n + groups - 1) / groups;   // first j with j*G/n >= g
  plan.lo = plan.group_lo[(size_t) plan.g0];
  plan.hi = plan.group_lo[(size_t) plan.g1];
  plan.ghosts.clear();
  if (world > 1) {
    const int nchunk = parallel_chunks(plan.hi - plan.lo);
    std::vector<std::vector<int32_t>> part((size_t) nchunk);
    parallel_for(plan.hi - plan.lo, [&](int64_t a, int64_t b, int t) {
      for (int64_t j = plan.lo + a; j < plan.lo + b; j++)
        for (int64_t e = ptr[j]; e < ptr[j + 1]; e++)
          if (idx[e] < plan.lo || idx[e] >= plan.hi) part[(size_t) t].push_back(idx[e]);
    });
    for (auto &v : part) plan.ghosts.insert(plan.ghosts.end(), v.begin(), v.end());
    std::sort(plan.ghosts.begin(), plan.ghosts.end());
    plan.ghosts.erase(std::unique(plan.ghosts.begin(), plan.ghosts.end()), plan.ghosts.end());
  }
  return BSB200_OK;
}

int make_halo_plan(const ShardPlan &plan, const int64_t *ptr, const int32_t *idx, const int32_t *colour, int ncol,
                   const std::vector<int32_t> &own_inv, int64_t ghost_base, HaloPlan &halo) {
  const int W = plan.world;
  std::vector<std::vector<int32_t>> send((size_t) ncol * W), recv((size_t) ncol * W);
  // what I receive: my ghosts, by (colour, owner), ascending original id (ghosts are sorted)
  for (size_t gi = 0; gi < plan.ghosts.size(); gi++) {
    const int32_t u = plan.ghosts[gi];
    recv[(size_t) colour[u] * W + plan.owner(u)].push_back((int32_t) (ghost_base + (int64_t) gi));
  }
  // what I send: own spots read by a spot of another rank
  // (every spot of the other ranks is looked at: the lists need not be symmetric; chunks on host threads)
  std::vector<std::vector<int32_t>> need((size_t) W);
  {
    const int nchunk = parallel_chunks(plan.n);
    std::vector<std::vector<std::vector<int32_t>>> part((size_t) nchunk, std::vector<std::vector<int32_t>>((size_t) W));
    parallel_for(plan.n, [&](int64_t a, int64_t b, int t) {
      for (int64_t v = a; v < b; v++) {
        if (v >= plan.lo && v < plan.hi) { v = plan.hi - 1; continue; }
        const int p = plan.owner(v);
        for (int64_t e = ptr[v]; e < ptr[v + 1]; e++)
          if (idx[e] >= plan.lo && idx[e] < plan.hi) part[(size_t) t][(size_t) p].push_back(idx[e]);
      }
    });
    for (int t = 0; t < nchunk; t++)
      for (int p = 0; p < W; p++) need[(size_t) p].insert(need[(size_t) p].end(), part[(size_t) t][(size_t) p].begin(), part[(size_t) t][(size_t) p].end());
  }
  for (int p = 0; p < W; p++) {
    auto &v = need[(size_t) p];
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    for (int32_t u : v) send[(size_t) colour[u] * W + p].push_back(own_inv[(size_t) (u - plan.lo)]);
  }
  halo.send_pos.clear(); halo.recv_pos.clear();
  halo.send_off.assign((size_t) ncol * W + 1, 0);
  halo.recv_off.assign((size_t) ncol * W + 1, 0);
  for (size_t k = 0; k < (size_t) ncol * W; k++) {
    halo.send_pos.insert(halo.send_pos.end(), send[k].begin(), send[k].end());
    halo.recv_pos.insert(halo.recv_pos.end(), recv[k].begin(), recv[k].end());
    halo.send_off[k + 1] = (int32_t) halo.send_pos.size();
    halo.recv_off[k + 1] = (int32_t) halo.recv_pos.size();
  }
  return BSB200_OK;
}

}   // namespace bsb

using namespace bsb;

// Host-only view of the plan, for tests of the multi-rank logic (no device needed).
// counts: ncol*world send counts then ncol*world receive counts.  ghosts: capacity cap entries.
extern "C" int bsb200_shard_plan(const int64_t *ptr, const int32_t *idx, int64_t n, const int32_t *colour,
                                 int32_t ncolours, int32_t groups, int32_t world, int32_t rank, int64_t *lo,
                                 int64_t *hi, int32_t *ghosts, int64_t cap, int64_t *nghosts, int32_t *counts,
                                 int32_t *send_ids, int32_t *recv_ids) {
  if (!ptr || !colour || !lo || !hi) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  ShardPlan plan;
  int rc = make_shard_plan(n, ptr, idx, groups, world, rank, plan);
  if (rc) return rc;
  *lo = plan.lo; *hi = plan.hi;
  if (nghosts) *nghosts = (int64_t) plan.ghosts.size();
  if (ghosts) {
    if (cap < (int64_t) plan.ghosts.size()) return fail(BSB200_ERR_BUFFER, "ghost buffer too small");
    std::copy(plan.ghosts.begin(), plan.ghosts.end(), ghosts);
  }
  if (counts || send_ids || recv_ids) {
    std::vector<int32_t> own_inv((size_t) plan.n_own());
    for (int64_t j = 0; j < plan.n_own(); j++) own_inv[(size_t) j] = (int32_t) j;   // identity: positions = original id - lo
    HaloPlan halo;
    rc = make_halo_plan(plan, ptr, idx, colour, ncolours, own_inv, plan.n_own(), halo);
    if (rc) return rc;
    const int K = ncolours * world;
    if (counts)
      for (int k = 0; k < K; k++) {
        counts[k] = halo.send_off[(size_t) k + 1] - halo.send_off[(size_t) k];
        counts[K + k] = halo.recv_off[(size_t) k + 1] - halo.recv_off[(size_t) k];
      }
    // original ids in list order (what the other side must agree with)
    if (send_ids) for (size_t i = 0; i < halo.send_pos.size(); i++) send_ids[i] = (int32_t) (halo.send_pos[i] + plan.lo);
    if (recv_ids) for (size_t i = 0; i < halo.recv_pos.size(); i++) recv_ids[i] = plan.ghosts[(size_t) (halo.recv_pos[i] - plan.n_own())];
  }
  return BSB200_OK;
}
