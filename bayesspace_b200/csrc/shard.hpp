// shard.hpp -- spot sharding of ONE chain over several GPUs (one process per GPU), SURVEY.md §8e.
//
// The spots are cut into kGroups fixed "virtual shards" = contiguous ranges of original spot ids (row
// strips of a row-major lattice).  Rank r of W owns shards [r*G/W, (r+1)*G/W).  Everything that enters
// the chain's arithmetic is defined on the virtual shards, never on the ranks:
//   * statistics: per-segment records -> per-shard sums (fixed order) -> allgather of the G shard
//     sums -> every rank folds the G sums in the same order  => identical bits for W = 1, 2, 4, 8;
//   * randomness: Philox counters are original spot ids;
//   * column means of Y: per-shard sums, same fold.
// What crosses ranks on the data path: after every colour phase the labels of the boundary spots of
// that colour (bytes), and once per iteration the allgather of G*E doubles.
#pragma once
#include <cstdint>
#include <vector>

namespace bsb {

struct HaloPlan {
  // per colour c and peer p: positions are LOCAL (own spots in internal order, then ghosts)
  std::vector<int32_t> send_pos, recv_pos;          // concatenated over (colour, peer)
  std::vector<int32_t> send_off, recv_off;          // (ncol * world + 1) offsets into send_pos / recv_pos
  int64_t send_total() const { return (int64_t) send_pos.size(); }
  int64_t recv_total() const { return (int64_t) recv_pos.size(); }
};

struct ShardPlan {
  int world = 1, rank = 0, groups = 1, g0 = 0, g1 = 1;
  int64_t n = 0, lo = 0, hi = 0;                    // own original ids [lo, hi)
  std::vector<int64_t> group_lo;                    // groups + 1 original-id boundaries of the virtual shards
  std::vector<int32_t> ghosts;                      // original ids of the ghost spots, ascending
  int64_t n_own() const { return hi - lo; }
  int64_t n_local() const { return n_own() + (int64_t) ghosts.size(); }
  int owner(int64_t j) const;                       // rank that owns original spot j
  int group_of(int64_t j) const { return (int) ((j * (int64_t) groups) / n); }
};

// Range of the rank's own spots and the ghost list (neighbours of own spots that live elsewhere).
int make_shard_plan(int64_t n, const int64_t *ptr, const int32_t *idx, int groups, int world, int rank,
                    ShardPlan &plan);

// Send / receive lists per (colour, peer), both ordered by original id so that the two sides agree
// without negotiation.  local_pos(j) maps an original id (own or ghost) to its local position.
int make_halo_plan(const ShardPlan &plan, const int64_t *ptr, const int32_t *idx, const int32_t *colour, int ncol,
                   const std::vector<int32_t> &own_inv /* original id - lo -> internal position */,
                   int64_t ghost_base /* local position of the first ghost */, HaloPlan &halo);

}   // namespace bsb
