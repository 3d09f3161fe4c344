// comm.hpp -- the two exchanges of a spot-sharded chain over NVLink 5 / NVSwitch.
//
// Two transports, chosen by the communicator id every rank receives (bsb200_comm_unique_id):
//
//   PEER (default)  one process per GPU of ONE box.  Every rank exports its label array, its shard-sum window and a
//                   flag page with CUDA IPC; the peers map them, and the exchanges become plain stores into peer
//                   memory from our own kernels, ordered by epoch flags:
//                     * halo: after a colour phase one small kernel stores the labels of the boundary spots straight
//                       into the neighbours' ghost slots, releases a flag, and waits for the neighbours' flags;
//                     * shard sums: one kernel stores the rank's G/W shard sums into every peer's window (double
//                       buffered by epoch parity), releases a flag per peer and waits for theirs.
//                   No NCCL, no staging buffers, no communicator start-up: the host rendezvous is a POSIX shared
//                   memory segment named after the id.  Synchronisation is neighbour-to-neighbour for the halo (a
//                   rank never waits for a GPU it shares no boundary with) and all-to-all once per iteration.
//   NCCL            (BSB200_COMM=nccl when the id is made)  pack -> grouped ncclSend/ncclRecv -> unpack and
//                   ncclAllGather; bound at run time (dlopen of libnccl.so.2), kept as the baseline the peer
//                   transport is measured against and for setups without peer access.
//
// Both are enqueued on the chain's stream and are captured into its CUDA graph with the kernels around them.
// Every spin wait is bounded (BSB_PEER_TIMEOUT_NS): a lost peer turns into BSB200_ERR_CUDA, never into a hang.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

namespace bsb {

constexpr int kMaxPeerWorld = 16;                       // ranks of one box
constexpr int BSB_DEVERR_PEER_TIMEOUT = 7;              // device error code: a peer's flag never arrived
constexpr unsigned long long BSB_PEER_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;

struct HaloPlan;
class ShmRendezvous;

// one (colour, peer) leg of the halo exchange, resident in device memory
struct HaloLeg {
  uint8_t *peer_z;            // peer's label array (mapped)
  uint32_t *peer_flag;        // peer's halo flag for this rank
  const uint32_t *my_flag;    // this rank's halo flag for the peer
  int32_t s0, s1;             // range in send_pos / send_dst
  int32_t nrecv;              // > 0: wait for the peer's labels of this colour
  int32_t peer;
};
struct GatherLeg {
  double *peer_gsum;          // peer's shard-sum window (both parities)
  uint32_t *peer_flag;
  const uint32_t *my_flag;
  int32_t peer, pad;
};

struct Comm {
  enum Transport { NONE = 0, NCCL = 1, PEER = 2 };
  Transport transport = NONE;
  int world = 1, rank = 0;
  // ---- NCCL
  void *nccl = nullptr;   // ncclComm_t
  // ---- PEER
  ShmRendezvous *shm = nullptr;
  // ONE exported allocation per rank: [flags page | shard sums parity 0 | parity 1 | receive list | labels].
  // (A single handle per rank: small cudaMalloc blocks may share an underlying allocation, which can be opened
  // only once per process and whose handle names its base, so the arena's offset in it travels with the handle.)
  uint8_t *arena = nullptr;
  size_t arena_bytes = 0, gsum_stride = 0, off_recv = 0, off_z = 0, recv_count = 0, z_bytes = 0;
  void *peer_base[kMaxPeerWorld] = {};      // what cudaIpcOpenMemHandle returned (closed at shutdown)
  uint8_t *peer_arena[kMaxPeerWorld] = {};  // the peer's arena inside that mapping
  uint64_t peer_timeout_ns = BSB_PEER_TIMEOUT_NS;
  int32_t *send_dst = nullptr;          // for every send entry: position in the peer's label array
  HaloLeg *halo_legs = nullptr;         // [ncol][world - 1]
  GatherLeg *gather_legs = nullptr;     // [world - 1]
  uint32_t *counters = nullptr;         // sent[world], waited[world], gather epoch, gather done-count
  std::vector<int> halo_leg_count;      // legs per colour
  int ncol = 0;

  ~Comm();
  int init(int world, int rank, const void *unique_id128);
  bool peer() const { return transport == PEER; }

  // host-side allgather of `bytes` per rank (set-up data: handles, offsets, column sums)
  int host_allgather(const void *mine, size_t bytes, void *all);

  // PEER: the exported arena (labels: z_bytes, receive list: recv_count int32, shard sums: gsum_doubles = G * Emax
  // per parity), zeroed.  The chain places its label array and receive list in it, then calls connect(), which
  // exchanges the handles, maps the peers and builds the leg tables.
  int alloc_arena(size_t z_bytes, size_t recv_count, size_t gsum_doubles, cudaStream_t s);
  int connect(const HaloPlan &halo, int ncol, cudaStream_t s);
  double *gsum() const { return reinterpret_cast<double *>(arena + 4096); }
  int32_t *recv_pos() const { return reinterpret_cast<int32_t *>(arena + off_recv); }
  uint8_t *z() const { return arena + off_z; }
  const uint32_t *gather_epoch() const { return counters + 2 * world; }

  // PEER data path
  void enqueue_halo_peer(int colour, const uint8_t *z, const int32_t *send_pos_dev, int *err, cudaStream_t s);
  void enqueue_gather_peer(size_t offset_doubles, size_t count_doubles, int *err, cudaStream_t s);
  // fused-halo mode: a colour phase in which this rank sweeps nothing still owes its neighbours the phase's epoch
  void enqueue_halo_release(int colour, const uint32_t *iter, cudaStream_t s);

  // NCCL data path.  in-place allgather: every rank contributes `count` doubles at buf + rank*count
  int allgather_f64(double *buf, size_t count, cudaStream_t s);
  // grouped send/recv of bytes with every peer that has a non-empty list
  int exchange_u8(const uint8_t *sendbuf, const int32_t *send_off, uint8_t *recvbuf, const int32_t *recv_off,
                  cudaStream_t s);   // offsets: world + 1 entries each (host)
  void shutdown();   // unmap the peers' buffers, meet the peers, release the window
};

int comm_unique_id(void *out128);

}   // namespace bsb
