// comm.cpp -- the transports of the spot-sharded chain (see comm.hpp): peer-memory stores with epoch flags
// (CUDA IPC, host rendezvous in POSIX shared memory) and the run-time-bound NCCL baseline.
#include "comm.hpp"
#include "peer_sync.cuh"

#include <dlfcn.h>
#include <fcntl.h>
#include <nccl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <new>

#include "common.hpp"
#include "shard.hpp"

namespace bsb {
namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *api() {
  static NcclApi a;
  static bool tried = false;
  if (!tried) {
    tried = true;
    a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (a.handle) {
#define BIND(field, sym) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, sym))
      BIND(GetUniqueId, "ncclGetUniqueId");
      BIND(CommInitRank, "ncclCommInitRank");
      BIND(CommDestroy, "ncclCommDestroy");
      BIND(GroupStart, "ncclGroupStart");
      BIND(GroupEnd, "ncclGroupEnd");
      BIND(Send, "ncclSend");
      BIND(Recv, "ncclRecv");
      BIND(AllGather, "ncclAllGather");
      BIND(AllReduce, "ncclAllReduce");
      BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
      if (!a.GetUniqueId || !a.CommInitRank || !a.Send || !a.Recv || !a.AllGather || !a.GroupStart || !a.GroupEnd) a.handle = nullptr;
    }
  }
  return a.handle ? &a : nullptr;
}

int nccl_fail(ncclResult_t r, const char *what) {
  NcclApi *a = api();
  return fail(BSB200_ERR_CUDA, "NCCL %s failed: %s", what, a && a->GetErrorString ? a->GetErrorString(r) : "?");
}

}   // namespace

#define NCK(call, what)                      \
  do {                                       \
    ncclResult_t r_ = (call);                \
    if (r_ != ncclSuccess) return nccl_fail(r_, what); \
  } while (0)


static const char kPeerMagic[8] = {'B', 'S', 'B', 'P', 'E', 'E', 'R', '1'};

int comm_unique_id(void *out128) {
  const char *want = std::getenv("BSB200_COMM");
  if (want && std::strcmp(want, "nccl") == 0) {
    NcclApi *a = api();
    if (!a) return fail(BSB200_ERR_CUDA, "libnccl.so.2 could not be loaded (BSB200_COMM=nccl)");
    ncclUniqueId id;
    NCK(a->GetUniqueId(&id), "GetUniqueId");
    static_assert(NCCL_UNIQUE_ID_BYTES == 128, "bsb200_comm_unique_id hands out 128 bytes");
    std::memcpy(out128, &id, NCCL_UNIQUE_ID_BYTES);
    return BSB200_OK;
  }
  // peer transport: magic + 16 random bytes that name the rendezvous segment
  unsigned char id[128] = {};
  std::memcpy(id, kPeerMagic, 8);
  bool ok = false;
  if (FILE *f = std::fopen("/dev/urandom", "rb")) {
    ok = std::fread(id + 8, 1, 16, f) == 16;
    std::fclose(f);
  }
  if (!ok) {
    const uint64_t t = (uint64_t) std::chrono::steady_clock::now().time_since_epoch().count();
    const uint64_t p = (uint64_t) getpid() * 0x9E3779B97F4A7C15ull ^ (uint64_t) (uintptr_t) &id;
    std::memcpy(id + 8, &t, 8);
    std::memcpy(id + 16, &p, 8);
  }
  std::memcpy(out128, id, 128);
  return BSB200_OK;
}

// ---- host rendezvous: one POSIX shared-memory segment per communicator ------------------------------------------
// W slots of kSlot bytes and two sequence counters per rank.  A fresh segment is all zeros, which is the valid
// initial state, so every rank may create it.  allgather round r: write the slot, publish seq_in = r, read the
// others once their seq_in reaches r, publish seq_out = r; a slot is rewritten only when every seq_out reached r.
class ShmRendezvous {
 public:
  static constexpr size_t kSlot = 64 * 1024;
  struct Header {
    std::atomic<uint32_t> seq_in[kMaxPeerWorld];
    std::atomic<uint32_t> seq_out[kMaxPeerWorld];
  };
  ~ShmRendezvous() {
    if (base) munmap(base, bytes);
    if (rank == 0 && !unlinked) shm_unlink(name);
  }
  int open(const unsigned char *id16, int world_, int rank_) {
    world = world_; rank = rank_;
    static const char *hex = "0123456789abcdef";
    std::snprintf(name, sizeof(name), "/bsb200-");
    size_t k = std::strlen(name);
    for (int i = 0; i < 16; i++) { name[k++] = hex[id16[i] >> 4]; name[k++] = hex[id16[i] & 15]; }
    name[k] = 0;
    bytes = 4096 + (size_t) world * kSlot;
    const int fd = shm_open(name, O_CREAT | O_RDWR, 0600);
    if (fd < 0) return fail(BSB200_ERR_CUDA, "shm_open(%s) failed: %s", name, std::strerror(errno));
    if (ftruncate(fd, (off_t) bytes) != 0) { close(fd); return fail(BSB200_ERR_CUDA, "ftruncate(%s) failed: %s", name, std::strerror(errno)); }
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return fail(BSB200_ERR_CUDA, "mmap(%s) failed: %s", name, std::strerror(errno));
    base = (unsigned char *) p;
    hdr = reinterpret_cast<Header *>(base);
    static_assert(sizeof(Header) <= 4096, "header page");
    return BSB200_OK;
  }
  static double default_timeout() {
    const char *e = std::getenv("BSB200_RENDEZVOUS_TIMEOUT");   // seconds a rank waits for the others on the host
    const double v = e ? std::atof(e) : 0.0;
    return v > 0.0 ? v : 120.0;
  }
  int allgather(const void *mine, size_t nbytes, void *all, double timeout_s = default_timeout()) {
    if (nbytes > kSlot) return fail(BSB200_ERR_INVALID, "rendezvous message of %zu bytes exceeds the slot", nbytes);
    const uint32_t r = ++round;
    int rc;
    for (int p = 0; p < world; p++)   // everyone is done reading the previous round
      if ((rc = wait(hdr->seq_out[p], r - 1, p, timeout_s))) return rc;
    std::memcpy(base + 4096 + (size_t) rank * kSlot, mine, nbytes);
    hdr->seq_in[rank].store(r, std::memory_order_release);
    for (int p = 0; p < world; p++) {
      if ((rc = wait(hdr->seq_in[p], r, p, timeout_s))) return rc;
      std::memcpy((unsigned char *) all + (size_t) p * nbytes, base + 4096 + (size_t) p * kSlot, nbytes);
    }
    hdr->seq_out[rank].store(r, std::memory_order_release);
    if (rank == 0 && !unlinked) {   // every rank has the segment mapped: the name can go
      shm_unlink(name);
      unlinked = true;
    }
    return BSB200_OK;
  }

 private:
  int wait(std::atomic<uint32_t> &a, uint32_t target, int peer, double timeout_s) {
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spin = 0; (int32_t) (a.load(std::memory_order_acquire) - target) < 0; spin++) {
      if (spin < 2000) continue;
      if ((spin & 63) == 0) {
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s)
          return fail(BSB200_ERR_CUDA, "rank %d waited %.0f s for rank %d at the rendezvous (different ids, or a rank died?)",
                      rank, timeout_s, peer);
        usleep(50);
      } else {
        sched_yield();
      }
    }
    return BSB200_OK;
  }
  char name[64] = {};
  unsigned char *base = nullptr;
  Header *hdr = nullptr;
  size_t bytes = 0;
  int world = 1, rank = 0;
  uint32_t round = 0;
  bool unlinked = false;
};

// ---- device side of the peer transport ----------------------------------------------------------------------------
namespace {

// One CTA per neighbouring rank: store the boundary labels of the colour just swept into the neighbour's ghost
// slots, release the neighbour's flag, then wait until the neighbour's own labels of this colour have landed here.
// WAR safety: the slots written belong to the colour just swept, which the neighbour reads only in LATER phases,
// and a rank can be at most one phase ahead of a neighbour because each phase ends with this wait.  The handshake
// is SYMMETRIC: two ranks that exchange labels in ANY colour release and wait on each other in EVERY colour phase,
// also when one direction (or both) carries no labels for that colour -- with one-sided legs (irregular graphs,
// DSATUR colourings) the sender would otherwise run ahead and overwrite ghost labels its neighbour still reads.
__global__ void __launch_bounds__(256) k_halo_peer(const HaloLeg *legs, const uint8_t *z, const int32_t *send_pos,
                                                   const int32_t *send_dst, uint32_t *counters, int world, int *err,
                                                   unsigned long long timeout_ns) {
  const HaloLeg L = legs[blockIdx.x];
  for (int k = L.s0 + (int) threadIdx.x; k < L.s1; k += 256) L.peer_z[send_dst[k]] = z[send_pos[k]];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t e = counters[L.peer] + 1u;   // one epoch per colour phase and neighbour, in both directions
    counters[L.peer] = e;
    st_release_sys(L.peer_flag, e);
    wait_flag(L.my_flag, e, err, timeout_ns);
  }
}

// Fused-halo mode (k_sweep_fast pushes and waits itself, layout.hpp HaloFuse): epoch of an empty colour phase
__global__ void k_halo_release(const HaloLeg *legs, int nlegs, const uint32_t *iter, int ncol, int colour) {
  const int l = threadIdx.x;
  if (l < nlegs) st_release_sys(legs[l].peer_flag, (*iter - 1u) * (uint32_t) ncol + (uint32_t) colour + 1u);
}

// One CTA per peer: store this rank's shard sums into the peer's window (parity of the new epoch), release, wait
// for the peer's sums.  The last CTA to finish publishes the epoch, which k_param reads as its parity.
__global__ void __launch_bounds__(256) k_gather_peer(const GatherLeg *legs, const double *my_gsum, size_t stride,
                                                     size_t offset, size_t count, uint32_t *counters, int world,
                                                     int *err, unsigned long long timeout_ns) {
  const GatherLeg L = legs[blockIdx.x];
  const uint32_t e = counters[2 * world] + 1u;
  const size_t base = (size_t) (e & 1u) * stride + offset;
  const double *src = my_gsum + base;
  double *dst = L.peer_gsum + base;
  for (size_t i = threadIdx.x; i < count; i += 256) dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    st_release_sys(L.peer_flag, e);
    wait_flag(L.my_flag, e, err, timeout_ns);
    __threadfence();
    const uint32_t done = atomicAdd(&counters[2 * world + 1], 1u);
    if (done == gridDim.x - 1) {
      counters[2 * world + 1] = 0u;
      counters[2 * world] = e;
    }
  }
}

}   // namespace

#define CKC(call)                                                                                              \
  do {                                                                                                         \
    cudaError_t e_ = (call);                                                                                   \
    if (e_ != cudaSuccess)                                                                                     \
      return fail(BSB200_ERR_CUDA, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__);      \
  } while (0)

Comm::~Comm() { shutdown(); }

void Comm::shutdown() {
  if (transport == NCCL) {
    NcclApi *a = api();
    if (nccl && a && a->CommDestroy) a->CommDestroy((ncclComm_t) nccl);
    nccl = nullptr;
  }
  if (transport == PEER) {
    bool mapped = false;
    for (int p = 0; p < world; p++) {
      if (peer_base[p]) { cudaIpcCloseMemHandle(peer_base[p]); mapped = true; }
      peer_base[p] = nullptr; peer_arena[p] = nullptr;
    }
    if (shm && mapped) {   // nobody frees exported memory while a peer still has it mapped
      char one = 1, all[kMaxPeerWorld];
      shm->allgather(&one, 1, all, 10.0);
    }
    if (arena) cudaFree(arena);
    if (send_dst) cudaFree(send_dst);
    if (halo_legs) cudaFree(halo_legs);
    if (gather_legs) cudaFree(gather_legs);
    if (counters) cudaFree(counters);
    arena = nullptr; send_dst = nullptr; halo_legs = nullptr; gather_legs = nullptr; counters = nullptr;
    delete shm;
    shm = nullptr;
  }
  transport = NONE;
}

int Comm::init(int world_, int rank_, const void *unique_id128) {
  if (!unique_id128) return fail(BSB200_ERR_INVALID, "world > 1 needs the communicator id of bsb200_comm_unique_id()");
  world = world_; rank = rank_;
  const unsigned char *id = (const unsigned char *) unique_id128;
  if (std::memcmp(id, kPeerMagic, 8) == 0) {
    if (world > kMaxPeerWorld) return fail(BSB200_ERR_UNSUPPORTED, "the peer transport serves up to %d ranks of one box", kMaxPeerWorld);
    if (const char *e = std::getenv("BSB200_PEER_TIMEOUT")) {   // seconds a kernel waits for a peer's flag
      const double v = std::atof(e);
      if (v > 0.0) peer_timeout_ns = (uint64_t) (v * 1e9);
    }
    shm = new (std::nothrow) ShmRendezvous();
    if (!shm) return fail(BSB200_ERR_BUFFER, "out of memory");
    int rc = shm->open(id + 8, world, rank);
    if (rc) return rc;
    transport = PEER;
    // every pair must be able to reach each other's memory
    int dev = 0;
    CKC(cudaGetDevice(&dev));
    std::vector<int> devs((size_t) world);
    if ((rc = shm->allgather(&dev, sizeof(int), devs.data()))) return rc;
    for (int p = 0; p < world; p++) {
      if (p == rank) continue;
      if (devs[(size_t) p] == dev)
        return fail(BSB200_ERR_INVALID, "ranks %d and %d use the same GPU %d (one process per GPU)", rank, p, dev);
      int can = 0;
      CKC(cudaDeviceCanAccessPeer(&can, dev, devs[(size_t) p]));
      if (!can)
        return fail(BSB200_ERR_UNSUPPORTED, "GPU %d cannot access GPU %d's memory: make the id with BSB200_COMM=nccl", dev,
                    devs[(size_t) p]);
    }
    return BSB200_OK;
  }
  NcclApi *a = api();
  if (!a) return fail(BSB200_ERR_CUDA, "libnccl.so.2 could not be loaded (needed for an NCCL communicator id)");
  ncclUniqueId nid;
  std::memcpy(&nid, unique_id128, NCCL_UNIQUE_ID_BYTES);
  ncclComm_t c = nullptr;
  NCK(a->CommInitRank(&c, world, nid, rank), "CommInitRank");
  nccl = c;
  transport = NCCL;
  return BSB200_OK;
}

int Comm::host_allgather(const void *mine, size_t bytes, void *all) {
  if (transport != PEER) return fail(BSB200_ERR_INVALID, "host_allgather needs the peer transport");
  return shm->allgather(mine, bytes, all);
}

namespace {
struct ConnectHead {
  cudaIpcMemHandle_t handle;
  uint64_t base_off;            // arena - base of the underlying allocation the handle names
  uint64_t off_recv, off_z, arena_bytes;
  uint32_t magic;
  int32_t n_off, recv_total, rank;
};

// offset of a device pointer inside the allocation cudaIpcGetMemHandle exports (driver API through the runtime)
int allocation_offset(const void *ptr, uint64_t *off) {
  typedef int (*range_fn)(unsigned long long *, size_t *, unsigned long long);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CKC(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
  if (!fn || qr != cudaDriverEntryPointSuccess) return fail(BSB200_ERR_CUDA, "cuMemGetAddressRange is not available");
  unsigned long long base = 0;
  size_t size = 0;
  const int r = reinterpret_cast<range_fn>(fn)(&base, &size, (unsigned long long) (uintptr_t) ptr);
  if (r != 0) return fail(BSB200_ERR_CUDA, "cuMemGetAddressRange failed (%d)", r);
  *off = (uint64_t) ((unsigned long long) (uintptr_t) ptr - base);
  return BSB200_OK;
}
}   // namespace

int Comm::alloc_arena(size_t z_bytes_, size_t recv_count_, size_t gsum_doubles, cudaStream_t s) {
  auto up = [](size_t v) { return (v + 255) & ~(size_t) 255; };
  z_bytes = z_bytes_; recv_count = recv_count_;
  gsum_stride = (gsum_doubles + 1) & ~(size_t) 1;
  off_recv = up(4096 + 2 * gsum_stride * sizeof(double));
  off_z = up(off_recv + recv_count * sizeof(int32_t));
  arena_bytes = up(off_z + z_bytes);
  CKC(cudaMalloc((void **) &arena, arena_bytes));
  CKC(cudaMemsetAsync(arena, 0, arena_bytes, s));
  CKC(cudaMalloc((void **) &counters, sizeof(uint32_t) * (2 * (size_t) world + 2)));
  CKC(cudaMemsetAsync(counters, 0, sizeof(uint32_t) * (2 * (size_t) world + 2), s));
  return BSB200_OK;
}

int Comm::connect(const HaloPlan &halo, int ncol_, cudaStream_t s) {
  if (transport != PEER) return BSB200_OK;
  if (!arena) return fail(BSB200_ERR_INVALID, "connect() before alloc_arena()");
  ncol = ncol_;
  const int W = world;
  const size_t n_off = (size_t) ncol * W + 1;
  if (sizeof(ConnectHead) + n_off * sizeof(int32_t) > ShmRendezvous::kSlot)
    return fail(BSB200_ERR_UNSUPPORTED, "%d colours x %d ranks exceed the rendezvous slot", ncol, W);
  const uint32_t magic = 0xB5B20000u + (uint32_t) rank;
  CKC(cudaMemcpyAsync(arena + 4092, &magic, 4, cudaMemcpyHostToDevice, s));
  CKC(cudaStreamSynchronize(s));   // the labels and the receive list the chain placed in the arena are there, too

  std::vector<unsigned char> msg(sizeof(ConnectHead) + n_off * sizeof(int32_t)), all(msg.size() * (size_t) W);
  ConnectHead head{};
  CKC(cudaIpcGetMemHandle(&head.handle, arena));
  int rc = allocation_offset(arena, &head.base_off);
  if (rc) return rc;
  head.off_recv = off_recv; head.off_z = off_z; head.arena_bytes = arena_bytes;
  head.magic = magic;
  head.n_off = (int32_t) n_off;
  head.recv_total = (int32_t) halo.recv_total();
  head.rank = rank;
  std::memcpy(msg.data(), &head, sizeof(head));
  std::memcpy(msg.data() + sizeof(head), halo.recv_off.data(), n_off * sizeof(int32_t));
  if ((rc = shm->allgather(msg.data(), msg.size(), all.data()))) return rc;

  std::vector<HaloLeg> legs((size_t) ncol * (size_t) (W - 1));
  halo_leg_count.assign((size_t) ncol, 0);
  const size_t send_total = (size_t) halo.send_total();
  if (send_total) CKC(cudaMalloc((void **) &send_dst, sizeof(int32_t) * send_total));
  std::vector<ConnectHead> heads((size_t) W);
  // neighbouring ranks: any label traffic in either direction, in any colour (the same set on both sides, because
  // the send list of one is the receive list of the other)
  std::vector<char> is_nbr((size_t) W, 0);
  for (int p = 0; p < W; p++)
    for (int c = 0; c < ncol && p != rank; c++) {
      const size_t e = (size_t) c * W + p;
      if (halo.send_off[e + 1] > halo.send_off[e] || halo.recv_off[e + 1] > halo.recv_off[e]) is_nbr[(size_t) p] = 1;
    }
  for (int p = 0; p < W; p++) {
    if (p == rank) continue;
    ConnectHead &ph = heads[(size_t) p];
    std::memcpy(&ph, all.data() + (size_t) p * msg.size(), sizeof(ph));
    const int32_t *poff = reinterpret_cast<const int32_t *>(all.data() + (size_t) p * msg.size() + sizeof(ph));
    if (ph.rank != p || ph.n_off != (int32_t) n_off)
      return fail(BSB200_ERR_INVALID, "rank %d announced itself as rank %d with %d halo offsets (expected %d)", p, ph.rank, ph.n_off, (int) n_off);
    CKC(cudaIpcOpenMemHandle(&peer_base[p], ph.handle, cudaIpcMemLazyEnablePeerAccess));
    peer_arena[p] = (uint8_t *) peer_base[p] + ph.base_off;
    uint32_t seen = 0;
    CKC(cudaMemcpy(&seen, peer_arena[p] + 4092, 4, cudaMemcpyDeviceToHost));
    if (seen != ph.magic) return fail(BSB200_ERR_CUDA, "the mapped arena of rank %d does not carry its tag (%08x)", p, seen);
    for (int c = 0; c < ncol; c++) {
      const size_t e = (size_t) c * W + p;
      const int32_t s0 = halo.send_off[e], s1 = halo.send_off[e + 1];
      const int32_t nrecv = halo.recv_off[e + 1] - halo.recv_off[e];
      // the peer's receive list from this rank, same colour: positions in ITS label array, same order as our sends
      const size_t pe = (size_t) c * W + rank;
      const int32_t r0 = poff[pe], r1 = poff[pe + 1];
      if (r1 - r0 != s1 - s0)
        return fail(BSB200_ERR_INVALID, "halo lists of ranks %d and %d disagree for colour %d (%d vs %d)", rank, p, c, s1 - s0, r1 - r0);
      if (!is_nbr[(size_t) p]) continue;   // every colour phase of a neighbouring pair gets a leg, empty or not
      if (s1 > s0)
        CKC(cudaMemcpyAsync(send_dst + s0, reinterpret_cast<const int32_t *>(peer_arena[p] + ph.off_recv) + r0,
                            sizeof(int32_t) * (size_t) (s1 - s0), cudaMemcpyDefault, s));
      HaloLeg L{};
      L.peer_z = peer_arena[p] + ph.off_z;
      L.peer_flag = reinterpret_cast<uint32_t *>(peer_arena[p]) + rank;
      L.my_flag = reinterpret_cast<const uint32_t *>(arena) + p;
      L.s0 = s0; L.s1 = s1; L.nrecv = nrecv; L.peer = p;
      legs[(size_t) c * (size_t) (W - 1) + (size_t) halo_leg_count[(size_t) c]++] = L;
    }
  }
  std::vector<GatherLeg> glegs((size_t) (W - 1));
  for (int i = 0; i < W - 1; i++) {
    const int p = (rank + 1 + i) % W;
    GatherLeg g{};
    g.peer_gsum = reinterpret_cast<double *>(peer_arena[p] + 4096);
    g.peer_flag = reinterpret_cast<uint32_t *>(peer_arena[p]) + 64 + rank;
    g.my_flag = reinterpret_cast<const uint32_t *>(arena) + 64 + p;
    g.peer = p;
    glegs[(size_t) i] = g;
  }
  CKC(cudaMalloc((void **) &halo_legs, sizeof(HaloLeg) * std::max<size_t>(legs.size(), 1)));
  CKC(cudaMalloc((void **) &gather_legs, sizeof(GatherLeg) * glegs.size()));
  CKC(cudaMemcpyAsync(halo_legs, legs.data(), sizeof(HaloLeg) * legs.size(), cudaMemcpyHostToDevice, s));
  CKC(cudaMemcpyAsync(gather_legs, glegs.data(), sizeof(GatherLeg) * glegs.size(), cudaMemcpyHostToDevice, s));
  CKC(cudaStreamSynchronize(s));
  // nobody starts the data path before every rank has read the receive lists it needs
  char one = 1, seen_all[kMaxPeerWorld];
  return shm->allgather(&one, 1, seen_all);
}

void Comm::enqueue_halo_peer(int colour, const uint8_t *z, const int32_t *send_pos_dev, int *err, cudaStream_t s) {
  const int nlegs = halo_leg_count[(size_t) colour];
  if (nlegs == 0) return;
  k_halo_peer<<<nlegs, 256, 0, s>>>(halo_legs + (size_t) colour * (size_t) (world - 1), z, send_pos_dev, send_dst, counters,
                                    world, err, (unsigned long long) peer_timeout_ns);
  g_launches++;
}

void Comm::enqueue_halo_release(int colour, const uint32_t *iter, cudaStream_t s) {
  const int nlegs = halo_leg_count[(size_t) colour];
  if (nlegs == 0) return;
  k_halo_release<<<1, 32, 0, s>>>(halo_legs + (size_t) colour * (size_t) (world - 1), nlegs, iter, ncol, colour);
  g_launches++;
}

void Comm::enqueue_gather_peer(size_t offset_doubles, size_t count_doubles, int *err, cudaStream_t s) {
  k_gather_peer<<<world - 1, 256, 0, s>>>(gather_legs, gsum(), gsum_stride, offset_doubles, count_doubles, counters, world,
                                          err, (unsigned long long) peer_timeout_ns);
  g_launches++;
}

int Comm::allgather_f64(double *buf, size_t count, cudaStream_t s) {
  NcclApi *a = api();
  NCK(a->AllGather(buf + (size_t) rank * count, buf, count, ncclFloat64, (ncclComm_t) nccl, s), "AllGather");
  return BSB200_OK;
}

int Comm::exchange_u8(const uint8_t *sendbuf, const int32_t *send_off, uint8_t *recvbuf, const int32_t *recv_off,
                      cudaStream_t s) {
  NcclApi *a = api();
  bool any = false;
  for (int p = 0; p < world; p++) any |= send_off[p + 1] > send_off[p] || recv_off[p + 1] > recv_off[p];
  if (!any) return BSB200_OK;
  NCK(a->GroupStart(), "GroupStart");
  for (int p = 0; p < world; p++) {
    if (p == rank) continue;
    if (send_off[p + 1] > send_off[p])
      NCK(a->Send(sendbuf + send_off[p], (size_t) (send_off[p + 1] - send_off[p]), ncclUint8, p, (ncclComm_t) nccl, s), "Send");
    if (recv_off[p + 1] > recv_off[p])
      NCK(a->Recv(recvbuf + recv_off[p], (size_t) (recv_off[p + 1] - recv_off[p]), ncclUint8, p, (ncclComm_t) nccl, s), "Recv");
  }
  NCK(a->GroupEnd(), "GroupEnd");
  return BSB200_OK;
}

}   // namespace bsb

// Host rendezvous alone (no GPU): `rounds` allgathers of `bytes` per rank; round r sends mine[i] + r.  Test hook.
extern "C" int bsb200_debug_rendezvous(const void *id128, int32_t world, int32_t rank, const uint8_t *mine, int64_t bytes,
                                       int32_t rounds, uint8_t *all_last) {
  using namespace bsb;
  if (!id128 || !mine || !all_last || world < 1 || world > kMaxPeerWorld || rank < 0 || rank >= world || bytes < 1)
    return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  if (std::memcmp(id128, kPeerMagic, 8) != 0) return fail(BSB200_ERR_INVALID, "not a peer-transport communicator id");
  ShmRendezvous shm;
  int rc = shm.open((const unsigned char *) id128 + 8, world, rank);
  if (rc) return rc;
  std::vector<uint8_t> msg((size_t) bytes), all((size_t) bytes * (size_t) world);
  for (int r = 0; r < rounds; r++) {
    for (int64_t i = 0; i < bytes; i++) msg[(size_t) i] = (uint8_t) (mine[i] + r);
    if ((rc = shm.allgather(msg.data(), (size_t) bytes, all.data()))) return rc;
  }
  std::memcpy(all_last, all.data(), all.size());
  return BSB200_OK;
}

extern "C" int bsb200_comm_unique_id(void *id128) {
  if (!id128) return bsb::fail(BSB200_ERR_INVALID, "Invalid arguments!");
  return bsb::comm_unique_id(id128);
}
