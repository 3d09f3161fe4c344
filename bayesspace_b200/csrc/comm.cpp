// comm.cpp -- run-time binding of NCCL for the spot-sharded chain (see comm.hpp).
#include "comm.hpp"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "common.hpp"

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

int comm_unique_id(void *out128) {
  NcclApi *a = api();
  if (!a) return fail(BSB200_ERR_CUDA, "libnccl.so.2 could not be loaded (needed for world > 1)");
  ncclUniqueId id;
  NCK(a->GetUniqueId(&id), "GetUniqueId");
  std::memcpy(out128, &id, NCCL_UNIQUE_ID_BYTES);
  return BSB200_OK;
}

Comm::~Comm() {
  NcclApi *a = api();
  if (nccl && a && a->CommDestroy) a->CommDestroy((ncclComm_t) nccl);
}

int Comm::init(int world_, int rank_, const void *unique_id128) {
  NcclApi *a = api();
  if (!a) return fail(BSB200_ERR_CUDA, "libnccl.so.2 could not be loaded (needed for world > 1)");
  if (!unique_id128) return fail(BSB200_ERR_INVALID, "world > 1 needs the communicator id of bsb200_comm_unique_id()");
  world = world_; rank = rank_;
  ncclUniqueId id;
  std::memcpy(&id, unique_id128, NCCL_UNIQUE_ID_BYTES);
  ncclComm_t c = nullptr;
  NCK(a->CommInitRank(&c, world, id, rank), "CommInitRank");
  nccl = c;
  return BSB200_OK;
}

int Comm::allgather_f64(double *buf, size_t count, cudaStream_t s) {
  NcclApi *a = api();
  NCK(a->AllGather(buf + (size_t) rank * count, buf, count, ncclFloat64, (ncclComm_t) nccl, s), "AllGather");
  return BSB200_OK;
}

int Comm::allreduce_max_f64(double *buf, size_t count, cudaStream_t s) {
  NcclApi *a = api();
  NCK(a->AllReduce(buf, buf, count, ncclFloat64, ncclMax, (ncclComm_t) nccl, s), "AllReduce");
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

extern "C" int bsb200_comm_unique_id(void *id128) {
  if (!id128) return bsb::fail(BSB200_ERR_INVALID, "Invalid arguments!");
  return bsb::comm_unique_id(id128);
}
