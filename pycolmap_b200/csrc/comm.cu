// comm.cu -- multi-GPU plumbing of libb200match.so: one NCCL communicator per context, the all-gather that makes
// the sharded descriptor set resident on every GPU (replaces upstream's per-worker reads of
// the host-side FeatureMatcherCache, U:controllers/feature_matching_utils.cc, for the gpu_index list of
// R:pipeline/match_features.h:76-81).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a PyTorch process that is the copy torch already
// loaded, otherwise the system library), so the library itself has no link-time dependency on it and a single-GPU
// user never touches it.  Only types come from <nccl.h>.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "internal.h"

namespace b2m {

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;  // load failure
};

NcclApi* nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      api.why = std::string("[comm.cu] libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
      return;
    }
    bool ok = true;
    auto sym = [&](const char* name) {
      void* p = dlsym(api.handle, name);
      if (!p) {
        ok = false;
        api.why = std::string("[comm.cu] libnccl lacks ") + name;
      }
      return p;
    };
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) {
      dlclose(api.handle);
      api.handle = nullptr;
    }
  });
  return &api;
}

int nccl_fail(b2m_ctx* ctx, const char* what, ncclResult_t r) {
  NcclApi* a = nccl();
  char b[384];
  snprintf(b, sizeof(b), "[comm.cu] %s failed: %s", what, a->GetErrorString ? a->GetErrorString(r) : "?");
  if (ctx) ctx->err = b; else g_noctx_err = b;
  return B2M_ECUDA;
}

int need_nccl(b2m_ctx* ctx) {
  NcclApi* a = nccl();
  if (a->handle) return B2M_OK;
  if (ctx) ctx->err = a->why; else g_noctx_err = a->why;
  return B2M_ENODEV;
}

}  // namespace

void comm_release(b2m_ctx* ctx) {
  if (ctx->comm) {
    NcclApi* a = nccl();
    if (a->handle) a->CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nullptr;
  }
  ctx->comm_size = 1;
  ctx->comm_rank = 0;
}

// In-place all-gather over the communicator of `ctx`: rank r owns bytes [off[r], off[r] + len[r]) of `base` (on
// every rank the same offsets into its own copy of the array).  Equal lengths at equal spacing -> ncclAllGather;
// ragged -> one ncclBroadcast per owner, fused in a group (an all-gather-v).  Enqueued on `st`.
int comm_allgather_regions(b2m_ctx* ctx, uint8_t* base, const std::vector<size_t>& off, const std::vector<size_t>& len,
                           cudaStream_t st) {
  if (!ctx->comm || ctx->comm_size <= 1) return B2M_OK;
  NcclApi* a = nccl();
  ncclComm_t comm = static_cast<ncclComm_t>(ctx->comm);
  const int n = ctx->comm_size;
  bool uniform = true;
  for (int r = 0; r < n; ++r) uniform = uniform && len[r] == len[0] && off[r] == off[0] + static_cast<size_t>(r) * len[0];
  if (uniform) {
    if (len[0] == 0) return B2M_OK;
    ncclResult_t rc = a->AllGather(base + off[ctx->comm_rank], base + off[0], len[0], ncclUint8, comm, st);
    return rc == ncclSuccess ? B2M_OK : nccl_fail(ctx, "ncclAllGather", rc);
  }
  ncclResult_t rc = a->GroupStart();
  if (rc != ncclSuccess) return nccl_fail(ctx, "ncclGroupStart", rc);
  for (int r = 0; r < n; ++r) {
    if (len[r] == 0) continue;
    rc = a->Broadcast(base + off[r], base + off[r], len[r], ncclUint8, r, comm, st);
    if (rc != ncclSuccess) {
      a->GroupEnd();
      return nccl_fail(ctx, "ncclBroadcast", rc);
    }
  }
  rc = a->GroupEnd();
  return rc == ncclSuccess ? B2M_OK : nccl_fail(ctx, "ncclGroupEnd", rc);
}

}  // namespace b2m

using namespace b2m;

extern "C" {

int b2m_comm_get_unique_id(b2m_comm_id* out) {
  if (!out) return B2M_EINVAL;
  if (int rc = need_nccl(nullptr)) return rc;
  static_assert(sizeof(ncclUniqueId) == B2M_COMM_ID_BYTES, "b2m_comm_id carries an ncclUniqueId");
  ncclUniqueId id;
  ncclResult_t r = nccl()->GetUniqueId(&id);
  if (r != ncclSuccess) return nccl_fail(nullptr, "ncclGetUniqueId", r);
  memcpy(out->bytes, &id, sizeof(id));
  return B2M_OK;
}

int b2m_comm_init_rank(b2m_ctx* ctx, int32_t n_ranks, int32_t rank, const b2m_comm_id* id) {
  if (!ctx) return B2M_EINVAL;
  if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) {
    ctx->err = "[comm.cu] Check Failed: id != NULL && 0 <= rank < n_ranks";
    return B2M_EINVAL;
  }
  if (int rc = need_nccl(ctx)) return rc;
  comm_release(ctx);
  cudaSetDevice(ctx->device);
  ncclUniqueId nid;
  memcpy(&nid, id->bytes, sizeof(nid));
  ncclComm_t comm = nullptr;
  ncclResult_t r = nccl()->CommInitRank(&comm, n_ranks, nid, rank);
  if (r != ncclSuccess) return nccl_fail(ctx, "ncclCommInitRank", r);
  ctx->comm = comm;
  ctx->comm_size = n_ranks;
  ctx->comm_rank = rank;
  return B2M_OK;
}

int b2m_comm_init_local(b2m_ctx* const* ctxs, int32_t n) {
  if (!ctxs || n < 1) return B2M_EINVAL;
  for (int i = 0; i < n; ++i)
    if (!ctxs[i]) return B2M_EINVAL;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j)
      if (ctxs[i]->device == ctxs[j]->device) {
        ctxs[0]->err = "[comm.cu] Check Failed: the contexts of one communicator live on distinct devices";
        return B2M_EINVAL;
      }
  if (int rc = need_nccl(ctxs[0])) return rc;
  for (int i = 0; i < n; ++i) comm_release(ctxs[i]);
  if (n == 1) return B2M_OK;
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
  std::vector<ncclComm_t> comms(n, nullptr);
  ncclResult_t r = nccl()->CommInitAll(comms.data(), n, devs.data());
  if (r != ncclSuccess) return nccl_fail(ctxs[0], "ncclCommInitAll", r);
  for (int i = 0; i < n; ++i) {
    ctxs[i]->comm = comms[i];
    ctxs[i]->comm_size = n;
    ctxs[i]->comm_rank = i;
  }
  return B2M_OK;
}

int b2m_comm_destroy(b2m_ctx* ctx) {
  if (!ctx) return B2M_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  comm_release(ctx);
  return B2M_OK;
}

void b2m_comm_image_range(int32_t n_images, int32_t n_ranks, int32_t rank, int32_t* first, int32_t* count) {
  if (n_ranks < 1) n_ranks = 1;
  const int64_t per = (static_cast<int64_t>(n_images) + n_ranks - 1) / n_ranks;
  const int64_t lo = std::min<int64_t>(per * rank, n_images), hi = std::min<int64_t>(per * (rank + 1), n_images);
  if (first) *first = static_cast<int32_t>(lo);
  if (count) *count = static_cast<int32_t>(hi - lo);
}

}  // extern "C"
