// nccl_dyn.cu -- NCCL bound at run time.  libmcp.so has no link-time dependency on NCCL: the portfolio-sharded path
// needs no collective at all, and only the optional trial-sharded mode (mcp_run_trial_sharded) exchanges data.  The
// library is looked up with dlopen("libnccl.so.2"), which also finds a copy the host process has already loaded
// (e.g. the one bundled with PyTorch), so both sides share one NCCL.
#include <dlfcn.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace mcp {

namespace {

struct UniqueId { char internal[128]; };
typedef void *Comm;
typedef int (*fn_get_unique_id)(UniqueId *);
typedef int (*fn_comm_init_rank)(Comm *, int, UniqueId, int);
typedef int (*fn_comm_destroy)(Comm);
typedef int (*fn_all_gather)(const void *, void *, size_t, int /*ncclDataType_t*/, Comm, cudaStream_t);
typedef const char *(*fn_error_string)(int);

struct Api {
    void *handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_error_string error_string = nullptr;
    std::string why;
};

void load_api(Api &a)
{
    // MCP_NCCL_LIB names the library explicitly (tests point it at a path that does not exist to exercise the failure)
    const char *forced = getenv("MCP_NCCL_LIB");
    const char *names[] = {forced && *forced ? forced : "libnccl.so.2", forced && *forced ? forced : "libnccl.so"};
    std::string first_error;
    for (const char *n : names) {
        a.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (a.handle) break;
        const char *e = dlerror();  // glibc clears the message on read: fetch it exactly once per failure
        if (first_error.empty()) first_error = e ? e : "not found";
    }
    if (!a.handle) { a.why = std::string("dlopen(") + names[0] + "): " + first_error; return; }
    a.get_unique_id = (fn_get_unique_id)dlsym(a.handle, "ncclGetUniqueId");
    a.comm_init_rank = (fn_comm_init_rank)dlsym(a.handle, "ncclCommInitRank");
    a.comm_destroy = (fn_comm_destroy)dlsym(a.handle, "ncclCommDestroy");
    a.all_gather = (fn_all_gather)dlsym(a.handle, "ncclAllGather");
    a.error_string = (fn_error_string)dlsym(a.handle, "ncclGetErrorString");
    if (!a.get_unique_id || !a.comm_init_rank || !a.comm_destroy || !a.all_gather) {
        a.why = "libnccl.so.2 lacks a required symbol";
        a.handle = nullptr;
    }
}

Api *api()
{
    static Api a;
    static std::once_flag once;  // contexts live on different threads: bind NCCL exactly once
    std::call_once(once, [] { load_api(a); });
    return &a;
}

struct CommState { Comm comm = nullptr; int rank = 0, world = 1; };

int nccl_fail(mcp_ctx *ctx, const char *what, int rc)
{
    Api *a = api();
    std::string msg = std::string(what) + ": " + (a->error_string ? a->error_string(rc) : "NCCL error");
    return fail(ctx, MCP_E_NCCL, msg.c_str());
}

} // namespace

int nccl_unique_id(unsigned char id[128], std::string *err)
{
    Api *a = api();
    if (!a->handle) { if (err) *err = a->why; return MCP_E_NCCL; }
    UniqueId u;
    const int rc = a->get_unique_id(&u);
    if (rc != 0) { if (err) *err = a->error_string ? a->error_string(rc) : "ncclGetUniqueId failed"; return MCP_E_NCCL; }
    memcpy(id, u.internal, 128);
    return MCP_OK;
}

int nccl_comm_init(mcp_ctx *ctx, const unsigned char id[128], int rank, int world)
{
    Api *a = api();
    if (!a->handle) return fail(ctx, MCP_E_NCCL, a->why.c_str());
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, MCP_E_ARG, "comm_init: bad rank / world");
    nccl_comm_destroy(ctx);
    CommState *st = new CommState();
    st->rank = rank; st->world = world;
    UniqueId u;
    memcpy(u.internal, id, 128);
    const int rc = a->comm_init_rank(&st->comm, world, u, rank);
    if (rc != 0) { delete st; return nccl_fail(ctx, "ncclCommInitRank", rc); }
    ctx->nccl = st;
    return MCP_OK;
}

void nccl_comm_destroy(mcp_ctx *ctx)
{
    if (!ctx->nccl) return;
    CommState *st = reinterpret_cast<CommState *>(ctx->nccl);
    Api *a = api();
    if (a->handle && st->comm) a->comm_destroy(st->comm);
    delete st;
    ctx->nccl = nullptr;
}

int nccl_rank_world(mcp_ctx *ctx, int *rank, int *world)
{
    if (!ctx->nccl) return fail(ctx, MCP_E_STATE, "no communicator: call mcp_comm_init first");
    CommState *st = reinterpret_cast<CommState *>(ctx->nccl);
    *rank = st->rank; *world = st->world;
    return MCP_OK;
}

// all-gather of `bytes` bytes per rank on the context's stream
int nccl_all_gather(mcp_ctx *ctx, const void *send, void *recv, size_t bytes)
{
    if (!ctx->nccl) return fail(ctx, MCP_E_STATE, "no communicator: call mcp_comm_init first");
    CommState *st = reinterpret_cast<CommState *>(ctx->nccl);
    const int rc = api()->all_gather(send, recv, bytes, 0 /* ncclInt8 */, st->comm, ctx->stream);
    if (rc != 0) return nccl_fail(ctx, "ncclAllGather", rc);
    return MCP_OK;
}

} // namespace mcp
