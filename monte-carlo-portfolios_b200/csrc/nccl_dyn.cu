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
typedef int (*fn_send)(const void *, size_t, int /*ncclDataType_t*/, int /*peer*/, Comm, cudaStream_t);
typedef int (*fn_recv)(void *, size_t, int, int, Comm, cudaStream_t);
typedef int (*fn_group)(void);

struct Api {
    void *handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_error_string error_string = nullptr;
    fn_send send = nullptr;         // point-to-point (optional: without them the slice exchange falls back to the all-gather)
    fn_recv recv = nullptr;
    fn_group group_start = nullptr, group_end = nullptr;
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
    a.send = (fn_send)dlsym(a.handle, "ncclSend");
    a.recv = (fn_recv)dlsym(a.handle, "ncclRecv");
    a.group_start = (fn_group)dlsym(a.handle, "ncclGroupStart");
    a.group_end = (fn_group)dlsym(a.handle, "ncclGroupEnd");
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

bool nccl_has_send_recv()
{
    Api *a = api();
    return a->handle && a->send && a->recv && a->group_start && a->group_end;
}

// Every rank receives, from every other rank's block, only what it will read: the regions' heads and ITS rows of the row
// regions, at the addresses an all-gather of the whole blocks would have put them (recv_base + peer * block_bytes + ...), so
// the consumers do not care which of the two ran.  One group of ncclSend / ncclRecv pairs over NVLink / NVSwitch; the rank's
// own part is a device copy.
int nccl_slice_exchange(mcp_ctx *ctx, const unsigned char *send_block, unsigned char *recv_base, size_t block_bytes,
                        const ExchRegion *regions, int nregions, const int64_t *p0)
{
    if (!ctx->nccl) return fail(ctx, MCP_E_STATE, "no communicator: call mcp_comm_init first");
    if (!nccl_has_send_recv()) return fail(ctx, MCP_E_NCCL, "this NCCL has no ncclSend / ncclRecv");
    CommState *st = reinterpret_cast<CommState *>(ctx->nccl);
    Api *a = api();
    const int me = st->rank, world = st->world;
    int rc = a->group_start();
    if (rc != 0) return nccl_fail(ctx, "ncclGroupStart", rc);
    for (int peer = 0; peer < world && rc == 0; ++peer) {
        for (int r = 0; r < nregions && rc == 0; ++r) {
            const ExchRegion &g = regions[r];
            // what `peer` reads of my block, and what I read of `peer`'s
            const size_t s_off = g.head ? g.off : g.off + (size_t)p0[peer] * g.row_bytes;
            const size_t s_len = g.head ? g.head : (size_t)(p0[peer + 1] - p0[peer]) * g.row_bytes;
            const size_t r_off = g.head ? g.off : g.off + (size_t)p0[me] * g.row_bytes;
            const size_t r_len = g.head ? g.head : (size_t)(p0[me + 1] - p0[me]) * g.row_bytes;
            unsigned char *dst = recv_base + (size_t)peer * block_bytes + r_off;
            if (peer == me) {
                if (r_len && cudaMemcpyAsync(dst, send_block + r_off, r_len, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) rc = -1;
                continue;
            }
            if (s_len) rc = a->send(send_block + s_off, s_len, 0 /* ncclInt8 */, peer, st->comm, ctx->stream);
            if (rc == 0 && r_len) rc = a->recv(dst, r_len, 0, peer, st->comm, ctx->stream);
        }
    }
    const int rc2 = a->group_end();
    if (rc == -1) return fail(ctx, MCP_E_CUDA, "slice exchange: device copy of the rank's own part");
    if (rc != 0) return nccl_fail(ctx, "ncclSend / ncclRecv", rc);
    if (rc2 != 0) return nccl_fail(ctx, "ncclGroupEnd", rc2);
    return MCP_OK;
}

} // namespace mcp
