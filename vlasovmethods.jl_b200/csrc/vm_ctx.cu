// vm_ctx.cu -- context lifetime, error reporting, device scratch, NCCL plumbing,
// and the particle container (device SoA replacing ParticleMethods.ParticleList).
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <unordered_set>

#include "vm_internal.cuh"

static thread_local std::string g_last_error;

void vm_set_error(vm_ctx* ctx, const std::string& msg)
{
    g_last_error = msg;
    if (ctx) ctx->last_error = msg;
}

void vm_use(vm_ctx* ctx) { VM_CUDA(cudaSetDevice(ctx->device)); }

namespace {
std::mutex g_live_mu;
std::unordered_set<vm_ctx*> g_live;
}  // namespace

bool vm_ctx_alive(vm_ctx* ctx)
{
    std::lock_guard<std::mutex> lk(g_live_mu);
    return g_live.count(ctx) != 0;
}

void vm_child_quiesce(vm_ctx* ctx, int device)
{
    cudaSetDevice(device);
    if (vm_ctx_alive(ctx)) {
        cudaStreamSynchronize(ctx->stream);
        if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    } else {
        cudaDeviceSynchronize();       // orphan: its context (and streams) are gone, the device memory is not
    }
}

double* vm_partials(vm_ctx* ctx, size_t elems)
{
    if (elems > ctx->partials_elems) {
        VM_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->partials) VM_CUDA(cudaFree(ctx->partials));
        ctx->partials = nullptr;
        ctx->partials_elems = 0;
        size_t want = elems + elems / 2;
        VM_CUDA(cudaMalloc(&ctx->partials, want * sizeof(double)));
        ctx->partials_elems = want;
    }
    return ctx->partials;
}

double* vm_pinned(vm_ctx* ctx, size_t elems)
{
    if (elems > ctx->pinned_elems) {
        VM_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->pinned) VM_CUDA(cudaFreeHost(ctx->pinned));
        ctx->pinned = nullptr;
        ctx->pinned_elems = 0;
        size_t want = elems < 4096 ? 4096 : elems + elems / 2;
        VM_CUDA(cudaMallocHost(&ctx->pinned, want * sizeof(double)));
        ctx->pinned_elems = want;
    }
    return ctx->pinned;
}

void vm_launch_geometry(vm_ctx* ctx, int* grid, int* threads)
{
    int t = ctx->threads_per_cta > 0 ? ctx->threads_per_cta : 512;
    int c = ctx->ctas_per_sm > 0 ? ctx->ctas_per_sm : 2;
    *threads = t;
    *grid = ctx->sm_count * c;
}

void vm_check_peer_error(vm_ctx* ctx)
{
    if (!ctx->peers_connected) return;
    unsigned e = 0;
    VM_CUDA(cudaMemcpy(&e, ctx->xerr, sizeof(e), cudaMemcpyDeviceToHost));
    if (e) throw vm_error(VM_ERR_NCCL, "peer-memory exchange timed out waiting for another rank (ranks out of step?)");
}

void vm_prof_mark(vm_ctx* ctx)
{
    if (!ctx->profile) return;
    if (ctx->prof_used == ctx->prof_events.size()) {
        cudaEvent_t e;
        VM_CUDA(cudaEventCreate(&e));
        ctx->prof_events.push_back(e);
    }
    VM_CUDA(cudaEventRecord(ctx->prof_events[ctx->prof_used++], ctx->stream));
}

// ------------------------------------------------------------------ NCCL ----
// libnccl.so.2 is resolved at run time (the copy bundled with PyTorch is already
// mapped in a torchrun process; a Julia host would have NCCL_jll's).  Only the
// handful of entry points the path needs are bound.
namespace {
struct NcclId { char internal[128]; };   // ncclUniqueId (nccl.h): 128 opaque bytes, passed by value
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi g_nccl;
std::mutex g_nccl_mu;

NcclApi& nccl_api()
{
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_nccl.handle) return g_nccl;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) throw vm_error(VM_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce)
        throw vm_error(VM_ERR_NCCL, "libnccl is missing required symbols");
    g_nccl.handle = h;
    return g_nccl;
}

void nccl_check(NcclApi& api, int rc, const char* what)
{
    if (rc != 0) {
        std::string msg = std::string(what) + " failed: ";
        msg += api.GetErrorString ? api.GetErrorString(rc) : "nccl error";
        throw vm_error(VM_ERR_NCCL, msg);
    }
}
}  // namespace

void vm_allreduce_sum(vm_ctx* ctx, double* dev, size_t count)
{
    if (ctx->nranks <= 1) return;
    NcclApi& api = nccl_api();
    const int ncclDouble = 8, ncclSum = 0;   // ncclDataType_t / ncclRedOp_t values (nccl.h)
    nccl_check(api, api.AllReduce(dev, dev, count, ncclDouble, ncclSum, ctx->nccl_comm, ctx->stream), "ncclAllReduce");
}

void vm_allreduce_max(vm_ctx* ctx, double* dev, size_t count)
{
    if (ctx->nranks <= 1) return;
    NcclApi& api = nccl_api();
    const int ncclDouble = 8, ncclMax = 2;
    nccl_check(api, api.AllReduce(dev, dev, count, ncclDouble, ncclMax, ctx->nccl_comm, ctx->stream), "ncclAllReduce(max)");
}

// ------------------------------------------------------------------- API ----
extern "C" {

int vm_abi_version(void) { return VM_ABI_VERSION; }

const char* vm_last_error(vm_ctx* ctx) { return (ctx && vm_ctx_alive(ctx)) ? ctx->last_error.c_str() : g_last_error.c_str(); }

int vm_ctx_create(int device, vm_ctx** out)
{
    vm_ctx* ctx__ = nullptr;
    try {
        VM_REQUIRE(out != nullptr, "vm_ctx_create: out is NULL");
        *out = nullptr;
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) {
            (void)cudaGetLastError();
            throw vm_error(VM_ERR_NO_DEVICE,
                           "no CUDA device visible: libvlasov_b200 has no CPU path (sm_100a kernels only)");
        }
        VM_REQUIRE(device >= 0 && device < ndev, "vm_ctx_create: device ordinal out of range");
        VM_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        VM_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            throw vm_error(VM_ERR_UNSUPPORTED, std::string("device ") + prop.name +
                                                   " is not Blackwell (sm_100a code only)");
        vm_ctx* c = new vm_ctx();
        c->device = device;
        c->sm_count = prop.multiProcessorCount;
        c->cc_major = prop.major;
        c->cc_minor = prop.minor;
        c->smem_optin = prop.sharedMemPerBlockOptin;
        VM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        for (int i = 0; i < VM_MAX_EVENTS; ++i) VM_CUDA(cudaEventCreate(&c->events[i]));
        VM_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        VM_CUDA(cudaEventCreateWithFlags(&c->snap_ready, cudaEventDisableTiming));
        VM_CUDA(cudaEventCreateWithFlags(&c->snap_done, cudaEventDisableTiming));
        VM_CUDA(cudaMalloc(&c->ticket, (1 + VM_MAX_GROUPS) * sizeof(unsigned)));
        VM_CUDA(cudaMemset(c->ticket, 0, (1 + VM_MAX_GROUPS) * sizeof(unsigned)));
        {
            std::lock_guard<std::mutex> lk(g_live_mu);
            g_live.insert(c);
        }
        *out = c;
    }
    catch (const vm_error& e) { vm_set_error(ctx__, e.what()); return e.code; }
    catch (const std::exception& e) { vm_set_error(ctx__, e.what()); return VM_ERR_INVALID; }
    return VM_OK;
}

int vm_ctx_destroy(vm_ctx* ctx)
{
    if (!ctx) return VM_OK;
    {
        std::lock_guard<std::mutex> lk(g_live_mu);
        if (g_live.erase(ctx) == 0) return VM_OK;      // not (or no longer) a context: a second destroy is a no-op
    }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->peers_connected)
        for (int r = 0; r < ctx->nranks; ++r)
            if (r != ctx->rank && ctx->peer_inbox[r]) cudaIpcCloseMemHandle(ctx->peer_inbox[r]);
    if (ctx->inbox) cudaFree(ctx->inbox);
    if (ctx->xerr) cudaFree(ctx->xerr);
    if (ctx->nccl_comm) {
        try { nccl_api().CommDestroy(ctx->nccl_comm); } catch (...) {}
    }
    for (int i = 0; i < VM_MAX_EVENTS; ++i) if (ctx->events[i]) cudaEventDestroy(ctx->events[i]);
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->snap_ready) cudaEventDestroy(ctx->snap_ready);
    if (ctx->snap_done) cudaEventDestroy(ctx->snap_done);
    if (ctx->partials) cudaFree(ctx->partials);
    if (ctx->ticket) cudaFree(ctx->ticket);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return VM_OK;
}

int vm_sync(vm_ctx* ctx)
{
    VM_API_BEGIN(ctx)
    VM_REQUIRE(ctx != nullptr, "vm_sync: ctx is NULL");
    VM_CUDA(cudaStreamSynchronize(ctx->stream));
    vm_check_peer_error(ctx);
    VM_API_END
}

int vm_ctx_device_info(vm_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* free_bytes,
                       size_t* total_bytes)
{
    VM_API_BEGIN(ctx)
    VM_REQUIRE(ctx != nullptr, "vm_ctx_device_info: ctx is NULL");
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    size_t f = 0, t = 0;
    VM_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    VM_API_END
}

int vm_ctx_set_tuning(vm_ctx* ctx, const char* key, int value)
{
    VM_API_BEGIN(ctx)
    VM_REQUIRE(ctx != nullptr && key != nullptr, "vm_ctx_set_tuning: NULL argument");
    std::string k(key);
    if (k == "ctas_per_sm") { VM_REQUIRE(value >= 0 && value <= 8, "ctas_per_sm out of range"); ctx->ctas_per_sm = value; }
    else if (k == "threads_per_cta") {
        VM_REQUIRE(value == 0 || (value >= 64 && value <= 1024 && value % 32 == 0), "threads_per_cta must be a multiple of 32 in 64..1024");
        ctx->threads_per_cta = value;
    }
    else if (k == "replicas") {
        VM_REQUIRE(value == 0 || (value >= 1 && value <= 32 && (value & (value - 1)) == 0), "replicas must be a power of two <= 32");
        ctx->replicas = value;
    }
    else if (k == "pairs") {          // pairs of particles in flight per thread in the lane-private passes (0 = auto)
        VM_REQUIRE(value == 0 || value == 1 || value == 2 || value == 4 || value == 8, "pairs must be 0, 1, 2, 4 or 8");
        ctx->pairs = value;
    }
    else if (k == "priv_min_warps") {  // fewest warps per SM for which the lane-private deposit is still chosen (0 = auto)
        VM_REQUIRE(value >= 0 && value <= 32, "priv_min_warps out of range");
        ctx->priv_min_warps = value;
    }
    else if (k == "bankq") { VM_REQUIRE(value >= -1 && value <= 1, "bankq must be -1, 0 or 1"); ctx->bankq = value; }
    else if (k == "af") { VM_REQUIRE(value >= -1 && value <= 1, "af must be -1, 0 or 1"); ctx->af = value; }
    else if (k == "no_presolve") ctx->no_presolve = value;   // 1: separate solve kernel between the fused passes of large meshes (A/B)
    else if (k == "af_replicas") {
        VM_REQUIRE(value == 0 || (value >= 1 && value <= 32 && (value & (value - 1)) == 0), "af_replicas must be a power of two <= 32");
        ctx->af_replicas = value;
    }
    else if (k == "af_ctas") { VM_REQUIRE(value >= 0 && value <= 4 && value != 3, "af_ctas must be 0, 1, 2 or 4"); ctx->af_ctas = value; }
    else if (k == "no_repg") ctx->no_repg = value;   // 1: single (bank-conflicting) gather table in the fused pass (A/B)
    else if (k == "profile") ctx->profile = value;
    else if (k == "force_match") ctx->force_match = value;   // 1: MATCH.ANY grouping instead of xor rounds (A/B)
    else if (k == "no_uniform_w") ctx->no_uniform_w = value;   // 1: always stream per-particle weights (A/B)
    else if (k == "no_pdl") ctx->no_pdl = value;     // 1: plain stream serialization for the pass kernels (A/B)
    else if (k == "no_fuse") ctx->no_fuse = value;   // 1: separate reduce/solve kernels even for small grids (A/B)
    else throw vm_error(VM_ERR_INVALID, "unknown tuning key: " + k);
    VM_API_END
}

int vm_comm_unique_id(void* out128)
{
    vm_ctx* ctx__ = nullptr;
    try {
        VM_REQUIRE(out128 != nullptr, "vm_comm_unique_id: out is NULL");
        NcclApi& api = nccl_api();
        nccl_check(api, api.GetUniqueId(out128), "ncclGetUniqueId");
    }
    catch (const vm_error& e) { vm_set_error(ctx__, e.what()); return e.code; }
    return VM_OK;
}

int vm_ctx_comm_init(vm_ctx* ctx, int rank, int nranks, const void* id128)
{
    VM_API_BEGIN(ctx)
    VM_REQUIRE(ctx != nullptr, "vm_ctx_comm_init: ctx is NULL");
    VM_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "vm_ctx_comm_init: bad rank/nranks");
    VM_REQUIRE(ctx->nccl_comm == nullptr, "vm_ctx_comm_init: communicator already initialised");
    if (nranks > 1) {
        VM_REQUIRE(id128 != nullptr, "vm_ctx_comm_init: id is NULL");
        NcclApi& api = nccl_api();
        NcclId id;
        std::memcpy(id.internal, id128, 128);
        nccl_check(api, api.CommInitRank(&ctx->nccl_comm, nranks, id, rank), "ncclCommInitRank");
    }
    ctx->rank = rank;
    ctx->nranks = nranks;
    VM_API_END
}

int vm_ctx_peer_handle(vm_ctx* ctx, void* out64)
{
    VM_API_BEGIN(ctx)
    VM_REQUIRE(ctx != nullptr && out64 != nullptr, "vm_ctx_peer_handle: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    if (!ctx->inbox) {
        const size_t bytes = (size_t)VM_XINBOX_WORDS * sizeof(unsigned long long);
        VM_CUDA(cudaMalloc(&ctx->inbox, bytes));
        VM_CUDA(cudaMemset(ctx->inbox, 0, bytes));
        VM_CUDA(cudaMalloc(&ctx->xerr, sizeof(unsigned)));
        VM_CUDA(cudaMemset(ctx->xerr, 0, sizeof(unsigned)));
    }
    cudaIpcMemHandle_t h;
    VM_CUDA(cudaIpcGetMemHandle(&h, ctx->inbox));
    std::memcpy(out64, &h, 64);
    VM_API_END
}

int vm_ctx_peer_connect(vm_ctx* ctx, const void* handles)
{
    VM_API_BEGIN(ctx)
    VM_REQUIRE(ctx != nullptr && handles != nullptr, "vm_ctx_peer_connect: NULL argument");
    VM_REQUIRE(ctx->inbox != nullptr, "vm_ctx_peer_connect: call vm_ctx_peer_handle first");
    VM_REQUIRE(ctx->nranks > 1 && ctx->nranks <= VM_MAX_PEERS, "vm_ctx_peer_connect: needs 2..8 ranks (vm_ctx_comm_init first)");
    VM_REQUIRE(!ctx->peers_connected, "vm_ctx_peer_connect: already connected");
    for (int r = 0; r < ctx->nranks; ++r) {
        if (r == ctx->rank) { ctx->peer_inbox[r] = ctx->inbox; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const char*)handles + 64 * r, 64);
        void* ptr = nullptr;
        VM_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_inbox[r] = (unsigned long long*)ptr;
    }
    ctx->peers_connected = true;
    VM_API_END
}

int vm_ctx_comm_info(vm_ctx* ctx, int* rank, int* nranks)
{
    VM_API_BEGIN(ctx)
    VM_REQUIRE(ctx != nullptr, "vm_ctx_comm_info: ctx is NULL");
    if (rank) *rank = ctx->rank;
    if (nranks) *nranks = ctx->nranks;
    VM_API_END
}

int vm_event_record(vm_ctx* ctx, int slot)
{
    VM_API_BEGIN(ctx)
    VM_REQUIRE(ctx != nullptr && slot >= 0 && slot < VM_MAX_EVENTS, "vm_event_record: bad slot");
    VM_CUDA(cudaEventRecord(ctx->events[slot], ctx->stream));
    VM_API_END
}

int vm_event_elapsed_ms(vm_ctx* ctx, int a, int b, double* ms)
{
    VM_API_BEGIN(ctx)
    VM_REQUIRE(ctx != nullptr && ms != nullptr && a >= 0 && a < VM_MAX_EVENTS && b >= 0 && b < VM_MAX_EVENTS,
               "vm_event_elapsed_ms: bad argument");
    VM_CUDA(cudaEventSynchronize(ctx->events[b]));
    float f = 0.f;
    VM_CUDA(cudaEventElapsedTime(&f, ctx->events[a], ctx->events[b]));
    *ms = (double)f;
    VM_API_END
}

unsigned long long vm_launch_count(vm_ctx* ctx) { return (ctx && vm_ctx_alive(ctx)) ? ctx->launches : 0ull; }

int vm_profile_read(vm_ctx* ctx, long* launches, double* total_ms)
{
    VM_API_BEGIN(ctx)
    VM_REQUIRE(ctx != nullptr, "vm_profile_read: ctx is NULL");
    VM_CUDA(cudaStreamSynchronize(ctx->stream));
    double tot = 0.0;
    const size_t pairs = ctx->prof_used / 2;
    for (size_t i = 0; i < pairs; ++i) {
        float ms = 0.f;
        VM_CUDA(cudaEventElapsedTime(&ms, ctx->prof_events[2 * i], ctx->prof_events[2 * i + 1]));
        tot += ms;
    }
    ctx->prof_used = 0;
    if (launches) *launches = (long)pairs;
    if (total_ms) *total_ms = tot;
    VM_API_END
}

}  // extern "C"

// -------------------------------------------------------------- particles ---
namespace {

// AoS (3 x N column-major: x,v,w interleaved) <-> SoA through shared memory so that
// both the global reads and the global writes are coalesced.
__global__ void __launch_bounds__(256) k_aos_to_soa(const double* __restrict__ z, double* __restrict__ x,
                                                    double* __restrict__ v, double* __restrict__ w, long n)
{
    __shared__ double tile[3 * 256];
    for (long base = (long)blockIdx.x * 256; base < n; base += (long)gridDim.x * 256) {
        const long cnt = (n - base < 256) ? (n - base) : 256;
        for (int i = threadIdx.x; i < 3 * cnt; i += 256) tile[i] = z[3 * base + i];
        __syncthreads();
        if (threadIdx.x < cnt) {
            x[base + threadIdx.x] = tile[3 * threadIdx.x + 0];
            v[base + threadIdx.x] = tile[3 * threadIdx.x + 1];
            w[base + threadIdx.x] = tile[3 * threadIdx.x + 2];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_soa_to_aos(const double* __restrict__ x, const double* __restrict__ v,
                                                    const double* __restrict__ w, double* __restrict__ z, long n)
{
    __shared__ double tile[3 * 256];
    for (long base = (long)blockIdx.x * 256; base < n; base += (long)gridDim.x * 256) {
        const long cnt = (n - base < 256) ? (n - base) : 256;
        if (threadIdx.x < cnt) {
            tile[3 * threadIdx.x + 0] = x[base + threadIdx.x];
            tile[3 * threadIdx.x + 1] = v[base + threadIdx.x];
            tile[3 * threadIdx.x + 2] = w[base + threadIdx.x];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * cnt; i += 256) z[3 * base + i] = tile[i];
        __syncthreads();
    }
}

// flag[0] stays 1 iff every weight has the same bits as w[0]
__global__ void __launch_bounds__(256) k_weights_uniform(const double* __restrict__ w, long n, int* __restrict__ flag)
{
    const long long first = __double_as_longlong(w[0]);
    bool same = true;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        same = same && (__double_as_longlong(w[i]) == first);
    if (!__all_sync(VM_FULL_MASK, same) && (threadIdx.x & 31) == 0) atomicExch(flag, 0);
}

__global__ void __launch_bounds__(256) k_fill_const(double* __restrict__ a, long n, double c)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] = c;
}

// out[0] = max |w| (bit pattern of a non-negative double orders like an integer: atomicMax is exact and order-free)
__global__ void __launch_bounds__(256) k_wabs_max(const double* __restrict__ w, long n, unsigned long long* __restrict__ out)
{
    double m = 0.0;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) m = fmax(m, fabs(w[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(VM_FULL_MASK, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// out[0] += sum_p floor(|w_p| * q): integer sum -- exact, so independent of the order and of the sharding
__global__ void __launch_bounds__(256) k_wabs_qsum(const double* __restrict__ w, long n, double q, unsigned long long* __restrict__ out)
{
    unsigned long long s = 0ull;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s += (unsigned long long)(fabs(w[i]) * q);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(VM_FULL_MASK, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// Host <-> device transfers of big arrays go through a pair of pinned bounce buffers so that
// pageable host memory (a Julia Array) still streams at PCIe rate and overlaps with the copy engine.
void copy_h2d(vm_ctx* ctx, double* dst, const double* src, size_t n)
{
    VM_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
}
void copy_d2h(vm_ctx* ctx, double* dst, const double* src, size_t n)
{
    VM_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
}

}  // namespace

bool vm_particles_uniform_weight(vm_particles* p, double* w0)
{
    vm_ctx* ctx = p->ctx;
    if (ctx->no_uniform_w || p->n == 0) return false;
    if (p->w_dirty) {
        int* flag = nullptr;
        VM_CUDA(cudaMalloc(&flag, sizeof(int)));
        const int one = 1;
        VM_CUDA(cudaMemcpyAsync(flag, &one, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        k_weights_uniform<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(p->w, p->n, flag);
        ++ctx->launches;
        int h = 0;
        double first = 0.0;
        VM_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        VM_CUDA(cudaMemcpyAsync(&first, p->w, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        VM_CUDA(cudaStreamSynchronize(ctx->stream));
        VM_CUDA(cudaFree(flag));
        p->uniform_w = (h == 1);
        p->w0 = first;
        p->w_dirty = false;
    }
    if (w0) *w0 = p->w0;
    return p->uniform_w;
}

// Scale exponent S of the fixed-point deposit: every |w_p| B_j is rounded to a multiple of 2^-S and summed as a 64-bit
// integer.  S must be the same on every rank and for every sharding of the same particles, so it is derived from
// quantities that are EXACT: max |w| (a maximum), sum_p floor(|w_p| 2^20 / max|w|) (an integer sum below 2^53) and the
// particle count.  U = (qsum + N) max|w| 2^-20 >= sum |w| bounds every row sum; S = min(60 - ilogb(U), 50 - ilogb(max|w|))
// keeps row sums below 2^62 and single contributions below 2^51 (the magic-number rounding of fix_of).
int vm_particles_fixed_scale(vm_particles* p)
{
    vm_ctx* ctx = p->ctx;
    if (!p->fix_dirty && p->fix_nranks == ctx->nranks) return p->fix_S;
    double* d = vm_partials(ctx, 8);                       // [0] max bits / max, [1] qsum, [2] n
    double* host = vm_pinned(ctx, 8);
    VM_CUDA(cudaMemsetAsync(d, 0, 8 * sizeof(double), ctx->stream));
    if (p->n > 0) {
        k_wabs_max<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(p->w, p->n, (unsigned long long*)d);
        ++ctx->launches;
    }
    vm_allreduce_max(ctx, d, 1);                           // (the bit pattern of a non-negative double IS the double)
    VM_CUDA(cudaMemcpyAsync(host, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    VM_CUDA(cudaStreamSynchronize(ctx->stream));
    const double wmax = host[0];
    int S = 0;
    if (wmax > 0.0 && wmax < 1e300) {
        const double q = 1048576.0 / wmax;
        VM_CUDA(cudaMemsetAsync(d + 1, 0, sizeof(double), ctx->stream));
        if (p->n > 0) {
            k_wabs_qsum<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(p->w, p->n, q, (unsigned long long*)(d + 1));
            ++ctx->launches;
        }
        unsigned long long qs = 0ull;
        VM_CUDA(cudaMemcpyAsync(&qs, d + 1, sizeof(qs), cudaMemcpyDeviceToHost, ctx->stream));
        VM_CUDA(cudaStreamSynchronize(ctx->stream));
        host[1] = (double)qs;                              // < 2^33 * 2^20 = 2^53: exact
        host[2] = (double)p->n;
        VM_CUDA(cudaMemcpyAsync(d + 1, host + 1, 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        vm_allreduce_sum(ctx, d + 1, 2);                   // integers below 2^53: exact in any order
        VM_CUDA(cudaMemcpyAsync(host + 1, d + 1, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        VM_CUDA(cudaStreamSynchronize(ctx->stream));
        const double U = (host[1] + host[2]) * wmax * (1.0 / 1048576.0);
        const int s_sum = 60 - ilogb(U), s_one = 50 - ilogb(wmax);
        S = s_sum < s_one ? s_sum : s_one;
        if (S > 1000) S = 1000;
        if (S < -1000) S = -1000;
    }
    p->fix_S = S;
    p->fix_ok = (wmax == 0.0) || (wmax > 0.0 && wmax < 1e300);     // (NaN / infinite weights: the fp64 layouts propagate them)
    p->fix_dirty = false;
    p->fix_nranks = ctx->nranks;
    return S;
}

extern "C" {

int vm_particles_create(vm_ctx* ctx, long n, vm_particles** out)
{
    VM_API_BEGIN(ctx)
    VM_REQUIRE(ctx != nullptr && out != nullptr, "vm_particles_create: NULL argument");
    VM_REQUIRE(n >= 0, "vm_particles_create: negative size");
    VM_REQUIRE(n < (1L << 33) - (1L << 24), "vm_particles_create: at most 2^33 - 2^24 particles per GPU (32-bit pair indices)");
    *out = nullptr;
    vm_particles* p = new vm_particles();
    p->ctx = ctx;
    p->device = ctx->device;
    p->n = n;
    size_t bytes = (size_t)(n > 0 ? n : 1) * sizeof(double);
    cudaError_t e1 = cudaMalloc(&p->x, bytes), e2 = cudaMalloc(&p->v, bytes), e3 = cudaMalloc(&p->w, bytes);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        (void)cudaGetLastError();
        if (p->x) cudaFree(p->x);
        if (p->v) cudaFree(p->v);
        if (p->w) cudaFree(p->w);
        delete p;
        throw vm_error(VM_ERR_NOMEM, "vm_particles_create: cudaMalloc failed");
    }
    VM_CUDA(cudaMemsetAsync(p->x, 0, bytes, ctx->stream));
    VM_CUDA(cudaMemsetAsync(p->v, 0, bytes, ctx->stream));
    VM_CUDA(cudaMemsetAsync(p->w, 0, bytes, ctx->stream));
    *out = p;
    VM_API_END
}

int vm_particles_destroy(vm_particles* p)
{
    if (!p) return VM_OK;
    vm_child_quiesce(p->ctx, p->device);
    cudaFree(p->x); cudaFree(p->v); cudaFree(p->w);
    if (p->a) cudaFree(p->a);
    if (p->snap) cudaFree(p->snap);
    for (double* q : p->work) if (q) cudaFree(q);
    delete p;
    return VM_OK;
}

long vm_particles_size(vm_particles* p) { return p ? p->n : -1; }

int vm_particles_upload_soa(vm_particles* p, const double* x, const double* v, const double* w)
{
    VM_API_BEGIN(p ? p->ctx : nullptr)
    VM_REQUIRE(p != nullptr, "vm_particles_upload_soa: NULL handle");
    if (p->n > 0) {
        if (x) copy_h2d(p->ctx, p->x, x, (size_t)p->n);
        if (v) copy_h2d(p->ctx, p->v, v, (size_t)p->n);
        if (w) { copy_h2d(p->ctx, p->w, w, (size_t)p->n); p->w_dirty = true; p->fix_dirty = true; }
        VM_CUDA(cudaStreamSynchronize(p->ctx->stream));   // host buffers are only borrowed for the call
    }
    VM_API_END
}

int vm_particles_download_soa(vm_particles* p, double* x, double* v, double* w)
{
    VM_API_BEGIN(p ? p->ctx : nullptr)
    VM_REQUIRE(p != nullptr, "vm_particles_download_soa: NULL handle");
    if (p->n > 0) {
        if (x) copy_d2h(p->ctx, x, p->x, (size_t)p->n);
        if (v) copy_d2h(p->ctx, v, p->v, (size_t)p->n);
        if (w) copy_d2h(p->ctx, w, p->w, (size_t)p->n);
    }
    VM_CUDA(cudaStreamSynchronize(p->ctx->stream));
    vm_check_peer_error(p->ctx);
    VM_API_END
}

int vm_particles_set_uniform_weight(vm_particles* p, double w0)
{
    VM_API_BEGIN(p ? p->ctx : nullptr)
    VM_REQUIRE(p != nullptr, "vm_particles_set_uniform_weight: NULL handle");
    vm_ctx* ctx = p->ctx;
    if (p->n > 0) {
        // the device array is filled too (8 B/particle written at HBM rate instead of uploaded over PCIe): passes
        // that stream per-particle weights and downloads keep working
        k_fill_const<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(p->w, p->n, w0);
        VM_LAUNCHED(ctx);
    }
    p->uniform_w = true;
    p->w0 = w0;
    p->w_dirty = false;
    p->fix_dirty = true;
    VM_API_END
}

int vm_particles_upload_aos(vm_particles* p, const double* z)
{
    VM_API_BEGIN(p ? p->ctx : nullptr)
    VM_REQUIRE(p != nullptr && z != nullptr, "vm_particles_upload_aos: NULL argument");
    if (p->n > 0) {
        vm_ctx* ctx = p->ctx;
        p->w_dirty = true;
        p->fix_dirty = true;
        // stage in chunks through the scratch buffer to bound the extra device memory
        const long chunk = 1L << 22;   // particles per chunk (96 MiB of AoS)
        double* stage = vm_partials(ctx, (size_t)3 * (size_t)(p->n < chunk ? p->n : chunk));
        for (long off = 0; off < p->n; off += chunk) {
            long cnt = p->n - off < chunk ? p->n - off : chunk;
            copy_h2d(ctx, stage, z + 3 * off, (size_t)3 * cnt);
            int grid = (int)((cnt + 255) / 256);
            if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
            k_aos_to_soa<<<grid, 256, 0, ctx->stream>>>(stage, p->x + off, p->v + off, p->w + off, cnt);
            VM_LAUNCHED(ctx);
        }
        VM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    VM_API_END
}

int vm_particles_download_aos(vm_particles* p, double* z)
{
    VM_API_BEGIN(p ? p->ctx : nullptr)
    VM_REQUIRE(p != nullptr && z != nullptr, "vm_particles_download_aos: NULL argument");
    if (p->n > 0) {
        vm_ctx* ctx = p->ctx;
        const long chunk = 1L << 22;
        double* stage = vm_partials(ctx, (size_t)3 * (size_t)(p->n < chunk ? p->n : chunk));
        for (long off = 0; off < p->n; off += chunk) {
            long cnt = p->n - off < chunk ? p->n - off : chunk;
            int grid = (int)((cnt + 255) / 256);
            if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
            k_soa_to_aos<<<grid, 256, 0, ctx->stream>>>(p->x + off, p->v + off, p->w + off, stage, cnt);
            VM_LAUNCHED(ctx);
            copy_d2h(ctx, z + 3 * off, stage, (size_t)3 * cnt);
        }
    }
    VM_CUDA(cudaStreamSynchronize(p->ctx->stream));
    VM_API_END
}

int vm_particles_snapshot_begin(vm_particles* p, double* x_host, double* v_host)
{
    VM_API_BEGIN(p ? p->ctx : nullptr)
    VM_REQUIRE(p != nullptr, "vm_particles_snapshot_begin: NULL handle");
    vm_ctx* ctx = p->ctx;
    if (p->n > 0 && (x_host || v_host)) {
        const size_t bytes = (size_t)p->n * sizeof(double);
        if (!p->snap) VM_CUDA(cudaMalloc(&p->snap, 2 * bytes));
        // the staging buffer may still be draining to the host from the previous snapshot
        if (p->snap_pending) VM_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->snap_done, 0));
        if (x_host) VM_CUDA(cudaMemcpyAsync(p->snap, p->x, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        if (v_host) VM_CUDA(cudaMemcpyAsync(p->snap + p->n, p->v, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        VM_CUDA(cudaEventRecord(ctx->snap_ready, ctx->stream));
        VM_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->snap_ready, 0));
        if (x_host) VM_CUDA(cudaMemcpyAsync(x_host, p->snap, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
        if (v_host) VM_CUDA(cudaMemcpyAsync(v_host, p->snap + p->n, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
        VM_CUDA(cudaEventRecord(ctx->snap_done, ctx->copy_stream));
        p->snap_pending = true;
    }
    VM_API_END
}

int vm_particles_snapshot_wait(vm_particles* p)
{
    VM_API_BEGIN(p ? p->ctx : nullptr)
    VM_REQUIRE(p != nullptr, "vm_particles_snapshot_wait: NULL handle");
    if (p->snap_pending) {
        VM_CUDA(cudaEventSynchronize(p->ctx->snap_done));
        p->snap_pending = false;
        vm_check_peer_error(p->ctx);
    }
    VM_API_END
}

int vm_host_alloc(size_t bytes, void** out)
{
    vm_ctx* ctx__ = nullptr;
    try {
        VM_REQUIRE(out != nullptr, "vm_host_alloc: out is NULL");
        *out = nullptr;
        if (cudaMallocHost(out, bytes ? bytes : 1) != cudaSuccess) {
            (void)cudaGetLastError();
            throw vm_error(VM_ERR_NOMEM, "vm_host_alloc: cudaMallocHost failed");
        }
    }
    catch (const vm_error& e) { vm_set_error(ctx__, e.what()); return e.code; }
    return VM_OK;
}

int vm_host_free(void* ptr)
{
    if (ptr) cudaFreeHost(ptr);
    return VM_OK;
}

int vm_particles_copy(vm_particles* dst, vm_particles* src)
{
    VM_API_BEGIN(dst ? dst->ctx : nullptr)
    VM_REQUIRE(dst != nullptr && src != nullptr, "vm_particles_copy: NULL handle");
    VM_REQUIRE(dst->ctx == src->ctx && dst->n == src->n, "vm_particles_copy: handles differ in context or size");
    size_t bytes = (size_t)dst->n * sizeof(double);
    if (bytes) {
        VM_CUDA(cudaMemcpyAsync(dst->x, src->x, bytes, cudaMemcpyDeviceToDevice, dst->ctx->stream));
        VM_CUDA(cudaMemcpyAsync(dst->v, src->v, bytes, cudaMemcpyDeviceToDevice, dst->ctx->stream));
        VM_CUDA(cudaMemcpyAsync(dst->w, src->w, bytes, cudaMemcpyDeviceToDevice, dst->ctx->stream));
    }
    dst->w_dirty = true;
    dst->fix_dirty = true;
    VM_API_END
}

}  // extern "C"
