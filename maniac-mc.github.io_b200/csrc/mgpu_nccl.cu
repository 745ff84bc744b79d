// mgpu_nccl.cu -- end-of-run reduction of per-isotherm-point averages across ranks.
// Walkers are independent, so this is the only inter-GPU exchange of the whole path
// (SURVEY 8e): a sum of a few hundred doubles.  NCCL is resolved at run time with
// dlopen so the library has no link-time dependency on a particular NCCL build (the
// one torch.distributed already loaded is reused when present).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstring>
#include <string>
#include "../../include/maniac_gpu.h"

namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8, ncclSum = 0 };
struct Nccl {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    cudaStream_t stream = nullptr;
    double *dbuf = nullptr; size_t dcap = 0;
} N;
std::string n_err;
int nfail(const std::string &m) { n_err = m; return 1; }
int load()
{
    if (N.lib) return 0;
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char *nm : names) { N.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (N.lib) break; }
    if (!N.lib) return nfail(std::string("mgpu_nccl: cannot dlopen libnccl: ") + dlerror());
    N.GetUniqueId = (decltype(N.GetUniqueId))dlsym(N.lib, "ncclGetUniqueId");
    N.CommInitRank = (decltype(N.CommInitRank))dlsym(N.lib, "ncclCommInitRank");
    N.AllReduce = (decltype(N.AllReduce))dlsym(N.lib, "ncclAllReduce");
    N.CommDestroy = (decltype(N.CommDestroy))dlsym(N.lib, "ncclCommDestroy");
    N.GetErrorString = (decltype(N.GetErrorString))dlsym(N.lib, "ncclGetErrorString");
    if (!N.GetUniqueId || !N.CommInitRank || !N.AllReduce || !N.CommDestroy) return nfail("mgpu_nccl: missing NCCL symbols");
    return 0;
}
}

extern "C" {
const char *mgpu_nccl_last_error(void) { return n_err.c_str(); }

int mgpu_nccl_unique_id(char id[128])
{
    if (load()) return 1;
    ncclUniqueId u;
    ncclResult_t r = N.GetUniqueId(&u);
    if (r) return nfail(std::string("ncclGetUniqueId: ") + (N.GetErrorString ? N.GetErrorString(r) : "error"));
    std::memcpy(id, u.internal, 128);
    return 0;
}

int mgpu_nccl_init(const char id[128], int32_t nranks, int32_t rank)
{
    N.nranks = nranks; N.rank = rank;
    if (nranks <= 1) return 0;
    if (load()) return 1;
    ncclUniqueId u;
    std::memcpy(u.internal, id, 128);
    ncclResult_t r = N.CommInitRank(&N.comm, nranks, u, rank);
    if (r) return nfail(std::string("ncclCommInitRank: ") + (N.GetErrorString ? N.GetErrorString(r) : "error"));
    if (cudaStreamCreateWithFlags(&N.stream, cudaStreamNonBlocking) != cudaSuccess) return nfail("mgpu_nccl: stream");
    return 0;
}

int mgpu_reduce_averages(double *buf, int32_t count)
{
    if (N.nranks <= 1 || count <= 0) return 0;
    if (!N.comm) return nfail("mgpu_reduce_averages: mgpu_nccl_init was not called");
    if ((size_t)count > N.dcap) {
        if (N.dbuf) cudaFree(N.dbuf);
        if (cudaMalloc(&N.dbuf, sizeof(double) * count) != cudaSuccess) return nfail("mgpu_reduce_averages: cudaMalloc");
        N.dcap = count;
    }
    if (cudaMemcpyAsync(N.dbuf, buf, sizeof(double) * count, cudaMemcpyHostToDevice, N.stream) != cudaSuccess) return nfail("mgpu_reduce_averages: H2D");
    ncclResult_t r = N.AllReduce(N.dbuf, N.dbuf, count, ncclFloat64, ncclSum, N.comm, N.stream);
    if (r) return nfail(std::string("ncclAllReduce: ") + (N.GetErrorString ? N.GetErrorString(r) : "error"));
    if (cudaMemcpyAsync(buf, N.dbuf, sizeof(double) * count, cudaMemcpyDeviceToHost, N.stream) != cudaSuccess) return nfail("mgpu_reduce_averages: D2H");
    if (cudaStreamSynchronize(N.stream) != cudaSuccess) return nfail("mgpu_reduce_averages: sync");
    return 0;
}

void mgpu_nccl_finalize(void)
{
    if (N.comm && N.CommDestroy) N.CommDestroy(N.comm);
    N.comm = nullptr;
    if (N.dbuf) cudaFree(N.dbuf);
    N.dbuf = nullptr; N.dcap = 0;
    if (N.stream) cudaStreamDestroy(N.stream);
    N.stream = nullptr;
    N.nranks = 1; N.rank = 0;
}
}
