// comm.cu -- the multi-GPU exchange step of the drop-in renderer: an NCCL communicator behind the C ABI.
// libnccl.so.2 is resolved at RUN time (dlopen): in a torch process that is the library torch already loaded, in the
// stand-alone CLI the system one; the build needs neither nccl.h nor -lnccl.  Only the five entry points below are used.
#include "../../../include/pbrtgpu.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstring>

#include <mutex>
#include <string>

extern "C" int pbrtgpu_set_error_text(int code, const char *text);

namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;   // NCCL_UNIQUE_ID_BYTES
enum { NCCL_SUCCESS = 0, NCCL_FLOAT = 7, NCCL_SUM = 0 };
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string why;
};
NcclApi &Nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.why = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return; }
        api.GetUniqueId = (int (*)(ncclUniqueId *))dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(api.handle, "ncclCommInitRank");
        api.CommDestroy = (int (*)(ncclComm_t))dlsym(api.handle, "ncclCommDestroy");
        api.Reduce = (int (*)(const void *, void *, size_t, int, int, int, ncclComm_t, cudaStream_t))dlsym(api.handle, "ncclReduce");
        api.GetErrorString = (const char *(*)(int))dlsym(api.handle, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Reduce) {
            api.why = "libnccl.so.2 lacks an expected entry point";
            api.handle = nullptr;
        }
    });
    return api;
}
int NcclFail(const char *what, int rc) {
    NcclApi &n = Nccl();
    std::string msg = std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(rc) : "NCCL error");
    return pbrtgpu_set_error_text(PBRTGPU_ECUDA, msg.c_str());
}
}  // namespace

struct PbrtGpuComm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
};

// (for pbrtgpu.cu) sum `count` floats of every rank's `buf` onto rank 0's `buf`, on `stream`
extern "C" int PbrtGpuCommReduceToRoot(PbrtGpuComm *c, float *buf, size_t count, cudaStream_t stream) {
    if (c->world <= 1) return PBRTGPU_OK;
    const int rc = Nccl().Reduce(buf, buf, count, NCCL_FLOAT, NCCL_SUM, 0, c->comm, stream);
    if (rc != NCCL_SUCCESS) return NcclFail("ncclReduce", rc);
    return PBRTGPU_OK;
}
extern "C" int PbrtGpuCommRank(const PbrtGpuComm *c) { return c->rank; }
extern "C" int PbrtGpuCommWorld(const PbrtGpuComm *c) { return c->world; }

extern "C" {

int pbrtgpu_comm_unique_id(void *id) {
    if (!id) return pbrtgpu_set_error_text(PBRTGPU_EINVAL, "null argument");
    NcclApi &n = Nccl();
    if (!n.handle) return pbrtgpu_set_error_text(PBRTGPU_EUNSUPPORTED, n.why.c_str());
    ncclUniqueId u;
    const int rc = n.GetUniqueId(&u);
    if (rc != NCCL_SUCCESS) return NcclFail("ncclGetUniqueId", rc);
    memcpy(id, &u, PBRTGPU_COMM_ID_BYTES);
    return PBRTGPU_OK;
}

int pbrtgpu_comm_create(const void *id, int rank, int worldSize, PbrtGpuComm **out) {
    if (!id || !out || worldSize < 1 || rank < 0 || rank >= worldSize) return pbrtgpu_set_error_text(PBRTGPU_EINVAL, "bad communicator arguments");
    *out = nullptr;
    NcclApi &n = Nccl();
    if (!n.handle) return pbrtgpu_set_error_text(PBRTGPU_EUNSUPPORTED, n.why.c_str());
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
        return pbrtgpu_set_error_text(PBRTGPU_ENODEVICE, "no CUDA device visible; libpbrtgpu has no CPU fallback");
    PbrtGpuComm *c = new PbrtGpuComm;
    c->rank = rank; c->world = worldSize;
    cudaGetDevice(&c->device);
    ncclUniqueId u;
    memcpy(&u, id, PBRTGPU_COMM_ID_BYTES);
    const int rc = n.CommInitRank(&c->comm, worldSize, u, rank);
    if (rc != NCCL_SUCCESS) { delete c; return NcclFail("ncclCommInitRank", rc); }
    *out = c;
    return PBRTGPU_OK;
}

void pbrtgpu_comm_destroy(PbrtGpuComm *comm) {
    if (!comm) return;
    if (comm->comm) Nccl().CommDestroy(comm->comm);
    delete comm;
}

}  // extern "C"
