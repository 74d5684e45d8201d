// symm.cpp -- symmetric device buffers for the GPUs of one NVSwitch box: every rank allocates the
// same number of bytes with cuMemCreate, the ranks exchange the allocations' file descriptors
// (comm.hpp), and every rank maps every peer's allocation into its own address space (peer-mapped
// pointers: loads / stores go over NVLink).  Where the box supports it (NVLS), the allocations are
// also bound to ONE multicast object whose mapping is a "multicast pointer": a multimem.st to it is
// delivered by the switch to the same offset of every rank's buffer.
//
// This is what torch.distributed._symmetric_memory did for the round-1 Python loop; here it is
// ~200 lines of CUDA driver API in the library itself, so a C / C++ caller needs no Python.  The
// driver entry points are looked up at run time (cudaGetDriverEntryPoint): no link-time dependency
// on libcuda.so, and a box with an older driver simply reports "no multicast".
#include "symm.hpp"

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

namespace spmv {
namespace b200 {
namespace {

struct Driver {
    bool ok = false;
    bool has_multicast = false;
    CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
    CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
    CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
};

template <class F>
bool load_entry(const char* name, F* fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    *fn = reinterpret_cast<F>(p);
    return true;
}

const Driver& driver() {
    static const Driver d = [] {
        Driver r;
        r.ok = load_entry("cuMemGetAllocationGranularity", &r.MemGetAllocationGranularity) &&
               load_entry("cuMemCreate", &r.MemCreate) && load_entry("cuMemRelease", &r.MemRelease) &&
               load_entry("cuMemExportToShareableHandle", &r.MemExportToShareableHandle) &&
               load_entry("cuMemImportFromShareableHandle", &r.MemImportFromShareableHandle) &&
               load_entry("cuMemAddressReserve", &r.MemAddressReserve) && load_entry("cuMemAddressFree", &r.MemAddressFree) &&
               load_entry("cuMemMap", &r.MemMap) && load_entry("cuMemUnmap", &r.MemUnmap) &&
               load_entry("cuMemSetAccess", &r.MemSetAccess) && load_entry("cuDeviceGetAttribute", &r.DeviceGetAttribute);
        r.has_multicast = r.ok && load_entry("cuMulticastCreate", &r.MulticastCreate) &&
                          load_entry("cuMulticastAddDevice", &r.MulticastAddDevice) &&
                          load_entry("cuMulticastBindMem", &r.MulticastBindMem) &&
                          load_entry("cuMulticastGetGranularity", &r.MulticastGetGranularity) &&
                          load_entry("cuMulticastUnbind", &r.MulticastUnbind);
        return r;
    }();
    return d;
}

size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

// every rank reports a status; the collective step succeeded only if it did everywhere
bool all_ok(Comm* comm, bool mine) {
    std::vector<int> all(comm->world(), 0);
    const int v = mine ? 1 : 0;
    if (comm->allgather(&v, all.data(), sizeof(int)) != 0) return false;
    for (int x : all) if (!x) return false;
    return true;
}

bool map_handle(const Driver& d, CUmemGenericAllocationHandle h, size_t size, size_t gran, int device, void** out) {
    CUdeviceptr va = 0;
    if (d.MemAddressReserve(&va, size, gran, 0, 0) != CUDA_SUCCESS) return false;
    if (d.MemMap(va, size, 0, h, 0) != CUDA_SUCCESS) {
        d.MemAddressFree(va, size);
        return false;
    }
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (d.MemSetAccess(va, size, &acc, 1) != CUDA_SUCCESS) {
        d.MemUnmap(va, size);
        d.MemAddressFree(va, size);
        return false;
    }
    *out = reinterpret_cast<void*>(va);
    return true;
}

void unmap(const Driver& d, void* p, size_t size) {
    if (!p) return;
    d.MemUnmap(reinterpret_cast<CUdeviceptr>(p), size);
    d.MemAddressFree(reinterpret_cast<CUdeviceptr>(p), size);
}

}  // namespace

bool symm_supported() { return driver().ok; }

void symm_free(SymmBuffer* b) {
    if (!b) return;
    const Driver& d = driver();
    if (d.ok) {
        if (b->mc) unmap(d, b->mc, b->mapped_bytes);
        for (int p = 0; p < kSymmMaxRanks; ++p) {
            if (b->peer[p]) unmap(d, b->peer[p], b->mapped_bytes);
            if (b->handles[p]) d.MemRelease(static_cast<CUmemGenericAllocationHandle>(b->handles[p]));
        }
        if (b->mc_handle) {
            if (b->mc_bound) d.MulticastUnbind(static_cast<CUmemGenericAllocationHandle>(b->mc_handle), b->device, 0, b->mapped_bytes);
            d.MemRelease(static_cast<CUmemGenericAllocationHandle>(b->mc_handle));
        }
    }
    *b = SymmBuffer();
}

// Collective over `comm`.  Returns 0 when every rank holds a peer-mapped buffer; b->mc is set on
// every rank or on none.  On failure nothing is left allocated (on this rank).
int symm_alloc(Comm* comm, size_t bytes, bool want_multicast, SymmBuffer* b) {
    *b = SymmBuffer();
    const Driver& d = driver();
    const int world = comm->world(), rank = comm->rank();
    int device = 0;
    bool ok = d.ok && world <= kSymmMaxRanks && cudaGetDevice(&device) == cudaSuccess && cudaFree(nullptr) == cudaSuccess;
    b->device = device;
    b->world = world;
    b->rank = rank;

    CUmemAllocationProp prop{};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 2u << 20;
    int mc_supported = 0;
    CUmulticastObjectProp mcp{};
    if (ok) {
        size_t g = 0;
        if (d.MemGetAllocationGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && g > gran) gran = g;
        if (want_multicast && d.has_multicast && world > 1 &&
            d.DeviceGetAttribute(&mc_supported, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, device) != CUDA_SUCCESS)
            mc_supported = 0;
        if (mc_supported) {
            mcp.numDevices = static_cast<unsigned>(world);
            mcp.size = round_up(bytes, gran);
            mcp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
            size_t mg = 0;
            if (d.MulticastGetGranularity(&mg, &mcp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > gran) gran = mg;
        }
    }
    const size_t size = round_up(bytes > 0 ? bytes : 1, gran);
    b->bytes = bytes;
    b->mapped_bytes = size;

    // ---- 1. allocate, export, exchange descriptors, import, map -------------------------------
    CUmemGenericAllocationHandle mine = 0;
    int my_fd = -1;
    if (ok) ok = d.MemCreate(&mine, size, &prop, 0) == CUDA_SUCCESS;
    if (ok) {
        b->handles[rank] = static_cast<unsigned long long>(mine);
        ok = d.MemExportToShareableHandle(&my_fd, mine, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) == CUDA_SUCCESS;
    }
    if (!ok && my_fd < 0) my_fd = ::open("/dev/null", O_RDONLY | O_CLOEXEC);  // keep the collective well-formed
    std::vector<int> fds(world, -1);
    if (comm->allgather_fds(my_fd, fds.data()) != 0) ok = false;
    if (my_fd >= 0) ::close(my_fd);
    ok = all_ok(comm, ok);
    for (int p = 0; p < world && ok; ++p) {
        CUmemGenericAllocationHandle h = mine;
        if (p != rank) {
            if (d.MemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fds[p])),
                                               CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS) {
                ok = false;
                break;
            }
            b->handles[p] = static_cast<unsigned long long>(h);
        }
        ok = map_handle(d, h, size, gran, device, &b->peer[p]);
    }
    for (int fd : fds) if (fd >= 0) ::close(fd);
    ok = all_ok(comm, ok);
    if (!ok) {
        symm_free(b);
        return -1;
    }
    b->local = b->peer[rank];

    // ---- 2. multicast object (NVLS): created by rank 0, every device added, every buffer bound ----
    std::vector<int> sup(world, 0);
    comm->allgather(&mc_supported, sup.data(), sizeof(int));
    bool mc_ok = true;
    for (int s : sup) mc_ok = mc_ok && s != 0;
    if (mc_ok) {
        mcp.size = size;
        CUmemGenericAllocationHandle mc = 0;
        int mc_fd = -1;
        bool step = true;
        if (rank == 0) {
            step = d.MulticastCreate(&mc, &mcp) == CUDA_SUCCESS &&
                   d.MemExportToShareableHandle(&mc_fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) == CUDA_SUCCESS;
        }
        if (mc_fd < 0) mc_fd = ::open("/dev/null", O_RDONLY | O_CLOEXEC);
        std::vector<int> mfds(world, -1);
        if (comm->allgather_fds(mc_fd, mfds.data()) != 0) step = false;
        if (mc_fd >= 0) ::close(mc_fd);
        step = all_ok(comm, step);
        if (step && rank != 0)
            step = d.MemImportFromShareableHandle(&mc, reinterpret_cast<void*>(static_cast<uintptr_t>(mfds[0])),
                                                  CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) == CUDA_SUCCESS;
        for (int fd : mfds) if (fd >= 0) ::close(fd);
        if (mc) b->mc_handle = static_cast<unsigned long long>(mc);
        if (step) step = d.MulticastAddDevice(mc, device) == CUDA_SUCCESS;
        step = all_ok(comm, step);  // binding needs every device added first
        if (step) {
            step = d.MulticastBindMem(mc, 0, mine, 0, size, 0) == CUDA_SUCCESS;
            b->mc_bound = step;
        }
        step = all_ok(comm, step);
        if (step) step = map_handle(d, mc, size, gran, device, &b->mc);
        step = all_ok(comm, step);
        if (!step) {  // fall back to peer stores everywhere
            if (b->mc) unmap(d, b->mc, size);
            b->mc = nullptr;
            if (b->mc_handle) {
                if (b->mc_bound) d.MulticastUnbind(mc, device, 0, size);
                d.MemRelease(mc);
            }
            b->mc_handle = 0;
            b->mc_bound = false;
        }
    }
    return 0;
}

}  // namespace b200
}  // namespace spmv
