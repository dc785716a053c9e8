// np_peer.h -- host-side plumbing for NVLink peer memory between the ranks of one node (one process per GPU).
//
// The slab-decomposed solver keeps every device array of a handle in one arena and lets its kernels store /
// load halo planes directly in the neighbours' arenas.  Two ways to map a neighbour's arena:
//   (1) CUDA virtual memory management: the arena is a cuMemCreate allocation exported as a POSIX file
//       descriptor; the descriptor travels to the neighbour over an abstract-namespace UNIX datagram socket
//       (SCM_RIGHTS) and is mapped there with cuMemImportFromShareableHandle + cuMemMap + cuMemSetAccess.
//       This is what NCCL itself does for its peer buffers (NCCL_CUMEM_ENABLE).
//   (2) legacy CUDA IPC (cudaIpcGetMemHandle / cudaIpcOpenMemHandle) on a cudaMalloc arena.
// Both work on this pool's 2-, 4- and 8-GPU boxes for 128 MB .. 2.5 GB allocations (tools/ipc_probe.cu,
// profiles/r2_ipc_probe_4gpu.txt); (1) is the default, (2) the fallback when cuMem* is unavailable.
// Driver entry points are taken from the runtime (cudaGetDriverEntryPoint): no link dependency on libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <string>

namespace b200np_peer {

struct DriverApi {
    bool ok = false;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    bool load()
    {
        if (ok) return true;
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult st;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess && *fn;
        };
        ok = get("cuMemCreate", (void**)&MemCreate) && get("cuMemRelease", (void**)&MemRelease) &&
             get("cuMemAddressReserve", (void**)&MemAddressReserve) && get("cuMemAddressFree", (void**)&MemAddressFree) &&
             get("cuMemMap", (void**)&MemMap) && get("cuMemUnmap", (void**)&MemUnmap) && get("cuMemSetAccess", (void**)&MemSetAccess) &&
             get("cuMemExportToShareableHandle", (void**)&MemExportToShareableHandle) &&
             get("cuMemImportFromShareableHandle", (void**)&MemImportFromShareableHandle) &&
             get("cuMemGetAllocationGranularity", (void**)&MemGetAllocationGranularity) && get("cuGetErrorString", (void**)&GetErrorString);
        cudaGetLastError();
        return ok;
    }
    const char* err(CUresult r) const
    {
        const char* s = nullptr;
        if (GetErrorString) GetErrorString(r, &s);
        return s ? s : "unknown driver error";
    }
};

// a cuMemCreate allocation mapped into this process: the rank's own arena, or a neighbour's
struct VmmMapping {
    CUdeviceptr ptr = 0;
    size_t size = 0;                       // mapped (granularity-rounded) size
    CUmemGenericAllocationHandle handle = 0;
    bool mapped = false;
};

inline CUmemAllocationProp vmm_prop(int device)
{
    CUmemAllocationProp p{};
    p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    p.location.id = device;
    p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return p;
}

// map `handle` (size bytes, already rounded) read/write for `device`
inline bool vmm_map(DriverApi& D, int device, VmmMapping& m, std::string* why)
{
    CUresult r = D.MemAddressReserve(&m.ptr, m.size, 0, 0, 0);
    if (r != CUDA_SUCCESS) { if (why) *why = std::string("cuMemAddressReserve: ") + D.err(r); m.ptr = 0; return false; }
    r = D.MemMap(m.ptr, m.size, 0, m.handle, 0);
    if (r != CUDA_SUCCESS) {
        if (why) *why = std::string("cuMemMap: ") + D.err(r);
        D.MemAddressFree(m.ptr, m.size); m.ptr = 0;
        return false;
    }
    CUmemAccessDesc a{};
    a.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    a.location.id = device;
    a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    r = D.MemSetAccess(m.ptr, m.size, &a, 1);
    if (r != CUDA_SUCCESS) {
        if (why) *why = std::string("cuMemSetAccess: ") + D.err(r);
        D.MemUnmap(m.ptr, m.size); D.MemAddressFree(m.ptr, m.size); m.ptr = 0;
        return false;
    }
    m.mapped = true;
    return true;
}

// allocate `bytes` on `device` as an exportable allocation and map it locally
inline bool vmm_alloc(DriverApi& D, int device, size_t bytes, VmmMapping& m, std::string* why)
{
    const CUmemAllocationProp p = vmm_prop(device);
    size_t gran = 0;
    CUresult r = D.MemGetAllocationGranularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    if (r != CUDA_SUCCESS || gran == 0) { if (why) *why = std::string("cuMemGetAllocationGranularity: ") + D.err(r); return false; }
    m.size = (bytes + gran - 1) / gran * gran;
    r = D.MemCreate(&m.handle, m.size, &p, 0);
    if (r != CUDA_SUCCESS) { if (why) *why = std::string("cuMemCreate: ") + D.err(r); m.handle = 0; return false; }
    if (!vmm_map(D, device, m, why)) { D.MemRelease(m.handle); m.handle = 0; return false; }
    return true;
}

inline void vmm_free(DriverApi& D, VmmMapping& m)
{
    if (m.mapped) { D.MemUnmap(m.ptr, m.size); D.MemAddressFree(m.ptr, m.size); }
    if (m.handle) D.MemRelease(m.handle);
    m = VmmMapping{};
}

inline int vmm_export_fd(DriverApi& D, const VmmMapping& m, std::string* why)
{
    int fd = -1;
    CUresult r = D.MemExportToShareableHandle(&fd, m.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) { if (why) *why = std::string("cuMemExportToShareableHandle: ") + D.err(r); return -1; }
    return fd;
}

// import a neighbour's allocation from a received file descriptor and map it for `device`
inline bool vmm_import(DriverApi& D, int device, int fd, size_t size, VmmMapping& m, std::string* why)
{
    CUresult r = D.MemImportFromShareableHandle(&m.handle, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    if (r != CUDA_SUCCESS) { if (why) *why = std::string("cuMemImportFromShareableHandle: ") + D.err(r); m.handle = 0; return false; }
    m.size = size;
    if (!vmm_map(D, device, m, why)) { D.MemRelease(m.handle); m.handle = 0; return false; }
    return true;
}

// ---- file descriptors between processes: abstract-namespace UNIX datagram sockets + SCM_RIGHTS ----
inline void fd_addr(sockaddr_un& a, socklen_t& len, long long pid, unsigned long long serial)
{
    memset(&a, 0, sizeof(a));
    a.sun_family = AF_UNIX;
    char name[64];
    const int n = snprintf(name, sizeof(name), "b200np.%lld.%llu", pid, serial);
    a.sun_path[0] = '\0';                       // abstract namespace: no file system entry, gone with the socket
    memcpy(a.sun_path + 1, name, n);
    len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
}
inline int fd_socket_bind(long long pid, unsigned long long serial)
{
    int s = socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
    if (s < 0) return -1;
    sockaddr_un a; socklen_t len;
    fd_addr(a, len, pid, serial);
    if (bind(s, (sockaddr*)&a, len) != 0) { close(s); return -1; }
    timeval tv{20, 0};                          // a lost peer must not hang the rank forever
    setsockopt(s, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    return s;
}
inline bool fd_send(int sock, long long to_pid, unsigned long long to_serial, int fd, int tag)
{
    sockaddr_un a; socklen_t len;
    fd_addr(a, len, to_pid, to_serial);
    msghdr msg{};
    iovec io{&tag, sizeof(tag)};
    alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int))];
    memset(ctl, 0, sizeof(ctl));
    msg.msg_name = &a; msg.msg_namelen = len;
    msg.msg_iov = &io; msg.msg_iovlen = 1;
    msg.msg_control = ctl; msg.msg_controllen = sizeof(ctl);
    cmsghdr* c = CMSG_FIRSTHDR(&msg);
    c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(c), &fd, sizeof(int));
    return sendmsg(sock, &msg, 0) == (ssize_t)sizeof(tag);
}
// returns the received descriptor (>= 0) and the sender's tag, or -1
inline int fd_recv(int sock, int* tag)
{
    msghdr msg{};
    int t = -1;
    iovec io{&t, sizeof(t)};
    alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int))];
    memset(ctl, 0, sizeof(ctl));
    msg.msg_iov = &io; msg.msg_iovlen = 1;
    msg.msg_control = ctl; msg.msg_controllen = sizeof(ctl);
    if (recvmsg(sock, &msg, 0) != (ssize_t)sizeof(t)) return -1;
    cmsghdr* c = CMSG_FIRSTHDR(&msg);
    if (!c || c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS) return -1;
    int fd = -1;
    memcpy(&fd, CMSG_DATA(c), sizeof(int));
    *tag = t;
    return fd;
}

}  // namespace b200np_peer
