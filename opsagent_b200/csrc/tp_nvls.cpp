// tp_nvls.cpp — NVLink-switch multicast memory for the tensor-parallel all-reduce (in-switch reduction, "NVLS").
//
// One multicast object spans the t GPUs of a tensor-parallel group; every rank binds one physical allocation of its own device to it and
// maps (a) that allocation at a unicast address (plain loads/stores hit local HBM) and (b) the multicast object at a second address:
// `multimem.ld_reduce` on it returns the SUM of the t bound copies, reduced inside the NVSwitch, `multimem.st` writes all t copies.
// The all-reduce kernel (tp_comm.cu) then moves 2/t of a partial per rank instead of t-1 partials.
//
// Ranks are separate processes, so the multicast handle travels as a POSIX file descriptor: the leader creates the object, exports it and
// hands the fd to each follower over an abstract-namespace unix socket (SCM_RIGHTS); stage counters in the group's shm segment keep the
// ranks in lockstep (every device must be added before anybody binds) and carry the verdict: if ANY rank fails at ANY stage, ALL ranks fall
// back to the peer-memory all-reduce.  The CUDA driver entry points are resolved through the runtime (no link-time libcuda dependency).
#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>

#include "tp_comm.hpp"

namespace oa {

namespace {
template <typename Fn>
Fn drv(const char* name) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<Fn>(p);
}
struct Drv {
    decltype(&cuMulticastCreate) mcCreate = drv<decltype(&cuMulticastCreate)>("cuMulticastCreate");
    decltype(&cuMulticastAddDevice) mcAddDevice = drv<decltype(&cuMulticastAddDevice)>("cuMulticastAddDevice");
    decltype(&cuMulticastBindMem) mcBindMem = drv<decltype(&cuMulticastBindMem)>("cuMulticastBindMem");
    decltype(&cuMulticastGetGranularity) mcGran = drv<decltype(&cuMulticastGetGranularity)>("cuMulticastGetGranularity");
    decltype(&cuMemCreate) memCreate = drv<decltype(&cuMemCreate)>("cuMemCreate");
    decltype(&cuMemRelease) memRelease = drv<decltype(&cuMemRelease)>("cuMemRelease");
    decltype(&cuMemExportToShareableHandle) memExport = drv<decltype(&cuMemExportToShareableHandle)>("cuMemExportToShareableHandle");
    decltype(&cuMemImportFromShareableHandle) memImport = drv<decltype(&cuMemImportFromShareableHandle)>("cuMemImportFromShareableHandle");
    decltype(&cuMemAddressReserve) addrReserve = drv<decltype(&cuMemAddressReserve)>("cuMemAddressReserve");
    decltype(&cuMemAddressFree) addrFree = drv<decltype(&cuMemAddressFree)>("cuMemAddressFree");
    decltype(&cuMemMap) memMap = drv<decltype(&cuMemMap)>("cuMemMap");
    decltype(&cuMemUnmap) memUnmap = drv<decltype(&cuMemUnmap)>("cuMemUnmap");
    decltype(&cuMemSetAccess) setAccess = drv<decltype(&cuMemSetAccess)>("cuMemSetAccess");
    decltype(&cuDeviceGet) devGet = drv<decltype(&cuDeviceGet)>("cuDeviceGet");
    decltype(&cuDeviceGetAttribute) devAttr = drv<decltype(&cuDeviceGetAttribute)>("cuDeviceGetAttribute");
    bool ok() const { return mcCreate && mcAddDevice && mcBindMem && mcGran && memCreate && memRelease && memExport && memImport && addrReserve && addrFree && memMap && memUnmap && setAccess && devGet && devAttr; }
};

sockaddr_un abstract_addr(const std::string& name, socklen_t& len) {
    sockaddr_un a{}; a.sun_family = AF_UNIX;
    const size_t n = std::min(name.size(), sizeof(a.sun_path) - 2);
    std::memcpy(a.sun_path + 1, name.data(), n);          // sun_path[0] == 0: abstract namespace, nothing to unlink
    len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
    return a;
}
bool send_fd(int sock, int fd) {
    char byte = 'F'; iovec iov{&byte, 1};
    alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int))] = {};
    msghdr m{}; m.msg_iov = &iov; m.msg_iovlen = 1; m.msg_control = ctl; m.msg_controllen = sizeof ctl;
    cmsghdr* c = CMSG_FIRSTHDR(&m); c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int));
    std::memcpy(CMSG_DATA(c), &fd, sizeof(int));
    return sendmsg(sock, &m, 0) == 1;
}
int recv_fd(int sock) {
    char byte = 0; iovec iov{&byte, 1};
    alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int))] = {};
    msghdr m{}; m.msg_iov = &iov; m.msg_iovlen = 1; m.msg_control = ctl; m.msg_controllen = sizeof ctl;
    if (recvmsg(sock, &m, 0) != 1) return -1;
    cmsghdr* c = CMSG_FIRSTHDR(&m);
    if (!c || c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS) return -1;
    int fd = -1; std::memcpy(&fd, CMSG_DATA(c), sizeof(int));
    return fd;
}
}  // namespace

// stage barrier over the shm counters: every rank reports ok / not ok, everybody learns whether ALL were ok
static bool nvls_stage(TpShm* shm, int stage, int t, bool ok) {
    if (!ok) shm->nvls_fail.fetch_add(1, std::memory_order_acq_rel);
    shm->nvls_stage[stage].fetch_add(1, std::memory_order_acq_rel);
    const auto t0 = std::chrono::steady_clock::now();
    while (shm->nvls_stage[stage].load(std::memory_order_acquire) < (uint32_t)t) {
        std::this_thread::sleep_for(std::chrono::microseconds(200));
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 60.0) return false;
    }
    return shm->nvls_fail.load(std::memory_order_acquire) == 0;
}

bool TpComm::nvls_setup(size_t bytes_wanted, uint64_t nonce) {
    static const Drv D;
    const int t = t_, rank = rank_;
    bool ok = D.ok();
    int dev_ord = 0; cudaGetDevice(&dev_ord);
    CUdevice dev = 0; int supported = 0;
    if (ok) ok = D.devGet(&dev, dev_ord) == CUDA_SUCCESS && D.devAttr(&supported, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS && supported;
    CUmulticastObjectProp mp{}; mp.numDevices = (unsigned)t; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR; mp.flags = 0; mp.size = 2u << 20;
    size_t gran = 2u << 20;
    if (ok) ok = D.mcGran(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && gran > 0;
    const size_t size = (bytes_wanted + gran - 1) / gran * gran;
    mp.size = size;
    CUmemGenericAllocationHandle mc = 0; bool have_mc = false;
    const std::string sock_name = "oa_tp_nvls" + shm_name_ + "_" + std::to_string((unsigned long long)nonce) + "_" + std::to_string((long long)shm_->leader_pid.load());

    // ---- stage 0: the multicast object exists in every process ----
    std::thread server;
    if (rank == 0) {
        int fd = -1, lsock = -1;
        if (ok) ok = D.mcCreate(&mc, &mp) == CUDA_SUCCESS;
        have_mc = ok;
        if (ok) ok = D.memExport(&fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) == CUDA_SUCCESS && fd >= 0;
        if (ok) {
            lsock = socket(AF_UNIX, SOCK_STREAM, 0);
            socklen_t len; sockaddr_un a = abstract_addr(sock_name, len);
            ok = lsock >= 0 && bind(lsock, reinterpret_cast<sockaddr*>(&a), len) == 0 && listen(lsock, t) == 0;
        }
        shm_->nvls_leader_ready.store(ok ? 1u : 2u, std::memory_order_release);
        if (ok) server = std::thread([lsock, fd, t] {          // hand the fd to each follower, then drop our copies
            for (int i = 1; i < t; ++i) {
                pollfd p{lsock, POLLIN, 0};
                if (poll(&p, 1, 60000) <= 0) break;
                const int c = accept(lsock, nullptr, nullptr);
                if (c < 0) break;
                send_fd(c, fd); close(c);
            }
            close(lsock); close(fd);
        });
        else { if (lsock >= 0) close(lsock); if (fd >= 0) close(fd); }
    } else {
        const auto t0 = std::chrono::steady_clock::now();
        uint32_t st = 0;
        while ((st = shm_->nvls_leader_ready.load(std::memory_order_acquire)) == 0 &&
               std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 60.0) std::this_thread::sleep_for(std::chrono::microseconds(200));
        if (st != 1) ok = false;
        if (ok) {
            const int s = socket(AF_UNIX, SOCK_STREAM, 0);
            socklen_t len; sockaddr_un a = abstract_addr(sock_name, len);
            int fd = -1;
            if (s >= 0 && connect(s, reinterpret_cast<sockaddr*>(&a), len) == 0) fd = recv_fd(s);
            if (s >= 0) close(s);
            ok = fd >= 0 && D.memImport(&mc, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) == CUDA_SUCCESS;
            have_mc = ok;
            if (fd >= 0) close(fd);
        }
    }
    if (ok) ok = D.mcAddDevice(mc, dev) == CUDA_SUCCESS;
    const bool all0 = nvls_stage(shm_, 0, t, ok);
    if (server.joinable()) server.join();
    // ---- stage 1: every device is part of the object -> bind local memory, map both views ----
    CUmemGenericAllocationHandle mem = 0; bool have_mem = false; CUdeviceptr va_uc = 0, va_mc = 0; bool mapped_uc = false, mapped_mc = false;
    bool ok1 = all0;
    if (ok1) {
        CUmemAllocationProp ap{}; ap.type = CU_MEM_ALLOCATION_TYPE_PINNED; ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ap.location.id = dev_ord;
        ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        ok1 = D.memCreate(&mem, size, &ap, 0) == CUDA_SUCCESS; have_mem = ok1;
        if (ok1) ok1 = D.mcBindMem(mc, 0, mem, 0, size, 0) == CUDA_SUCCESS;
        CUmemAccessDesc ad{}; ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ad.location.id = dev_ord; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        if (ok1) ok1 = D.addrReserve(&va_uc, size, gran, 0, 0) == CUDA_SUCCESS;
        if (ok1) { ok1 = D.memMap(va_uc, size, 0, mem, 0) == CUDA_SUCCESS; mapped_uc = ok1; }
        if (ok1) ok1 = D.setAccess(va_uc, size, &ad, 1) == CUDA_SUCCESS;
        if (ok1) ok1 = D.addrReserve(&va_mc, size, gran, 0, 0) == CUDA_SUCCESS;
        if (ok1) { ok1 = D.memMap(va_mc, size, 0, mc, 0) == CUDA_SUCCESS; mapped_mc = ok1; }
        if (ok1) ok1 = D.setAccess(va_mc, size, &ad, 1) == CUDA_SUCCESS;
        if (ok1) ok1 = cudaMemset(reinterpret_cast<void*>(va_uc), 0, size) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
    }
    const bool all1 = nvls_stage(shm_, 1, t, ok1);
    if (!all1) {          // somebody failed: nobody uses it
        if (mapped_mc) D.memUnmap(va_mc, size);
        if (va_mc) D.addrFree(va_mc, size);
        if (mapped_uc) D.memUnmap(va_uc, size);
        if (va_uc) D.addrFree(va_uc, size);
        if (have_mem) D.memRelease(mem);
        if (have_mc) D.memRelease(mc);
        if (rank == 0) std::fprintf(stderr, "opsagent_b200: NVLS multicast not available for this tensor-parallel group — using the peer-memory all-reduce\n");
        return false;
    }
    nvls_uc_ = reinterpret_cast<void*>(va_uc); nvls_mc_ = reinterpret_cast<void*>(va_mc); nvls_bytes_ = size;
    nvls_mem_ = (unsigned long long)mem; nvls_obj_ = (unsigned long long)mc;
    return true;
}

void TpComm::nvls_teardown() {
    if (!nvls_uc_) return;
    static const Drv D;
    if (!D.ok()) return;
    D.memUnmap(reinterpret_cast<CUdeviceptr>(nvls_mc_), nvls_bytes_); D.addrFree(reinterpret_cast<CUdeviceptr>(nvls_mc_), nvls_bytes_);
    D.memUnmap(reinterpret_cast<CUdeviceptr>(nvls_uc_), nvls_bytes_); D.addrFree(reinterpret_cast<CUdeviceptr>(nvls_uc_), nvls_bytes_);
    D.memRelease((CUmemGenericAllocationHandle)nvls_mem_); D.memRelease((CUmemGenericAllocationHandle)nvls_obj_);
    nvls_uc_ = nvls_mc_ = nullptr;
}

}  // namespace oa
