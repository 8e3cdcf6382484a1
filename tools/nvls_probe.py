"""Is in-switch (NVLS / multimem) reduction available to a user process on this box?  Queries CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED and tries to
create a multicast object over all visible GPUs with a POSIX-fd shareable handle (what a multi-process tensor-parallel group would need), bind
one allocation per device and map it.  Pure cuda-python driver calls, no kernels."""
import sys

from cuda import cuda


def ck(res, what):
    err = res[0]
    if err != cuda.CUresult.CUDA_SUCCESS:
        name = cuda.cuGetErrorName(err)[1]
        print(f"FAIL {what}: {name}")
        sys.exit(0)
    return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)


ck(cuda.cuInit(0), "cuInit")
n = ck(cuda.cuDeviceGetCount(), "count")
print("devices:", n)
devs, ctxs = [], []
for i in range(n):
    d = ck(cuda.cuDeviceGet(i), "get")
    mc = ck(cuda.cuDeviceGetAttribute(cuda.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, d), "attr")
    fd = ck(cuda.cuDeviceGetAttribute(cuda.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, d), "attr")
    print(f"device {i}: multicast_supported={mc} posix_fd_handles={fd}")
    devs.append(d)
    ctxs.append(ck(cuda.cuDevicePrimaryCtxRetain(d), "ctx"))
if n < 2:
    print("need >= 2 GPUs for a multicast group"); sys.exit(0)
prop = cuda.CUmulticastObjectProp()
prop.numDevices = n
prop.handleTypes = cuda.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
prop.flags = 0
prop.size = 2 << 20
gran = ck(cuda.cuMulticastGetGranularity(prop, cuda.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED), "granularity")
print("recommended granularity:", gran)
prop.size = max(int(gran), 2 << 20)
mc = ck(cuda.cuMulticastCreate(prop), "cuMulticastCreate")
print("multicast object created")
for d in devs:
    ck(cuda.cuMulticastAddDevice(mc, d), "cuMulticastAddDevice")
print("all devices added")
fdh = ck(cuda.cuMemExportToShareableHandle(mc, cuda.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "export fd")
print("exported as fd", int(fdh))
for i, d in enumerate(devs):
    ck(cuda.cuCtxSetCurrent(ctxs[i]), "setctx")
    ap = cuda.CUmemAllocationProp()
    ap.type = cuda.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    ap.location.type = cuda.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    ap.location.id = i
    ap.requestedHandleTypes = cuda.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    h = ck(cuda.cuMemCreate(prop.size, ap, 0), "cuMemCreate")
    ck(cuda.cuMulticastBindMem(mc, 0, h, 0, prop.size, 0), f"cuMulticastBindMem dev {i}")
print("memory bound on every device")
ck(cuda.cuCtxSetCurrent(ctxs[0]), "setctx")
va = ck(cuda.cuMemAddressReserve(prop.size, int(gran), 0, 0), "reserve")
ck(cuda.cuMemMap(va, prop.size, 0, mc, 0), "map multicast")
ad = cuda.CUmemAccessDesc()
ad.location.type = cuda.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
ad.location.id = 0
ad.flags = cuda.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
ck(cuda.cuMemSetAccess(va, prop.size, [ad], 1), "set access")
print("NVLS_OK: multicast address mapped at", hex(int(va)))
