"""One-off probe (gpurun): does this box expose NVLink multicast (NVLS) to user code, and what does the fabric look
like? Writes gpurun_out/<tag>_multicast_probe.txt. Driver API through ctypes — no kernels."""
import ctypes
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "probe"
out = open("gpurun_out/%s_multicast_probe.txt" % tag, "w")


def p(*a):
    print(*a, file=out, flush=True)
    print(*a, flush=True)


cu = ctypes.CDLL("libcuda.so.1")
p("cuInit", cu.cuInit(0))
n = ctypes.c_int()
cu.cuDeviceGetCount(ctypes.byref(n))
p("devices", n.value)
ATTR = {"MULTICAST_SUPPORTED": 132, "VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED": 102,
        "HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED": 103, "HANDLE_TYPE_FABRIC_SUPPORTED": 128,
        "GPU_DIRECT_RDMA_WITH_CUDA_VMM_SUPPORTED": 116, "MEMORY_POOLS_SUPPORTED": 115}
for d in range(n.value):
    dev = ctypes.c_int()
    cu.cuDeviceGet(ctypes.byref(dev), d)
    for k, a in ATTR.items():
        v = ctypes.c_int(-1)
        rc = cu.cuDeviceGetAttribute(ctypes.byref(v), a, dev)
        p("dev", d, k, "rc", rc, "value", v.value)
for cmd in (["nvidia-smi", "topo", "-m"], ["nvidia-smi", "nvlink", "-s", "-i", "0"], ["nvidia-smi", "-q", "-i", "0", "-d", "COMPUTE"],
            ["bash", "-c", "ls /dev/nvidia* ; ls /dev/nvidia-caps* 2>/dev/null; cat /proc/driver/nvidia/version; nproc; free -g | head -2"]):
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=30)
        p("$", " ".join(cmd)); p(r.stdout[-3000:]); p(r.stderr[-500:])
    except Exception as e:  # noqa: BLE001
        p("failed", cmd, e)
