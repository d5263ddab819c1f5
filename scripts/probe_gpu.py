"""Print what the GPU box looks like (topology, peer access, multicast support, NCCL's algorithm choice)."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60).stdout
    except Exception as e:  # noqa: BLE001
        return "ERR %s" % e


print(sh("nvidia-smi -L"))
print(sh("nvidia-smi topo -m"))
print(sh("nvidia-smi --query-gpu=index,name,memory.total,clocks.sm,clocks.max.sm,power.limit --format=csv"))
print(sh("nproc; free -g | head -2; df -h /dev/shm | tail -1"))
import torch  # noqa: E402

print("torch", torch.__version__, "cuda", torch.version.cuda, "devices", torch.cuda.device_count())
n = torch.cuda.device_count()
for i in range(n):
    print(i, torch.cuda.get_device_name(i), torch.cuda.get_device_capability(i),
          [int(torch.cuda.can_device_access_peer(i, j)) for j in range(n) if j != i])
try:
    from cuda import cuda as cu
    cu.cuInit(0)
    for i in range(n):
        err, dev = cu.cuDeviceGet(i)
        for name in ("CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED", "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED",
                     "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED", "CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED"):
            attr = getattr(cu.CUdevice_attribute, name, None)
            if attr is not None:
                print(i, name, cu.cuDeviceGetAttribute(attr, dev))
except Exception as e:  # noqa: BLE001
    print("cuda-python probe failed:", e)
