from cuda.bindings import runtime as rt
import torch
torch.cuda.init()
for name in ("cudaDevAttrL2CacheSize", "cudaDevAttrMaxPersistingL2CacheSize", "cudaDevAttrMaxAccessPolicyWindowSize", "cudaDevAttrMultiProcessorCount", "cudaDevAttrMaxSharedMemoryPerBlockOptin"):
    err, v = rt.cudaDeviceGetAttribute(getattr(rt.cudaDeviceAttr, name), 0)
    print(name, err, v)
err, v = rt.cudaDeviceGetLimit(rt.cudaLimit.cudaLimitPersistingL2CacheSize)
print("limit persisting", err, v)
import os
print("cpus", os.cpu_count())
os.system("free -g | head -2; nvidia-smi topo -m | head -12")
