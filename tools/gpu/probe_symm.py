import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
try:
    t = sm.empty(1 << 20, dtype=torch.uint8, device=dev)
    h = sm.rendezvous(t, dist.group.WORLD)
    print(rank, "multicast_ptr", hex(h.multicast_ptr), "buffers", [hex(p) for p in h.buffer_ptrs][:3], "signal", [hex(p) for p in h.signal_pad_ptrs][:2], "world", h.world_size, flush=True)
except Exception as e:
    print(rank, "symm_mem failed:", repr(e)[:400], flush=True)
from cuda.bindings import driver as cu
try:
    err, v = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, lr)
    print(rank, "MULTICAST_SUPPORTED", err, v, flush=True)
    err, v = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, lr)
    print(rank, "FABRIC_HANDLE", err, v, flush=True)
except Exception as e:
    print("attr query failed", repr(e)[:200])
dist.barrier(); dist.destroy_process_group()
