"""Probe: does this node expose NVLS multicast (NCCL log + torch symmetric memory + driver attribute)?"""
import os

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
x = torch.ones(1 << 24, device="cuda")
dist.all_reduce(x)
torch.cuda.synchronize()
try:
    from cuda.bindings import driver as cu
except Exception:  # older cuda-python layout
    from cuda import cuda as cu
cu.cuInit(0)
err, dev = cu.cuDeviceGet(torch.cuda.current_device())
err, mc = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev)
err, fab = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev)
err, pfd = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev)
if rank == 0:
    print(f"PROBE multicast_supported={mc} fabric_handles={fab} posix_fd_handles={pfd}")
try:
    import torch.distributed._symmetric_memory as symm_mem

    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD.group_name)
    if rank == 0:
        print(f"PROBE symm_mem ok: world={hdl.world_size} multicast_ptr={hdl.multicast_ptr:#x} "
              f"buffer_ptrs={[hex(p) for p in hdl.buffer_ptrs]} signal_pad_ptrs={len(hdl.signal_pad_ptrs)}")
except Exception as e:  # noqa: BLE001
    if rank == 0:
        print(f"PROBE symm_mem failed: {type(e).__name__}: {e}")
dist.barrier()
os._exit(0)
