"""Probe (2+ GPUs): does torch's symmetric memory rendezvous work on this box, and does the NVSwitch offer multicast (NVLS)?
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/probe_symm.py"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
t.fill_(rank + 1)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
info = dict(rank=rank, world=hdl.world_size, buffer_ptrs=[hex(int(p)) for p in hdl.buffer_ptrs], multicast_ptr=hex(int(hdl.multicast_ptr or 0)),
            has_multicast=getattr(hdl, "has_multicast_support", None), signal_pad_size=hdl.signal_pad_size, local_ptr=hex(t.data_ptr()))
try:
    info["has_multicast"] = bool(type(hdl).has_multicast_support(torch._C._distributed_c10d.DeviceType.CUDA if hasattr(torch._C._distributed_c10d, "DeviceType") else dev.type, local))
except Exception as e:
    info["has_multicast"] = f"? ({type(e).__name__})"
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (8,), torch.float32)
info["peer_value"] = float(peer[0])
print(info, flush=True)
dist.barrier()
dist.destroy_process_group()
