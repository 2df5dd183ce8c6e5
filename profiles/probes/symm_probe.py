import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
print(rank, "can access peer", [torch.cuda.can_device_access_peer(local, p) for p in range(world) if p != local], flush=True)
try:
    buf = symm_mem.empty((world, 1024), dtype=torch.float64, device=torch.device("cuda", local))
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD.group_name)
    print(rank, "rendezvous ok", type(hdl).__name__, [hex(p) for p in hdl.buffer_ptrs][:4], len(hdl.signal_pad_ptrs), hdl.signal_pad_size if hasattr(hdl,'signal_pad_size') else None, flush=True)
    buf.fill_(rank + 1)
    hdl.barrier(channel=0)
    peer = hdl.get_buffer((rank + 1) % world, (world, 1024), torch.float64)
    print(rank, "peer value", float(peer[0, 0].item()), flush=True)
    hdl.barrier(channel=0)
    print(rank, [a for a in dir(hdl) if not a.startswith('_')], flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
dist.barrier(); dist.destroy_process_group()
