import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(4096, dtype=torch.uint8, device=dev)
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
    print(rank, "OK", type(hdl).__name__, [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs][:2], hdl.rank, hdl.world_size, flush=True)
    t.fill_(rank + 1)
    dist.barrier(); torch.cuda.synchronize()
    peer = hdl.get_buffer((rank + 1) % world, (4096,), torch.uint8)
    print(rank, "peer value", int(peer[0].item()), flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
dist.barrier(); dist.destroy_process_group()
