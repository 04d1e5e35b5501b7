"""Does torch symmetric memory (peer-mapped buffers + signal pads) work on this box?  (development probe, 2+ GPUs under torchrun)"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
t = symm.empty(1024, dtype=torch.int32, device=dev)
t.fill_(-1)
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok: world", hdl.world_size, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
hdl.barrier(channel=0)
peer = (rank + 1) % world
pb = hdl.get_buffer(peer, (1024,), torch.int32)
pb[rank * 4:(rank + 1) * 4] = rank + 100            # a kernel on this GPU storing into the peer's memory
hdl.barrier(channel=0)
torch.cuda.synchronize()
src = (rank - 1) % world
print(rank, "got from", src, t[src * 4:(src + 1) * 4].tolist(), flush=True)
assert t[src * 4:(src + 1) * 4].tolist() == [src + 100] * 4
dist.barrier()
dist.destroy_process_group()
