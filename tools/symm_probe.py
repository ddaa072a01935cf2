"""Does torch's symmetric memory (peer-mapped buffers over NVLink) work on this box?  torchrun --nproc-per-node 2."""
import os, sys
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.uint8, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "ok", [hex(p) for p in hdl.buffer_ptrs], [hex(p) for p in hdl.signal_pad_ptrs], hdl.signal_pad_size,
          flush=True)
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (16,), torch.uint8)
    print(rank, "peer value", int(peer[0]), flush=True)
    hdl.barrier()
except Exception as e:
    print(rank, "FAILED", repr(e)[:500], flush=True)
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
os._exit(0)
