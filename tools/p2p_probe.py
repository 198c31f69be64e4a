"""2-GPU probe: map the neighbour's tensor through torch's CUDA-IPC tensor sharing and write into it from a kernel
running on THIS device (fdtd2d_efield with ez in peer memory)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.multiprocessing.reductions import reduce_tensor
from simulation_b200 import fd2d, _lib
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank, world = dist.get_rank(), dist.get_world_size()
n = 256
mine = torch.zeros((n, n), dtype=torch.float32, device="cuda")
allh = [None] * world
dist.all_gather_object(allh, reduce_tensor(mine))
fn, args = allh[(rank + 1) % world]
peer = fn(*args)
print(rank, "peer tensor on", peer.device, "ptr", hex(peer.data_ptr()), "can access:",
      torch.cuda.can_device_access_peer(local, peer.device.index), flush=True)
_lib.check(_lib.lib().fdtd_enable_peer_access(int(peer.device.index)), "enable")
naz = torch.full((n, n), 2.0, device="cuda"); dz = torch.full((n, n), float(rank + 1), device="cuda")
fd2d.efield(n, n, naz, dz, dz.new_empty((n, n)))          # local warm-up
md = _lib.Medium2D(naz.data_ptr(), None)
import ctypes as C
rc = _lib.lib().fdtd2d_efield(_lib.F32, n, n, C.byref(md), C.c_void_p(dz.data_ptr()), None, C.c_void_p(peer.data_ptr()),
                              C.c_void_p(torch.cuda.current_stream().cuda_stream))
print(rank, "launch rc", rc, flush=True)
torch.cuda.synchronize()
dist.barrier()
print(rank, "my tensor now holds", float(mine[0, 0]), "(expected", 2.0 * (((rank - 1) % world) + 1), ")", flush=True)
dist.barrier()
dist.destroy_process_group()
