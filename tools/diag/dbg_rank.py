import os, sys
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gen = torch.Generator(device="cuda"); gen.manual_seed(2002 + rank)
x = torch.rand((4, 2), dtype=torch.float64, device="cuda", generator=gen)
print("rank", rank, "local", local, "dev", torch.cuda.current_device(), x.flatten()[:3].tolist(), flush=True)
mine = torch.tensor([float(rank), 1.0], dtype=torch.float64, device="cuda")
g = [torch.zeros(2, dtype=torch.float64, device="cuda") for _ in range(dist.get_world_size())]
dist.all_gather(g, mine)
print("rank", rank, "gathered", torch.stack(g).cpu().tolist(), flush=True)
dist.destroy_process_group()
