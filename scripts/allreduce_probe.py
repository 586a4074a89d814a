"""All-reduce of the flat gradient block alone (127 MB fp32), under torchrun: prints ms, algorithm / bus bandwidth.  NCCL reads its
tuning variables (NCCL_ALGO, NCCL_PROTO, NCCL_MIN_NCHANNELS, ...) when the communicator is created, so run once per setting."""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n = 31744480
g = torch.ones(n, device='cuda')
for _ in range(5):
    dist.all_reduce(g)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    dist.all_reduce(g)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 20], device='cuda'); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    b = n * 4
    print('%-60s %.3f ms  algbw %.0f GB/s  busbw %.0f GB/s' % (os.environ.get('TAG', ''), ms.item(), b / ms.item() / 1e6, b / ms.item() / 1e6 * 2 * (world - 1) / world), flush=True)
dist.destroy_process_group()
