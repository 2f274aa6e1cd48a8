"""Developer tool (torchrun, N GPUs): two DDP training steps; checks that every rank ends with bit-identical parameters and
that the all-reduced gradient equals the mean of the per-rank gradients."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rethinking-flow-and-diffusion-bridge-models-for-speech-enhancement_b200"))
import torch, torch.distributed as dist
from fdbm_b200 import BackboneRegistry, Bridge, SpecsDataModule, sensitise_
from fdbm_b200.training import TrainStep
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
net = sensitise_(BackboneRegistry.get_by_name("ncsnpp_v2")(), 0).to(dev)
dm = SpecsDataModule(n_fft=512, hop_length=256, window="sqrthann")
ts = TrainStep(net, Bridge("sb"), dm, batch=2, n_frames=64, loss_scale=1024.0, lr=1e-3)
g = torch.Generator(device=dev).manual_seed(100 + rank)                     # every rank has its own data
x = torch.view_as_complex(torch.randn(2, 1, 257, 64, 2, device=dev, generator=g)) * 0.1
y = x + 0.05 * torch.view_as_complex(torch.randn(2, 1, 257, 64, 2, device=dev, generator=g))
for step in range(2):
    loss = ts.loss_and_backward(x, y)
    local_g = ts.flat_grads.clone()
    ts.optimizer_step()                                                      # all-reduce inside
    summed = ts.flat_grads.clone()
    ref = local_g.clone(); dist.all_reduce(ref)
    assert torch.equal(summed, ref), "flat gradient buffer is not the sum over ranks"
p = ts.flat_params.clone()
p0 = p.clone(); dist.broadcast(p0, 0)
same = torch.equal(p, p0)
print(f"rank {rank}: loss {float(loss):.4f}, params identical to rank 0: {same}, |grad| {float(summed.norm()):.4e}")
assert same
dist.destroy_process_group()
