"""Developer tool: per-op timing of one training backward (CUDA events around every recorded op)."""
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rethinking-flow-and-diffusion-bridge-models-for-speech-enhancement_b200"))
import torch
from fdbm_b200 import BackboneRegistry, Bridge, SpecsDataModule, sensitise_, _lib
from fdbm_b200.training import TrainStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
net = sensitise_(BackboneRegistry.get_by_name("ncsnpp_v2")(), 0).cuda()
dm = SpecsDataModule(n_fft=512, hop_length=256, window="sqrthann")
ts = TrainStep(net, Bridge("sb"), dm, batch=B, n_frames=256, loss_scale=1024.0)
x = torch.view_as_complex(torch.randn(B, 1, 257, 256, 2, device="cuda")) * 0.1
y = x + 0.05 * torch.view_as_complex(torch.randn(B, 1, 257, 256, 2, device="cuda"))
for _ in range(2): ts.loss_and_backward(x, y)
g = torch.view_as_real(ts._keep)
n_max = 8192
ms = (C.c_float * n_max)(); kinds = (C.c_int * n_max)()
n = ts.lib.fdbm_plan_profile_backward(ts.plan, g.data_ptr(), 1024.0, ms, kinds, n_max, torch.cuda.current_stream().cuda_stream)
assert n > 0, ts.lib.fdbm_last_error()
names = {0: "dgrad conv", 1: "norm/elementwise", 3: "skinny", 4: "attn", 5: "small", 6: "wgrad"}
agg = {}
for i in range(n):
    a = agg.setdefault(kinds[i], [0, 0.0]); a[0] += 1; a[1] += ms[i]
tot = sum(v[1] for v in agg.values())
print(f"B={B}: backward {n} ops, {tot:.3f} ms")
for k, (c, t) in sorted(agg.items()): print(f"  {names.get(k, k):18s} n={c:4d} {t:8.3f} ms {100*t/tot:5.1f}%")
top = sorted(range(n), key=lambda i: -ms[i])[:25]
for i in top: print(f"   op {i:4d} {names.get(kinds[i], kinds[i]):18s} {ms[i]*1e3:8.1f} us")
