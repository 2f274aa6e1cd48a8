"""Developer tool: per-launch timing of one backbone forward (CUDA events around every kernel)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rethinking-flow-and-diffusion-bridge-models-for-speech-enhancement_b200"))
import torch
from fdbm_b200 import BackboneRegistry, sensitise_
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
net = sensitise_(BackboneRegistry.get_by_name("ncsnpp_v2")(), 0).cuda().eval()
x = torch.view_as_complex(torch.randn(B, 1, 257, T, 2, device="cuda"))
t = torch.full((B,), 0.5, device="cuda")
for _ in range(3): prof = net.profile_forward(x, x, t)
names = {0: "conv", 1: "gn", 2: "stats", 3: "skinny", 4: "attn", 5: "temb"}
tot = sum(p[0] for p in prof)
print(f"B={B} T={T}: {len(prof)} launches, {tot:.3f} ms")
agg = {}
for ms, k, fl in prof:
    agg.setdefault(k, [0, 0.0, 0.0]); agg[k][0] += 1; agg[k][1] += ms; agg[k][2] += fl
for k, (n, ms, fl) in sorted(agg.items()):
    print(f"  {names[k]:7s} n={n:3d} {ms:8.3f} ms {100*ms/tot:5.1f}%" + (f"  {fl/ms/1e9:7.1f} TFLOP/s" if fl else ""))
print("convs by time:")
convs = sorted([(ms, fl, i) for i, (ms, k, fl) in enumerate(prof) if k == 0], reverse=True)
for ms, fl, i in convs[:40]:
    print(f"   op {i:3d}: {ms*1e3:8.1f} us  {fl/1e9:7.2f} GFLOP  {fl/ms/1e9:7.1f} TFLOP/s")
print("non-conv ops > 60 us:")
for i, (ms, k, fl) in enumerate(prof):
    if k != 0 and ms > 0.06: print(f"   op {i:3d} {names[k]:7s} {ms*1e3:8.1f} us")
# where the conv time goes relative to the sustained tensor peak: buckets by algorithmic GFLOP of the launch
PEAK = 1398.0
buckets = {}
for ms, fl, i in convs:
    g = fl / 1e9
    key = ">=1000" if g >= 1000 else (">=300" if g >= 300 else (">=100" if g >= 100 else (">=30" if g >= 30 else (">=10" if g >= 10 else "<10"))))
    b = buckets.setdefault(key, [0, 0.0, 0.0]); b[0] += 1; b[1] += ms; b[2] += fl
conv_ms = sum(c[0] for c in convs)
print("conv launches by size (GFLOP per launch): count, ms, share of conv time, TFLOP/s, ms lost against the sustained peak")
for key in (">=1000", ">=300", ">=100", ">=30", ">=10", "<10"):
    if key in buckets:
        n, ms, fl = buckets[key]
        print(f"   {key:7s} n={n:3d} {ms:8.3f} ms {100*ms/conv_ms:5.1f}%  {fl/ms/1e9:7.1f} TFLOP/s  lost {ms - fl/1e9/PEAK:7.3f} ms")
