"""Developer tool: TF-GridNet forward timing by batch size (4 s utterances, 257 x 256), CUDA events."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rethinking-flow-and-diffusion-bridge-models-for-speech-enhancement_b200"))
import torch
from fdbm_b200 import BackboneRegistry
dev = torch.device("cuda:0")
net = BackboneRegistry.get_by_name("tfgridnet_5l32c100")().to(dev).eval()
for B in [int(a) for a in sys.argv[1:]] or [4, 16, 32]:
    Y = torch.view_as_complex(torch.randn(B, 1, 257, 256, 2, device=dev)) * 0.3
    t = torch.full((B,), 0.5, device=dev)
    for _ in range(2):
        net(Y, Y, t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        net(Y, Y, t)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    gflop = 2 * 5 * 2 * (B * 262 * 260 + B * 263 * 259) * 2 * (228 * 400 + 100 * 128) / 1e9 / 2   # 2 dirs x (gates + deconv) MACs x 2
    print(json.dumps({"B": B, "forward_ms": ms, "audio_s_per_s_predictive": B * 4 / (ms * 1e-3), "lstm_tflops": gflop / ms}))
