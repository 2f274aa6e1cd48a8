"""Developer tool: the HBM-bound kernels (STFT+compress, decompress+iSTFT, bridge update) timed alone (bench.py's leg)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rethinking-flow-and-diffusion-bridge-models-for-speech-enhancement_b200"))
import torch
import bench
from fdbm_b200 import EnhancementModel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
model = EnhancementModel("ncsnpp_v2", "sb", bridge_kwargs=dict(N=5, sampler_type="ode_ei")).to(dev).eval()
waves = bench.synth_batch(n, dev)
print(json.dumps(bench.hbm_kernel_leg(model, waves, dev), indent=1))
