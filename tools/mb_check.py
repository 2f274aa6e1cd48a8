"""Developer tool: the same utterances enhanced at two micro-batch sizes must give the same waveforms (per-utterance independence)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rethinking-flow-and-diffusion-bridge-models-for-speech-enhancement_b200"))
sys.path.insert(0, ROOT)
import torch
from fdbm_b200 import EnhancementModel, sensitise_
from bench import synth_batch
a, b = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
model = EnhancementModel("ncsnpp_v2", "sb", bridge_kwargs=dict(N=5, sampler_type="ode_ei"))
sensitise_(model.dnn, seed=0)
model = model.to(dev).eval()
waves = synth_batch(max(a, b), dev, seed=1234)
oa = model.enhance_many(waves, micro_batch=a).clone()
ob = model.enhance_many(waves, micro_batch=b).clone()
d = (oa - ob).abs().max().item()
rel = ((oa - ob).norm() / oa.norm()).item()
print(f"micro-batch {a} vs {b}: max |diff| {d:.3e}, rel L2 {rel:.3e}, finite {bool(torch.isfinite(ob).all())}, per-utterance max rel "
      f"{((oa - ob).norm(dim=1) / oa.norm(dim=1)).max().item():.3e}")
