"""Generates csrc/tw512.inc: the 512 twiddle factors exp(-2 pi i m / 512) of the STFT / iSTFT kernels, rounded from double."""
import math
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "rethinking-flow-and-diffusion-bridge-models-for-speech-enhancement_b200", "csrc", "tw512.inc")
lines = ["{%.9ef, %.9ef}" % (math.cos(2 * math.pi * m / 512), -math.sin(2 * math.pi * m / 512)) for m in range(512)]
body = ",\n".join("  " + ", ".join(lines[i:i + 4]) for i in range(0, 512, 4))
open(out, "w").write("// exp(-2 pi i m / 512), m = 0..511, rounded from double precision (generated: see tools/gen_tw512.py)\n" + body + "\n")
