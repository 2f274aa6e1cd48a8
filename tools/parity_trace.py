"""Where does a sampler's deviation from the fp32 oracle come from?  (GPU box; test infrastructure, not product.)

For each sampler the oracle runs with a trace of (x_t before the step, D = model(x_t, y, t)); the CUDA backbone is then
evaluated TEACHER-FORCED on the oracle's own x_t of every step, so the per-step one-pass error is separated from the
loop's amplification.  Also prints the oracle's own TF32-emulated (operand_rounding) deviation per step."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "rethinking-flow-and-diffusion-bridge-models-for-speech-enhancement_b200"), os.path.join(ROOT, "oracle"),
          os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import fdbm_oracle as O                        # noqa: E402
from helpers import load_npz, rel_l2           # noqa: E402
from fdbm_b200 import BackboneRegistry, Bridge  # noqa: E402


def main():
    cfg = O.NcsnppConfig()
    sd = O.sensitised_state_dict(cfg, seed=0)
    net = BackboneRegistry.get_by_name("ncsnpp_v2")()
    net.load_state_dict(sd)
    net = net.cuda().eval()
    g = load_npz(os.path.join(ROOT, "tests", "golden", "bridge_T64.npz"))
    Y = torch.from_numpy(g["Y"])
    om = lambda a, b, c: O.ncsnpp_forward(sd, cfg, a, b, c)
    for path, st in (("sb", "ode_ei"), ("sb", "sde_ei"), ("fm", "ode_ei")):
        key = f"noise_{path}_{st}" if f"noise_{path}_{st}" in g else "noise_sb_sde_ei"
        zs = [torch.from_numpy(z) for z in g[key]]
        ob = O.Bridge(path, N=5, sampler_type=st)
        trace = []
        with torch.no_grad():
            xt = ob.prior_sampling(Y, zs[0])
            ref = ob.sampler(om, Y, z0=zs[0], zs=zs[1:], trace=trace)
        ts = ob.time_grid()
        x_prev = xt
        print(f"== {path}/{st}")
        for i, (est, x_after) in enumerate(trace):
            t = ts[i] * torch.ones(1)
            D = net(x_prev.cuda(), Y.cuda(), t.cuda()).cpu()
            with torch.no_grad(), O.operand_rounding("tf32"):
                D_tf = om(x_prev, Y, t)
            w = ob.coefficient_table()[i].tolist()
            print(f"  step {i}: t={float(ts[i]):.4f}  |x| {float(x_prev.abs().pow(2).mean().sqrt()):.3f}  |D| {float(est.abs().pow(2).mean().sqrt()):.3f}  "
                  f"one-pass err ours {rel_l2(D, est):.3e}  tf32-emulated {rel_l2(D_tf, est):.3e}   w = ({w[0]:.4g}, {w[1]:.4g}, {w[2]:.4g})")
            x_prev = x_after
        br = Bridge(path, N=5, sampler_type=st, match_torch_rng=True)
        seq = iter([z.cuda() for z in zs])
        orig = torch.randn_like
        torch.randn_like = lambda x, **k: next(seq)
        try:
            got = br.sampler(net, Y.cuda())
        finally:
            torch.randn_like = orig
        print(f"  whole loop: ours vs oracle {rel_l2(got, ref):.3e}")


if __name__ == "__main__":
    main()
