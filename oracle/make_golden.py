"""Generate tests/golden/*.npz by running the REFERENCE itself (imported from /root/reference
with stubs for its missing optional dependencies) on seeded inputs and
the oracle's deterministic sensitised weights.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py            # writes tests/golden/, prints oracle-vs-reference errors

The GPU box never runs this; it only reads the committed fixtures.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("FDBM_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)


def import_reference():
    sys.path.insert(0, REF)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    stub("pesq", pesq=lambda *a, **k: 0.0)
    stub("pystoi", stoi=lambda *a, **k: 0.0)

    class _LM(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

    stub("pytorch_lightning", LightningModule=_LM,
         LightningDataModule=type("LDM", (), {"__init__": lambda s: None}))
    from fdbm.bridge import Bridge
    from fdbm.backbones import BackboneRegistry
    from fdbm.data_module import SpecsDataModule
    from fdbm.util.other import pad_spec, si_sdr
    return Bridge, BackboneRegistry, SpecsDataModule, pad_spec, si_sdr


def rel(a, b):
    a = torch.as_tensor(a); b = torch.as_tensor(b)
    return float((a - b).abs().pow(2).sum().sqrt() / b.abs().pow(2).sum().sqrt().clamp_min(1e-30))


def c2n(t):
    return t.detach().cpu().numpy()


def main():
    import fdbm_oracle as O
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    Bridge, BackboneRegistry, SpecsDataModule, ref_pad_spec, ref_si_sdr = import_reference()
    os.makedirs(OUT, exist_ok=True)
    dm = SpecsDataModule(base_dir="/unused", n_fft=512, hop_length=256, num_frames=256,
                         window="sqrthann", gpu=False)
    scfg = O.SpecConfig()

    # ---- (1) spectral front/back end on synthetic utterances -------------------------------
    clean, noisy = O.synth_pair(0, n_samples=16000)           # 1 s -> 63 frames -> padded 64
    y = (noisy / noisy.abs().max())[None]
    S_ref = dm.stft(y)
    Y_ref = dm.spec_fwd(S_ref)
    Yp_ref = ref_pad_spec(Y_ref[None], mode="reflection")
    Yz_ref = ref_pad_spec(Y_ref[None], mode="zero_pad")
    wav_ref = dm.istft(dm.spec_back(Yp_ref.squeeze()), 16000)
    S_or = O.stft(y, scfg)
    Y_or = O.spec_fwd(S_or, scfg)
    Yp_or = O.pad_spec(Y_or[None], "reflection")
    wav_or = O.istft(O.spec_back(Yp_or.squeeze(), scfg), scfg, 16000)
    print("stft      oracle vs ref: max abs", float((S_or - S_ref).abs().max()), "bit-equal", bool(torch.equal(S_or, S_ref)))
    print("spec_fwd  oracle vs ref:", rel(Y_or, Y_ref))
    print("pad_spec  oracle vs ref: equal", bool(torch.equal(Yp_or, Yp_ref)),
          bool(torch.equal(O.pad_spec(Y_or[None], 'zero_pad'), O.pad_spec(Y_ref[None], 'zero_pad'))))
    print("istft     oracle vs ref:", rel(wav_or, wav_ref), float((wav_or - wav_ref).abs().max()))
    np.savez_compressed(os.path.join(OUT, "spectral_1s.npz"),
                        wave=c2n(y), stft=c2n(S_ref), spec=c2n(Y_ref), spec_reflect=c2n(Yp_ref),
                        spec_zero=c2n(Yz_ref), wave_back=c2n(wav_ref))

    # ---- (2) coefficient tables --------------------------------------------------------------
    tables = {}
    for path, kw in (("sb", dict(noise_schedule="bb")), ("sb", dict(noise_schedule="ve")),
                     ("sb", dict(noise_schedule="vp")), ("sb", dict(noise_schedule="gmax")),
                     ("fm", dict())):
        for st in ("ode_ei", "sde_ei"):
            if path == "fm" and st == "sde_ei":
                continue
            for N in (1, 5, 10, 30):
                rb = Bridge(path, N=N, sampler_type=st, **kw)
                ob = O.Bridge(path, N=N, sampler_type=st, **kw)
                ts = torch.linspace(rb.start_time, rb.end_time, N + 1)
                rows = []
                tp = ts[0] * torch.ones(1)
                for t in ts[1:]:
                    tt = t * torch.ones(1)
                    if st == "ode_ei":
                        w = rb.path.sampling_param_ode_ei(tt, tp, 1, "cpu")
                    else:
                        w = list(rb.path.sampling_param_sde_ei(tt, tp, 1, "cpu"))
                        if t == ts[-1]:
                            w[2] = torch.zeros_like(w[2])
                    rows.append(torch.stack([w[0][0], w[1][0], w[2][0]]))
                    tp = tt
                tab = torch.stack(rows)
                key = f"{path}_{kw.get('noise_schedule', 'ot')}_{st}_N{N}"
                tables[key] = c2n(tab)
                same = torch.equal(tab, ob.coefficient_table())
                if not same:
                    print("coefficient table MISMATCH", key, (tab - ob.coefficient_table()).abs().max())
                # prior + path params
                tq = torch.tensor([0.03, 0.25, 0.5, 0.9999, 1.0])
                pr = rb.path.path_param(tq); po = ob.path.path_param(tq)
                assert all(torch.equal(a, b) for a, b in zip(pr, po)), key
                tables[key + "_pathparam"] = c2n(torch.stack(pr))
    print("coefficient tables:", len(tables), "entries, oracle bit-identical")
    np.savez_compressed(os.path.join(OUT, "coeff_tables.npz"), **tables)

    # ---- (3) FIR resampling identities vs the reference's upfirdn2d_native ---------------------
    from fdbm.backbones.ncsnpp_utils import up_or_down_sampling as uds
    xx = torch.randn(2, 3, 8, 12)
    print("fir_down2 vs ref:", float((O.fir_down2(xx) - uds.downsample_2d(xx, (1, 3, 3, 1), factor=2)).abs().max()))
    print("fir_up2   vs ref:", float((O.fir_up2(xx) - uds.upsample_2d(xx, (1, 3, 3, 1), factor=2)).abs().max()))
    np.savez_compressed(os.path.join(OUT, "fir.npz"), x=c2n(xx),
                        down=c2n(uds.downsample_2d(xx, (1, 3, 3, 1), factor=2)),
                        up=c2n(uds.upsample_2d(xx, (1, 3, 3, 1), factor=2)))

    # ---- (4) backbone + sampler at T=64 frames with sensitised weights -------------------------
    for name, pred in (("ncsnpp_v2", False), ("ncsnpp_v2_predictive", True)):
        cfg = O.NcsnppConfig(predictive=pred)
        sd = O.sensitised_state_dict(cfg, seed=0)
        net = BackboneRegistry.get_by_name(name)().eval()
        ref_sd = net.state_dict()
        assert set(ref_sd.keys()) == set(sd.keys()), (set(ref_sd) ^ set(sd))
        for k in ref_sd:
            assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
        net.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(7)
        Y = Yp_ref.clone()                                                     # [1,1,257,64]
        with torch.no_grad():
            if pred:
                D_ref = net(Y)
                D_or = O.ncsnpp_forward(sd, cfg, Y)
                print(name, "forward oracle vs ref:", rel(D_or, D_ref), "out std", float(D_ref.abs().std()))
                np.savez_compressed(os.path.join(OUT, "predictive_T64.npz"), Y=c2n(Y), D=c2n(D_ref))
                continue
            xt = Y + 0.3 * torch.view_as_complex(torch.randn(1, 1, 257, 64, 2, generator=g))
            t = torch.tensor([0.6])
            D_ref = net(xt, Y, t)
            D_or = O.ncsnpp_forward(sd, cfg, xt, Y, t)
            print(name, "forward oracle vs ref:", rel(D_or, D_ref), "out std", float(D_ref.abs().std()))
            out = dict(Y=c2n(Y), xt=c2n(xt), t=c2n(t), D=c2n(D_ref))
            for path, st in (("sb", "ode_ei"), ("sb", "sde_ei"), ("fm", "ode_ei")):
                rb = Bridge(path, N=5, sampler_type=st)
                ob = O.Bridge(path, N=5, sampler_type=st)
                # deterministic noise: patch torch.randn_like through a shared generator sequence
                zg = torch.Generator().manual_seed(11)
                zs = [torch.view_as_complex(torch.randn(1, 1, 257, 64, 2, generator=zg)) * (0.5 ** 0.5)
                      for _ in range(6)]
                seq = iter(zs)
                orig = torch.randn_like
                torch.randn_like = lambda x, **k: next(seq)
                try:
                    s_ref = rb.sampler(lambda a, b, c: net(a, b, c), Y)
                finally:
                    torch.randn_like = orig
                s_or = ob.sampler(lambda a, b, c: O.ncsnpp_forward(sd, cfg, a, b, c), Y, z0=zs[0], zs=zs[1:])
                w_ref = dm.istft(dm.spec_back(s_ref.squeeze()), 16000)
                w_or = O.istft(O.spec_back(s_or.squeeze(), scfg), scfg, 16000)
                print(f"{name} sampler {path}/{st}: spec rel {rel(s_or, s_ref):.3e}  wave rel {rel(w_or, w_ref):.3e}")
                out[f"sample_{path}_{st}"] = c2n(s_ref)
                out[f"wave_{path}_{st}"] = c2n(w_ref)
                if st != "ode_ei" or path == "fm":
                    out[f"noise_{path}_{st}"] = c2n(torch.stack(zs))
            np.savez_compressed(os.path.join(OUT, "bridge_T64.npz"), **out)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  ", f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
