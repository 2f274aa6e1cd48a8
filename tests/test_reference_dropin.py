"""The Python seams of the drop-in boundary, checked against the REFERENCE's own objects (CPU, no CUDA call).

Runs only where /root/reference is present (the build container); the GPU box has no copy of the reference.
SURVEY.md section 8(b): (1) BackboneRegistry.register / get_by_name + add_argparse_args + **kwargs tolerance,
(2) Bridge attributes and path maths, (3) SpecsDataModule attributes, state_dict compatibility in both directions."""
import argparse
import os
import sys

import pytest
import torch

REF = os.environ.get("FDBM_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "fdbm")), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import make_golden as MG
    Bridge, BackboneRegistry, SpecsDataModule, pad_spec, si_sdr = MG.import_reference()
    import fdbm.bridge as ref_bridge
    return dict(Bridge=Bridge, BackboneRegistry=BackboneRegistry, SpecsDataModule=SpecsDataModule, pad_spec=pad_spec,
                si_sdr=si_sdr, bridge_mod=ref_bridge)


def test_our_backbones_register_into_the_reference_registry(ref):
    """`BackboneRegistry.register(name)(cls)` + `get_by_name(name)(**kwargs)` (util/registry.py:17-30, model.py:47-48) with the
    kwargs soup train.py passes (every argparse group's values at once)."""
    import fdbm_b200
    R = ref["BackboneRegistry"]
    ref_cls = R.get_by_name("ncsnpp_v2")
    R.register("ncsnpp_v2_b200")(fdbm_b200.NCSNpp_v2)
    R.register("ncsnpp_v2_predictive_b200")(fdbm_b200.NCSNpp_v2_predictive)
    cls = R.get_by_name("ncsnpp_v2_b200")
    assert cls is fdbm_b200.NCSNpp_v2
    # argparse seam (train.py:92-93): same flags, same defaults
    pa, pb = argparse.ArgumentParser(), argparse.ArgumentParser()
    ref_cls.add_argparse_args(pa); cls.add_argparse_args(pb)
    assert vars(pa.parse_args([])) == vars(pb.parse_args([]))
    # constructor tolerates the reference's full kwargs dictionary (model.py:47-48 passes **kwargs of ALL groups)
    soup = dict(vars(pa.parse_args([])), lr=1e-4, ema_decay=0.999, N=5, T=1.0, sampler_type="ode_ei", noise_schedule="bb",
                base_dir="/x", batch_size=8, n_fft=512, hop_length=256, gpus=1, loss_type="data_prediction_hybrid")
    ours = cls(**soup)
    theirs = ref_cls(**soup)
    # state_dict compatibility in both directions, strict
    sd_ref = theirs.state_dict()
    assert list(sd_ref.keys()).sort() == list(ours.state_dict().keys()).sort()
    assert {k: tuple(v.shape) for k, v in sd_ref.items()} == {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    ours.load_state_dict(sd_ref, strict=True)
    theirs.load_state_dict(ours.state_dict(), strict=True)
    # requires_grad flags agree (the Fourier projection is frozen, layerspp.py:37): torch_ema / Adam see the same parameter list
    assert [(n, p.requires_grad) for n, p in theirs.named_parameters()] == [(n, p.requires_grad) for n, p in ours.named_parameters()]
    # same initialisation law: per-tensor standard deviations of a fresh model agree statistically
    torch.manual_seed(0); a = ref_cls()
    torch.manual_seed(0); b = cls()
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        if p.numel() > 4096:
            assert abs(float(p.std()) - float(q.std())) < 0.08 * float(p.std()) + 1e-9, n
    # unsupported architecture variants are refused loudly, not silently ignored
    for bad in (dict(dropout=0.1), dict(fir_kernel=[1, 2, 1]), dict(resamp_with_conv=False), dict(progressive="none")):
        with pytest.raises(NotImplementedError):
            cls(**bad)
    pred = R.get_by_name("ncsnpp_v2_predictive_b200")()
    pred.load_state_dict(R.get_by_name("ncsnpp_v2_predictive")().state_dict(), strict=True)


def test_forward_without_cuda_fails_loudly(ref):
    import fdbm_b200
    net = fdbm_b200.NCSNpp_v2()
    x = torch.zeros(1, 1, 257, 64, dtype=torch.complex64)
    with pytest.raises(RuntimeError):
        net(x, x, torch.ones(1))                     # CPU tensors: no fallback


@pytest.mark.parametrize("path,kw", [("sb", dict(noise_schedule="bb")), ("sb", dict(noise_schedule="ve")),
                                     ("sb", dict(noise_schedule="vp", c=0.3)), ("sb", dict(noise_schedule="gmax")), ("fm", dict())])
def test_bridge_host_maths_equal_the_reference(ref, path, kw):
    """Attributes callers read / mutate (infer_single.py:55-56, model.py:451-452) and every scalar function of the paths."""
    import fdbm_b200
    rb = ref["Bridge"](path, N=7, sampler_type="sde_ei", sampling_eps=1e-3, **kw)
    ob = fdbm_b200.Bridge(path, N=7, sampler_type="sde_ei", sampling_eps=1e-3, **kw)
    for attr in ("N", "T", "sampler_type", "start_time", "end_time"):
        assert getattr(rb, attr) == getattr(ob, attr), attr
    assert rb.path.sampling_direction == ob.path.sampling_direction and rb.path.T == ob.path.T
    ob.N, ob.sampler_type = 3, "ode_ei"                                           # mutable after construction
    assert ob.time_grid().numel() == 4
    t = torch.tensor([0.03, 0.2, 0.55, 0.9999, 1.0])
    for a, b in zip(rb.path.path_param(t), ob.path.path_param(t)):
        assert torch.equal(a, b)
    assert torch.equal(rb._std(t), ob._std(t))
    g = torch.Generator().manual_seed(1)
    s, y, x = (torch.view_as_complex(torch.randn(5, 1, 9, 4, 2, generator=g)) for _ in range(3))
    mr, sr = rb.probability_path(s, y, t); mo, so = ob.probability_path(s, y, t)
    assert torch.equal(mr, mo) and torch.equal(sr, so)
    assert torch.equal(rb.score_fn(t, x, s, y), ob.score_fn(t, x, s, y))
    # ode / sde: the reference's [B]-into-[B,1,F,T] product is only meaningful for B = 1 (its callers); equal there
    t1 = torch.tensor([0.4])
    x1, s1, y1 = x[:1], s[:1], y[:1]
    assert torch.allclose(rb.path.ode(t1, x1, s1, y1), ob.path.ode(t1, x1, s1, y1), rtol=1e-6, atol=1e-7)
    if path == "sb":
        dr, gr = rb.path.sde(t1, x1, s1, y1); do, go = ob.path.sde(t1, x1, s1, y1)
        assert torch.allclose(dr, do, rtol=1e-6, atol=1e-7) and torch.allclose(torch.as_tensor(gr), torch.as_tensor(go))
        fr, g_r = rb.path.auxiliary_param(t1); fo, g_o = ob.path.auxiliary_param(t1)
        assert torch.allclose(torch.as_tensor(fr, dtype=torch.float32), torch.as_tensor(fo, dtype=torch.float32))
        assert torch.allclose(torch.as_tensor(g_r), torch.as_tensor(g_o))
        for st in ("ode_ei", "sde_ei"):
            fn = "sampling_param_" + st
            for a, b in zip(getattr(rb.path, fn)(t1 * 0.5, t1, 1, "cpu"), getattr(ob.path, fn)(t1 * 0.5, t1, 1, "cpu")):
                assert torch.equal(a, b)
    # argparse seam of Bridge and of the path classes
    for r_cls, o_cls in ((ref["Bridge"], fdbm_b200.Bridge), (type(rb.path), type(ob.path))):
        pa, pb = argparse.ArgumentParser(), argparse.ArgumentParser()
        r_cls.add_argparse_args(pa); o_cls.add_argparse_args(pb)
        assert vars(pa.parse_args([])) == vars(pb.parse_args([]))


def test_oracle_pc_and_ode_int_samplers_equal_the_reference(ref):
    """The oracle's restatement of pc_sampler / ode_sampler_int (the GPU tests' yardstick) against the reference's own
    Bridge, with a cheap stand-in model (B = 1, as the reference's broadcasting requires)."""
    import fdbm_oracle as O
    g = torch.Generator().manual_seed(3)
    y = torch.view_as_complex(torch.randn(1, 1, 17, 8, 2, generator=g)) * 0.3
    model = lambda x, yy, t: 0.6 * x + 0.3 * yy * torch.cos(t)[:, None, None, None] + 0.05 * x.abs()
    zs = [torch.view_as_complex(torch.randn(1, 1, 17, 8, 2, generator=g)) * 0.5 ** 0.5 for _ in range(40)]
    for corrector, steps in (("ald", 1), ("langevin", 2), ("none", 1)):
        for predictor in ("euler_maruyama", "none"):
            rb = ref["Bridge"]("sb", N=4, sampler_type="pc")
            ob = O.Bridge("sb", N=4, sampler_type="pc")
            seq = iter(zs)
            orig = torch.randn_like
            torch.randn_like = lambda x, **k: next(seq)
            try:
                want = rb.sampler(model, y, predictor_name=predictor, corrector_name=corrector, snr=0.3, corrector_steps=steps)
            finally:
                torch.randn_like = orig
            got = ob.pc_sampler(model, y, predictor_name=predictor, corrector_name=corrector, snr=0.3, corrector_steps=steps,
                                z0=zs[0], zs=zs[1:])
            assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), (corrector, predictor)
    with pytest.raises(ValueError):
        ref["Bridge"]("sb", N=4, sampler_type="pc").sampler(model, y)             # 'reverse_diffusion' is not registered
    orig = torch.randn_like
    torch.randn_like = lambda x, **k: zs[0]
    try:
        want = ref["Bridge"]("fm", sampler_type="ode_int").sampler(model, y, rtol=1e-6, atol=1e-6)
    finally:
        torch.randn_like = orig
    got = O.Bridge("fm", sampler_type="ode_int").ode_sampler_int(model, y, rtol=1e-6, atol=1e-6, z0=zs[0])
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)


def test_data_module_attributes_and_pad_spec(ref):
    import fdbm_b200
    rd = ref["SpecsDataModule"](base_dir="/unused", n_fft=512, hop_length=256, num_frames=256, window="sqrthann", gpu=False)
    od = fdbm_b200.SpecsDataModule(base_dir="/unused", n_fft=512, hop_length=256, num_frames=256, window="sqrthann", gpu=False)
    for attr in ("normalize", "n_fft", "hop_length", "num_frames", "spec_factor", "spec_abs_exponent", "transform_type"):
        assert getattr(rd, attr) == getattr(od, attr), attr
    assert torch.equal(rd.window, od.window)
    assert set(rd.stft_kwargs) == set(od.stft_kwargs)
    with pytest.raises(NotImplementedError):
        fdbm_b200.pad_spec(torch.zeros(1, 1, 257, 10, dtype=torch.complex64), mode="bogus")
