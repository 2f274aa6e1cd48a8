"""The CPU oracle against the golden vectors produced by the reference itself
(oracle/make_golden.py, run in the build container where /root/reference exists)."""
import pytest
import numpy as np
import torch

import fdbm_oracle as O
from helpers import load_npz, rel_l2


def test_spectral_chain(golden_dir):
    g = load_npz(f"{golden_dir}/spectral_1s.npz")
    cfg = O.SpecConfig()
    y = torch.from_numpy(g["wave"])
    S = O.stft(y, cfg)
    assert torch.equal(S, torch.from_numpy(g["stft"]))                 # bit-identical to torch.stft on CPU
    Y = O.spec_fwd(S, cfg)
    assert rel_l2(Y, g["spec"]) < 1e-7
    assert torch.equal(O.pad_spec(torch.from_numpy(g["spec"])[None], "reflection"), torch.from_numpy(g["spec_reflect"]))
    assert torch.equal(O.pad_spec(torch.from_numpy(g["spec"])[None], "zero_pad"), torch.from_numpy(g["spec_zero"]))
    w = O.istft(O.spec_back(torch.from_numpy(g["spec_reflect"]).squeeze(), cfg), cfg, 16000)
    assert rel_l2(w, g["wave_back"].reshape(-1)) < 1e-6


def test_frame_index_matches_unfold():
    n = 5000
    x = torch.arange(n, dtype=torch.float32)
    idx = O.frame_index(n, 512, 256)
    ref = torch.nn.functional.pad(x[None, None], (256, 256), mode="reflect")[0, 0].unfold(-1, 512, 256)
    assert torch.equal(x[torch.from_numpy(idx)], ref)


def test_coefficient_tables_bit_exact(golden_dir):
    g = load_npz(f"{golden_dir}/coeff_tables.npz")
    n = 0
    for key, ref in g.items():
        if key.endswith("_pathparam"):
            continue
        path, sched, st1, st2, N = key.split("_")
        kw = {} if path == "fm" else {"noise_schedule": sched}
        b = O.Bridge(path, N=int(N[1:]), sampler_type=f"{st1}_{st2}", **kw)
        assert np.array_equal(b.coefficient_table().numpy(), ref), key
        n += 1
    assert n == 36
    # the values SURVEY.md section 8 (A6) quotes for the default sampler
    t = O.Bridge("sb", N=5, sampler_type="ode_ei").coefficient_table()
    assert abs(float(t[0, 0]) - 3999.45) < 0.01 and abs(float(t[4, 1]) - 0.979906) < 1e-5


def test_fir_identities(golden_dir):
    g = load_npz(f"{golden_dir}/fir.npz")
    x = torch.from_numpy(g["x"])
    assert float((O.fir_down2(x) - torch.from_numpy(g["down"])).abs().max()) < 5e-7
    assert float((O.fir_up2(x) - torch.from_numpy(g["up"])).abs().max()) < 5e-7


def test_backbone_forward_golden(golden_dir):
    torch.set_num_threads(max(1, torch.get_num_threads()))
    g = load_npz(f"{golden_dir}/bridge_T64.npz")
    cfg = O.NcsnppConfig()
    sd = O.sensitised_state_dict(cfg, seed=0)
    assert len(sd) == 647 and sum(v.numel() for v in sd.values()) == 65590822
    with torch.no_grad():
        D = O.ncsnpp_forward(sd, cfg, torch.from_numpy(g["xt"]), torch.from_numpy(g["Y"]), torch.from_numpy(g["t"]))
    assert rel_l2(D, g["D"]) < 2e-5
    assert float(torch.from_numpy(g["D"]).abs().std()) > 0.5           # sensitised weights: O(1) output, not a constant


def test_predictive_forward_golden(golden_dir):
    g = load_npz(f"{golden_dir}/predictive_T64.npz")
    cfg = O.NcsnppConfig(predictive=True)
    sd = O.sensitised_state_dict(cfg, seed=0)
    with torch.no_grad():
        D = O.ncsnpp_forward(sd, cfg, torch.from_numpy(g["Y"]))
    assert rel_l2(D, g["D"]) < 2e-5


def test_sampler_golden_sb_ode(golden_dir):
    g = load_npz(f"{golden_dir}/bridge_T64.npz")
    cfg = O.NcsnppConfig()
    sd = O.sensitised_state_dict(cfg, seed=0)
    b = O.Bridge("sb", N=5, sampler_type="ode_ei")
    s = b.sampler(lambda a, c, t: O.ncsnpp_forward(sd, cfg, a, c, t), torch.from_numpy(g["Y"]))
    assert rel_l2(s, g["sample_sb_ode_ei"]) < 2e-4
    w = O.istft(O.spec_back(s.squeeze(), O.SpecConfig()), O.SpecConfig(), 16000)
    assert O.si_sdr(g["wave_sb_ode_ei"].reshape(-1), w.numpy()) > 60.0


def test_si_sdr_and_synth():
    c, n = O.synth_pair(0, 16000)
    assert c.shape == (16000,) and n.shape == (16000,)
    assert -1.0 < O.si_sdr(c.numpy(), n.numpy()) < 16.0                # SNR drawn from U(0, 15) dB
    c2, _ = O.synth_pair(0, 16000)
    assert torch.equal(c, c2)


@pytest.mark.parametrize("pred", [False, True])
def test_tfgridnet_oracle_matches_reference_golden(golden_dir, pred):
    """The oracle's restatement of tfgridnet.py / tfgridnet_predictive.py against the reference's own output (make_golden.py)."""
    g = load_npz(f"{golden_dir}/tfgridnet_T24.npz")
    cfg = O.TFGridNetConfig(predictive=pred)
    sd = O.tfgridnet_state_dict(cfg, seed=0)
    Y, X, t = (torch.from_numpy(g[k]) for k in ("Y", "X", "t"))
    with torch.no_grad():
        got = O.tfgridnet_forward(sd, cfg, Y) if pred else O.tfgridnet_forward(sd, cfg, X, Y, t)
    assert rel_l2(got, g["D_pred" if pred else "D"]) < 1e-4


def test_hybrid_loss_oracle_matches_reference_golden(golden_dir):
    """O.hybrid_loss against BridgeModel._loss of the reference (value and autograd gradient, tests/golden/hybrid_loss.npz)."""
    g = load_npz(f"{golden_dir}/hybrid_loss.npz")
    x, xh = torch.from_numpy(g["x"]), torch.from_numpy(g["x_hat"]).requires_grad_(True)
    loss = O.hybrid_loss(xh, x, O.SpecConfig())
    (grad,) = torch.autograd.grad(loss, xh)
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    assert rel_l2(torch.nan_to_num(grad), g["grad"]) < 1e-5


def test_data_prediction_loss_oracle_matches_reference_golden(golden_dir):
    """O.data_prediction_loss against BridgeModel._loss 'data_prediction' of the reference (value and autograd gradient,
    tests/golden/data_prediction_loss.npz)."""
    g = load_npz(f"{golden_dir}/data_prediction_loss.npz")
    x, xh = torch.from_numpy(g["x"]), torch.from_numpy(g["x_hat"]).requires_grad_(True)
    loss = O.data_prediction_loss(xh, x, O.SpecConfig(), 0.001)
    (grad,) = torch.autograd.grad(loss, xh)
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    assert rel_l2(torch.nan_to_num(grad), g["grad"]) < 1e-5


def test_mel_loss_oracle_matches_reference_golden(golden_dir):
    """O.data_prediction_mel_loss (mel and mel + phase) against BridgeModel._loss 'data_prediction_mel' / '_melphase' of the
    reference run through its own MelSpectrogramLoss / PhaseLoss (tests/golden/mel_loss.npz; librosa's filterbank is the oracle's
    restatement there, so the fixture pins everything except the filterbank formula)."""
    g = load_npz(f"{golden_dir}/mel_loss.npz")
    for tag, with_phase in (("mel", False), ("melphase", True)):
        x, xh = torch.from_numpy(g["x"]), torch.from_numpy(g["x_hat"]).requires_grad_(True)
        loss = O.data_prediction_mel_loss(xh, x, O.SpecConfig(), with_phase)
        (grad,) = torch.autograd.grad(loss, xh)
        assert abs(float(loss) - float(g["loss_" + tag])) < 1e-5 * abs(float(g["loss_" + tag]))
        assert rel_l2(torch.nan_to_num(grad), g["grad_" + tag]) < 1e-5


def test_mel_filterbank_properties():
    """The restated librosa.filters.mel (Slaney scale, area-normalised triangles): shape, non-negativity, every filter non-empty at
    the seven resolutions of model.py:77-92, unit area in Hz, and the known centre frequencies of the Slaney scale (linear 66.67 Hz
    per mel below 1 kHz)."""
    for n_mels, n_fft in zip(O.MEL_N_MELS, O.MEL_N_FFTS):
        w = O.mel_filterbank(16000, n_fft, n_mels)
        assert w.shape == (n_mels, n_fft // 2 + 1) and bool((w >= 0).all())
        if n_fft >= 4 * n_mels:
            assert bool((w.sum(1) > 0).all())
    w = O.mel_filterbank(16000, 2048, 80).double()
    area = w.sum(1) * (8000.0 / 1024)                       # integral over frequency of an area-normalised triangle = 1
    assert float((area - 1).abs().max()) < 0.05
    peaks = w.argmax(1).double() * (8000.0 / 1024)
    mel_step = (15 + 27 * __import__("math").log(8.0) / __import__("math").log(6.4)) / 81      # hz_to_mel(8000) / (n_mels + 1)
    assert abs(float(peaks[0]) - mel_step * 200.0 / 3) < 8.0 and abs(float(peaks[9]) - 10 * mel_step * 200.0 / 3) < 8.0
