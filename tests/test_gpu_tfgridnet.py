"""TF-GridNet on the GPU (csrc/tfgridnet.cu, fdbm_b200/tfgridnet.py) against the oracle's restatement of
fdbm/backbones/tfgridnet.py and the reference's own outputs (tests/golden/tfgridnet_T24.npz).

Whole-network bound: LSTM recurrences amplify operand rounding -- with these random weights the SAME fp32 algorithm run with
10-bit-mantissa operands (what cuDNN's TF32 does to the reference on a GPU) deviates 1-3e-2 from fp32
(tests/golden/ref_tf32_deviation.json), so as in test_gpu_parity.py the bound is max(5e-3, 3 x that deviation); the
kernel-level tests below pin each piece tightly, where nothing amplifies."""
import ctypes as C
import json
import os

import pytest
import torch
import torch.nn.functional as F

from helpers import load_npz, rel_l2

pytestmark = pytest.mark.gpu


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.fixture(scope="module")
def env():
    import fdbm_oracle as O
    from fdbm_b200 import _lib
    return O, _lib.load(), _lib


def _ck(_lib, rc):
    _lib.check(rc, "tfgridnet kernel")


def test_state_dict_and_registry():
    import fdbm_oracle as O
    from fdbm_b200 import BackboneRegistry
    for name, cfg in (("tfgridnet_5l32c100", O.TFGridNetConfig()), ("tfgridnet_5l32c100_predictive", O.TFGridNetConfig(predictive=True)),
                      ("tfgridnet_4l32c80", O.TFGridNetConfig(n_layers=4, lstm_hidden_units=80))):
        net = BackboneRegistry.get_by_name(name)(unused_kwarg=1)
        assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == O.tfgridnet_param_shapes(cfg), name
    with pytest.raises(NotImplementedError):
        BackboneRegistry.get_by_name("tfgridnet_5l32c100")(emb_ks=4, emb_hs=4)


@pytest.mark.parametrize("inter,B", [(False, 2), (True, 2), (False, 10)])
def test_lstm_sweep_and_post_vs_oracle(env, inter, B):
    """One BiLSTM path (LayerNorm -> unfold 4 -> BiLSTM(128 -> 100 x 2) -> ConvTranspose1d -> + residual, tfgridnet.py:335-375)
    on a small padded tensor: sequences along Q (intra) or along T (inter), including a partial 128-sequence tile (B = 10: 150
    sequences = one full tile + 22)."""
    O, lib, _lib = env
    cfg = O.TFGridNetConfig()
    sd = O.tfgridnet_state_dict(cfg, seed=3)
    g = torch.Generator().manual_seed(7)
    T, Q = 9, 14                                             # padded 15 x 20
    Tp, Qp = T + 6, Q + 6
    xp = torch.randn(B, Tp, Qp, 32, generator=g)
    name = "inter" if inter else "intra"
    p = "blocks.1."
    with torch.no_grad():
        ref = O._rnn_path(xp.transpose(1, 2), sd, p, name, cfg).transpose(1, 2) if inter else O._rnn_path(xp, sd, p, name, cfg)
    xpd = xp.cuda()
    xn = F.layer_norm(xpd, (32,), sd[p + name + "_norm.weight"].cuda(), sd[p + name + "_norm.bias"].cuda(), cfg.eps).half().contiguous()
    w = [sd[p + f"{name}_rnn.{n}_l0{sfx}"].cuda().contiguous() for sfx in ("", "_reverse") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    wl = sd[p + f"{name}_linear.weight"].cuda().contiguous()
    n_seq, L = (B * Qp, Tp - 3) if inter else (B * Tp, Qp - 3)
    yf = torch.zeros(n_seq * (L + 3) * 32, dtype=torch.float16, device="cuda"); yb = torch.zeros_like(yf)
    buf = torch.empty(int(lib.fdbm_tfg_lstm_pack_bytes()), dtype=torch.uint8, device="cuda")
    arr = (C.c_void_p * 8)(*[t.data_ptr() for t in w])
    _ck(_lib, lib.fdbm_tfg_lstm_pack(arr, wl.data_ptr(), 100, buf.data_ptr(), _stream()))
    xs = xn.transpose(1, 2).contiguous() if inter else xn              # sequences contiguous: [n_seq][L + 3][32]
    _ck(_lib, lib.fdbm_tfg_lstm_sweep(xs.data_ptr(), n_seq, L, buf.data_ptr(), yf.data_ptr(), yb.data_ptr(), _stream()))
    torch.cuda.synchronize()
    lb = sd[p + name + "_linear.bias"].cuda()
    if inter:
        out = torch.empty(B, T, Q, 32, device="cuda")
        _ck(_lib, lib.fdbm_tfg_sweep_post(yf.data_ptr(), yb.data_ptr(), lb.data_ptr(), xpd.data_ptr(), B, T, Q, 1, None, None, cfg.eps, None, None,
                                          out.data_ptr(), 0, _stream()))
        want = ref[:, 3:3 + T, 3:3 + Q]
    else:
        out = torch.empty(B, Tp, Qp, 32, device="cuda")
        xn2 = torch.empty(B, Tp, Qp, 32, dtype=torch.float16, device="cuda")
        xn2t = torch.empty(B, Qp, Tp, 32, dtype=torch.float16, device="cuda")
        gam, bet = sd[p + "inter_norm.weight"].cuda(), sd[p + "inter_norm.bias"].cuda()
        for buf2, tr in ((xn2, 0), (xn2t, 1)):
            _ck(_lib, lib.fdbm_tfg_sweep_post(yf.data_ptr(), yb.data_ptr(), lb.data_ptr(), xpd.data_ptr(), B, T, Q, 0, gam.data_ptr(), bet.data_ptr(),
                                              cfg.eps, out.data_ptr(), buf2.data_ptr(), None, tr, _stream()))
        want = ref
        ln = F.layer_norm(out, (32,), gam, bet, cfg.eps)
        assert rel_l2(xn2.float(), ln) < 1e-3
        assert torch.equal(xn2t, xn2.transpose(1, 2).contiguous())
    torch.cuda.synchronize()
    # the residual dominates `out`; compare the LSTM path's own contribution
    err = rel_l2(out.cpu() - (want - want + (xp[:, 3:3 + T, 3:3 + Q] if inter else xp)), want - (xp[:, 3:3 + T, 3:3 + Q] if inter else xp))
    print(f"{name} BiLSTM path ({n_seq} sequences x {L} steps): rel L2 of the path's contribution {err:.3e}")
    assert err < 5e-3


def test_pad_add_norm_and_input_and_output_kernels(env):
    O, lib, _lib = env
    cfg = O.TFGridNetConfig()
    sd = O.tfgridnet_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(9)
    B, T, Q = 2, 11, 257
    x = torch.view_as_complex(torch.randn(B, 1, Q, T, 2, generator=g)); y = torch.view_as_complex(torch.randn(B, 1, Q, T, 2, generator=g))
    inp = torch.cat((x.real, x.imag, y.real, y.imag), dim=1).permute(0, 1, 3, 2)
    ref = F.group_norm(F.conv2d(inp, sd["conv.0.weight"], sd["conv.0.bias"], padding=1), 1, sd["conv.1.weight"], sd["conv.1.bias"], cfg.eps)
    d = {k: v.cuda() for k, v in sd.items()}
    h = torch.empty(B, T, Q, 32, device="cuda"); sums = torch.empty(2 * B, dtype=torch.float64, device="cuda")
    xd, yd = x.cuda().contiguous(), y.cuda().contiguous()
    _ck(_lib, lib.fdbm_tfg_input(xd.data_ptr(), yd.data_ptr(), d["conv.0.weight"].data_ptr(), d["conv.0.bias"].data_ptr(), d["conv.1.weight"].data_ptr(),
                                 d["conv.1.bias"].data_ptr(), B, T, Q, 4, cfg.eps, sums.data_ptr(), h.data_ptr(), _stream()))
    assert rel_l2(h.permute(0, 3, 1, 2), ref) < 2e-6
    # time embedding
    t = torch.tensor([0.6, 0.0123])
    proj = torch.log(t)[:, None] * sd["get_time_emb.W"][None] * 2 * 3.141592653589793
    te = torch.cat([torch.sin(proj), torch.cos(proj)], -1)
    te = F.silu(F.linear(F.silu(F.linear(te, sd["time_emb_fc.0.weight"], sd["time_emb_fc.0.bias"])), sd["time_emb_fc.2.weight"], sd["time_emb_fc.2.bias"]))
    want_emb = torch.stack([F.linear(te, sd[f"time_emb_blocks.{b}.weight"], sd[f"time_emb_blocks.{b}.bias"]) for b in range(5)])
    wb = torch.stack([d[f"time_emb_blocks.{b}.weight"] for b in range(5)]).contiguous(); bb = torch.stack([d[f"time_emb_blocks.{b}.bias"] for b in range(5)]).contiguous()
    emb = torch.empty(5, B, 32, device="cuda"); td = t.cuda()
    _ck(_lib, lib.fdbm_tfg_time_embedding(td.data_ptr(), 1, d["get_time_emb.W"].data_ptr(), d["time_emb_fc.0.weight"].data_ptr(), d["time_emb_fc.0.bias"].data_ptr(),
                                          d["time_emb_fc.2.weight"].data_ptr(), d["time_emb_fc.2.bias"].data_ptr(), wb.data_ptr(), bb.data_ptr(), 5, B,
                                          emb.data_ptr(), _stream()))
    assert rel_l2(emb, want_emb) < 1e-4
    # pad + add + LayerNorm
    xp = torch.empty(B, T + 6, Q + 6, 32, device="cuda"); xn = torch.empty(B, T + 6, Q + 6, 32, dtype=torch.float16, device="cuda")
    _ck(_lib, lib.fdbm_tfg_pad_add_norm(h.data_ptr(), emb[2].data_ptr(), d["blocks.2.intra_norm.weight"].data_ptr(), d["blocks.2.intra_norm.bias"].data_ptr(),
                                        B, T, Q, cfg.eps, xp.data_ptr(), xn.data_ptr(), _stream()))
    want_xp = F.pad(h + emb[2][:, None, None, :], (0, 0, 3, 3, 3, 3))
    assert torch.equal(xp, want_xp)
    assert rel_l2(xn.float(), F.layer_norm(want_xp, (32,), d["blocks.2.intra_norm.weight"], d["blocks.2.intra_norm.bias"], cfg.eps)) < 1e-3
    # output deconv
    out = torch.empty(B, 1, Q, T, dtype=torch.complex64, device="cuda")
    _ck(_lib, lib.fdbm_tfg_output(h.data_ptr(), d["deconv.weight"].data_ptr(), d["deconv.bias"].data_ptr(), B, T, Q, out.data_ptr(), _stream()))
    r = F.conv_transpose2d(h.permute(0, 3, 1, 2).cpu(), sd["deconv.weight"], sd["deconv.bias"], padding=1)          # [B,2,T,Q]
    want = torch.view_as_complex(r.reshape(B, 1, 2, T, Q).permute(0, 1, 4, 3, 2).contiguous())
    assert rel_l2(out, want) < 2e-6


@pytest.mark.parametrize("T", [24, 61])
def test_attention_vs_oracle(env, T):
    O, lib, _lib = env
    cfg = O.TFGridNetConfig()
    sd = O.tfgridnet_state_dict(cfg, seed=1)
    g = torch.Generator().manual_seed(T)
    B, Q = 2, 257
    z = torch.randn(B, 32, T, Q, generator=g)
    p = "blocks.3."
    with torch.no_grad():
        ref = O._gridnet_attention(z, sd, p, cfg)
    names = [p + n for n in ("attn_conv_Q.weight", "attn_conv_Q.bias", "attn_conv_K.weight", "attn_conv_K.bias", "attn_conv_V.weight", "attn_conv_V.bias",
                              "attn_norm_Q.act.weight", "attn_norm_K.act.weight", "attn_norm_V.act.weight", "attn_norm_Q.gamma", "attn_norm_Q.beta",
                              "attn_norm_K.gamma", "attn_norm_K.beta", "attn_norm_V.gamma", "attn_norm_V.beta", "attn_concat_proj.0.weight",
                              "attn_concat_proj.0.bias", "attn_concat_proj.1.weight", "attn_concat_proj.2.gamma", "attn_concat_proj.2.beta")]
    tensors = [sd[n].cuda().contiguous() for n in names]
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    zd = z.permute(0, 2, 3, 1).contiguous().cuda()
    ws = torch.empty(int(lib.fdbm_tfg_attention_workspace_bytes(B, T, Q)) + 256, dtype=torch.uint8, device="cuda")
    base = (ws.data_ptr() + 255) // 256 * 256
    out = torch.empty_like(zd)
    _ck(_lib, lib.fdbm_tfg_attention(zd.data_ptr(), arr, B, T, Q, cfg.eps, base, out.data_ptr(), _stream()))
    err = rel_l2(out.permute(0, 3, 1, 2) - zd.permute(0, 3, 1, 2), ref - z)
    print(f"attention T={T}: rel L2 of the attention branch {err:.3e}")
    assert err < 5e-3


@pytest.mark.parametrize("name,key", [("tfgridnet_5l32c100", "D"), ("tfgridnet_5l32c100_predictive", "D_pred")])
def test_forward_vs_reference_golden(golden_dir, name, key):
    import fdbm_oracle as O
    from fdbm_b200 import BackboneRegistry
    pred = key == "D_pred"
    cfg = O.TFGridNetConfig(predictive=pred)
    sd = O.tfgridnet_state_dict(cfg, seed=0)
    net = BackboneRegistry.get_by_name(name)()
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    g = load_npz(f"{golden_dir}/tfgridnet_T24.npz")
    dev = json.load(open(os.path.join(golden_dir, "ref_tf32_deviation.json")))
    Y, X, t = (torch.from_numpy(g[k]).cuda() for k in ("Y", "X", "t"))
    D = net(Y) if pred else net(X, Y, t)
    err, yard = rel_l2(D, g[key]), dev[name + "_forward_T24"]
    bound = max(5e-3, 3 * yard)
    print(f"{name} T=24, B=2: rel L2 vs reference golden {err:.3e}  (bound {bound:.1e}; 10-bit-operand run of the same algorithm deviates {yard:.3e})")
    assert D.shape == g[key].shape and err < bound
    # batch rows are independent
    D1 = net(Y[1:]) if pred else net(X[1:], Y[1:], t[1:])
    assert rel_l2(D1[0], D[1]) < 1e-5


def test_enhance_with_tfgridnet_4s():
    """config.yaml's model end to end at BASELINE's size: a 4 s utterance through STFT -> 5-step SB sampler on tfgridnet_5l32c100 ->
    iSTFT; zero padding (infer_single.py:64-69 pads with zeros for every backbone but 'ncsnpp_v2')."""
    import fdbm_oracle as O
    from fdbm_b200 import EnhancementModel
    cfg = O.TFGridNetConfig()
    sd = O.tfgridnet_state_dict(cfg, seed=0)
    model = EnhancementModel("tfgridnet_5l32c100", "sb", bridge_kwargs=dict(N=2, sampler_type="ode_ei"))
    model.dnn.load_state_dict(sd)
    model = model.cuda().eval()
    assert model.pad_mode == "zero_pad"
    _, noisy = O.synth_pair(3, n_samples=64000)
    got = model.enhance_batch(torch.stack([noisy, noisy * 0.5]).cuda())
    assert got.shape == (2, 64000) and torch.isfinite(got).all()
    sc = O.SpecConfig()
    Yo = O.pad_spec(O.spec_fwd(O.stft(noisy[None] / noisy.abs().max(), sc), sc)[:, None], "zero_pad")
    with torch.no_grad():
        D_ref = O.tfgridnet_forward(sd, cfg, Yo, Yo, torch.tensor([1.0]))
    D = model.dnn(Yo.cuda(), Yo.cuda(), torch.ones(1, device="cuda"))
    err = rel_l2(D, D_ref)
    print(f"tfgridnet_5l32c100 forward, 4 s (T=256): rel L2 vs oracle {err:.3e}")
    assert err < 0.1
