"""Host-side mirror of the reference interfaces (no GPU): schedule maths, registries, module table,
state_dict names, sharding, and the world_size-2 gather over gloo."""
import os
import sys

import numpy as np
import pytest
import torch

import fdbm_oracle as O
from helpers import load_npz


def test_coefficient_tables_match_reference_bit_exact(golden_dir):
    from fdbm_b200 import Bridge
    g = load_npz(f"{golden_dir}/coeff_tables.npz")
    for key, ref in g.items():
        if key.endswith("_pathparam"):
            continue
        path, sched, st1, st2, N = key.split("_")
        kw = {} if path == "fm" else {"noise_schedule": sched}
        b = Bridge(path, N=int(N[1:]), sampler_type=f"{st1}_{st2}", **kw)
        assert np.array_equal(b.coefficient_table().numpy(), ref), key
        tq = torch.tensor([0.03, 0.25, 0.5, 0.9999, 1.0])
        assert np.array_equal(torch.stack(b.path.path_param(tq)).numpy(), g[key + "_pathparam"]), key


def test_bridge_api_surface():
    from fdbm_b200 import Bridge, BridgeRegistry
    assert set(BridgeRegistry.get_all_names()) >= {"sb", "fm"}
    b = Bridge("sb", N=5)
    assert (b.start_time, b.end_time, b.path.sampling_direction) == (1.0, 1e-4, "reverse")
    f = Bridge("fm", N=30, sampler_type="ode_ei")
    assert (f.start_time, f.end_time, f.path.sampling_direction) == (1e-4, 1.0, "forward")
    b.N, b.sampler_type = 10, "sde_ei"                       # infer_single.py:55-56 mutates these after construction
    assert b.coefficient_table().shape == (10, 3) and float(b.coefficient_table()[-1, 2]) == 0.0
    with pytest.raises(ValueError):                          # the default predictor 'reverse_diffusion' is unregistered, as in
        Bridge("sb", sampler_type="pc").sampler(None, None)  # the reference (util/predictors.py registers euler_maruyama, none)
    with pytest.raises(ValueError):
        Bridge("sb", sampler_type="pc").sampler(None, None, predictor_name="euler_maruyama", corrector_name="bogus")
    assert Bridge("sb", sampler_type="bogus").sampler(None, None) is None      # bridge.py:56-64 falls through
    p = Bridge("sb", noise_schedule="ve").path
    t = torch.tensor([0.4])
    wx, ws, wy = p.ode_weights(t); sx, ss, sy, gd = p.sde_weights(t)
    ox = O.PathSB(noise_schedule="ve")
    one = torch.ones(1, 1, 1, 1)
    assert torch.allclose(ox.ode(t, one, 0 * one, 0 * one).reshape(-1), wx) and torch.allclose(ox.ode(t, 0 * one, one, 0 * one).reshape(-1), ws)
    assert torch.allclose(ox.sde(t, 0 * one, 0 * one, one)[0].reshape(-1), sy) and torch.allclose(ox.sde(t, one, one, one)[1], gd)
    with pytest.raises(ValueError):
        Bridge("nope")
    s = torch.randn(2, 1, 3, 4, dtype=torch.complex64); y = torch.randn(2, 1, 3, 4, dtype=torch.complex64)
    mean, sig = b.probability_path(s, y, torch.tensor([0.3, 1.0]))
    om, osig = O.Bridge("sb").probability_path(s, y, torch.tensor([0.3, 1.0]))
    assert torch.equal(mean, om) and torch.equal(sig, osig)


def test_backbone_registry_and_state_dict_names():
    from fdbm_b200 import BackboneRegistry
    assert set(BackboneRegistry.get_all_names()) >= {"ncsnpp_v2", "ncsnpp_v2_predictive", "ncsnpp_v2_16M", "ncsnpp_v2_5M", "ncsnpp_v2_37M",
                                                     "tfgridnet_5l32c100", "tfgridnet_4l32c80", "tfgridnet_5l32c100_predictive"}   # backbones/__init__.py
    n16 = BackboneRegistry.get_by_name("ncsnpp_v2_16M")(nf=128)                # the variant fixes its own size (ncsnpp_v2.py:418-427)
    assert {k: tuple(v.shape) for k, v in n16.state_dict().items()} == O.param_shapes(O.NcsnppConfig(nf=64, attn_resolutions=(0,)))
    for name, pred, n_par in (("ncsnpp_v2", False, 65590822), ("ncsnpp_v2_predictive", True, None)):
        net = BackboneRegistry.get_by_name(name)(unused_option=1)           # accepts/ignores extra kwargs
        want = O.param_shapes(O.NcsnppConfig(predictive=pred))
        got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        assert got == want
        if n_par:
            assert sum(p.numel() for p in net.parameters()) == n_par
        assert not dict(net.named_parameters())["all_modules.0.W"].requires_grad if not pred else True
    # the nf = 96 variants: reference parameter names / shapes, and the zero-padded image the nf = 128 plan is loaded with
    from fdbm_b200.backbones import _pad_channel_blocks
    for name, cfg in (("ncsnpp_v2_5M", O.NcsnppConfig(nf=96, ch_mult=(1, 1, 1, 1), num_res_blocks=1, attn_resolutions=(0,))),
                      ("ncsnpp_v2_37M", O.NcsnppConfig(nf=96))):
        net = BackboneRegistry.get_by_name(name)(nf=128, ch_mult=(1,))
        assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == O.param_shapes(cfg)
        big = O.param_shapes(O.NcsnppConfig(nf=128, ch_mult=cfg.ch_mult, num_res_blocks=cfg.num_res_blocks, attn_resolutions=cfg.attn_resolutions))
        arch = net._arch()
        assert (arch.nf, arch.channel_block_real) == (128, 96)
        for k, v in net.state_dict().items():
            p = _pad_channel_blocks(k, v)
            assert tuple(p.shape) == big[k], k
            scale = (128 / 96) ** 0.5 if k.endswith(("NIN_0.W", "NIN_0.b")) else 1.0
            assert abs(float(p.double().sum()) - scale * float(v.double().sum())) <= 1e-6 * float(v.double().abs().sum()) + 1e-12   # only zeros were added
    w = torch.arange(192.0)
    p = _pad_channel_blocks("x", w)
    assert torch.equal(p[:96], w[:96]) and torch.equal(p[128:224], w[96:]) and not p[96:128].any() and not p[224:].any()
    with pytest.raises(ValueError):
        BackboneRegistry.get_by_name("ncsnpp")                              # unregistered name, as in the reference
    with pytest.raises(NotImplementedError):
        BackboneRegistry.get_by_name("ncsnpp_v2")(resblock_type="ddpm")


def test_default_init_statistics():
    from fdbm_b200 import BackboneRegistry
    torch.manual_seed(0)
    net = BackboneRegistry.get_by_name("ncsnpp_v2")()
    sd = net.state_dict()
    w = sd["all_modules.4.Conv_0.weight"]                                    # fan_avg uniform, scale 1
    fan = (128 * 9 + 128 * 9) / 2
    assert abs(float(w.var()) - 1.0 / fan) < 0.05 / fan
    assert float(sd["all_modules.4.Conv_1.weight"].abs().max()) < 1e-5       # init_scale = 0 -> 1e-10
    assert float(sd["all_modules.4.Conv_0.bias"].abs().max()) == 0.0
    assert abs(float(sd["all_modules.0.W"].std()) - 16.0) < 3.0              # Fourier scale 16


def test_forward_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fdbm_b200 import BackboneRegistry
    net = BackboneRegistry.get_by_name("ncsnpp_v2")()
    x = torch.zeros(1, 1, 257, 64, dtype=torch.complex64)
    with pytest.raises(RuntimeError):
        net(x, x, torch.ones(1))


def test_split_list_matches_infer_folder():
    from fdbm_b200 import split_list, shard_for_rank, padded_frames
    items = list(range(10))
    assert split_list(items, 4) == [[0, 1, 2], [3, 4, 5], [6, 7], [8, 9]]
    assert split_list(list(range(256)), 8)[3] == list(range(96, 128))
    assert sum(split_list(items, 3), []) == items
    assert shard_for_rank(items, 1, 2) == [5, 6, 7, 8, 9]
    assert split_list([], 2) == [[], []]
    assert [padded_frames(t) for t in (251, 256, 1876, 63, 64)] == [256, 256, 1920, 64, 64]


def test_spec_data_module_attributes():
    from fdbm_b200 import SpecsDataModule, get_window
    dm = SpecsDataModule(base_dir="/unused", n_fft=512, hop_length=256, num_frames=256, window="sqrthann", gpu=False)
    assert (dm.spec_factor, dm.spec_abs_exponent, dm.transform_type, dm.normalize) == (0.15, 0.5, "exponent", "noisy")
    assert torch.equal(dm.window, O.make_window("sqrthann", 512))
    assert dm.stft_kwargs["return_complex"] and dm.istft_kwargs["center"]
    with pytest.raises(NotImplementedError):
        get_window("blackman", 512)


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "rethinking-flow-and-diffusion-bridge-models-for-speech-enhancement_b200"))
    from fdbm_b200 import gather_waveforms, shard_for_rank
    utts = list(range(5))                                                    # 5 utterances over 2 ranks: 3 + 2
    mine = shard_for_rank(utts, rank, world)
    local = torch.stack([torch.full((8,), float(u)) for u in mine])         # "enhanced" = utterance id
    counts = [len(shard_for_rank(utts, r, world)) for r in range(world)]
    out = gather_waveforms(local, counts)
    q.put((rank, out[:, 0].tolist()))
    dist.destroy_process_group()


def test_utterance_sharding_and_gather_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, order in res:
        assert order == [0.0, 1.0, 2.0, 3.0, 4.0]                             # every rank sees all utterances, in order


def _allreduce_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "rethinking-flow-and-diffusion-bridge-models-for-speech-enhancement_b200"))
    from fdbm_b200.training import allreduce_gradients_
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)               # rank r holds (r + 1) * [0..9]
    w = allreduce_gradients_(flat)
    q.put((rank, w, flat.tolist()))
    dist.destroy_process_group()


def test_ddp_gradient_allreduce_world2():
    """The training step's DDP exchange: one all-reduce(sum) of the flat gradient buffer; 1/world goes into grad_div."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, w, flat in res:
        assert w == 2
        assert flat == [3.0 * i for i in range(10)]


def test_length_buckets_group_only_equal_padded_lengths():
    """Variable-length batching (callers' edge of infer_folder.py:91-146): utterances may share a batch only when pad_spec
    gives them the same padded frame count; batches are bounded by the micro-batch and cover every utterance once."""
    from fdbm_b200 import length_buckets, padded_frames
    lens = [19200, 30400, 19911, 40000, 16000, 64000, 63999, 64256, 16384]
    b = length_buckets(lens, 256, 2)
    seen = sorted(i for _, g in b for i in g)
    assert seen == list(range(len(lens)))
    for T_pad, g in b:
        assert 1 <= len(g) <= 2
        assert all(padded_frames(1 + lens[i] // 256) == T_pad for i in g)
        assert [lens[i] for i in g] == sorted(lens[i] for i in g)
    assert [T for T, _ in b] == sorted(T for T, _ in b)
    assert padded_frames(251) == 256 and padded_frames(256) == 256 and padded_frames(257) == 320


def test_rk45_tableau_is_dormand_prince():
    """The constants fdbm_b200.rk45 integrates with are scipy's RK45 tableau (consistency conditions + scipy itself)."""
    from fdbm_b200 import rk45
    from scipy.integrate import RK45
    import numpy as np
    for s, row in enumerate(rk45._A):
        assert abs(sum(row) - rk45._C[s]) < 1e-15
        assert np.allclose(row, RK45.A[s][:s])
    assert np.allclose(rk45._B, RK45.B) and np.allclose(rk45._C, RK45.C) and np.allclose(rk45._E, RK45.E)
    assert abs(sum(rk45._B) - 1) < 1e-15 and abs(sum(rk45._E)) < 1e-15
    assert (rk45.SAFETY, rk45.MIN_FACTOR, rk45.MAX_FACTOR) == (0.9, 0.2, 10.0) and rk45.ERR_EXP == RK45.error_exponent if hasattr(RK45, "error_exponent") else True


def test_on_device_rejects_mixed_devices_and_passes_cpu_through():
    from fdbm_b200._lib import on_device

    @on_device
    def f(a, b=None):
        return "ran"
    assert f(torch.zeros(1), b=torch.zeros(1)) == "ran" and f(None) == "ran"
