"""Training-step tail: Adam + GradScaler unscale / skip + zero_grad + EMA in one sweep.

CPU part: the oracle (oracle.c o_adam_step / o_ema_update) is pinned against torch.optim.Adam itself -- the third-party
implementation the reference calls (main.py:204, nerf/utils.py:1171-1173) -- run on the CPU here, and the host-side
logic (chunk layout, descriptor sizes, argument errors, no CPU path).  GPU part: the kernel against the oracle, BIT FOR BIT
(both spell every rounding: fmaf / IEEE sqrt / division), and FusedAdam / ParamEMA against torch's CUDA Adam through
torch.amp.GradScaler within fp32 round-off (torch's foreach kernels contract differently: tolerance 2e-6 relative, stated
below)."""
import copy
import ctypes as C

import numpy as np
import pytest
import torch

BETAS, EPS = (0.9, 0.99), 1e-15            # main.py:204


def _grads(rng, n):
    """sparse-ish gradients over 6 decades, like the rows of a hash table that a batch touches"""
    return (rng.standard_normal(n) * (rng.random(n) < 0.3) * 10 ** rng.uniform(-6, 0, n)).astype(np.float32)


@pytest.mark.parametrize("n", [1, 5, 4097, 50000])
def test_oracle_adam_is_torch_adam(oracle, n):
    """30 steps under the reference's LambdaLR decay: moments bit-identical to torch (CPU, single-tensor path), parameters
    identical except where torch's vectorised CPU sqrt (Sleef) is not correctly rounded: <= 0.2% of elements, <= 1.2e-7"""
    rng = np.random.default_rng(n)
    torch.manual_seed(n)
    p0 = torch.randn(n)
    pt = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([pt], lr=5e-3, betas=BETAS, eps=EPS, foreach=False)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 0.1 ** (it / 30))       # main.py:219
    p, m, v = p0.numpy().copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for step in range(1, 31):
        g = _grads(rng, n)
        pt.grad = torch.from_numpy(g.copy())
        lr = opt.param_groups[0]["lr"]
        opt.step()
        sched.step()
        oracle.adam_step(p, g, m, v, lr, BETAS, EPS, 0.0, step)
    st = opt.state[pt]
    assert np.array_equal(m, st["exp_avg"].numpy()) and np.array_equal(v, st["exp_avg_sq"].numpy())
    ref = pt.detach().numpy()
    assert (p != ref).mean() <= 2e-3 and np.abs(p - ref).max() <= 1.2e-7


def test_oracle_adam_weight_decay_and_loss_scale(oracle):
    """weight decay (get_params(..., wd), nerf/network.py:330) and a GradScaler-scaled gradient (scale 2^16: unscaling by a
    power of two is exact, so the step equals the unscaled one)"""
    rng = np.random.default_rng(7)
    n = 3000
    p0 = rng.standard_normal(n).astype(np.float32)
    pt = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([pt], lr=5e-4, betas=BETAS, eps=EPS, weight_decay=0.01, foreach=False)
    p, m, v = p0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for step in range(1, 11):
        g = _grads(rng, n)
        pt.grad = torch.from_numpy(g.copy())
        opt.step()
        oracle.adam_step(p, g * np.float32(65536.0), m, v, 5e-4, BETAS, EPS, 0.01, step, inv_scale=1.0 / 65536.0)
    assert np.abs(m - opt.state[pt]["exp_avg"].numpy()).max() <= 1e-9      # the scalar tail of torch's add(alpha) is not an fma
    assert np.abs(p - pt.detach().numpy()).max() <= 2e-7


def test_oracle_ema_is_the_torch_ema_update(oracle):
    """torch_ema `update`: tmp = s - p; tmp.mul_(1 - decay); s.sub_(tmp), decay = min(decay, (1 + k) / (10 + k))"""
    torch.manual_seed(3)
    s, p = torch.randn(5000), torch.randn(5000)
    shadow = s.numpy().copy()
    for k in range(1, 6):
        decay = min(0.95, (1 + k) / (10 + k))
        tmp = s - p
        tmp.mul_(1.0 - decay)
        s.sub_(tmp)
        oracle.ema_update(shadow, p.numpy(), decay)
        p = p + 0.01
    assert np.array_equal(shadow, s.numpy())


def test_chunk_layout_and_descriptor_sizes():
    from radnerf_b200 import abi, optim
    first, total = optim.chunk_layout([1, 4096, 4097, 0, 10])
    assert first == [0, 1, 2, 4, 4] and total == 5
    L = abi.lib()
    L.rn_sizeof.restype = C.c_uint32
    L.rn_sizeof.argtypes = [C.c_char_p]
    assert L.rn_sizeof(b"rn_adam_tensor") == C.sizeof(optim.AdamTensor) == 64
    assert L.rn_sizeof(b"rn_adam_group") == C.sizeof(optim.AdamGroup) == 40


def test_argument_errors_without_a_gpu():
    from radnerf_b200 import abi, optim
    L = abi.lib()
    g = (optim.AdamGroup * 1)()
    g[0].beta1, g[0].beta2 = 0.9, 0.99
    assert L.rn_adam_step(None, 0, 0, g, 1, None, None, 0, None) == 0             # nothing to do
    assert L.rn_adam_step(None, 1, 1, g, 1, None, None, 0, None) == -1 and b"null pointer" in L.rn_last_error_string()
    assert L.rn_adam_step(C.c_void_p(16), 1, 1, g, 0, None, None, 0, None) == -1 and b"parameter groups" in L.rn_last_error_string()
    g[0].beta2 = 1.0
    assert L.rn_adam_step(C.c_void_p(16), 1, 1, g, 1, None, None, 0, None) == -1 and b"betas" in L.rn_last_error_string()
    assert L.rn_ema_update(C.c_void_p(16), 1, 1, 1.5, None) == -1 and b"decay" in L.rn_last_error_string()


def test_no_cpu_path():
    from radnerf_b200.optim import FusedAdam, ParamEMA
    p = torch.nn.Parameter(torch.zeros(8))
    p.grad = torch.ones(8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        FusedAdam([p], lr=1e-3).step()
    with pytest.raises(RuntimeError, match="no CPU path"):
        ParamEMA([p], 0.95)
    with pytest.raises(ValueError):
        FusedAdam([p], lr=1e-3, betas=(1.0, 0.99))


# ----------------------------------------------------------------------------------------------------------------- GPU

def _tensor_set(rng, dev):
    """ragged sizes, one deliberately 4-byte-aligned view (scalar path), three groups with their own lr / weight decay"""
    sizes = [1, 7, 4096, 4097, 64 * 64, 3 * 4096 + 5, 200001, 33]
    groups = [0, 0, 1, 1, 2, 2, 1, 0]
    params = []
    for i, n in enumerate(sizes):
        if i == 5:      # a view that starts one float into its storage: not 16-byte aligned
            t = torch.from_numpy(rng.standard_normal(n + 1).astype(np.float32)).to(dev)[1:]
        else:
            t = torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(dev)
        params.append(torch.nn.Parameter(t))
    return params, groups


@pytest.mark.gpu
def test_kernel_is_bit_identical_to_the_oracle(oracle):
    from radnerf_b200.optim import FusedAdam
    dev = "cuda"
    rng = np.random.default_rng(11)
    params, gidx = _tensor_set(rng, dev)
    hyper = [dict(lr=5e-3, weight_decay=0.0), dict(lr=5e-4, weight_decay=0.0), dict(lr=2.5e-3, weight_decay=0.01)]
    opt = FusedAdam([dict(params=[p for p, g in zip(params, gidx) if g == k], **hyper[k]) for k in range(3)],
                    betas=BETAS, eps=EPS, zero_grads=True)
    state = [(p.detach().cpu().numpy().copy(), np.zeros(p.numel(), np.float32), np.zeros(p.numel(), np.float32)) for p in params]
    scale = torch.tensor(1024.0, device=dev)
    taken = 0
    for it in range(1, 9):
        gs = [_grads(rng, p.numel()) for p in params]
        for p, g in zip(params, gs):
            g32 = torch.from_numpy(g * np.float32(1024.0)).to(dev)
            if p.grad is None:
                p.grad = g32
            else:
                assert not p.grad.any()                   # the sweep left the gradients zeroed
                p.grad.copy_(g32)
        skipped = it == 4
        opt.grad_scale, opt.found_inf = scale, torch.tensor(1.0 if skipped else 0.0, device=dev)
        for grp in opt.param_groups:                      # a schedule: the group table is rebuilt every step
            grp["lr"] *= 0.97
        opt.step()
        if not skipped:
            taken += 1
            for (p, m, v), g, k in zip(state, gs, gidx):
                oracle.adam_step(p, g * np.float32(1024.0), m, v, opt.param_groups[k]["lr"], BETAS, EPS,
                                 hyper[k]["weight_decay"], taken, inv_scale=1.0 / 1024.0)
        for prm, (p, m, v) in zip(params, state):
            st = opt.state[prm]
            assert float(st["step"]) == taken
            assert np.array_equal(st["exp_avg"].cpu().numpy().ravel(), m)
            assert np.array_equal(st["exp_avg_sq"].cpu().numpy().ravel(), v)
            assert np.array_equal(prm.detach().cpu().numpy().ravel(), p), (it, prm.numel())
            assert not prm.grad.any()


@pytest.mark.gpu
def test_fused_adam_follows_torch_adam_through_gradscaler():
    """same model, same batches, torch.optim.Adam (CUDA foreach) vs FusedAdam, both driven by torch.amp.GradScaler with an
    overflow in step 3 (skipped by both, scale halved).  fp32 tolerance: torch's foreach functors and this kernel round
    a few operations differently (fma contraction), 12 steps at lr 5e-3 stay within 2e-6 relative / 1e-7 absolute"""
    from radnerf_b200.optim import FusedAdam
    dev = "cuda"

    def make():
        torch.manual_seed(5)
        return torch.nn.Sequential(torch.nn.Linear(32, 64), torch.nn.ReLU(), torch.nn.Linear(64, 3, bias=False)).to(dev)

    nets = [make(), make()]
    opts = [torch.optim.Adam(nets[0].parameters(), lr=5e-3, betas=BETAS, eps=EPS),
            FusedAdam(nets[1].parameters(), lr=5e-3, betas=BETAS, eps=EPS, zero_grads=True)]
    scalers = [torch.amp.GradScaler("cuda", init_scale=2.0 ** 16), torch.amp.GradScaler("cuda", init_scale=2.0 ** 16)]
    torch.manual_seed(6)
    xs = torch.randn(12, 256, 32, device=dev)
    for it in range(12):
        for net, opt, sc in zip(nets, opts, scalers):
            opt.zero_grad(set_to_none=False)
            loss = net(xs[it]).square().mean()
            sc.scale(loss).backward()
            if it == 3:
                next(net.parameters()).grad[0, 0] = float("inf")
            sc.step(opt)
            sc.update()
    assert float(scalers[0].get_scale()) == float(scalers[1].get_scale()) == 2.0 ** 15
    for a, b in zip(nets[0].parameters(), nets[1].parameters()):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), float((a - b).abs().max())
    for a, b in zip(nets[0].parameters(), nets[1].parameters()):
        assert float(opts[0].state[a]["step"]) == float(opts[1].state[b]["step"]) == 11
    # a checkpoint written by torch.optim.Adam loads into FusedAdam (and the run continues identically)
    third = make()
    third.load_state_dict(nets[0].state_dict())
    opt3 = FusedAdam(third.parameters(), lr=5e-3, betas=BETAS, eps=EPS)
    opt3.load_state_dict(copy.deepcopy(opts[0].state_dict()))     # as a checkpoint round trip would (load_state_dict shares tensors)
    for net, opt in ((nets[0], opts[0]), (third, opt3)):
        opt.zero_grad(set_to_none=False)
        net(xs[0]).square().mean().backward()
        opt.step()
    for a, b in zip(nets[0].parameters(), third.parameters()):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7)


@pytest.mark.gpu
def test_param_ema_matches_the_oracle_and_restores(oracle):
    from radnerf_b200.optim import ParamEMA
    dev = "cuda"
    rng = np.random.default_rng(2)
    params = [torch.nn.Parameter(torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(dev)) for n in (5, 4096, 70001)]
    ema = ParamEMA(params, decay=0.95)
    shadows = [p.detach().cpu().numpy().copy() for p in params]
    for k in range(1, 5):
        with torch.no_grad():
            for p in params:
                p.add_(0.1 * torch.randn_like(p))
        ema.update()
        for s, p in zip(shadows, params):
            oracle.ema_update(s, p.detach().cpu().numpy(), min(0.95, (1 + k) / (10 + k)))
    for s, t in zip(shadows, ema.shadow_params):
        assert np.array_equal(s, t.cpu().numpy())
    live = [p.detach().clone() for p in params]
    ema.store()
    ema.copy_to()
    assert all(torch.equal(p, s) for p, s in zip(params, ema.shadow_params))
    ema.restore()
    assert all(torch.equal(p, l) for p, l in zip(params, live))


def test_every_chunk_finds_its_tensor():
    """the kernel's per-CTA binary search over `first_chunk` (csrc/optim_tail.cu find_tensor), restated: for random tensor sizes
    including empty tensors, chunk c belongs to the tensor whose chunk range contains it"""
    from radnerf_b200 import optim
    rng = np.random.default_rng(9)
    for _ in range(200):
        sizes = [int(s) for s in rng.choice([0, 1, 5, 4095, 4096, 4097, 10000, 70001], size=int(rng.integers(1, 12)))]
        first, total = optim.chunk_layout(sizes)
        owner = []
        for t, n in enumerate(sizes):
            owner += [t] * ((n + optim.CHUNK - 1) // optim.CHUNK)
        assert len(owner) == total
        for c in range(total):
            lo, hi = 0, len(sizes)                      # last tensor with first_chunk <= c
            while hi - lo > 1:
                mid = (lo + hi) >> 1
                lo, hi = (mid, hi) if first[mid] <= c else (lo, mid)
            assert lo == owner[c], (sizes, c)
            assert 0 <= (c - first[lo]) * optim.CHUNK < sizes[lo]
