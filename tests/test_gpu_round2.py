"""Round-2 kernels in isolation: the MN-major tcgen05 contraction (weight gradients) and the restructured 3-D table backward."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("Kx,Ky,passes", [(16, 16, 1), (32, 64, 1), (64, 64, 1), (64, 16, 1), (80, 64, 1), (64, 80, 3), (128, 128, 2), (32, 48, 5)])
def test_tcgen05_mn_major_contraction_matches_torch_fp32(Kx, Ky, passes):
    """out[m, n] = passes * sum_s X[s, m] Y[s, n]: both operands are the SAME interleaved tiles the forward writes, read MN-major
    (samples = the MMA's K dimension), accumulated over `passes` issues in TMEM -- vs a plain PyTorch fp32 matmul."""
    from radnerf_b200 import abi as L
    g = torch.Generator(device="cpu").manual_seed(Kx * 1000 + Ky)
    X = torch.randn(128, Kx, generator=g).half().to(DEV)
    Y = torch.randn(128, Ky, generator=g).half().to(DEV)
    out = torch.full((128, Ky), float("nan"), device=DEV)
    L.check(L.lib().rn_selftest_umma_mn(L.ptr(X), L.ptr(Y), L.ptr(out), Kx, Ky, passes, L.cur_stream()))
    torch.cuda.synchronize()
    ref = passes * (X.float().t() @ Y.float())
    got = out[:Kx]
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item()) * passes


def _bwd3_case(B, seed, clustered):
    from gridencoder import GridEncoder
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=16, desired_resolution=2048,
                      gridtype='tiled').to(DEV)
    g = torch.Generator(device="cpu").manual_seed(seed)
    if clustered:   # rays: runs of 16 consecutive samples 0.0135 apart (in [0,1] units), like march_rays_train's packing
        o = torch.rand(B // 16, 1, 3, generator=g) * 0.3 + 0.3
        d = torch.nn.functional.normalize(torch.randn(B // 16, 1, 3, generator=g), dim=-1)
        x = (o + d * torch.arange(16).view(1, 16, 1) * 0.0135).reshape(-1, 3)
    else:
        x = torch.rand(B, 3, generator=g)
    x[::97] = 1.5   # out-of-range samples are skipped
    grad = (torch.randn(x.shape[0], 32, generator=g) * 0.1).half()
    return enc, x.to(DEV).contiguous(), grad.to(DEV).contiguous()


@pytest.mark.parametrize("clustered", [False, True])
@pytest.mark.parametrize("variant", [1, 3, 5, 7])
def test_grid_backward3_variants_match_the_generic_kernel(variant, clustered):
    from radnerf_b200 import abi
    enc, x, grad = _bwd3_case(16 * 4000 + 16, 7, clustered)
    B, rows = x.shape[0], enc.embeddings.shape[0]
    S, H = float(np.log2(enc.per_level_scale)), int(enc.base_resolution)
    ref = torch.zeros(rows, 2, device=DEV, dtype=torch.float64)
    # double-precision restatement of the scatter with torch ops, per level
    offs = enc.offsets.cpu().numpy()
    xv = x.double()
    ok = ((xv >= 0) & (xv <= 1)).all(1)
    scales = torch.empty(16, device=DEV)
    abi.call("rn_grid_level_geometry", S, H, 16, scales, None)
    for l in range(16):
        sc = float(scales[l].item())
        res = int(np.ceil(sc)) + 1
        size = int(offs[l + 1] - offs[l])
        pos = (x.double() * sc + 0.5).float()        # one rounding, as the kernel's fmaf
        pg = pos.floor()
        fr = (pos - pg).double()
        pg = pg.long()
        s1, s2 = res + 1, (res + 1) ** 2
        use = [True, s1 <= size, s2 <= size]
        for k in range(8):
            w = torch.ones(B, device=DEV, dtype=torch.float64)
            idx = torch.zeros(B, device=DEV, dtype=torch.long)
            for d, st in enumerate((1, s1, s2)):
                bit = (k >> d) & 1
                w = w * (fr[:, d] if bit else 1 - fr[:, d])
                if use[d] and all(use[:d + 1]):
                    idx = idx + (pg[:, d] + bit) * st
            idx = idx % size + int(offs[l])
            contrib = (w[:, None] * grad[:, 2 * l:2 * l + 2].double())[ok]
            ref.index_add_(0, idx[ok], contrib)
    out = torch.zeros(rows, 2, device=DEV)
    sizes = offs[1:] - offs[:-1]
    abi.check(abi.lib().rn_grid_backward3(abi.ptr(grad), abi.ptr(x), abi.ptr(enc.offsets), abi.ptr(out), B, 16, S, H, 1, 1, variant, 0xffff, 6,
                                          2, int(sizes[0] + sizes[1]), abi.cur_stream()))
    gen = torch.zeros(rows, 2, device=DEV)
    abi.call("rn_grid_encode_backward", grad, x, None, enc.offsets, gen, B, 3, 2, 16, S, H, None, None, 1, 0, 0, 1, 1, 0)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    assert (gen.double() - ref).abs().max().item() <= 1e-5 * scale      # the generic kernel vs the double restatement
    assert (out.double() - ref).abs().max().item() <= 1e-5 * scale, variant


def test_occupancy_merge_and_device_threshold_packbits():
    """rn_occupancy_merge + rn_packbits_min (the occupancy update without host synchronisation) against the reference's expressions
    (nerf/renderer.py:463-471): masked max(grid * decay, fresh), mean(clamp(grid, 0)), packbits at min(mean, density_thresh)"""
    import raymarching as rm
    from radnerf_b200 import abi
    g = torch.Generator().manual_seed(5)
    n = 128 ** 3
    grid = (torch.rand(n, generator=g) * 30).cuda()
    grid[torch.rand(n, generator=g).cuda() < 0.2] = -1.0           # untrained cells
    fresh = (torch.rand(n, generator=g) * 30).cuda()
    fresh[torch.rand(n, generator=g).cuda() < 0.1] = -1.0
    want = grid.clone()
    both = (want >= 0) & (fresh >= 0)
    want[both] = torch.maximum(want[both] * 0.95, fresh[both])
    want_mean = want.clamp(min=0).double().mean()
    ws = torch.zeros(int(abi.lib().rn_occupancy_merge_workspace_bytes()) // 8 + 1, dtype=torch.float64, device="cuda")
    mean = torch.empty(1, device="cuda")
    for rep in range(2):        # twice: the kernel re-arms its ticket
        got = grid.clone()
        abi.call("rn_occupancy_merge", got, fresh, n, 0.95, ws, mean)
        assert torch.equal(got, want)
        assert abs(float(mean) - float(want_mean)) <= 1e-6 * float(want_mean)
    for thresh in (1e9, 5.0):   # the mean wins / density_thresh wins
        bits = torch.empty(n // 8, dtype=torch.uint8, device="cuda")
        abi.call("rn_packbits_min", got, n // 8, float(thresh), mean, bits)
        ref = rm.packbits(got.view(1, -1), min(float(mean), thresh))
        assert torch.equal(bits, ref.view(-1))


@pytest.mark.parametrize("C,H,fill", [(1, 128, "head"), (2, 64, "blobs"), (1, 32, "empty"), (1, 64, "full")])
def test_occupancy_pack_is_the_bitfield_restricted_to_its_boxes(C, H, fill):
    """rn_occupancy_pack (the copy of the bitfield the fused marcher stages in shared memory): per cascade the box is exactly the
    bounding box of the set bits, every cell inside it carries the bitfield's bit, and the world box is the inflated hull"""
    import numpy as np
    from radnerf_b200 import abi, synthetic as syn
    rng = np.random.default_rng(C * 1000 + H)
    occ = np.zeros((C, H, H, H), bool)                      # [level, x, y, z]
    if fill == "head":
        grid = syn.head_density_grid(H, semi_axes=(0.34, 0.24, 0.37))          # Morton-ordered density, one cascade
        bits_np = syn.packbits_np(grid, min(float(np.clip(grid, 0, None).mean()), 10.0))
    else:
        if fill == "blobs":
            occ[0, 10:30, 5:9, 40:64] = rng.random((20, 4, 24)) < 0.5
            occ[1, 3:5, 60:64, 0:2] = True
        elif fill == "full":
            occ[:] = True
        from oracle import oracle as O
        xs = np.stack(np.meshgrid(np.arange(H), np.arange(H), np.arange(H), indexing="ij"), -1).reshape(-1, 3).astype(np.int32)
        m = O.morton3D(xs).astype(np.int64)
        flat = np.zeros((C, H ** 3), np.uint8)
        for l in range(C):
            flat[l, m] = occ[l].reshape(-1)
        bits_np = np.packbits(flat.reshape(-1), bitorder="little")
    bf = torch.from_numpy(bits_np).cuda()
    pack = torch.zeros(int(abi.lib().rn_occupancy_pack_bytes()), dtype=torch.uint8, device="cuda")
    bound = float(2 ** (C - 1))
    abi.check(abi.lib().rn_occupancy_pack(abi.ptr(bf), C, H, bound, abi.ptr(pack), abi.cur_stream()))
    torch.cuda.synchronize()
    raw = pack.cpu().numpy()
    hdr = raw[:16].view(np.int32)
    box = raw[16:40].view(np.float32)
    lv = raw[48:48 + 32 * 16].view(np.int32).reshape(16, 8)
    words = raw[560:].view(np.uint32)
    # decode the bitfield on the host
    from oracle import oracle as O
    xs = np.stack(np.meshgrid(np.arange(H), np.arange(H), np.arange(H), indexing="ij"), -1).reshape(-1, 3).astype(np.int32)
    m = O.morton3D(xs).astype(np.int64)
    allbits = np.unpackbits(bits_np, bitorder="little").reshape(C, H ** 3)
    want_lo, want_hi, total = np.full(3, np.inf), np.full(3, -np.inf), 0
    for l in range(C):
        cells = allbits[l][m].reshape(H, H, H).astype(bool)
        lo, dim, off = lv[l, 0:3], lv[l, 3:6], lv[l, 6]
        if not cells.any():
            assert (dim == 0).all()
            continue
        idx = np.argwhere(cells)
        assert (lo == idx.min(0)).all() and (lo + dim - 1 == idx.max(0)).all()
        n = int(dim.prod())
        sub = cells[lo[0]:lo[0] + dim[0], lo[1]:lo[1] + dim[1], lo[2]:lo[2] + dim[2]]
        got = np.unpackbits(words[off:off + (n + 31) // 32].view(np.uint8), bitorder="little")[:n].reshape(dim[2], dim[1], dim[0])
        assert hdr[2] == 1 and np.array_equal(got.astype(bool), sub.transpose(2, 1, 0))      # x fastest
        b = min(2.0 ** l, bound)
        cell = 2 * b / H
        want_lo = np.minimum(want_lo, idx.min(0) * cell - b - cell)
        want_hi = np.maximum(want_hi, (idx.max(0) + 1) * cell - b + cell)
        assert off == total
        total += (n + 31) // 32
    assert hdr[0] == C and hdr[1] == total
    if total:
        assert np.allclose(box[:3], want_lo, atol=1e-6) and np.allclose(box[3:], want_hi, atol=1e-6)
    else:
        assert hdr[2] == 0 and (box[:3] > box[3:]).all()        # empty grid: every ray is pruned


def test_gather_to_root_protocol_with_two_ranks_emulated_on_one_gpu():
    """rn_scatter_rows_to_root + rn_stage_frame_at_root (arrival counter / consumed flags, csrc/peer_gather.cu) with TWO ranks emulated
    on one device: both 'ranks' own a frame buffer and a control block in this process's memory, their kernels run on separate streams.
    Every staged frame must be exactly frame i (fp32 and uint8 staging), the arrival counter and the consumed flags must end at their
    expected values.  The real
    multi-process run is tools/gather_to_root_check.py (profiles/r02_gather_to_root_2gpu.json)."""
    import numpy as np
    from radnerf_b200 import abi
    from radnerf_b200.sharding import local_pixel_ids
    L = abi.lib()
    H = W = 64
    world, n_slots, n_frames = 2, 2, 12
    ids = [local_pixel_ids(H, W, world, r).to(torch.int32).cuda() for r in range(world)]
    n_local = ids[0].numel()
    frames_buf = [[torch.zeros(H * W, 3, device="cuda") for _ in range(n_slots)] for _ in range(world)]
    ctrl = [torch.zeros(2 * n_slots + 2, dtype=torch.int64, device="cuda") for _ in range(world)]
    frame_peers = [torch.tensor([frames_buf[r][k].data_ptr() for r in range(world)], dtype=torch.int64, device="cuda") for k in range(n_slots)]
    ctrl_peers = torch.tensor([c.data_ptr() for c in ctrl], dtype=torch.int64, device="cuda")
    tickets = torch.zeros(n_slots, dtype=torch.int32, device="cuda")
    ctas = int(L.rn_scatter_signal_ctas(n_local, W))
    assert ctas >= 1
    base = torch.rand(H * W, 3, generator=torch.Generator().manual_seed(1)).cuda()
    rows = [base.index_select(0, i.long()).contiguous() for i in ids]
    staged32 = [torch.zeros(H * W, 3, device="cuda") for _ in range(n_frames)]
    staged8 = [torch.zeros(H * W * 3, dtype=torch.uint8, device="cuda") for _ in range(n_frames)]
    s_root, s_peer = torch.cuda.Stream(), torch.cuda.Stream()
    # every buffer exists before the first launch: an allocation (cudaMalloc) while the peer's kernel spins on a flag that only later
    # launches of this same thread can set would stall the host against the device
    locs = [[(rows[r] * ((i + 1) / 16.0)).contiguous() for i in range(n_frames)] for r in range(world)]
    torch.cuda.synchronize()
    # launches interleaved in frame order (as the two processes would issue them): every wait is on a launch that precedes it even if
    # both streams share a hardware queue; the peer still runs a frame ahead of the root's staging, so frame i+2 of a buffer is only
    # written after the consumed flag of frame i
    for i in range(n_frames):
        k, seq = i % n_slots, i // n_slots + 1
        with torch.cuda.stream(s_peer):
            abi.check(L.rn_scatter_rows_to_root(abi.ptr(locs[1][i]), abi.ptr(ids[1]), n_local, W, abi.ptr(frame_peers[k]), abi.ptr(ctrl_peers), world, 1, 0, k,
                                                seq, C.c_void_p(s_peer.cuda_stream)))
        with torch.cuda.stream(s_root):
            abi.check(L.rn_scatter_rows_to_root(abi.ptr(locs[0][i]), abi.ptr(ids[0]), n_local, W, abi.ptr(frame_peers[k]), abi.ptr(ctrl_peers), world, 0, 0, k,
                                                seq, C.c_void_p(s_root.cuda_stream)))
            u8 = i % 2 == 1
            abi.check(L.rn_stage_frame_at_root(abi.ptr(frames_buf[0][k]), abi.ptr(staged8[i] if u8 else staged32[i]), H * W * 3, 1 if u8 else 0,
                                               abi.ptr(ctrl_peers), world, 0, k, seq, ctas, tickets[k:k + 1].data_ptr(), C.c_void_p(s_root.cuda_stream)))
    torch.cuda.synchronize()
    for i in range(n_frames):
        want = base * ((i + 1) / 16.0)
        if i % 2 == 1:
            assert torch.equal(staged8[i].view(-1, 3), (want * 255.0).to(torch.uint8)), i
        else:
            assert torch.equal(staged32[i], want), i
    assert int(ctrl[0][0]) == (n_frames // n_slots) * ctas and int(ctrl[1][1]) == n_frames // n_slots       # arrivals at the root, frames consumed


@pytest.mark.parametrize("D,dtype,kb", [(3, torch.float16, 24), (3, torch.float16, 100), (3, torch.float32, 160), (2, torch.float16, 64),
                                        (2, torch.float32, 200)])
def test_grid_forward_with_coarse_levels_staged_in_shared_memory_is_bit_identical(D, dtype, kb, monkeypatch):
    """RADNERF_GRID_STAGE_KB: the persistent forward that bulk-copies the leading dense levels into shared memory
    (gridencoder_impl.cuh, STAGE) must return exactly the words of the plain kernel -- outputs and dy_dx, in-range, out-of-range
    and ragged-tail rows included"""
    from gridencoder import GridEncoder
    torch.manual_seed(D * 7 + kb)
    enc = GridEncoder(input_dim=D, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=16, desired_resolution=2048).cuda()
    with torch.no_grad():
        enc.embeddings.uniform_(-1, 1)
    B = 4 * 148 * 512 + 1237
    x = (torch.rand(B, D, device="cuda") * 2.2 - 1.1)          # ~9 % of the rows per axis fall outside [-1, 1]

    def run():
        xi = x.clone().requires_grad_(True)                     # dy_dx is produced when the input needs a gradient
        with torch.autocast("cuda", dtype=torch.float16, enabled=dtype == torch.float16):
            out = enc(xi, bound=1)
        g, = torch.autograd.grad(out.float().square().sum(), xi)
        return out.detach().clone(), g.detach().clone()

    monkeypatch.delenv("RADNERF_GRID_STAGE_KB", raising=False)
    plain, plain_g = run()
    monkeypatch.setenv("RADNERF_GRID_STAGE_KB", str(kb))
    staged, staged_g = run()
    assert plain.dtype == dtype
    assert torch.equal(plain.view(torch.int16 if dtype == torch.float16 else torch.int32),
                       staged.view(torch.int16 if dtype == torch.float16 else torch.int32))
    assert torch.equal(plain_g, staged_g)
