"""Streaming audio hand-off: ASR feature ring -> the renderer's [8, dim, 16] window.

Golden blocks come from the reference's own ASR.get_next_feat (tests/golden/make_feature_ring_golden.py).  CPU: the oracle
restatement and the host-side window bookkeeping (WindowBook, with the device gather emulated in numpy) reproduce them
exactly; GPU: radnerf_b200.audio_ring.FeatureRing (rn_feature_window) does, bit for bit -- the path only moves fp32 values."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import golden
from feature_ring_case import CASES, script


def _check(case, blocks, g):
    blocks = np.stack(blocks)
    assert np.array_equal(blocks[:, -1], g[case + "/newest"])
    assert np.array_equal(blocks[:12], g[case + "/head"]) and np.array_equal(blocks[-4:], g[case + "/tail"])
    assert np.array_equal(blocks.astype(np.float64).sum(axis=(2, 3)), g[case + "/sums"])


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_ring_is_the_reference_ring(case):
    from oracle.feature_ring import Ring
    slots, context, dim, ops = script(case)
    ring, blocks = Ring(slots, context, dim), []
    for op in ops:
        if op[0] == "write":
            ring.write(op[1])
        else:
            blocks.append(ring.next_window())
    _check(case, blocks, golden("feature_ring"))


def _emulated_kernel(queue, snapshots, start, snapshot, fresh):
    """numpy statement of rn_feature_window's contract (include/radnerf_b200.h)"""
    size, dim = queue.shape
    out = np.zeros((8, dim, 16), np.float32)
    for k in range(8):
        if start[k] < 0:
            continue
        rows = queue[(start[k] + np.arange(16)) % size]
        if snapshot[k] >= 0:
            if fresh[k]:
                snapshots[snapshot[k]] = rows
            rows = snapshots[snapshot[k]]
        out[k] = rows.T
    return out


@pytest.mark.parametrize("case", list(CASES))
def test_window_book_reproduces_the_reference_blocks(case):
    """host logic only: live windows, snapshot slots for wrapping windows, the four start-up zero windows"""
    from radnerf_b200.audio_ring import WindowBook
    slots, context, dim, ops = script(case)
    queue, snaps = np.zeros((slots * context, dim), np.float32), np.zeros((8, 16, dim), np.float32)
    book, slot, blocks = WindowBook(slots * context), 0, []
    for op in ops:
        if op[0] == "write":
            queue[slot * context: slot * context + op[1].shape[0]] = op[1]
            slot = (slot + 1) % slots
        else:
            start, snapshot, fresh = book.advance()
            live = [s for s in snapshot if s >= 0]
            assert len(live) == len(set(live))                       # no two windows share a snapshot slot
            blocks.append(_emulated_kernel(queue, snaps, start, snapshot, fresh))
    _check(case, blocks, golden("feature_ring"))


def test_ring_abi_without_a_gpu():
    from radnerf_b200 import abi, audio_ring
    L = abi.lib()
    L.rn_sizeof.restype = C.c_uint32
    L.rn_sizeof.argtypes = [C.c_char_p]
    assert L.rn_sizeof(b"rn_ring_windows") == C.sizeof(audio_ring.RingWindows) == 96
    w = audio_ring.RingWindows()
    w.start[:], w.snapshot[:] = [-1] * 8, [-1] * 8
    assert L.rn_feature_window(None, 200, 44, C.byref(w), None, None, None) == -1 and b"null pointer" in L.rn_last_error_string()
    p = C.c_void_p(256)
    assert L.rn_feature_window(p, 8, 44, C.byref(w), None, p, None) == -1 and b"at least one window" in L.rn_last_error_string()
    w.start[3] = 200
    assert L.rn_feature_window(p, 200, 44, C.byref(w), None, p, None) == -1 and b"outside the ring" in L.rn_last_error_string()
    w.start[3], w.snapshot[3] = 190, 2
    assert L.rn_feature_window(p, 200, 44, C.byref(w), None, p, None) == -1 and b"needs the snapshot buffer" in L.rn_last_error_string()
    with pytest.raises(RuntimeError, match="no CPU path"):
        audio_ring.FeatureRing(4, 50, 44, "cpu")
    with pytest.raises(ValueError):
        audio_ring.WindowBook(8)


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_device_ring_is_the_reference_ring(case):
    from radnerf_b200.audio_ring import FeatureRing
    slots, context, dim, ops = script(case)
    ring, blocks = FeatureRing(slots, context, dim, "cuda"), []
    # every other frame is delivered into a slice of a larger block, the way a frame lane's input block receives it
    block = torch.full((24 + 8 * dim * 16 + 5,), -7.0, device="cuda")
    for i, op in enumerate(ops):
        if op[0] == "write":
            ring.push(torch.from_numpy(op[1]).cuda() if i % 3 else torch.from_numpy(op[1]).pin_memory())
        elif i % 2:
            blocks.append(ring.next_window().cpu().numpy())
        else:
            ring.next_window(out=block[24:24 + 8 * dim * 16])
            blocks.append(block[24:24 + 8 * dim * 16].view(8, dim, 16).cpu().numpy())
            assert float(block[23]) == -7.0 and float(block[24 + 8 * dim * 16]) == -7.0
    _check(case, blocks, golden("feature_ring"))


@pytest.mark.gpu
def test_streamed_frames_from_the_device_ring_equal_frames_from_host_windows():
    """FrameStreamer.submit(head, ring=...) -- the audio window gathered on the device straight into the lane's input block --
    renders the same images as submitting complete host blocks whose windows come from the oracle ring (first frames through
    the Python path, later ones through the one-call C path; lip smoothing on, so the frames are chained)"""
    import bench
    from oracle.feature_ring import Ring
    from radnerf_b200.audio_ring import FeatureRing
    from radnerf_b200.stream import FrameStreamer, pack_inputs
    hw, n, slots, context, dim = 64, 24, 4, 10, 44
    frames, intr, bg = bench.make_frames(hw, n)
    bg_t = torch.from_numpy(bg).cuda()
    rng = np.random.default_rng(5)
    writes = {i: (rng.standard_normal((context, dim)) * 3).astype(np.float32) for i in range(0, n, 5)}
    cpu_ring, blocks = Ring(slots, context, dim), []
    for i in range(n):
        if i in writes:
            cpu_ring.write(writes[i])
        blocks.append(cpu_ring.next_window())
    assert np.abs(blocks[-1]).max() > 0
    images = []
    for mode in ("host", "ring"):
        model = bench.make_model("cuda", seed=9)
        kw = model.opt.render_kwargs()
        model.enc_a = None
        streamer = FrameStreamer(model, hw, hw, intr, bg_t, (8, dim, 16), use_eye=True, depth=2, **kw)
        if mode == "host":
            got = [img.clone() for img in streamer.render_all([pack_inputs(f["pose"], b, f["pose6"], f["eye"]) for f, b in zip(frames, blocks)])]
        else:
            ring, got = FeatureRing(slots, context, dim, "cuda"), []
            for i, f in enumerate(frames):
                if i in writes:
                    ring.push(torch.from_numpy(writes[i]).cuda())
                streamer.submit(pack_inputs(f["pose"], np.zeros(0, np.float32), f["pose6"], f["eye"]), ring=ring)
                if streamer.in_flight() == 2:
                    got.append(streamer.collect().clone())
            while streamer.in_flight():
                got.append(streamer.collect().clone())
            assert any(x is not None for x in streamer.fast)        # the one-call path was reached
        streamer.close()
        images.append(got)
    assert len(images[0]) == len(images[1]) == n
    for i, (a, b) in enumerate(zip(*images)):
        assert torch.equal(a, b), i
    assert not torch.equal(images[1][3], images[1][20])


def test_window_book_random_schedules_against_the_oracle_ring():
    """property check over random ring geometries and write/read schedules (ring sizes from one window up, writes of ragged
    lengths at arbitrary times): host bookkeeping + the kernel's contract == the restated reference, window for window"""
    from oracle.feature_ring import Ring
    from radnerf_b200.audio_ring import WindowBook
    rng = np.random.default_rng(2024)
    for trial in range(40):
        slots, context, dim = int(rng.integers(1, 6)), int(rng.integers(4, 40)), int(rng.integers(1, 9))
        if slots * context < 16:
            slots = -(-16 // context)
        size = slots * context
        ref, book = Ring(slots, context, dim), WindowBook(size)
        queue, snaps, slot = np.zeros((size, dim), np.float32), np.zeros((8, 16, dim), np.float32), 0
        for step in range(int(rng.integers(20, 120))):
            if rng.random() < 0.3:
                feats = rng.standard_normal((int(rng.integers(1, context + 1)), dim)).astype(np.float32)
                ref.write(feats)
                queue[slot * context: slot * context + feats.shape[0]] = feats
                slot = (slot + 1) % slots
            else:
                start, snapshot, fresh = book.advance()
                got = _emulated_kernel(queue, snaps, start, snapshot, fresh)
                assert np.array_equal(got, ref.next_window()), (trial, step, slots, context)
