"""Multi-GPU inference: each frame's rays are split by interleaved row tiles, every rank renders its tiles with replicated
weights / hash tables / occupancy bitfield, and one all-gather over NCCL (NVLink) reassembles the image.

Row tile t (TILE_ROWS image rows) goes to rank t mod world: the head sits in the middle rows, so interleaving balances the
number of occupied samples per rank.  There is no collective on the data path other than the final image all-gather;
`encode_audio` and the lip-smoothing state are recomputed identically on every rank (deterministic, ~1 MFLOP).

Implementations of the gather:
  * NCCL `all_gather_into_tensor` + an un-permute gather kernel (default, works everywhere): `gather()`;
  * `enable_peer_gather()` + `gather()`: every rank stores its finished rows directly at their final position in every rank's
    frame buffer (torch symmetric memory = the same allocation mapped into all processes over NVLink/NVSwitch;
    csrc/peer_gather.cu) followed by a symmetric-memory barrier;
  * `enable_peer_gather()` + `gather_to_root()` (what FrameStreamer uses): frames are delivered by ONE rank, so only the root's
    buffer is assembled.  No barrier and no host call between rendering and delivery: a non-root rank's scatter kernel waits on a
    `consumed` flag (the root has staged the previous frame of that buffer -- the write-after-read hazard between frames in flight),
    stores its rows into the root's buffer and bumps the root's arrival counter; the root's staging kernel waits for the arrivals,
    copies the frame out and releases the buffer (flags in a symmetric control block, release/acquire at system scope).
"""
import torch
import torch.distributed as dist

TILE_ROWS = 8


def local_pixel_ids(H, W, world, rank, tile_rows=TILE_ROWS):
    rows = torch.arange(H)
    mine = rows[(rows // tile_rows) % world == rank]
    return (mine[:, None] * W + torch.arange(W)[None, :]).reshape(-1)


class FrameSharder:
    def __init__(self, H, W, world, rank, device, tile_rows=TILE_ROWS):
        self.H, self.W, self.world, self.rank, self.device = H, W, world, rank, device
        if world > 1:
            n_tiles = (H + tile_rows - 1) // tile_rows
            if H % tile_rows or n_tiles % world:
                raise ValueError(f"H={H} must split into a multiple of {world} tiles of {tile_rows} rows")
            self.ids = local_pixel_ids(H, W, world, rank, tile_rows).to(device)
            # position of every pixel inside the rank-major concatenation the all-gather produces
            order = torch.cat([local_pixel_ids(H, W, world, r, tile_rows) for r in range(world)])
            inv = torch.empty_like(order)
            inv[order] = torch.arange(order.numel())
            self.unpermute = inv.to(device)
        else:
            self.ids = None
        self.n_local = H * W // world
        self.peer = None   # (buffers, handles, ids32) once enable_peer_gather() succeeded
        self.parity = 0
        self.root = 0
        self.ctrl = None   # (control block, handle, tickets, per-buffer frame counters, scatter CTAs) for gather_to_root()

    def enable_peer_gather(self, group=None, n_buffers=2):
        """switch gather() to direct peer stores.  Returns False (and keeps NCCL) if symmetric memory is unavailable.
        n_buffers frame buffers are rotated (or addressed by `slot`): one per frame that can be in flight."""
        if self.world == 1:
            return False
        if (self.W * 3) % 4:
            return False
        try:
            import torch.distributed._symmetric_memory as symm
            bufs, hdls = [], []
            for _ in range(max(2, n_buffers)):
                t = symm.empty(self.H * self.W, 3, dtype=torch.float32, device=self.device)
                hdls.append(symm.rendezvous(t, group if group is not None else dist.group.WORLD))
                bufs.append(t)
            self.peer = (bufs, hdls, self.ids.to(torch.int32).contiguous())
            # control block of the flag-based gather-to-root: word 2k = arrivals at the root for buffer k, word 2k+1 = frames of buffer k
            # the root has consumed (lives on every rank, only the root writes it)
            from . import abi
            ctrl = symm.empty(2 * len(bufs) + 2, dtype=torch.int64, device=self.device)
            ctrl.zero_()
            ctrl_hdl = symm.rendezvous(ctrl, group if group is not None else dist.group.WORLD)
            torch.cuda.synchronize(self.device)
            ctrl_hdl.barrier(channel=0)      # every control block is zero before anybody signals
            torch.cuda.synchronize(self.device)
            ctas = int(abi.lib().rn_scatter_signal_ctas(self.n_local, self.W))
            if ctas > 0:
                self.ctrl = (ctrl, ctrl_hdl, torch.zeros(len(bufs), dtype=torch.int32, device=self.device), [0] * len(bufs), ctas)
            return True
        except Exception as e:  # noqa: BLE001 -- any failure of the optional path leaves the NCCL path in place
            self.peer = None
            self.peer_error = repr(e)
            return False

    def shard(self, t):
        """[H*W, c] -> this rank's rows [n_local, c] (contiguous)"""
        return t if self.world == 1 else t.index_select(0, self.ids).contiguous()

    def gather(self, local, slot=None):
        """[n_local, c] per rank -> [H*W, c] on every rank.  slot: which peer frame buffer to use (frame lanes); default rotates."""
        if self.world == 1:
            return local
        local = local.contiguous()
        if self.peer is not None and local.dtype == torch.float32 and local.dim() == 2 and local.shape[1] == 3:
            from . import abi
            bufs, hdls, ids32 = self.peer
            if slot is None:
                k, self.parity = self.parity, (self.parity + 1) % len(bufs)
            else:
                k = slot % len(bufs)
            # buffer k may still be read on another rank (its staging copy of the previous frame that went through this buffer is
            # ordered on ITS stream, not on ours): nobody stores into it before every rank has reached this point of its stream
            hdls[k].barrier(channel=1)
            abi.check(abi.lib().rn_scatter_rows_to_peers(abi.ptr(local), abi.ptr(ids32), local.shape[0], self.W, hdls[k].buffer_ptrs_dev,
                                                         self.world, abi.cur_stream()))
            hdls[k].barrier(channel=0)   # every rank's rows have landed in every rank's buffer k
            return bufs[k]               # valid until this buffer's next gather
        out = torch.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local)
        return out.index_select(0, self.unpermute)

    def next_seq(self, k):
        """frame counter of buffer k (1, 2, ...): every rank calls this once per frame that goes through the buffer, in the same order"""
        self.ctrl[3][k] += 1
        return self.ctrl[3][k]

    def gather_to_root(self, local, slot=None, stage_to=None, to_uint8=False):
        """[n_local, 3] fp32 per rank -> the root's frame buffer [H*W, 3] (returned; its content is valid ON THE ROOT ONLY, after this
        call in stream order, until the buffer's next use).  stage_to: on the root, additionally copy the assembled frame there (fp32, or
        uint8 when to_uint8) inside the same kernel that waits for the other ranks.  Falls back to gather() without the control block."""
        if self.world == 1:
            return local
        if self.peer is None or self.ctrl is None:
            return self.gather(local, slot)
        from . import abi
        local = local.contiguous()
        bufs, hdls, ids32 = self.peer
        ctrl, ctrl_hdl, tickets, _, ctas = self.ctrl
        if slot is None:
            k, self.parity = self.parity, (self.parity + 1) % len(bufs)
        else:
            k = slot % len(bufs)
        seq = self.next_seq(k)
        L = abi.lib()
        abi.check(L.rn_scatter_rows_to_root(abi.ptr(local), abi.ptr(ids32), local.shape[0], self.W, hdls[k].buffer_ptrs_dev, ctrl_hdl.buffer_ptrs_dev,
                                            self.world, self.rank, self.root, k, seq, abi.cur_stream()), "rn_scatter_rows_to_root")
        if self.rank == self.root:
            abi.check(L.rn_stage_frame_at_root(abi.ptr(bufs[k]), abi.ptr(stage_to), bufs[k].numel(), 1 if to_uint8 else 0, ctrl_hdl.buffer_ptrs_dev,
                                               self.world, self.root, k, seq, ctas, tickets[k:k + 1].data_ptr(), abi.cur_stream()),
                      "rn_stage_frame_at_root")
        return bufs[k]
