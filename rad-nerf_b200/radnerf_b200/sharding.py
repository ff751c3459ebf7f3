"""Multi-GPU inference: each frame's rays are split by interleaved row tiles, every rank renders its tiles with replicated
weights / hash tables / occupancy bitfield, and one all-gather over NCCL (NVLink) reassembles the image.

Row tile t (TILE_ROWS image rows) goes to rank t mod world: the head sits in the middle rows, so interleaving balances the
number of occupied samples per rank.  There is no collective on the data path other than the final image all-gather;
`encode_audio` and the lip-smoothing state are recomputed identically on every rank (deterministic, ~1 MFLOP).

Two implementations of the gather:
  * NCCL `all_gather_into_tensor` + an un-permute gather kernel (default, works everywhere);
  * `enable_peer_gather()`: every rank stores its finished rows directly at their final position in every rank's frame
    buffer (torch symmetric memory = the same allocation mapped into all processes over NVLink/NVSwitch;
    csrc/peer_gather.cu) followed by a symmetric-memory barrier.  One small kernel + one barrier instead of a collective
    launch + a 3 MB permutation; frames alternate between two buffers so a fast rank can never overwrite a frame a slow
    rank is still reading.
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
            abi.check(abi.lib().rn_scatter_rows_to_peers(abi.ptr(local), abi.ptr(ids32), local.shape[0], self.W, hdls[k].buffer_ptrs_dev,
                                                         self.world, abi.cur_stream()))
            hdls[k].barrier(channel=0)   # every rank's rows have landed in every rank's buffer k
            return bufs[k]               # valid until this buffer's next gather
        out = torch.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local)
        return out.index_select(0, self.unpermute)
