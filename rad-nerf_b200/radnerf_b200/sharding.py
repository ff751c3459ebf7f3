"""Multi-GPU inference: each frame's rays are split by interleaved row tiles, every rank renders its tiles with replicated
weights / hash tables / occupancy bitfield, and one all-gather over NCCL (NVLink) reassembles the image.

Row tile t (TILE_ROWS image rows) goes to rank t mod world: the head sits in the middle rows, so interleaving balances the
number of occupied samples per rank.  There is no collective on the data path other than the final image all-gather;
`encode_audio` and the lip-smoothing state are recomputed identically on every rank (deterministic, ~1 MFLOP).
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

    def shard(self, t):
        """[H*W, c] -> this rank's rows [n_local, c] (contiguous)"""
        return t if self.world == 1 else t.index_select(0, self.ids).contiguous()

    def gather(self, local):
        """[n_local, c] per rank -> [H*W, c] on every rank"""
        if self.world == 1:
            return local
        local = local.contiguous()
        out = torch.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local)
        return out.index_select(0, self.unpermute)
