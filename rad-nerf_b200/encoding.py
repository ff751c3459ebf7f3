"""`get_encoder` -- the encoder factory nerf/network.py builds its five encoders with (reference: encoding.py).

Same call contract: get_encoder(name, input_dim=3, multires=6, degree=4, num_levels=16, level_dim=2, base_resolution=16,
log2_hashmap_size=19, desired_resolution=2048, align_corners=False, **kw) -> (encoder, encoder.output_dim); 'None' yields an
identity callable."""


def _frequency(a):
    from freqencoder import FreqEncoder
    return FreqEncoder(input_dim=a["input_dim"], degree=a["multires"])


def _spherical(a):
    from shencoder import SHEncoder
    return SHEncoder(input_dim=a["input_dim"], degree=a["degree"])


def _grid(kind):
    def make(a):
        from gridencoder import GridEncoder
        return GridEncoder(input_dim=a["input_dim"], num_levels=a["num_levels"], level_dim=a["level_dim"],
                           base_resolution=a["base_resolution"], log2_hashmap_size=a["log2_hashmap_size"],
                           desired_resolution=a["desired_resolution"], gridtype=kind, align_corners=a["align_corners"])
    return make


def _ash(a):
    # the reference forwards this name to a package that is not part of its tree; keep the same ImportError
    from ashencoder import AshEncoder
    return AshEncoder(input_dim=a["input_dim"], output_dim=16, log2_hashmap_size=a["log2_hashmap_size"], resolution=a["desired_resolution"])


_FACTORIES = {"frequency": _frequency, "spherical_harmonics": _spherical, "hashgrid": _grid("hash"), "tiledgrid": _grid("tiled"),
              "ash": _ash}


def get_encoder(encoding, input_dim=3, multires=6, degree=4, num_levels=16, level_dim=2, base_resolution=16,
                log2_hashmap_size=19, desired_resolution=2048, align_corners=False, **kwargs):
    if encoding == 'None':
        return (lambda x, **kw: x), input_dim
    if encoding not in _FACTORIES:
        raise NotImplementedError('Unknown encoding mode, choose from [None, frequency, spherical_harmonics, hashgrid, tiledgrid]')
    enc = _FACTORIES[encoding](dict(input_dim=input_dim, multires=multires, degree=degree, num_levels=num_levels, level_dim=level_dim,
                                    base_resolution=base_resolution, log2_hashmap_size=log2_hashmap_size,
                                    desired_resolution=desired_resolution, align_corners=align_corners))
    return enc, enc.output_dim
