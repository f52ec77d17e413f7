"""Algorithmic (compulsory) HBM bytes per sample for each kernel of the path,
exactly the formulas of SURVEY.md §8(d) / BASELINE.md §3: every input read once,
every output written once; a "map" is one 64x64 float32 plane = 16 384 B."""
MAP = 64 * 64 * 4


def sfr_build_bytes(J):
    """train mode: the 128x128 float32 crop's worth of depth in; img + label +
    mask + J heat maps + J depth maps + uvd (12 J) + (box, cube, com) out."""
    return 65536 + (65536 + 2 * MAP) + 2 * J * MAP + 12 * J + 20


def sfr_crop_bytes():
    return 65536 + (65536 + 2 * MAP) + 20


def decoder_fwd_bytes(J, store_heat=True, with_targets=False):
    """read z, D (2J maps) + L, m (2 maps) + w; write H (J maps) + uvd."""
    b = 2 * J * MAP + 2 * MAP + 16 * J
    if store_heat:
        b += J * MAP
    if with_targets:
        b += 2 * J * MAP
    return b


def decoder_bwd_bytes(J, with_targets=True, upstream_maps=False):
    """read z, D (+ heat_gt, dmap_gt) (+ dense gH_up, gD_up) + L, m; write gz, gD."""
    maps = 4 * J + (2 * J if with_targets else 0) + (2 * J if upstream_maps else 0)
    return maps * MAP + 2 * MAP


TAPS = 64    # sizeof(pwr_joint_taps)


def sfr_build_sparse_bytes(J):
    """train mode with compact targets: no dense heat / depth maps, 64 B of taps per joint."""
    return 65536 + (65536 + 2 * MAP) + TAPS * J + 12 * J + 20


def decoder_bwd_sparse_bytes(J):
    """read z, D + L, m + 64 B of taps per joint; write gz, gD."""
    return 4 * J * MAP + 2 * MAP + TAPS * J


def decoder_fused_bytes(J, store_heat=True, sparse=False):
    """pwr_decoder_fwd_bwd_loss: read z, D (+ heat_gt, dmap_gt, or 64 B of taps per joint) + L, m once;
    write gz, gD (+ H) + uvd."""
    maps = 2 * J + (0 if sparse else 2 * J) + 2 * J + (J if store_heat else 0)
    return maps * MAP + 2 * MAP + (TAPS * J if sparse else 0) + 16 * J


def step_one_pass_bytes(J, sparse=False):
    """SFR build + one-pass last stage (the logits and targets are visited once)."""
    return (sfr_build_sparse_bytes(J) if sparse else sfr_build_bytes(J)) + decoder_fused_bytes(J, sparse=sparse)


def step_sparse_bytes(J):
    return sfr_build_sparse_bytes(J) + decoder_fwd_bytes(J) + decoder_bwd_sparse_bytes(J)


def step_bytes(J):
    """SFR build + decoder forward + last-stage backward+loss (BASELINE config 2)."""
    return sfr_build_bytes(J) + decoder_fwd_bytes(J) + decoder_bwd_bytes(J)


assert sfr_build_bytes(14) == 622780 and sfr_build_bytes(21) == 852240
assert decoder_fwd_bytes(14) == 721120 and decoder_fwd_bytes(21) == 1065296
assert decoder_bwd_bytes(14) == 1409024 and decoder_bwd_bytes(14, upstream_maps=True) == 1867776
assert step_bytes(14) == 2752924
assert decoder_fused_bytes(14) == 1638624 and step_one_pass_bytes(14) == 2261404
