"""Row-block sharding of a frame across the GPUs of one box — the host-side mirror of what
gvt_render_frame does internally (csrc/gvt_api.cu): equal blocks of ceil(H / world) rows (the all-gather needs equal
counts; the last block may be short or empty), one redundant halo row on each interior edge when TAA is on."""


def rows_per_rank(height, world):
    return (height + world - 1) // world


def shard_rows(height, rank, world):
    """-> (row_begin, row_end) owned by `rank`."""
    rpr = rows_per_rank(height, world)
    r0 = min(height, rank * rpr)
    return r0, min(height, r0 + rpr)


def traced_rows(height, rank, world, taa):
    """Rows the trace kernel covers on `rank`: its block plus the TAA halo."""
    r0, r1 = shard_rows(height, rank, world)
    if taa and r1 > r0:
        return (r0 - 1 if r0 > 0 else r0), (r1 + 1 if r1 < height else r1)
    return r0, r1


def gather_counts(width, height, world):
    """floats each rank contributes to the single ncclAllGather, and the padded frame height it implies."""
    rpr = rows_per_rank(height, world)
    return rpr * width * 4, rpr * world


def interleaved_rows(height, rank, world, taa=False):
    """GVT_FLAG_ROW_INTERLEAVE (peer-store gather only): stripes dealt round-robin, stripe j to rank j % world. Without
    TAA a stripe is one row (rows k, k + world, ...); with TAA it is 16 rows (each traced with a one-row halo)."""
    s = 16 if taa else 1
    return [y for y in range(height) if (y // s) % world == rank]
