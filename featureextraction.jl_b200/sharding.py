"""Host-side shard arithmetic shared by the binding, the tests and bench.py.
Mirrors the layout rule in csrc/api_kpca.cu: K is split into contiguous row
blocks of ``ld = roundup(ceil(n / world), 128)`` rows; only the last non-empty
rank is ragged; padding sits at the very end of the global index space."""


def round_up(a, b):
    return (a + b - 1) // b * b


def row_partition(n, world):
    """-> (ld, n_ld, [(row0, rows_valid) per rank])."""
    if n < 1 or world < 1:
        raise ValueError("n and world must be positive")
    ld = round_up((n + world - 1) // world, 128)
    parts = []
    for r in range(world):
        row0 = r * ld
        parts.append((row0, max(0, min(ld, n - row0))))
    return ld, ld * world, parts


def query_partition(m, world):
    """Contiguous split of m query rows across ranks for transform(): [(q0, count)]."""
    base, rem = divmod(m, world)
    out, q0 = [], 0
    for r in range(world):
        cnt = base + (1 if r < rem else 0)
        out.append((q0, cnt))
        q0 += cnt
    return out


def gram_tiles_executed(n, world, rank=0, peer=False):
    """-> (executed, total): 128x64 output tiles of this rank's row block that the symmetric Gram
    kernel actually computes (the rest arrive by mirror stores).  Restates tile_role() in
    csrc/gram.cu: inside the rank's diagonal block only tiles on / right of the diagonal; with
    ``peer`` (NVLink peer mappings) the off-diagonal blocks are shared out between the two ranks of
    each pair as well (ring distance < world/2 owns; opposite blocks split in a checkerboard of
    128x128 super-tiles)."""
    import numpy as np
    ld = row_partition(n, world)[0]
    tiles_y, bw = ld // 128, 2 * (ld // 128)
    ty = np.arange(tiles_y)[:, None]
    executed = 0
    for c in range(world):
        rel = np.arange(bw)[None, :]
        if c == rank:
            executed += int((rel >= 2 * ty).sum())
        elif not peer:
            executed += tiles_y * bw
        else:
            delta = (c - rank) % world
            if 2 * delta < world:
                executed += tiles_y * bw
            elif 2 * delta == world:
                executed += int((((ty + rel // 2) & 1) == (0 if rank < c else 1)).sum())
    return executed, tiles_y * bw * world
