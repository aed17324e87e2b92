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
