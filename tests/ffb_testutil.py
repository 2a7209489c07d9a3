"""Shared helpers for the test-suite (kept out of conftest so they can be imported by name)."""


def norm_reads(n, length, seed):
    """Synthetic squiggles pushed through the host signal prep, as calculate_post would."""
    from flappie_b200.model import synthetic_reads
    from flappie_b200.signal import prepare_read
    out = []
    for r in synthetic_reads(n, length, seed=seed):
        x = prepare_read(r)
        assert x is not None
        out.append(x)
    return out
