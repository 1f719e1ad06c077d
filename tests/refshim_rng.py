"""Pure-python definition of the tie-break stream (no reference import needed)."""
M64 = (1 << 64) - 1


def rng_u32(seed: int, tree: int, ctr: int) -> int:
    z = (seed + 0x9E3779B97F4A7C15 * (((tree & 0xFFFFFFFF) << 32) | (ctr & 0xFFFFFFFF))) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return (z >> 32) & 0xFFFFFFFF
