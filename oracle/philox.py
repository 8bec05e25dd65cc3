"""
ORACLE (test infrastructure, NOT product code) -- Philox4x32-10 counter-based RNG (Salmon et al., SC'11;
Random123), vectorised in NumPy.  The reference uses Python's global Mersenne Twister
(OML/optimize_SA.py:76,114); SURVEY 9.5#8 replaces it by step-indexed draws so any generator is admissible.
Production draws: key = (seed_lo, seed_hi), counter = (step_lo, step_hi, chain_lo, chain_hi);
  proposal site = mulhi(out[0], N),  uniform = (out[1] >> 8) * 2**-24   (exactly representable in fp32).
Known-answer vectors (Random123 kat_vectors) are checked in tests/test_oracle_philox.py.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: 4 arrays (uint32-valued), key: 2 arrays.  Returns 4 uint32 arrays."""
    c = [np.asarray(v, dtype=np.uint64) & MASK for v in ctr]
    k0 = np.asarray(key[0], dtype=np.uint64) & MASK
    k1 = np.asarray(key[1], dtype=np.uint64) & MASK
    for r in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [(hi1 ^ c[1] ^ k0) & MASK, lo1, (hi0 ^ c[3] ^ k1) & MASK, lo0]
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return [v.astype(np.uint32) for v in c]


def draws(seed, chains, steps, n_sites):
    """(proposal int32, uniform float32) for broadcastable arrays of global chain ids and step indices."""
    chains = np.asarray(chains, dtype=np.uint64)
    steps = np.asarray(steps, dtype=np.uint64)
    chains, steps = np.broadcast_arrays(chains, steps)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    out = philox4x32_10(
        (steps & MASK, steps >> np.uint64(32), chains & MASK, chains >> np.uint64(32)),
        (np.uint64(seed & 0xFFFFFFFF), np.uint64(seed >> 32)))
    prop = ((out[0].astype(np.uint64) * np.uint64(n_sites)) >> np.uint64(32)).astype(np.int32)
    u = ((out[1] >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
    return prop, u
