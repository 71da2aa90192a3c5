"""The libsodium calls the tests' own gossip host (tests/host_sim.py) needs (signing, BLAKE2b ids, CSPRNG).
Test infrastructure: crypto is outside the accelerated path -- the consensus
kernels only ever consume signature *bytes*.  `pysodium` is used when it is installed (the reference's
dependency); otherwise the same primitives come from PyNaCl's bundled libsodium.
"""
from __future__ import annotations

import hashlib

try:  # the reference's own dependency, if present
    from pysodium import (crypto_generichash, crypto_sign, crypto_sign_detached,  # noqa: F401
                          crypto_sign_keypair, crypto_sign_open,
                          crypto_sign_verify_detached, randombytes)
except Exception:  # PyNaCl fallback (same libsodium underneath)
    from nacl import bindings as _b
    from nacl import exceptions as _e

    def crypto_sign_keypair():
        return _b.crypto_sign_keypair()

    def crypto_sign(m, sk):
        return _b.crypto_sign(m, sk)

    def crypto_sign_open(sm, pk):
        try:
            return _b.crypto_sign_open(sm, pk)
        except _e.CryptoError as exc:
            raise ValueError(str(exc))

    def crypto_sign_detached(m, sk):
        return _b.crypto_sign(m, sk)[:_b.crypto_sign_BYTES]

    def crypto_sign_verify_detached(sig, m, pk):
        try:
            _b.crypto_sign_open(sig + m, pk)
        except _e.CryptoError as exc:
            raise ValueError(str(exc))

    def crypto_generichash(m, k=b"", outlen=32):
        return hashlib.blake2b(m, digest_size=outlen, key=k).digest()

    def randombytes(n):
        return _b.randombytes(n)


def randrange(n: int) -> int:
    """Uniform integer in [0, n) by rejection sampling on CSPRNG bytes
    (same contract as /root/reference/utils.py:49-55)."""
    nbytes = (n.bit_length() + 7) // 8
    shift = 8 * nbytes - n.bit_length()
    while True:
        r = int.from_bytes(randombytes(nbytes), "big") >> shift
        if r < n:
            return r
