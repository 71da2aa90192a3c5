"""Test-infrastructure shim: the seven `pysodium` names the reference imports
(/root/reference/swirld.py:10-12, utils.py:5, viz.py:16), mapped onto PyNaCl's
bundled libsodium (pysodium itself is not installed in this image).  Put this
directory on sys.path *before* /root/reference to import the reference
unmodified.  Not product code; never imported by the engine."""
import hashlib

from nacl import bindings as _b
from nacl import exceptions as _e


def crypto_sign_keypair():
    return _b.crypto_sign_keypair()


def crypto_sign(m, sk):
    return _b.crypto_sign(m, sk)


def crypto_sign_open(sm, pk):
    try:
        return _b.crypto_sign_open(sm, pk)
    except _e.CryptoError as exc:          # swirld.py:100 expects ValueError
        raise ValueError(str(exc))


def crypto_sign_detached(m, sk):
    return _b.crypto_sign(m, sk)[:64]


def crypto_sign_verify_detached(sig, m, pk):
    try:
        _b.crypto_sign_open(sig + m, pk)
    except _e.CryptoError as exc:
        raise ValueError(str(exc))


def crypto_generichash(m, k=b'', outlen=32):
    return hashlib.blake2b(m, digest_size=outlen, key=k).digest()


def randombytes(n):
    return _b.randombytes(n)
