"""CPU restatement of the device-side generators (qgrad_b200/csrc/rand.cu).  TEST INFRASTRUCTURE (see
``oracle/__init__.py``).

The reference draws its synthetic data on the host: ``rand_ket`` / ``rand_dm`` / ``rand_unitary``
(qgrad/qgrad_qutip.py:558-620, ``jax.random`` threefry), QuTiP ``rand_ket`` and tenpy ``GUE`` in the examples
(examples/Unitary-Learning-qgrad.py:63,88; examples/Unitary-Learning-no-qgrad.py:106-107).  None of those streams can
be reproduced here (no JAX / QuTiP / tenpy) and the reference's tests pin only PROPERTIES of the draws, so the CUDA
generators use their own counter-based stream, Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as
easy as 1, 2, 3", SC'11 -- the Random123 library, not vendored anywhere in /root/reference).  This file restates that
published algorithm in NumPy and the layout rand.cu gives it; it is pinned against Random123's known-answer vectors
(``PHILOX4X32_10_KAT``) in tests/test_oracle_golden.py, and the GPU tests compare the kernels with it word for word.
"""
from __future__ import annotations

import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
# Random123 kat_vectors, philox4x32 10 rounds: (counter, key, expected)
PHILOX4X32_10_KAT = (
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF),
     (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
)
STREAM_KET, STREAM_HERM, STREAM_UNIFORM = 1, 2, 3
TAG = 0x51677261


def philox4x32_10(ctr, key):
    """ctr: 4 arrays of uint64 holding 32-bit words (vectorised over any shape), key: 2 Python ints.  Returns 4 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & np.uint64(MASK) for c in ctr]
    k0, k1 = int(key[0]) & MASK, int(key[1]) & MASK
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(MASK)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(MASK)
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def draw(seed, stream, index):
    """the four words rand.cu draws at (seed, stream, index); index: array of non-negative ints < 2^64"""
    index = np.asarray(index, dtype=np.uint64)
    z = np.zeros_like(index)
    return philox4x32_10((index & np.uint64(MASK), index >> np.uint64(32), z + np.uint64(stream), z + np.uint64(TAG)),
                         (int(seed) & MASK, (int(seed) >> 32) & MASK))


def unit_open(a, b, single):
    """(0, 1) uniform from one (float32: top 24 bits) or two (float64: 53 bits) words, as rand.cu"""
    if single:
        return ((a >> np.uint64(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    v = (a << np.uint64(21)) ^ (b >> np.uint64(11))
    return (v.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def normal_pair(w, single):
    u1, u2 = unit_open(w[0], w[1], single), unit_open(w[2], w[3], single)
    r = np.sqrt(-2.0 * np.log(u1.astype(np.float64)))
    ang = 2.0 * np.pi * u2.astype(np.float64)
    return r * np.cos(ang) + 1j * r * np.sin(ang)


def rand_ket(batch, n, seed, first_ket=0, kind=0, single=False):
    """(batch, n) kets first_ket .. first_ket + batch - 1 of sequence ``seed``: Haar (kind 0) or the reference's
    rand_ket structure (kind 1: re == im uniform, qgrad_qutip.py:570-573), normalised."""
    idx = (np.arange(batch, dtype=np.uint64)[:, None] + np.uint64(first_ket)) * np.uint64(n) + np.arange(n, dtype=np.uint64)[None, :]
    w = draw(seed, STREAM_KET, idx)
    if kind == 1:
        u = unit_open(w[0], w[1], single).astype(np.float64)
        k = u + 1j * u
    else:
        k = normal_pair(w, single)
    return k / np.linalg.norm(k, axis=1, keepdims=True)


def rand_hermitian(batch, n, seed, first=0, scale=1.0, single=False):
    """(batch, n, n): scale (X + X^H) / 2 with X iid complex standard normal drawn at (matrix, r, c)"""
    nn = n * n
    base = (np.arange(batch, dtype=np.uint64) + np.uint64(first))[:, None, None] * np.uint64(nn)
    r, c = np.arange(n, dtype=np.uint64)[None, :, None], np.arange(n, dtype=np.uint64)[None, None, :]
    X = normal_pair(draw(seed, STREAM_HERM, base + r * np.uint64(n) + c), single)
    H = 0.5 * scale * (X + np.conj(np.swapaxes(X, -1, -2)))
    d = np.arange(n)
    H[:, d, d] = scale * X[:, d, d].real
    return H


def rand_uniform(count, seed, lo, hi, single=False):
    w = draw(seed, STREAM_UNIFORM, np.arange(count, dtype=np.uint64))
    u = unit_open(w[0], w[1], single)
    return (u.astype(np.float64) * (hi - lo) + lo) if not single else (u * np.float32(hi - lo) + np.float32(lo))
