"""The random stream: Philox4x32-10 known-answer vectors (Random123's kat_vectors,
Salmon et al. SC'11) and the uniform / normal maps of oracle/philox_ref.h."""
import ctypes as C
import math

import numpy as np
import pytest

KAT = [  # counter, key, expected
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox(lib, ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib.mo_philox(c, k, o)
    return tuple(o)


def philox_py(ctr, key, rounds):
    """Philox4x32-R straight from the paper (Salmon et al. SC'11, fig. 2 + the Weyl key schedule)."""
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(rounds):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = (p1 >> 32) ^ c1 ^ k0, p1 & 0xffffffff, (p0 >> 32) ^ c3 ^ k1, p0 & 0xffffffff
        k0, k1 = (k0 + 0x9E3779B9) & 0xffffffff, (k1 + 0xBB67AE85) & 0xffffffff
    return (c0, c1, c2, c3)


@pytest.fixture
def stream(port, request):
    """Select the oracle's stream version for one test, restore afterwards."""
    was = port.stream
    request.addfinalizer(lambda: port.set_stream(was))
    return port.set_stream


def test_philox_known_answers(port, stream):
    lib = port.lib._lib
    stream(1)                                   # stream v1 = Philox4x32-10: Random123's vectors
    for ctr, key, want in KAT:
        assert philox(lib, ctr, key) == want == philox_py(ctr, key, 10)
    stream(2)                                   # stream v2 = Philox4x32-7
    rng = np.random.default_rng(0)
    for ctr, key, _ in KAT + [(tuple(int(x) for x in rng.integers(0, 2 ** 32, 4)),
                               tuple(int(x) for x in rng.integers(0, 2 ** 32, 2)), None) for _ in range(50)]:
        assert philox(lib, ctr, key) == philox_py(ctr, key, 7) != philox_py(ctr, key, 10)


def test_uniform_and_normal_maps(port):
    lib = port.lib._lib
    lib.mo_uniform_at.restype = C.c_double
    lib.mo_normal_at.restype = C.c_double
    seed, gene, chain = 0x0123456789abcdef, 17, 3
    key = (seed & 0xffffffff, seed >> 32)
    for n in (0, 1, 2, 3, 4, 5, 1000003):
        w = philox(lib, (n >> 2, 0, gene, chain), key)[n & 3]
        u = lib.mo_uniform_at(C.c_uint64(seed), gene, chain, C.c_uint64(n))
        assert u == (w + 0.5) * 2.0 ** -32 and 0.0 < u < 1.0
    for n in (0, 1, 77):
        x = philox(lib, (n, 1, gene, chain), key)
        a = (x[0] << 21) | (x[1] >> 11)
        b = (x[2] << 21) | (x[3] >> 11)
        want = math.sqrt(-2.0 * math.log((a + 1) * 2.0 ** -53)) * math.cos(6.283185307179586 * (b * 2.0 ** -53))
        got = lib.mo_normal_at(C.c_uint64(seed), gene, chain, C.c_uint64(n))
        assert abs(got - want) <= 4e-16 * max(1.0, abs(want))
    z = np.array([lib.mo_normal_at(C.c_uint64(1), 0, 0, C.c_uint64(i)) for i in range(20000)])
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1.0) < 0.03
