"""Known-answer vectors of Philox4x32-10 (Salmon et al., SC'11; the Random123 distribution's kat_vectors file, entries
`philox4x32 10`): the counter-based generator that replaces the reference's random_device-seeded mt19937
(/root/reference/include/mccompletepathv2.h:32-34). The CPU oracle and the device function must both reproduce them --
so the generator is pinned to its published definition, not only to its twin."""
import ctypes as C

import numpy as np
import pytest

import oracle_bindings as ob

# (counter[4], key[2]) -> output[4]
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_oracle_philox_matches_the_random123_known_answers():
    lib = ob.oracle()
    for ctr, key, want in KAT:
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        out = (C.c_uint32 * 4)()
        lib.oracle_philox4x32_10(c, k, out)
        assert tuple(out) == want, (ctr, key, [hex(x) for x in out])


@pytest.mark.gpu
def test_device_philox_matches_the_random123_known_answers_and_the_oracle():
    from approximated_personalized_pagerank_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(7)
    extra = rng.integers(0, 2 ** 32, size=(500, 6), dtype=np.uint64).astype(np.uint32)
    blocks = np.concatenate([np.array([list(c) + list(k) for c, k, _ in KAT], dtype=np.uint32), extra])
    out = np.zeros((len(blocks), 4), dtype=np.uint32)
    _lib.check(lib.pprb200_debug_philox(_lib.ptr(np.ascontiguousarray(blocks)), _lib.ptr(out), len(blocks)))
    for i, (_, _, want) in enumerate(KAT):
        assert tuple(int(x) for x in out[i]) == want
    olib = ob.oracle()
    for i in range(len(KAT), len(blocks)):
        c = (C.c_uint32 * 4)(*[int(x) for x in blocks[i, :4]])
        k = (C.c_uint32 * 2)(*[int(x) for x in blocks[i, 4:]])
        o = (C.c_uint32 * 4)()
        olib.oracle_philox4x32_10(c, k, o)
        assert tuple(o) == tuple(int(x) for x in out[i])
