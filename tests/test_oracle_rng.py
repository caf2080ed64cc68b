"""Oracle of the device random numbers against the published known answers of Philox4x32-10 (Random123 kat_vectors),
and basic statistics of the Gaussian mapping."""
import numpy as np

from oracle import ref_rng


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = ref_rng.philox4x32_10(*[np.array([c]) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want


def test_counter_layout_and_statistics():
    w = ref_rng.words(seed=(7 << 32) | 5, stream=(3 << 32) | 9, ncalls=4)
    one = ref_rng.philox4x32_10(np.array([2]), np.array([0]), np.array([9]), np.array([3]), 5, 7)
    assert [int(x[0]) for x in one] == [int(v) for v in w[2]]
    z = ref_rng.randn(1234, 1, 200001)
    assert z.size == 200001 and abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs(np.mean(z ** 4) - 3.0) < 0.1
    a = ref_rng.randn_alm(99, 0, 300)
    assert np.all(a[:301].imag == 0) and abs(np.var(a[:301].real) - 1) < 0.2
    assert abs(np.mean(np.abs(a[301:]) ** 2) - 1) < 0.02
    assert not np.array_equal(ref_rng.randn(1234, 2, 10), z[:10])
