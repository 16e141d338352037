"""The Philox4x32-10 restatement the MD tests rely on, against the known-answer vectors published
with Random123 (kat_vectors): the device generator is held to this restatement bit for bit by
tests/test_gpu_md.py, so these vectors pin both."""
import numpy as np

from md_ref import normals3, philox4x32_10

KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
     [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


def test_philox_known_answers():
    for ctr, key, want in KAT:
        got = philox4x32_10(np.array([ctr], dtype=np.uint32), np.array(key, dtype=np.uint32))[0]
        assert [int(x) for x in got] == want


def test_normals_are_standard_and_streams_differ():
    n = normals3(seed=12345, n=200000, step=7, stream=1)
    assert n.shape == (200000, 3)
    assert abs(n.mean()) < 0.01 and abs(n.std() - 1.0) < 0.01
    assert abs(np.corrcoef(n[:, 0], n[:, 1])[0, 1]) < 0.01
    assert not np.array_equal(n, normals3(seed=12345, n=200000, step=8, stream=1))
    assert not np.array_equal(n, normals3(seed=12345, n=200000, step=7, stream=0))
    assert np.array_equal(n, normals3(seed=12345, n=200000, step=7, stream=1))
