"""-m "not gpu": CRC-32C (SPEC X1/X2) against a witness nobody here wrote - the CPU's own `crc32` instruction - beside
the RFC 3720 known answers: the C oracle's table and bitwise forms, its combine, its 64 KiB chunk table and the
pure-Python restatement must all give what the hardware gives."""
import numpy as np
import pytest

from oracle import oracle as o, spec_py as sp
import hwcrc

SIZES = [0, 1, 2, 3, 7, 8, 9, 15, 16, 17, 63, 64, 65, 4095, 4096, 4097, 65535, 65536, 65537, 3 * 65536 + 11, 1_000_003, 8 << 20]


@pytest.fixture(scope="module")
def hw():
    if hwcrc.lib() is None:
        pytest.skip("no SSE4.2 crc32 instruction / gcc here")
    return hwcrc


def test_hardware_gives_the_rfc3720_answers(hw):
    assert hw.crc32c(b"123456789") == 0xE3069283
    assert hw.crc32c(bytes(32)) == 0x8A9136AA
    assert hw.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert hw.crc32c(bytes(range(32))) == 0x46DD794E


@pytest.mark.parametrize("n", SIZES)
def test_oracle_forms_equal_the_instruction(hw, n):
    rng = np.random.default_rng(1000 + n)
    d = rng.integers(0, 256, n, dtype=np.uint8)
    want = hw.crc32c(d)
    assert o.crc32c(d) == want
    if n <= 1_000_003:
        assert o.crc32c_bitwise(d) == want
    if n <= 65537:
        assert sp.crc32c(d.tobytes()) == want
    chunks, whole = o.crc32c_chunks(d)
    assert whole == want
    for i, c in enumerate(chunks):
        assert int(c) == hw.crc32c(d[i * 65536:(i + 1) * 65536])
    if n >= 2:                                                       # X2: combine(crc(A), crc(B), |B|) = crc(A || B)
        cut = n // 3
        assert o.crc32c_combine(hw.crc32c(d[:cut]), hw.crc32c(d[cut:]), n - cut) == want
    assert hw.crc32c(d[n // 2:], hw.crc32c(d[:n // 2])) == want      # the instruction's own running form
