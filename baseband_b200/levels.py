"""Reconstruction levels, generated on the host with the reference formulas
so that the tables uploaded to the GPU are bit-identical by construction.

Follows baseband/base/encoding.py:14, :45-56 (levels), :131-144 (8 bit),
baseband/mark5b/payload.py:27-75 and baseband/mark4/payload.py:88-115
(sign/magnitude ordering).  Tables are indexed by the raw code as it sits in
the payload (LSB first).
"""
import numpy as np

OPTIMAL_2BIT_HIGH = 3.316505
TWO_BIT_1_SIGMA = 2.174564
FOUR_BIT_1_SIGMA = 2.95
EIGHT_BIT_1_SIGMA = 71.0 / 2.0

decoder_levels = {
    1: np.array([-1.0, 1.0], dtype=np.float32),
    2: np.array([-OPTIMAL_2BIT_HIGH, -1.0, 1.0, OPTIMAL_2BIT_HIGH],
                dtype=np.float32),
    4: (np.arange(16, dtype=np.float32) - 8.0) / FOUR_BIT_1_SIGMA,
}


def _eight_bit():
    lv = np.arange(256, dtype=np.uint8).astype(np.float32)
    lv -= 127.5
    lv /= EIGHT_BIT_1_SIGMA
    return lv


def offset_binary(bps):
    """VDIF: code 0 lowest ... all ones highest."""
    if bps == 8:
        return _eight_bit()
    return decoder_levels[bps].copy()


def mark5b(bps):
    """Mark 5B: 1 bit set = -1; 2 bit code = sign | magnitude << 1."""
    if bps == 1:
        return decoder_levels[1][::-1].copy()
    code = np.arange(4)
    return decoder_levels[2][2 * (code & 1) + (code >> 1)]


def sign_magnitude():
    """Mark 4: indexed by 2 * sign + magnitude."""
    return decoder_levels[2].copy()
