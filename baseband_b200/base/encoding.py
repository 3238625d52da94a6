"""Reference names of baseband/base/encoding.py, backed by the CUDA codec.

Constants and ``decoder_levels`` (:14, :45-56) are the host-generated tables
of `baseband_b200.levels`.  ``encode_{1,2,4}bit_base`` (:63-128) return one
uint8 code per value -- the quantisation step on its own: the values go
through the GPU quantise-and-pack kernel and the packed codes are unpacked
again by the decode kernel with the identity table, so the thresholds,
clipping and rounding are exactly those of the packed encoders.
``decode_8bit`` / ``encode_8bit`` (:131-158) are the 8-bit offset-binary
codec.  There is no CPU arithmetic here.
"""
import numpy as np

from .. import codecs, kernels
from ..levels import (OPTIMAL_2BIT_HIGH, TWO_BIT_1_SIGMA,  # noqa: F401
                      FOUR_BIT_1_SIGMA, EIGHT_BIT_1_SIGMA, decoder_levels)

__all__ = ['OPTIMAL_2BIT_HIGH', 'TWO_BIT_1_SIGMA', 'FOUR_BIT_1_SIGMA',
           'EIGHT_BIT_1_SIGMA', 'decoder_levels', 'encode_1bit_base',
           'encode_2bit_base', 'encode_4bit_base', 'decode_8bit',
           'encode_8bit']


def _codes(values, bps):
    values = np.asarray(values)
    shape = values.shape
    flat = values.reshape(-1)
    per_word = 32 // bps
    pad = (-flat.size) % per_word
    if pad:
        flat = np.concatenate([flat, np.zeros(pad, flat.dtype)])
    packed = codecs.encode_flat(flat, bps, kernels.QUANT_OFFSET_BINARY)
    identity = np.arange(1 << bps, dtype=np.float32)
    codes = codecs.decode_flat(packed, bps, identity)
    return codes[:values.size].astype(np.uint8).reshape(shape)


def encode_1bit_base(values):
    """Sign of each value as 0 / 1 (base/encoding.py:63-74)."""
    return _codes(values, 1)


def encode_2bit_base(values):
    """Two-bit code 0..3 of each value (base/encoding.py:77-102)."""
    return _codes(values, 2)


def encode_4bit_base(values):
    """Four-bit code 0..15 of each value (base/encoding.py:105-128)."""
    return _codes(values, 4)


decode_8bit = codecs.VDIF_DECODERS[8]
encode_8bit = codecs.VDIF_ENCODERS[8]
