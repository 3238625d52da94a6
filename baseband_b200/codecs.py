"""Operator tables: the callables behind every ``Payload._decoders`` /
``_encoders`` entry, dispatching to the CUDA library.

Signatures follow the reference's tables (baseband/base/payload.py:46-50,
:314-325): ``decode(words: ndarray) -> ndarray[float32]`` and
``encode(values: ndarray[float]) -> ndarray[uint8 / int8 / word]``, so a
payload class can hold them in a dict keyed by ``bps`` (VDIF, Mark 5B, GUPPI,
DADA, GSB) or by ``(nchan, bps-or-magbits, fanout)`` (Mark 4).  Each call
uploads its input, runs one kernel and downloads the result; the batched
stream readers bypass these and keep everything on the device.

``*_device`` variants take and return CUDA tensors.
"""
import numpy as np
import torch

from . import device as _device
from . import kernels, levels

__all__ = ['decode_flat', 'encode_flat', 'make_decoder', 'make_encoder',
           'VDIF_DECODERS', 'VDIF_ENCODERS', 'MARK5B_DECODERS',
           'MARK5B_ENCODERS', 'INT8_DECODERS', 'INT8_ENCODERS',
           'GSB_DECODERS', 'GSB_ENCODERS', 'MARK4_DECODERS',
           'MARK4_ENCODERS']

_zero_offset = {}


def _offset0(dev):
    key = str(dev)
    if key not in _zero_offset:
        _zero_offset[key] = kernels.zeros(1, torch.int64, dev)
    return _zero_offset[key]


def decode_flat_device(raw, bps, levels_table, codec):
    """uint8 CUDA tensor (multiple of 4 bytes) -> flat float32 CUDA tensor."""
    nbytes = raw.numel()
    out = kernels.decode_bitfield(raw, _offset0(raw.device), 1, 1, nbytes,
                                  bps, 1, False, codec, levels_table)
    return out.reshape(-1)


def decode_flat(words, bps, levels_table=None, codec=kernels.CODEC_LEVELS,
                dev=None):
    """Decode every ``bps``-bit code of ``words`` (any dtype), LSB first."""
    dev = _device.resolve(dev)
    arr = np.ascontiguousarray(words).view(np.uint8).reshape(-1)
    nbytes = arr.size
    if nbytes == 0:
        return np.empty(0, np.float32)
    pad = (-nbytes) % 4
    if pad:
        arr = np.concatenate([arr, np.zeros(pad, np.uint8)])
    raw = _device.upload(arr, dev)
    out = decode_flat_device(raw, bps, levels_table, codec)
    return _device.download(out)[:nbytes * 8 // bps]


def encode_flat_device(values, bps, quantiser):
    """float32/float64 CUDA tensor whose size*bps is a multiple of 32 ->
    packed uint8 CUDA tensor."""
    flat = values.reshape(-1)
    nbytes = flat.numel() * bps // 8
    dst = torch.empty(nbytes, dtype=torch.uint8, device=values.device)
    kernels.encode_bitfield(flat, dst, _offset0(values.device), 1, 1, nbytes,
                            bps, 1, quantiser)
    return dst


def encode_flat(values, bps, quantiser, dev=None):
    """Quantise and pack a float array; arithmetic in the input's width
    (float32 stays float32, everything else is done in float64, as numpy
    would)."""
    dev = _device.resolve(dev)
    values = np.asarray(values)
    if values.dtype != np.float32:
        values = values.astype(np.float64, copy=False)
    flat = np.ascontiguousarray(values).reshape(-1)
    n = flat.size
    if n * bps % 8:
        raise ValueError('number of values does not fill whole bytes')
    if n == 0:
        return np.empty(0, np.uint8)
    per_word = 32 // bps
    pad = (-n) % per_word
    if pad:
        flat = np.concatenate([flat, np.zeros(pad, flat.dtype)])
    t = _device.upload(flat, dev).view(
        torch.float32 if flat.dtype == np.float32 else torch.float64)
    out = encode_flat_device(t, bps, quantiser)
    return _device.download(out)[:n * bps // 8]


def make_decoder(bps, levels_table, codec=kernels.CODEC_LEVELS, name=None):
    table = (None if levels_table is None
             else np.ascontiguousarray(levels_table, np.float32))

    def decode(words):
        return decode_flat(words, bps, table, codec)
    decode.__name__ = name or 'decode_{}bit'.format(bps)
    decode.bps, decode.levels, decode.codec = bps, table, codec
    return decode


def make_encoder(bps, quantiser, out_dtype=np.uint8, name=None):
    def encode(values):
        return encode_flat(values, bps, quantiser).view(out_dtype)
    encode.__name__ = name or 'encode_{}bit'.format(bps)
    encode.bps, encode.quantiser = bps, quantiser
    return encode


# VDIF: baseband/vdif/payload.py:137-145
VDIF_DECODERS = {bps: make_decoder(bps, levels.offset_binary(bps))
                 for bps in (1, 2, 4, 8)}
VDIF_ENCODERS = {bps: make_encoder(bps, kernels.QUANT_OFFSET_BINARY)
                 for bps in (1, 2, 4, 8)}
# Mark 5B: baseband/mark5b/payload.py:127-130
MARK5B_DECODERS = {bps: make_decoder(bps, levels.mark5b(bps))
                   for bps in (1, 2)}
MARK5B_ENCODERS = {bps: make_encoder(bps, kernels.QUANT_MARK5B)
                   for bps in (1, 2)}
# GUPPI / DADA: baseband/guppi/payload.py:43-46, dada/payload.py:40-43
INT8_DECODERS = {8: make_decoder(8, None, kernels.CODEC_SINT)}
INT8_ENCODERS = {8: make_encoder(8, kernels.QUANT_SINT, np.int8)}
# GSB: baseband/gsb/payload.py:72-75
GSB_DECODERS = {4: make_decoder(4, None, kernels.CODEC_SINT),
                8: make_decoder(8, None, kernels.CODEC_SINT)}
GSB_ENCODERS = {4: make_encoder(4, kernels.QUANT_SINT, np.int8),
                8: make_encoder(8, kernels.QUANT_SINT, np.int8)}

# Mark 4: baseband/mark4/payload.py:333-342, keyed (nchan, bps | magbits,
# fanout); the Fortaleza key is the packed non-standard magnitude-bit mask.
M4_FT_MAGBITS = 0xf0faf050f0faf05
_M4_MODES = {(2, 2, 4): (2, 4, False), (4, 2, 4): (4, 4, False),
             (8, 2, 2): (8, 2, False), (8, 2, 4): (8, 4, False),
             (16, M4_FT_MAGBITS, 2): (16, 2, True)}
_M4_WORD = {16: '<u2', 32: '<u4', 64: '<u8'}


def _make_mark4(nchan, fanout, ft):
    wdtype = np.dtype(_M4_WORD[nchan * 2 * fanout])

    def decode(words):
        dev = _device.resolve(None)
        w = np.ascontiguousarray(words).view(wdtype).reshape(-1)
        if w.size == 0:
            return np.empty((0, nchan), np.float32)
        raw = _device.upload(w, dev)
        out = kernels.mark4_decode_words(raw, w.size, nchan, fanout, ft,
                                         levels.sign_magnitude())
        return _device.download(out)

    def encode(values):
        dev = _device.resolve(None)
        values = np.asarray(values)
        if values.dtype != np.float32:
            values = values.astype(np.float64, copy=False)
        flat = np.ascontiguousarray(values).reshape(-1, nchan)
        nword = flat.shape[0] // fanout
        if nword == 0:
            return np.empty(0, wdtype)
        t = _device.upload(flat, dev).view(
            torch.float32 if flat.dtype == np.float32 else torch.float64)
        words = torch.empty(nword * wdtype.itemsize, dtype=torch.uint8,
                            device=dev)
        kernels.mark4_encode_words(t, words, nword, nchan, fanout, ft)
        return _device.download(words).view(wdtype)

    tag = '{}chan_2bit_fanout{}{}'.format(nchan, fanout, '_ft' if ft else '')
    decode.__name__, encode.__name__ = 'decode_' + tag, 'encode_' + tag
    return decode, encode


MARK4_DECODERS, MARK4_ENCODERS = {}, {}
for _key, _mode in _M4_MODES.items():
    MARK4_DECODERS[_key], MARK4_ENCODERS[_key] = _make_mark4(*_mode)
