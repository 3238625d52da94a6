"""Minimal 80-column ``KEY = value`` card headers (GUPPI raw files).

GUPPI headers look like FITS headers (fixed-format cards, ``END`` card) but
are not padded to 2880 bytes.  The reference parses them with
``astropy.io.fits.Header`` (baseband/guppi/header.py:17-186); here a small
ordered mapping does the same job without astropy.  Cards read from a file
keep their original text, so an unmodified header is written back byte for
byte; new or changed cards are formatted in FITS fixed format (strings
quoted, left-justified and padded to 8 characters; numbers right-justified to
column 30).
"""
__all__ = ['CardHeader']


def parse_value(text):
    text = text.split('/')[0].strip() if not text.lstrip().startswith("'") \
        else text.strip()
    if text.startswith("'"):
        end = text.find("'", 1)
        while end != -1 and text[end:end + 2] == "''":
            end = text.find("'", end + 2)
        return text[1:end if end != -1 else None].replace("''", "'").rstrip()
    if text in ('T', 'F'):
        return text == 'T'
    try:
        return int(text)
    except ValueError:
        pass
    try:
        return float(text.replace('D', 'E'))
    except ValueError:
        return text


def format_card(key, value):
    if isinstance(value, bool):
        body = '{:>20s}'.format('T' if value else 'F')
    elif isinstance(value, str):
        body = "'{:<8s}'".format(value.replace("'", "''"))
    elif isinstance(value, float):
        text = repr(float(value))
        if 'e' in text:
            text = text.upper()
        elif '.' not in text and 'inf' not in text and 'nan' not in text:
            text += '.0'
        body = '{:>20s}'.format(text)
    else:
        body = '{:>20d}'.format(int(value))
    return '{:<8s}= {}'.format(key, body).ljust(80)[:80]


class CardHeader:
    """Ordered, case-insensitive mapping of header cards."""

    def __init__(self, cards=None):
        self._values = {}
        self._text = {}
        if cards:
            for key, value in (cards.items() if hasattr(cards, 'items')
                               else cards):
                self._set(key, value)

    # ------------------------------------------------------------ mapping
    def _set(self, key, value, text=None):
        key = key.upper()
        if hasattr(value, 'item') and getattr(value, 'ndim', 1) == 0:
            value = value.item()
        self._values[key] = value
        if text is None:
            self._text.pop(key, None)
        else:
            self._text[key] = text

    def __getitem__(self, key):
        return self._values[key.upper()]

    def __setitem__(self, key, value):
        self._set(key, value)

    def __delitem__(self, key):
        del self._values[key.upper()]
        self._text.pop(key.upper(), None)

    def __contains__(self, key):
        return isinstance(key, str) and key.upper() in self._values

    def __len__(self):
        return len(self._values)

    def __iter__(self):
        return iter(self._values)

    def keys(self):
        return self._values.keys()

    def items(self):
        return self._values.items()

    def get(self, key, default=None):
        return self._values.get(key.upper(), default)

    # ---------------------------------------------------------------- text
    @classmethod
    def parse(cls, text):
        self = cls()
        for i in range(0, len(text), 80):
            card = text[i:i + 80]
            if card[:3] == 'END' and card[3:].strip() == '':
                break
            key = card[:8].strip()
            if not key or card[8:10] != '= ':
                continue
            self._set(key, parse_value(card[10:]), card.ljust(80))
        return self

    def tostring(self):
        cards = [self._text.get(k) or format_card(k, v)
                 for k, v in self._values.items()]
        return ''.join(cards) + 'END'.ljust(80)
