"""Card-level value types of the reference API: ``Barva``, ``Karta``, ``Roka``, ``Tip_igre``.

Same names, constructor arguments and observable behaviour as the reference's ``Karta.py``,
``Roka.py`` and ``Tip_igre.py`` (citations per item), so player code written against the reference
keeps working.  Internally everything is table/bitboard driven: a card is its id 0..53 and a hand is
a 64-bit mask on the device; these objects exist for the callback protocol (``Igralec``) and tests.
"""
from __future__ import annotations

import enum
from typing import Iterable, List


class Barva(enum.IntEnum):
    """Suits (Karta.py:69-74)."""
    KARA = 0
    SRCE = 1
    PIK = 2
    KRIZ = 3
    TAROK = 4

    # The reference slices the default str() (Karta.py:76-77), which yields the bare member name on the
    # Python it targets (3.6) and '' on >= 3.11; the intended spelling is kept here.
    def __str__(self):
        return self.name

    __repr__ = __str__


class Tip_igre(enum.IntEnum):
    """Contracts; the integer is also the game value used in scoring (Tip_igre.py:4-15)."""
    Naprej = -10
    Klop = 0
    Tri = 10
    Dve = 20
    Ena = 30
    Solo_tri = 40
    Solo_dve = 50
    Solo_ena = 60
    Berac = 70
    Solo_brez = 80
    Odprti_berac = 90

    def __str__(self):
        return self.name

    __repr__ = __str__

    @property
    def code(self) -> int:
        """Device contract code (value / 10); Naprej is -1."""
        return int(self) // 10

    @staticmethod
    def iz_kode(code: int) -> "Tip_igre":
        return Tip_igre(int(code) * 10)


_SLIKE = {5: "J", 6: "K", 7: "D", 8: "KR"}
# buggy discardability value (Karta.py:10-16): trula 5, any rank > 4 -> rank-3, else 1
_VREDNOST = [5 if i in (32, 52, 53) else (((i - 31) if i > 31 else (i % 8 + 1)) - 3
             if ((i - 31) if i > 31 else (i % 8 + 1)) > 4 else 1) for i in range(54)]
# counting points (Roka.py:76-91): suit ranks 1-4 -> 1, J C Q K -> 2 3 4 5; taroks 1 except trula 5
_TOCKE = [(5 if i in (32, 52, 53) else 1) if i > 31 else (1, 1, 1, 1, 2, 3, 4, 5)[i % 8] for i in range(54)]


class Karta:
    """One card: ``Karta(barva, st)``; id = suit*8 + st-1, taroks 32 + st-1 (Karta.py:5-47)."""

    __slots__ = ("barva", "st")

    def __init__(self, barva, st):
        self.barva = barva
        self.st = st

    def v_id(self) -> int:
        base = 32 if self.barva == Barva.TAROK else int(self.barva) * 8
        return base + self.st - 1

    @staticmethod
    def iz_id(id) -> "Karta":
        id = int(id)
        if id > 31:
            return Karta(Barva.TAROK, id - 31)
        return Karta(Barva(id >> 3), (id & 7) + 1)

    def vrednost(self) -> int:
        return _VREDNOST[self.v_id()]

    def __eq__(self, other):
        return isinstance(other, Karta) and self.barva == other.barva and self.st == other.st

    __hash__ = None  # the reference defines __eq__ only, so cards are unhashable there too

    def __lt__(self, other):
        assert isinstance(other, Karta)
        return (self.barva, self.st) < (other.barva, other.st)

    def __str__(self):
        m = _SLIKE[self.st] if self.barva != Barva.TAROK and self.st > 4 else self.st
        return "%s_%s" % (self.barva, m)

    __repr__ = __str__


def karte_iz_maske(mask: int) -> List[Karta]:
    """Cards of a bitboard, ascending by id (= ascending by (suit, rank))."""
    out, mask = [], int(mask)
    while mask:
        low = mask & -mask
        out.append(Karta.iz_id(low.bit_length() - 1))
        mask ^= low
    return out


def maska_iz_kart(karte: Iterable[Karta]) -> int:
    m = 0
    for k in karte:
        m |= 1 << k.v_id()
    return m


class Roka:
    """A hand: ``.karte`` maps each ``Barva`` to a list, sorted at construction; cards picked up later are
    appended unsorted (Roka.py:4-21).  The player owns and mutates it, exactly as in the reference."""

    def __init__(self, karte):
        self.karte = {b: [] for b in Barva}
        for k in karte:
            self.karte[k.barva].append(k)
        for lst in self.karte.values():
            lst.sort()

    @classmethod
    def iz_maske(cls, mask: int) -> "Roka":
        return cls(karte_iz_maske(mask))

    def maska(self) -> int:
        return maska_iz_kart(self)

    def igraj_karto(self, k):
        self.karte[k.barva].remove(k)

    def dodaj_karte(self, karte):
        for k in karte:
            self.karte[k.barva].append(k)

    def mozno_zalozit(self):
        """Cards that may be laid down: buggy value < 5 (Roka.py:23-27)."""
        return [k for lst in self.karte.values() for k in lst if k.vrednost() < 5]

    def __contains__(self, karta):
        return isinstance(karta, Karta) and karta in self.karte[karta.barva]

    def __iter__(self):
        for lst in self.karte.values():
            yield from lst

    def __len__(self):
        return sum(len(lst) for lst in self.karte.values())

    def __str__(self):
        return str(sorted(self))

    __repr__ = __str__

    # ---- counting (Roka.py:55-98) ----
    @staticmethod
    def tri_po_tri(kupcek):
        n = len(kupcek)
        for i in range(0, n - n % 3, 3):
            yield kupcek[i:i + 3]
        if n % 3:
            yield kupcek[n - n % 3:]

    @staticmethod
    def vrednost_stiha(stih):
        total = sum(_TOCKE[k.v_id()] for k in stih)
        return total - (1 if len(stih) in (1, 2) else 2)

    @staticmethod
    def prestej(kupcek):
        return sum(Roka.vrednost_stiha(s) for s in Roka.tri_po_tri(kupcek))
