"""Player protocol of the reference (``Igralec.py:32-171``): the callbacks the engine invokes, the
bidding filter, and the uniform-random ``Bot_igralec``.

Only the caller contract is mirrored here (same method names, argument order and return values), so
existing ``Igralec`` subclasses can be driven by the batched CUDA engine (``tarok_b200.igra``).  The
neural players / training code of the reference (``Igralec.py:173-887``) stay upstream: they plug in
unchanged through this protocol.
"""
from __future__ import annotations

import random
import warnings

import numpy as np

from .karte import Barva, Tip_igre

_stevec = 0


class Igralec:
    """Base player.  Per-game state lives in dicts keyed by ``id_igre`` (Igralec.py:40-49)."""

    #: set by subclasses whose decisions the device can take itself (no host callbacks needed)
    device_policy = None

    def __init__(self, ime=None):
        global _stevec
        if ime is None:
            self.ime = str(_stevec)
            _stevec += 1
        else:
            self.ime = str(ime)
        self.roka = dict()
        self.igra = dict()
        self.kupcek = dict()
        self.napovedi = []

    # -- deal -------------------------------------------------------------------------------------
    def nova_igra(self, roka, igralci, id_igre):
        self.roka[id_igre] = roka
        self.igra[id_igre] = None
        self.kupcek[id_igre] = []
        self.napovedi = []

    # -- auction ----------------------------------------------------------------------------------
    def pripavi_licitiram(self, id_igre):
        pass

    def predict_licitiram(self):
        pass

    def licitiram(self, licitiram, min_igra, id_igre, obvezno=None, prednost=False):
        """The bid filter (Igralec.py:58-74): keep the wanted game if it is high enough (>= with
        priority, > without), else the obligatory game if there is one, else pass."""
        dovolj = licitiram >= min_igra if prednost else licitiram > min_igra
        if dovolj:
            return licitiram
        return Tip_igre.Naprej if obvezno is None else obvezno

    def izberi_barvo_kralja(self, id_igre):
        raise NotImplementedError()

    def konec_licitiranja(self, igralec_ki_igra, tip_igre, id_igre, barva_kralja=None):
        pass

    # -- talon ------------------------------------------------------------------------------------
    def pripravi_izbral_iz_talona(self, talon, st_kupcka, id_igre):
        pass

    def predict_izberi_iz_talona(self):
        pass

    def izbral_iz_talona(self, talon, st_kupcka, id_igre):
        pass

    def menjaj_iz_talona(self, kupcki, st_kart, id_igre):
        raise NotImplementedError()

    # -- play -------------------------------------------------------------------------------------
    def pripravi_igraj_karto(self, karte_na_mizi, mozne, zgodovina, id_igre):
        pass

    def predict_igraj_karto(self):
        pass

    def igraj_karto(self, karta, id_igre):
        """Helper subclasses call from their own ``igraj_karto(stih, mozne, zgodovina, id)``: removes the
        card from the player's hand (Igralec.py:82-85)."""
        self.roka[id_igre].igraj_karto(karta)
        return karta

    def rezultat_stiha(self, stih, sem_pobral, id_igre):
        pass

    def rezultat_igre(self, st_tock, povzetek_igre, id_igre):
        pass

    def poglej_karte_odprtega_beraca(self, roka, id_igre):
        warnings.warn("Ne uporablam podatka za odprtega beraca")

    def __contains__(self, item):
        return item in self.roka

    def __str__(self):
        return "Igralec_" + str(self.ime)

    __repr__ = __str__


#: the methods whose override means "this is no longer the plain bot" (decisions and result callbacks)
ODLOCITVE = ("licitiram", "izberi_barvo_kralja", "igraj_karto", "menjaj_iz_talona", "nova_igra", "pripavi_licitiram",
             "konec_licitiranja", "pripravi_izbral_iz_talona", "izbral_iz_talona", "pripravi_igraj_karto", "rezultat_stiha",
             "rezultat_igre", "poglej_karte_odprtega_beraca", "predict_licitiram", "predict_izberi_iz_talona",
             "predict_igraj_karto")


class Bot_igralec(Igralec):
    """Uniform-random legal-move player (Igralec.py:142-171).

    Bids Naprej/Tri/Dve/Ena with p = 1/2, 1/6, 1/6, 1/6 afresh at every call, calls a uniform king suit,
    always takes talon group 0 and lays down a uniform subset of the discardable cards, plays a uniform
    legal card.  Four of these in a ``Tarok`` run entirely on the device (``device_policy``), with the
    same distributions drawn from Philox instead of the host RNG streams."""

    device_policy = "bot"

    def licitiram(self, min_igra, id_igre, obvezno=None, prednost=False):
        zelim = np.random.choice([Tip_igre.Naprej, Tip_igre.Tri, Tip_igre.Dve, Tip_igre.Ena],
                                 p=[0.5, 0.5 / 3, 0.5 / 3, 0.5 / 3])
        return super().licitiram(zelim, min_igra, id_igre, obvezno, prednost)

    def izberi_barvo_kralja(self, id_game):
        return random.choice([Barva.SRCE, Barva.KRIZ, Barva.KARA, Barva.PIK])

    def igraj_karto(self, karte_na_mizi, mozne, zgodovina, id_igre):
        return super().igraj_karto(random.choice(mozne), id_igre)

    def menjaj_iz_talona(self, kupcki, st_kart, id_igre):
        izbrani = 0
        roka = self.roka[id_igre]
        roka.dodaj_karte(kupcki[izbrani])
        zalozi = random.sample(roka.mozno_zalozit(), k=st_kart)
        self.kupcek[id_igre].extend(zalozi)
        for k in zalozi:
            roka.igraj_karto(k)
        return izbrani
