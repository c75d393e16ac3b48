"""``Igra`` and ``Tarok`` of the reference API on top of the batched CUDA environment.

Reference seam being replaced: ``Tarok.paralel_start`` (Tarok.py:30-62) advances N
``Igra`` generators (Igra.py:26-62) in lock-step and each generator calls back into the players
(``Igralec`` protocol, Igralec.py:32-127).  Here the N games are ONE ``TarokEnv``; the rules run in
the sm_100a kernels and this module only (a) translates between the device bitboards and the
``Karta``/``Roka`` objects the callbacks expect and (b) invokes the callbacks in the reference's order:

    nova_igra -> pripavi_licitiram -> [predict_licitiram] -> licitiram* -> izberi_barvo_kralja ->
    konec_licitiranja -> pripravi_izbral_iz_talona -> [predict_izberi_iz_talona] -> menjaj_iz_talona ->
    izbral_iz_talona -> 48 x (pripravi_igraj_karto -> [predict_igraj_karto] -> igraj_karto) with
    rezultat_stiha after every 4th card -> rezultat_igre

Two speeds:
* four ``Bot_igralec`` (``device_policy == "bot"``): no callbacks at all, the whole batch is dealt, bid,
  played and scored on the device (Philox draws with Bot_igralec's distributions) -- the fast path;
* anything else: the callback path above, one device launch per phase for the whole batch.
"""
from __future__ import annotations

import os
from copy import deepcopy
from itertools import cycle
from typing import List, Optional

import numpy as np
import torch

from . import env as E
from .igralec import Igralec
from .karte import Barva, Karta, Roka, Tip_igre, maska_iz_kart

#: Deal-injection hook with the reference's semantics: when set to a callable it is applied to
#: ``list(range(54))`` exactly like ``random.shuffle`` at Igra.py:66-67 and the result is dealt
#: (seat i <- karte[12i:12i+12], talon <- karte[48:54]).  ``None`` = Philox deal on the device.
shuffle = None

_KRALJ_IGRE = (Tip_igre.Tri, Tip_igre.Dve, Tip_igre.Ena)
_VELJAVNE = tuple(t for t in Tip_igre if t != Tip_igre.Naprej)


def _u64(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy().view(np.uint64)


def _st_za_menjat(tip: Tip_igre) -> int:
    """Cards taken from / laid into the talon (Navadna_igra.py:36-58); 0 = no exchange."""
    return {Tip_igre.Tri: 3, Tip_igre.Solo_tri: 3, Tip_igre.Dve: 2, Tip_igre.Solo_dve: 2,
            Tip_igre.Ena: 1, Tip_igre.Solo_ena: 1}.get(tip, 0)


class Partije:
    """Lock-step engine: N concurrent games, arbitrary ``Igralec`` objects, rules on the device.

    ``sedezi[g]`` lists the four players of game g in seat order, ``ids[g]`` is the ``id_igre`` the
    players are called with."""

    def __init__(self, sedezi: List[List[Igralec]], ids: List[int], device: int = 0, seed: Optional[int] = None,
                 prvi_id: int = 0, reference_lockstep_quirks: bool = False):
        self.sedezi, self.ids, self.n = sedezi, list(ids), len(sedezi)
        self.quirks = bool(reference_lockstep_quirks)
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")      # the reference is unseeded (main.py:174)
        self.env = E.TarokEnv(self.n, seed=seed, device=device, history=False)
        self.prvi_id = prvi_id
        self.zgodovina = [[] for _ in range(self.n)]
        self.stih = [[] for _ in range(self.n)]
        self.rezultati: List[Optional[dict]] = [None] * self.n
        self.tip: List[Optional[Tip_igre]] = [None] * self.n
        self.kdo = [0] * self.n
        self.kralj: List[Optional[Barva]] = [None] * self.n
        self.talon: List[List[Karta]] = [[] for _ in range(self.n)]
        self._kupcki = {}
        self._mozne = {}
        self._meta = None

    # ---------------------------------------------------------------- helpers
    def _igralci_vseh(self):
        seen, out = set(), []
        for s in self.sedezi:
            for p in s:
                if id(p) not in seen:
                    seen.add(id(p))
                    out.append(p)
        return out

    def _beri_meta(self):
        m = _u64(self.env.meta[: self.n])
        f = lambda sh, b: ((m >> np.uint64(sh)) & np.uint64((1 << b) - 1)).astype(np.int64)
        self._meta = dict(faza=f(E.M_PHASE, 2), zacne=f(E.M_LEADER, 2), pos=f(E.M_POS, 2), stihi=f(E.M_TRICKS, 4),
                          zmaga=f(E.M_WINNER, 2), konec_stiha=f(E.M_TRICKDONE, 1), napaka=f(E.M_ERR, 1))
        return self._meta

    def _preveri_napake(self, kje):
        if int(self._meta["napaka"].sum()):
            g = int(np.nonzero(self._meta["napaka"])[0][0])
            raise Exception("tarok_b200: device rejected the action (%s) in game id %s" % (kje, self.ids[g]))

    # ---------------------------------------------------------------- phases
    def razdeli(self):
        """Igra.razdeli (Igra.py:65-73) for the whole batch + ``nova_igra`` callbacks."""
        if shuffle is not None:
            perm = np.empty((self.n, 54), np.uint8)
            for g in range(self.n):
                karte = list(range(54))
                shuffle(karte)
                perm[g] = karte
            self.env.set_deals(perm, self.prvi_id)
        else:
            self.env.deal(self.prvi_id)
        roke = _u64(self.env.hands[:, : self.n])
        red = _u64(self.env.talon_order[: self.n])
        for g in range(self.n):
            self.talon[g] = [Karta.iz_id((int(red[g]) >> (6 * i)) & 63) for i in range(6)]
            for s, igralec in enumerate(self.sedezi[g]):
                igralec.nova_igra(Roka.iz_maske(int(roke[s, g])), self.sedezi[g], self.ids[g])

    def pripravi_licitiranje(self):
        for g in range(self.n):
            for igralec in self.sedezi[g]:
                igralec.pripavi_licitiram(self.ids[g])

    @staticmethod
    def licitiraj(igralci, id_igre):
        """The bidding round of Igra.licitacija (Igra.py:81-114) over the players' ``licitiram``."""
        Naprej = Tip_igre.Naprej
        v_igri, najvisja = set(), Tip_igre.Tri
        for s in (1, 2, 3):
            n = igralci[s].licitiram(najvisja, id_igre)
            if n != Naprej:
                v_igri.add(s)
            najvisja = max(najvisja, n)
        if najvisja == Tip_igre.Tri:                       # nobody bid: forehand's own game or Klop
            return 0, igralci[0].licitiram(Naprej, id_igre, Tip_igre.Klop)
        n = igralci[0].licitiram(najvisja, id_igre, prednost=True)
        if n != Naprej:
            v_igri.add(0)
        najvisja = max(najvisja, n)
        ima_igro = min(v_igri)
        while len(v_igri) != 1:
            ostali = set()
            vrsta = sorted(v_igri - {0}) + ([0] if 0 in v_igri else [])
            for s in vrsta:
                if s == ima_igro:
                    n = igralci[s].licitiram(najvisja, id_igre, najvisja)
                else:
                    n = igralci[s].licitiram(najvisja, id_igre)
                if n != Naprej:
                    ostali.add(s)
                    ima_igro, najvisja = s, n
            v_igri = ostali
        return ima_igro, najvisja

    def licitacija(self):
        """Auction via callbacks, king call, contract start on the device, ``konec_licitiranja``."""
        c = np.zeros(self.n, np.uint8); d = np.zeros(self.n, np.uint8); k = np.full(self.n, E.NO_KING, np.uint8)
        for g in range(self.n):
            igralci, idg = self.sedezi[g], self.ids[g]
            kdo, tip = self.licitiraj(igralci, idg)
            if int(tip) not in [int(t) for t in _VELJAVNE]:
                raise Exception("Igra ni definirana:" + str(tip))                       # Igra.py:55
            tip = Tip_igre(int(tip))
            barva = None
            if tip in _KRALJ_IGRE:
                barva = igralci[kdo].izberi_barvo_kralja(idg)                           # Igra.py:42-43
                assert barva != Barva.TAROK                                             # Navadna_igra.py:21
                k[g] = int(barva)
            self.tip[g], self.kdo[g], self.kralj[g] = tip, kdo, barva
            c[g], d[g] = tip.code, kdo
        self.env.force_contract(c, d, k)
        for g in range(self.n):
            for igralec in self.sedezi[g]:
                igralec.konec_licitiranja(self.sedezi[g][self.kdo[g]], self.tip[g], self.ids[g], self.kralj[g])
        self._beri_meta()
        self._preveri_napake("contract")

    def pripravi_menjavo(self):
        """``pripravi_izbral_iz_talona`` for every declarer that exchanges (Navadna_igra.py:59-60)."""
        for g in range(self.n):
            k = _st_za_menjat(self.tip[g])
            if k:
                kupcki = [self.talon[g][i:i + k] for i in range(0, 6, k)]               # odpri_talon
                self._kupcki[g] = kupcki
                self.sedezi[g][self.kdo[g]].pripravi_izbral_iz_talona(deepcopy(kupcki), k, self.ids[g])

    def menjaj(self):
        """``menjaj_iz_talona`` + ``izbral_iz_talona`` (Navadna_igra.py:62-66), applied on the device."""
        if not self._kupcki:
            return
        skupina = np.zeros(self.n, np.uint8)
        zalozil = np.zeros(self.n, np.uint64)
        for g, kupcki in self._kupcki.items():
            idg, igralec = self.ids[g], self.sedezi[g][self.kdo[g]]
            k = _st_za_menjat(self.tip[g])
            st = igralec.menjaj_iz_talona(deepcopy(kupcki), k, idg)
            self.zgodovina[g].append(("Talon", (st, deepcopy(kupcki))))
            for i in self.sedezi[g]:
                i.izbral_iz_talona(deepcopy(kupcki), st, idg)
            skupina[g] = st
            zalozil[g] = maska_iz_kart(igralec.kupcek[idg])       # the player moved the cards itself
        self.env.exchange(skupina, zalozil)
        self._beri_meta()
        self._preveri_napake("talon exchange")
        roke = _u64(self.env.hands[:, : self.n])
        for g in self._kupcki:
            igralec = self.sedezi[g][self.kdo[g]]
            if igralec.roka[self.ids[g]].maska() != int(roke[self.kdo[g], g]):
                raise Exception("tarok_b200: player's hand and device hand differ after the talon exchange")

    def _mozne_karte(self, roka: Roka, spodnja: Optional[Karta], maska: int) -> List[Karta]:
        """The legal set as the list the reference would hand to the player: the player's own suit list
        order when following suit / trumping, all cards sorted otherwise (Navadna_igra.py:158-168,
        Klop.py:96-133); membership comes from the device mask."""
        if spodnja is not None and roka.karte[spodnja.barva]:
            kand = roka.karte[spodnja.barva]
        elif spodnja is not None and roka.karte[Barva.TAROK]:
            kand = roka.karte[Barva.TAROK]
        else:
            kand = sorted(roka)
        mozne = [k for k in kand if (maska >> k.v_id()) & 1]
        if maska_iz_kart(mozne) != maska:
            raise Exception("tarok_b200: player's hand and device hand differ")
        return mozne

    def pripravi_poteze(self, samo=None):
        """``pripravi_igraj_karto`` for the seat to move of every live game (of the games in ``samo`` only, if given)."""
        m = self._beri_meta()
        maske = _u64(self.env.mask[: self.n])
        self._mozne = {}
        for g in np.nonzero(m["faza"] == E.PH_PLAY)[0]:
            g = int(g)
            if samo is not None and g not in samo:
                continue
            idg, igralci = self.ids[g], self.sedezi[g]
            if self.tip[g] == Tip_igre.Odprti_berac and m["stihi"][g] == 1 and m["pos"][g] == 0:
                berac = igralci[self.kdo[g]]
                for i in igralci:                                                       # Berac.py:22-25
                    if i is not berac:
                        i.poglej_karte_odprtega_beraca(berac.roka, idg)
            igralec = igralci[int(m["zacne"][g] + m["pos"][g]) & 3]
            spodnja = self.stih[g][0] if self.stih[g] else None
            mozne = self._mozne_karte(igralec.roka[idg], spodnja, int(maske[g]))
            self._mozne[g] = (igralec, mozne, deepcopy(mozne))
            igralec.pripravi_igraj_karto(deepcopy(self.stih[g]), mozne, self.zgodovina[g], idg)

    def igraj_poteze(self):
        """``igraj_karto`` for every game prepared by the last ``pripravi_poteze``; the other live games keep their state
        (TAROK_CARD_SKIP)."""
        karte = np.full(self.n, 0xFE, np.uint8)
        for g, (igralec, mozne, kopija) in self._mozne.items():
            idg = self.ids[g]
            karta = igralec.igraj_karto(deepcopy(self.stih[g]), mozne, self.zgodovina[g], idg)
            if karta not in kopija:
                raise Exception(str(igralec) + str(igralec.__class__) + " Karte ne mores igarti. Karta: " + str(karta)
                                + " karte na mizi:" + str(self.stih[g]) + " Roka" + str(igralec.roka) + "Mozne"
                                + str(mozne) + "deep mozne:" + str(kopija))
            self.zgodovina[g].append((igralec, karta))
            self.stih[g].append(karta)
            karte[g] = karta.v_id()
        prej = self._meta
        self.env.step(karte)
        m = self._beri_meta()
        self._preveri_napake("card")
        koncane = []
        for g in self._mozne:
            if not m["konec_stiha"][g]:
                continue
            idg, igralci, stih = self.ids[g], self.sedezi[g], self.stih[g]
            if self.tip[g] == Tip_igre.Klop and prej["stihi"][g] < 6:                   # Klop.py:67-71
                tk = self.talon[g][5 - int(prej["stihi"][g])]
                stih.append(tk)
                self.zgodovina[g].append((None, tk))
            zmaga = int(m["zmaga"][g])
            igralci[zmaga].kupcek[idg].extend(stih)
            for s in range(4):
                igralci[s].rezultat_stiha(stih, s == zmaga, idg)
            self.stih[g] = []
            if m["faza"][g] == E.PH_DONE:
                koncane.append(g)
        if koncane:
            tocke = self.env.score().cpu().numpy()
            for g in koncane:
                pisejo = {self.sedezi[g][s]: int(tocke[g, s]) for s in range(4)}
                for i in self.sedezi[g]:
                    i.rezultat_igre(pisejo[i], self.zgodovina[g], self.ids[g])
                self.rezultati[g] = pisejo

    def zive(self) -> bool:
        return bool((self._meta["faza"] == E.PH_PLAY).any())

    def _solo_brez(self):
        return {g for g in range(self.n) if self.tip[g] == Tip_igre.Solo_brez}

    def faze(self):
        """Generator over the lock-step phases; yields the reference's marker strings and finally the
        list of per-game result dicts.

        ``reference_lockstep_quirks``: in the reference a Solo_brez game yields one item fewer (Navadna_igra.py:48,59-68: no
        'Pripravljen menjat'), so under ``Tarok.paralel_start`` (Tarok.py:42-56) it runs ONE STEP AHEAD of every other game:
        its first card is prepared while the others prepare the talon exchange and played while they exchange, and it
        finishes one iteration early.  With the flag the phases reproduce exactly that interleaving (which
        callbacks fall between which predict_* calls); without it every game plays card t in iteration t."""
        self.razdeli()
        self.pripravi_licitiranje()
        yield "Pripravljen_licitirat"
        self.licitacija()
        self.pripravi_menjavo()
        naprej = self._solo_brez() if self.quirks else set()
        if naprej:
            self.pripravi_poteze(samo=naprej)            # Tarok.py:42: next(inner) of a Solo_brez game reaches its first card
        yield "Pripravljen menjat"
        self.menjaj()
        if naprej:
            self.igraj_poteze()                          # Tarok.py:45: ... and plays it while the others exchange
        self.pripravi_poteze()
        while self.zive():
            yield "Pripravljen igrat karto"
            self.igraj_poteze()
            if self.zive():
                self.pripravi_poteze()
        yield self.rezultati

    def zapri(self):
        self.env.close()


def _pozeni(partije: Partije, igralci):
    """Runs the phases with the scheduler's predict calls in between (Tarok.py:36-56)."""
    rezultat = None
    for korak in partije.faze():
        if korak == "Pripravljen_licitirat":
            for i in igralci:
                i.predict_licitiram()
        elif korak == "Pripravljen menjat":
            for i in igralci:
                i.predict_izberi_iz_talona()
        elif korak == "Pripravljen igrat karto":
            for i in igralci:
                i.predict_igraj_karto()
        else:
            rezultat = korak
    partije.zapri()
    return rezultat


class Igra:
    """One deal: ``Igra(igralci, multi_games=False, id=0)`` (Igra.py:18-25)."""

    def __init__(self, igralci, multi_games=False, id=0, device=0, seed=None):
        self.igralci = igralci
        self.zgodovina = []
        self.multi_games = multi_games
        if self.multi_games and id is None:
            raise Exception("Can not have multiple games without id")
        self.id = id
        self._device, self._seed = device, seed
        self._partije = None

    def _nove_partije(self):
        self._partije = Partije([self.igralci], [self.id], device=self._device, seed=self._seed)
        return self._partije

    def start(self):
        """Generator with the reference's yield shape (Igra.py:26-62): a single game yields exactly one item,
        the ``{player: score}`` dict; with ``multi_games`` it yields ``'Pripravljen_licitirat'`` and then the
        inner generator of the contract (``'Pripravljen menjat'``, ``'Pripravljen igrat karto'`` ..., dict)."""
        p = self._nove_partije()
        faze = p.faze()
        if not self.multi_games:
            for korak in faze:
                if isinstance(korak, list):
                    self.zgodovina = p.zgodovina[0]
                    p.zapri()
                    yield korak[0]
            return
        yield next(faze)

        def notranja():
            for korak in faze:
                if isinstance(korak, list):
                    self.zgodovina = p.zgodovina[0]
                    p.zapri()
                    yield korak[0]
                else:
                    yield korak
        yield notranja()

    def razdeli(self):
        """Deals and returns the ordered talon as a list of ``Karta`` (Igra.py:65-73)."""
        p = self._partije or self._nove_partije()
        p.razdeli()
        return list(p.talon[0])

    def licitacija(self):
        """Generator: ``'Pripravljen_licitirat'`` then ``(seat, contract)`` (Igra.py:75-114)."""
        for i in self.igralci:
            i.pripavi_licitiram(self.id)
        yield "Pripravljen_licitirat"
        yield Partije.licitiraj(self.igralci, self.id)


class Tarok:
    """``Tarok(igralci, st_iger)``: N concurrent games, game i seats the players rotated by i % 4
    (Tarok.py:7-62); ``.rezultati`` accumulates the score of every player object."""

    izpis = True

    def __init__(self, igralci, st_iger=None, device=0, seed=None, reference_lockstep_quirks=False):
        assert len({i.ime for i in igralci}) == 4
        #: reproduce the reference scheduler's Solo_brez one-step lead (see Partije.faze)
        self.reference_lockstep_quirks = bool(reference_lockstep_quirks)
        self.igralci = igralci
        self.rezultati = {i: 0 for i in igralci}
        self.radelci = {i: 0 for i in igralci}
        self.st_iger = st_iger
        self.device, self.seed = device, seed
        self.statistika = None

    def stream(self):
        if self.st_iger is None:
            yield from cycle([0, 1, 2, 3])
        else:
            yield from range(self.st_iger)

    def _vsi_na_napravi(self):
        """The device fast path replaces the players' decisions by the in-kernel uniform bot, so it is only taken when
        every player IS that bot: ``Bot_igralec`` itself, or a subclass that overrides none of its decision methods
        or callbacks (an override would silently never run)."""
        from .igralec import Bot_igralec, ODLOCITVE
        for i in self.igralci:
            if getattr(i, "device_policy", None) != "bot" or not isinstance(i, Bot_igralec):
                return False
            if any(getattr(type(i), m) is not getattr(Bot_igralec, m) for m in ODLOCITVE):
                return False
        return True

    def start(self):
        """Sequential single games, all with id 0 and no seat rotation (Tarok.py:23-28, Q14)."""
        if self._vsi_na_napravi():
            self._na_napravi(self.st_iger, rotacija=False)
        else:
            for i in range(self.st_iger):
                # every sequential game is its own deal: game i draws from Philox counter (seed, i), like game i of the
                # device path (the reference shuffles afresh for every Igra, Igra.py:66-67)
                r = _pozeni(Partije([list(self.igralci)], [0], device=self.device, seed=self.seed, prvi_id=i), self.igralci)
                for k, v in r[0].items():
                    self.rezultati[k] += v
        if self.izpis:
            print(self.rezultati)

    def paralel_start(self):
        """All games at once (Tarok.py:30-62)."""
        if self.st_iger is None:
            raise ValueError("paralel_start needs st_iger")
        if self._vsi_na_napravi():
            self._na_napravi(self.st_iger, rotacija=True)
        else:
            sedezi = [self.igralci[i % 4:] + self.igralci[:i % 4] for i in self.stream()]
            r = _pozeni(Partije(sedezi, list(range(self.st_iger)), device=self.device, seed=self.seed,
                                reference_lockstep_quirks=self.reference_lockstep_quirks), self.igralci)
            for d in r:
                for k, v in d.items():
                    self.rezultati[k] += v
        if self.izpis:
            print(self.rezultati)

    def _na_napravi(self, st_iger, rotacija):
        """Fast path: four Bot_igralec -> deal, Bot bidding, exchange, 48 random plays and scoring entirely
        on the device (the fused rollout kernel: bit-identical to the 50-launch stepwise pipeline, one launch); only the
        32-entry statistics vector comes back."""
        seed = self.seed if self.seed is not None else int.from_bytes(os.urandom(8), "little")
        env = _okolje_na_napravi(st_iger, seed, self.device)
        env.reset_stats()
        env.rollout(E.MODE_AUCTION_BOT, first_game_id=0, fused=True)
        st = env.stats()
        self.statistika = st
        if int(st[E.S_ERRORS]):
            # the reference raises out of random.sample when fewer than k cards can be laid down (Igralec.py:166, Q19)
            raise ValueError("Sample larger than population or is negative (%d of %d deals: a Bot_igralec declarer "
                             "could not lay down enough cards)" % (int(st[E.S_ERRORS]), st_iger))
        vsote = st[E.S_PLAYER:E.S_PLAYER + 4] if rotacija else st[E.S_SEAT:E.S_SEAT + 4]
        for p, igralec in enumerate(self.igralci):
            self.rezultati[igralec] += int(vsote[p])


# The device environment of the Bot fast path is kept between calls (one per device, the last batch size): main.py's loop
# builds a new Tarok(igralci, num_games) for every iteration (main.py:111-116), and allocating + clearing ~170 MB of state per
# million games each time costs far more than playing them (20 ms against 0.3 ms).  ``sprosti_okolja()`` gives the memory back.
_OKOLJA = {}


def _okolje_na_napravi(st_iger, seed, device):
    env = _OKOLJA.get(device)
    if env is not None and (env.n != st_iger or not env._h):
        env.close()
        env = None
    if env is None:
        env = E.TarokEnv(st_iger, seed=seed, device=device)
        env.set_materialise(False)                       # only the result sums are needed (Tarok.py:59-61)
        _OKOLJA[device] = env
    else:
        env.reseed(seed)
    return env


def sprosti_okolja():
    """Releases the device environments kept by the Bot fast path of ``Tarok.start`` / ``paralel_start``."""
    for env in _OKOLJA.values():
        env.close()
    _OKOLJA.clear()


# ------------------------------------------------------------------------------------------------------
# Per-contract entry points (the reference's L1 classes).  They take players that already hold their hands
# (``igralec.roka[id_igre]``, e.g. after ``Igra(P).razdeli()``) and the ordered talon, bypass the bidding and
# play the contract on the device; ``list(obj.start())[-1]`` is the ``{player: score}`` dict.
# ------------------------------------------------------------------------------------------------------
class _Pogodba:
    def __init__(self, igralci, tip, kdo, barva_kralja, talon, id_igre, device=0, seed=None):
        self.igralci, self.talon, self.id_igre = igralci, talon, id_igre
        self._tip, self._kdo, self._barva = tip, kdo, barva_kralja
        self._device, self._seed = device, seed
        self.zgodovina = []

    def start(self):
        p = Partije([self.igralci], [self.id_igre], device=self._device, seed=self._seed)
        idg = self.id_igre
        perm = []
        for i in self.igralci:
            perm.extend(k.v_id() for k in i.roka[idg])
        perm.extend(k.v_id() for k in self.talon)
        if sorted(perm) != list(range(54)):
            raise Exception("tarok_b200: the four hands and the talon do not form a 54-card deck")
        p.env.set_deals(np.array([perm], np.uint8))
        p.talon[0] = list(self.talon)
        p.tip[0], p.kdo[0], p.kralj[0] = self._tip, self._kdo, self._barva
        p.env.force_contract([self._tip.code], [self._kdo], [E.NO_KING if self._barva is None else int(self._barva)])
        p._beri_meta()
        p._preveri_napake("contract")
        self.zgodovina = p.zgodovina[0]
        p.pripravi_menjavo()
        if self._tip != Tip_igre.Solo_brez:              # Solo_brez yields one item fewer (Q17)
            yield "Pripravljen menjat"
        p.menjaj()
        p.pripravi_poteze()
        while p.zive():
            yield "Pripravljen igrat karto"
            p.igraj_poteze()
            if p.zive():
                p.pripravi_poteze()
        p.zapri()
        yield p.rezultati[0]


class Navadna_igra(_Pogodba):
    """``Navadna_igra(igralci, igra, barva_kralja, igralec, talon, id_igre)`` (Navadna_igra.py:15)."""

    def __init__(self, igralci, igra, barva_kralja, igralec, talon, id_igre, **kw):
        if igra in _KRALJ_IGRE:
            assert barva_kralja != Barva.TAROK
        else:
            barva_kralja = None
        super().__init__(igralci, Tip_igre(int(igra)), igralci.index(igralec), barva_kralja, talon, id_igre, **kw)
        self.igra, self.igralec, self.barva_kralja = igra, igralec, barva_kralja


class Klop(_Pogodba):
    """``Klop(igralci, talon, id_igre)`` (Klop.py:16)."""

    def __init__(self, igralci, talon, id_igre, **kw):
        super().__init__(igralci, Tip_igre.Klop, 0, None, talon, id_igre, **kw)


class Berac(_Pogodba):
    """``Berac(igralci, berac, talon, odprti, id_igre)`` (Berac.py:5)."""

    def __init__(self, igralci, berac, talon, odprti, id_igre, **kw):
        tip = Tip_igre.Odprti_berac if odprti else Tip_igre.Berac
        super().__init__(igralci, tip, igralci.index(berac), None, talon, id_igre, **kw)
        self.berac, self.odprti = berac, odprti
