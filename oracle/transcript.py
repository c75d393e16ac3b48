"""TEST INFRASTRUCTURE ONLY -- callback-transcript recorder shared by the fixture generator (built on the REFERENCE's
``Igralec`` base class, oracle/make_golden_transcripts.py) and by the GPU tests (built on ``tarok_b200.Igralec``).

The reference's engine talks to its players through fifteen callbacks (``Igralec.py:32-127``); results of that path were
pinned in round 1, the SEQUENCE and ARGUMENTS were not.  A ``TranscriptPlayer`` records every callback it receives --
name, id_igre, and the arguments reduced to ints (card ids, masks, seat indices) -- into one list per game, and takes its
decisions from a deterministic rule keyed by (seed, id_igre, seat, decision counter), choosing by POSITION in the lists the
engine hands over (``mozne``, ``mozno_zalozit()``, ``kupcki``), so both engines make the same decisions exactly as long as
they show the players the same things.

Event formats (all ints; ``hist`` = [length, 12-hex sha256 of the canonical history]):
    nova_igra [seat, hand ids in Roka iteration order, player indices of `igralci`]      pripavi_licitiram [seat]
    licitiram [seat, min_igra, obvezno | -99, prednost, returned]                        izberi_barvo_kralja [seat, returned]
    konec_licitiranja [seat, declarer's player index, tip, king | 7]
    pripravi_izbral_iz_talona [seat, groups, st]     menjaj_iz_talona [seat, groups, st, returned, discards in pile order]
    izbral_iz_talona [seat, groups, chosen]
    pripravi_igraj_karto [seat, stih, mozne (in list order), hist]      igraj_karto [seat, stih, mozne, hist, returned]
    rezultat_stiha [seat, stih (5 cards in Klop's first six tricks), sem_pobral]         rezultat_igre [seat, score, hist]
    poglej_karte_odprtega_beraca [seat, the declarer's hand ids for this game]
    predict_licitiram / predict_izberi_iz_talona / predict_igraj_karto [player index]    (scheduler level, no id_igre)
"""
from __future__ import annotations

import hashlib
import json
import random


def _ids(cards):
    return [int(k.v_id()) for k in cards]


class Journal:
    """The shared sink: one event list per id_igre plus the global sequence (for the lock-step scheduler tests)."""

    def __init__(self):
        self.per_game = {}
        self.everything = []

    def add(self, id_igre, name, *args):
        ev = [name] + list(args)
        if id_igre is not None:
            self.per_game.setdefault(int(id_igre), []).append(ev)
        self.everything.append([None if id_igre is None else int(id_igre)] + ev)

    @staticmethod
    def digest(events) -> str:
        return hashlib.sha256(json.dumps(events, separators=(",", ":")).encode()).hexdigest()[:16]

    def phases(self):
        """The global sequence cut at the scheduler's predict_* calls (Tarok.py:39-56): a list of {id_igre: [events]} --
        inside a phase the engine may serve the games in any order, every game's own order counts."""
        out, cur = [], {}
        for ev in self.everything:
            if ev[0] is None:
                if cur:
                    out.append(cur)
                    cur = {}
                continue
            cur.setdefault(ev[0], []).append(ev[1:])
        if cur:
            out.append(cur)
        return out


def make_transcript_player(Base, Tip, Barva):
    """``Base`` = the engine's Igralec base class (reference or tarok_b200), ``Tip`` / ``Barva`` = its enums."""

    class TranscriptPlayer(Base):
        def __init__(self, index, journal, seed, intents=None, name=None):
            super().__init__(name if name is not None else "p%d" % index)
            self.index, self.j, self.seed = index, journal, seed
            self.intents = intents or {}          # {id_igre: [index2igra-like (tip value, king suit | None)] per seat}
            self.sedez, self.stevec, self.namen = {}, {}, {}
            self.vsi = {}

        # ---- helpers
        def _rng(self, id_igre, what):
            k = self.stevec.get((id_igre, what), 0)
            self.stevec[(id_igre, what)] = k + 1
            return random.Random("%d/%d/%d/%s/%d" % (self.seed, id_igre, self.sedez[id_igre], what, k))

        def _hist(self, id_igre, zgodovina):
            canon = []
            for kdo, kaj in zgodovina:
                if isinstance(kdo, str):
                    canon.append(["T", int(kaj[0]), [_ids(g) for g in kaj[1]]])
                elif kdo is None:
                    canon.append([9, int(kaj.v_id())])
                else:
                    canon.append([self.vsi[id_igre].index(kdo), int(kaj.v_id())])
            return [len(canon), hashlib.sha256(json.dumps(canon, separators=(",", ":")).encode()).hexdigest()[:12]]

        # ---- deal
        def nova_igra(self, roka, igralci, id_igre):
            super().nova_igra(roka, igralci, id_igre)
            self.sedez[id_igre] = igralci.index(self)
            self.vsi[id_igre] = list(igralci)
            self.j.add(id_igre, "nova_igra", self.sedez[id_igre], _ids(roka), [p.index for p in igralci])

        # ---- auction
        def pripavi_licitiram(self, id_igre):
            self.j.add(id_igre, "pripavi_licitiram", self.sedez[id_igre])

        def predict_licitiram(self):
            self.j.add(None, "predict_licitiram", self.index)

        def licitiram(self, min_igra, id_igre, obvezno=None, prednost=False):
            if id_igre not in self.namen:
                self.namen[id_igre] = Tip(self.intents[id_igre][self.sedez[id_igre]][0])
            r = super().licitiram(self.namen[id_igre], min_igra, id_igre, obvezno, prednost)
            self.namen[id_igre] = r                # Nevronski_igralec overwrites its intent (Igralec.py:304)
            self.j.add(id_igre, "licitiram", self.sedez[id_igre], int(min_igra), -99 if obvezno is None else int(obvezno),
                       int(bool(prednost)), int(r))
            return r

        def izberi_barvo_kralja(self, id_igre):
            b = Barva(self.intents[id_igre][self.sedez[id_igre]][1])
            self.j.add(id_igre, "izberi_barvo_kralja", self.sedez[id_igre], int(b))
            return b

        def konec_licitiranja(self, igralec_ki_igra, tip_igre, id_igre, barva_kralja=None):
            self.j.add(id_igre, "konec_licitiranja", self.sedez[id_igre], self.vsi[id_igre].index(igralec_ki_igra),
                       int(tip_igre), 7 if barva_kralja is None else int(barva_kralja))

        # ---- talon
        def pripravi_izbral_iz_talona(self, talon, st_kupcka, id_igre):
            self.j.add(id_igre, "pripravi_izbral_iz_talona", self.sedez[id_igre], [_ids(g) for g in talon], int(st_kupcka))

        def predict_izberi_iz_talona(self):
            self.j.add(None, "predict_izberi_iz_talona", self.index)

        def menjaj_iz_talona(self, kupcki, st_kart, id_igre):
            # the hand <-> pile bookkeeping of Bot_igralec.menjaj_iz_talona (Igralec.py:161-171), decisions by position
            g = self._rng(id_igre, "group").randrange(len(kupcki))
            roka = self.roka[id_igre]
            roka.dodaj_karte(kupcki[g])
            mozno = roka.mozno_zalozit()
            zalozi = self._rng(id_igre, "discard").sample(list(mozno), st_kart)
            self.kupcek[id_igre].extend(zalozi)
            for k in zalozi:
                roka.igraj_karto(k)
            self.j.add(id_igre, "menjaj_iz_talona", self.sedez[id_igre], [_ids(x) for x in kupcki], int(st_kart), int(g),
                       _ids(zalozi))
            return g

        def izbral_iz_talona(self, talon, st_kupcka, id_igre):
            self.j.add(id_igre, "izbral_iz_talona", self.sedez[id_igre], [_ids(g) for g in talon], int(st_kupcka))

        # ---- play
        def pripravi_igraj_karto(self, karte_na_mizi, mozne, zgodovina, id_igre):
            self.j.add(id_igre, "pripravi_igraj_karto", self.sedez[id_igre], _ids(karte_na_mizi), _ids(mozne),
                       self._hist(id_igre, zgodovina))

        def predict_igraj_karto(self):
            self.j.add(None, "predict_igraj_karto", self.index)

        def igraj_karto(self, karte_na_mizi, mozne, zgodovina, id_igre):
            karta = mozne[self._rng(id_igre, "card").randrange(len(mozne))]
            self.j.add(id_igre, "igraj_karto", self.sedez[id_igre], _ids(karte_na_mizi), _ids(mozne),
                       self._hist(id_igre, zgodovina), int(karta.v_id()))
            return super().igraj_karto(karta, id_igre)

        def rezultat_stiha(self, stih, sem_pobral, id_igre):
            self.j.add(id_igre, "rezultat_stiha", self.sedez[id_igre], _ids(stih), int(bool(sem_pobral)))

        def rezultat_igre(self, st_tock, povzetek_igre, id_igre):
            self.j.add(id_igre, "rezultat_igre", self.sedez[id_igre], int(st_tock), self._hist(id_igre, povzetek_igre))

        def poglej_karte_odprtega_beraca(self, roka, id_igre):
            self.j.add(id_igre, "poglej_karte_odprtega_beraca", self.sedez[id_igre], sorted(_ids(roka[id_igre])))

    return TranscriptPlayer


# index2igra of Nevronski_igralec (Igralec.py:717-745) by value: (Tip value, king suit or None)
def index2igra(idx):
    if idx == 0:
        return -10, None
    if idx <= 12:
        return 10 * (1 + (idx - 1) // 4), (idx - 1) % 4
    return [40, 50, 60, 70, 80][idx - 13], None


def bot_like_intents(rng):
    """One intent index per seat with Bot_igralec-like weight on Naprej/Tri/Dve/Ena (so Klop and Tri/Dve/Ena games are
    frequent) and some weight on every higher contract."""
    pool = [0] * 8 + list(range(1, 13)) + [13, 14, 15, 16, 17]
    return [rng.choice(pool) for _ in range(4)]


# ----------------------------------------------------------------------------------------------------------------------
# Engine-agnostic drivers: `eng` exposes Igralec, Tip_igre, Barva, Igra, Klop, Berac, Navadna_igra, Tarok and
# set_shuffle(fn) (the deal-injection hook: Igra.shuffle of the reference, tarok_b200.igra.shuffle here).
# ----------------------------------------------------------------------------------------------------------------------
class Engine:
    def __init__(self, Igralec, Tip_igre, Barva, Igra, Klop, Berac, Navadna_igra, Tarok, set_shuffle, tarok_kwargs=None):
        self.Igralec, self.Tip_igre, self.Barva = Igralec, Tip_igre, Barva
        self.Igra, self.Klop, self.Berac, self.Navadna_igra, self.Tarok = Igra, Klop, Berac, Navadna_igra, Tarok
        self.set_shuffle = set_shuffle
        self.tarok_kwargs = tarok_kwargs or {}
        self.Player = make_transcript_player(Igralec, Tip_igre, Barva)


def reference_engine():
    from . import ref_harness as H
    ref = H.load_reference()

    def set_shuffle(fn):
        ref.Igra.shuffle = fn
    return Engine(ref.Igralec.Igralec, ref.Tip_igre.Tip_igre, ref.Karta.Barva, ref.Igra.Igra, ref.Klop.Klop, ref.Berac.Berac,
                  ref.Navadna_igra.Navadna_igra, ref.Tarok.Tarok, set_shuffle)


def _inject(eng, perms):
    it = iter(perms)

    def sh(lst):
        lst[:] = [int(x) for x in next(it)]
    eng.set_shuffle(sh)


def run_forced(eng, perm, contract, declarer, king, seed):
    """Igra(P).razdeli() + the per-contract constructor; returns (events of game 0, scores by seat)."""
    j = Journal()
    P = [eng.Player(i, j, seed) for i in range(4)]
    _inject(eng, [perm])
    talon = eng.Igra(P).razdeli()
    if contract == 0:
        g = eng.Klop(P, talon, 0)
    elif contract in (7, 9):
        g = eng.Berac(P, P[declarer], talon, contract == 9, 0)
    else:
        g = eng.Navadna_igra(P, eng.Tip_igre(contract * 10), eng.Barva(king) if king != 7 else None, P[declarer], talon, 0)
    res = list(g.start())[-1]
    return j.per_game[0], [int(res[p]) for p in P]


def run_full(eng, perm, intent_idx, seed):
    """The whole Igra.start() of a single game with fixed intents (index2igra indices per seat)."""
    j = Journal()
    intents = {0: [index2igra(i) for i in intent_idx]}
    P = [eng.Player(i, j, seed, intents) for i in range(4)]
    _inject(eng, [perm])
    res = list(eng.Igra(P).start())[-1]
    return j.per_game[0], [int(res[p]) for p in P]


def run_paralel(eng, perms, intent_idx, seed):
    """Tarok(P, n).paralel_start() (Tarok.py:30-62): n lock-step games, seats rotated by i % 4; returns (journal, rezultati)."""
    import contextlib
    import io
    j = Journal()
    n = len(perms)
    intents = {i: [index2igra(k) for k in intent_idx[i]] for i in range(n)}
    P = [eng.Player(i, j, seed, intents) for i in range(4)]
    _inject(eng, perms)
    t = eng.Tarok(P, n, **eng.tarok_kwargs)
    with contextlib.redirect_stdout(io.StringIO()):
        t.paralel_start()
    return j, [int(t.rezultati[p]) for p in P]
