"""TEST INFRASTRUCTURE ONLY -- harness that drives the *real* anzeA/Tarok reference.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.

This module imports the unmodified reference from ``/root/reference`` (read-only, only
present in the build container -- never on the GPU box) and drives it with

* **deal injection**: ``Igra.shuffle`` (the name imported at ``Igra.py:10`` and called at
  ``Igra.py:67``) is rebound to a function that overwrites the list with a chosen
  permutation, so seat *i* receives ``perm[12i:12i+12]`` and the talon ``perm[48:54]``
  (``Igra.py:65-73``);
* **teacher forcing**: a recording ``Igralec`` subclass whose decisions are supplied by a
  ``policy`` object; it records what the engine showed it (legal set ``mozne`` at every
  ``igraj_karto``, the trick and winner flag at ``rezultat_stiha``, the final score and
  history at ``rezultat_igre``).

The records it produces are what ``oracle/make_golden.py`` freezes into ``tests/golden``.
The import needs three modules that the reference names but does not ship
(``pytorch_lightning``, ``pytorch_lightning.callbacks``, ``torch_models`` --
``Igralec.py:19,24,26-27``); they are stubbed in ``sys.modules`` first.
"""
from __future__ import annotations

import hashlib
import os
import sys
import types

def _find_reference():
    """TAROK_REFERENCE_DIR, else ``baseline/_ref`` (the documented install target -- ``pip install --target baseline/_ref
    /root/reference`` fails: the upstream tree has neither setup.py nor pyproject.toml), else ``/root/reference``."""
    env = os.environ.get("TAROK_REFERENCE_DIR")
    if env:
        return env
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.path.join(root, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(cand, "Igra.py")):
            return cand
    return "/root/reference"


REFERENCE_DIR = _find_reference()

_ref = None


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "Igra.py"))


def load_reference():
    """Import the reference modules (once) and return them in a namespace."""
    global _ref
    if _ref is not None:
        return _ref
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_DIR)
    sys.dont_write_bytecode = True  # the reference dir is read-only
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.Trainer = object
        pl.LightningModule = object
        cb = types.ModuleType("pytorch_lightning.callbacks")
        cb.EarlyStopping = object
        pl.callbacks = cb
        sys.modules["pytorch_lightning"] = pl
        sys.modules["pytorch_lightning.callbacks"] = cb
    if "torch_models" not in sys.modules:
        import torch

        tm = types.ModuleType("torch_models")
        tm.torch = torch
        tm.TensorDataset = torch.utils.data.TensorDataset
        sys.modules["torch_models"] = tm
    import Berac  # noqa: E402
    import Igra  # noqa: E402
    import Igralec  # noqa: E402
    import Karta  # noqa: E402
    import Klop  # noqa: E402
    import Navadna_igra  # noqa: E402
    import Roka  # noqa: E402
    import Tarok  # noqa: E402
    import Tip_igre  # noqa: E402

    ns = types.SimpleNamespace(
        Berac=Berac, Igra=Igra, Igralec=Igralec, Karta=Karta, Klop=Klop,
        Navadna_igra=Navadna_igra, Roka=Roka, Tarok=Tarok, Tip_igre=Tip_igre,
    )
    _ref = ns
    return ns


# contract code used everywhere in this repo: code = Tip_igre value / 10  (Klop 0 ... Odprti_berac 9)
CONTRACT_NAMES = ["Klop", "Tri", "Dve", "Ena", "Solo_tri", "Solo_dve", "Solo_ena", "Berac",
                  "Solo_brez", "Odprti_berac"]
NO_KING = 7


def inject_deal(ref, perm):
    """Make the next ``Igra.razdeli`` deal exactly ``perm`` (Igra.py:66-67)."""
    perm = [int(x) for x in perm]
    assert sorted(perm) == list(range(54))

    def _shuffle(lst):
        lst[:] = perm

    ref.Igra.shuffle = _shuffle


def cards_to_mask(cards) -> int:
    m = 0
    for k in cards:
        m |= 1 << k.v_id()
    return m


class GameRecord:
    """Everything the engine exposed to the players during one deal."""

    def __init__(self):
        self.perm = None
        self.contract = None      # code 0..9
        self.declarer = None      # seat 0..3 (0 for Klop)
        self.king = NO_KING       # suit 0..3 or NO_KING
        self.group = 0xFF         # chosen talon group or 0xFF
        self.discard_mask = 0
        self.seats = []           # per play
        self.masks = []
        self.cards = []
        self.winners = []         # per trick, absolute seat
        self.scores = None        # by seat
        self.history = None       # reference zgodovina (objects)
        self.bid_calls = 0
        self.hands_after_deal = None  # 4 masks
        self.talon_after_deal = None  # list of ids in order

    def history_hash(self, players) -> str:
        """sha256 over (seat, card) of every non-"Talon" history entry; seat 9 = Klop talon card
        (definition of SURVEY.md A.6)."""
        b = bytearray()
        for who, what in self.history:
            if isinstance(who, str):
                continue
            seat = 9 if who is None else players.index(who)
            b += bytes((seat, what.v_id()))
        return hashlib.sha256(bytes(b)).hexdigest()[:16]


def make_player_class(ref):
    Base = ref.Igralec.Igralec
    Tip = ref.Tip_igre.Tip_igre

    class RecordingPlayer(Base):
        """Teacher-forced player.  ``policy`` supplies decisions; ``rec`` collects evidence."""

        def __init__(self, seat, policy, rec):
            super().__init__(ime="seat%d" % seat)
            self.seat = seat
            self.policy = policy
            self.rec = rec
            self.intent = {}

        # --- auction ---------------------------------------------------------------------
        def licitiram(self, min_igra, id_igre, obvezno=None, prednost=False):
            self.rec.bid_calls += 1
            want = self.policy.bid(self.seat, self.intent.get(id_igre), min_igra, obvezno, prednost)
            r = super().licitiram(Tip(want), min_igra, id_igre, obvezno, prednost)
            if self.policy.fixed_intent:
                # Nevronski_igralec overwrites its intent with every returned value (Igralec.py:304)
                self.intent[id_igre] = int(r)
            return r

        def izberi_barvo_kralja(self, id_igre):
            return ref.Karta.Barva(self.policy.king(self.seat))

        # --- talon -----------------------------------------------------------------------
        def menjaj_iz_talona(self, kupcki, st_kart, id_igre):
            # same hand<->pile bookkeeping as Bot_igralec.menjaj_iz_talona (Igralec.py:161-171)
            roka = self.roka[id_igre]
            g = self.policy.group(self.seat, kupcki)
            roka.dodaj_karte(kupcki[g])
            mozno = roka.mozno_zalozit()
            zalozi = self.policy.discards(self.seat, mozno, st_kart)
            assert len(zalozi) == st_kart
            self.kupcek[id_igre].extend(zalozi)
            for k in zalozi:
                roka.igraj_karto(k)
            self.rec.group = g
            self.rec.discard_mask = cards_to_mask(zalozi)
            return g

        # --- play ------------------------------------------------------------------------
        def igraj_karto(self, karte_na_mizi, mozne, zgodovina, id_igre):
            self.rec.seats.append(self.seat)
            self.rec.masks.append(cards_to_mask(mozne))
            karta = self.policy.card(self.seat, karte_na_mizi, mozne, zgodovina)
            self.rec.cards.append(karta.v_id())
            return super().igraj_karto(karta, id_igre)

        def rezultat_stiha(self, stih, sem_pobral, id_igre):
            if sem_pobral:
                self.rec.winners.append(self.seat)

        def poglej_karte_odprtega_beraca(self, roka, id_igre):
            pass

        def rezultat_igre(self, st_tock, povzetek_igre, id_igre):
            self.rec.history = povzetek_igre

    return RecordingPlayer


class Policy:
    """Default deterministic policy; subclass or pass callables."""

    fixed_intent = True

    def __init__(self, intents=None, king=0, group=0, card="lo", discard="lo", rng=None):
        self.intents = intents
        self._king = king
        self._group = group
        self._card = card
        self._discard = discard
        self.rng = rng

    def bid(self, seat, current, min_igra, obvezno, prednost):
        if current is not None:
            return current
        return self.intents[seat]

    def king(self, seat):
        return self._king[seat] if isinstance(self._king, (list, tuple)) else self._king

    def group(self, seat, kupcki):
        if self._group == "rand":
            return self.rng.randrange(len(kupcki))
        return self._group

    def discards(self, seat, mozno, k):
        if self._discard == "rand":
            return self.rng.sample(mozno, k)
        ordered = sorted(mozno, key=lambda c: c.v_id())
        return ordered[:k]

    def card(self, seat, stih, mozne, zgodovina):
        if self._card == "lo":
            return min(mozne, key=lambda c: c.v_id())
        if self._card == "hi":
            return max(mozne, key=lambda c: c.v_id())
        return mozne[self.rng.randrange(len(mozne))]


class ScriptedBidPolicy(Policy):
    """Bot_igralec-style bidding: a fresh draw at *every* ``licitiram`` call (Igralec.py:151),
    taken in call order from ``draws`` (Tip_igre values)."""

    fixed_intent = False

    def __init__(self, draws, **kw):
        super().__init__(**kw)
        self.draws = list(draws)
        self.used = 0

    def bid(self, seat, current, min_igra, obvezno, prednost):
        v = self.draws[self.used]
        self.used += 1
        return v


def _finish(rec, players, result):
    rec.scores = [int(result[p]) for p in players]
    return rec


def _setup(ref, perm, policy):
    rec = GameRecord()
    rec.perm = [int(x) for x in perm]
    P = make_player_class(ref)
    players = [P(s, policy, rec) for s in range(4)]
    inject_deal(ref, perm)
    return rec, players


def run_forced(perm, contract, declarer, king, policy):
    """Deal ``perm``, force the contract through the per-contract constructors
    (``Klop(P,talon,0)``, ``Navadna_igra(P,tip,king,P[d],talon,0)``,
    ``Berac(P,P[d],talon,odprti,0)``) and play it out."""
    ref = load_reference()
    Tip = ref.Tip_igre.Tip_igre
    rec, players = _setup(ref, perm, policy)
    talon = ref.Igra.Igra(players).razdeli()
    rec.hands_after_deal = [cards_to_mask(p.roka[0]) for p in players]
    rec.talon_after_deal = [k.v_id() for k in talon]
    rec.contract, rec.declarer, rec.king = contract, declarer, king
    name = CONTRACT_NAMES[contract]
    if name == "Klop":
        g = ref.Klop.Klop(players, talon, 0)
    elif name in ("Berac", "Odprti_berac"):
        g = ref.Berac.Berac(players, players[declarer], talon, name == "Odprti_berac", 0)
    else:
        barva = ref.Karta.Barva(king) if king != NO_KING else None
        g = ref.Navadna_igra.Navadna_igra(players, Tip(contract * 10), barva, players[declarer], talon, 0)
    result = list(g.start())[-1]
    return _finish(rec, players, result), players


def run_full(perm, policy):
    """Deal ``perm`` and run the whole ``Igra.start()`` (auction + king call + play)."""
    ref = load_reference()
    rec, players = _setup(ref, perm, policy)
    outcome = {}
    P0 = type(players[0])
    orig = P0.konec_licitiranja

    def konec(self, igralec_ki_igra, tip_igre, id_igre, barva_kralja=None):
        outcome["c"] = int(tip_igre) // 10
        outcome["d"] = players.index(igralec_ki_igra)
        outcome["k"] = NO_KING if barva_kralja is None else int(barva_kralja)

    P0.konec_licitiranja = konec
    try:
        result = list(ref.Igra.Igra(players).start())[-1]
    finally:
        P0.konec_licitiranja = orig
    rec.contract, rec.declarer, rec.king = outcome["c"], outcome["d"], outcome["k"]
    if rec.contract == 0:
        rec.declarer = 0
    return _finish(rec, players, result), players


def run_auction_only(policy):
    """Run only ``Igra.licitacija`` (Igra.py:75-114); returns (declarer, contract code, bid calls)."""
    ref = load_reference()
    rec = GameRecord()
    P = make_player_class(ref)
    players = [P(s, policy, rec) for s in range(4)]
    gen = ref.Igra.Igra(players).licitacija()
    next(gen)
    seat, tip = next(gen)
    return int(seat), int(tip) // 10, rec.bid_calls


def ref_prestej(ids):
    """``Roka.prestej`` (Roka.py:96-98) on a list of card ids, in the given order."""
    ref = load_reference()
    K = ref.Karta.Karta
    return ref.Roka.Roka.prestej([K.iz_id(i) for i in ids])
