"""TEST INFRASTRUCTURE ONLY -- freezes outputs of the REAL reference into tests/golden/.

Run in the build container (the only place /root/reference exists):

    python -m oracle.make_golden

Every array below is produced by the unmodified anzeA/Tarok Python engine driven through
oracle/ref_harness.py (deal injection + teacher-forced recording players).  The fixtures travel
to the GPU box, where they pin both the C oracle (tests -m "not gpu") and the CUDA path (-m gpu).
"""
from __future__ import annotations

import hashlib
import itertools
import json
import os
import random

import numpy as np

from . import ref_harness as H

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
TIP_VALUES = [-10, 0, 10, 20, 30, 40, 50, 60, 70, 80, 90]


def index2igra(idx):
    """Nevronski_igralec.generete_igra2index_and_index2igra (Igralec.py:717-745), by value."""
    if idx == 0:
        return -10, H.NO_KING
    if idx <= 12:
        return 10 * (1 + (idx - 1) // 4), (idx - 1) % 4
    return [40, 50, 60, 70, 80][idx - 13], H.NO_KING


def gen_kat():
    ident = list(range(54))
    r = list(range(54))
    random.Random(12345).shuffle(r)
    rows = []
    for name, deal in (("identity", ident), ("R12345", r)):
        for pol in ("lo", "hi"):
            for c, d, k in ((0, 0, 7), (1, 0, 0), (2, 1, 1), (3, 2, 2), (1, 3, 3), (4, 0, 7), (5, 1, 7),
                            (6, 2, 7), (8, 3, 7), (7, 0, 7), (7, 3, 7), (9, 2, 7)):
                rec, players = H.run_forced(deal, c, d, k, H.Policy(card=pol))
                rows.append(dict(deal=name, perm=deal, pol=pol, contract=c, declarer=d, king=k,
                                 scores=rec.scores, plays=len(rec.cards), hash=rec.history_hash(players),
                                 cards=rec.cards, group=rec.group, discard_mask=rec.discard_mask))
    with open(os.path.join(OUT, "kat.json"), "w") as f:
        json.dump(rows, f)
    print("kat.json", len(rows))


def _pack(recs, extra=None):
    n = len(recs)
    a = dict(
        perm=np.zeros((n, 54), np.uint8), contract=np.zeros(n, np.uint8), declarer=np.zeros(n, np.uint8),
        king=np.zeros(n, np.uint8), group=np.zeros(n, np.uint8), discard_mask=np.zeros(n, np.uint64),
        seat=np.full((n, 48), 0xFF, np.uint8), mask=np.zeros((n, 48), np.uint64),
        card=np.full((n, 48), 0xFF, np.uint8), winner=np.full((n, 12), 0xFF, np.uint8),
        scores=np.zeros((n, 4), np.int16), plays=np.zeros(n, np.uint8),
        hands=np.zeros((n, 4), np.uint64), piles=np.zeros((n, 4), np.uint64),
        hands0=np.zeros((n, 4), np.uint64),
    )
    for i, (rec, players) in enumerate(recs):
        a["perm"][i] = rec.perm
        a["contract"][i], a["declarer"][i], a["king"][i] = rec.contract, rec.declarer, rec.king
        a["group"][i], a["discard_mask"][i] = rec.group, rec.discard_mask
        p = len(rec.cards)
        a["plays"][i] = p
        a["seat"][i, :p], a["mask"][i, :p], a["card"][i, :p] = rec.seats, rec.masks, rec.cards
        a["winner"][i, :len(rec.winners)] = rec.winners
        a["scores"][i] = rec.scores
        for s, pl in enumerate(players):
            a["hands"][i, s] = H.cards_to_mask(pl.roka[0])
            a["piles"][i, s] = H.cards_to_mask(pl.kupcek[0])
            if rec.hands_after_deal is not None:
                a["hands0"][i, s] = rec.hands_after_deal[s]
    if extra:
        a.update(extra)
    return a


def build_traces(per_contract, seed):
    """Forced-contract games played by the real reference with seeded random decisions -> packed arrays."""
    rng = random.Random(seed)
    recs = []
    for c in range(10):
        for _ in range(per_contract):
            perm = list(range(54))
            rng.shuffle(perm)
            d = 0 if c == 0 else rng.randrange(4)
            k = rng.randrange(4) if 1 <= c <= 3 else H.NO_KING
            pol = H.Policy(card="rand", discard="rand", group="rand", rng=rng)
            try:
                recs.append(H.run_forced(perm, c, d, k, pol))
            except ValueError:      # random.sample: fewer than k discardable cards (Q19)
                continue
    return _pack(recs)


def gen_traces(per_contract=150, seed=20261018):
    a = build_traces(per_contract, seed)
    np.savez_compressed(os.path.join(OUT, "traces_forced.npz"), **a)
    print("traces_forced.npz", len(a["contract"]))


def build_full(n, seed):
    """Whole Igra.start(): fixed-intent auction (Nevronski model) + king call + dispatch + play."""
    rng = random.Random(seed)
    recs, intents = [], []
    while len(recs) < n:
        perm = list(range(54))
        rng.shuffle(perm)
        # two thirds of the deals bid like Bot_igralec (Naprej 1/2, Tri / Dve / Ena 1/6 each, Igralec.py:151) -- the auctions
        # in which Klop / Tri / Dve / Ena are decided --, one third uniformly over index2igra with extra weight on Naprej
        if len(recs) % 3:
            idx = [rng.choice([0, 0, 0, rng.randrange(1, 5), rng.randrange(5, 9), rng.randrange(9, 13)]) for _ in range(4)]
        else:
            idx = [rng.choice([0, 0, 0] + list(range(18))) for _ in range(4)]
        tips = [index2igra(i)[0] for i in idx]
        kings = [index2igra(i)[1] for i in idx]
        pol = H.Policy(intents=tips, king=kings, card="rand", discard="rand", group="rand", rng=rng)
        try:
            recs.append(H.run_full(perm, pol))
        except ValueError:
            continue
        intents.append(idx)
    return _pack(recs, dict(intent=np.array(intents, np.uint8)))


def gen_full(n=1500, seed=7):
    a = build_full(n, seed)
    np.savez_compressed(os.path.join(OUT, "traces_full.npz"), **a)
    print("traces_full.npz", len(a["contract"]), np.bincount(a["contract"], minlength=10))


def gen_auction():
    combos = list(itertools.product(TIP_VALUES, repeat=4))
    decl = np.zeros(len(combos), np.uint8)
    con = np.zeros(len(combos), np.uint8)
    calls = np.zeros(len(combos), np.uint8)
    for i, c in enumerate(combos):
        d, k, n = H.run_auction_only(H.Policy(intents=list(c)))
        decl[i], con[i], calls[i] = d, k, n
    np.savez_compressed(os.path.join(OUT, "auction_fixed.npz"),
                        intents=(np.array(combos, np.int16) // 10).astype(np.int8),
                        declarer=decl, contract=con, calls=calls)
    # the hash quoted in SURVEY.md A.3 (9 index2igra values, product order of sorted values)
    vals9 = [-10, 10, 20, 30, 40, 50, 60, 70, 80]
    h = hashlib.sha256()
    lut = {c: (int(decl[i]), int(con[i])) for i, c in enumerate(combos)}
    for c in itertools.product(vals9, repeat=4):
        d, k = lut[c]
        h.update(bytes((d, k * 10 + 10)))
    print("auction_fixed.npz", len(combos), "max calls", calls.max(), "survey-hash", h.hexdigest()[:16])

    rng = random.Random(99)
    n = 6000
    draws = np.zeros((n, 16), np.int8)
    d2, c2, k2 = np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    for i in range(n):
        if i < 4000:   # Bot_igralec distribution (Igralec.py:151)
            seq = [rng.choice([-10, -10, -10, 10, 20, 30]) for _ in range(16)]
        else:          # anything goes
            seq = [rng.choice(TIP_VALUES) for _ in range(16)]
        pol = H.ScriptedBidPolicy(seq + [-10] * 64)
        d, k, nc = H.run_auction_only(pol)
        assert nc <= 16
        draws[i] = [v // 10 for v in seq]
        d2[i], c2[i], k2[i] = d, k, nc
    np.savez_compressed(os.path.join(OUT, "auction_scripted.npz"), draws=draws, declarer=d2, contract=c2,
                        calls=k2)
    print("auction_scripted.npz", n, "max calls", k2.max())


def gen_units(seed=5):
    ref = H.load_reference()
    K = ref.Karta.Karta
    rng = random.Random(seed)
    # Roka.prestej on ordered piles
    piles, vals = [], []
    for _ in range(3000):
        n = rng.randrange(0, 55)
        ids = rng.sample(range(54), n)
        piles.append(ids)
        vals.append(H.ref_prestej(ids))
    stride = 54
    arr = np.full((len(piles), stride), 0xFF, np.uint8)
    for i, p in enumerate(piles):
        arr[i, :len(p)] = p
    # legal moves: (hand, lead) -> Navadna_igra.mozne_karte / Klop.mozne_karte; discardable set
    nav = ref.Navadna_igra.Navadna_igra.__new__(ref.Navadna_igra.Navadna_igra)
    klop = ref.Klop.Klop.__new__(ref.Klop.Klop)
    hands, leads, m_nav, m_klop, m_disc = [], [], [], [], []
    for _ in range(20000):
        n = rng.randrange(1, 16)
        ids = rng.sample(range(54), n)
        roka = ref.Roka.Roka([K.iz_id(i) for i in ids])
        lead = rng.choice([None] + list(range(54)))
        if lead is not None and lead in ids:
            lead = None
        lk = None if lead is None else K.iz_id(lead)
        hands.append(sum(1 << i for i in ids))
        leads.append(0xFF if lead is None else lead)
        m_nav.append(H.cards_to_mask(nav.mozne_karte(lk, roka)))
        m_klop.append(H.cards_to_mask(klop.mozne_karte(lk, roka)))
        m_disc.append(H.cards_to_mask(roka.mozno_zalozit()))
    # card tables
    ids = list(range(54))
    np.savez_compressed(
        os.path.join(OUT, "units.npz"),
        pile_ids=arr, pile_len=np.array([len(p) for p in piles], np.uint8), pile_val=np.array(vals, np.int32),
        hand=np.array(hands, np.uint64), lead=np.array(leads, np.uint8),
        mozne_navadna=np.array(m_nav, np.uint64), mozne_klop=np.array(m_klop, np.uint64),
        mozno_zalozit=np.array(m_disc, np.uint64),
        card_barva=np.array([int(K.iz_id(i).barva) for i in ids], np.uint8),
        card_st=np.array([K.iz_id(i).st for i in ids], np.uint8),
        card_vrednost=np.array([K.iz_id(i).vrednost() for i in ids], np.uint8),
        card_roundtrip=np.array([K.iz_id(i).v_id() for i in ids], np.uint8),
    )
    print("units.npz")


def main():
    os.makedirs(OUT, exist_ok=True)
    gen_kat()
    gen_units()
    gen_auction()
    gen_traces()
    gen_full()


if __name__ == "__main__":
    main()
