"""TEST INFRASTRUCTURE ONLY -- freezes OBSERVATIONS of the real reference into tests/golden/obs.npz.

    python -m oracle.make_golden_obs

The observation encoder is the reference's own ``Nevronski_igralec.stanje_v_vektor_rek_navadna``
(Igralec.py:453-533), called unmodified through ``pripravi_igraj_karto`` (Igralec.py:312-314) at every
decision of every seat.  Only the *decisions* are scripted (seeded random legal cards / groups / discards);
the nets are never built (``ignor_models=True``).  Per observation the fixture stores the bit-packed
concatenation of the reference's input arrays (all entries are 0/1) in the reference's list order.
"""
from __future__ import annotations

import os
import random

import numpy as np

from . import ref_harness as H

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
NET_TYPES = {"Klop": 0, "Navadna_igra": 1, "Solo": 2, "Berac": 3}     # Nevronski_igralec.Tipi_NN values


def make_obs_player(ref, rng, log):
    NN = ref.Igralec.Nevronski_igralec
    Base = ref.Igralec.Igralec

    class ObsPlayer(NN):
        def __init__(self, seat):
            super().__init__(ignor_models=True, ime="seat%d" % seat)
            self.seat = seat
            self.game_row = 0

        # ---- decisions are scripted; everything else is the reference's own code ----
        def pripravi_izbral_iz_talona(self, talon, st_kart, id_igre):
            NN.pripravi_izbral_iz_talona(self, talon, st_kart, id_igre)      # reference: menjaj_talon_v_vektor (Igralec.py:357-360)
            st = self.zalaganje2tocke[id_igre][0]
            flat = np.concatenate([np.asarray(a, np.float64).reshape(-1) for a in st[:3]])
            assert set(np.unique(flat)) <= {0.0, 1.0}
            log["xobs"] = np.packbits(flat.astype(np.uint8))

        def predict_izberi_iz_talona(self):
            self.predict_queue["Zalaganje"] = []

        def predict_igraj_karto(self):
            for k in self.predict_queue:
                self.predict_queue[k] = []

        def menjaj_iz_talona(self, kupcki, st_kart, id_igre):
            # the REFERENCE's own decision code (Igralec.py:365-385) on a fake, tie-prone network output
            import torch
            # card scores are distinct (np.argsort's order among equal values depends on the numpy build: SIMD sorts are
            # not stable), group scores are tie-prone (np.argmax = first maximum is well defined)
            cards_part = list(range(54))
            rng.shuffle(cards_part)
            pvec = np.array(cards_part + [rng.randrange(3) for _ in range(6)], np.float32)
            self.predicted_resoult["Zalaganje"][id_igre] = torch.from_numpy(pvec)
            self.random_card = 0.0
            g = int(NN.menjaj_iz_talona(self, kupcki, st_kart, id_igre))
            log["group"], log["discard"] = g, H.cards_to_mask(self.zalaganje2tocke[id_igre][1])
            log["xp"] = pvec
            return g

        def pripravi_igraj_karto(self, karte_na_mizi, mozne, zgodovina, id_igre):
            super().pripravi_igraj_karto(karte_na_mizi, mozne, zgodovina, id_igre)   # builds self.stanje[id]
            arrs = self.stanje[id_igre]
            flat = np.concatenate([np.asarray(a, np.float64).reshape(-1) for a in arrs])
            assert set(np.unique(flat)) <= {0.0, 1.0}
            log["obs"].append(dict(seat=self.seat, T=int(arrs[0].shape[1]), kind=NET_TYPES[self.tip_igre[id_igre]],
                                   shapes=[tuple(a.shape[1:]) for a in arrs], bits=np.packbits(flat.astype(np.uint8))))

        def igraj_karto(self, karte_na_mizi, mozne, zgodovina, id_igre):
            # the card is chosen by the REFERENCE's own action selection (Igralec.py:344-355) from a fake network
            # output with few distinct values, so that argmax ties (broken by the order of `mozne`) are exercised
            import torch
            q = np.array([rng.randrange(3) for _ in range(54)], np.float32)
            self.predicted_resoult[self.tip_igre[id_igre]][id_igre] = torch.from_numpy(q)
            self.random_card = 0.0
            karta = NN.igraj_karto(self, karte_na_mizi, mozne, zgodovina, id_igre)
            log["cards"].append(karta.v_id())
            log["q"].append(q)
            log["qmax"].append(float(self.next_Q_max[id_igre]))
            return karta

        # rezultat_stiha / rezultat_igre are the REFERENCE's own (Igralec.py:387-446): they build the replay samples
        def rezultat_igre(self, st_tock, povzetek_igre, id_igre):
            try:
                NN.rezultat_igre(self, st_tock, povzetek_igre, id_igre)
            except KeyError:          # Odprti_berac is not in igra2index (Q20): the reference cannot finish such a game
                log["targets_ok"] = False

        def poglej_karte_odprtega_beraca(self, roka, id_igre):
            pass

    return ObsPlayer


def run_game(ref, rng, perm, contract, declarer, king):
    Tip = ref.Tip_igre.Tip_igre
    log = dict(obs=[], cards=[], q=[], qmax=[], group=0xFF, discard=0, xobs=None, xp=None, targets_ok=True)
    P = make_obs_player(ref, rng, log)
    players = [P(s) for s in range(4)]
    H.inject_deal(ref, perm)
    talon = ref.Igra.Igra(players).razdeli()
    name = H.CONTRACT_NAMES[contract]
    barva = ref.Karta.Barva(king) if king != H.NO_KING else None
    for p in players:      # what Igra.start does right after the auction (Igra.py:57-58)
        p.lic[0] = Tip(contract * 10)
        p.konec_licitiranja(players[declarer], Tip(contract * 10), 0, barva)
    if name == "Klop":
        g = ref.Klop.Klop(players, talon, 0)
    elif name in ("Berac", "Odprti_berac"):
        g = ref.Berac.Berac(players, players[declarer], talon, name == "Odprti_berac", 0)
    else:
        g = ref.Navadna_igra.Navadna_igra(players, Tip(contract * 10), barva, players[declarer], talon, 0)
    list(g.start())
    # the samples the reference stored: zgodovina[(tip, T)] = [(stanje, dy), ...] per player, in trick order
    log["dy"] = []
    for p in players:
        rows = []
        for key in sorted(p.zgodovina, key=lambda k: k[1]):
            rows.extend(dy for _, dy in p.zgodovina[key])
        log["dy"].append(np.array(rows, np.float32).reshape(-1, 54))
    return log


def main(per_contract=6, seed=424242):
    ref = H.load_reference()
    rng = random.Random(seed)
    games, obs_rows, blobs, qs, qmaxs, dys = [], [], [], [], [], []
    for c in range(10):
        done = 0
        while done < per_contract:
            perm = list(range(54))
            rng.shuffle(perm)
            d = 0 if c == 0 else rng.randrange(4)
            k = rng.randrange(4) if 1 <= c <= 3 else H.NO_KING
            try:
                log = run_game(ref, rng, perm, c, d, k)
            except ValueError:
                continue
            gi = len(games)
            cards = log["cards"] + [0xFF] * (48 - len(log["cards"]))
            games.append(dict(perm=perm, contract=c, declarer=d, king=k, group=log["group"], discard=log["discard"],
                              cards=cards, xobs=log["xobs"] if log["xobs"] is not None else np.zeros(50, np.uint8),
                              xp=log["xp"] if log["xp"] is not None else np.zeros(60, np.float32)))
            tgt = np.full((48, 54), np.nan, np.float32)           # row t = the dy of card play t (seat = obs seat)
            if log["targets_ok"]:
                nxt = [0, 0, 0, 0]
                for t, o in enumerate(log["obs"]):
                    seat = o["seat"]
                    if nxt[seat] < len(log["dy"][seat]):
                        tgt[t] = log["dy"][seat][nxt[seat]]
                    nxt[seat] += 1
            dys.append(tgt)
            for t, o in enumerate(log["obs"]):
                obs_rows.append((gi, t, o["seat"], o["T"], o["kind"], len(o["bits"])))
                blobs.append(o["bits"])
                qs.append(log["q"][t])
                qmaxs.append(log["qmax"][t])
            done += 1
    off = np.concatenate([[0], np.cumsum([len(b) for b in blobs])]).astype(np.int64)
    np.savez_compressed(
        os.path.join(OUT, "obs.npz"),
        perm=np.array([g["perm"] for g in games], np.uint8), contract=np.array([g["contract"] for g in games], np.uint8),
        declarer=np.array([g["declarer"] for g in games], np.uint8), king=np.array([g["king"] for g in games], np.uint8),
        group=np.array([g["group"] for g in games], np.uint8), discard_mask=np.array([g["discard"] for g in games], np.uint64),
        card=np.array([g["cards"] for g in games], np.uint8),
        exch_obs_bits=np.array([g["xobs"] for g in games], np.uint8), exch_p=np.array([g["xp"] for g in games], np.float32),
        obs_index=np.array(obs_rows, np.int32),          # (game, play, seat, T, kind, packed bytes)
        obs_offset=off, obs_bits=np.concatenate(blobs),
        q=np.array(qs, np.float32), qmax=np.array(qmaxs, np.float32),
        dy=np.array(dys, np.float32))                     # [games,48,54] the reference's replay targets (NaN = none)       # fake net outputs -> reference's choice = card
    print("obs.npz: %d games, %d observations, %d packed bytes" % (len(games), len(obs_rows), off[-1]))


if __name__ == "__main__":
    main()
