"""TEST INFRASTRUCTURE ONLY -- freezes CALLBACK TRANSCRIPTS of the real reference into tests/golden/transcripts.json.

    python -m oracle.make_golden_transcripts          (build container: needs /root/reference)

For every game the unmodified engine is driven with oracle/transcript.py's TranscriptPlayer (built on the reference's own
Igralec base class); what is frozen is, per game, the sha256 digest of the full event list (name, id_igre and every argument
of each of the callbacks, in order), the number of events, the scores -- and the full event list of the first games of each
group, so a mismatch can be read.  The GPU test (tests/test_gpu_transcripts.py) replays the same deals with the same
deterministic decision rule through tarok_b200's Igra / Klop / Berac / Navadna_igra / Tarok and compares.

Groups: 200 forced games per contract (all ten: the 5-card rezultat_stiha of Klop, poglej_karte_odprtega_beraca of Odprti_berac,
the talon callbacks of the six exchanging contracts, Solo_brez without any), 300 full Igra.start() games with a Bot-like
intent mix (licitiram / izberi_barvo_kralja / konec_licitiranja), and two Tarok.paralel_start batches of 96 games whose
phase structure (what happens between consecutive predict_* calls) includes the Solo_brez one-step lead.
"""
from __future__ import annotations

import json
import os
import random

from . import transcript as T

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "transcripts.json")
KEEP_FULL = 2          # games per group whose whole event list is stored


def main():
    eng = T.reference_engine()
    rng = random.Random(20261019)
    out = {"seed": 4711, "forced": [], "full": [], "paralel": []}
    per_contract = {}
    for c in range(10):
        done = 0
        while done < 200:
            perm = list(range(54)); rng.shuffle(perm)
            d = 0 if c == 0 else rng.randrange(4)
            k = rng.randrange(4) if 1 <= c <= 3 else 7
            try:
                ev, sc = T.run_forced(eng, perm, c, d, k, out["seed"])
            except ValueError:                      # fewer than k discardable cards (Q19)
                continue
            row = {"perm": perm, "contract": c, "declarer": d, "king": k, "scores": sc, "n": len(ev), "digest": T.Journal.digest(ev)}
            if done < KEEP_FULL:
                row["events"] = ev
            out["forced"].append(row)
            done += 1
        per_contract[c] = done
    hist = [0] * 10
    while len(out["full"]) < 300:
        perm = list(range(54)); rng.shuffle(perm)
        idx = T.bot_like_intents(rng)
        try:
            ev, sc = T.run_full(eng, perm, idx, out["seed"])
        except ValueError:
            continue
        c = next(e for e in ev if e[0] == "konec_licitiranja")[3] // 10
        hist[c] += 1
        row = {"perm": perm, "intent": idx, "scores": sc, "n": len(ev), "digest": T.Journal.digest(ev), "contract": c}
        if len(out["full"]) < KEEP_FULL:
            row["events"] = ev
        out["full"].append(row)
    for batch in range(2):
        n = 96
        while True:
            perms, idxs = [], []
            for i in range(n):
                perm = list(range(54)); rng.shuffle(perm)
                perms.append(perm)
                idxs.append(T.bot_like_intents(rng) if i % 3 else [rng.choice([0, 0, 17, 17, 16, 13]) for _ in range(4)])
            try:
                j, rez = T.run_paralel(eng, perms, idxs, out["seed"])
            except ValueError:
                continue
            break
        phases = j.phases()
        contracts = [next(e for e in j.per_game[i] if e[0] == "konec_licitiranja")[3] // 10 for i in range(n)]
        out["paralel"].append({
            "perms": perms, "intent": idxs, "rezultati": rez, "contracts": contracts,
            "game_digests": [T.Journal.digest(j.per_game[i]) for i in range(n)],
            "phase_digests": [T.Journal.digest(sorted([[g, ev] for g, ev in ph.items()])) for ph in phases],
            "phase_games": [sorted(ph) for ph in phases][:6],
            "n_phases": len(phases)})
        print("paralel batch", batch, "contracts", [contracts.count(c) for c in range(10)], "phases", len(phases))
    with open(OUT, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("transcripts.json: forced", per_contract, "full", hist, "bytes", os.path.getsize(OUT))


if __name__ == "__main__":
    main()
