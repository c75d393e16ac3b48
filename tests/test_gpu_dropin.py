"""GPU: the reference-API layer (Igra / Tarok / contract classes / Igralec callbacks) driven exactly like
the reference would be, against traces frozen from the real reference (tests/golden)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _index2igra(idx):
    from tarok_b200 import Barva, Tip_igre
    if idx == 0:
        return Tip_igre.Naprej, None
    if idx <= 12:
        return Tip_igre(10 * (1 + (idx - 1) // 4)), Barva((idx - 1) % 4)
    return Tip_igre([40, 50, 60, 70, 80][idx - 13]), None


def _make_player_class():
    from tarok_b200 import Igralec, Karta
    from tarok_b200.karte import karte_iz_maske, maska_iz_kart

    class Posnetek(Igralec):
        """Replays a golden record (keyed by id_igre -> row of the fixture) and checks what the engine shows."""

        def __init__(self, ime, g, vrstica=None):
            super().__init__(ime)
            self.g, self.vrstica = g, (vrstica or (lambda idg: idg))
            self.sedez, self.namen, self.zmage, self.koncno = {}, {}, {}, {}

        def nova_igra(self, roka, igralci, id_igre):
            super().nova_igra(roka, igralci, id_igre)
            self.sedez[id_igre] = igralci.index(self)

        def licitiram(self, min_igra, id_igre, obvezno=None, prednost=False):
            r = self.vrstica(id_igre)
            if id_igre not in self.namen:
                self.namen[id_igre] = _index2igra(int(self.g["intent"][r][self.sedez[id_igre]]))[0]
            out = super().licitiram(self.namen[id_igre], min_igra, id_igre, obvezno, prednost)
            self.namen[id_igre] = out            # Nevronski_igralec overwrites its intent (Igralec.py:304)
            return out

        def izberi_barvo_kralja(self, id_igre):
            r = self.vrstica(id_igre)
            return _index2igra(int(self.g["intent"][r][self.sedez[id_igre]]))[1]

        def menjaj_iz_talona(self, kupcki, st_kart, id_igre):
            r = self.vrstica(id_igre)
            st = int(self.g["group"][r])
            roka = self.roka[id_igre]
            roka.dodaj_karte(kupcki[st])
            zalozi = karte_iz_maske(int(self.g["discard_mask"][r]))
            assert len(zalozi) == st_kart
            self.kupcek[id_igre].extend(zalozi)
            for k in zalozi:
                roka.igraj_karto(k)
            return st

        def igraj_karto(self, stih, mozne, zgodovina, id_igre):
            r = self.vrstica(id_igre)
            t = sum(1 for kdo, _ in zgodovina if kdo is not None and not isinstance(kdo, str))
            assert int(self.g["seat"][r][t]) == self.sedez[id_igre]
            assert maska_iz_kart(mozne) == int(self.g["mask"][r][t]), (r, t)
            return super().igraj_karto(Karta.iz_id(int(self.g["card"][r][t])), id_igre)

        def rezultat_stiha(self, stih, sem_pobral, id_igre):
            if sem_pobral:
                self.zmage.setdefault(id_igre, []).append(self.sedez[id_igre])

        def rezultat_igre(self, st_tock, povzetek_igre, id_igre):
            self.koncno[id_igre] = st_tock

        def poglej_karte_odprtega_beraca(self, roka, id_igre):
            pass

    return Posnetek


def _inject(perms):
    """Deal-injection hook with the semantics of patching Igra.shuffle in the reference."""
    it = iter(perms)

    def shuffle(lst):
        lst[:] = [int(x) for x in next(it)]
    return shuffle


def test_igra_start_single_games_match_reference(golden):
    from tarok_b200 import Igra
    from tarok_b200 import igra as igra_mod
    g = golden("traces_full.npz")
    P = _make_player_class()
    rows = list(range(0, 400, 7))
    try:
        for r in rows:
            players = [P("p%d" % s, g, vrstica=lambda idg, r=r: r) for s in range(4)]
            igra_mod.shuffle = _inject([g["perm"][r]])
            out = list(Igra(players).start())
            assert len(out) == 1                                   # a single game yields exactly one item
            assert [out[0][p] for p in players] == g["scores"][r].tolist(), r
            zm = sorted((i, s) for p in players for i, s in enumerate(p.zmage.get(0, [])))
            assert len(zm) == sum(1 for w in g["winner"][r] if w != 0xFF)
            for p in players:
                assert p.koncno[0] == g["scores"][r][p.sedez[0]]
    finally:
        igra_mod.shuffle = None


def test_multi_games_generator_shape(golden):
    from tarok_b200 import Igra
    from tarok_b200 import igra as igra_mod
    g = golden("traces_full.npz")
    P = _make_player_class()
    r = int(np.nonzero(g["contract"] == 3)[0][0])                  # an Ena game: exchange + 48 plays
    players = [P("p%d" % s, g, vrstica=lambda idg: r) for s in range(4)]
    try:
        igra_mod.shuffle = _inject([g["perm"][r]])
        gen = Igra(players, multi_games=True, id=5).start()
        assert next(gen) == "Pripravljen_licitirat"
        inner = next(gen)
        items = list(inner)
        assert items[0] == "Pripravljen menjat"
        assert items[1:-1] == ["Pripravljen igrat karto"] * 48
        assert [items[-1][p] for p in players] == g["scores"][r].tolist()
    finally:
        igra_mod.shuffle = None
    with pytest.raises(Exception, match="Can not have multiple games without id"):
        Igra(players, multi_games=True, id=None)


def test_contract_classes_match_reference(golden):
    """Klop(P,talon,0) / Navadna_igra(P,tip,king,P[d],talon,0) / Berac(P,P[d],talon,odprti,0) after Igra(P).razdeli()."""
    from tarok_b200 import Barva, Berac, Igra, Klop, Navadna_igra, Tip_igre
    from tarok_b200 import igra as igra_mod
    g = golden("traces_forced.npz")
    P = _make_player_class()
    try:
        for r in range(0, 1500, 29):
            players = [P("p%d" % s, g, vrstica=lambda idg, r=r: r) for s in range(4)]
            igra_mod.shuffle = _inject([g["perm"][r]])
            talon = Igra(players).razdeli()
            assert [k.v_id() for k in talon] == g["perm"][r][48:].tolist()
            c, d, k = int(g["contract"][r]), int(g["declarer"][r]), int(g["king"][r])
            if c == 0:
                igra = Klop(players, talon, 0)
            elif c in (7, 9):
                igra = Berac(players, players[d], talon, c == 9, 0)
            else:
                igra = Navadna_igra(players, Tip_igre(10 * c), Barva(k) if k != 7 else None, players[d], talon, 0)
            pisejo = list(igra.start())[-1]
            assert [pisejo[p] for p in players] == g["scores"][r].tolist(), (r, c)
            plays = sum(1 for kdo, _ in igra.zgodovina if kdo is not None and not isinstance(kdo, str))
            assert plays == g["plays"][r]
            if c == 0:       # Klop: the talon cards enter the history as (None, card), last talon card first
                tk = [kk.v_id() for kdo, kk in igra.zgodovina if kdo is None]
                assert tk == g["perm"][r][48:][::-1].tolist()
    finally:
        igra_mod.shuffle = None


def test_tarok_paralel_start_lockstep_with_rotation(golden):
    from tarok_b200 import Tarok
    from tarok_b200 import igra as igra_mod
    g = golden("traces_full.npz")
    P = _make_player_class()
    n = 48
    players = [P("p%d" % s, g) for s in range(4)]
    try:
        igra_mod.shuffle = _inject(g["perm"][:n])
        t = Tarok(players, n)
        t.izpis = False
        t.paralel_start()
    finally:
        igra_mod.shuffle = None
    want = [0, 0, 0, 0]
    for i in range(n):
        for s in range(4):
            want[(s + i) % 4] += int(g["scores"][i][s])           # Tarok.py:34: seat s of game i is player (s+i)%4
    assert [t.rezultati[p] for p in players] == want


def test_tarok_with_four_bots_runs_on_device_and_matches_oracle(oracle):
    from tarok_b200 import Bot_igralec, Tarok
    n, seed = 50000, 31337
    bots = [Bot_igralec() for _ in range(4)]
    t = Tarok(bots, n, seed=seed)
    t.izpis = False
    t.paralel_start()
    ref = oracle.rollout(seed, 0, n, oracle.MODE_AUCTION_BOT, full=False)
    assert [t.rezultati[b] for b in bots] == ref["stats"][4:8].tolist()
    assert t.statistika[19] == ref["stats"][8]
    # contract mix of Bot bidding (SURVEY.md section 6 probe: ~15% Klop, 5% Tri, 27% Dve, 52% Ena)
    mix = t.statistika[8:18] / n
    assert abs(mix[0] - 0.148) < 0.01 and abs(mix[3] - 0.53) < 0.015


def test_host_side_bots_through_the_callback_path():
    """Bot_igralec forced through the callback protocol (host RNG): every game ends with a legal score."""
    from tarok_b200 import Bot_igralec, Tarok

    class HostBot(Bot_igralec):
        device_policy = None

    bots = [HostBot() for _ in range(4)]
    t = Tarok(bots, 32)
    t.izpis = False
    t.paralel_start()
    assert all(isinstance(v, int) for v in t.rezultati.values())
    assert all(len(b.roka[i]) == 0 for b in bots for i in range(32))     # Bot bids never reach Berac: all 48 cards played


def test_reference_exceptions_are_reproduced():
    """Illegal card -> the reference's exception text (Navadna_igra.py:125-126); four distinct names required (Tarok.py:9)."""
    from tarok_b200 import Bot_igralec, Igra, Karta, Tarok

    class Goljuf(Bot_igralec):
        device_policy = None

        def igraj_karto(self, karte_na_mizi, mozne, zgodovina, id_igre):
            held = {k.v_id() for k in self.roka[id_igre]}
            return Karta.iz_id(next(i for i in range(54) if i not in held))      # a card the player does not hold

    class Tih(Bot_igralec):
        device_policy = None

        def licitiram(self, min_igra, id_igre, obvezno=None, prednost=False):
            from tarok_b200 import Igralec, Tip_igre
            return Igralec.licitiram(self, Tip_igre.Naprej, min_igra, id_igre, obvezno, prednost)

    players = [Goljuf(), Tih(), Tih(), Tih()]
    with pytest.raises(Exception, match="Karte ne mores igarti"):
        list(Igra(players).start())
    same = [Bot_igralec(ime="a"), Bot_igralec(ime="a"), Bot_igralec(ime="b"), Bot_igralec(ime="c")]
    with pytest.raises(AssertionError):
        Tarok(same, 4)
    # everybody passes -> forehand must play Klop (Igra.py:92-94)
    quiet = [Tih() for _ in range(4)]
    t = Tarok(quiet, 8)
    t.izpis = False
    t.paralel_start()
    assert all(v <= 0 for v in t.rezultati.values())                 # Klop scores are never positive (Klop.py:36-42)


def test_bot_fast_path_keeps_its_environment_between_calls():
    """main.py builds a new Tarok(igralci, num_games) every iteration (main.py:111-116): the Bot fast path re-uses one device
    environment (re-seeded) instead of allocating a new one -- same results as before for a given seed, different deals for
    different seeds, and the memory can be handed back."""
    from tarok_b200 import Tarok, Bot_igralec
    from tarok_b200 import igra as I
    I.sprosti_okolja()
    res = []
    for seed in (5, 6, 5):
        bots = [Bot_igralec() for _ in range(4)]
        t = Tarok(bots, 4096, seed=seed)
        t.izpis = False
        t.paralel_start()
        res.append([t.rezultati[b] for b in bots])
    assert res[0] == res[2] and res[0] != res[1]
    assert len(I._OKOLJA) == 1
    env = next(iter(I._OKOLJA.values()))
    t = Tarok([Bot_igralec() for _ in range(4)], 1000, seed=5)        # another batch size: a new environment replaces the old one
    t.izpis = False
    t.paralel_start()
    assert next(iter(I._OKOLJA.values())) is not env and not env._h
    I.sprosti_okolja()
    assert not I._OKOLJA
