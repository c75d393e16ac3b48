"""CPU: the reference-API value types and host logic (tarok_b200.karte / igralec / Partije.licitiraj)
against vectors frozen from the real reference (tests/golden/units.npz, auction_*.npz)."""
import pytest

from tarok_b200 import Barva, Bot_igralec, Igralec, Karta, Roka, Tip_igre
from tarok_b200.karte import maska_iz_kart


def test_karta_tables(golden):
    u = golden("units.npz")
    for i in range(54):
        k = Karta.iz_id(i)
        assert int(k.barva) == u["card_barva"][i] and k.st == u["card_st"][i]
        assert k.vrednost() == u["card_vrednost"][i]
        assert k.v_id() == i == u["card_roundtrip"][i]
        assert k == Karta(Barva(int(u["card_barva"][i])), int(u["card_st"][i]))
    cards = [Karta.iz_id(i) for i in range(54)]
    assert sorted(cards, reverse=True)[::-1] == cards                  # __lt__ orders by (suit, rank) = id
    assert str(Karta(Barva.KARA, 8)) == "KARA_KR" and str(Karta(Barva.TAROK, 21)) == "TAROK_21"
    with pytest.raises(TypeError):
        hash(cards[0])                                                 # __eq__ without __hash__, as upstream


def test_tip_igre_values():
    assert [int(t) for t in Tip_igre] == [-10, 0, 10, 20, 30, 40, 50, 60, 70, 80, 90]
    assert Tip_igre.Solo_brez.code == 8 and Tip_igre.iz_kode(7) is Tip_igre.Berac


def test_roka_counting_and_discards(golden):
    u = golden("units.npz")
    for ids, n, want in zip(u["pile_ids"][:800], u["pile_len"][:800], u["pile_val"][:800]):
        assert Roka.prestej([Karta.iz_id(int(i)) for i in ids[:n]]) == want
    assert Roka.prestej([Karta.iz_id(i) for i in range(54)]) == 70
    for hand, want in zip(u["hand"][:3000], u["mozno_zalozit"][:3000]):
        r = Roka.iz_maske(int(hand))
        assert maska_iz_kart(r.mozno_zalozit()) == int(want)
        assert r.maska() == int(hand) and len(r) == bin(int(hand)).count("1")
    r = Roka([Karta.iz_id(i) for i in (40, 3, 33, 1)])
    assert [k.v_id() for k in r] == [1, 3, 33, 40]                      # sorted per suit at construction
    r.dodaj_karte([Karta.iz_id(0), Karta.iz_id(32)])
    assert [k.v_id() for k in r] == [1, 3, 0, 33, 40, 32]               # pick-ups are appended unsorted
    r.igraj_karto(Karta.iz_id(3))
    assert Karta.iz_id(3) not in r and Karta.iz_id(0) in r
    assert [len(s) for s in Roka.tri_po_tri(list(range(8)))] == [3, 3, 2]


def test_bid_filter():
    p = Igralec("x")
    N, K, T, D = Tip_igre.Naprej, Tip_igre.Klop, Tip_igre.Tri, Tip_igre.Dve
    assert p.licitiram(D, T, 0) == D and p.licitiram(T, T, 0) == N and p.licitiram(T, T, 0, prednost=True) == T
    assert p.licitiram(N, N, 0, K) == K and p.licitiram(T, D, 0, D) == D and p.licitiram(T, D, 0) == N


class _Fixed(Igralec):
    def __init__(self, tip):
        super().__init__()
        self.tip = tip

    def licitiram(self, min_igra, id_igre, obvezno=None, prednost=False):
        self.tip = super().licitiram(self.tip, min_igra, id_igre, obvezno, prednost)
        return self.tip


class _Scripted(Igralec):
    def __init__(self, seq):
        super().__init__()
        self.seq = seq

    def licitiram(self, min_igra, id_igre, obvezno=None, prednost=False):
        return super().licitiram(Tip_igre(10 * int(self.seq.pop(0))), min_igra, id_igre, obvezno, prednost)


def test_host_auction_state_machine_fixed_intents(golden):
    from tarok_b200.igra import Partije
    a = golden("auction_fixed.npz")
    for i in range(0, len(a["intents"]), 5):
        players = [_Fixed(Tip_igre(10 * int(c))) for c in a["intents"][i]]
        kdo, tip = Partije.licitiraj(players, 0)
        assert (kdo, int(tip) // 10) == (int(a["declarer"][i]), int(a["contract"][i]))


def test_host_auction_state_machine_scripted(golden):
    from tarok_b200.igra import Partije
    a = golden("auction_scripted.npz")
    for i in range(0, len(a["draws"]), 3):
        seq = list(a["draws"][i]) + [-1] * 8
        players = [_Scripted(seq) for _ in range(4)]      # one shared call-ordered script
        kdo, tip = Partije.licitiraj(players, 0)
        assert (kdo, int(tip) // 10) == (int(a["declarer"][i]), int(a["contract"][i]))
        assert 24 - len(seq) == int(a["calls"][i])


def test_bot_is_device_capable():
    assert Bot_igralec.device_policy == "bot" and Igralec.device_policy is None
