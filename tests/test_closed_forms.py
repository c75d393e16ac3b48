"""CPU checks of closed forms the device code uses in place of the reference's loops."""
import numpy as np


def test_trick_winner_closed_form():
    """The device computes the trick winner as an arg-max over (eligible ? id : none) instead of the reference's sequential
    scan (pobere_stih / primerjaj_karti, Navadna_igra.py:143-156): identical on every ordered trick of four distinct cards."""
    idx = np.indices((54, 54, 54, 54), dtype=np.int8).reshape(4, -1).T
    ok = ((idx[:, 0] != idx[:, 1]) & (idx[:, 0] != idx[:, 2]) & (idx[:, 0] != idx[:, 3]) & (idx[:, 1] != idx[:, 2])
          & (idx[:, 1] != idx[:, 3]) & (idx[:, 2] != idx[:, 3]))
    c = idx[ok].astype(np.int32)
    assert len(c) == 54 * 53 * 52 * 51

    def suit(x):
        return np.where(x >= 32, 4, x >> 3)
    best, w = c[:, 0].copy(), np.zeros(len(c), np.int32)
    for i in (1, 2, 3):                                            # the scan of the reference
        ch = c[:, i]
        beats = np.where(suit(best) == suit(ch), ch > best, suit(ch) == 4)
        best, w = np.where(beats, ch, best), np.where(beats, i, w)
    lead = c[:, 0]
    m = ((lead + 1) << 2)
    for i in (1, 2, 3):                                            # tarok_rules.cuh: trick_winner
        ci = c[:, i]
        m = np.maximum(m, np.where((ci >= 32) | ((ci ^ lead) < 8), ((ci + 1) << 2) | i, 0))
    assert (w == (m & 3)).all()


def _points(c):
    """Per-card points of Roka.vrednost_stiha (Roka.py:76-91): suit ranks 1-4 -> 1, J 2, C 3, Q 4, K 5; taroks 1, trula 5."""
    c = np.asarray(c)
    return np.where(c < 32, np.where((c & 7) < 4, 1, (c & 7) - 2), np.where((c == 32) | (c >= 52), 5, 1))


def _all_tricks():
    c = np.arange(54, dtype=np.uint32)
    g = np.stack(np.meshgrid(c, c, c, c, indexing="ij"), -1).reshape(-1, 4)
    return g, (g[:, 0] | (g[:, 1] << 6) | (g[:, 2] << 12) | (g[:, 3] << 18)).astype(np.uint32)


def test_trick_points_closed_form():
    """tarok_rules.cuh trick_points: the card points of a trick computed on the four packed 6-bit ids at once; identical to the
    per-card table on all 54^4 tuples (distinct or not)."""
    g, f = _all_tricks()
    u = np.uint32
    B5 = u(0x820820)
    hr = ~f & (f << u(3)) & B5
    v = ((f & u(0x0C30C3)) + u(0x041041)) & ((hr >> u(5)) * u(7))
    a = f & (f << u(1)) & (f << u(3))
    nz = (f & u(0x7DF7DF)) + u(0x7DF7DF)
    tr = f & (a | ~nz) & B5
    pc = sum(((tr >> u(6 * j + 5)) & u(1)) for j in range(4))
    got = u(4) + (((v * u(0x041041)) >> u(18)) & u(0x3F)) + u(4) * pc
    assert (got == _points(g).sum(1)).all()
    assert int(got.max()) == 20                                   # fits the 5 bits (24-28) of the trick-log entry


def test_trick_has_closed_form():
    """tarok_rules.cuh trick_has: zero-field detection on the XOR with the broadcast id."""
    g, f = _all_tricks()
    u = np.uint32
    for kid in (7, 15, 23, 31, 0, 32, 53):
        x = f ^ (u(kid) * u(0x041041))
        z = ((x & u(0x7DF7DF)) + u(0x7DF7DF)) | x
        assert ((((~z) & u(0x820820)) != 0) == (g == kid).any(1)).all()


def test_card_points_sum_to_the_reference_total():
    """Full deck = 106 card points -> 70 after 3-card grouping (SURVEY a5 probe): 106 - 2 * 18."""
    assert int(_points(np.arange(54)).sum()) == 106


def test_batched_draws_are_exactly_uniform():
    """philox.cuh bdraw: k draws from one word = Lemire's multiply-shift for the product bound read in mixed radix.  On a
    16-bit word (same algebra, exhaustive): every accepted tuple has exactly floor(2^16 / B) preimages."""
    L = 16
    for bounds in [(6, 5, 4), (54, 53), (3, 4, 4), (18, 18, 18), (12, 11, 10)]:
        B = int(np.prod(bounds))
        x = np.arange(1 << L, dtype=np.uint64)
        digits = []
        for n in bounds:
            m = x * np.uint64(n)
            digits.append(m >> np.uint64(L))
            x = m & np.uint64((1 << L) - 1)
        ok = x >= np.uint64((1 << L) % B)
        code = np.zeros(1 << L, np.int64)
        for d, n in zip(digits, bounds):
            code = code * n + d.astype(np.int64)
        counts = np.bincount(code[ok], minlength=B)
        assert len(counts) == B and (counts == (1 << L) // B).all()
