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
