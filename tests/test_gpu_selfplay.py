"""GPU: BASELINE config 4 pipeline -- restated policy nets + device env / observations / action selection."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_neural_selfplay_rollout_is_legal_and_complete():
    import torch
    from tarok_b200.samoigra import Samoigra
    torch.manual_seed(0)
    n = 3000
    for eps in (0.0, 0.3):
        s = Samoigra(n, seed=77, random_card=eps)
        st, ms = s.odigraj(first_game_id=0, meri=True)
        assert st[18] + st[20] == n                       # every deal finished or was flagged ...
        assert st[20] <= 3 and st[21] == st[20]           # ... and only for Q19 hands (no illegal card ever selected)
        assert st[8:18].sum() == st[18]
        assert 10 * n < st[19] <= 48 * n                  # Berac games stop early
        assert set(ms) == {"env", "obs", "forward", "select"}
        if eps > 0:
            assert (st[8:18] > 0).sum() >= 5              # exploring bids reach most contracts
        s.zapri()


def test_nets_follow_the_reference_call_contract():
    import torch
    from tarok_b200.mreze import ustvari_mreze
    nets = ustvari_mreze(torch.device("cuda"))
    B, T = 7, 16
    z = lambda *s: torch.zeros(s, device="cuda")
    for m in nets.values():
        m.eval()
    assert nets["Navadna_igra"]([z(B, T, 3, 54), z(B, 4), z(B, T, 54), z(B, 6, 55), z(B, 4), z(B, 54)]).shape == (B, 54)
    assert nets["Solo"]([z(B, T, 3, 54), z(B, T, 54), z(B, 6, 55), z(B, 4), z(B, 54)]).shape == (B, 54)
    assert nets["Klop"]([z(B, T, 3, 54), z(B, T, 54), z(B, 54)]).shape == (B, 54)
    assert nets["Berac"]([z(B, T, 3, 54), z(B, T, 54), z(B, 4)]).shape == (B, 54)
    assert nets["Vrednotenje_roke"](z(B, 54)).shape == (B, 18)
    assert nets["Zalaganje"]([z(B, 54), z(B, 54, 6), z(B, 15)]).shape == (B, 60)


def test_replay_samples_and_one_training_pass():
    import torch
    from tarok_b200.samoigra import Samoigra
    torch.manual_seed(1)
    n = 2048
    s = Samoigra(n, seed=5, random_card=0.2)
    st, _ = s.odigraj()
    total = 0
    for ime, T, stanje, dy in s.vzorci():
        B = dy.shape[0]
        assert stanje[0].shape == (B, T, 3, 54) and dy.shape == (B, 54) and T % 8 == 0
        assert bool(((dy == -70) | (dy > -70)).all())
        total += B
    assert total == st[19]                                   # one sample per card play (Igralec.py:416)
    before = [p.detach().clone() for p in s.mreze["Navadna_igra"].parameters()]
    loss = s.nauci()
    assert loss and all(np.isfinite(v) for v in loss.values())
    after = list(s.mreze["Navadna_igra"].parameters())
    assert any(not torch.equal(a, b) for a, b in zip(before, after))
    s.zapri()
