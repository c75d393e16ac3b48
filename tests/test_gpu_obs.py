"""GPU: observation expansion (tarok_obs_expand) against observations frozen from the REAL reference's
Nevronski_igralec.stanje_v_vektor_rek_navadna (tests/golden/obs.npz, made by oracle/make_golden_obs.py).
Bit-exact: every entry of every input array is 0/1."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_observations_match_reference(golden):
    import torch
    from tarok_b200.env import TarokEnv
    g = golden("obs.npz")
    n = len(g["perm"])
    env = TarokEnv(n, history=True)
    env.set_deals(g["perm"])
    env.force_contract(g["contract"], g["declarer"], g["king"])
    env.exchange(g["group"], g["discard_mask"])
    assert env.errors() == 0
    idx = g["obs_index"]                                   # (game, play, seat, T, kind, packed bytes)
    row_of = {(int(r[0]), int(r[1])): k for k, r in enumerate(idx)}
    off, bits = g["obs_offset"], g["obs_bits"]
    checked = 0
    for t in range(48):
        kinds, rows = env.obs_shape()
        kinds, rows = kinds.cpu().numpy(), rows.cpu().numpy()
        live = kinds != 255
        for kind in range(4):
            for T in np.unique(rows[live & (kinds == kind)]):
                sel = np.nonzero(live & (kinds == kind) & (rows == T))[0].astype(np.int32)
                arrs, ok = env.obs_expand(kind, int(T), sel)
                assert ok.cpu().numpy().all()
                arrs = [a.cpu().numpy() for a in arrs]
                for j, game in enumerate(sel):
                    k = row_of[(int(game), t)]
                    assert idx[k][3] == T and idx[k][4] == kind, (game, t)
                    flat = np.concatenate([a[j].reshape(-1) for a in arrs])
                    assert set(np.unique(flat)) <= {0.0, 1.0}
                    want = np.unpackbits(bits[off[k]:off[k + 1]])[:flat.size]
                    assert (flat.astype(np.uint8) == want).all(), (int(game), t, kind, int(T))
                    checked += 1
        # the reference chose its card from the (tie-prone) fake net output q: the device selection must agree
        live_games = np.nonzero(live)[0].astype(np.int32)
        if len(live_games):
            qrows = np.stack([g["q"][row_of[(int(game), t)]] for game in live_games])
            cards, qmax = env.select_action(torch.from_numpy(qrows), live_games)
            got = cards[:n].cpu().numpy()
            assert (got[live_games] == g["card"][live_games, t]).all(), t
            want_q = np.array([g["qmax"][row_of[(int(game), t)]] for game in live_games], np.float32)
            assert (qmax.cpu().numpy()[live_games] == want_q).all()
            env.step(cards)
    assert env.errors() == 0
    assert checked == len(idx)
    # ---- replay samples (Igralec.py:387-446): targets dy and the observation "as of play t"
    env.score()
    dy, seat, rows = env.targets(final_reword_factor=0.1)
    dy, seat, rows = dy.cpu().numpy(), seat.cpu().numpy(), rows.cpu().numpy()
    want = g["dy"]
    have = ~np.isnan(want[:, :, 0])
    played = g["card"] != 0xFF
    assert (seat[played] != 0xFF).all() and (seat[~played] == 0xFF).all() and (dy[~played] == 0).all()
    assert have.sum() > 2000
    # the only floating-point arithmetic on the path: value + 0.1 * next.  The reference does it in float64 (numpy), the
    # kernel in fp32 -> tolerance 1e-4 absolute (values are |x| <= 200); the -70 (illegal card) pattern must be exact
    assert np.allclose(dy[have], want[have], rtol=0.0, atol=1e-4)
    assert np.array_equal(dy[have] == -70.0, want[have] == -70.0)
    for (game, t), k in list(row_of.items())[::7]:
        kind, T = int(idx[k][4]), int(idx[k][3])
        assert rows[game, t] == T and seat[game, t] == idx[k][2]
        arrs, ok = env.obs_expand(kind, T, np.array([game], np.int32), play=t)
        assert bool(ok.all())
        flat = np.concatenate([a[0].cpu().numpy().reshape(-1) for a in arrs]).astype(np.uint8)
        ref = np.unpackbits(bits[off[k]:off[k + 1]])[:flat.size]
        assert (flat[:-54] == ref[:-54]).all(), (game, t)   # everything but the legal-mask vector (not a net input)
    # a bucket that does not match is reported, not silently filled
    arrs, ok = env.obs_expand(1, 8, np.arange(4, dtype=np.int32))
    assert not ok.cpu().numpy().any() and float(arrs[0].abs().sum()) == 0.0
    env.close()


def test_epsilon_greedy_explores_uniformly_over_legal_cards():
    import torch
    from tarok_b200.env import TarokEnv
    n = 200000
    env = TarokEnv(n, seed=12, history=True)
    env.setup_synth(0)                                   # Klop: seat 0 leads with 12 cards (11 legal if it holds the pagat)
    q = torch.zeros((n, 54), dtype=torch.float32, device="cuda")
    greedy, _ = env.select_action(q, random_card=0.0)
    mixed, _ = env.select_action(q, random_card=0.5)
    mask = env.mask[:n].cpu().numpy().view(np.uint64)
    gc, mc = greedy[:n].cpu().numpy(), mixed[:n].cpu().numpy()
    assert ((mask >> gc.astype(np.uint64)) & np.uint64(1)).all() and ((mask >> mc.astype(np.uint64)) & np.uint64(1)).all()
    lowest = np.array([int(x & -x).bit_length() - 1 for x in mask.astype(object)], np.uint8)
    assert (gc == lowest).all()                          # all-equal q: first card in mozne order = lowest id when leading
    frac_changed = (mc != gc).mean()                     # explore w.p. 0.5, then uniform over ~11.8 legal cards
    assert 0.42 < frac_changed < 0.49
    env.close()


def test_observations_need_history():
    from tarok_b200.env import TarokEnv
    env = TarokEnv(64)
    with pytest.raises(ValueError):
        env.obs_expand(0, 8)
    env.close()


def test_hand_and_exchange_observations_and_exchange_decision(golden):
    """pripavi_licitiram / menjaj_talon_v_vektor / menjaj_iz_talona against the reference (obs.npz)."""
    import torch
    from tarok_b200.env import TarokEnv
    g = golden("obs.npz")
    n = len(g["perm"])
    env = TarokEnv(n, history=True)
    env.set_deals(g["perm"])
    hands = env.obs_hands().cpu().numpy()
    for s in range(4):                                   # one-hot of perm[12s:12s+12]
        want = np.zeros((n, 54), np.float32)
        np.put_along_axis(want, g["perm"][:, 12 * s:12 * s + 12].astype(np.int64), 1.0, axis=1)
        assert (hands[:, s] == want).all()
    env.force_contract(g["contract"], g["declarer"], g["king"])
    ex = np.nonzero((g["contract"] >= 1) & (g["contract"] <= 6))[0].astype(np.int32)
    (hand, talon, game), ok = env.obs_exchange(ex)
    assert ok.cpu().numpy().all()
    flat = torch.cat([hand.flatten(1), talon.flatten(1), game.flatten(1)], dim=1).cpu().numpy().astype(np.uint8)
    want = np.unpackbits(g["exch_obs_bits"][ex], axis=1)[:, :393]
    assert (flat == want).all()
    _, ok_all = env.obs_exchange()                       # games without an exchange are flagged, not filled
    assert (ok_all.cpu().numpy() == ((g["contract"] >= 1) & (g["contract"] <= 6))).all()
    group, discard = env.select_exchange(torch.from_numpy(g["exch_p"][ex]), ex)
    assert (group.cpu().numpy()[ex] == g["group"][ex]).all()
    assert (discard.cpu().numpy().view(np.uint64)[ex] == g["discard_mask"][ex]).all()
    env.exchange(group, discard)
    assert env.errors() == 0
    # exploring decisions stay valid exchanges
    env2 = TarokEnv(50000, seed=4, history=True)
    env2.deal(); env2.force_contract_synth(16)
    gr, di = env2.select_exchange(torch.zeros((50000, 60)), random_card=1.0)
    env2.exchange(gr, di)
    assert env2.errors() <= 2                            # only hands with fewer than k discardable cards (Q19)
    assert len(torch.unique(gr)) >= 2
    env.close(); env2.close()


def test_observation_invariants_at_config4_size():
    """65,536 envs, random play: structural invariants of the expanded observations at every 7th decision."""
    import torch
    from tarok_b200.env import TarokEnv
    n = 65536
    env = TarokEnv(n, seed=9, history=True)
    env.setup_synth(17)                                   # all contracts
    for t in range(48):
        if t % 7 == 3:
            kinds, rows = env.obs_shape()
            live = kinds != 255
            key = kinds.to(torch.int32) * 64 + rows.to(torch.int32)
            for k in torch.unique(key[live]).tolist():
                kind, T = k // 64, k % 64
                sel = torch.nonzero(key == k).flatten().to(torch.int32)
                arrs, ok = env.obs_expand(kind, T, sel)
                assert bool(ok.all())
                opp, hand, mozne = arrs[0], arrs[2] if kind == 1 else arrs[1], arrs[-1]
                assert bool((opp.sum(dim=(2, 3)) <= 1).all())                    # at most one card per row
                own_rows = hand.sum(dim=2) > 0
                assert bool(((opp.sum(dim=(2, 3)) > 0) & own_rows).sum() == 0)    # a row is an opponent's play or mine
                assert bool(((opp.sum(dim=(2, 3)) > 0) | own_rows).sum(dim=1).eq(t).all())   # t plays so far
                assert bool((hand.sum(dim=2)[own_rows] <= 12).all())
                legal = env.mask[:n][sel.long()]
                bits = ((legal.unsqueeze(1) >> torch.arange(54, device="cuda")) & 1).float()
                assert bool((mozne == bits).all())
        env.step_random(1)
    env.close()
