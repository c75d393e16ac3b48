"""Direct large-N parity GPU -> reference (BASELINE config 1 "exactly": 10,000 Klop deals; 2,000 deals of each other contract).

tests/golden/gpu_traces.npz holds games PLAYED BY THE CUDA KERNELS on a B200 (tools/export_gpu_traces.py): Philox deals,
uniform-random legal-move players, per play the card and a digest of the legal mask the device showed.
* build container (`reference` marker): every game is replayed through the imported, unmodified Python reference -- the
  recorded cards are teacher-forced, the reference's own `mozne` at every play, its trick winners and its scores must equal
  what the GPU produced;
* everywhere (CPU): the same replay through the C oracle;
* GPU box (`gpu` marker): the kernels regenerate the traces from the same seed and must reproduce the committed file byte
  for byte, which ties the CURRENT kernels to the replayed games."""
import hashlib
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "golden", "gpu_traces.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="tests/golden/gpu_traces.npz not generated yet")


@pytest.fixture(scope="module")
def tr():
    return dict(np.load(PATH))


def _digest(masks):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(masks, np.uint64).tobytes()).digest()[:16], np.uint8)


def test_gpu_traces_replay_through_the_c_oracle(tr, oracle):
    n = len(tr["contract"])
    assert n == 28000 and (tr["contract"][:10000] == 0).all()                       # config 1 exactly + 9 x 2,000
    ok = tr["err"] == 0
    rep = oracle.replay(tr["perm"], tr["contract"], tr["declarer"], tr["king"], tr["group"], tr["discard"], tr["card"],
                        want_state=False)
    assert (rep["err"][ok] == 0).all()
    assert (rep["scores"][ok] == tr["scores"][ok]).all()
    assert (rep["plays"][ok] == tr["plays"][ok]).all()
    played = tr["card"] != 0xFF
    assert (rep["seat"][played & ok[:, None]] == tr["seat"][played & ok[:, None]]).all()
    tricks = (tr["plays"] // 4)
    for k in range(12):
        sel = ok & (tricks > k)
        assert (rep["winner"][sel, k] == tr["winner"][sel, k]).all(), k
    masks = np.where(played, rep["mask"], 0).astype(np.uint64)
    dig = np.stack([_digest(masks[i]) for i in range(n)])
    assert (dig[ok] == tr["mask_digest"][ok]).all()
    for c in range(10):                                                              # and the masks stored whole
        a = int(tr["first"][c])
        m = tr["mask_full"][c]
        assert (np.where(played[a:a + len(m)], m, 0) == masks[a:a + len(m)])[ok[a:a + len(m)]].all()


def _replay_chunk(args):
    """Worker: replays games [a, b) through the reference; returns the list of mismatching game indices."""
    a, b = args
    from oracle import ref_harness as H
    tr = dict(np.load(PATH))
    bad = []

    class Forced(H.Policy):
        def __init__(self, i):
            super().__init__()
            self.i, self.t = i, 0

        def group(self, seat, kupcki):
            return int(tr["group"][self.i])

        def discards(self, seat, mozno, k):
            want = int(tr["discard"][self.i])
            out = [c for c in mozno if (want >> c.v_id()) & 1]
            assert len(out) == k and sum(1 << c.v_id() for c in out) == want
            return out

        def card(self, seat, stih, mozne, zgodovina):
            c = int(tr["card"][self.i][self.t])
            self.t += 1
            return next(k for k in mozne if k.v_id() == c)          # StopIteration = the GPU played a card the reference forbids

    for i in range(a, b):
        if tr["err"][i]:
            continue
        try:
            rec, players = H.run_forced(tr["perm"][i], int(tr["contract"][i]), int(tr["declarer"][i]), int(tr["king"][i]), Forced(i))
        except Exception as ex:                                     # noqa: BLE001 - any divergence is a failure of this game
            bad.append((i, repr(ex)[:80]))
            continue
        p = int(tr["plays"][i])
        masks = np.zeros(48, np.uint64)
        masks[:p] = np.array(rec.masks, np.uint64)
        good = (len(rec.cards) == p and list(tr["seat"][i][:p]) == rec.seats and list(tr["scores"][i]) == rec.scores
                and list(tr["winner"][i][:p // 4]) == rec.winners and (_digest(masks) == tr["mask_digest"][i]).all())
        if not good:
            bad.append((i, "differs"))
    return bad


@pytest.mark.reference
def test_gpu_traces_replay_through_the_python_reference(tr):
    """Every GPU-played game through the imported anzeA/Tarok engine (per-contract constructors, deal injection)."""
    import multiprocessing as mp
    n = len(tr["contract"])
    workers = min(8, os.cpu_count() or 1)
    step = (n + 8 * workers - 1) // (8 * workers)
    chunks = [(a, min(n, a + step)) for a in range(0, n, step)]
    with mp.get_context("spawn").Pool(workers) as pool:
        bad = [x for part in pool.map(_replay_chunk, chunks) for x in part]
    assert not bad, bad[:5]


@pytest.mark.gpu
def test_current_kernels_regenerate_the_committed_traces(tr, tmp_path):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
    import export_gpu_traces as X
    got = X.export(str(tmp_path / "again.npz"))
    for k in ("perm", "contract", "declarer", "king", "group", "discard", "seat", "card", "winner", "scores", "plays", "err",
              "mask_digest", "mask_full"):
        assert np.array_equal(np.asarray(got[k]), tr[k]), k
