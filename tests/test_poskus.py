"""The reference's experiment files (scores / loss / kwargs pickles + per-player network directories, main.py:48-150) around the
device self-play loop: tarok_b200.poskus."""
import os
import pickle

import pytest
import torch


def test_experiment_directory_has_the_reference_layout(tmp_path):
    from tarok_b200 import poskus as P
    d = str(tmp_path / "poskus")
    igralci, sf, kf, lf = P.naredi_nove_igralce(d, random_card=0.25, final_reword_factor=0.3, learning_rate=0.01)
    assert [str(i) for i in igralci] == ["Igralec_1", "Igralec_2", "Igralec_3", "Igralec_4"]           # Igralec.py:118-119
    assert sorted(os.listdir(d)) == ["1", "2", "3", "4", "kwargs.pickle", "loss.pickle", "scores.pickle"]   # main.py:73-87
    assert pickle.load(open(sf, "rb")) == [] and pickle.load(open(lf, "rb")) == [] and pickle.load(open(kf, "rb")) == {}
    for i in igralci:
        i.save_models()
    assert sorted(os.listdir(os.path.join(d, "2"))) == ["Berac_A.pth", "Klop_A.pth", "Navadna_igra_A.pth", "Solo_A.pth",
                                                        "Vrednotenje_roke.pth", "Zalaganje.pth"]              # Igralec.py:802-808
    with open(kf, "wb") as f:
        pickle.dump({"final_reword_factor": 0.3, "random_card": 0.25}, f)                                 # main.py:144-147
    # a second network set left by the reference's Double_Nevronski_Igralec is tolerated; anything else is an error
    torch.save(igralci[0].models["Klop"].state_dict(), os.path.join(d, "1", "Klop_B.pth"))
    again, *_ = P.load_igralce(d)
    for a, b in zip(igralci, again):
        assert str(a) == str(b) and b.random_card == 0.25 and b.final_reword_factor == 0.3
        for k in a.models:
            for x, y in zip(a.models[k].state_dict().values(), b.models[k].state_dict().values()):
                assert torch.equal(x, y), (str(a), k)
    open(os.path.join(d, "3", "nekaj.txt"), "w").write("x")
    with pytest.raises(IOError):
        P.load_igralce(d)


@pytest.mark.gpu
def test_experiment_loop_writes_and_continues_the_pickles(tmp_path):
    from tarok_b200 import poskus as P
    d = str(tmp_path / "tek")
    scores, loss = P.main(d, iteracij=2, num_games=768, seed=5, random_card=0.1, learning_rate=0.01)
    assert len(scores) == 2 and len(loss) == 2
    for r in scores + loss:
        assert sorted(r) == ["Igralec_1", "Igralec_2", "Igralec_3", "Igralec_4"]
    assert pickle.load(open(os.path.join(d, "scores.pickle"), "rb")) == scores
    assert pickle.load(open(os.path.join(d, "kwargs.pickle"), "rb")) == {"final_reword_factor": 0.1, "random_card": 0.1}
    assert all(isinstance(v, int) for r in scores for v in r.values())
    assert any(v is not None and v == v for r in loss for v in r.values())            # somebody learnt something (not NaN)
    scores2, loss2 = P.main(d, iteracij=1, num_games=768, seed=6, uci=False)             # continues from the directory (main.py:100-101)
    assert len(scores2) == 3 and scores2[:2] == scores and loss2[2] == {}
