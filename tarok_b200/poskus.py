"""The reference's experiment files around the device self-play loop.

``main.py`` keeps an experiment in a directory: ``scores.pickle`` -- a list with one ``{str(player): score}`` dict per
iteration (``main.py:136``), ``loss.pickle`` -- a list with one ``{str(player): loss}`` dict per iteration
(``main.py:118-123``), ``kwargs.pickle`` -- ``{'final_reword_factor', 'random_card'}`` of the neural players
(``main.py:144-147``) and one sub-directory ``1`` .. ``4`` per player with its networks' ``state_dict`` files
(``Igralec.py:802-810``: ``Vrednotenje_roke.pth``, ``Zalaganje.pth``, ``<net>_A.pth`` for the four play nets).  The functions
here have the reference's names, arguments and return shapes (``naredi_nove_igralce`` ``main.py:73-87``, ``load_igralce``
``main.py:89-97``, ``main`` ``main.py:99-150``) and write the same files, so plots and scripts made for the reference's
experiment directories keep working; the games themselves run through ``Samoigra`` on the GPU.

Not reproduced (out of scope): the second network set of ``Double_Nevronski_Igralec`` (``*_B.pth``, Double-DQN) and
the Lightning fits of ``nauci`` -- ``Samoigra.nauci`` is a plain PyTorch pass that closes the loop.
"""
from __future__ import annotations

import os
import pickle
from random import shuffle
from typing import List, Optional

import torch

from .mreze import ustvari_mreze

_PLAY_NETS = ("Navadna_igra", "Klop", "Solo", "Berac")


class Igralec_na_napravi:
    """What the experiment files know of one ``Nevronski_igralec`` (``Igralec.py:190-233``): its name, its six networks, its
    epsilon and reward factor, and where its networks live on disk."""

    def __init__(self, load_path=None, save_path=None, random_card=0.05, learning_rate=0.1, final_reword_factor=0.1, ime=None,
                 device=None):
        self.ime = str(ime)
        self.random_card = float(random_card)
        self.learning_rate = float(learning_rate)
        self.final_reword_factor = float(final_reword_factor)
        self.save_path = save_path
        self.load_path = load_path if load_path is not None else None
        self.models = ustvari_mreze(device)
        if self.load_path is not None:
            self.load_models()
            self.load_path = self.save_path                                    # Igralec.py:229

    def __str__(self):                                                         # Igralec.py:118-119: the key in scores / loss
        return "Igralec_" + self.ime

    __repr__ = __str__

    def save_models(self):
        """``Double_Nevronski_Igralec.save_models`` file names for the first network set (``Igralec.py:802-808``)."""
        for k, v in self.models.items():
            name = k + ".pth" if k in ("Vrednotenje_roke", "Zalaganje") else k + "_A.pth"
            torch.save(v.state_dict(), os.path.join(self.save_path, name))

    def load_models(self):
        """``Igralec.py:812-825``; files of a second network set (``*_B.pth``) are accepted and ignored."""
        for f in sorted(os.listdir(self.load_path)):
            path = os.path.join(self.load_path, f)
            if f[-6:-4] == "_A" and f[:-6] in self.models:
                self.models[f[:-6]].load_state_dict(torch.load(path, map_location="cpu"))
            elif f[-6:-4] == "_B":
                continue
            elif f[:-4] in ("Vrednotenje_roke", "Zalaganje"):
                self.models[f[:-4]].load_state_dict(torch.load(path, map_location="cpu"))
            else:
                raise IOError("Napaka pri loadanju modelov. " + str(path))


def _datoteke(path):
    return os.path.join(path, "scores.pickle"), os.path.join(path, "kwargs.pickle"), os.path.join(path, "loss.pickle")


def naredi_nove_igralce(path, device=None, **kwargs):
    """``main.py:73-87``: a fresh experiment directory with empty pickles and four players ``1`` .. ``4``."""
    os.mkdir(path)
    scores_file, kwargs_file, loss_file = _datoteke(path)
    for f, prazno in ((scores_file, []), (loss_file, []), (kwargs_file, dict())):
        with open(f, "wb") as out:
            pickle.dump(prazno, out)
    nn = []
    for i in range(1, 5):
        p = os.path.join(path, str(i))
        os.mkdir(p)
        nn.append(Igralec_na_napravi(load_path=None, save_path=p, ime=i, device=device, **kwargs))
    return nn, scores_file, kwargs_file, loss_file


def load_igralce(path, device=None):
    """``main.py:89-97``: the four players of an existing experiment with the parameters its ``kwargs.pickle`` holds."""
    scores_file, kwargs_file, loss_file = _datoteke(path)
    with open(kwargs_file, "rb") as f:
        kwargs = pickle.load(f)
    nn = []
    for i in range(1, 5):
        p = os.path.join(path, str(i))
        ima_mreze = os.path.isdir(p) and any(x.endswith(".pth") for x in os.listdir(p))
        nn.append(Igralec_na_napravi(load_path=p if ima_mreze else None, save_path=p, ime=i, device=device, **kwargs))
    return nn, scores_file, kwargs_file, loss_file


def main(dir, iteracij: int = 1000, num_games: int = 2000, device: int = 0, seed: Optional[int] = None, uci: bool = True, **kwargs):
    """``main.py:99-150`` with the games on the GPU: per iteration shuffle the players (``main.py:113``), play ``num_games``
    concurrent games with the four of them (``Tarok(igralci, num_games).paralel_start()``), let every player learn from its
    samples, append ``{str(player): score}`` / ``{str(player): loss}`` and rewrite the three pickles and the networks.
    Returns (scores, loss)."""
    from .samoigra import Samoigra
    dev = torch.device("cuda", device)
    if os.path.isdir(dir):
        igralci, scores_file, kwargs_file, loss_file = load_igralce(dir, device=dev)
    else:
        kwargs.pop("debug", None)
        igralci, scores_file, kwargs_file, loss_file = naredi_nove_igralce(dir, device=dev, **kwargs)
    with open(scores_file, "rb") as f:
        scores = pickle.load(f)
    with open(loss_file, "rb") as f:
        loss = pickle.load(f)
    s = Samoigra(num_games, mreze=[i.models for i in igralci], device=device, seed=seed,
                 random_card=[i.random_card for i in igralci], igralci=4)
    try:
        for it in range(iteracij):
            shuffle(igralci)                                                   # main.py:113: who is player 0..3 of this batch
            s.mreze = [i.models for i in igralci]
            s.random_card = [i.random_card for i in igralci]
            for d in s.mreze:
                for m in d.values():
                    m.eval()                                                   # Igralec.py:335
            prvi = (len(scores) * num_games)                                    # fresh deals every iteration
            st, _ = s.odigraj(first_game_id=prvi)
            rezultati = {igr: int(st[4 + p]) for p, igr in enumerate(igralci)}     # Tarok.rezultati (Tarok.py:59-61), by player
            loss.append({})
            if uci:
                po_mrezi = s.nauci(final_reword_factor=igralci[0].final_reword_factor, lr=igralci[0].learning_rate, first_game_id=prvi)
                for p, igr in enumerate(igralci):
                    moje = [v for (q, _), v in po_mrezi.items() if q == p]
                    loss[-1][str(igr)] = (sum(moje) / len(moje)) if moje else None  # mean loss of the player's nets (Igralec.py:606)
            for igr in igralci:
                igr.save_models()
            scores.append({str(k): v for k, v in rezultati.items()})          # main.py:136
            with open(scores_file, "wb") as out:
                pickle.dump(scores, out)
            with open(loss_file, "wb") as out:
                pickle.dump(loss, out)
            with open(kwargs_file, "wb") as out:                               # main.py:144-147
                pickle.dump({"final_reword_factor": igralci[0].final_reword_factor, "random_card": igralci[0].random_card}, out)
    finally:
        s.zapri()
    return scores, loss
