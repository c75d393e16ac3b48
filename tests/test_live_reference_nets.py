"""CPU, build container only: the reference's own ``Nevronski_igralec`` (imported from /root/reference) built on top of the
restated networks through ``tarok_b200/compat/torch_models.py`` -- its ``create_models`` / ``predict_licitiram`` /
``predict_igraj_karto`` / ``predict_izberi_iz_talona`` drive them with the tensors its own encoders produce, in real games of
the reference engine.  This pins the call contract of ``tarok_b200.mreze`` (argument lists, shapes, dtypes); the numerics have
no upstream counterpart to be pinned against."""
import os
import random
import sys
import types

import numpy as np
import pytest

pytestmark = pytest.mark.reference


def _load_reference_with_our_nets():
    from oracle import ref_harness as H
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tarok_b200", "compat")
    spec_name = "torch_models"
    import importlib.util
    spec = importlib.util.spec_from_file_location(spec_name, os.path.join(compat, "torch_models.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    stale = [m for m in ("Igralec", "torch_models") if m in sys.modules]
    saved = {m: sys.modules.pop(m) for m in stale}
    sys.modules[spec_name] = mod                      # instead of the harness's empty stub
    H._ref = None
    try:
        ref = H.load_reference()
    finally:
        H._ref = None
    return ref, saved


def test_reference_neural_player_runs_on_the_restated_nets(tmp_path, monkeypatch):
    import torch
    ref, saved = _load_reference_with_our_nets()
    try:
        monkeypatch.chdir(tmp_path)                   # the player creates its save directory relative to the cwd
        random.seed(3); np.random.seed(3); torch.manual_seed(3)
        N = ref.Igralec.Nevronski_igralec
        players = []
        for i in range(4):
            p = N.__new__(N)
            try:
                N.__init__(p, ime="n%d" % i)
            except TypeError:
                N.__init__(p)
            players.append(p)
        for p in players:
            assert set(p.models) == {"Navadna_igra", "Klop", "Solo", "Berac", "Vrednotenje_roke", "Zalaganje"}
            p.random_card = 0.2
        t = ref.Tarok.Tarok(players, 12)
        t.izpis = False
        with open(os.devnull, "w") as devnull:
            old, sys.stdout = sys.stdout, devnull
            try:
                t.paralel_start()                     # bidding, talon exchange and 48 plays per game through the nets
            finally:
                sys.stdout = old
        assert all(isinstance(v, (int, np.integer)) for v in t.rezultati.values())
    finally:
        for m in ("Igralec", "torch_models"):
            sys.modules.pop(m, None)
        sys.modules.update(saved)
