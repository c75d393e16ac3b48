"""The six policy/value networks of the reference, restated in PyTorch.

The reference imports them from ``torch_models`` (``Igralec.py:24``), a file that is NOT in the upstream
repository; the only architecture source is the legacy Keras builders in ``train.py``
(``train.py:4-272``).  These modules restate that architecture with the call contract ``Igralec.py`` uses:
``model(list_of_tensors) -> [B, 54]`` (``Igralec.py:336``), ``.eval()/.train()/.to(...)``, ``state_dict`` I/O
(``Igralec.py:808-823``), and the class names of ``Nevronski_igralec.create_models`` (``Igralec.py:235-240``).
No weights exist upstream, so numerics are **unpinned** (random init); they are outside the accelerated path
and exist so that BASELINE config 4 (policy forward + GPU env step) can run end to end.

Input lists (the trailing legal-mask vector is dropped before the forward, ``Igralec.py:333``):
    Navadna_igra  [opp (B,T,3,54), king (B,4), hand (B,T,54), talon (B,6,55), decl (B,4), discard (B,54)]
    Solo          [opp, hand, talon (B,6,55), decl, discard]
    Klop          [opp, hand, talon (B,54)]
    Berac         [opp, hand, decl]
    Vrednotenje_roke  hand (B,54) -> 18          Zalaganje  [hand (B,54), talon (B,54,6), game (B,15)] -> 60
"""
from __future__ import annotations

import torch
from torch import nn
import torch.nn.functional as F


class _Glava(nn.Module):
    """Dense32-ELU -> BatchNorm -> Dropout(.2) -> Dense32-ELU -> BatchNorm [-> Dropout(.2)] -> Dense(out)
    (``train.py:6-22, 35-41, 94-105``)."""

    def __init__(self, n_in, n_out, drugi_dropout):
        super().__init__()
        self.d1, self.b1 = nn.Linear(n_in, 32), nn.BatchNorm1d(32)
        self.d2, self.b2 = nn.Linear(32, 32), nn.BatchNorm1d(32)
        self.izhod = nn.Linear(32, n_out)
        self.drugi_dropout = drugi_dropout

    def forward(self, x):
        x = F.dropout(self.b1(F.elu(self.d1(x))), 0.2, self.training)
        x = self.b2(F.elu(self.d2(x)))
        if self.drugi_dropout:
            x = F.dropout(x, 0.2, self.training)
        return self.izhod(x)


class Net_vrednotenje_roke(nn.Module):
    """Hand evaluation for bidding: 54 -> 32 -> 32 -> 18 (``train.py:4-27``)."""

    def __init__(self):
        super().__init__()
        self.glava = _Glava(54, 18, True)

    def forward(self, x):
        if isinstance(x, (list, tuple)):
            x = x[0]
        return self.glava(x)


class Net_zalaganje(nn.Module):
    """Talon exchange: concat(flatten talon 324, hand 54, game 15) = 393 -> 32 -> 32 -> 60 (``train.py:29-49``)."""

    def __init__(self):
        super().__init__()
        self.glava = _Glava(393, 60, True)

    def forward(self, x):
        roka, talon, igra = x[0], x[1], x[2]
        return self.glava(torch.cat([talon.flatten(1), roka, igra], dim=1))


class _Igra(nn.Module):
    """Shared trunk of the four play nets: per-step Dense on the flattened opponents' one-hots -> LSTM over the
    rows, an LSTM over the own-hand rows, concat -> LSTM -> last state."""

    def __init__(self, h_opp, h_roka, h_skupaj, roka_dense=0, elu_opp=False, elu_roka=False, elu_skupaj=False):
        super().__init__()
        self.opp_dense = nn.Linear(162, 32)
        self.opp_lstm = nn.LSTM(32, h_opp, batch_first=True)
        self.roka_dense = nn.Linear(54, roka_dense) if roka_dense else None
        self.roka_lstm = nn.LSTM(roka_dense or 54, h_roka, batch_first=True)
        self.skupaj = nn.LSTM(h_opp + h_roka, h_skupaj, batch_first=True)
        self.elu_opp, self.elu_roka, self.elu_skupaj = elu_opp, elu_roka, elu_skupaj

    def forward(self, opp, roka):
        a, _ = self.opp_lstm(self.opp_dense(opp.flatten(2)))
        if self.elu_opp:
            a = F.elu(a)
        r = F.elu(self.roka_dense(roka)) if self.roka_dense is not None else roka
        r, _ = self.roka_lstm(r)
        if self.elu_roka:
            r = F.elu(r)
        y, _ = self.skupaj(torch.cat([a, r], dim=2))
        y = y[:, -1]
        return F.elu(y) if self.elu_skupaj else y


class Net_Navadna_igra(nn.Module):
    """``train.py:51-117``: LSTM32 || LSTM32 -> LSTM32 -> concat[32, king 4, talon 330, decl 4, discard 54] -> head."""

    def __init__(self):
        super().__init__()
        self.trup = _Igra(32, 32, 32)
        self.glava = _Glava(32 + 4 + 330 + 4 + 54, 54, True)

    def forward(self, x):
        opp, kralj, roka, talon, kdo, zalozil = x
        return self.glava(torch.cat([self.trup(opp, roka), kralj, talon.flatten(1), kdo, zalozil], dim=1))


class Net_Klop(nn.Module):
    """``train.py:119-175``: LSTM16-ELU || LSTM16-ELU -> LSTM16 -> concat talon 54 -> head (no second dropout)."""

    def __init__(self):
        super().__init__()
        self.trup = _Igra(16, 16, 16, elu_opp=True, elu_roka=True)
        self.glava = _Glava(16 + 54, 54, False)

    def forward(self, x):
        opp, roka, talon = x
        return self.glava(torch.cat([self.trup(opp, roka), talon], dim=1))


class Net_Solo(nn.Module):
    """``train.py:177-224``: LSTM32-ELU || Dense30-ELU -> LSTM20 -> LSTM32 -> concat[32, 330, 4, 54] -> head."""

    def __init__(self):
        super().__init__()
        self.trup = _Igra(32, 20, 32, roka_dense=30, elu_opp=True)
        self.glava = _Glava(32 + 330 + 4 + 54, 54, False)

    def forward(self, x):
        opp, roka, talon, kdo, zalozil = x
        return self.glava(torch.cat([self.trup(opp, roka), talon.flatten(1), kdo, zalozil], dim=1))


class Net_Berac(nn.Module):
    """``train.py:226-272``: LSTM32-ELU || Dense32-ELU -> LSTM32-ELU -> LSTM32-ELU -> concat decl 4 -> head."""

    def __init__(self):
        super().__init__()
        self.trup = _Igra(32, 32, 32, roka_dense=32, elu_opp=True, elu_roka=True, elu_skupaj=True)
        self.glava = _Glava(32 + 4, 54, False)

    def forward(self, x):
        opp, roka, kdo = x
        return self.glava(torch.cat([self.trup(opp, roka), kdo], dim=1))


def ustvari_mreze(device=None):
    """The dict ``Nevronski_igralec.create_models`` builds (``Igralec.py:235-240``)."""
    d = {"Navadna_igra": Net_Navadna_igra(), "Klop": Net_Klop(), "Solo": Net_Solo(), "Berac": Net_Berac(),
         "Vrednotenje_roke": Net_vrednotenje_roke(), "Zalaganje": Net_zalaganje()}
    if device is not None:
        for v in d.values():
            v.to(device=device)
    return d
