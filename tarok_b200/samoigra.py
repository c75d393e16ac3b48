"""Batched neural self-play on the device (BASELINE config 4).

What ``Tarok.paralel_start`` does with four ``Nevronski_igralec`` (``Tarok.py:30-62``; ``Igralec.py:278-385``) --
queue one observation per game, one batched forward per net type, then per game a ``.cpu()`` sync, an argmax over the
legal cards and a Python state update -- becomes, per phase, a handful of launches over the whole batch:

    deal -> obs_hands -> Vrednotenje_roke -> argmax/epsilon -> auction
         -> obs_exchange -> Zalaganje -> select_exchange -> exchange
         -> 48 x [ obs_shape -> per (net, T) bucket: obs_expand -> net forward -> select_action ] -> step
         -> score

The environment, the observation expansion and the action selection are the CUDA kernels of this repository; the
forward passes stay in the reference's networks (here the PyTorch restatement in ``tarok_b200.mreze``, because the
upstream ``torch_models.py`` is missing).  One shared set of nets plays all four seats.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import env as E
from .mreze import ustvari_mreze

_NET_OF_KIND = {0: "Klop", 1: "Navadna_igra", 2: "Solo", 3: "Berac"}          # Nevronski_igralec.Tipi_NN


class _Ura:
    """CUDA-event stopwatch per section (summed after one synchronise)."""

    def __init__(self, on):
        self.on, self.pairs = on, {}

    def __call__(self, name):
        return _Odsek(self, name)

    def ms(self) -> Dict[str, float]:
        torch.cuda.synchronize()
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in self.pairs.items()}


class _Odsek:
    def __init__(self, ura, name):
        self.ura, self.name = ura, name

    def __enter__(self):
        if self.ura.on:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if self.ura.on:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            self.ura.pairs.setdefault(self.name, []).append((self.a, b))


class Samoigra:
    def __init__(self, n_envs: int, mreze: Optional[dict] = None, device: int = 0, seed: Optional[int] = None,
                 random_card: float = 0.05):
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        self.env = E.TarokEnv(n_envs, seed=seed, device=device, history=True)
        self.n = n_envs
        self.dev = self.env.torch_device
        self.mreze = mreze if mreze is not None else ustvari_mreze(self.dev)
        for m in self.mreze.values():
            m.eval()                                                   # Igralec.py:335
        self.random_card = float(random_card)
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(seed & 0x7FFFFFFFFFFFFFFF)

    @torch.no_grad()
    def odigraj(self, first_game_id: int = 0, meri: bool = False):
        """Plays the whole batch once.  Returns (stats int64[32], per-section milliseconds or None)."""
        env, n, eps, ura = self.env, self.n, self.random_card, _Ura(meri)
        env.reset_stats()
        with ura("env"):
            env.deal(first_game_id)
        # ---- bidding: Nevronski_igralec.pripavi_licitiram / predict_licitiram / licitiram (Igralec.py:278-306)
        with ura("obs"):
            x = env.obs_hands()
        with ura("forward"):
            y = self.mreze["Vrednotenje_roke"](x.view(n * 4, 54)).view(n, 4, 18)
        with ura("select"):
            namen = y.argmax(dim=2)
            if eps > 0:
                razisci = torch.rand((n, 4), device=self.dev, generator=self.gen) < eps
                nakljucno = torch.randint(0, 18, (n, 4), device=self.dev, generator=self.gen)
                namen = torch.where(razisci, nakljucno, namen)
        with ura("env"):
            env.auction(namen.to(torch.uint8))
        # ---- talon exchange (Igralec.py:357-385)
        faza = E.meta_field(env.meta[:n], E.M_PHASE, 2)
        menjajo = torch.nonzero(faza == E.PH_EXCHANGE).flatten().to(torch.int32)
        if menjajo.numel():
            with ura("obs"):
                vhod, _ = env.obs_exchange(menjajo)
            with ura("forward"):
                p = self.mreze["Zalaganje"](vhod)
            with ura("select"):
                kupcek, zalozi = env.select_exchange(p, menjajo, eps)
            with ura("env"):
                env.exchange(kupcek, zalozi)
        # ---- 48 card plays (Igralec.py:312-355)
        karte = torch.full((env.n_alloc,), 0xFF, dtype=torch.uint8, device=self.dev)
        qmax = torch.zeros(n, dtype=torch.float32, device=self.dev)
        for _ in range(48):
            with ura("obs"):
                vrsta, vrstice = env.obs_shape()
                kljuc = vrsta.to(torch.int32) * 64 + vrstice.to(torch.int32)
                zivi = vrsta != 255
                if not bool(zivi.any()):
                    break
                skupine = torch.unique(kljuc[zivi]).tolist()           # the (net, T) buckets of predict_igraj_karto
            karte.fill_(0xFF)
            for k in skupine:
                vr, T = k // 64, k % 64
                with ura("obs"):
                    sel = torch.nonzero(kljuc == k).flatten().to(torch.int32)
                    vhod, _ = env.obs_expand(vr, T, sel)
                with ura("forward"):
                    q = self.mreze[_NET_OF_KIND[vr]](vhod[:-1])        # the legal-mask vector is not a net input (Igralec.py:333)
                with ura("select"):
                    env.select_action(q, sel, eps, cards=karte, qmax=qmax)
            with ura("env"):
                env.step(karte)
        with ura("env"):
            env.score()
        st = env.stats()
        return st, (ura.ms() if meri else None)

    @torch.no_grad()
    def vzorci(self, final_reword_factor: float = 0.1):
        """Replay samples of the batch just played, bucketed like ``Nevronski_igralec.zgodovina[(tip, T)]``
        (Igralec.py:441-443): yields (net name, T, stanje list without the legal-mask vector, dy [B,54])."""
        env, n = self.env, self.n
        dy, seat, rows = env.targets(final_reword_factor=final_reword_factor)
        vrsta = E.meta_field(env.meta[:n], E.M_CONTRACT, 4)
        kind = torch.full_like(vrsta, 2)                                   # Solo_* ...
        kind = torch.where(vrsta == 0, torch.zeros_like(kind), kind)       # Klop
        kind = torch.where((vrsta >= 1) & (vrsta <= 3), torch.ones_like(kind), kind)
        kind = torch.where((vrsta == 7) | (vrsta == 9), torch.full_like(kind, 3), kind)
        for t in range(48):
            igrano = seat[:, t] != 0xFF
            if not bool(igrano.any()):
                break
            kljuc = kind * 64 + rows[:, t].to(kind.dtype)
            for k in torch.unique(kljuc[igrano]).tolist():
                vr, T = k // 64, k % 64
                sel = torch.nonzero(igrano & (kljuc == k)).flatten().to(torch.int32)
                stanje, ok = env.obs_expand(vr, T, sel, play=t)
                yield _NET_OF_KIND[vr], T, stanje[:-1], dy[sel.long(), t]

    def nauci(self, final_reword_factor: float = 0.1, lr: float = 1e-3):
        """One optimisation pass over the samples of the last batch: Huber(delta=25) + Adam per play net, the loss and
        optimiser of the reference's models (train.py:23-25; Igralec.py:545-607 fits them with Lightning).  Training is
        outside the accelerated path; this exists so the self-play loop closes.  Returns {net name: mean loss}."""
        opt = {k: torch.optim.Adam(self.mreze[k].parameters(), lr=lr) for k in _NET_OF_KIND.values()}
        izguba, stevec = {}, {}
        for ime, T, stanje, dy in self.vzorci(final_reword_factor):
            m = self.mreze[ime]
            m.train()
            with torch.enable_grad():
                for a in range(0, dy.shape[0], 4096):
                    x = [s[a:a + 4096] for s in stanje]
                    if x[0].shape[0] < 2:
                        continue                                           # BatchNorm needs more than one sample
                    l = torch.nn.functional.huber_loss(m(x), dy[a:a + 4096], delta=25.0)
                    opt[ime].zero_grad(set_to_none=True)
                    l.backward()
                    opt[ime].step()
                    izguba[ime] = izguba.get(ime, 0.0) + float(l.detach())
                    stevec[ime] = stevec.get(ime, 0) + 1
            m.eval()
        return {k: izguba[k] / stevec[k] for k in izguba}

    def zapri(self):
        self.env.close()
