"""Batched neural self-play on the device (BASELINE config 4).

What ``Tarok.paralel_start`` does with four ``Nevronski_igralec`` (``Tarok.py:30-62``; ``main.py:48-66``;
``Igralec.py:278-385``) -- every player queues one observation per game it has to act in, runs one batched forward per
net type, then per game a ``.cpu()`` sync, an argmax over the legal cards and a Python state update -- becomes, per
phase, a handful of launches over the whole batch:

    deal -> obs_hands -> Vrednotenje_roke (per player) -> argmax/epsilon -> auction
         -> obs_exchange -> Zalaganje (per player) -> select_exchange -> exchange
         -> 48 x [ obs_buckets -> per (player, net, T) bucket: obs_expand -> net forward -> select_action ] -> step
         -> score

The reference seats FOUR players, each with its own six networks and its own epsilon (``main.py:48-66``,
``Igralec.py:190,235-248``); game i seats them rotated by i % 4 (``Tarok.py:34``), so the player of seat s in game i is
(s + i) % 4.  The buckets of ``predict_igraj_karto`` (one queue per net type per player, stacked per history length T,
``Igralec.py:316-342``) are formed on the DEVICE by a counting sort (``tarok_obs_buckets``): the host reads 128 counts
once per step -- the only synchronisation of a step -- and walks the non-empty ranges.

The environment, the bucketing, the observation expansion and the action selection are the CUDA kernels of this
repository; the forward passes stay in the reference's networks (here the PyTorch restatement in ``tarok_b200.mreze``,
because the upstream ``torch_models.py`` is missing).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Union

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


class _Arena:
    """Output buffers of the observation kernels, allocated once for the largest bucket (n games, T = 56) and re-used by
    every bucket of every step (the launches of one stream are ordered, so a bucket's forward has consumed its inputs before
    the next expansion overwrites them)."""

    def __init__(self, n, dev):
        f = torch.float32
        self.opp = torch.empty(n * 56 * 162, dtype=f, device=dev)
        self.hand = torch.empty(n * 56 * 54, dtype=f, device=dev)
        self.talon = torch.empty(n * 330, dtype=f, device=dev)
        self.talon_klop = torch.empty(n * 54, dtype=f, device=dev)
        self.king = torch.empty(n * 4, dtype=f, device=dev)
        self.decl = torch.empty(n * 4, dtype=f, device=dev)
        self.disc = torch.empty(n * 54, dtype=f, device=dev)


class Samoigra:
    """``igralci`` = 4 (default): four players with their own networks and epsilon, as in the reference; ``igralci`` = 1:
    one shared set plays every seat.  ``mreze``: None (fresh random nets), one dict (shared, implies one player) or a
    sequence of four dicts; ``random_card``: one epsilon or one per player (``Nevronski_igralec.random_card``)."""

    def __init__(self, n_envs: int, mreze: Union[None, dict, Sequence[dict]] = None, device: int = 0,
                 seed: Optional[int] = None, random_card: Union[float, Sequence[float]] = 0.05, igralci: int = 4):
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        self.env = E.TarokEnv(n_envs, seed=seed, device=device, history=True)
        self.n = n_envs
        self.dev = self.env.torch_device
        if isinstance(mreze, dict):
            mreze, igralci = [mreze], 1
        if igralci not in (1, 4):
            raise ValueError("igralci must be 1 (one shared set of nets) or 4 (the reference's four players)")
        self.igralci = igralci
        self.mreze: List[dict] = list(mreze) if mreze is not None else [ustvari_mreze(self.dev) for _ in range(igralci)]
        if len(self.mreze) != igralci:
            raise ValueError("need %d sets of nets" % igralci)
        for d in self.mreze:
            for m in d.values():
                m.eval()                                               # Igralec.py:335
        eps = [float(random_card)] * igralci if isinstance(random_card, (int, float)) else [float(x) for x in random_card]
        if len(eps) != igralci:
            raise ValueError("need one random_card per player")
        self.random_card = eps
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(seed & 0x7FFFFFFFFFFFFFFF)
        self._arena = None
        self.zadnji_koraki = None        # per step: number of buckets walked (diagnostics)

    # ------------------------------------------------------------------ helpers
    def _igralec_sedeza(self, first_game_id: int) -> torch.Tensor:
        """int64 [n,4]: the player of every seat, (seat + global game id) % 4 (Tarok.py:34); zeros for one shared player."""
        if self.igralci == 1:
            return torch.zeros((self.n, 4), dtype=torch.int64, device=self.dev)
        gid = torch.arange(self.n, dtype=torch.int64, device=self.dev) + int(first_game_id)
        return (gid[:, None] + torch.arange(4, dtype=torch.int64, device=self.dev)[None, :]) & 3

    def _vhod(self, vr: int, T: int, B: int, off: int, row: int):
        """The input list of one bucket (reference order, A.4) as views of the arenas ``obs_expand_buckets`` filled:
        ``off`` = the bucket's offset in games, ``row`` = its offset in observation rows."""
        a = self._arena
        opp = a.opp[row * 162: (row + B * T) * 162].view(B, T, 3, 54)
        hand = a.hand[row * 54: (row + B * T) * 54].view(B, T, 54)
        if vr == 0:
            return [opp, hand, a.talon_klop[off * 54: (off + B) * 54].view(B, 54)]
        decl = a.decl[off * 4: (off + B) * 4].view(B, 4)
        if vr == 3:
            return [opp, hand, decl]
        talon = a.talon[off * 330: (off + B) * 330].view(B, 6, 55)
        disc = a.disc[off * 54: (off + B) * 54].view(B, 54)
        if vr == 2:
            return [opp, hand, talon, decl, disc]
        return [opp, a.king[off * 4: (off + B) * 4].view(B, 4), hand, talon, decl, disc]

    # ------------------------------------------------------------------ one rollout
    @torch.no_grad()
    def odigraj(self, first_game_id: int = 0, meri: bool = False):
        """Plays the whole batch once.  Returns (stats int64[32], per-section milliseconds or None)."""
        env, n, ura, P = self.env, self.n, _Ura(meri), self.igralci
        if self._arena is None:
            self._arena = _Arena(n, self.dev)
        env.reset_stats()
        with ura("env"):
            env.deal(first_game_id)
        kdo = self._igralec_sedeza(first_game_id)                              # [n,4] player of each seat
        # ---- bidding: Nevronski_igralec.pripavi_licitiram / predict_licitiram / licitiram (Igralec.py:278-306)
        with ura("obs"):
            x = env.obs_hands()                                                # [n,4,54]
        namen = torch.empty((n, 4), dtype=torch.int64, device=self.dev)
        vrsta_iger = torch.arange(n, device=self.dev)
        for p in range(P):
            if P == 4:                                                         # one seat per game belongs to player p
                with ura("bucket"):
                    sedez = (p - (vrsta_iger + int(first_game_id))) & 3
                    xp = x[vrsta_iger, sedez]
                with ura("forward"):
                    y = self.mreze[p]["Vrednotenje_roke"](xp)                  # [n,18]
                with ura("select"):
                    namen[vrsta_iger, sedez] = y.argmax(dim=-1)
            else:
                with ura("forward"):
                    y = self.mreze[0]["Vrednotenje_roke"](x.view(n * 4, 54)).view(n, 4, 18)
                with ura("select"):
                    namen = y.argmax(dim=-1)
        with ura("select"):                                                    # random.random() < random_card (Igralec.py:300-301)
            eps = torch.tensor(self.random_card, dtype=torch.float32, device=self.dev)[kdo]      # the seat's player's epsilon
            if max(self.random_card) > 0:
                razisci = torch.rand((n, 4), device=self.dev, generator=self.gen) < eps
                nakljucno = torch.randint(0, 18, (n, 4), device=self.dev, generator=self.gen)
                namen = torch.where(razisci, nakljucno, namen)
        with ura("env"):
            env.auction(namen.to(torch.uint8))
        # ---- talon exchange (Igralec.py:357-385): the declarer's player decides
        with ura("bucket"):
            meta = env.meta[:n]
            menja = E.meta_field(meta, E.M_PHASE, 2) == E.PH_EXCHANGE
            decl = E.meta_field(meta, E.M_DECL, 2)
            lastnik = torch.gather(kdo, 1, decl[:, None]).squeeze(1)
        kupcek = torch.full((n,), 0xFF, dtype=torch.uint8, device=self.dev)
        zalozi = torch.zeros(n, dtype=torch.int64, device=self.dev)
        for p in range(P):
            with ura("bucket"):
                sel = torch.nonzero(menja & (lastnik == p)).flatten().to(torch.int32)
            if sel.numel() == 0:
                continue
            with ura("obs"):
                vhod, _ = env.obs_exchange(sel)
            with ura("forward"):
                q = self.mreze[p]["Zalaganje"](vhod)
            with ura("select"):
                env.select_exchange(q, sel, self.random_card[p], group_out=kupcek, discard_out=zalozi)
        with ura("env"):
            env.exchange(kupcek, zalozi)
        # ---- 48 card plays (Igralec.py:312-355)
        karte = torch.full((env.n_alloc,), 0xFF, dtype=torch.uint8, device=self.dev)
        qmax = torch.zeros(n, dtype=torch.float32, device=self.dev)
        eps4 = (self.random_card * 4)[:4]
        koraki, ar = [], self._arena
        for _ in range(48):
            with ura("bucket"):                                                # counting sort on the device + 1.5 KB to the host
                sel_all, stevci = env.obs_buckets(4 if P == 4 else 1)
            skupaj = int(stevci[255])
            if skupaj == 0:
                break
            with ura("obs"):                                                   # every bucket's inputs in one launch
                env.obs_expand_buckets(skupaj, ar.opp, ar.hand, ar.talon, ar.talon_klop, ar.king, ar.decl, ar.disc)
            polni = [k for k in range(127) if stevci[k]]
            koraki.append(len(polni))
            q = {}
            with ura("forward"):
                for k in polni:
                    p, vr, T = k // 28, (k % 28) // 7, 8 * (k % 7 + 1)
                    vhod = self._vhod(vr, T, int(stevci[k]), int(stevci[128 + k]), int(stevci[256 + k]))
                    q[k] = self.mreze[p][_NET_OF_KIND[vr]](vhod).contiguous()      # the legal-mask vector is not a net input (Igralec.py:333)
            with ura("select"):                                                # every bucket's argmax / epsilon in one launch
                env.select_action_buckets(skupaj, q, eps4, karte, qmax)
            with ura("env"):
                env.step(karte)
        with ura("env"):
            env.score()
        self.zadnji_koraki = koraki
        st = env.stats()
        return st, (ura.ms() if meri else None)

    # ------------------------------------------------------------------ replay samples (Igralec.py:441-443)
    @torch.no_grad()
    def vzorci(self, final_reword_factor: float = 0.1, first_game_id: int = 0):
        """Replay samples of the batch just played, bucketed like ``Nevronski_igralec.zgodovina[(tip, T)]`` of each
        player (Igralec.py:441-443): yields (player, net name, T, stanje list without the legal-mask vector, dy [B,54])."""
        env, n = self.env, self.n
        dy, seat, rows = env.targets(final_reword_factor=final_reword_factor)
        vrsta = E.meta_field(env.meta[:n], E.M_CONTRACT, 4)
        kind = torch.full_like(vrsta, 2)                                   # Solo_* ...
        kind = torch.where(vrsta == 0, torch.zeros_like(kind), kind)       # Klop
        kind = torch.where((vrsta >= 1) & (vrsta <= 3), torch.ones_like(kind), kind)
        kind = torch.where((vrsta == 7) | (vrsta == 9), torch.full_like(kind, 3), kind)
        gid = torch.arange(n, dtype=torch.int64, device=self.dev) + int(first_game_id)
        for t in range(48):
            igrano = seat[:, t] != 0xFF
            if not bool(igrano.any()):
                break
            igralec = ((seat[:, t].to(torch.int64) + gid) & 3) if self.igralci == 4 else torch.zeros_like(gid)
            kljuc = igralec * 256 + kind * 64 + rows[:, t].to(kind.dtype)
            for k in torch.unique(kljuc[igrano]).tolist():
                p, vr, T = k // 256, (k % 256) // 64, k % 64
                sel = torch.nonzero(igrano & (kljuc == k)).flatten().to(torch.int32)
                stanje, ok = env.obs_expand(vr, T, sel, play=t)
                yield p, _NET_OF_KIND[vr], T, stanje[:-1], dy[sel.long(), t]

    def nauci(self, final_reword_factor: float = 0.1, lr: float = 1e-3, first_game_id: int = 0):
        """One optimisation pass over the samples of the last batch: Huber(delta=25) + Adam per play net of each player, the
        loss and optimiser of the reference's models (train.py:23-25; Igralec.py:545-607 fits them with Lightning).
        Training is outside the accelerated path; this exists so the self-play loop closes.  Returns {(player, net): loss}."""
        opt = {(p, k): torch.optim.Adam(self.mreze[p][k].parameters(), lr=lr)
               for p in range(self.igralci) for k in _NET_OF_KIND.values()}
        izguba, stevec = {}, {}
        for p, ime, T, stanje, dy in self.vzorci(final_reword_factor, first_game_id):
            m = self.mreze[p][ime]
            m.train()
            with torch.enable_grad():
                for a in range(0, dy.shape[0], 4096):
                    x = [s[a:a + 4096] for s in stanje]
                    if x[0].shape[0] < 2:
                        continue                                           # BatchNorm needs more than one sample
                    l = torch.nn.functional.huber_loss(m(x), dy[a:a + 4096], delta=25.0)
                    opt[(p, ime)].zero_grad(set_to_none=True)
                    l.backward()
                    opt[(p, ime)].step()
                    izguba[(p, ime)] = izguba.get((p, ime), 0.0) + float(l.detach())
                    stevec[(p, ime)] = stevec.get((p, ime), 0) + 1
            m.eval()
        return {k: izguba[k] / stevec[k] for k in izguba}

    def zapri(self):
        self.env.close()
