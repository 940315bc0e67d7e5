"""Fused trainer step: the caller side of the hot path.

``model/universal_model_trainer.py:80-90`` materialises ``pulses.repeat_interleave(M)``, samples
``M*B`` errors on the CPU, copies them H2D and calls generator -> loss_fn -> backward.
:class:`FusedStepMixin` overrides exactly those two methods (``train_epoch`` / ``evaluate``) with
ONE call to :func:`fused_propagate_loss`: pulses stay ``(B, L, P)``, errors are generated on-chip by
Philox (counter = global sample index, key = seed, offset = step counter), the evaluation fidelity
comes from the same kind of pass, and with a process group the Monte-Carlo samples are sharded
over ranks with a single all-reduce.  Everything else of the reference trainer (curriculum loop,
best-state tracking, persistence) is reused unchanged when the mixin is combined with it:

    from model.universal_model_trainer import UniversalModelTrainer          # the reference's
    FusedTrainer = fused_trainer_class(UniversalModelTrainer)
    trainer = FusedTrainer(model, monte_carlo=1000, device="cuda")
    trainer.train(...)                                                       # trainer.py:137-231 as is

:class:`FusedTrainer` below is the same thing with a minimal stand-alone curriculum loop for
environments where the reference package is not importable.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import ops


class SigmaSpec:
    """What ``get_error_distribution`` returns in fused mode: the (delta_std, epsilon_std) of the
    current curriculum stage.  Still callable like the reference's sampler closure
    (``trainer.py:127-131``) for code that wants explicit samples."""

    def __init__(self, delta_std=1.0, epsilon_std=0.05, **_):
        self.delta_std = float(delta_std)
        self.epsilon_std = float(epsilon_std)

    def __call__(self, batch_size: int) -> torch.Tensor:
        return ops.get_ore_ple_error_distribution(batch_size, self.delta_std, self.epsilon_std)

    @property
    def sigma(self):
        return (self.delta_std, self.epsilon_std)


def sync_replicas(model: torch.nn.Module, group, device) -> None:
    """Sample sharding keeps the model REPLICATED: every rank runs the same forward / backward / optimiser step on
    bit-identical ``[G | Fsum]`` (no parameter all-reduce).  That only holds if the replicas start identical and draw
    the same dropout masks, so rank 0's parameters, buffers and CUDA RNG state are broadcast once at construction."""
    import torch.distributed as dist
    if group is None or dist.get_world_size(group) == 1:
        return
    dev = torch.device(device)
    src = dist.get_global_rank(group, 0)
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src, group=group)
    state = torch.cuda.get_rng_state(dev).to(dev)
    dist.broadcast(state, src, group=group)
    torch.cuda.set_rng_state(state.cpu(), dev)


class FusedStepMixin:
    """train_epoch / evaluate of ``UniversalModelTrainer`` over the fused op."""

    fused_loss: str = "sharp"
    fused_seed: int = 0
    fused_group = None
    fused_dtype: Optional[torch.dtype] = None
    fused_autotune: bool = False                 # measure the launch shape once per (B, L, M) (ops.autotune_flags)
    _fused_step: int = 0
    clip_norm: float = 1.0                       # trainer.py:91

    def get_error_distribution(self, *, error_params: Dict):      # trainer.py:127-131
        return SigmaSpec(**error_params)

    def _fused_call(self, pulses, U_target, spec, loss):
        if pulses.shape[-1] == 3:
            fn = ops.fused_propagate_loss_su4
        else:
            fn = ops.fused_propagate_loss
        sigma = spec.sigma if isinstance(spec, SigmaSpec) else (1.0, 0.05)
        error = None
        if not isinstance(spec, SigmaSpec):               # a reference-style sampler closure: explicit errors
            error = spec(self.monte_carlo * U_target.shape[0]).to(pulses.device)
        self._fused_step += 1
        group = self._exchange_for(pulses)
        kw = {}
        if self.fused_autotune and pulses.shape[-1] == 2 and group is None:
            kw["flags"] = ops.autotune_flags(pulses.shape[0], pulses.shape[1], self.monte_carlo,
                                             self.fused_dtype or torch.float32, pulses.device)
        return fn(pulses, U_target, error=error, monte_carlo=self.monte_carlo, sigma=sigma, seed=self.fused_seed,
                  offset=self._fused_step, loss=loss, dtype=self.fused_dtype, group=group, **kw)

    fused_peer_exchange: bool = True             # multi-GPU: small exchange vectors go over NVLink peer memory

    def _exchange_for(self, pulses):
        """The process group itself (NCCL all-reduce), or - for the SU(2) path with a small [Fsum | G] vector - a
        cached :class:`PeerExchange` over it (exchange fused into the partials reduction, peer.py)."""
        group = self.fused_group
        if group is None or not self.fused_peer_exchange or pulses.shape[-1] != 2 or isinstance(group, ops.PeerExchange):
            return group
        B, L, P = pulses.shape
        from . import peer as _peer
        if B + B * L * P > _peer.MAX_N:
            return group
        key = (B, L, P, self.fused_dtype or torch.float32)
        cache = self.__dict__.setdefault("_peer_cache", {})
        if key not in cache:
            try:
                cache[key] = ops.PeerExchange(group, B, L, P, key[3], pulses.device)
            except Exception as e:          # no peer mapping on this box (still a GPU path: NCCL)
                import warnings
                warnings.warn(f"PeerExchange unavailable ({e!r}); using the NCCL all-reduce")
                cache[key] = group
        return cache[key]

    fused_head: bool = True                      # models that expose logits()/uqoc_head (heads.Headless*) skip the pulses tensor

    def _head_call(self, U_emb, U_target, spec, loss):
        """The step with the model's element-wise tail folded into the fused kernel: single GPU,
        single-qubit models wrapped by ``heads.HeadlessGRAPE`` / ``heads.HeadlessTransformer``.  None: not applicable."""
        hs = getattr(self.model, "uqoc_head", None)
        if not self.fused_head or hs is None or self.fused_group is not None or not isinstance(spec, SigmaSpec):
            return None
        logits, azimuth = self.model.logits(U_emb)
        self._fused_step += 1
        return ops.fused_head_propagate_loss(logits, U_target, head=hs.kind, pulse_ranges=hs.pulse_ranges, phi_offset=azimuth,
                                             base_pulse=hs.base_pulse if hs.kind == "transformer" else None, scale=hs.scale,
                                             monte_carlo=self.monte_carlo, sigma=spec.sigma, seed=self.fused_seed,
                                             offset=self._fused_step, loss=loss, dtype=self.fused_dtype)

    def train_epoch(self, U_emb_batch, U_target_batch, error_distribution) -> float:   # trainer.py:58-94
        self.model.train()
        self.optimizer.zero_grad()
        U_emb = U_emb_batch.to(self.device)
        U_target = U_target_batch.to(self.device)
        out = self._head_call(U_emb, U_target, error_distribution, self.fused_loss)
        if out is None:
            pulses = self.model(U_emb)                              # (B, L, P)
            out = self._fused_call(pulses, U_target, error_distribution, self.fused_loss)
        loss = out[0]
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=self.clip_norm)
        self.optimizer.step()
        return float(loss.detach().item())

    @torch.no_grad()
    def evaluate(self, U_emb_batch, U_target_batch, error_distribution) -> float:        # trainer.py:101-121
        self.model.eval()
        U_emb = U_emb_batch.to(self.device)
        U_target = U_target_batch.to(self.device)
        out = self._head_call(U_emb, U_target, error_distribution, "none")
        if out is None:
            out = self._fused_call(self.model(U_emb), U_target, error_distribution, "none")
        mean_fid = out[0]                                                                # loss "none" = pooled mean F
        return float(mean_fid.item())


def fused_trainer_class(base):
    """Combine the mixin with the reference's ``UniversalModelTrainer`` (or any class with its
    attributes) so that ``train()`` and persistence are the reference's own code."""

    class FusedTrainer(FusedStepMixin, base):
        def __init__(self, model, unitary_generator=ops.batched_unitary_generator,
                     error_sampler=ops.get_ore_ple_error_distribution, *, fidelity_fn=ops.fidelity,
                     loss_fn=ops.sharp_loss, loss: str = "sharp", seed: int = 0, process_group=None,
                     compute_dtype: Optional[torch.dtype] = None, **kw):
            super().__init__(model, unitary_generator, error_sampler, fidelity_fn=fidelity_fn, loss_fn=loss_fn, **kw)
            self.fused_loss, self.fused_seed, self.fused_group, self.fused_dtype = loss, seed, process_group, compute_dtype
            if process_group is not None:
                sync_replicas(self.model, getattr(process_group, "group", process_group), self.device)

    FusedTrainer.__name__ = f"Fused{base.__name__}"
    return FusedTrainer


class FusedTrainer(FusedStepMixin):
    """Stand-alone trainer with the reference's constructor defaults (Adam lr 3e-5 ``trainer.py:46``,
    ``monte_carlo`` 1000 ``trainer.py:34``) and a minimal curriculum loop (``trainer.py:168-231``
    without tqdm / plotting / file output)."""

    def __init__(self, model, *, monte_carlo: int = 1000, device="cuda", optimizer=None, loss: str = "sharp",
                 seed: int = 0, process_group=None, compute_dtype: Optional[torch.dtype] = None):
        self.model = model.to(device)
        self.monte_carlo = monte_carlo
        self.device = device
        self.optimizer = optimizer or torch.optim.Adam(self.model.parameters(), lr=3e-5)
        self.fused_loss, self.fused_seed, self.fused_group, self.fused_dtype = loss, seed, process_group, compute_dtype
        self.best_state = None
        self.best_fidelity = 0.0
        if process_group is not None:
            sync_replicas(self.model, getattr(process_group, "group", process_group), device)

    def train(self, train_inputs: torch.Tensor, train_unitaries: torch.Tensor, eval_inputs: torch.Tensor,
              eval_unitaries: torch.Tensor, error_params_list: List[Dict], epochs: int = 100, batch_size: int = 10,
              log=None) -> List[Dict]:
        history = []
        nb_t, nb_e = train_inputs.shape[0] // batch_size, eval_inputs.shape[0] // batch_size   # trainer.py:161-164
        tr_in = train_inputs[: nb_t * batch_size].reshape(nb_t, batch_size, *train_inputs.shape[1:])
        tr_U = train_unitaries[: nb_t * batch_size].reshape(nb_t, batch_size, *train_unitaries.shape[1:])
        ev_in = eval_inputs[: nb_e * batch_size].reshape(nb_e, batch_size, *eval_inputs.shape[1:])
        ev_U = eval_unitaries[: nb_e * batch_size].reshape(nb_e, batch_size, *eval_unitaries.shape[1:])
        for error_params in error_params_list:                         # curriculum, trainer.py:168
            self.best_fidelity = 0.0
            spec = self.get_error_distribution(error_params=error_params)
            for epoch in range(1, epochs + 1):
                losses = [self.train_epoch(a, b, spec) for a, b in zip(tr_in, tr_U)]
                fids = [self.evaluate(a, b, spec) for a, b in zip(ev_in, ev_U)]
                loss_m, fid_m = sum(losses) / max(len(losses), 1), sum(fids) / max(len(fids), 1)
                if fid_m > self.best_fidelity:                         # trainer.py:191-195
                    self.best_fidelity = fid_m
                    self.best_state = {k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()}
                rec = {"error_params": dict(error_params), "epoch": epoch, "loss": loss_m, "fid": fid_m, "best": self.best_fidelity}
                history.append(rec)
                if log is not None:
                    log(rec)
            if self.best_state is not None:                            # trainer.py:224-225
                self.model.load_state_dict(self.best_state)
        return history


class GraphedTrainStep:
    """The whole optimisation step of ``trainer.py:67-92`` as ONE CUDA-graph replay:
    ``model(U_emb)`` -> fused propagate/loss -> ``backward()`` -> ``clip_grad_norm_`` -> ``optimizer.step()``.

    At the reference's step sizes the fused op is a 60-200 us kernel while eager PyTorch spends milliseconds of
    host time launching the pulse generator's small kernels; a captured step costs one graph launch.  The Philox
    (seed, offset) pair lives in device memory and is advanced INSIDE the graph, so every replay draws fresh error
    samples (``UQOC_FLAG_RNG_FROM_DEVICE``); dropout keeps working through torch's graph-safe generator.  One graph
    is captured per (delta_std, epsilon_std) curriculum stage on first use.

        step = GraphedTrainStep(model, B=200, emb_shape=(4,), monte_carlo=1000, lr=3e-5)
        loss = step(U_emb_batch, U_target_batch, SigmaSpec(0.4, 0.05))     # device scalar; .item() when needed
    """

    def __init__(self, model, *, B: int, emb_shape: Sequence[int], monte_carlo: int = 1000, device="cuda", lr: float = 3e-5,
                 optimizer=None, loss: str = "sharp", seed: int = 0, clip_norm: float = 1.0, emb_dtype=torch.float32,
                 compute_dtype: Optional[torch.dtype] = None, target_dim: int = 2, flags: int = 0):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs a CUDA device: the uqoc ops have no CPU fallback")
        self.model = model.to(self.device)
        # Adam with device-side step counters: required for optimizer.step() inside a CUDA graph (trainer.py:46 lr)
        self.optimizer = optimizer or torch.optim.Adam(self.model.parameters(), lr=lr, capturable=True)
        self.M, self.loss, self.clip_norm, self.dtype, self.flags = int(monte_carlo), loss, clip_norm, compute_dtype, flags
        cdt = torch.complex128 if compute_dtype == torch.float64 else torch.complex64
        self.s_emb = torch.zeros(B, *emb_shape, dtype=emb_dtype, device=self.device)
        self.s_target = torch.zeros(B, target_dim, target_dim, dtype=cdt, device=self.device)
        self.s_target[:] = torch.eye(target_dim, dtype=cdt, device=self.device)
        self.d_rng = torch.tensor([seed, 0], dtype=torch.int64, device=self.device)
        self.s_loss = torch.zeros((), dtype=torch.float32, device=self.device)
        self.s_fid = torch.zeros(B, dtype=torch.float32, device=self.device)
        self._graphs = {}
        self.fused_head = True                                        # fold a Headless* model's tail into the fused kernel
        self._ws = None                                               # private workspace (its address is captured)
        self._stream = torch.cuda.Stream(self.device)

    def _step_body(self, sigma):
        self.d_rng[1] += 1                                            # fresh Philox offset on every replay
        kw = dict(monte_carlo=self.M, sigma=sigma, seed=self.d_rng.data_ptr(), loss=self.loss, dtype=self.dtype,
                  flags=self.flags | 8)                               # 8 = UQOC_FLAG_RNG_FROM_DEVICE
        hs = getattr(self.model, "uqoc_head", None)
        if hs is not None and self.fused_head:
            # model wrapped by heads.Headless*: its element-wise tail runs inside the fused kernel (no pulses tensor)
            logits, azimuth = self.model.logits(self.s_emb)
            if self._ws is None:
                self._ws = ops.su2_workspace(logits.shape[0], logits.shape[1], self.M, self.dtype or torch.float32,
                                             kw["flags"] | ops.FLAG_RAW_TARGET, self.device)
            loss, fid = ops.fused_head_propagate_loss(logits, self.s_target, head=hs.kind, pulse_ranges=hs.pulse_ranges,
                                                      phi_offset=azimuth, base_pulse=hs.base_pulse if hs.kind == "transformer" else None,
                                                      scale=hs.scale, workspace=self._ws, **kw)
            self._finish_step(loss, fid)
            return
        pulses = self.model(self.s_emb)
        fn = ops.fused_propagate_loss_su4 if pulses.shape[-1] == 3 else ops.fused_propagate_loss
        if self._ws is None:                                          # first warm-up step, outside capture
            if pulses.shape[-1] == 3:
                self._ws = ops.su4_workspace(pulses.shape[0], pulses.shape[1], self.M, self.dtype or torch.float32, kw["flags"],
                                             self.device)
            else:
                self._ws = ops.su2_workspace(pulses.shape[0], pulses.shape[1], self.M, self.dtype or torch.float32,
                                             kw["flags"] | ops.FLAG_RAW_TARGET, self.device)
        loss, fid = fn(pulses, self.s_target, workspace=self._ws, **kw)
        self._finish_step(loss, fid)

    def _finish_step(self, loss, fid):
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=self.clip_norm)
        self.optimizer.step()
        self.s_loss.copy_(loss.detach().float())
        self.s_fid.copy_(fid.detach().float())

    def _capture(self, sigma):
        self.model.train()
        with torch.cuda.stream(self._stream):
            rng0 = self.d_rng.clone()
            state = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
            opt_before = {id(t): t.detach().clone() for st in self.optimizer.state.values() for t in st.values()
                          if torch.is_tensor(t)}
            for _ in range(3):                                        # warm-up outside capture (lazy init, pools,
                self.optimizer.zero_grad(set_to_none=True)            # optimizer state tensors must exist before capture)
                self._step_body(sigma)
            self._stream.synchronize()
            # warm-up steps must not count: restore parameters, the optimizer's moments / step counters IN PLACE
            # (tensors created by the warm-up are zeroed = a fresh optimizer) and the Philox offset
            self.model.load_state_dict(state)
            for st in self.optimizer.state.values():
                for t in st.values():
                    if torch.is_tensor(t):
                        if id(t) in opt_before:
                            t.copy_(opt_before[id(t)])
                        else:
                            t.zero_()
            self.d_rng.copy_(rng0)
            graph = torch.cuda.CUDAGraph()
            self.optimizer.zero_grad(set_to_none=True)
            with torch.cuda.graph(graph, stream=self._stream):
                self._step_body(sigma)
            # the capture itself does not execute: nothing to undo
        self._stream.synchronize()
        return graph

    def __call__(self, U_emb: torch.Tensor, U_target: torch.Tensor, spec) -> torch.Tensor:
        sigma = tuple(float(x) for x in (spec.sigma if isinstance(spec, SigmaSpec) else spec))
        graph = self._graphs.get(sigma)
        if graph is None:
            graph = self._graphs[sigma] = self._capture(sigma)
        self.s_emb.copy_(U_emb, non_blocking=True)
        self.s_target.copy_(U_target, non_blocking=True)
        self._stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._stream):                         # CUDAGraph.replay() launches on the CURRENT stream
            graph.replay()
        torch.cuda.current_stream(self.device).wait_stream(self._stream)
        return self.s_loss

    @property
    def mean_fidelity(self) -> torch.Tensor:
        """Per-target mean fidelity of the last replayed step (device tensor)."""
        return self.s_fid
