"""Whole-step CUDA-graph replay for training the drop-in model.

An eager training step of the N_side=256 network is ~900 kernel launches; at the small stages the GPU waits for the host
between them (5 % of the step, profiles/r1s_step_composition.log).  ``GraphedTrainStep`` captures forward + loss +
backward once and replays it: one launch per step, no host work in between.

What makes the capture legal: every hs_* entry point takes its stream explicitly (``ops.current_stream()`` is the
capturing stream, also on the autograd thread), allocates nothing and never synchronises; TMA descriptors are kernel
parameters built on the host from addresses that torch serves from the graph's private pool, hence stable across replays;
the weight operands of the GEMMs (split / rounded copies of the parameters) are refreshed by ONE batched launch before
every replay (``ops.refresh_weight_splits``), so a replay always sees the weights the optimizer just wrote without ~200
split kernels inside the graph; the small zero-initialised accumulators of the backward (~360 per step) are slices of one
arena that the captured step clears with a single memset (``ops.ZeroArena``).

Data parallelism: the captured part is rank-local.  Gradients live in ONE flat buffer (each ``p.grad`` is a view), so the
exchange step after the replay is a single NCCL all-reduce over NVLink, followed by the (eager, fused) optimizer step.
The reference gets the same semantics from Lightning's DDP plugin (heal_swin/train.py:187).

Dropout: the element / attention-probability masks of this library are pure functions of (seed, position).  Inside a
capture ``ops._next_dropout_seed`` hands out INDIRECT seeds that refer to a device-side counter (csrc/hs_common.h:
``resolve_seed``); the captured step increments that counter first, so every replay draws fresh masks while forward and
backward of one replay agree.  Stochastic depth uses torch's own graph-safe CUDA generator.
"""
import torch
import torch.distributed as dist

from . import ops


def _active_drop_probability(model):
    worst = 0.0
    for m in model.modules():
        p = m.p if isinstance(m, torch.nn.Dropout) else getattr(m, "drop_prob", None)
        if p and m.training:
            worst = max(worst, float(p))
    return worst


class GraphedTrainStep:
    """``loss = step(x, t)``: copies the batch into static buffers, replays forward + loss + backward, all-reduces the flat
    gradient over the data-parallel ranks and runs ``optimizer.step()``.  ``loss`` is a static device tensor (read it after
    the call; it is overwritten by the next one).

    ``x`` / ``t`` may be HOST tensors (pinned memory): they are copied straight into the static device buffers.  With
    ``preprocess`` the static buffers hold the RAW batch (e.g. uint8 images and class ids, as the data pipeline delivers
    them) and ``x, t = preprocess(raw_x, raw_t)`` -- the ``x.float()`` of the Lightning wrapper,
    models_lightning/segmentation/model_lightning_swin_hp.py:61 -- runs inside the captured graph."""

    def __init__(self, model, loss_fn, optimizer, example_x, example_t, warmup=3, preprocess=None):
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        self.has_dropout = _active_drop_probability(model) > 0.0
        self.seed_counter = ops.seed_counter(example_x.device)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.params = [q for q in model.parameters() if q.requires_grad]
        pad = lambda n: (n + 63) // 64 * 64  # every slice starts on a 256-byte boundary
        total = sum(pad(q.numel()) for q in self.params)
        self.flat_grad = torch.zeros(total, device=example_x.device, dtype=torch.float32)
        off = 0
        for q in self.params:
            assert q.dtype == torch.float32
            q.grad = self.flat_grad[off:off + q.numel()].view_as(q)
            off += pad(q.numel())
        self.x = example_x.clone()
        self.t = example_t.clone()
        self.warmup = warmup
        self.preprocess = preprocess
        self.graph = None
        self.loss = torch.zeros((), device=example_x.device, dtype=torch.float32)
        self.arena = ops.ZeroArena(example_x.device)

    def capture(self):
        """Warm-up on a side stream, then capture (done lazily by the first replaying call)."""
        was_timing, ops.STATS.timing = ops.STATS.timing, False
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # eager warm-up on a side stream, as torch.cuda.graphs requires
            for _ in range(self.warmup):
                self._fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        # Inside the graph the weight operands are NOT re-split (one small kernel per linear and orientation): all of them
        # are refreshed by one batched launch before every replay (__call__), and the small zero-initialised accumulators
        # of the backward come out of one arena cleared by a single memset.
        ops.refresh_weight_splits(self.x.device)
        prev = ops.maintain_weight_splits(True)
        try:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                loss = self._fwd_bwd()
                self.loss.copy_(loss)
        finally:
            ops.maintain_weight_splits(prev)
        self.flat_grad.zero_()
        ops.STATS.timing = was_timing

    def _fwd_bwd(self):
        self.seed_counter.add_(1)  # (captured: every replay advances the counter behind the indirect dropout seeds)
        self.flat_grad.zero_()
        self.arena.begin()
        prev = ops.use_zero_arena(self.arena)
        try:
            x, t = (self.x, self.t) if self.preprocess is None else self.preprocess(self.x, self.t)
            loss = self.loss_fn(self.model(x), t)
            loss.backward()  # accumulates in place into the views of flat_grad
            return loss.detach().clone()  # (the loss sums live in the arena)
        finally:
            ops.use_zero_arena(prev)

    def __call__(self, x, t, eager=False):
        """``eager=True`` runs the very same step without the graph (per-kernel timing, debugging)."""
        self.x.copy_(x, non_blocking=True)
        self.t.copy_(t, non_blocking=True)
        if eager:
            self.loss.copy_(self._fwd_bwd())
        else:
            if self.graph is None:
                self.capture()
            ops.refresh_weight_splits(self.x.device)  # the operands the replay reads, from the current parameters
            self.graph.replay()
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
            self.flat_grad.mul_(1.0 / self.world)
        self.optimizer.step()
        return self.loss
