"""The benchmarked training step of the hot path, host side.

Mirrors the data flow of the reference's joint stage (train_and_test.py:547-603) with the flow-based
transfer replaced by AdaIN + Gram loss as BASELINE.json's north_star prescribes:

    tf  = FE_t(xt)                 target extractor          (OS_CNN_res)
    sf  = FE_s(xs)                 source extractor          (OS_CNN_res)
    ssf = DimensionUnification(sf) source -> target shape    (torch, "next" row)
    s2t = AdaIN(ssf, tf)           feature-level style transfer
    L_style = Gram(s2t, tf)
    logits_t = CL_t(tf); logits_s = CL_s(ssf)                (OS_CNN)
    loss = CE_t + CE_s + style_weight * L_style ; backward ; RMSprop (reference learning rates)

Data parallel: one process per GPU, batch sharded by rank, per-rank BatchNorm statistics (DDP semantics),
one all-reduce over a flat fp32 gradient bucket per step followed by an explicit 1/N scale (RMSprop is not
linear in the gradient, so the scale is NOT folded into the learning rate).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as TF
from . import ops
from .OS_CNN import OS_CNN as OSM
from .OS_CNN.OS_CNN import OS_CNN, OS_CNN_res, layer_parameter_list_input_change
from .OS_CNN.OS_CNN_Structure_build import generate_layer_parameter_list
from .C_DAN import CDAN, RandomLayer
from .widgets import AdversarialNetworkforCDAN, DimensionUnification

MAX_KERNEL_SIZE = 89        # train_and_test.py:40
LEARNING_RATES = dict(fe_t=0.001, cl_t=0.003, fe_s=0.001, du=0.001, cl_s=0.003)    # train_and_test.py:97-101


def trainer_layer_lists(C: int, L: int):
    """Extractor / classifier layer lists exactly as train_and_test.py:38-53 derives them."""
    budgets = [8 * 128 * C, 5 * 128 * 256 + 2 * 256 * 128]
    ext = generate_layer_parameter_list(1, min(int(L / 4), MAX_KERNEL_SIZE), budgets, C)
    cf = sum(p[1] for p in ext[-1])
    return ext, layer_parameter_list_input_change(ext, cf), cf


class StyleTransferModelSet(nn.Module):
    """The five modules of the step, constructed in the reference's order (train_and_test.py:47-67) so that one
    ``torch.manual_seed`` reproduces the reference's initial parameters."""

    def __init__(self, Ct: int, Lt: int, Kt: int, Cs: int, Ls: int, Ks: int):
        super().__init__()
        lpl_t, lpl_c, cf_t = trainer_layer_lists(Ct, Lt)
        lpl_s, _, cf_s = trainer_layer_lists(Cs, Ls)
        self.fe_t = OS_CNN_res(lpl_t)
        self.cl_t = OS_CNN(lpl_c, Kt)
        self.fe_s = OS_CNN_res(lpl_s)
        self.du = DimensionUnification(cf_s, cf_t, Ls, Lt)
        self.cl_s = OS_CNN(lpl_c, Ks)                 # the reference reuses the target's list (train_and_test.py:67)
        self.feature_channels = cf_t

    two_streams = True      # run the target and the source branch on two CUDA streams (they are independent up to AdaIN)
    grads_final_callback = None   # set by the data-parallel trainer: called with a module name ("cl_t", "cl_s") during backward
                                  # as soon as that module's parameter gradients are complete (its slice of the flat
                                  # bucket can be all-reduced while the extractors' backward is still running)

    def _announce_final(self, feature: torch.Tensor, name: str) -> None:
        """The gradient of an extractor output is complete only after every consumer has run its backward -- among them
        the classifier `name` (and, in the C-DAN pair, its second call on the generated features, whose gradient reaches
        the target feature through AdaIN's style input)."""
        cb = self.grads_final_callback
        if cb is not None and feature.requires_grad:
            def hook(grad, _name=name, _cb=cb):
                _cb(_name)
                return None
            feature.register_hook(hook)

    def forward(self, xt, yt, xs, ys, style_weight: float = 1.0) -> Dict[str, torch.Tensor]:
        if not (self.two_streams and xt.is_cuda):
            tf = self.fe_t(xt)
            sf = self.fe_s(xs)
            ssf = self.du(sf)
            s2t = TF.adain(ssf, tf)
            l_style = TF.gram_style_loss(s2t, tf)
            logits_t, _, ce_t = self.cl_t.forward_loss(tf, yt)
            logits_s, _, ce_s = self.cl_s.forward_loss(ssf, ys)
        else:
            # Every kernel of one branch is a single wave of <= 148 CTAs with long load / epilogue phases, and the conv,
            # wgrad and BatchNorm kernels keep to half of an SM's shared memory and TMEM: the two branches' launches are
            # co-resident and hide each other's latency.  Autograd replays the same stream assignment in backward.
            main = torch.cuda.current_stream()
            side = getattr(self, "_side_stream", None)
            if side is None:
                side = self._side_stream = torch.cuda.Stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):             # the longer branch: source extractor + DimensionUnification
                sf = self.fe_s(xs)
                ssf = self.du(sf)
            tf = self.fe_t(xt)
            logits_t, _, ce_t = self.cl_t.forward_loss(tf, yt)      # needs tf only: runs while the source branch is still busy
            main.wait_stream(side)
            ssf.record_stream(main)
            s2t = TF.adain(ssf, tf)
            l_style = TF.gram_style_loss(s2t, tf)
            with torch.cuda.stream(side):
                logits_s, _, ce_s = self.cl_s.forward_loss(ssf, ys)
            main.wait_stream(side)
            logits_s.record_stream(main)
            ce_s.record_stream(main)
        self._announce_final(tf, "cl_t")
        self._announce_final(ssf, "cl_s")
        loss = TF.weighted_loss_sum([ce_t, ce_s, l_style], [1.0, 1.0, float(style_weight)]) if xt.is_cuda \
            else ce_t + ce_s + style_weight * l_style
        return dict(loss=loss, ce_t=ce_t, ce_s=ce_s, l_style=l_style,
                    logits_t=logits_t, logits_s=logits_s, tf=tf, ssf=ssf, s2t=s2t)


class TransferPairModelSet(StyleTransferModelSet):
    """One (source, target) pair of BASELINE configuration 3: the five modules of the cfg2 step plus the
    C-DAN critic and its random layer (train_and_test.py:74-76).  The step adds the target classifier on the generated
    features in **eval-BatchNorm mode with gradients** (train_and_test.py:584-586) and the C-DAN loss on
    (target feature, generated feature, their logits) (train_and_test.py:590-591):

        loss = CE_t + CE_s + style_weight * L_style + cdan_weight * CDAN      (cdan_weight = 3: train_and_test.py:660)
    """

    LRS = dict(fe_t=0.001, cl_t=0.003, fe_s=0.001, du=0.001, cl_s=0.003, ad_net=0.001)   # train_and_test.py:97-101,105
    CLAMPS = dict(ad_net=0.0005)                                                          # train_and_test.py:763-764
    cdan_weight = 3.0

    def __init__(self, Ct: int, Lt: int, Kt: int, Cs: int, Ls: int, Ks: int, critic_hidden: int = 1024):
        super().__init__(Ct, Lt, Kt, Cs, Ls, Ks)
        self.random_layer = RandomLayer([self.feature_channels * Lt, Kt], with_nvidia=False)   # follows .cuda()
        self.ad_net = AdversarialNetworkforCDAN(1024, critic_hidden)

    def forward(self, xt, yt, xs, ys, style_weight: float = 1.0) -> Dict[str, torch.Tensor]:
        out = super().forward(xt, yt, xs, ys, style_weight)
        was_training = self.cl_t.training
        self.cl_t.eval()                                  # train_and_test.py:584: running statistics, gradients still flow
        try:
            logits_s2t, _ = self.cl_t(out["s2t"])
        finally:
            self.cl_t.train(was_training)
        cdan = CDAN(out["tf"], out["s2t"], out["logits_t"], logits_s2t, self.ad_net, self.random_layer)
        out.update(loss=out["loss"] + self.cdan_weight * cdan, cdan=cdan, logits_s2t=logits_s2t)
        return out

    def advance_schedules(self):
        """What the two critic calls of one step do to the reversal schedule, for steps replayed from a CUDA graph."""
        self.ad_net.reversal_coefficients(calls=2)

    def schedule_state(self):
        return (self.ad_net.iter_num, self.ad_net.coeff)

    def set_schedule_state(self, st):
        self.ad_net.iter_num, self.ad_net.coeff = st
        self.ad_net._coeff_host = None


class MultiSourceModelSet(nn.Module):
    """BASELINE configuration 3: several source domains, one target.  As in the reference, every (source_i, target)
    pair is an independent training with its own module set (main.py:7-11, multi_source_voting.py:265-267); one step
    advances all of them on the same target batch.  Pairs are independent, so each runs on its own pair of streams.

    forward(xt, yt, xs_0, ys_0, xs_1, ys_1, ..., style_weight) -> dict(loss = sum of the pair losses, pairs = [...])."""

    def __init__(self, target, sources, critic_hidden: int = 1024):
        super().__init__()
        Ct, Lt, Kt = target
        self.pairs = nn.ModuleList([TransferPairModelSet(Ct, Lt, Kt, Cs, Ls, Ks, critic_hidden) for (Cs, Ls, Ks) in sources])
        self._streams = None

    multi_stream = True      # one stream (pair) per source; False runs the pairs one after the other

    def parameter_groups(self):
        return [(list(getattr(pair, name).parameters()), lr, pair.CLAMPS.get(name, 0.0))
                for pair in self.pairs for name, lr in pair.LRS.items()]

    def forward(self, xt, yt, *rest) -> Dict[str, torch.Tensor]:
        n = len(self.pairs)
        style_weight = rest[2 * n] if len(rest) > 2 * n else 1.0
        outs = [None] * n
        if xt.is_cuda and n > 1 and self.multi_stream:
            main = torch.cuda.current_stream()
            if self._streams is None:
                self._streams = [torch.cuda.Stream() for _ in range(n - 1)]
            for i in range(1, n):
                st = self._streams[i - 1]
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    outs[i] = self.pairs[i](xt, yt, rest[2 * i], rest[2 * i + 1], style_weight)
            outs[0] = self.pairs[0](xt, yt, rest[0], rest[1], style_weight)
            for i in range(1, n):
                main.wait_stream(self._streams[i - 1])
                outs[i]["loss"].record_stream(main)
        else:
            for i in range(n):
                outs[i] = self.pairs[i](xt, yt, rest[2 * i], rest[2 * i + 1], style_weight)
        loss = outs[0]["loss"]
        for o in outs[1:]:
            loss = loss + o["loss"]
        return dict(loss=loss, pairs=outs)

    def advance_schedules(self):
        for pair in self.pairs:
            pair.advance_schedules()

    def schedule_state(self):
        return [pair.schedule_state() for pair in self.pairs]

    def set_schedule_state(self, st):
        for pair, s in zip(self.pairs, st):
            pair.set_schedule_state(s)


class SingleDomainModelSet(nn.Module):
    """Extractor + classifier of one domain (the pre-training stages, train_and_test.py:143-220, and BASELINE config 4:
    long-series OS-CNN forward + backward).  forward(x, y) -> dict(loss, logits)."""

    LRS = dict(fe=0.001, cl=0.003)          # train_and_test.py:97-98

    def __init__(self, C: int, L: int, K: int):
        super().__init__()
        lpl, lpl_c, cf = trainer_layer_lists(C, L)
        self.fe = OS_CNN_res(lpl)
        self.cl = OS_CNN(lpl_c, K)
        self.feature_channels = cf

    def forward(self, x, y, style_weight: float = 1.0) -> Dict[str, torch.Tensor]:
        logits, _, ce = self.cl.forward_loss(self.fe(x), y)
        return dict(loss=ce, logits=logits)


class FlatParameters:
    """Every trainable parameter, its gradient and its RMSprop state as views into three flat fp32 buffers
    (group by group), so that the data-parallel exchange is ONE all-reduce and the optimizer ONE kernel launch
    per step.  ``p.data`` / ``p.grad`` are re-bound to their slices; module code is unaffected."""

    def __init__(self, groups):
        """groups: list of (params, lr) or (params, lr, clamp) -- clamp > 0 clips the group to [-clamp, clamp] after its
        update (the WGAN clipping of the C-DAN critic, train_and_test.py:763-764)."""
        groups = [(g[0], g[1], g[2] if len(g) > 2 else 0.0) for g in groups]
        params = [p for ps, _, _ in groups for p in ps if p.requires_grad]
        dev = params[0].device
        # 4-element alignment of every tensor keeps the float4 path of the optimizer kernel on group boundaries
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]
        n = sum(sizes)
        self.params = params
        self.flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.group_end, self.group_lr, self.group_clamp = [], [], []
        off = 0
        it = iter(sizes)
        for ps, lr, clamp in groups:
            for p in ps:
                if not p.requires_grad:
                    continue
                sz = next(it)
                view = self.flat_p[off: off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_g[off: off + p.numel()].view_as(p)
                off += sz
            self.group_end.append(off)
            self.group_lr.append(lr)
            self.group_clamp.append(clamp)

    def zero_grad(self):
        self.flat_g.zero_()

    def all_reduce_sum(self, group=None) -> int:
        """Sum the flat gradient bucket over the ranks; returns the world size (the 1/N is applied inside the
        optimizer kernel)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=group)
            return dist.get_world_size(group)
        return 1

    def rmsprop(self, grad_scale: float = 1.0, alpha: float = 0.99, eps: float = 1e-8):
        ops.rmsprop_step(self.flat_p, self.flat_g, self.flat_v, self.group_end, self.group_lr, alpha, eps, grad_scale,
                         self.group_clamp)


def remaining_ranges(done, n: int):
    """The parts of [0, n) that the half-open ranges in `done` do not cover, as few contiguous ranges as possible and in
    ascending order (what is left of the flat gradient bucket after the early all-reduces of a step)."""
    pos, rest = 0, []
    for lo, hi in sorted(done) + [(n, n)]:
        if lo > pos:
            rest.append((pos, lo))
        pos = max(pos, hi)
    return rest


class Trainer:
    """forward + backward + (all-reduce) + fused RMSprop of the cfg2 step.

    ``use_graph=True`` captures forward + backward of one step in a CUDA graph (fixed shapes): the step is
    launch-bound at cfg2 size (about 200 small kernels), and replaying one graph removes the host from the loop."""

    def __init__(self, model: nn.Module, style_weight: float = 1.0, group=None, use_graph: bool = False, lrs=None):
        self.model = model
        self.style_weight = style_weight
        self.group = group
        if lrs is None and hasattr(model, "parameter_groups"):
            groups = model.parameter_groups()
        else:
            lrs = lrs if lrs is not None else getattr(model, "LRS", LEARNING_RATES)
            clamps = getattr(model, "CLAMPS", {})
            groups = [(list(getattr(model, name).parameters()), lr, clamps.get(name, 0.0)) for name, lr in lrs.items()]
        self.flat = FlatParameters(groups)
        # slice of the flat bucket of every named group (for the early, overlapped part of the gradient exchange)
        self._group_range = {}
        if not (lrs is None and hasattr(model, "parameter_groups")):
            lo = 0
            for name, hi in zip(lrs.keys(), self.flat.group_end):
                self._group_range[name] = (lo, hi)
                lo = hi
        # TSC_DP_OVERLAP=1: all-reduce a classifier's slice as soon as its gradients are final, during backward.  Off by default:
        # measured on 2 x B200 (profiles/r2_dp_exchange.md) the early all-reduces cost more than they hide -- the compute
        # kernels are single waves sized to the SM count, and NCCL's resident CTAs turn them into two waves (1.256 ms per step
        # against 1.223 with one all-reduce after backward; capping NCCL_MAX_CTAS makes both slower)
        self.overlap_exchange = os.environ.get("TSC_DP_OVERLAP", "0") == "1"
        self._comm_stream = None
        self._early_works, self._early_done = [], []
        if hasattr(model, "grads_final_callback"):
            model.grads_final_callback = self._on_grads_final
        self.use_graph = use_graph
        self._graph = None
        self._static_in = None
        self._static_loss = None

    def broadcast_parameters(self, src: int = 0):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.broadcast(self.flat.flat_p, src=src, group=self.group)
            for t in self.model.buffers():
                dist.broadcast(t.data, src=src, group=self.group)
            # RandomLayer keeps its projections as a plain tensor list (as the reference does, C_DAN.py:15): neither
            # parameters nor buffers, but every rank must train the shared critic against the SAME projections
            for m in self.model.modules():
                if isinstance(m, RandomLayer):
                    for t in m.random_matrix:
                        dist.broadcast(t, src=src, group=self.group)

    def _world(self) -> int:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group)
        return 1

    def _model_streams(self):
        out = []
        for m in self.model.modules():
            st = getattr(m, "_side_stream", None)
            if st is not None:
                out.append(st)
            out.extend(getattr(m, "_streams", None) or [])
        return out

    def _on_grads_final(self, name: str) -> None:
        """Autograd hook (see StyleTransferModelSet._announce_final): the gradients of group `name` are complete -- start
        the all-reduce of its slice on the communication stream while the rest of backward keeps the compute streams busy."""
        if not self.overlap_exchange or name not in self._group_range or name in self._early_done or self._world() <= 1:
            return
        import torch.distributed as dist
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream()
        comm = self._comm_stream
        comm.wait_stream(torch.cuda.current_stream())
        for st in self._model_streams():          # the group's wgrad / BatchNorm-backward kernels ran on its branch's stream
            comm.wait_stream(st)
        lo, hi = self._group_range[name]
        with torch.cuda.stream(comm):
            self._early_works.append(dist.all_reduce(self.flat.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self._early_done.append(name)

    def _exchange(self) -> int:
        """Sum the flat gradient bucket over the ranks: whatever the backward hooks have not already put on the wire, in as
        few contiguous all-reduces as possible.  Returns the world size (the 1/N is applied inside the optimizer kernel)."""
        world = self._world()
        if world <= 1:
            return 1
        import torch.distributed as dist
        if not self._early_done:
            dist.all_reduce(self.flat.flat_g, op=dist.ReduceOp.SUM, group=self.group)
            return world
        for lo, hi in remaining_ranges([self._group_range[n] for n in self._early_done], self.flat.flat_g.numel()):
            dist.all_reduce(self.flat.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
        for w in self._early_works:
            w.wait()                               # the current stream waits for the early slices
        torch.cuda.current_stream().wait_stream(self._comm_stream)
        self._early_works, self._early_done = [], []
        return world

    def _step_body(self, *inputs):
        """forward + backward + gradient exchange + optimizer: what one CUDA graph replays (NCCL collectives are capturable,
        so a data-parallel step is still ONE graph launch, and the exchange overlaps the tail of backward inside it)."""
        loss = self._fwd_bwd(*inputs)
        world = self._exchange()
        self.flat.rmsprop(grad_scale=1.0 / world)
        return loss

    def _fwd_bwd(self, *inputs):
        self._early_works, self._early_done = [], []
        self.flat.zero_grad()
        # every parameter owns a slice of the flat gradient bucket: the wgrad / BatchNorm-backward kernels add into it
        # in place (no AccumulateGrad kernels), so the bucket is complete the moment backward returns
        prev = TF.direct_grads()
        TF.set_direct_grads(True)
        # the dense GEMMs of DimensionUnification and of the classifier heads stay with cuBLAS ("next" rows); next to
        # the bf16 tensor-core engine they may use TF32 (the reference's own GPU default for its convolutions), the
        # fp32 engine keeps them exact
        tf32_prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = ops.engine_name() == "tcgen05"
        OSM.defer_batch_counters(True)
        try:
            out = self.model(*inputs, self.style_weight)
            counters = OSM.defer_batch_counters(False)
            if counters:
                torch._foreach_add_(counters, 1)          # one launch for every BatchNorm's num_batches_tracked
            out["loss"].backward()
            # direct_grads: the wgrad / BatchNorm-backward kernels of the side branches wrote into the flat bucket on their
            # own streams and reported None to autograd -- order the all-reduce / optimizer behind them explicitly instead
            # of relying on autograd's end-of-backward leaf-stream synchronisation
            cur = torch.cuda.current_stream()
            for st in self._model_streams():
                cur.wait_stream(st)
        finally:
            OSM.defer_batch_counters(False)
            TF.set_direct_grads(prev)
            torch.backends.cuda.matmul.allow_tf32 = tf32_prev
        return out["loss"].detach()

    def _capture(self, *inputs):
        self._static_in = [t.clone() for t in inputs]
        # the warm-up passes must leave no trace: BatchNorm running statistics are forward side effects
        saved = [b.detach().clone() for b in self.model.buffers()]
        saved_p, saved_v = self.flat.flat_p.clone(), self.flat.flat_v.clone()      # the warm-up steps update the parameters
        sched = self.model.schedule_state() if hasattr(self.model, "schedule_state") else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):                          # warm-up on a side stream, as graph capture requires (NCCL included)
                self._step_body(*self._static_in)
        torch.cuda.current_stream().wait_stream(side)
        with torch.no_grad():
            self.flat.flat_p.copy_(saved_p)
            self.flat.flat_v.copy_(saved_v)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_loss = self._step_body(*self._static_in)
        # capturing does not execute: parameters and optimizer state are untouched; BatchNorm statistics were moved by the
        # warm-up steps only
        with torch.no_grad():
            for b, v in zip(self.model.buffers(), saved):
                b.copy_(v)
        if sched is not None:
            self.model.set_schedule_state(sched)

    def release_graph(self) -> None:
        """Drop the captured step.  A CUDA graph that holds NCCL collectives must be destroyed BEFORE the process group is:
        ``destroy_process_group`` waits for the communicator's captured work to be released and hangs otherwise."""
        if self._graph is not None:
            torch.cuda.synchronize()
            self._graph = None
            self._static_loss = None
            torch.cuda.synchronize()

    def step(self, *inputs) -> torch.Tensor:
        if self.use_graph:
            if self._graph is None:
                self._capture(*inputs)
            for dst, src in zip(self._static_in, inputs):
                dst.copy_(src, non_blocking=True)
            if hasattr(self.model, "advance_schedules"):
                self.model.advance_schedules()          # host-side schedules (C-DAN reversal strength) -> device scalars
            self._graph.replay()
            return self._static_loss
        return self._step_body(*inputs)
