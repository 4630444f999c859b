"""CUDA-graph front ends of the hot path.

The hot path is a chain of ~60 (inference) / ~400 (training step) small launches behind one large persistent kernel; on a
B200 the launch gaps and the Python/ctypes enqueue cost are comparable to the kernels themselves.  Capturing the chain once
and replaying it removes both, and lets independent branches (the audio AU_former next to the video chain) run on parallel
graph branches.  No tracing compiler is involved: the captured work is exactly the library's kernels.

    GraphedHotPath(model, stage3, frame, audio)      inference: replay(stage3, frame, audio) -> (sformer_out, out21, decisions)
    GraphedTrainStep(model, optimizer, ...)          zero_grad + forward + AULoss + backward in one graph, then FusedAdam.step()
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _lib
from .encoder import Transformer


def hot_path_modules(model) -> list:
    """The sub-modules whose weights the transformer hot path reads: SFormer stack, TFormer, the two AU_formers,
    the fusion head — flattened to the list of modules that own parameters or buffers.  Conv backbones are not among them."""
    vm = model.video_model.video_model
    roots = [vm.s_former, vm.t_former, model.video_model.au_head, model.audio_model.au_head, model.au_head]
    mods = []
    for r in roots:
        for mod in r.modules():
            if isinstance(mod, torch.nn.Conv2d) or isinstance(mod, torch.nn.BatchNorm2d):
                continue                                   # the conv stages inside SFormer's wrapper stay torch modules outside the path
            if mod._parameters or mod._buffers:
                mods.append(mod)
    return mods


def weights_signature(model, modules: Optional[list] = None) -> tuple:
    """Changes whenever a parameter or buffer may have changed: the optimiser epoch (FusedAdam writes parameters through raw
    pointers) plus every tensor's version counter, per-parameter update count and address (load_state_dict, in-place edits, .to(),
    a replaced Parameter object).  A captured graph holds raw pointers to the PACKED (bf16 / stacked / BN-folded) copies of the
    weights, which the eager path re-packs — and frees — on such a change: a graph replayed afterwards would read stale weights or
    recycled memory.

    ``modules`` = the modules to watch (default: every module of ``model``).  This runs on the host in front of EVERY graph replay:
    the full model's 600 tensors cost about 0.9 ms per call — as long as the whole 1.09 ms forward step, so the replay loop became
    host-bound on a slower host (1.40 ms per step measured) — which is why GraphedHotPath watches only the hot path's modules,
    through their ``_parameters`` / ``_buffers`` dicts directly (no module-tree walk, no generator chain)."""
    from . import functional as AF
    if modules is None:
        modules = list(model.modules())
    v = 0
    for mod in modules:
        for t in mod._parameters.values():
            if t is not None:
                v = (v * 1000003 + (t._version + getattr(t, "_avf_wver", 0)) * 31 + t.data_ptr()) & 0xFFFFFFFFFFFFFFFF
        for t in mod._buffers.values():
            if t is not None:
                v = (v * 1000003 + t._version * 31 + t.data_ptr()) & 0xFFFFFFFFFFFFFFFF
    return (AF.WEIGHTS_EPOCH, v)


def _packed_refs(model) -> list:
    """Strong references to every packed weight copy the model's modules currently cache (kept by a graph for as long as it may
    replay with them)."""
    keep = []
    for mod in model.modules():
        for name in ("_packed", "_front", "_last"):
            obj = getattr(mod, name, None)
            if obj is not None and not isinstance(obj, torch.nn.Module):
                keep.append(obj)
    return keep


def _side_warmup(fn, iters: int = 3) -> None:
    """Warm-up on a side stream (allocations, lazy kernel attributes, weight packing) as torch.cuda.graph requires."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(iters):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()


class GraphedHotPath:
    """model.hot_path(...) captured for fixed input shapes.  Inputs are copied into static buffers (device-to-device or
    host-to-device), outputs are the graph's static tensors (valid until the next replay)."""

    def __init__(self, model, stage3: torch.Tensor, frame: torch.Tensor, audio: torch.Tensor, sm_split=None, gather_into: Optional[torch.Tensor] = None):
        """gather_into: the evaluation-time gather of the logits INSIDE the captured graph.
        A dp.PeerLogitGather: its push kernel (NVLink peer stores, no rendezvous) sits right behind the fusion head and its wait at
        the very end, behind the SFormer kernel — the default of bench.py at N > 1.  A [world * n_clips, 21] fp32 tensor: an NCCL
        all-gather as a graph branch from the fusion head to the end (measured SLOWER than the in-stream call on 8 GPUs: NCCL's CTAs
        and the persistent SFormer grid race for SMs and a rank whose gather started late stalls the others'; kept for A/B).
        sm_split = (sformer_sms, chain_sms[, frames_beside]): run the persistent SFormer kernel on `sformer_sms` SMs NEXT TO the
        TFormer / AU_former / fusion-head chain, whose persistent GEMMs are capped at `chain_sms` CTAs (every persistent CTA of
        either family owns a whole SM, so the two grids partition the GPU).  With `frames_beside` only that many leading
        stage-3 maps go through the side-by-side launch — sized to last about as long as the chain, half of whose kernels
        (fusion head: 64 CTAs, AU_former GEMMs: 32-96) leave SMs idle — and the rest follows on all SMs.
        None = the two run one after the other on all SMs.  Default from the environment variable AVF_SM_SPLIT="s,c[,f]"."""
        if model.training:
            raise RuntimeError("GraphedHotPath captures the inference kernels: call model.eval() first")
        if sm_split is None and os.environ.get("AVF_SM_SPLIT", "") not in ("", "0", "off"):
            sm_split = tuple(int(v) for v in os.environ["AVF_SM_SPLIT"].split(","))
        self.sm_split = tuple(sm_split) if sm_split else None
        self.gather_into = gather_into
        self._peer = gather_into if hasattr(gather_into, "push") else None
        if gather_into is not None:
            import torch.distributed as dist
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("GraphedHotPath(gather_into=...) needs an initialised torch.distributed process group")
        self.side2 = torch.cuda.Stream()
        self.model = model
        self.stage3, self.frame = stage3.clone(), frame.clone()
        self.audio = audio.float().clone()
        self.side = torch.cuda.Stream()
        self.recaptures = 0
        self._capture()

    def _capture(self) -> None:
        with torch.no_grad():
            _side_warmup(self._run)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = self._run()
        self._keep = _packed_refs(self.model)          # the graph reads these through raw pointers
        self._watch = hot_path_modules(self.model)
        self._sig = weights_signature(self.model, self._watch)

    def _run(self):
        m = self.model
        vm = m.video_model.video_model
        cur = torch.cuda.current_stream()
        n_clips = self.frame.numel() // (vm.t_former.num_patches * vm.t_former.dim)
        fused = torch.empty((n_clips * 12, 256), dtype=torch.float32, device=self.frame.device)
        L = _lib.lib()
        old_cap = 0
        if self.sm_split is not None:
            # branch: the SFormer (independent of the chain below) on its own share of the SMs
            self.side2.wait_stream(cur)
            old_cap = L.avf_set_sm_cap(self.sm_split[0])
            n_beside = self.stage3.shape[0] if len(self.sm_split) < 3 else min(int(self.sm_split[2]), self.stage3.shape[0])
            s_out = torch.empty_like(self.stage3)
            with torch.cuda.stream(self.side2):
                vm.s_former.sformer(self.stage3[:n_beside], out=s_out[:n_beside])
            L.avf_set_sm_cap(self.sm_split[1])
        # branch: the audio AU_former is independent of everything up to the fusion head
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            m.audio_model.au_head.tokens_into(self.audio, self.audio.shape[1], n_clips, out=fused, ld_out=256)
        cls = vm.t_former.cls_features(self.frame)
        m.video_model.au_head.tokens_into(cls, cls.shape[1], n_clips, out=fused[:, 128:], ld_out=256)
        cur.wait_stream(self.side)
        out21, dec = m.au_head.logits21_(fused, n_clips, True)
        work = None
        if self._peer is not None:
            self._peer.push(out21)
        elif self.gather_into is not None:
            import torch.distributed as dist
            work = dist.all_gather_into_tensor(self.gather_into, out21, async_op=True)      # a graph branch from here ...
        if self.sm_split is not None:
            L.avf_set_sm_cap(old_cap)
            cur.wait_stream(self.side2)
            if n_beside < self.stage3.shape[0]:
                vm.s_former.sformer(self.stage3[n_beside:], out=s_out[n_beside:])
        else:
            s_out = vm.s_former.sformer(self.stage3)
        if self._peer is not None:
            self._peer.wait()
        if work is not None:
            work.wait()                                                                      # ... joined here, behind the SFormer
        return s_out, out21, dec

    def release(self) -> None:
        """Destroy the captured graph.  With gather_into the graph holds NCCL kernel nodes, and the communicator cannot be torn down
        (dist.destroy_process_group() blocks) while such a graph is alive: call this first."""
        torch.cuda.synchronize()
        self.graph.reset()
        self.graph = None
        self.out = None

    def replay(self, stage3: Optional[torch.Tensor] = None, frame: Optional[torch.Tensor] = None, audio: Optional[torch.Tensor] = None):
        if self.graph is None:
            raise RuntimeError("GraphedHotPath.replay() after release()")
        if stage3 is not None and stage3.data_ptr() != self.stage3.data_ptr():
            self.stage3.copy_(stage3, non_blocking=True)
        if frame is not None and frame.data_ptr() != self.frame.data_ptr():
            self.frame.copy_(frame, non_blocking=True)
        if audio is not None and audio.data_ptr() != self.audio.data_ptr():
            self.audio.copy_(audio, non_blocking=True)
        if weights_signature(self.model, self._watch) != self._sig:   # optimizer.step(), load_state_dict(), in-place edits: re-pack and capture again
            if self.model.training:
                raise RuntimeError("GraphedHotPath captures the inference kernels: call model.eval() first")
            self.recaptures += 1
            self._capture()
        self.graph.replay()
        if self._peer is not None:
            self._peer.note_replay()
        return self.out


class GraphedTrainStep:
    """One training step of the hot path (train.py:206-236 on the transformer stack) as a single graph replay:

        zero_grad -> hot_path_train -> AULoss (+ optional gradient of the SFormer output) -> backward      [captured]
        FusedAdam.step()  (gradient all-reduce + one update kernel; bias corrections depend on the step number) [eager]

    Dropout masks are drawn afresh on every replay through a device-side salt word (see avf_encoder_stack_fwd_train); the
    weight re-packing (fp32 master -> bf16 operands) is part of the captured forward, so replays always see the parameters
    the optimiser just wrote."""

    def __init__(self, model, optimizer, stage3, frame, audio, labels, d_sformer_out: Optional[torch.Tensor] = None):
        self.model, self.opt = model, optimizer
        self.stage3 = stage3.detach().clone().requires_grad_(stage3.requires_grad)
        self.frame = frame.detach().clone().requires_grad_(frame.requires_grad)
        self.audio = audio.detach().float().clone().requires_grad_(audio.requires_grad)
        self.labels = labels.detach().clone()
        self.d_s = d_sformer_out.detach().clone() if d_sformer_out is not None else None
        self.salt = torch.zeros(1, dtype=torch.int32, device=labels.device)
        for mod in model.modules():
            if isinstance(mod, Transformer):
                mod.dropout_salt = self.salt
        # eager steps first: they build the optimiser's flat buckets (p.grad become persistent views) and warm every kernel up
        _side_warmup(lambda: (self._fwd_bwd(), self.opt.step()))
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._fwd_bwd()

    def _fwd_bwd(self):
        self.salt.add_(0x61C88647)                      # new dropout masks for this step (int32 wrap-around is fine)
        self.opt.zero_grad()
        for t in (self.stage3, self.frame, self.audio):
            if t.grad is not None:
                t.grad = None
        s_out, out21 = self.model.hot_path_train(self.stage3, self.frame, self.audio)
        loss = self.model.get_au_loss(out21, self.labels)
        if self.d_s is not None and s_out.requires_grad:
            torch.autograd.backward([loss, s_out], [None, self.d_s])
        else:
            loss.backward()
        if hasattr(self.opt, "finish_reductions"):
            self.opt.finish_reductions()                # data parallel: the segment all-reduces started during backward join here (inside the capture)
        return loss.detach()

    def step(self, stage3=None, frame=None, audio=None, labels=None):
        with torch.no_grad():
            for dst, src in ((self.stage3, stage3), (self.frame, frame), (self.audio, audio), (self.labels, labels)):
                if src is not None and src.data_ptr() != dst.data_ptr():
                    dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.opt.step()
        return self.loss
