"""The core training step as ONE CUDA graph.

The step (train.d_update + train.g_update on fixed shapes, audiogan.py:706-788 + :816-921 with ``--critic_iter 1
--gencatchup 1``) contains no host synchronisation once the per-layer length tables are cached: every kernel of both
forward passes, both backward passes (autograd runs inside the capture), the weight-norm / packing kernels and the per-tensor
clip + RMSprop launches are recorded once and replayed with one ``cudaGraphLaunch`` per step.  Data-parallel, the capture is
cut at the two gradient all-reduces (after the discriminator's backward, after the generator's): three graph segments that
share one memory pool, with the bucketed NCCL all-reduces issued between them in stream order (no host synchronisation; the
collectives are deliberately NOT captured -- graph-captured NCCL kernels deadlocked against eager collectives on this stack).  That removes the host's per-launch cost (≈170 C-ABI launches + ≈240 small torch
fills / copies per step, 6 ms of one host core: profiles/r1_host_profile.txt) and the launch gaps between small kernels.

Inputs live in static device buffers (``GraphedStep.inputs``); ``run(batch)`` copies a batch into them (device-to-device
or host-to-device, on the current stream) and replays.  Per-sample lengths are host metadata that decides shapes and masks:
they are fixed at capture time (a loop with ragged batches keeps one GraphedStep per length pattern, or runs eagerly).
"""
import torch

from . import _abi as A
from . import engine
from . import train


class GraphedStep:
    def __init__(self, g, d, opt_d, opt_g, example, clip_d=1.0, clip_g=0.1, grad_sync=None, warmup=2,
                 d_kwargs=None, g_kwargs=None):
        """example: a step_inputs()-style dict (tensors on the device or host; ``*_len`` entries stay host tensors)."""
        self.g, self.d, self.opt_d, self.opt_g = g, d, opt_d, opt_g
        self.clip_d, self.clip_g, self.grad_sync = clip_d, clip_g, grad_sync
        self.d_kwargs, self.g_kwargs = dict(d_kwargs or {}), dict(g_kwargs or {})
        dev = next(g.parameters()).device
        self.device = dev
        self.inputs = {}
        for k, v in example.items():
            if not torch.is_tensor(v):
                continue
            self.inputs[k] = v if k.endswith("_len") else v.to(dev, copy=True)
        self.graph = None
        self.out = None
        self.launches = 0
        self._capture(warmup)

    # the step on the static buffers ------------------------------------------------------------------------------
    def _step(self, grad_sync):
        di = dict(self.inputs)
        di["u_stop"] = None
        if not self.d_kwargs and not self.g_kwargs:
            return train.core_step(self.g, self.d, self.opt_d, self.opt_g, di, clip_d=self.clip_d, clip_g=self.clip_g,
                                   grad_sync=grad_sync)
        m1 = train.d_update(self.g, self.d, self.opt_d, di, clip=self.clip_d, grad_sync=grad_sync, **self.d_kwargs)
        gb = {"c_g": di["g_c_g"], "c_d": di["g_c_d"], "z": di["g_z"], "noise_fake": di["g_noise_fake"], "u_stop": None}
        for k in ("real", "real_len", "noise_real", "noise_adv"):
            if k in di:
                gb[k] = di[k]
        m2 = train.g_update(self.g, self.d, self.opt_g, gb, clip=self.clip_g, grad_sync=grad_sync, **self.g_kwargs)
        return m1, m2

    def _capture(self, warmup):
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            # eager warm-up on the capture's side stream: one-time kernel attributes, tensor maps' first use, NCCL
            # communicators, the length-table and constant caches are all populated before the capture starts
            for _ in range(max(1, warmup)):
                self._step(self.grad_sync)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        # NCCL all-reduces are issued between graph segments; the peer-memory all-reduce (dist.PeerGradSync) is plain kernels
        distributed = (self.grad_sync is not None and getattr(self.grad_sync, "world", 1) > 1
                       and not getattr(self.grad_sync, "capturable", False))
        if distributed and any(getattr(m, "_plan", None) is not None and m._plan.early_sync is not None for m in (self.g, self.d)):
            raise RuntimeError("GraphedStep: GradSync.attach (all-reduce inside backward) cannot be combined with graph capture")
        self.graphs, self.syncs = [], []
        state = {}

        def begin():
            gr = torch.cuda.CUDAGraph()
            # thread_local: the autograd engine's worker thread (and NCCL's watchdog) issue CUDA calls while the capture runs
            ctx = torch.cuda.graph(gr, pool=(self.graphs[0].pool() if self.graphs else None), stream=side,
                                   capture_error_mode="thread_local")
            ctx.__enter__()
            self.graphs.append(gr)
            state["ctx"] = ctx

        def cut(params):
            """the all-reduce point between backward and the optimizer step: end this graph segment, reduce eagerly (on
            whatever the gradient buffers hold: nothing has executed yet), start the next segment"""
            engine.shadow_join(self.device)                      # side-stream work must rejoin before a capture segment ends
            state["ctx"].__exit__(None, None, None)
            params = list(params)
            scale = self.grad_sync(params)
            self.syncs.append(params)
            begin()
            return scale

        l0 = A.launches
        begin()
        try:
            m1, m2 = self._step(cut if distributed else self.grad_sync)
            self.out = {"loss_d": m1["loss_d"], "loss_g": m1["loss_g"], "loss": m2["loss"],
                        "losses": torch.stack([m1["loss_d"], m1["loss_g"], m2["loss"]]),
                        "d_grad_norm": m1["d_grad_norm"], "g_grad_norm": m2["g_grad_norm"],
                        "stats_d": m1["stats_d"], "stats_g": m1["stats_g"]}
        finally:
            state["ctx"].__exit__(None, None, None)
        self.launches = A.launches - l0
        self.graph = self.graphs[0]

    # ---------------------------------------------------------------------------------------------------------------
    def load(self, batch, non_blocking=True):
        """copy a batch (host or device tensors of the captured shapes) into the static input buffers, on the current stream"""
        for k, dst in self.inputs.items():
            if k.endswith("_len"):
                src = batch.get(k)
                if src is not None and not torch.equal(src.cpu().to(dst.dtype), dst.cpu()):
                    raise ValueError("GraphedStep: lengths differ from the captured ones (%s); capture a new step" % k)
                continue
            dst.copy_(batch[k], non_blocking=non_blocking)

    def replay(self):
        for i, gr in enumerate(self.graphs):
            gr.replay()
            if i < len(self.syncs):
                self.grad_sync(self.syncs[i])          # bucketed NCCL all-reduce of that net's gradients, in stream order
        return self.out

    def run(self, batch=None):
        if batch is not None:
            self.load(batch)
        return self.replay()
