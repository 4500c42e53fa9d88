"""The host side of the step's inputs and of its checkpoints.

``make_sample`` / ``collate`` mirror what ``dataset.py`` hands the loop for one minibatch (dataset.py:48-91): every waveform is
placed at the start of a zero row of ``maxlen`` samples, peak-normalised to max|x| = 1 (:68-71), its length rounded UP to the
generator frame (:57), words as zero-padded character codes (:43-46).  ``StepFeed`` is the pipeline around the step that the
reference does synchronously on the training thread (numpy -> ``tovar`` -> ``.cuda()`` for every tensor, audiogan.py:94-97,
:719-751): pinned host staging, one copy stream, two device slots -- batch i+1's host-to-device copies run while step i
computes, and nothing is allocated per step.  ``save_checkpoint`` / ``load_checkpoint`` keep the reference's file naming
(audiogan.py:696-701, :936-939) with state_dicts whose keys are the reference's (``rnn.0.module.weight_hh_v`` ...), so a
checkpoint moves between the two code bases in either direction.
"""
import os

import numpy as np
import torch

from .modules import div_roundup


def roundup(x, d):                                               # utiltf.roundup, used by dataset.py:57
    return div_roundup(x, d) * d


def make_sample(wave, maxlen, frame_size=None):
    """dataset.py:48-60, :68-71 for one waveform: (row of `maxlen` float32 samples, length) or (None, None) when the waveform is
    longer than maxlen or silent (the reference re-draws in both cases)."""
    wave = np.asarray(wave, dtype=np.float64)
    nz = np.nonzero(wave)[0]
    n = int(nz[-1]) + 1 if len(nz) else 0                         # :54 -- length without the trailing zeros
    if n > maxlen or n == 0:
        return None, None
    out = np.zeros(maxlen, dtype=np.float64)
    out[:n] = wave[:n]
    out /= np.abs(out).max()                                      # :68-71
    return out.astype(np.float32), (n if frame_size is None else roundup(n, frame_size))   # :57


def word_to_seq(word, maxcharlen):                               # dataset.py:43-46
    seq = np.zeros(maxcharlen, dtype=np.int64)
    seq[:len(word)] = [ord(c) for c in word]
    return seq


def collate(samples, words=None, maxcharlen=None):
    """[(row, length)] (+ words) -> dict of CPU tensors: real (B, maxlen) float32, real_len (B,) int64, [chars (B, maxcharlen)
    int64, char_len (B,) int64] -- what `dataloader.next()` returns (dataset.py:91), as tensors."""
    out = {"real": torch.from_numpy(np.stack([s for s, _ in samples])),
           "real_len": torch.tensor([int(l) for _, l in samples], dtype=torch.int64)}
    if words is not None:
        mc = maxcharlen or max(len(w) for w in words)
        out["chars"] = torch.from_numpy(np.stack([word_to_seq(w, mc) for w in words]))
        out["char_len"] = torch.tensor([len(w) for w in words], dtype=torch.int64)
    return out


class StepFeed:
    """Double-buffered host-to-device pipeline for step inputs.

    ``source``: an iterator of dicts of CPU tensors with fixed shapes (entries whose key ends in ``_len`` are host metadata --
    Discriminator.forward needs their host values -- and are passed through).  ``next()`` returns the batch whose copies were
    issued one call earlier, already ordered behind them on the current stream, and issues the copies of the following one:

        feed = StepFeed(batches, device)
        for _ in range(steps):
            batch = feed.next()          # device tensors (slot k); valid until the call after next
            core_step(g, d, opt_d, opt_g, batch)
    """

    def __init__(self, source, device, slots=2):
        self.source = iter(source)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.nslots = slots
        self.pinned, self.dev, self.ready, self.consumed = [None] * slots, [None] * slots, [None] * slots, [None] * slots
        self.i = 0
        self.h2d_bytes = 0
        self._issue(0)

    def _issue(self, i):
        k = i % self.nslots
        try:
            host = next(self.source)
        except StopIteration:
            self.ready[k] = None
            return
        if self.pinned[k] is None:                                # first use of the slot: the only allocations this class makes
            self.pinned[k] = {n: torch.empty_like(v).pin_memory() for n, v in host.items() if not n.endswith("_len")}
            self.dev[k] = {n: torch.empty(v.shape, dtype=v.dtype, device=self.device) for n, v in self.pinned[k].items()}
            self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.pinned[k].values())
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[k] is not None:
                self.consumed[k].synchronize()                    # host: the pinned staging of this slot is free again
                self.copy_stream.wait_event(self.consumed[k])     # device: the step that read the slot has finished
            for n, pv in self.pinned[k].items():
                pv.copy_(host[n])                                 # pageable -> pinned (a dataset would decode straight into it)
                self.dev[k][n].copy_(pv, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        batch = dict(self.dev[k])
        batch.update({n: v for n, v in host.items() if n.endswith("_len")})
        self.ready[k] = (batch, ev)

    def next(self):
        k = self.i % self.nslots
        if self.ready[k] is None:
            raise StopIteration
        batch, ev = self.ready[k]
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        if self.i >= 1:
            kp = (self.i - 1) % self.nslots                      # the previous batch's consumer work has been enqueued by now
            self.consumed[kp] = torch.cuda.Event()
            self.consumed[kp].record(cur)
        self.i += 1
        self._issue(self.i)
        return batch

    __next__ = next

    def __iter__(self):
        return self


# ------------------------------------------------------------------------------------------------- checkpoints
_KINDS = (("dis", "d"), ("gen", "g"), ("eg", "e_g"), ("ed", "e_d"))


def checkpoint_paths(prefix, iteration):
    """audiogan.py:936-939 / :698-701: '<prefix>-dis-00500', '-gen-', '-eg-', '-ed-'."""
    return {name: "%s-%s-%05d" % (prefix, tag, iteration) for tag, name in _KINDS}


def save_checkpoint(prefix, iteration, **modules):
    """modules: any of d=, g=, e_g=, e_d=.  Writes one file per module under the reference's names holding the module's
    state_dict (CPU tensors, the reference's key names).  The reference pickles whole module objects (`T.save(d, ...)`): its
    class definitions are py2 code that cannot be unpickled here, the state_dict is the part that carries over."""
    paths = checkpoint_paths(prefix, iteration)
    for name, m in modules.items():
        sd = {k: v.detach().to("cpu").clone() for k, v in m.state_dict().items()}
        torch.save(sd, paths[name])
    return {n: paths[n] for n in modules}


def load_checkpoint(prefix, iteration, **modules):
    """Load the files written by save_checkpoint -- or a reference checkpoint's state_dict dumped with
    ``T.save(module.state_dict(), path)`` -- into the given modules (strict key match) and invalidate their packed operands."""
    paths = checkpoint_paths(prefix, iteration)
    for name, m in modules.items():
        obj = torch.load(paths[name], map_location="cpu", weights_only=False)
        sd = obj.state_dict() if hasattr(obj, "state_dict") else obj
        m.load_state_dict(sd)
        if hasattr(m, "invalidate_packed"):
            m.invalidate_packed()
    return {n: paths[n] for n in modules}


def checkpoint_exists(prefix, iteration, names=("d", "g")):
    paths = checkpoint_paths(prefix, iteration)
    return all(os.path.exists(paths[n]) for n in names)
