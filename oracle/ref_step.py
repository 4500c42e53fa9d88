"""The core training step executed by the REFERENCE'S OWN classes (Generator / Discriminator / helper functions of
audiogan.py:1-552, exec'd by oracle/ref_loader.py) on stock torch CPU kernels.  TEST / BASELINE INFRASTRUCTURE ONLY:
only bench.py's `--impl reference` arm and tests/ call this.

The loop body itself (audiogan.py:703-921) is py2 script code interleaved with data loading, TensorBoard writers and .cuda()
calls and cannot be executed; what is restated here is its call ORDER for the core step of SURVEY 8(d) -- even-iteration
D-update (:723-728, :748-751, :761-766, :780-788) + G-update (:816-864, :897-921 without the REINFORCE / feature-matching
extras) -- while every number is produced by the reference's modules (weight-norm pre-hooks on every call, per-frame
LSTMCell loop, dynamic_rnn sort / pack / unpack, its own clip_grad / check_grad) and torch.optim.RMSprop (:693-694).

Where the definitions come from: /root/reference/audiogan.py in the build container; on the GPU box (no /root/reference) the
copy of lines 1-552 that __graft_entry__.build() drops into oracle/_ref/ (git-ignored build artefact, like a compiled
reference binary would be)."""
import os
import time

import torch as T

from . import ref_loader as R
from . import restated as O

_REF_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "audiogan_defs.py")


def locate():
    """-> path of the reference definitions (the original, else the build-time copy), or None."""
    if os.path.exists(R.REFERENCE_FILE):
        return R.REFERENCE_FILE
    if os.path.exists(_REF_COPY):
        return _REF_COPY
    return None


def write_ref_copy():
    """build(): keep lines 1-552 of the reference beside the oracle so that the reference arm can run where /root/reference
    does not exist.  Returns the path, or None when the reference is not present (GPU box: the prebuilt copy is used)."""
    if not os.path.exists(R.REFERENCE_FILE):
        return _REF_COPY if os.path.exists(_REF_COPY) else None
    os.makedirs(os.path.dirname(_REF_COPY), exist_ok=True)
    with open(R.REFERENCE_FILE) as f:
        lines = f.readlines()[:R._N_DEF_LINES]
    with open(_REF_COPY, "w") as f:
        f.writelines(lines)
    return _REF_COPY


def load_namespace():
    path = locate()
    if path is None:
        raise FileNotFoundError("reference definitions not found (neither %s nor %s)" % (R.REFERENCE_FILE, _REF_COPY))
    R.REFERENCE_FILE = path
    return R.load()


class ReferenceStep:
    """Reference Generator / Discriminator (default nets), RMSprop(lr 1e-4) for both, per-tensor clip 1 / 0.1."""

    def __init__(self, g_seed=11, d_seed=12, gk=None, dk=None):
        self.ns = ns = load_namespace()
        self.g = ns["Generator"](embed_size=100, **(gk or {}))
        self.d = ns["Discriminator"](embed_size=100, **(dk or {}))
        self.g.load_state_dict(O.pin_stopper(O.init_generator(g_seed, **(gk or {}))))
        self.d.load_state_dict(O.init_discriminator(d_seed, **(dk or {})))
        self.param_g, self.param_d = list(self.g.parameters()), list(self.d.parameters())
        self.opt_g = T.optim.RMSprop(self.param_g, lr=1e-4)                       # :693-694
        self.opt_d = T.optim.RMSprop(self.param_d, lr=1e-4)

    def step(self, inp, dgradclip=1.0, ggradclip=0.1):
        ns, g, d = self.ns, self.g, self.d
        bce, lm = ns["binary_cross_entropy_with_logits_per_sample"], ns["length_mask"]
        with R.py2_tensor_semantics():
            # ---- D-update, even iteration (audiogan.py:706-788)
            for p in self.param_g:
                p.requires_grad = False
            for p in self.param_d:
                p.requires_grad = True
            real = inp["real"] + inp["noise_real"]                                                    # :724-725
            cls_d, _, _, nf_d = d(real, inp["real_len"], inp["c_real"])
            loss_d = (bce(cls_d, T.full_like(cls_d, 0.9), weight=lm(cls_d.size(), nf_d)) / nf_d.float()).mean()   # :727-740
            with T.no_grad():
                fake, _, _, fake_len = g(z=inp["z"], c=inp["c_g"])                                   # :748
            fake = (fake + inp["noise_fake"][:, :fake.shape[1]]).detach()                            # :750-751
            cls_g, _, _, nf_g = d(fake, fake_len, inp["c_d2"])                                       # :761
            loss_g = (bce(cls_g, T.zeros_like(cls_g), weight=lm(cls_g.size(), nf_g)) / nf_g.float()).mean()       # :762-766, :780
            loss = loss_d + loss_g                                                                   # :783
            self.opt_d.zero_grad()
            loss.backward()
            ns["check_grad"](self.param_d)                                                           # :786
            ns["clip_grad"](self.param_d, dgradclip)                                                 # :787
            self.opt_d.step()                                                                        # :788
            # ---- G-update (audiogan.py:816-921, core: no adversarial z, feature penalty, REINFORCE)
            for p in self.param_g:
                p.requires_grad = True
            for p in self.param_d:
                p.requires_grad = False
            fake, _, _, fake_len = g(z=inp["g_z"], c=inp["g_c_g"])                                   # :841
            fake = fake + inp["g_noise_fake"][:, :fake.shape[1]]                                     # :842-843
            cls_g2, _, _, nf_g2 = d(fake, fake_len, inp["g_c_d"])                                    # :845
            lg = (bce(cls_g2, T.full_like(cls_g2, 0.5), weight=lm(cls_g2.size(), nf_g2)) / nf_g2.float()).mean()  # :857-864, :897
            self.opt_g.zero_grad()
            lg.backward()                                                                            # :902-903
            ns["check_grad"](self.param_g)                                                           # :909
            ns["clip_grad"](self.param_g, ggradclip)                                                 # :910
            self.opt_g.step()                                                                        # :921
        return float(loss_d), float(loss_g), float(lg)


def time_steps(inp, steps, warmup, **kw):
    rs = ReferenceStep(**kw)
    times, losses = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        losses = rs.step(inp)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), losses
